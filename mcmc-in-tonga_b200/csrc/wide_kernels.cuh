// wide_kernels.cuh -- the "wide" sampler: the proposal loop of TD_inversion_function.jl:70-302 for ray sets / model sizes
// that do not fit the shared-memory-resident sampler (sampler_kernel.cuh: Ppad bytes of owners per chain in shared
// memory, max_cells <= 126).  BASELINE.json config 3 (100k rays, 2e7 points, up to 2000 nuclei) is the target.
//
// Every iteration is three launches on the context's stream, all chains in lock step:
//   tg_wide_propose_kernel   draw / replay the proposal, a-priori checks (proposal.cuh), build the CANDIDATE model in global memory
//   tg_eval_kernel           the full forward model of the candidates (evaluate.cu; candidates with K < 0 are skipped)
//   tg_wide_accept_kernel    canonical phi, acceptance rule, commit (candidate -> current), counters, traces, thinning
// i.e. exactly what the reference does (a full `evaluate` per proposal, MCsub.jl:123-185), batched over chains.  The
// proposal and acceptance code is shared with the resident sampler and t* / phi use the same canonical summation orders, so
// the two samplers produce bit-identical chains (tests/test_gpu_parity.py::test_wide_matches_resident).
#pragma once
#include "proposal.cuh"
#include "tonga_internal.cuh"

namespace tg {

struct WideArgs {
    tonga_params prm;
    int R, Rp, KC, mode, hist_cap;
    long long iter, it, nIter;
    unsigned long long seed;
    long long chain_id0;
    const int32_t *ray_orig;
    const double *tS, *sig;  // sorted ray order
    // current state
    int32_t *K;
    double *cells, *phi, *noise, *beta, *tstar;  // cells [n][4][KC]; tstar [n][Rp] sorted ray order
    long long *counts;
    int32_t *pending_slot;
    // candidate
    int32_t *Kc;      // [n]; -1 = nothing to evaluate
    double *cells_c;  // [n][4][KC]
    double *ptS_c;    // [n][R] caller's ray order (written by tg_eval_kernel)
    Prop *props;      // [n]
    // streams / traces
    const tonga_proposal *recs_in;
    tonga_proposal *recs_out;
    int8_t *tr_accept;
    double *tr_phi;
    int32_t *tr_K;
    // history
    int32_t *n_hist;
    long long *model_num;
    int32_t *hist_K;
    double *hist_cells, *hist_phi, *hist_ptS;
    long long *hist_iter;
    int32_t *hist_action, *hist_accept, *hist_next;
};

constexpr int WIDE_PROPOSE_THREADS = 256;

__global__ void __launch_bounds__(WIDE_PROPOSE_THREADS) tg_wide_propose_kernel(const WideArgs a) {
    __shared__ Prop s_prop;
    const int chain = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int KC = a.KC;
    const int K = a.K[chain];
    const double *cur = a.cells + (size_t)chain * 4 * KC;
    if (tid < 32) {
        Prop pr;
        draw_proposal<1>(pr, a.mode, a.mode == 1 ? a.recs_in + (size_t)chain * a.nIter + a.it : nullptr, a.seed, a.iter,
                      (unsigned long long)(a.chain_id0 + chain), cur, cur + KC, cur + 2 * KC, cur + 3 * KC, K, a.noise[chain], a.prm,
                         a.prm.zeta_scale * a.prm.sig / 100, lane);
        if (lane == 0) {
            s_prop = pr;
            a.props[chain] = pr;
            if (a.mode == 0 && a.recs_out) {
                tonga_proposal rec;
                rec.action = pr.action; rec.idx = pr.idx; rec.x = pr.x; rec.y = pr.y; rec.z = pr.z; rec.zeta = pr.zeta; rec.u = pr.u;
                a.recs_out[(size_t)chain * a.nIter + a.it] = rec;
            }
            const int ps = a.pending_slot[chain];
            if (ps >= 0) { a.hist_next[(size_t)chain * a.hist_cap + ps] = pr.action; a.pending_slot[chain] = -1; }
        }
    }
    __syncthreads();
    const int act = s_prop.action, idx = s_prop.idx;
    const int Kn = (act == 1) ? K + 1 : (act == 2 ? K - 1 : K);
    // sigma move / prior sampling (debug_prior): t* does not change or is not needed -> no forward model
    const bool geometry = s_prop.do_eval && act != 5 && !a.prm.debug_prior;
    if (tid == 0) a.Kc[chain] = geometry ? Kn : -1;
    if (!s_prop.do_eval || act == 5) return;
    double *cand = a.cells_c + (size_t)chain * 4 * KC;
    const double nv[4] = {s_prop.x, s_prop.y, s_prop.z, s_prop.zeta};
    for (int i = tid; i < KC; i += WIDE_PROPOSE_THREADS) {
        const int src = (act == 2 && i >= idx) ? i + 1 : i;  // deleteat!, TD_inversion_function.jl:132-135
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double v = (i < Kn && src < KC) ? cur[c * KC + src] : 0.0;
            if (act == 1 && i == K) v = nv[c];                  // append!, :85-88
            else if (act == 3 && i == idx && c == 3) v = nv[3];  // :189
            else if (act == 4 && i == idx && c < 3) v = nv[c];   // :233-235
            cand[c * KC + i] = v;
        }
    }
}

__global__ void __launch_bounds__(TG_PHI_LANES) tg_wide_accept_kernel(const WideArgs a) {
    __shared__ double scratch[4];
    __shared__ int s_accept;
    const int chain = blockIdx.x, tid = threadIdx.x;
    const int KC = a.KC, R = a.R;
    const tonga_params &pm = a.prm;
    const Prop pr = a.props[chain];
    const int act = pr.action, do_eval = pr.do_eval;
    int K = a.K[chain];
    double phi = a.phi[chain], noise = a.noise[chain];
    const double beta = a.beta[chain];
    double *ts = a.tstar + (size_t)chain * a.Rp;
    const double *tc = a.ptS_c + (size_t)chain * R;
    int accepted = 0;
    if (do_eval) {
        double phin;
        if (pm.debug_prior) {
            phin = 1.0;  // MCsub.jl:128-136
        } else {
            const double nz = (act == 5) ? pr.zeta : noise;
            phin = phi_canonical_128(R, tid, scratch, [&](int r) {
                const double t = (act == 5) ? ts[r] : tc[a.ray_orig[r]];
                return misfit_term(t, a.tS[r], a.sig[r], nz);
            });
        }
        if (tid == 0) s_accept = accept_decision(pr, K, phi, phin, (act == 2 || act == 3) ? a.cells[(size_t)chain * 4 * KC + 3 * KC + pr.idx] : 0.0, noise, beta, R, pm,
                                                     pm.zeta_scale * pm.sig / 100);
        __syncthreads();
        accepted = s_accept;
        if (accepted) {
            if (act != 5) {
                double *cur = a.cells + (size_t)chain * 4 * KC;
                const double *cand = a.cells_c + (size_t)chain * 4 * KC;
                for (int i = tid; i < 4 * KC; i += TG_PHI_LANES) cur[i] = cand[i];
                if (!pm.debug_prior)
                    for (int r = tid; r < R; r += TG_PHI_LANES) ts[r] = tc[a.ray_orig[r]];
            }
            phi = phin;
            if (act == 1) K += 1;
            else if (act == 2) K -= 1;
            else if (act == 5) noise = pr.zeta;
        }
    }
    // bookkeeping, traces, thinning (:275-281) -- as step G of the resident sampler
    long long model_num = a.model_num[chain];
    int n_hist = a.n_hist[chain];
    int keep = 0;
    if ((double)a.iter >= pm.burn_in) {
        model_num += 1;
        if (fmod((double)model_num, pm.keep_each) == 0) keep = 1;
    }
    __syncthreads();  // the committed state is visible to the whole CTA before the history copy
    if (keep && beta == 1.0) {
        if (n_hist < a.hist_cap) {
            const size_t h = (size_t)chain * a.hist_cap + n_hist;
            double *hc = a.hist_cells + h * 4 * KC;
            const double *cur = a.cells + (size_t)chain * 4 * KC;
            for (int i = tid; i < 4 * KC; i += TG_PHI_LANES) hc[i] = cur[i];
            double *hp = a.hist_ptS + h * R;
            for (int r = tid; r < R; r += TG_PHI_LANES) hp[a.ray_orig[r]] = ts[r];
            if (tid == 0) {
                a.hist_K[h] = K; a.hist_phi[h] = phi; a.hist_iter[h] = a.iter;
                a.hist_action[h] = act; a.hist_accept[h] = accepted; a.hist_next[h] = 0;
                a.pending_slot[chain] = n_hist;
            }
        }
        n_hist += 1;
    }
    if (tid == 0) {
        if (act >= 1 && act <= 5) {
            long long *c = a.counts + (size_t)chain * 15 + (act - 1);
            c[0] += 1;
            if (accepted) c[5] += 1;
            if (do_eval) c[10] += 1;
        }
        if (a.tr_accept) a.tr_accept[(size_t)chain * a.nIter + a.it] = (int8_t)accepted;
        if (a.tr_phi) a.tr_phi[(size_t)chain * a.nIter + a.it] = phi;
        if (a.tr_K) a.tr_K[(size_t)chain * a.nIter + a.it] = K;
        a.K[chain] = K; a.phi[chain] = phi; a.noise[chain] = noise;
        a.n_hist[chain] = n_hist; a.model_num[chain] = model_num;
    }
}

}  // namespace tg
