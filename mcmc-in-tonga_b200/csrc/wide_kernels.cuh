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
#include <type_traits>

#include "proposal.cuh"
#include "tonga_internal.cuh"

namespace tg {

// ---- ray sharding of the streamed sampler ("alternative for config 3": the seam is the ray loop MCsub.jl:142).
// Rank g of `world` owns a contiguous range of the streamed sampler's tiles (whole rays) and keeps the per-point state of
// those points only up to date; proposals, models, phi, t*, counters and history are replicated (every rank draws the same
// Philox proposals and takes the same decisions).  The one exchange step per proposal -- t* and the misfit term of every ray
// under the candidate models -- is fused into the candidate pass: the kernel stores each ray's (t*, term) into the exchange
// block of EVERY rank over peer memory (NVLink P2P stores; plain stores for the own block), tg_shard_signal_kernel then
// raises this rank's sequence flag in every rank's block, and the accept kernel waits until all ranks' flags have reached
// the iteration's sequence number before it sums the terms in the canonical order.  Buffers alternate with the parity of
// the sequence number: a rank can be at most one exchange ahead of the slowest one (it needs that rank's flag to pass its own
// accept kernel), so the buffer being filled is never the one a slower rank still reads.
constexpr int TG_MAX_SHARDS = 16;
struct ShardArgs {
    int rank, world;                  // world <= 1: not sharded
    int tile0, tile1;                 // own tiles of the streamed sampler
    unsigned long long seq;           // sequence number of this iteration's exchange (monotonic over the batch's lifetime)
    unsigned long long timeout_ns;    // bound of the accept kernel's wait (a missing peer must not hang the GPU)
    unsigned long long *flags;        // own exchange block: flags[16 * w] = last sequence number rank w has published
    int *err;                         // own exchange block: set when a wait timed out
    unsigned char *peer_base[TG_MAX_SHARDS];  // exchange blocks of all ranks (peer_base[rank] = own)
    double *peer_tsc[TG_MAX_SHARDS];          // [n][Rp] candidate t* buffer of this parity in rank w's block
    double *peer_term[TG_MAX_SHARDS];         // [n][Rp] candidate misfit terms, same
    // culled point pass (stream_cull.cuh): only the DIRTY rays travel.  Block layout behind the header: for parity 0, 1 and source
    // rank 0..world-1 one section {cnt[n] | ray[n][cap] | t*[n][cap] | term[n][cap]}; a rank writes its own section of every block.
    int mb;                   // 1: mailbox protocol
    int mb_cap, mb_n;         // records per chain and section (>= own rays of any rank); chains
    const int32_t *ndirty;    // this rank's dirty counts (published by the signal kernel)
    unsigned long long mb_hdr, mb_sec, mb_cnt_bytes;
};
__device__ __forceinline__ unsigned char *mb_section(const ShardArgs &sh, int block_of, int src) {
    return sh.peer_base[block_of] + sh.mb_hdr + ((sh.seq & 1ull) * (unsigned long long)sh.world + (unsigned long long)src) * sh.mb_sec;
}
__device__ __forceinline__ int32_t *mb_cnt(unsigned char *sec) { return reinterpret_cast<int32_t *>(sec); }
__device__ __forceinline__ int32_t *mb_ray(const ShardArgs &sh, unsigned char *sec) { return reinterpret_cast<int32_t *>(sec + sh.mb_cnt_bytes); }
__device__ __forceinline__ double *mb_t(const ShardArgs &sh, unsigned char *sec) {
    return reinterpret_cast<double *>(sec + sh.mb_cnt_bytes + 4ull * sh.mb_n * sh.mb_cap);
}
__device__ __forceinline__ double *mb_term(const ShardArgs &sh, unsigned char *sec) {
    return reinterpret_cast<double *>(sec + sh.mb_cnt_bytes + 12ull * sh.mb_n * sh.mb_cap);
}
__device__ __forceinline__ unsigned long long tg_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// every rank's flag in every rank's block: "my stores of exchange `seq` are performed" (stream order puts this kernel after the
// candidate pass; the system-scope fence + release store make the pass's peer stores visible before the flag)
__global__ void tg_shard_signal_kernel(const ShardArgs sh) {
    if (sh.mb) {  // mailbox protocol: how many records this rank has left in every block
        for (int w = 0; w < sh.world; w++) {
            if (w == sh.rank) continue;
            int32_t *c = mb_cnt(mb_section(sh, w, sh.rank));
            for (int i = threadIdx.x; i < sh.mb_n; i += blockDim.x) c[i] = sh.ndirty[i];
        }
        __syncthreads();
    }
    __threadfence_system();
    const int w = threadIdx.x;
    if (w < sh.world) {
        unsigned long long *f = reinterpret_cast<unsigned long long *>(sh.peer_base[w]) + 16 * sh.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(sh.seq) : "memory");
    }
}
// wait (threads 0..world-1 of the CTA, then a CTA barrier) until every rank has published exchange `seq`
__device__ __forceinline__ void shard_wait(const ShardArgs &sh, int tid) {
    if (sh.world <= 1) return;
    if (tid < sh.world) {
        const unsigned long long *f = sh.flags + 16 * tid;
        const unsigned long long t0 = tg_globaltimer();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= sh.seq) break;
            if (tg_globaltimer() - t0 > sh.timeout_ns) { atomicExch(sh.err, 1); break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
}

struct WideArgs {
    tonga_params prm;
    int R, Rp, KC, mode, hist_cap;
    long long iter, it, nIter;
    unsigned long long seed;
    long long chain_id0;
    const int32_t *ray_orig, *ray_rank;
    const double *tS, *sig;  // sorted ray order
    // current state
    int32_t *K;
    double *cells, *phi, *noise, *beta, *tstar;  // cells [n][4][KC]; tstar [n][Rp] sorted ray order
    long long *counts;
    int32_t *pending_slot;
    // candidate
    int32_t *Kc;      // [n]; -1 = nothing to evaluate
    double *cells_c;  // [n][4][KC]
    float *cells_cf;  // [n][3][KC] fl32 copy of the candidate nuclei, +inf beyond Kc (streamed sampler; NULL otherwise)
    double *ptS_c;    // [n][R] caller's ray order (written by tg_eval_kernel)
    Prop *props;      // [n]
    // streamed sampler: t* of the candidates in sorted ray order, accept flags for the commit pass
    int streamed;
    double *tstar_c;        // [n][Rp]
    const double *term_c;   // [n][Rp] misfit terms of the candidate (written by tg_stream_kernel<false>)
    int32_t *active;        // [n] chains whose candidate needs the point pass this iteration (any order)
    int32_t *n_active;      // [1], zeroed by the host before the propose kernel
    int32_t *accept_flag;   // [n]
    ShardArgs sh;           // ray sharding (streamed sampler): the accept kernel waits for every rank's candidate pass
    // streamed sampler with culling (stream_cull.cuh): per-chain candidate / dirty ray lists, misfit terms of the current state
    int culled;
    int32_t *ncand, *ndirty;   // [n], zeroed by the propose kernel
    int32_t *nclist, *ncommit; // [n] rays with something to commit (candidate pass) -> the commit pass's count (accept kernel)
    const int32_t *dirty;      // [n][R]
    double *term;              // [n][Rp]
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
            if (pr.do_eval && (pr.action == 2 || pr.action == 3)) {  // position of the killed / changed nucleus (a move carries its old position already):
                pr.ox = cur[pr.idx]; pr.oy = cur[KC + pr.idx]; pr.oz = cur[2 * KC + pr.idx];  // the culled point pass needs it after the model was committed
            }
            s_prop = pr;
            a.props[chain] = pr;
            if (a.mode == 0 && a.recs_out) {
                tonga_proposal rec;
                rec.action = pr.action; rec.idx = pr.idx; rec.x = pr.x; rec.y = pr.y; rec.z = pr.z; rec.zeta = pr.zeta; rec.u = pr.u;
                a.recs_out[(size_t)chain * a.nIter + a.it] = rec;
            }
            const int ps = a.pending_slot[chain];
            if (ps >= 0) { a.hist_next[(size_t)chain * a.hist_cap + ps] = pr.action; a.pending_slot[chain] = -1; }
            if (a.culled) { a.ncand[chain] = 0; a.ndirty[chain] = 0; a.nclist[chain] = 0; a.ncommit[chain] = 0; }
        }
    }
    __syncthreads();
    const int act = s_prop.action, idx = s_prop.idx;
    const int Kn = (act == 1) ? K + 1 : (act == 2 ? K - 1 : K);
    // sigma move / prior sampling (debug_prior): t* does not change or is not needed -> no forward model
    const bool geometry = s_prop.do_eval && act != 5 && (a.streamed || !a.prm.debug_prior);
    if (tid == 0) {
        a.Kc[chain] = geometry ? Kn : -1;
        if (geometry && a.active) a.active[atomicAdd(a.n_active, 1)] = chain;
    }
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
            if (a.cells_cf && c < 3) a.cells_cf[((size_t)chain * 3 + c) * KC + i] = (i < Kn) ? (float)v : __int_as_float(0x7f800000);
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
    const double *tc = a.ptS_c + (size_t)chain * R;        // wide: candidate t*, caller's ray order
    const double *tsc = a.tstar_c + (size_t)chain * a.Rp;  // streamed: candidate t*, sorted ray order
    int accepted = 0;
    shard_wait(a.sh, tid);  // ray-sharded: (t*, term) of the other ranks' rays have arrived in this rank's exchange block
    if (a.sh.world > 1 && a.sh.mb && do_eval && act != 5) {  // culled point pass: the other ranks' dirty rays -> tstar_c / term_c
        double *tsc_w = a.tstar_c + (size_t)chain * a.Rp, *trm_w = const_cast<double *>(a.term_c) + (size_t)chain * a.Rp;
        for (int w = 0; w < a.sh.world; w++) {
            if (w == a.sh.rank) continue;
            unsigned char *sec = mb_section(a.sh, a.sh.rank, w);
            const int cnt = __ldcg(mb_cnt(sec) + chain);
            const int32_t *rr = mb_ray(a.sh, sec) + (size_t)chain * a.sh.mb_cap;
            const double *tt = mb_t(a.sh, sec) + (size_t)chain * a.sh.mb_cap, *mm = mb_term(a.sh, sec) + (size_t)chain * a.sh.mb_cap;
            for (int d = tid; d < cnt; d += TG_PHI_LANES) {
                const int r = __ldcg(rr + d);
                tsc_w[r] = __ldcg(tt + d); trm_w[r] = __ldcg(mm + d);
            }
        }
        __syncthreads();
    }
    if (do_eval) {
        double phin;
        if (pm.debug_prior) {
            phin = 1.0;  // MCsub.jl:128-136
        } else {
            const double nz = (act == 5) ? pr.zeta : noise;
            if (a.streamed && act != 5) {  // the candidate pass left the per-ray misfit terms: canonical sum only
                const double *trm = a.term_c + (size_t)chain * a.Rp;
                double acc = 0.0;
                int c = 0;
                for (; (c + 16) * TG_PHI_LANES <= R; c += 16) {  // 16 loads in flight, then the ordered adds (canonical phi order, phi_ray)
                    double v[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) v[u] = __ldcg(trm + phi_ray(c + u, tid));  // L2: peers write these rows
#pragma unroll
                    for (int u = 0; u < 16; u++) acc = __dadd_rn(acc, v[u]);
                }
                for (; c * TG_PHI_LANES < R; c++) {
                    const int r = phi_ray(c, tid);
                    if (r < R) acc = __dadd_rn(acc, __ldcg(trm + r));
                }
                acc = warp_sum_canonical(acc);
                if ((tid & 31) == 0) scratch[tid >> 5] = acc;
                __syncthreads();
                phin = phi_warp_sums(scratch);
                __syncthreads();
            } else {
                phin = phi_canonical(R, tid, scratch, [&](int r) {
                    const double t = (act == 5) ? ts[r] : (a.streamed ? tsc[r] : tc[a.ray_orig[r]]);
                    return misfit_term(t, a.tS[r], a.sig[r], nz);
                });
            }
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
                if (!a.streamed && !pm.debug_prior)  // streamed: the commit pass copies t* tile by tile
                    for (int r = tid; r < R; r += TG_PHI_LANES) ts[r] = tc[a.ray_orig[r]];
                if (a.streamed && a.sh.world > 1 && !a.sh.mb)  // ray-sharded: t* is replicated, the commit pass only sees this rank's tiles
                    for (int r = tid; r < R; r += TG_PHI_LANES) ts[r] = __ldcg(tsc + r);
            }
            phi = phin;
            if (act == 1) K += 1;
            else if (act == 2) K -= 1;
            else if (act == 5) noise = pr.zeta;
        }
        if (a.culled) {
            // the candidate pass changed tstar_c / term_c for the DIRTY rays only: commit them to the chain's t* / term, or restore
            // them (invariant between iterations: tstar_c == t*, term_c == term)
            double *tsc_w = a.tstar_c + (size_t)chain * a.Rp, *trm_w = const_cast<double *>(a.term_c) + (size_t)chain * a.Rp;
            double *trm_cur = a.term + (size_t)chain * a.Rp;
            if (act != 5) {
                const int nd = a.ndirty[chain];
                const int32_t *dl = a.dirty + (size_t)chain * R;
                for (int d = tid; d < nd; d += TG_PHI_LANES) {
                    const int r = dl[d];
                    if (accepted) { ts[r] = tsc_w[r]; trm_cur[r] = trm_w[r]; }
                    else { tsc_w[r] = ts[r]; trm_w[r] = trm_cur[r]; }
                }
                if (a.sh.world > 1 && a.sh.mb) {  // ... and the other ranks' dirty rays (t*, term are replicated)
                    for (int w = 0; w < a.sh.world; w++) {
                        if (w == a.sh.rank) continue;
                        unsigned char *sec = mb_section(a.sh, a.sh.rank, w);
                        const int cnt = __ldcg(mb_cnt(sec) + chain);
                        const int32_t *rr = mb_ray(a.sh, sec) + (size_t)chain * a.sh.mb_cap;
                        for (int d = tid; d < cnt; d += TG_PHI_LANES) {
                            const int r = __ldcg(rr + d);
                            if (accepted) { ts[r] = tsc_w[r]; trm_cur[r] = trm_w[r]; }
                            else { tsc_w[r] = ts[r]; trm_w[r] = trm_cur[r]; }
                        }
                    }
                }
            } else if (accepted) {  // the noise level changed: every term does
                for (int r = tid; r < R; r += TG_PHI_LANES) {
                    const double v = misfit_term(ts[r], a.tS[r], a.sig[r], noise);
                    trm_cur[r] = v; trm_w[r] = v;
                }
            }
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
            const double *tcur = (a.streamed && accepted && act != 5) ? tsc : ts;  // streamed: the commit pass has not copied t* yet
            for (int i = tid; i < R; i += TG_PHI_LANES) hp[i] = __ldcg(tcur + a.ray_rank[i]);  // caller's ray order, contiguous stores
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
        if (a.accept_flag) a.accept_flag[chain] = accepted;
        if (a.culled) a.ncommit[chain] = (accepted && do_eval && act != 5 && act != 3) ? a.nclist[chain] : 0;
    }
}

// ================================================================================================ streamed sampler
// Per-point chain state in HBM (u16 owner + fl32 squared distance to the owner per ray point), updated incrementally:
// one pass over the points per proposal instead of the O(P K) forward model.  CTA = (tile of whole rays, chain).
//   COMMIT = false  candidate pass: classify every point of the tile against the proposal (birth / move: fl32 distance
//                   to the new position vs the cached owner distance, exact FP64 inside the error band; death / move:
//                   points of the killed / moved nucleus are rescanned over the candidate model), re-integrate the rays that
//                   hold a changed point (left to right, canonical order) and write t* of every ray of the tile to tstar_c.
//                   Nothing of the chain state is modified.
//   COMMIT = true   accepted chains only: the same classification again (deterministic), now written to owner / dcache;
//                   a death also renumbers the owners above the killed index (deleteat!, TD_inversion_function.jl:132-135).
// Algorithmic traffic of a pass: 18 B per point (3 x fl32 coordinates, fl32 owner distance, u16 owner).
struct StreamArgs {
    const Tile *tiles;
    const float *pxf, *pyf, *pzf;
    const double *px, *py, *pz, *dt;
    const int32_t *ray_off;
    float tol_alpha, tol_beta2;
    int exact_only, KC, Rp, tile_pts;
    long long Ppad;
    const Prop *props;
    const int32_t *Kc;        // candidate nCells (-1: nothing to evaluate)
    const double *cells_c;    // candidate models [n][4][KC]
    const float *cells_cf;    // their fl32 copies [n][3][KC], +inf beyond Kc
    int n_chains;
    uint16_t *owner;          // [n][Ppad]
    float *dcache;            // [n][Ppad]
    double *tstar;            // [n][Rp]
    double *tstar_c;          // [n][Rp] t* under the candidate model
    double *term_c;           // [n][Rp] misfit term of every ray under the candidate model (summed by the accept kernel)
    const double *tS, *sig, *noise;
    const int32_t *accept_flag;
    const int32_t *active, *n_active;  // chains with a candidate to process (tg_wide_propose_kernel)
    uint8_t *tile_changed;             // [n_chains][n_tiles]: the candidate pass found a point of the tile that changes owner (birth / move)
    int n_tiles;
    ShardArgs sh;                      // ray sharding: the launch covers tiles [sh.tile0, sh.tile1); (t*, term) go to every rank's block
};

constexpr int STREAM_THREADS = 256;
constexpr int STREAM_GROUP = 1;  // active chains a CTA processes one after the other on its tile (coordinates re-read from L1)
constexpr int STREAM_PPT = 4;  // points per thread and batch of phase 1 (all loads of a batch are issued before the first use)
#define TG_NONE16S 0xFFFFu

template <bool COMMIT>
__global__ void __launch_bounds__(STREAM_THREADS, 5) tg_stream_kernel(const StreamArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // CTA = (tile, group of STREAM_GROUP active chains), processed one after the other: the tile's coordinates come from HBM once,
    // from L2 for the other groups and from L1 for the other chains of the group.  The launch covers n_tiles x ceil(n_chains /
    // GROUP) CTAs; those beyond the active list exit at once.  (A persistent-CTA variant with a static item loop was measured
    // 40 % slower.)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_active = *a.n_active;
    const int n_groups = (a.n_chains + STREAM_GROUP - 1) / STREAM_GROUP;
    const int grp = blockIdx.x % n_groups;
    if (grp * STREAM_GROUP >= n_active) return;
    const int tidx = a.sh.tile0 + blockIdx.x / n_groups;  // (tile0 = 0 unless ray-sharded)
    const Tile tile = a.tiles[tidx];
    for (int ci = grp * STREAM_GROUP; ci < min(n_active, (grp + 1) * STREAM_GROUP); ci++) {
    __syncthreads();  // the previous chain's shared-memory state is free
    const int chain = a.active[ci];
    const Prop pr = a.props[chain];
    const int act = pr.action;
    if (COMMIT && !a.accept_flag[chain]) continue;
    const int Kn = a.Kc[chain];  // nuclei of the candidate model
    if (COMMIT) {  // the accepted candidate's t* becomes the chain's
        const double *tsc = a.tstar_c + (size_t)chain * a.Rp;
        double *ts = a.tstar + (size_t)chain * a.Rp;
        if (a.sh.world <= 1)  // (ray-sharded: the accept kernel has copied the whole replicated row)
            for (int r = tile.r0 + tid; r < tile.r1; r += STREAM_THREADS) ts[r] = tsc[r];
        if (act == 3) continue;  // change: no owner moves
        if ((act == 1 || act == 4) && !a.tile_changed[(size_t)chain * a.n_tiles + tidx]) continue;  // no point of this tile switches
    }
    // shared layout: owner16[tile_pts] | changed bitmap [tile_pts/32] | queue u16[tile_pts] (orphans, then dirty rays) | counters
    const int cap = a.tile_pts + 8;  // the aligned groups may start up to 3 points before / end up to 3 points after the tile
    uint16_t *s_owner = reinterpret_cast<uint16_t *>(smem_raw);
    uint32_t *s_chg = reinterpret_cast<uint32_t *>(s_owner + cap);
    uint16_t *s_queue = reinterpret_cast<uint16_t *>(s_chg + cap / 32 + 1);
    int *s_cnt = reinterpret_cast<int *>(s_queue + cap);

    const double *cc = a.cells_c + (size_t)chain * 4 * a.KC;
    const float *cf = a.cells_cf + (size_t)chain * 3 * a.KC;  // fl32 copy of the candidate nuclei, +inf beyond Kn
    uint16_t *own = a.owner + (size_t)chain * a.Ppad;
    float *dc = a.dcache + (size_t)chain * a.Ppad;
    const int idx = pr.idx;
    const double cx = pr.x, cy = pr.y, cz = pr.z;
    const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
    const int newo = (act == 1) ? Kn - 1 : idx;  // index a switched point takes: the appended nucleus / the moved one
    const float ta = a.tol_alpha, tb = a.tol_beta2;

    if (!COMMIT) for (int i = tid; i < cap / 32 + 1; i += STREAM_THREADS) s_chg[i] = 0u;
    if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    __syncthreads();

    // ---- phase 1: one flat pass over the tile's points, 4 consecutive points (one aligned 8 / 16 B vector per array) per thread
    // and step.  The first / last group may reach into the neighbouring tiles: those points are masked out (never written).
    // ACT is a compile-time constant so that every action carries only its own logic.
    const long long p0a = tile.p0 & ~3ll;                     // aligned start; s_owner / s_chg are indexed relative to it
    const int lo = (int)(tile.p0 - p0a), hi = (int)(tile.p1 - p0a);  // valid relative range [lo, hi)
    const int ngroups = (hi + 3) >> 2;
    auto phase1 = [&](auto actc) {
        constexpr int ACT = decltype(actc)::value;
#pragma unroll 2
        for (int g = tid; g < ngroups; g += STREAM_THREADS) {
            const long long p = p0a + 4 * g;
            const ushort4 ov = *reinterpret_cast<const ushort4 *>(own + p);
            const uint32_t o[4] = {ov.x, ov.y, ov.z, ov.w};
            uint32_t no[4] = {ov.x, ov.y, ov.z, ov.w};
            uint32_t valid = 0xFu;  // bit q: point 4g+q belongs to this tile
            if (4 * g < lo) valid &= 0xFu << (lo - 4 * g);
            if (4 * g + 4 > hi) valid &= 0xFu >> (4 * g + 4 - hi);
            uint32_t chg = 0;       // bit q: the owner or the zeta of the point changes under the candidate model
            uint32_t wr = 0;        // bit q: COMMIT must write the owner
            float dnew[4];
            if (ACT == 1 || ACT == 4) {
                const float4 dv = *reinterpret_cast<const float4 *>(dc + p);
                const float4 xv = *reinterpret_cast<const float4 *>(a.pxf + p), yv = *reinterpret_cast<const float4 *>(a.pyf + p),
                             zv = *reinterpret_cast<const float4 *>(a.pzf + p);
                const float d_o[4] = {dv.x, dv.y, dv.z, dv.w}, x[4] = {xv.x, xv.y, xv.z, xv.w}, y[4] = {yv.x, yv.y, yv.z, yv.w}, z[4] = {zv.x, zv.y, zv.z, zv.w};
                uint32_t amb = 0, orph = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float d_c = dist2_f32(cxf, cyf, czf, x[q], y[q], z[q]);
                    dnew[q] = d_c;
                    const float diff = d_c - d_o[q], tol = fmaf(ta, d_c + d_o[q], tb);
                    const bool mine = (ACT == 4) && ((int)o[q] == idx);  // owned by the moved nucleus: full rescan below
                    orph |= mine ? (1u << q) : 0u;
                    chg |= (!mine && diff < -tol) ? (1u << q) : 0u;
                    amb |= (!mine && (a.exact_only || !(fabsf(diff) > tol))) ? (1u << q) : 0u;  // inside the error band (or NaN)
                }
                amb &= valid; orph &= valid; chg &= valid;
                if (amb) {  // exact FP64 comparison, MCsub.jl:254-255
#pragma unroll 1
                    for (int q = 0; q < 4; q++) {
                        if (!((amb >> q) & 1u)) continue;
                        const double xe = a.px[p + q], ye = a.py[p + q], ze = a.pz[p + q];
                        const uint32_t oo = o[q];
                        const double de_o = (oo == TG_NONE16S) ? 1e9 : dist2_exact(cc[oo], cc[a.KC + oo], cc[2 * a.KC + oo], xe, ye, ze);
                        const double de_c = dist2_exact(cx, cy, cz, xe, ye, ze);
                        // birth: the new nucleus has the highest index -> strict <.  move: index idx also wins exact ties against o > idx.
                        const bool sw = (de_c < de_o) || (ACT == 4 && de_c == de_o && idx < (int)oo && oo != TG_NONE16S);
                        chg = sw ? (chg | (1u << q)) : (chg & ~(1u << q));
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) no[q] = ((chg >> q) & 1u) ? (uint32_t)newo : o[q];
                wr = chg;
                if (ACT == 4 && orph) {
#pragma unroll 1
                    for (int q = 0; q < 4; q++)
                        if ((orph >> q) & 1u) s_queue[atomicAdd(&s_cnt[0], 1)] = (uint16_t)(4 * g + q);
                }
            } else if (ACT == 2) {
                uint32_t orph = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    orph |= ((int)o[q] == idx) ? (1u << q) : 0u;
                    const bool dec = o[q] != TG_NONE16S && (int)o[q] > idx;  // deleteat! renumbering
                    no[q] = dec ? o[q] - 1 : o[q];
                    wr |= dec ? (1u << q) : 0u;
                }
                orph &= valid; wr &= valid;
                if (orph) {
#pragma unroll 1
                    for (int q = 0; q < 4; q++)
                        if ((orph >> q) & 1u) s_queue[atomicAdd(&s_cnt[0], 1)] = (uint16_t)(4 * g + q);
                }
            } else {  // change: owners stay, the rays through the cell are re-integrated
#pragma unroll
                for (int q = 0; q < 4; q++) chg |= ((int)o[q] == idx) ? (1u << q) : 0u;
                chg &= valid;
            }
            if (!COMMIT) {
                *reinterpret_cast<ushort4 *>(s_owner + 4 * g) = make_ushort4((uint16_t)no[0], (uint16_t)no[1], (uint16_t)no[2], (uint16_t)no[3]);
                if (chg) atomicOr(&s_chg[g >> 3], chg << ((4 * g) & 31));
            } else if (wr) {
                if (valid == 0xFu) {
                    *reinterpret_cast<ushort4 *>(own + p) = make_ushort4((uint16_t)no[0], (uint16_t)no[1], (uint16_t)no[2], (uint16_t)no[3]);
                } else {  // tile boundary: the neighbouring tile's CTA owns the other points of this vector
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if ((wr >> q) & 1u) own[p + q] = (uint16_t)no[q];
                }
                if (ACT == 1 || ACT == 4) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if ((wr >> q) & 1u) dc[p + q] = dnew[q];
                }
            }
        }
    };
    if (act == 1) phase1(std::integral_constant<int, 1>{});
    else if (act == 4) phase1(std::integral_constant<int, 4>{});
    else if (act == 2) phase1(std::integral_constant<int, 2>{});
    else phase1(std::integral_constant<int, 3>{});
    __syncthreads();
    // ---- orphans (death / move): nearest nucleus of the candidate model, fl32 screening over the chain's fl32 nucleus table
    // (L1/L2-resident) + exact recheck
    const int nq = s_cnt[0];
    for (int e = warp; e < nq; e += STREAM_THREADS / 32) {  // one warp per orphan: the lanes split the nuclei
        const int j = s_queue[e];  // relative to p0a
        const long long p = p0a + j;
        int bi = -2;
        float dbest = 1e9f;
        if (!a.exact_only) {
            const float x = a.pxf[p], y = a.pyf[p], z = a.pzf[p];
            float d1 = 1e9f, d2 = 1e9f;
            int i1 = -1;
            for (int i = 4 * lane; i < Kn; i += 128) {
                const float4 fx = __ldg(reinterpret_cast<const float4 *>(cf + i)), fy = __ldg(reinterpret_cast<const float4 *>(cf + a.KC + i)),
                             fz = __ldg(reinterpret_cast<const float4 *>(cf + 2 * a.KC + i));
                const float d[4] = {dist2_f32(fx.x, fy.x, fz.x, x, y, z), dist2_f32(fx.y, fy.y, fz.y, x, y, z), dist2_f32(fx.z, fy.z, fz.z, x, y, z),
                                    dist2_f32(fx.w, fy.w, fz.w, x, y, z)};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool lt = d[u] < d1;
                    d2 = lt ? d1 : fminf(d2, d[u]);
                    i1 = lt ? i + u : i1;
                    d1 = lt ? d[u] : d1;
                }
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {  // merge (best, second best) across the lanes
                const float e1 = __shfl_xor_sync(0xffffffffu, d1, m), e2 = __shfl_xor_sync(0xffffffffu, d2, m);
                const int ei = __shfl_xor_sync(0xffffffffu, i1, m);
                const bool lt = e1 < d1;
                d2 = fminf(fminf(d2, e2), lt ? d1 : e1);
                i1 = lt ? ei : i1;
                d1 = lt ? e1 : d1;
            }
            dbest = d1;
            const float tol = fmaf(ta, d1 + d2, tb);
            if (d2 - d1 > tol) bi = i1;  // unambiguous (an exact fl32 tie between two lanes has d2 == d1 -> exact path)
        }
        if (bi == -2) {  // exact FP64: each lane keeps its first minimum, the merge prefers the lower index on ties (MCsub.jl:255)
            const double x = a.px[p], y = a.py[p], z = a.pz[p];
            double best = 1e9;
            int b = 0x7fffffff;
            for (int i = lane; i < Kn; i += 32) {
                const double d = dist2_exact(cc[i], cc[a.KC + i], cc[2 * a.KC + i], x, y, z);
                if (d < best) { best = d; b = i; }
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, m);
                const int oi = __shfl_xor_sync(0xffffffffu, b, m);
                if (ob < best || (ob == best && oi < b)) { best = ob; b = oi; }
            }
            bi = (b == 0x7fffffff) ? -1 : b;
            dbest = (float)best;
        }
        if (lane == 0) {
            const uint16_t no = bi < 0 ? (uint16_t)TG_NONE16S : (uint16_t)bi;
            if (COMMIT) {
                own[p] = no;
                dc[p] = bi < 0 ? 1e9f : dbest;
            } else {
                s_owner[j] = no;
                if (act == 2 || bi != idx) atomicOr(&s_chg[j >> 5], 1u << (j & 31));  // a point that stays with the moved nucleus keeps its zeta
            }
        }
    }
    if (COMMIT) continue;
    __syncthreads();
    if (act == 1 || act == 4) {  // tell the commit pass whether this tile holds any switching point at all
        int any = 0;
        for (int i = tid; i < cap / 32 + 1; i += STREAM_THREADS) any |= (s_chg[i] != 0u);
        any = __syncthreads_or(any | (act == 4 && s_cnt[0] > 0));  // orphans of a move always count (their owner distance changes)
        if (tid == 0) a.tile_changed[(size_t)chain * a.n_tiles + tidx] = (uint8_t)(any != 0);
    }
    // ---- phase 2: t* of the tile's rays.  Rays without a changed point keep their t*; the others are listed and re-integrated
    // by the warps in the canonical order.
    const double *zc = cc + 3 * (size_t)a.KC;
    auto zeta_of = [&](uint16_t o) -> double { return o == TG_NONE16S ? 0.0 : zc[o]; };
    const double *ts = a.tstar + (size_t)chain * a.Rp;
    double *tsc = a.tstar_c + (size_t)chain * a.Rp;
    double *trm = a.term_c + (size_t)chain * a.Rp;
    const double nz = a.noise[chain];
    // (t*, term) of a ray: own buffers, and -- ray-sharded -- the same row of every other rank's exchange block (P2P stores)
    auto publish = [&](int r, double t, double term) {
        tsc[r] = t; trm[r] = term;
        for (int w = 0; w < a.sh.world; w++)
            if (w != a.sh.rank) { a.sh.peer_tsc[w][(size_t)chain * a.Rp + r] = t; a.sh.peer_term[w][(size_t)chain * a.Rp + r] = term; }
    };
    for (int r = tile.r0 + tid; r < tile.r1; r += STREAM_THREADS) {
        const int q0 = a.ray_off[r], n = a.ray_off[r + 1] - q0;
        const int j0 = (int)(q0 - p0a), j1 = j0 + n;  // bit range [j0, j1) of the changed bitmap
        bool dirty = false;
        for (int w = j0 >> 5; w <= (j1 - 1) >> 5 && n > 0; w++) {
            uint32_t bits = s_chg[w];
            if (w == (j0 >> 5)) bits &= 0xFFFFFFFFu << (j0 & 31);
            if (w == ((j1 - 1) >> 5) && (j1 & 31)) bits &= 0xFFFFFFFFu >> (32 - (j1 & 31));
            dirty |= bits != 0u;
        }
        if (dirty) s_queue[atomicAdd(&s_cnt[1], 1)] = (uint16_t)(r - tile.r0);
        else { const double t = ts[r]; publish(r, t, misfit_term(t, a.tS[r], a.sig[r], nz)); }
    }
    __syncthreads();
    const int nd = s_cnt[1];
    for (int eb = 4 * warp; eb < nd; eb += 4 * (STREAM_THREADS / 32)) {  // canonical t* (tstar_g8): 8 lanes per ray, 4 rays per warp
        const int e = eb + (lane >> 3);
        const bool on = e < nd;
        const int r = tile.r0 + (on ? s_queue[e] : 0);
        const int q0 = on ? a.ray_off[r] : 0, n = on ? a.ray_off[r + 1] - q0 : 0;
        const int nseg = n > 1 ? n - 1 : 0;
        const int trip = __reduce_max_sync(0xffffffffu, (nseg + 7) >> 3);
        const uint16_t *ow = s_owner + (q0 - p0a);
        const double t = tstar_g8(nseg, trip, lane & 7, [&](int j) { return seg_term(a.dt[q0 + j], zeta_of(ow[j]), zeta_of(ow[j + 1])); });
        if (on && (lane & 7) == 0) publish(r, t, misfit_term(t, a.tS[r], a.sig[r], nz));
    }
    }
}

}  // namespace tg
