// sampler.cu -- the device-resident RJ-MCMC proposal loop of TD_inversion_function.jl:70-302 for a batch of
// independent chains, with INCREMENTAL Voronoi maintenance instead of a full evaluate() per proposal.
//
// One CTA (128 threads) owns one chain for the whole launch.  The chain's state lives in shared memory:
//   owner8[Ppad]   nearest-nucleus index of every ray point (u8; 0x7F = none; bit 7 = pending tag)
//   tstar[R]       predicted t* per ray (model.ptS)            tnew[R]  proposal's t* for touched rays
//   nuclei SoA     x, y, z, zeta [KC]                          zlut[256] owner byte -> zeta under the proposal
//   mask32 / dirty pending-overwrite bits per point / touched bits per ray
// It is loaded once per launch with TMA bulk copies (cp.async.bulk + mbarrier), nIter iterations run without any
// global synchronisation, and it is written back once.  The ray geometry (shared by all chains) is streamed from L2
// with 128-bit loads.  Per iteration:
//   A  warp 0 draws (Philox4x32-10) or replays the proposal, applies the a-priori checks of the reference, evaluates
//      v_nearest at the new / killed nucleus (warp argmin, lowest index wins ties) and builds zlut;
//   B  all threads: the point pass of the action -- birth: d(p,new) < d(p,owner)?; death: orphans of the killed
//      nucleus rescan the survivors; move: both; change: only marks rays.  Distances are exact FP64 (no FMA);
//   C  one warp per touched ray re-integrates t* in the canonical order;
//   D  canonical phi over all rays;  E  thread 0: alpha exactly as TD_inversion_function.jl:96-97,151-152,196,241;
//   F  commit (accept) or roll back (reject) the pending owner changes;  G  traces / thinning / history.
// The same canonical t* / phi reductions are used by evaluate.cu, so incremental state == full evaluate bit for bit
// (tonga_chains_verify checks that on the device).
#include <cmath>
#include <cstring>

#include "tonga_internal.cuh"

namespace tg {

constexpr int ST = 128;  // threads per chain CTA (4 warps) == TG_PHI_LANES
static_assert(ST == TG_PHI_LANES, "the canonical phi reduction is defined over 128 lanes");

struct SamplerArgs {
    // geometry
    const double *px, *py, *pz, *dt, *tS, *sig;
    const int32_t *rayid, *ray_off;
    int R, Rp, KC;
    int P, Ppad;
    tonga_params prm;
    // chain state (global)
    int32_t *K;
    double *cells;  // [n][4][KC]
    double *phi, *noise, *beta;
    uint8_t *owner;  // [n][Ppad]
    double *tstar;   // [n][Rp]
    long long *counts;  // [n][3][5] proposed / accepted / evaluated
    int32_t *pending_slot;  // [n] history slot awaiting its next_action, or -1
    // run
    long long iter0, nIter;
    int mode;  // 0 generate, 1 replay
    const tonga_proposal *recs_in;
    tonga_proposal *recs_out;
    int8_t *tr_accept;
    double *tr_phi;
    int32_t *tr_K;
    unsigned long long seed;
    long long chain_id0;
    // history
    int hist_cap;
    int32_t *n_hist;
    long long *model_num;
    int32_t *hist_K;
    double *hist_cells, *hist_phi, *hist_ptS;
    long long *hist_iter;
    int32_t *hist_action, *hist_accept, *hist_next;
};

struct Prop {  // proposal of the current iteration, broadcast through shared memory
    int action, idx, do_eval, valid, accept, K, keep, pad;
    double x, y, z, zeta, u;
    double aux;    // birth: czeta (:81); death: zetanew (:146)
    double phi, phin, noise, beta;
};

struct SmemLayout {
    size_t o_owner, o_mask, o_tstar, o_tnew, o_dirty, o_nuc, o_zlut, o_scr, o_prop, o_bar, total;
};
__host__ __device__ inline SmemLayout smem_layout(int Ppad, int Rp, int KC) {
    SmemLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
    L.o_owner = take((size_t)Ppad);
    L.o_mask = take((size_t)Ppad / 8);
    L.o_tstar = take(8 * (size_t)Rp);
    L.o_tnew = take(8 * (size_t)Rp);
    L.o_dirty = take(4 * (size_t)((Rp + 31) / 32));
    L.o_nuc = take(8 * 4 * (size_t)KC);
    L.o_zlut = take(8 * 256);
    L.o_scr = take(8 * 8);
    L.o_prop = take(sizeof(Prop));
    L.o_bar = take(16);
    L.total = o;
    return L;
}

// ---- TMA bulk copies (SASS: UBLKCP) -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s2u(dst)),
                 "l"(src), "r"(bytes), "r"(s2u(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s2u(src)), "r"(bytes) : "memory");
}

// ---- warp-cooperative v_nearest (MCsub.jl:247-263) over the nuclei in shared memory, skipping index `skip` ---------
__device__ __forceinline__ int warp_nearest(const double *nx, const double *ny, const double *nz, int K, int skip,
                                            double x, double y, double z, int lane) {
    double best = 1e9;
    int bi = 0x7fffffff;
    for (int i = lane; i < K; i += 32) {
        if (i == skip) continue;
        const double d = dist2_exact(nx[i], ny[i], nz[i], x, y, z);
        if (d < best) { best = d; bi = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
    return bi == 0x7fffffff ? -1 : bi;
}

__device__ __forceinline__ double jl_min1(double a) {  // min([1 a]...) in Julia: NaN propagates
    return (a != a) ? a : (a < 1.0 ? a : 1.0);
}

__device__ __forceinline__ void mark_dirty(uint32_t *dirty, const int32_t *__restrict__ rayid, int p) {
    const int r = rayid[p];
    atomicOr(&dirty[r >> 5], 1u << (r & 31));
}

__global__ void __launch_bounds__(ST, 7) tg_sampler_kernel(const SamplerArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout L = smem_layout(a.Ppad, a.Rp, a.KC);
    uint8_t *s_owner = smem + L.o_owner;
    uint32_t *s_own32 = reinterpret_cast<uint32_t *>(s_owner);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + L.o_mask);
    double *s_tstar = reinterpret_cast<double *>(smem + L.o_tstar);
    double *s_tnew = reinterpret_cast<double *>(smem + L.o_tnew);
    uint32_t *s_dirty = reinterpret_cast<uint32_t *>(smem + L.o_dirty);
    double *s_nx = reinterpret_cast<double *>(smem + L.o_nuc);
    double *s_ny = s_nx + a.KC, *s_nz = s_ny + a.KC, *s_zeta = s_nz + a.KC;
    double *s_zlut = reinterpret_cast<double *>(smem + L.o_zlut);
    double *s_scr = reinterpret_cast<double *>(smem + L.o_scr);
    Prop *s_prop = reinterpret_cast<Prop *>(smem + L.o_prop);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.o_bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chain = blockIdx.x;
    const int KC = a.KC, R = a.R;
    const int nOwnWords = a.Ppad / 4, nMaskWords = a.Ppad / 32, nDirtyWords = (a.Rp + 31) / 32;

    // ---- load the chain state: three TMA bulk copies on one mbarrier
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t b_owner = (uint32_t)a.Ppad, b_ts = (uint32_t)(8 * a.Rp), b_nuc = (uint32_t)(32 * KC);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(s_bar)), "r"(b_owner + b_ts + b_nuc) : "memory");
        bulk_load(s_owner, a.owner + (size_t)chain * a.Ppad, b_owner, s_bar);
        bulk_load(s_tstar, a.tstar + (size_t)chain * a.Rp, b_ts, s_bar);
        bulk_load(s_nx, a.cells + (size_t)chain * 4 * KC, b_nuc, s_bar);
    }
    for (int i = tid; i < nMaskWords; i += ST) s_mask[i] = 0u;
    for (int i = tid; i < nDirtyWords; i += ST) s_dirty[i] = 0u;
    {
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
                s2u(s_bar))
            : "memory");
    }
    __syncthreads();

    int K = a.K[chain];
    double phi = a.phi[chain];
    double noise = a.noise[chain];
    const double beta = a.beta[chain];
    int n_hist = a.n_hist[chain];
    long long model_num = a.model_num[chain];
    int pending_slot = a.pending_slot[chain];
    long long cnt_prop[5] = {0, 0, 0, 0, 0}, cnt_acc[5] = {0, 0, 0, 0, 0}, cnt_eval[5] = {0, 0, 0, 0, 0};  // thread 0 only

    const tonga_params &pm = a.prm;
    // TD_inversion_function.jl:22-23,30-32
    const double sig_zeta = pm.zeta_scale * pm.sig / 100;
    const double sig_sig = pm.max_sig * pm.sig / 100;
    const double xr = (pm.sig / 100) * (pm.xmax - pm.xmin);
    const double yr = (pm.sig / 100) * (pm.ymax - pm.ymin);
    const double zr = (pm.sig / 100) * (pm.zmax - pm.zmin);
    const int nact = pm.n_actions >= 4 ? pm.n_actions : 4;
    const double PI = 3.141592653589793;
    const unsigned long long gid = (unsigned long long)(a.chain_id0 + chain);
    const Philox philox{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};

    for (long long it = 0; it < a.nIter; it++) {
        const long long iter = a.iter0 + it;
        // ================================================================ A: proposal (warp 0, warp-uniform values)
        if (warp == 0) {
            Prop pr;
            pr.do_eval = 0; pr.valid = 0; pr.accept = 0; pr.K = K; pr.keep = 0; pr.pad = 0;
            pr.idx = 0; pr.x = pr.y = pr.z = pr.zeta = pr.u = pr.aux = 0.0;
            pr.phi = phi; pr.phin = phi; pr.noise = noise; pr.beta = beta;
            double uu[8];
            if (a.mode == 0) {
                uint32_t w[4] = {0, 0, 0, 0};
                if (lane < 4) philox((uint32_t)iter, (uint32_t)((unsigned long long)iter >> 32), (uint32_t)gid, (uint32_t)lane | ((uint32_t)(gid >> 32) << 8), w);
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const uint32_t w0 = __shfl_sync(0xffffffffu, w[0], s), w1 = __shfl_sync(0xffffffffu, w[1], s);
                    const uint32_t w2 = __shfl_sync(0xffffffffu, w[2], s), w3 = __shfl_sync(0xffffffffu, w[3], s);
                    uu[2 * s] = u53(w0, w1);
                    uu[2 * s + 1] = u53(w2, w3);
                }
                int act = 1 + (int)floor(uu[0] * nact);  // rand(1:4), TD_inversion_function.jl:72
                pr.action = act > nact ? nact : act;
            } else {
                const tonga_proposal rec = a.recs_in[(size_t)chain * a.nIter + it];
                pr.action = rec.action; pr.idx = rec.idx; pr.x = rec.x; pr.y = rec.y; pr.z = rec.z; pr.zeta = rec.zeta; pr.u = rec.u;
            }
            // Box-Muller normals from (open) uniforms; only evaluated in generate mode
            auto normal_pair = [&](double u1, double u2, double &n0, double &n1) {
                const double rr = sqrt(-2.0 * log(u1 + 0x1.0p-54));
                double sn, cs;
                sincospi(2.0 * u2, &sn, &cs);
                n0 = rr * cs;
                n1 = rr * sn;
            };
            const int act = pr.action;
            if (act == 1) {  // ---- birth :76-125
                if (K < pm.max_cells) {
                    if (a.mode == 0) {
                        pr.x = uu[2] * (pm.xmax - pm.xmin) + pm.xmin;  // :78
                        pr.y = uu[3] * (pm.ymax - pm.ymin) + pm.ymin;  // :79
                        pr.z = uu[4] * (pm.zmax - pm.zmin) + pm.zmin;  // :80
                    }
                    const int ci = warp_nearest(s_nx, s_ny, s_nz, K, -1, pr.x, pr.y, pr.z, lane);  // :81
                    const double czeta = ci < 0 ? 0.0 : s_zeta[ci];
                    pr.aux = czeta;
                    if (a.mode == 0) {
                        double n0, n1;
                        normal_pair(uu[5], uu[6], n0, n1);
                        pr.zeta = czeta + sig_zeta * n0;  // :82
                        pr.u = uu[7];                     // :121
                    }
                    if (pm.prior == 1) pr.valid = (pr.zeta > 0 && pr.zeta < pm.zeta_scale);  // :92
                    else if (pm.prior == 2) pr.valid = 1;
                    else pr.valid = (pr.zeta > 0);  // :111
                    pr.do_eval = pr.valid;
                }
            } else if (act == 2) {  // ---- death :126-181
                if (K > pm.min_cells) {
                    if (a.mode == 0) {
                        int k = (int)floor(uu[1] * K);  // :128
                        pr.idx = k >= K ? K - 1 : k;
                        pr.u = uu[7];  // :176
                    }
                    const int kill = pr.idx;
                    if (kill >= 0 && kill < K) {
                        const int zi = warp_nearest(s_nx, s_ny, s_nz, K, kill, s_nx[kill], s_ny[kill], s_nz[kill], lane);  // :146
                        pr.aux = zi < 0 ? 0.0 : s_zeta[zi];
                        pr.valid = (pm.prior == 3) ? (pr.aux > 0) : 1;  // :165
                        pr.do_eval = pr.valid;
                    }
                }
            } else if (act == 3) {  // ---- change :183-218
                if (a.mode == 0) {
                    int k = (int)floor(uu[1] * K);  // :184
                    pr.idx = k >= K ? K - 1 : k;
                    double n0, n1;
                    normal_pair(uu[2], uu[3], n0, n1);
                    pr.zeta = s_zeta[pr.idx] + sig_zeta * n0;  // :188
                    pr.u = uu[7];                              // :214
                }
                if (pr.idx >= 0 && pr.idx < K) {
                    if (pm.prior == 1) pr.valid = (pr.zeta > 0 && pr.zeta < pm.zeta_scale);  // :195
                    else if (pm.prior == 2) pr.valid = 1;
                    else pr.valid = (pr.zeta > 0);  // :206 (alpha = 0 otherwise)
                    pr.do_eval = pr.valid;  // the reference evaluates first (:191) but discards the result when invalid
                }
            } else if (act == 4) {  // ---- move :220-251
                if (K > 0) {
                    if (a.mode == 0) {
                        int k = (int)floor(uu[1] * K);  // :222
                        pr.idx = k >= K ? K - 1 : k;
                        double n0, n1, n2, n3;
                        normal_pair(uu[2], uu[3], n0, n1);
                        normal_pair(uu[4], uu[5], n2, n3);
                        pr.x = s_nx[pr.idx] + xr * n0;  // :226
                        pr.y = s_ny[pr.idx] + yr * n1;  // :227
                        pr.z = s_nz[pr.idx] + zr * n2;  // :228
                        pr.u = uu[7];                   // :247
                    }
                    if (pr.idx >= 0 && pr.idx < K)
                        pr.valid = (pr.x >= pm.xmin && pr.x <= pm.xmax && pr.y >= pm.ymin && pr.y <= pm.ymax && pr.z >= pm.zmin &&
                                    pr.z <= pm.zmax);  // :230-232
                    pr.do_eval = pr.valid;
                }
            } else if (act == 5) {  // ---- sigma :252-272 (dead code in the reference; extension)
                if (a.mode == 0) {
                    double n0, n1;
                    normal_pair(uu[2], uu[3], n0, n1);
                    pr.zeta = noise + sig_sig * n0;  // :254
                    pr.u = uu[7];
                }
                pr.valid = (pr.zeta > 0 && pr.zeta < pm.max_sig);  // :257
                pr.do_eval = pr.valid;
            }
            // zlut: owner byte -> zeta under the proposed model
            if (pr.do_eval && act != 5) {
                const double ztag = (act == 1) ? pr.zeta : (act == 4 ? s_zeta[pr.idx] : 0.0);
                for (int o = lane; o < 128; o += 32) {
                    double zv = (o < K) ? s_zeta[o] : 0.0;
                    if (act == 3 && o == pr.idx) zv = pr.zeta;
                    s_zlut[o] = zv;
                    s_zlut[128 + o] = ztag;
                }
            }
            if (lane == 0) {
                *s_prop = pr;
                if (a.mode == 0 && a.recs_out) {
                    tonga_proposal rec;
                    rec.action = pr.action; rec.idx = pr.idx; rec.x = pr.x; rec.y = pr.y; rec.z = pr.z; rec.zeta = pr.zeta; rec.u = pr.u;
                    a.recs_out[(size_t)chain * a.nIter + it] = rec;
                }
                if (pending_slot >= 0) a.hist_next[(size_t)chain * a.hist_cap + pending_slot] = pr.action;
            }
        }
        __syncthreads();
        pending_slot = -1;
        const int act = s_prop->action;
        const int do_eval = s_prop->do_eval;
        const int pidx = s_prop->idx;
        double phin = phi;
        int accepted = 0;

        if (do_eval) {
            if (act != 5) {
                const double cx = s_prop->x, cy = s_prop->y, cz = s_prop->z;
                // ======================================================== B: point pass
                if (act == 1 || act == 4) {
                    const int mv = (act == 4) ? pidx : -1;
                    for (int w = tid; w < nOwnWords; w += ST) {
                        uint32_t ow = s_own32[w];
                        const double2 xa = *reinterpret_cast<const double2 *>(a.px + 4 * w), xb = *reinterpret_cast<const double2 *>(a.px + 4 * w + 2);
                        const double2 ya = *reinterpret_cast<const double2 *>(a.py + 4 * w), yb = *reinterpret_cast<const double2 *>(a.py + 4 * w + 2);
                        const double2 za = *reinterpret_cast<const double2 *>(a.pz + 4 * w), zb = *reinterpret_cast<const double2 *>(a.pz + 4 * w + 2);
                        const double X[4] = {xa.x, xa.y, xb.x, xb.y}, Y[4] = {ya.x, ya.y, yb.x, yb.y}, Z[4] = {za.x, za.y, zb.x, zb.y};
                        uint32_t tags = 0, mbits = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int o = (ow >> (8 * q)) & 0xFF;
                            if (o == mv) {
                                // move, type A: the point belongs to the moved nucleus -> rescan all nuclei (moved one at its new place)
                                double best = 1e9;
                                int bi = TG_OWNER_NONE;
                                for (int i = 0; i < K; i++) {
                                    const double d = (i == mv) ? dist2_exact(cx, cy, cz, X[q], Y[q], Z[q])
                                                               : dist2_exact(s_nx[i], s_ny[i], s_nz[i], X[q], Y[q], Z[q]);
                                    if (d < best) { best = d; bi = i; }
                                }
                                if (bi != mv) {
                                    ow = (ow & ~(0xFFu << (8 * q))) | ((uint32_t)bi << (8 * q));
                                    mbits |= 1u << q;
                                }
                            } else {
                                const double d_o = (o == TG_OWNER_NONE) ? 1e9 : dist2_exact(s_nx[o], s_ny[o], s_nz[o], X[q], Y[q], Z[q]);
                                const double d_c = dist2_exact(cx, cy, cz, X[q], Y[q], Z[q]);
                                // birth: the new nucleus has the highest index -> strict <.  move: index mv also wins exact ties against o > mv.
                                const bool sw = (d_c < d_o) || (act == 4 && d_c == d_o && mv < o && o != TG_OWNER_NONE);
                                if (sw) tags |= 0x80u << (8 * q);
                            }
                        }
                        if (tags | mbits) {
                            s_own32[w] = ow | tags;
                            if (mbits) atomicOr(&s_mask[w >> 3], mbits << ((w & 7) * 4));
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if ((tags >> (8 * q + 7) & 1u) | (mbits >> q & 1u)) mark_dirty(s_dirty, a.rayid, 4 * w + q);
                        }
                    }
                } else if (act == 2) {
                    const int kill = pidx;
                    const uint32_t kk = (uint32_t)kill * 0x01010101u;
                    for (int w = tid; w < nOwnWords; w += ST) {
                        uint32_t ow = s_own32[w];
                        const uint32_t eq = __vcmpeq4(ow, kk);
                        if (eq) {
                            uint32_t mbits = 0;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                if ((eq >> (8 * q)) & 1u) {
                                    const int p = 4 * w + q;
                                    const double x = a.px[p], y = a.py[p], z = a.pz[p];
                                    double best = 1e9;
                                    int bi = TG_OWNER_NONE;
                                    for (int i = 0; i < K; i++) {
                                        if (i == kill) continue;
                                        const double d = dist2_exact(s_nx[i], s_ny[i], s_nz[i], x, y, z);
                                        if (d < best) { best = d; bi = i; }
                                    }
                                    ow = (ow & ~(0xFFu << (8 * q))) | ((uint32_t)bi << (8 * q));  // old numbering; renumbered on accept
                                    mbits |= 1u << q;
                                    mark_dirty(s_dirty, a.rayid, p);
                                }
                            }
                            s_own32[w] = ow;
                            atomicOr(&s_mask[w >> 3], mbits << ((w & 7) * 4));
                        }
                    }
                } else {  // act == 3: owners unchanged; rays through the changed cell are touched
                    const uint32_t kk = (uint32_t)pidx * 0x01010101u;
                    for (int w = tid; w < nOwnWords; w += ST) {
                        const uint32_t eq = __vcmpeq4(s_own32[w], kk);
                        if (eq) {
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if ((eq >> (8 * q)) & 1u) mark_dirty(s_dirty, a.rayid, 4 * w + q);
                        }
                    }
                }
                __syncthreads();
                // ======================================================== C: re-integrate touched rays (canonical order)
                auto zeta_of = [&](uint8_t o) -> double { return s_zlut[o]; };
                for (int r = warp; r < R; r += ST / 32) {
                    if ((s_dirty[r >> 5] >> (r & 31)) & 1u) {
                        const int q0 = a.ray_off[r], n = a.ray_off[r + 1] - q0;
                        const double t = ray_tstar_canonical<uint8_t>(s_owner, a.dt, q0, n, lane, zeta_of);
                        if (lane == 0) s_tnew[r] = t;
                    }
                }
                __syncthreads();
            }
            // ============================================================ D: phi of the proposed model (canonical order)
            const double nz = (act == 5) ? s_prop->zeta : noise;
            phin = phi_canonical_128(R, tid, s_scr, [&](int r) {
                const double t = ((s_dirty[r >> 5] >> (r & 31)) & 1u) ? s_tnew[r] : s_tstar[r];
                return misfit_term(t, a.tS[r], a.sig[r], nz);
            });
            // ============================================================ E: acceptance (thread 0)
            if (tid == 0) {
                const double K0 = (double)K;
                const double zn = s_prop->zeta, aux = s_prop->aux, u = s_prop->u;
                const double dphi2 = beta * ((phin - phi) / 2);
                double alpha = 0.0;
                int acc = 0;
                if (act == 1) {
                    const double g = ((aux - zn) * (aux - zn)) / (2 * (sig_zeta * sig_zeta));
                    if (pm.prior == 1)  // :96-97
                        alpha = ((K0) / (K0 + 1)) * ((sig_zeta * sqrt(2 * PI)) / (pm.zeta_scale)) * exp(g - dphi2);
                    else if (pm.prior == 2)  // :107-108
                        alpha = ((K0) / (K0 + 1)) * (sig_zeta / pm.zeta_scale) * exp(-(zn * zn) / (pm.zeta_scale * pm.zeta_scale) + g - dphi2);
                    else  // :113-114
                        alpha = ((K0) / (K0 + 1)) * (sqrt(2 * PI) * sig_zeta / pm.zeta_scale) * exp(-zn / pm.zeta_scale + g - dphi2);
                    alpha = jl_min1(alpha);
                    acc = (u < alpha);
                } else if (act == 2) {
                    const double zk = s_zeta[pidx];
                    const double g = ((zk - aux) * (zk - aux)) / (2 * (sig_zeta * sig_zeta));
                    if (pm.prior == 1)  // :151-152
                        alpha = ((K0) / (K0 - 1)) * ((pm.zeta_scale) / (sig_zeta * sqrt(2 * PI))) * exp(-g - dphi2);
                    else if (pm.prior == 2)  // :160-162
                        alpha = ((K0) / (K0 - 1)) * (pm.zeta_scale / sig_zeta) * exp((zk * zk) / (2 * (pm.zeta_scale * pm.zeta_scale)) - g - dphi2);
                    else  // :166-168
                        alpha = ((K0) / (K0 - 1)) * (pm.zeta_scale / (sqrt(2 * PI) * sig_zeta)) * exp(zk / pm.zeta_scale - g - dphi2);
                    alpha = jl_min1(alpha);
                    acc = (u < alpha);
                } else if (act == 3) {
                    const double zo = s_zeta[pidx];
                    if (pm.prior == 1) alpha = exp(-dphi2);  // :196
                    else if (pm.prior == 2) alpha = exp((zo * zo - zn * zn) / (2 * (pm.zeta_scale * pm.zeta_scale)) - dphi2);  // :202-203
                    else alpha = exp((zo - zn) / pm.zeta_scale - dphi2);  // :207-208
                    alpha = jl_min1(alpha);
                    acc = (u < alpha);
                } else if (act == 4) {
                    alpha = jl_min1(exp(-dphi2));  // :241-242
                    acc = (u < alpha);
                } else {  // sigma: log form :264-267
                    double la = log(noise / zn) * (double)R - dphi2;
                    la = (la != la) ? la : (la < 0.0 ? la : 0.0);
                    acc = (log(u) <= la);
                }
                s_prop->accept = acc;
                s_prop->phin = phin;
            }
            __syncthreads();
            accepted = s_prop->accept;
            // ============================================================ F: commit / roll back
            if (act == 1) {
                const uint32_t newb = (uint32_t)K * 0x01010101u;
                for (int w = tid; w < nOwnWords; w += ST) {
                    const uint32_t ow = s_own32[w], t = ow & 0x80808080u;
                    if (t) {
                        const uint32_t m = (t >> 7) * 0xFFu;
                        s_own32[w] = accepted ? ((ow & ~m) | (newb & m)) : (ow & 0x7F7F7F7Fu);
                    }
                }
                if (accepted && tid == 0) {  // append!, :85-88
                    s_nx[K] = s_prop->x; s_ny[K] = s_prop->y; s_nz[K] = s_prop->z; s_zeta[K] = s_prop->zeta;
                }
            } else if (act == 2) {
                const int kill = pidx;
                if (accepted) {
                    for (int i = tid; i < nMaskWords; i += ST) s_mask[i] = 0u;
                    const uint32_t kk = (uint32_t)kill * 0x01010101u;
                    for (int w = tid; w < nOwnWords; w += ST) {  // deleteat! renumbering: indices above `kill` shift down (:132-135)
                        const uint32_t ow = s_own32[w];
                        const uint32_t gt = __vcmpgtu4(ow, kk) & ~__vcmpeq4(ow, 0x7F7F7F7Fu);
                        if (gt) s_own32[w] = ow - (gt & 0x01010101u);
                    }
                    if (warp == 0) {  // order-preserving delete of the nucleus
                        double vx[4], vy[4], vz[4], vt[4];
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const int i = kill + lane + 32 * s;
                            if (i < K - 1) { vx[s] = s_nx[i + 1]; vy[s] = s_ny[i + 1]; vz[s] = s_nz[i + 1]; vt[s] = s_zeta[i + 1]; }
                        }
                        __syncwarp();
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const int i = kill + lane + 32 * s;
                            if (i < K - 1) { s_nx[i] = vx[s]; s_ny[i] = vy[s]; s_nz[i] = vz[s]; s_zeta[i] = vt[s]; }
                        }
                    }
                } else {
                    for (int i = tid; i < nMaskWords; i += ST) {
                        uint32_t b = s_mask[i];
                        if (b) {
                            s_mask[i] = 0u;
                            while (b) {
                                const int j = __ffs(b) - 1;
                                b &= b - 1;
                                s_owner[32 * i + j] = (uint8_t)kill;
                            }
                        }
                    }
                }
            } else if (act == 3) {
                if (accepted && tid == 0) s_zeta[pidx] = s_prop->zeta;
            } else if (act == 4) {
                const int mv = pidx;
                // masked bytes first (they never carry a tag), then tags
                if (!accepted) {
                    for (int i = tid; i < nMaskWords; i += ST) {
                        uint32_t b = s_mask[i];
                        if (b) {
                            s_mask[i] = 0u;
                            while (b) {
                                const int j = __ffs(b) - 1;
                                b &= b - 1;
                                s_owner[32 * i + j] = (uint8_t)mv;
                            }
                        }
                    }
                } else {
                    for (int i = tid; i < nMaskWords; i += ST) s_mask[i] = 0u;
                }
                __syncthreads();  // byte stores above and word updates below touch the same words
                const uint32_t newb = (uint32_t)mv * 0x01010101u;
                for (int w = tid; w < nOwnWords; w += ST) {
                    const uint32_t ow = s_own32[w], t = ow & 0x80808080u;
                    if (t) {
                        const uint32_t m = (t >> 7) * 0xFFu;
                        s_own32[w] = accepted ? ((ow & ~m) | (newb & m)) : (ow & 0x7F7F7F7Fu);
                    }
                }
                if (accepted && tid == 0) { s_nx[mv] = s_prop->x; s_ny[mv] = s_prop->y; s_nz[mv] = s_prop->z; }
            }
            if (act != 5) {
                if (accepted)
                    for (int r = tid; r < R; r += ST)
                        if ((s_dirty[r >> 5] >> (r & 31)) & 1u) s_tstar[r] = s_tnew[r];
                __syncthreads();
                for (int i = tid; i < nDirtyWords; i += ST) s_dirty[i] = 0u;
            }
            if (accepted) {
                phi = phin;
                if (act == 1) K += 1;
                else if (act == 2) K -= 1;
                else if (act == 5) noise = s_prop->zeta;
            }
        }
        // ================================================================ G: bookkeeping, traces, thinning (:275-281)
        int keep = 0;
        if ((double)iter >= pm.burn_in) {
            model_num += 1;
            if (fmod((double)model_num, pm.keep_each) == 0) keep = 1;
        }
        if (tid == 0) {
            if (act >= 1 && act <= 5) { cnt_prop[act - 1]++; cnt_acc[act - 1] += accepted; cnt_eval[act - 1] += do_eval; }
            if (a.tr_accept) a.tr_accept[(size_t)chain * a.nIter + it] = (int8_t)accepted;
            if (a.tr_phi) a.tr_phi[(size_t)chain * a.nIter + it] = phi;
            if (a.tr_K) a.tr_K[(size_t)chain * a.nIter + it] = K;
        }
        __syncthreads();  // state (nuclei, tstar, owners) consistent before the next proposal / the history copy
        if (keep) {
            if (n_hist < a.hist_cap) {
                const size_t h = (size_t)chain * a.hist_cap + n_hist;
                double *hc = a.hist_cells + h * 4 * KC;
                for (int i = tid; i < 4 * KC; i += ST) hc[i] = s_nx[i];
                double *hp = a.hist_ptS + h * R;
                for (int r = tid; r < R; r += ST) hp[r] = s_tstar[r];
                if (tid == 0) {
                    a.hist_K[h] = K; a.hist_phi[h] = phi; a.hist_iter[h] = iter;
                    a.hist_action[h] = act; a.hist_accept[h] = accepted; a.hist_next[h] = 0;
                }
                pending_slot = n_hist;
            }
            n_hist += 1;
        }
    }

    // ---- write the chain state back (TMA bulk stores for the arrays)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
        bulk_store(a.owner + (size_t)chain * a.Ppad, s_owner, (uint32_t)a.Ppad);
        bulk_store(a.tstar + (size_t)chain * a.Rp, s_tstar, (uint32_t)(8 * a.Rp));
        bulk_store(a.cells + (size_t)chain * 4 * KC, s_nx, (uint32_t)(32 * KC));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        a.K[chain] = K; a.phi[chain] = phi; a.noise[chain] = noise;
        a.n_hist[chain] = n_hist; a.model_num[chain] = model_num; a.pending_slot[chain] = pending_slot;
        for (int i = 0; i < 5; i++) {
            a.counts[(size_t)chain * 15 + i] += cnt_prop[i];
            a.counts[(size_t)chain * 15 + 5 + i] += cnt_acc[i];
            a.counts[(size_t)chain * 15 + 10 + i] += cnt_eval[i];
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---- build_starting (MCsub.jl:76-121) on the device: one thread per chain draws nCells, then K x 4 uniforms
__global__ void tg_build_starting_kernel(int n, int KC, tonga_params pm, unsigned long long seed, long long chain_id0, int32_t *K,
                                         double *cells, double *noise) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n) return;
    const unsigned long long gid = (unsigned long long)(chain_id0 + chain);
    const Philox philox{(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[4];
    unsigned ctr = 0;
    auto next2 = [&](double &u0, double &u1, bool open) {  // purpose 1 in counter word 1's top bit keeps this stream apart
        philox(ctr++, 0x80000000u, (uint32_t)gid, (uint32_t)(gid >> 32) << 8, w);
        u0 = open ? u53_open(w[0], w[1]) : u53(w[0], w[1]);
        u1 = open ? u53_open(w[2], w[3]) : u53(w[2], w[3]);
    };
    double u0, u1;
    next2(u0, u1, false);
    // :86-87 nCells = floor(exp(rand*log(max_cells/min_cells) + log(min_cells)))
    int k = (int)floor(exp(u0 * log((double)pm.max_cells / (double)pm.min_cells) + log((double)pm.min_cells)));
    if (k > KC) k = KC;
    if (k < 1) k = 1;
    K[chain] = k;
    noise[chain] = 1.0;
    double *c = cells + (size_t)chain * 4 * KC;
    for (int i = 0; i < 4 * KC; i++) c[i] = 0.0;
    for (int i = 0; i < k; i++) {
        double a0, a1, b0, b1;
        next2(a0, a1, false);
        next2(b0, b1, true);
        c[i] = pm.xmin + (pm.xmax - pm.xmin) * a0;           // :92
        c[KC + i] = pm.ymin + (pm.ymax - pm.ymin) * a1;      // :93
        next2(a0, a1, false);
        c[2 * KC + i] = pm.zmin + (pm.zmax - pm.zmin) * a0;  // :94
        double zt;
        if (pm.prior == 1) zt = a1 * pm.zeta_scale;  // :100
        else if (pm.prior == 2) {                    // :105
            double sn, cs;
            sincospi(2.0 * b1, &sn, &cs);
            zt = 0.0 + pm.zeta_scale * (sqrt(-2.0 * log(b0)) * cs);
        } else zt = -log(b0) * pm.zeta_scale;  // :108
        c[3 * KC + i] = zt;
    }
}

__global__ void tg_verify_kernel(int n, int64_t P, int64_t Ppad, int R, int Rp, const uint8_t *own_a, const uint8_t *own_b,
                                 const double *ts_a /* [n][Rp] */, const double *ts_b /* [n][R] */, const double *phi_a, const double *phi_b,
                                 unsigned long long *mism, double *maxd /* [2] as ordered uint64 bits */) {
    const int chain = blockIdx.y;
    unsigned long long local = 0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x)
        local += own_a[(size_t)chain * Ppad + p] != own_b[(size_t)chain * Ppad + p];
    if (local) atomicAdd(mism, local);
    if (blockIdx.x == 0) {
        double dm = 0.0;
        for (int r = threadIdx.x; r < R; r += blockDim.x) {
            const double d = fabs(ts_a[(size_t)chain * Rp + r] - ts_b[(size_t)chain * R + r]);
            dm = (d > dm || d != d) ? d : dm;
        }
        if (dm != dm) dm = INFINITY;
        atomicMax((unsigned long long *)&maxd[1], (unsigned long long)__double_as_longlong(dm));
        if (threadIdx.x == 0) {
            double dp = fabs(phi_a[chain] - phi_b[chain]);
            if (dp != dp) dp = INFINITY;
            atomicMax((unsigned long long *)&maxd[0], (unsigned long long)__double_as_longlong(dp));
        }
    }
}

__global__ void tg_copy_tstar_kernel(int n, int R, int Rp, const double *src /* [n][R] */, double *dst /* [n][Rp] */) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < (size_t)n * Rp) {
        const size_t c = i / Rp, r = i % Rp;
        dst[i] = r < (size_t)R ? src[c * R + r] : 0.0;
    }
}

}  // namespace tg

// ================================================================================================== host side
struct tonga_chains {
    tonga_ctx *ctx = nullptr;
    int n = 0, KC = 0, Rp = 0, hist_cap = 0;
    long long chain_id0 = 0;
    unsigned long long seed = 0;
    long long iter_done = 0;
    bool have_models = false;
    size_t smem = 0;
    // device state
    int32_t *d_K = nullptr;
    double *d_cells = nullptr, *d_phi = nullptr, *d_noise = nullptr, *d_beta = nullptr, *d_tstar = nullptr;
    uint8_t *d_owner = nullptr;
    long long *d_counts = nullptr;
    int32_t *d_pending = nullptr;
    int32_t *d_n_hist = nullptr;
    long long *d_model_num = nullptr;
    int32_t *d_hist_K = nullptr;
    double *d_hist_cells = nullptr, *d_hist_phi = nullptr, *d_hist_ptS = nullptr;
    long long *d_hist_iter = nullptr;
    int32_t *d_hist_action = nullptr, *d_hist_accept = nullptr, *d_hist_next = nullptr;
    // scratch
    double *d_ptS_tmp = nullptr;  // [n][R]
    double *d_phi_tmp = nullptr;
    uint8_t *d_owner_tmp = nullptr;
    unsigned long long *d_mism = nullptr;
    double *d_maxd = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
};

static inline size_t tg_nz(size_t b) { return b ? b : 1; }
#define TG_ALLOC(ptr, bytes) TG_CUDA(cudaMalloc((void **)&(ptr), tg_nz(bytes)))

extern "C" int tonga_chains_create(tonga_ctx *ctx, tonga_chains **out, int32_t nChains, int64_t chain_id0, uint64_t seed,
                                   int32_t hist_cap) {
    if (!ctx || !out || nChains < 1 || hist_cap < 0) return tg::fail(TONGA_ERR_ARG, "tonga_chains_create: bad argument");
    *out = nullptr;
    const tonga_params &pm = ctx->prm;
    if (pm.max_cells > TG_MAX_K_U8 || pm.max_cells < 1 || pm.min_cells < 1 || pm.min_cells > pm.max_cells)
        return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: need 1 <= min_cells <= max_cells <= 126 (u8 owner state)");
    if (ctx->R < 1 || ctx->P < 1) return tg::fail(TONGA_ERR_ARG, "tonga_chains_create: empty ray set");
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    tonga_chains *ch = new tonga_chains();
    ch->ctx = ctx;
    ch->n = nChains;
    ch->KC = ((pm.max_cells + 7) / 8) * 8;
    ch->Rp = (ctx->R + 1) & ~1;
    ch->hist_cap = hist_cap;
    ch->chain_id0 = chain_id0;
    ch->seed = seed;
    ch->smem = tg::smem_layout((int)ctx->Ppad, ch->Rp, ch->KC).total;
    if (ch->smem > ctx->smem_optin) {
        const size_t need = ch->smem;
        delete ch;
        return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: per-chain state (" + std::to_string(need) +
                                                " B) exceeds shared memory; the smem-resident sampler handles ray sets up to ~200k points");
    }
    const size_t n = (size_t)nChains, KC = (size_t)ch->KC, R = (size_t)ctx->R, Rp = (size_t)ch->Rp, Pp = (size_t)ctx->Ppad, H = (size_t)hist_cap;
    TG_ALLOC(ch->d_K, 4 * n);
    TG_ALLOC(ch->d_cells, 8 * n * 4 * KC);
    TG_ALLOC(ch->d_phi, 8 * n);
    TG_ALLOC(ch->d_noise, 8 * n);
    TG_ALLOC(ch->d_beta, 8 * n);
    TG_ALLOC(ch->d_tstar, 8 * n * Rp);
    TG_ALLOC(ch->d_owner, n * Pp);
    TG_ALLOC(ch->d_counts, 8 * n * 15);
    TG_ALLOC(ch->d_pending, 4 * n);
    TG_ALLOC(ch->d_n_hist, 4 * n);
    TG_ALLOC(ch->d_model_num, 8 * n);
    TG_ALLOC(ch->d_hist_K, 4 * n * H);
    TG_ALLOC(ch->d_hist_cells, 8 * n * H * 4 * KC);
    TG_ALLOC(ch->d_hist_phi, 8 * n * H);
    TG_ALLOC(ch->d_hist_ptS, 8 * n * H * R);
    TG_ALLOC(ch->d_hist_iter, 8 * n * H);
    TG_ALLOC(ch->d_hist_action, 4 * n * H);
    TG_ALLOC(ch->d_hist_accept, 4 * n * H);
    TG_ALLOC(ch->d_hist_next, 4 * n * H);
    TG_ALLOC(ch->d_ptS_tmp, 8 * n * R);
    TG_ALLOC(ch->d_phi_tmp, 8 * n);
    TG_ALLOC(ch->d_owner_tmp, n * Pp);
    TG_ALLOC(ch->d_mism, 8);
    TG_ALLOC(ch->d_maxd, 16);
    cudaStream_t s = ctx->stream;
    TG_CUDA(cudaMemsetAsync(ch->d_counts, 0, 8 * n * 15, s));
    TG_CUDA(cudaMemsetAsync(ch->d_pending, 0xFF, 4 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_n_hist, 0, 4 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_model_num, 0, 8 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_cells, 0, 8 * n * 4 * KC, s));
    {
        std::vector<double> ones(n, 1.0);
        TG_CUDA(cudaMemcpyAsync(ch->d_beta, ones.data(), 8 * n, cudaMemcpyHostToDevice, s));
        TG_CUDA(cudaMemcpyAsync(ch->d_noise, ones.data(), 8 * n, cudaMemcpyHostToDevice, s));
        TG_CUDA(cudaStreamSynchronize(s));
    }
    TG_CUDA(cudaFuncSetAttribute(tg::tg_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->smem));
    TG_CUDA(cudaEventCreate(&ch->ev0));
    TG_CUDA(cudaEventCreate(&ch->ev1));
    *out = ch;
    return TONGA_OK;
}

extern "C" void tonga_chains_destroy(tonga_chains *ch) {
    if (!ch) return;
    cudaSetDevice(ch->ctx->device);
    cudaStreamSynchronize(ch->ctx->stream);
    void *ptrs[] = {ch->d_K, ch->d_cells, ch->d_phi, ch->d_noise, ch->d_beta, ch->d_tstar, ch->d_owner, ch->d_counts, ch->d_pending,
                    ch->d_n_hist, ch->d_model_num, ch->d_hist_K, ch->d_hist_cells, ch->d_hist_phi, ch->d_hist_ptS, ch->d_hist_iter,
                    ch->d_hist_action, ch->d_hist_accept, ch->d_hist_next, ch->d_ptS_tmp, ch->d_phi_tmp, ch->d_owner_tmp, ch->d_mism,
                    ch->d_maxd};
    for (void *p : ptrs) cudaFree(p);
    if (ch->ev0) cudaEventDestroy(ch->ev0);
    if (ch->ev1) cudaEventDestroy(ch->ev1);
    delete ch;
}

// full evaluate of the current device-resident models -> owners, t*, phi of the chain state
static int establish_state(tonga_chains *ch) {
    tonga_ctx *ctx = ch->ctx;
    int rc = tg::launch_evaluate(ctx, ch->n, ch->KC, ch->d_K, ch->d_cells, ch->d_noise, ch->d_ptS_tmp, ch->d_phi, nullptr, ch->d_owner);
    if (rc != TONGA_OK) return rc;
    const size_t tot = (size_t)ch->n * ch->Rp;
    tg::tg_copy_tstar_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ch->n, ctx->R, ch->Rp, ch->d_ptS_tmp, ch->d_tstar);
    TG_CUDA(cudaGetLastError());
    ch->have_models = true;
    return TONGA_OK;
}

extern "C" int tonga_chains_build_starting(tonga_chains *ch) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_build_starting: NULL");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    tg::tg_build_starting_kernel<<<(ch->n + 127) / 128, 128, 0, ctx->stream>>>(ch->n, ch->KC, ctx->prm, ch->seed, ch->chain_id0, ch->d_K,
                                                                             ch->d_cells, ch->d_noise);
    TG_CUDA(cudaGetLastError());
    int rc = establish_state(ch);
    if (rc != TONGA_OK) return rc;
    TG_CUDA(cudaStreamSynchronize(ctx->stream));
    return TONGA_OK;
}

extern "C" int tonga_chains_set_models(tonga_chains *ch, int32_t Kcap, const int32_t *K, const double *cells, const double *noise) {
    if (!ch || Kcap < 1 || !K || !cells) return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_models: bad argument");
    tonga_ctx *ctx = ch->ctx;
    const size_t n = (size_t)ch->n, KC = (size_t)ch->KC;
    for (size_t i = 0; i < n; i++)
        if (K[i] < 1 || K[i] > Kcap || K[i] > (int)KC)
            return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_set_models: K[i] outside [1, min(Kcap, " + std::to_string(KC) + ")]");
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> packed(n * 4 * KC, 0.0);
    for (size_t i = 0; i < n; i++)
        for (int a = 0; a < 4; a++) std::memcpy(&packed[(i * 4 + a) * KC], cells + (i * 4 + a) * (size_t)Kcap, 8 * (size_t)K[i]);
    cudaStream_t s = ctx->stream;
    TG_CUDA(cudaMemcpyAsync(ch->d_K, K, 4 * n, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(ch->d_cells, packed.data(), 8 * n * 4 * KC, cudaMemcpyHostToDevice, s));
    std::vector<double> nz(n, 1.0);
    if (noise) std::memcpy(nz.data(), noise, 8 * n);
    TG_CUDA(cudaMemcpyAsync(ch->d_noise, nz.data(), 8 * n, cudaMemcpyHostToDevice, s));
    int rc = establish_state(ch);
    if (rc != TONGA_OK) return rc;
    TG_CUDA(cudaStreamSynchronize(s));
    return TONGA_OK;
}

extern "C" int tonga_chains_set_beta(tonga_chains *ch, const double *beta) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_beta: NULL");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    std::vector<double> b((size_t)ch->n, 1.0);
    if (beta) std::memcpy(b.data(), beta, 8 * (size_t)ch->n);
    TG_CUDA(cudaMemcpy(ch->d_beta, b.data(), 8 * (size_t)ch->n, cudaMemcpyHostToDevice));
    return TONGA_OK;
}

extern "C" int tonga_chains_run(tonga_chains *ch, int64_t nIter, int32_t mode, tonga_proposal *recs, int8_t *tr_accept,
                                double *tr_phi, int32_t *tr_K) {
    if (!ch || nIter < 0 || (mode != 0 && mode != 1) || (mode == 1 && !recs)) return tg::fail(TONGA_ERR_ARG, "tonga_chains_run: bad argument");
    if (!ch->have_models) return tg::fail(TONGA_ERR_STATE, "tonga_chains_run: no start models (call build_starting or set_models)");
    if (nIter == 0) return TONGA_OK;
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t n = (size_t)ch->n, N = n * (size_t)nIter;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_rec = 0, o_acc = o_rec + (recs ? al(sizeof(tonga_proposal) * N) : 0), o_phi = o_acc + (tr_accept ? al(N) : 0),
                 o_K = o_phi + (tr_phi ? al(8 * N) : 0), total = o_K + (tr_K ? al(4 * N) : 0);
    int rc = tg::ensure_scratch(ctx, total);
    if (rc != TONGA_OK) return rc;
    char *d = (char *)ctx->d_scratch;
    if (mode == 1) TG_CUDA(cudaMemcpyAsync(d + o_rec, recs, sizeof(tonga_proposal) * N, cudaMemcpyHostToDevice, s));

    tg::SamplerArgs a{};
    a.px = ctx->d_px; a.py = ctx->d_py; a.pz = ctx->d_pz; a.dt = ctx->d_dt; a.tS = ctx->d_tS; a.sig = ctx->d_sig;
    a.rayid = ctx->d_rayid; a.ray_off = ctx->d_ray_off;
    a.R = ctx->R; a.Rp = ch->Rp; a.KC = ch->KC; a.P = (int)ctx->P; a.Ppad = (int)ctx->Ppad;
    a.prm = ctx->prm;
    a.K = ch->d_K; a.cells = ch->d_cells; a.phi = ch->d_phi; a.noise = ch->d_noise; a.beta = ch->d_beta;
    a.owner = ch->d_owner; a.tstar = ch->d_tstar; a.counts = ch->d_counts; a.pending_slot = ch->d_pending;
    a.iter0 = ch->iter_done + 1; a.nIter = nIter; a.mode = mode;
    a.recs_in = (mode == 1) ? (const tonga_proposal *)(d + o_rec) : nullptr;
    a.recs_out = (mode == 0 && recs) ? (tonga_proposal *)(d + o_rec) : nullptr;
    a.tr_accept = tr_accept ? (int8_t *)(d + o_acc) : nullptr;
    a.tr_phi = tr_phi ? (double *)(d + o_phi) : nullptr;
    a.tr_K = tr_K ? (int32_t *)(d + o_K) : nullptr;
    a.seed = ch->seed; a.chain_id0 = ch->chain_id0;
    a.hist_cap = ch->hist_cap; a.n_hist = ch->d_n_hist; a.model_num = ch->d_model_num;
    a.hist_K = ch->d_hist_K; a.hist_cells = ch->d_hist_cells; a.hist_phi = ch->d_hist_phi; a.hist_ptS = ch->d_hist_ptS;
    a.hist_iter = ch->d_hist_iter; a.hist_action = ch->d_hist_action; a.hist_accept = ch->d_hist_accept; a.hist_next = ch->d_hist_next;

    TG_CUDA(cudaEventRecord(ch->ev0, s));
    tg::tg_sampler_kernel<<<ch->n, tg::ST, ch->smem, s>>>(a);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaEventRecord(ch->ev1, s));
    ch->iter_done += nIter;
    if (mode == 0 && recs) TG_CUDA(cudaMemcpyAsync(recs, d + o_rec, sizeof(tonga_proposal) * N, cudaMemcpyDeviceToHost, s));
    if (tr_accept) TG_CUDA(cudaMemcpyAsync(tr_accept, d + o_acc, N, cudaMemcpyDeviceToHost, s));
    if (tr_phi) TG_CUDA(cudaMemcpyAsync(tr_phi, d + o_phi, 8 * N, cudaMemcpyDeviceToHost, s));
    if (tr_K) TG_CUDA(cudaMemcpyAsync(tr_K, d + o_K, 4 * N, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    TG_CUDA(cudaEventElapsedTime(&ch->last_ms, ch->ev0, ch->ev1));
    return TONGA_OK;
}

extern "C" int tonga_chains_last_kernel_ms(tonga_chains *ch, float *ms) {
    if (!ch || !ms) return tg::fail(TONGA_ERR_ARG, "tonga_chains_last_kernel_ms: NULL");
    *ms = ch->last_ms;
    return TONGA_OK;
}

extern "C" int tonga_chains_reset(tonga_chains *ch) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_reset: NULL");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    cudaStream_t s = ch->ctx->stream;
    const size_t n = (size_t)ch->n;
    TG_CUDA(cudaMemsetAsync(ch->d_counts, 0, 8 * n * 15, s));
    TG_CUDA(cudaMemsetAsync(ch->d_pending, 0xFF, 4 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_n_hist, 0, 4 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_model_num, 0, 8 * n, s));
    ch->iter_done = 0;
    return TONGA_OK;
}

extern "C" int tonga_chains_get_state(tonga_chains *ch, int32_t Kcap, int32_t *K, double *cells, double *phi, double *ptS,
                                      double *noise, int32_t *owners) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_state: NULL");
    if (!ch->have_models) return tg::fail(TONGA_ERR_STATE, "tonga_chains_get_state: no models yet");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    TG_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t n = (size_t)ch->n, KC = (size_t)ch->KC, R = (size_t)ctx->R, Rp = (size_t)ch->Rp, P = (size_t)ctx->P, Pp = (size_t)ctx->Ppad;
    std::vector<int32_t> hK(n);
    TG_CUDA(cudaMemcpy(hK.data(), ch->d_K, 4 * n, cudaMemcpyDeviceToHost));
    if (K) std::memcpy(K, hK.data(), 4 * n);
    if (cells) {
        if (Kcap < 1) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_state: Kcap < 1");
        std::vector<double> hc(n * 4 * KC);
        TG_CUDA(cudaMemcpy(hc.data(), ch->d_cells, 8 * n * 4 * KC, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) {
            if (hK[i] > Kcap) return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_get_state: Kcap smaller than a chain's nCells");
            for (int a = 0; a < 4; a++) std::memcpy(cells + (i * 4 + a) * (size_t)Kcap, &hc[(i * 4 + a) * KC], 8 * (size_t)hK[i]);
        }
    }
    if (phi) TG_CUDA(cudaMemcpy(phi, ch->d_phi, 8 * n, cudaMemcpyDeviceToHost));
    if (noise) TG_CUDA(cudaMemcpy(noise, ch->d_noise, 8 * n, cudaMemcpyDeviceToHost));
    if (ptS) TG_CUDA(cudaMemcpy2D(ptS, 8 * R, ch->d_tstar, 8 * Rp, 8 * R, n, cudaMemcpyDeviceToHost));
    if (owners) {
        std::vector<uint8_t> ho(n * Pp);
        TG_CUDA(cudaMemcpy(ho.data(), ch->d_owner, n * Pp, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++)
            for (size_t p = 0; p < P; p++) {
                const uint8_t o = ho[i * Pp + p];
                owners[i * P + p] = (o == TG_OWNER_NONE) ? -1 : (int32_t)o;
            }
    }
    return TONGA_OK;
}

extern "C" int tonga_chains_get_stats(tonga_chains *ch, int64_t *iter, int64_t *counts) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_stats: NULL");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    TG_CUDA(cudaStreamSynchronize(ch->ctx->stream));
    if (iter) *iter = ch->iter_done;
    if (counts) TG_CUDA(cudaMemcpy(counts, ch->d_counts, 8 * (size_t)ch->n * 15, cudaMemcpyDeviceToHost));
    return TONGA_OK;
}

extern "C" int tonga_chains_get_history(tonga_chains *ch, int32_t Kcap, int32_t *n_hist, int32_t *hist_K, double *hist_cells,
                                        double *hist_phi, double *hist_ptS, int64_t *hist_iter, int32_t *hist_action,
                                        int32_t *hist_accept, int32_t *hist_next_action) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_history: NULL");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    TG_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t n = (size_t)ch->n, KC = (size_t)ch->KC, R = (size_t)ctx->R, H = (size_t)ch->hist_cap, nh = n * H;
    if (n_hist) TG_CUDA(cudaMemcpy(n_hist, ch->d_n_hist, 4 * n, cudaMemcpyDeviceToHost));
    if (nh == 0) return TONGA_OK;
    if (hist_K) TG_CUDA(cudaMemcpy(hist_K, ch->d_hist_K, 4 * nh, cudaMemcpyDeviceToHost));
    if (hist_cells) {
        if (Kcap == (int)KC) {
            TG_CUDA(cudaMemcpy(hist_cells, ch->d_hist_cells, 8 * nh * 4 * KC, cudaMemcpyDeviceToHost));
        } else {
            if (Kcap < 1) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_history: Kcap < 1");
            std::vector<double> hc(nh * 4 * KC);
            TG_CUDA(cudaMemcpy(hc.data(), ch->d_hist_cells, 8 * nh * 4 * KC, cudaMemcpyDeviceToHost));
            const size_t kc = std::min((size_t)Kcap, KC);
            for (size_t i = 0; i < nh * 4; i++) std::memcpy(hist_cells + i * (size_t)Kcap, &hc[i * KC], 8 * kc);
        }
    }
    if (hist_phi) TG_CUDA(cudaMemcpy(hist_phi, ch->d_hist_phi, 8 * nh, cudaMemcpyDeviceToHost));
    if (hist_ptS) TG_CUDA(cudaMemcpy(hist_ptS, ch->d_hist_ptS, 8 * nh * R, cudaMemcpyDeviceToHost));
    if (hist_iter) TG_CUDA(cudaMemcpy(hist_iter, ch->d_hist_iter, 8 * nh, cudaMemcpyDeviceToHost));
    if (hist_action) TG_CUDA(cudaMemcpy(hist_action, ch->d_hist_action, 4 * nh, cudaMemcpyDeviceToHost));
    if (hist_accept) TG_CUDA(cudaMemcpy(hist_accept, ch->d_hist_accept, 4 * nh, cudaMemcpyDeviceToHost));
    if (hist_next_action) TG_CUDA(cudaMemcpy(hist_next_action, ch->d_hist_next, 4 * nh, cudaMemcpyDeviceToHost));
    return TONGA_OK;
}

extern "C" int tonga_chains_verify(tonga_chains *ch, int64_t *owner_mismatch, double *max_dphi, double *max_dts) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_verify: NULL");
    if (!ch->have_models) return tg::fail(TONGA_ERR_STATE, "tonga_chains_verify: no models yet");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int rc = tg::launch_evaluate(ctx, ch->n, ch->KC, ch->d_K, ch->d_cells, ch->d_noise, ch->d_ptS_tmp, ch->d_phi_tmp, nullptr, ch->d_owner_tmp);
    if (rc != TONGA_OK) return rc;
    TG_CUDA(cudaMemsetAsync(ch->d_mism, 0, 8, s));
    TG_CUDA(cudaMemsetAsync(ch->d_maxd, 0, 16, s));
    dim3 grid(8, ch->n);
    tg::tg_verify_kernel<<<grid, 256, 0, s>>>(ch->n, ctx->P, ctx->Ppad, ctx->R, ch->Rp, ch->d_owner, ch->d_owner_tmp, ch->d_tstar, ch->d_ptS_tmp,
                                              ch->d_phi, ch->d_phi_tmp, ch->d_mism, ch->d_maxd);
    TG_CUDA(cudaGetLastError());
    unsigned long long mm = 0;
    double md[2] = {0, 0};
    TG_CUDA(cudaMemcpyAsync(&mm, ch->d_mism, 8, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaMemcpyAsync(md, ch->d_maxd, 16, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    if (owner_mismatch) *owner_mismatch = (int64_t)mm;
    if (max_dphi) *max_dphi = md[0];
    if (max_dts) *max_dts = md[1];
    return TONGA_OK;
}

extern "C" int tonga_chains_kcap(const tonga_chains *ch) { return ch ? ch->KC : 0; }

extern "C" int tonga_chains_device_ptrs(tonga_chains *ch, void **n_hist, void **hist_K, void **hist_cells, void **hist_phi,
                                        void **hist_ptS, void **state_K, void **state_cells, void **state_phi) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_device_ptrs: NULL");
    if (n_hist) *n_hist = ch->d_n_hist;
    if (hist_K) *hist_K = ch->d_hist_K;
    if (hist_cells) *hist_cells = ch->d_hist_cells;
    if (hist_phi) *hist_phi = ch->d_hist_phi;
    if (hist_ptS) *hist_ptS = ch->d_hist_ptS;
    if (state_K) *state_K = ch->d_K;
    if (state_cells) *state_cells = ch->d_cells;
    if (state_phi) *state_phi = ch->d_phi;
    return TONGA_OK;
}
