// sampler_kernel.cuh -- tg_sampler_kernel: the device-resident proposal loop (TD_inversion_function.jl:70-302).
// See sampler.cu for the overview.  One CTA of 4 warps = one chain; the chain's state lives in shared memory for the whole launch.
// Nothing in an iteration is done by one warp while the others wait:
//   A  EVERY warp assembles the proposal redundantly from the pre-generated raw draw (tg_pregen_kernel) or the replayed record:
//      a-priori checks, v_nearest at the new / killed nucleus -- identical registers in all warps, no broadcast, no barrier;
//   B  every warp owns the 128-point blocks b = warp, warp+4, ...:
//        B1 flat pass, 4 points per lane: birth/move screen d(p,new) against the cached d(p,owner) in FP32 (exact FP64 inside the
//           error band); death/move flag the points of the killed/moved nucleus ("orphans"); a ballot per block records which
//           4-point words changed (no per-point ray ids, no atomics);
//        B2 the warp compacts its orphan WORDS into a queue and rescans them 32 words at a time (4 points per lane, packed FP32,
//           nucleus index carried in the low mantissa bits so that best / second best are integer min / max);
//   --- barrier ---
//   C  warp w owns the (length-sorted) rays r = w (mod 4): it finds its touched rays from the word bitmap and re-integrates
//      them in the canonical order (tstar_g8: 8 lanes per ray, 4 rays at a time); t* of ray 128c + 4l + w ends in a register of
//      lane l -- the thread that needs it for phi and for the commit;
//   --- barrier ---
//   D  canonical phi: per-thread partial sums, one exchange of 4 warp sums  --- barrier ---
//   E  EVERY thread evaluates the acceptance rule;  F  every warp commits / rolls back its own blocks and rays;
//   --- barrier ---  G  counters, traces, thinning, history.
#pragma once
#include <type_traits>

#include "proposal.cuh"
#include "tonga_internal.cuh"

namespace tg {

struct SamplerArgs {
    // geometry (device order: rays sorted by length)
    const double *px, *py, *pz, *dt, *tS, *sig;
    const float *pxf, *pyf, *pzf;   // fl32 copies (screening)
    float tol_alpha, tol_beta2;     // screening band
    int exact_only;                 // 1 = skip the FP32 screening (pure FP64 path; used by tests)
    long long *prof;                // PROF instantiation only: [n][16] per-phase clock64 totals of thread 0
    const int32_t *ray_off, *ray_rank;
    int R, Rp, KC;
    int P, Ppad;
    const int32_t *perm;  // CTA -> chain: chains sorted by expected cost so that every SM hosts the same mix (NULL = identity)
    tonga_params prm;
    // chain state (global)
    int32_t *K;
    double *cells;  // [n][4][KC]
    double *phi, *noise, *beta;
    uint8_t *owner;  // [n][Ppad]
    float *dcache;   // [n][Ppad] fl32 squared distance of every point to its current owner (1e9 = none); L2-resident
    double *tstar;   // [n][Rp]  (sorted ray order)
    long long *counts;  // [n][3][5] proposed / accepted / evaluated
    int32_t *pending_slot;  // [n] history slot awaiting its next_action, or -1
    // run
    long long iter0, nIter;       // first iteration number and iteration count of THIS launch
    long long it0, trace_stride;  // records / traces are indexed [chain * trace_stride + it0 + it]
    int mode;  // 0 generate, 1 replay
    const RawDraw *raw;  // mode 0: [n][nIter] pre-generated raw draws (tg_pregen_kernel)
    const tonga_proposal *recs_in;
    tonga_proposal *recs_out;
    int8_t *tr_accept;
    double *tr_phi;
    int32_t *tr_K;
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

constexpr int SQ_CAP = 64;    // orphan-word queue entries per warp
constexpr int ZLUT_TAG = 128; // index of the "tagged byte" entry of the zeta look-up tables

struct SmemLayout {
    size_t o_owner, o_mask, o_tstar, o_nuc, o_nucf, o_zp, o_zqp, o_dirtyw, o_perm, o_queue, o_scr, o_cnt, o_bar, total;
};
__host__ __device__ inline SmemLayout smem_layout(int Ppad, int Rp, int KC) {
    SmemLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
    L.o_owner = take((size_t)Ppad);
    L.o_mask = take((size_t)Ppad / 8);
    L.o_tstar = take(8 * (size_t)Rp);
    L.o_nuc = take(8 * 4 * (size_t)KC);
    L.o_nucf = take(4 * 3 * (size_t)KC);   // fl32 nuclei, SoA with stride KC; unused slots and a killed nucleus hold +inf
    L.o_zp = take(8 * (ZLUT_TAG + 1));     // owner byte (clamped to 128) -> zeta under the PROPOSED model; [127] = 0 (none), [128] = tagged
    L.o_zqp = take(8 * (ZLUT_TAG + 1));    // ... -> zeta / 1000 (correctly rounded): the term of a segment inside one cell is dt * zq
    L.o_dirtyw = take(4 * (size_t)(Ppad / 128));  // bit per 4-point word: some point of the word changes owner / zeta under the proposal
    L.o_perm = take((size_t)(ST / 32) * 32);
    L.o_queue = take((size_t)(ST / 32) * SQ_CAP * 2);
    L.o_scr = take(8 * 8);
    L.o_cnt = take(4 * 16);
    L.o_bar = take(8);
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

// ---- raw draws for a whole launch: one thread per (chain, iteration) ------------------------------------------------
__global__ void __launch_bounds__(256) tg_pregen_kernel(int n, long long nIter, long long iter0, unsigned long long seed, long long chain_id0,
                                                         int n_actions, RawDraw *__restrict__ raw) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)n * nIter) return;
    const long long chain = i / nIter, it = i - chain * nIter;
    RawDraw rd;
    raw_draw_thread(rd, seed, iter0 + it, (unsigned long long)(chain_id0 + chain), n_actions);
    double *o = reinterpret_cast<double *>(raw + i);
    const double v[8] = {rd.act, rd.u1, rd.u2, rd.u3, rd.n0, rd.n1, rd.n2, rd.u7};
#pragma unroll
    for (int k = 0; k < 8; k += 2) __stcs(reinterpret_cast<double2 *>(o + k), make_double2(v[k], v[k + 1]));
}

// Orphan rescan, exact FP64: nearest nucleus of flat point p among the K nuclei except `skip` (strict <, ascending index,
// MCsub.jl:252-259); nucleus `mvi` (a proposed move) is taken at (cx, cy, cz).
__device__ __noinline__ int rescan_point(const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
                                         const double *nx, const double *ny, const double *nz, int K, int skip, int mvi, double cx, double cy,
                                         double cz, int p) {
    const double x = px[p], y = py[p], z = pz[p];
    double best = 1e9;
    int bi = TG_OWNER_NONE;
#pragma unroll 2
    for (int i = 0; i < K; i++) {
        const bool mv = (i == mvi);
        const double d = dist2_exact(mv ? cx : nx[i], mv ? cy : ny[i], mv ? cz : nz[i], x, y, z);
        if (d < best && i != skip) { best = d; bi = i; }
    }
    return bi;
}

template <int NCH, bool PROF>
__global__ void __launch_bounds__(ST, 7) tg_sampler_kernel(const SamplerArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout L = smem_layout(a.Ppad, a.Rp, a.KC);
    uint8_t *s_owner = smem + L.o_owner;
    uint32_t *s_own32 = reinterpret_cast<uint32_t *>(s_owner);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + L.o_mask);
    double *s_tstar = reinterpret_cast<double *>(smem + L.o_tstar);
    double *s_nx = reinterpret_cast<double *>(smem + L.o_nuc);
    double *s_ny = s_nx + a.KC, *s_nz = s_ny + a.KC, *s_zeta = s_nz + a.KC;
    float *s_fx = reinterpret_cast<float *>(smem + L.o_nucf);
    float *s_fy = s_fx + a.KC, *s_fz = s_fy + a.KC;
    double *s_zp = reinterpret_cast<double *>(smem + L.o_zp);
    double *s_zqp = reinterpret_cast<double *>(smem + L.o_zqp);
    uint32_t *s_dirtyw = reinterpret_cast<uint32_t *>(smem + L.o_dirtyw);
    double *s_scr = reinterpret_cast<double *>(smem + L.o_scr);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + L.o_cnt);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.o_bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, sub = lane & 7;
    const uint32_t FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    uint8_t *s_perm = smem + L.o_perm + warp * 32;
    uint16_t *s_queue = reinterpret_cast<uint16_t *>(smem + L.o_queue) + warp * SQ_CAP;
    const int chain = a.perm ? a.perm[blockIdx.x] : (int)blockIdx.x;  // launch order = cost order (tg_order_kernel)
    const int KC = a.KC, R = a.R;
    float *__restrict__ dcache = a.dcache + (size_t)chain * a.Ppad;
    const int nMaskWords = a.Ppad / 32;
    const int nBlocks = a.Ppad / 128;
    const float FINF = __int_as_float(0x7f800000);

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
    if (tid < 16) s_cnt[tid] = 0u;
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            s2u(s_bar))
        : "memory");
    __syncthreads();
    int K = a.K[chain];
    {   // fl32 copies of the nuclei (slots >= K hold +inf so that they never win a comparison) and the zeta tables
        for (int i = tid; i < 3 * KC; i += ST) {
            const int k = i % KC;
            s_fx[i] = (k < K) ? (float)s_nx[i] : FINF;
        }
        for (int o = tid; o <= ZLUT_TAG; o += ST) {
            const double zv = (o < K) ? s_zeta[o] : 0.0;
            s_zp[o] = zv;
            s_zqp[o] = div1000_exact(zv);
        }
    }
    // the rays of this thread: r = 128 c + 4 lane + warp (phi_ray); start and length packed in one register per chunk
    uint32_t ri[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int r = phi_ray(c, tid);
        ri[c] = 0u;
        if (r < R) {
            const int q0 = a.ray_off[r], n = a.ray_off[r + 1] - q0;
            ri[c] = (uint32_t)q0 | ((uint32_t)n << 18);
        }
    }
    __syncthreads();

    double phi = a.phi[chain];
    double noise = a.noise[chain];
    const double beta = a.beta[chain];
    int n_hist = a.n_hist[chain];
    long long model_num = a.model_num[chain];
    int pending_slot = a.pending_slot[chain];

    const tonga_params &pm = a.prm;
    const double sig_zeta = pm.zeta_scale * pm.sig / 100;  // TD_inversion_function.jl:22
    const float ta = a.tol_alpha, tb = a.tol_beta2;
    // B2 carries the nucleus index in the 7 low mantissa bits of the fl32 distance (truncation: relative 2^-16): widen the band
    const float ta2 = ta * (1.0f + 0x1.0p-16f) + 0x1.0p-16f + 0x1.0p-20f;

    long long pt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = PROF ? clock64() : 0;
    auto tick = [&](int ph) {
        if (PROF && tid == 0) { const long long t = clock64(); pt[ph] += t - tprev; tprev = t; }
    };
    // raw draw of the next iteration: lane l < 8 holds double l of the record (requested one iteration ahead)
    const double *rawp = (a.mode == 0) ? reinterpret_cast<const double *>(a.raw + (size_t)chain * a.nIter) + (lane & 7) : nullptr;
    double raw_next = (a.mode == 0 && a.nIter > 0) ? __ldcs(rawp) : 0.0;

#pragma unroll 1
    for (long long it = 0; it < a.nIter; it++) {
        const long long iter = a.iter0 + it;
        // ================================================================ A: the proposal, assembled by every warp
        Prop pr;
        pr.do_eval = 0; pr.accept = 0; pr.idx = 0; pr.action = 0;
        pr.x = pr.y = pr.z = pr.zeta = pr.u = pr.aux = pr.ox = pr.oy = pr.oz = pr.ztag = 0.0;
        {
            double u1 = 0, u2 = 0, u3 = 0, n0 = 0, n1 = 0, n2 = 0, u7 = 0;
            if (a.mode == 0) {
                const double rv = raw_next;
                if (it + 1 < a.nIter) raw_next = __ldcs(rawp + 8 * (it + 1));
                pr.action = (int)__shfl_sync(FULL, rv, 0);
                u1 = __shfl_sync(FULL, rv, 1); u2 = __shfl_sync(FULL, rv, 2); u3 = __shfl_sync(FULL, rv, 3);
                n0 = __shfl_sync(FULL, rv, 4); n1 = __shfl_sync(FULL, rv, 5); n2 = __shfl_sync(FULL, rv, 6);
                u7 = __shfl_sync(FULL, rv, 7);
            } else {
                const tonga_proposal rec = a.recs_in[(size_t)chain * a.trace_stride + a.it0 + it];
                pr.action = rec.action; pr.idx = rec.idx; pr.x = rec.x; pr.y = rec.y; pr.z = rec.z; pr.zeta = rec.zeta; pr.u = rec.u;
            }
            assemble_proposal<0>(pr, a.mode, u1, u2, u3, n0, n1, n2, u7, s_nx, s_ny, s_nz, s_zeta, K, noise, pm, sig_zeta, lane);
        }
        const int act = pr.action, do_eval = pr.do_eval, pidx = pr.idx;
        const double zold = (do_eval && (act == 2 || act == 3)) ? s_zeta[pidx] : 0.0;  // zeta[idx] of the CURRENT model (acceptance rule)
        if (tid == 0) {
            if (a.mode == 0 && a.recs_out) {
                tonga_proposal rec;
                rec.action = pr.action; rec.idx = pr.idx; rec.x = pr.x; rec.y = pr.y; rec.z = pr.z; rec.zeta = pr.zeta; rec.u = pr.u;
                a.recs_out[(size_t)chain * a.trace_stride + a.it0 + it] = rec;
            }
            if (pending_slot >= 0) a.hist_next[(size_t)chain * a.hist_cap + pending_slot] = pr.action;
            // zeta tables under the proposal (read by phase C only, i.e. after the barrier below)
            if (do_eval) {
                if (act == 1) { s_zp[ZLUT_TAG] = pr.zeta; s_zqp[ZLUT_TAG] = div1000_exact(pr.zeta); }
                else if (act == 4) { s_zp[ZLUT_TAG] = s_zp[pidx]; s_zqp[ZLUT_TAG] = s_zqp[pidx]; }
                else if (act == 3) { s_zp[pidx] = pr.zeta; s_zqp[pidx] = div1000_exact(pr.zeta); }
            }
        }
        pending_slot = -1;
        tick(0);
        const uint32_t kk = (uint32_t)pidx * 0x01010101u;
        double phin = phi;
        int accepted = 0;
        double tn[NCH];        // proposal's t* of this thread's touched rays
        uint32_t dmask[NCH];   // per chunk: which lanes of this warp hold a touched ray
#pragma unroll
        for (int c = 0; c < NCH; c++) { tn[c] = 0.0; dmask[c] = 0u; }

        if (do_eval) {
            unsigned long long blkmask = 0ull;  // own blocks holding orphans (bit = warp-iteration; iterations >= 64 are always scanned)
            unsigned long long tagmask = 0ull;  // own blocks holding tagged bytes
            const double cx = pr.x, cy = pr.y, cz = pr.z;
            const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
            if (act != 5) {
                // the fl32 nuclei under the proposal: every warp writes the same values before its own use (the FP64 arrays stay untouched)
                if (lane == 0) {
                    if (act == 2) s_fx[pidx] = FINF;  // the screening must not see the killed nucleus
                    else if (act == 4) { s_fx[pidx] = cxf; s_fy[pidx] = cyf; s_fz[pidx] = czf; }
                }
                __syncwarp();
                // ======================================================== B1: flat pass over this warp's 128-point blocks
                // birth / move: FP32 screening of d(p,new) against d(p,owner); ACT is a compile-time constant so that the birth
                // path carries no move logic.  Branch-free per point; near ties (inside the error band) go to the exact FP64 compare.
                auto b1_switch = [&](auto actc) {
                    constexpr int ACT = decltype(actc)::value;
                    const int mv = (ACT == 4) ? pidx : -1;
                    const float2 ncx = make_float2(-cxf, -cxf), ncy = make_float2(-cyf, -cyf), ncz = make_float2(-czf, -czf);
                    int bi = 0;
                    // software pipeline: the next block's coordinates are in flight while the current block is screened
                    float4 nxf = *reinterpret_cast<const float4 *>(a.pxf + 4 * (warp * 32 + lane));
                    float4 nyf = *reinterpret_cast<const float4 *>(a.pyf + 4 * (warp * 32 + lane));
                    float4 nzf = *reinterpret_cast<const float4 *>(a.pzf + 4 * (warp * 32 + lane));
                    float4 ndo = *reinterpret_cast<const float4 *>(dcache + 4 * (warp * 32 + lane));
#pragma unroll 1
                    for (int blk = warp; blk < nBlocks; blk += ST / 32, bi++) {
                        const int w = blk * 32 + lane;
                        const uint32_t ow = s_own32[w];
                        uint32_t tags = 0, amb = 0, mbits = 0;
                        const float4 xf = nxf, yf = nyf, zf = nzf;
                        const float4 dof = ndo;
                        if (blk + ST / 32 < nBlocks) {
                            const int wn = w + (ST / 32) * 32;
                            nxf = *reinterpret_cast<const float4 *>(a.pxf + 4 * wn);
                            nyf = *reinterpret_cast<const float4 *>(a.pyf + 4 * wn);
                            nzf = *reinterpret_cast<const float4 *>(a.pzf + 4 * wn);
                            ndo = *reinterpret_cast<const float4 *>(dcache + 4 * wn);
                        }
                        if (!a.exact_only) {
                            float2 ex = __fadd2_rn(make_float2(xf.x, xf.y), ncx), ey = __fadd2_rn(make_float2(yf.x, yf.y), ncy), ez = __fadd2_rn(make_float2(zf.x, zf.y), ncz);
                            const float2 dc01 = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                            ex = __fadd2_rn(make_float2(xf.z, xf.w), ncx); ey = __fadd2_rn(make_float2(yf.z, yf.w), ncy); ez = __fadd2_rn(make_float2(zf.z, zf.w), ncz);
                            const float2 dc23 = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                            const float DC[4] = {dc01.x, dc01.y, dc23.x, dc23.y};
                            const float DO[4] = {dof.x, dof.y, dof.z, dof.w};  // cached distance to the current owner: no nucleus gather
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const float d_o = DO[q];
                                const float diff = DC[q] - d_o;
                                const float tol = fmaf(ta, DC[q] + d_o, tb);
                                bool sw = diff < -tol, am = fabsf(diff) <= tol;
                                if (ACT == 4) {
                                    const bool mine = ((int)((ow >> (8 * q)) & 0xFF) == mv);  // move, type A: owned by the moved nucleus -> rescan in B2
                                    mbits |= mine ? (1u << q) : 0u;
                                    sw = sw && !mine;
                                    am = am && !mine;
                                }
                                tags |= sw ? (0x80u << (8 * q)) : 0u;
                                amb |= am ? (1u << q) : 0u;
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                if ((int)((ow >> (8 * q)) & 0xFF) == mv) mbits |= 1u << q;
                                else amb |= 1u << q;
                            }
                        }
                        if (amb) {  // exact FP64 comparison, MCsub.jl:254-255 semantics
#pragma unroll 1
                            for (int q = 0; q < 4; q++) {
                                if (!((amb >> q) & 1u)) continue;
                                const int o = (ow >> (8 * q)) & 0xFF;
                                const int p = 4 * w + q;
                                const double x = a.px[p], y = a.py[p], z = a.pz[p];
                                const double d_o = (o == TG_OWNER_NONE) ? 1e9 : dist2_exact(s_nx[o], s_ny[o], s_nz[o], x, y, z);
                                const double d_c = dist2_exact(cx, cy, cz, x, y, z);
                                // birth: the new nucleus has the highest index -> strict <.  move: index mv also wins exact ties against o > mv.
                                const bool sw = (d_c < d_o) || (ACT == 4 && d_c == d_o && mv < o && o != TG_OWNER_NONE);
                                if (sw) tags |= 0x80u << (8 * q);
                            }
                        }
                        if (tags) s_own32[w] = ow | tags;
                        const uint32_t dm = __ballot_sync(FULL, tags != 0u);
                        if (lane == 0) s_dirtyw[blk] = dm;
                        if (dm && bi < 64) tagmask |= 1ull << bi;
                        if (ACT == 4 && __any_sync(FULL, mbits != 0u)) {  // assemble the block's 4 mask words (8 lanes x 4 bits each)
                            uint32_t nib = mbits << (sub * 4);
                            nib |= __shfl_xor_sync(FULL, nib, 1);
                            nib |= __shfl_xor_sync(FULL, nib, 2);
                            nib |= __shfl_xor_sync(FULL, nib, 4);
                            if (sub == 0) s_mask[w >> 3] = nib;
                            if (bi < 64) blkmask |= 1ull << bi;
                        }
                    }
                };
                if (act == 1) b1_switch(std::integral_constant<int, 1>{});
                else if (act == 4) b1_switch(std::integral_constant<int, 4>{});
                else {  // death: flag the orphans; change: mark the words of the cell
                    int bi = 0;
#pragma unroll 1
                    for (int blk = warp; blk < nBlocks; blk += ST / 32, bi++) {
                        const int w = blk * 32 + lane;
                        const uint32_t eq = __vcmpeq4(s_own32[w], kk);  // bytes owned by the killed / changed nucleus
                        const uint32_t dm = __ballot_sync(FULL, eq != 0u);
                        if (lane == 0) s_dirtyw[blk] = dm;
                        if (act == 2 && dm) {
                            const uint32_t mbits = (eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u);
                            uint32_t nib = mbits << (sub * 4);
                            nib |= __shfl_xor_sync(FULL, nib, 1);
                            nib |= __shfl_xor_sync(FULL, nib, 2);
                            nib |= __shfl_xor_sync(FULL, nib, 4);
                            if (sub == 0) s_mask[w >> 3] = nib;
                            if (bi < 64) blkmask |= 1ull << bi;
                        }
                    }
                }
                tick(8);  // profile slot 8 = B1, slot 1 = B2 (+ barrier)
                // ======================================================== B2: rescan this warp's orphan words, 32 words at a time
                if (act == 2 || act == 4) {
                    __syncwarp();
                    const int skip = (act == 2) ? pidx : -1, mvi = (act == 4) ? pidx : -1;
                    const int Kr = (K + 1) & ~1;  // nuclei are visited in pairs; slot K (if any) holds +inf
                    int qn = 0, bi = 0, blk = warp;
#pragma unroll 1
                    for (;;) {
                        const bool last = blk >= nBlocks;
                        if (!last && (bi >= 64 || ((blkmask >> bi) & 1ull))) {
                            const int w = blk * 32 + lane;
                            const bool has = ((s_mask[w >> 3] >> (sub * 4)) & 0xFu) != 0u;
                            const uint32_t m = __ballot_sync(FULL, has);
                            if (has) s_queue[qn + __popc(m & lt_mask)] = (uint16_t)w;
                            qn += __popc(m);
                            __syncwarp();
                        }
                        if (qn >= 32 || (last && qn > 0)) {  // drain: lanes < cnt take the top `cnt` queue entries, one 4-point word each
                            const int cnt = qn < 32 ? qn : 32;
                            if (lane < cnt) {
                                const int w = (int)s_queue[qn - cnt + lane];
                                const uint32_t mb = (s_mask[w >> 3] >> ((w & 7) * 4)) & 0xFu;
                                const uint32_t ow = s_own32[w];
                                uint32_t nb[4] = {TG_OWNER_NONE, TG_OWNER_NONE, TG_OWNER_NONE, TG_OWNER_NONE};
                                uint32_t need = a.exact_only ? mb : 0u;
                                if (!a.exact_only) {
                                    const float4 X = *reinterpret_cast<const float4 *>(a.pxf + 4 * w), Y = *reinterpret_cast<const float4 *>(a.pyf + 4 * w),
                                                 Z = *reinterpret_cast<const float4 *>(a.pzf + 4 * w);
                                    const float2 X01 = make_float2(X.x, X.y), X23 = make_float2(X.z, X.w), Y01 = make_float2(Y.x, Y.y), Y23 = make_float2(Y.z, Y.w),
                                                 Z01 = make_float2(Z.x, Z.y), Z23 = make_float2(Z.z, Z.w);
                                    const uint32_t INIT = (__float_as_uint(1e9f) & 0xFFFFFF80u) | TG_OWNER_NONE;
                                    uint32_t d1[4] = {INIT, INIT, INIT, INIT}, d2[4] = {INIT, INIT, INIT, INIT};
#pragma unroll 1
                                    for (int i = 0; i < Kr; i += 2) {
                                        const float2 fx = *reinterpret_cast<const float2 *>(s_fx + i), fy = *reinterpret_cast<const float2 *>(s_fy + i),
                                                     fz = *reinterpret_cast<const float2 *>(s_fz + i);
#pragma unroll
                                        for (int u = 0; u < 2; u++) {
                                            const float ax = -(u ? fx.y : fx.x), ay = -(u ? fy.y : fy.x), az = -(u ? fz.y : fz.x);
                                            const float2 nax = make_float2(ax, ax), nay = make_float2(ay, ay), naz = make_float2(az, az);
                                            float2 ex = __fadd2_rn(X01, nax), ey = __fadd2_rn(Y01, nay), ez = __fadd2_rn(Z01, naz);
                                            const float2 da = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                                            ex = __fadd2_rn(X23, nax); ey = __fadd2_rn(Y23, nay); ez = __fadd2_rn(Z23, naz);
                                            const float2 db = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                                            const float d[4] = {da.x, da.y, db.x, db.y};
#pragma unroll
                                            for (int q = 0; q < 4; q++) {  // non-negative floats order like their bits: integer min / max, index in the low bits
                                                const uint32_t dp = (__float_as_uint(d[q]) & 0xFFFFFF80u) | (uint32_t)(i + u);
                                                const uint32_t t = max(d1[q], dp);
                                                d1[q] = min(d1[q], dp);
                                                d2[q] = min(d2[q], t);
                                            }
                                        }
                                    }
#pragma unroll
                                    for (int q = 0; q < 4; q++) {
                                        const float f1 = __uint_as_float(d1[q] & 0xFFFFFF80u), f2 = __uint_as_float(d2[q] & 0xFFFFFF80u);
                                        const float tol = fmaf(ta2, f1 + f2, tb);
                                        nb[q] = d1[q] & 0x7Fu;
                                        if (!(f2 - f1 > tol)) need |= 1u << q;  // ambiguous (also: nothing within the 1e9 threshold, NaN coordinates)
                                    }
                                    need &= mb;
                                }
                                if (need) {
#pragma unroll 1
                                    for (int q = 0; q < 4; q++)
                                        if ((need >> q) & 1u) nb[q] = (uint32_t)rescan_point(a.px, a.py, a.pz, s_nx, s_ny, s_nz, K, skip, mvi, cx, cy, cz, 4 * w + q);
                                }
                                uint32_t nw = ow, chg = 0u;
#pragma unroll
                                for (int q = 0; q < 4; q++)
                                    if ((mb >> q) & 1u) {
                                        nw = (nw & ~(0xFFu << (8 * q))) | (nb[q] << (8 * q));  // death: old numbering, renumbered on accept
                                        chg |= (act == 2 || (int)nb[q] != pidx) ? 1u : 0u;   // a point that stays with the moved nucleus keeps its zeta
                                    }
                                s_own32[w] = nw;
                                if (chg) atomicOr(&s_dirtyw[w >> 5], 1u << (w & 31));
                            }
                            qn -= cnt;
                            __syncwarp();
                        }
                        if (last) break;
                        blk += ST / 32; bi++;
                    }
                }
                __syncthreads();
                tick(1);
                // ======================================================== C: t* of this warp's touched rays (canonical order, tstar_g8)
                double tsv[NCH], sgv[NCH];  // observed t* and sigma of this thread's rays, requested now, used in D
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const int r = phi_ray(c, tid);
                    tsv[c] = (r < R) ? a.tS[r] : 0.0;
                    sgv[c] = (r < R) ? a.sig[r] : 1.0;
                }
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    if ((c << 7) < R) {  // warp-uniform
                        const uint32_t info = ri[c];
                        const int q0 = (int)(info & 0x3FFFFu), n = (int)(info >> 18);
                        bool dirty = false;
                        if (n > 0) {  // any changed word among the words the ray touches?
                            const int lo = q0 >> 2, hi = (q0 + n - 1) >> 2;
                            for (int wd = lo >> 5; wd <= (hi >> 5); wd++) {
                                uint32_t bits = s_dirtyw[wd];
                                if (wd == (lo >> 5)) bits &= 0xFFFFFFFFu << (lo & 31);
                                if (wd == (hi >> 5)) bits &= 0xFFFFFFFFu >> (31 - (hi & 31));
                                dirty |= bits != 0u;
                            }
                        }
                        const uint32_t dm = __ballot_sync(FULL, dirty);
                        dmask[c] = dm;
                        const int cnt = __popc(dm), rank = __popc(dm & lt_mask);
                        if (dirty) s_perm[rank] = (uint8_t)lane;
                        __syncwarp();
#pragma unroll 1
                        for (int t = 0; 4 * t < cnt; t++) {  // 4 touched rays at a time, 8 lanes each (ranks follow the length order)
                            const int e = 4 * t + grp;
                            const bool on = e < cnt;
                            const uint32_t inf = __shfl_sync(FULL, info, on ? (int)s_perm[e] : 0);
                            const int tq0 = (int)(inf & 0x3FFFFu), tnp = on ? (int)(inf >> 18) : 0;
                            const int nseg = tnp > 1 ? tnp - 1 : 0;
                            const int trip = __reduce_max_sync(FULL, (nseg + 7) >> 3);
                            const uint8_t *ow = s_owner + tq0;
                            const double *dtp = a.dt + tq0;
                            const double acc = tstar_g8(nseg, trip, sub, [&](int j) -> double {
                                const int oa = min((int)ow[j], ZLUT_TAG), ob = min((int)ow[j + 1], ZLUT_TAG);
                                const double d = dtp[j];
                                // both ends in one cell: 0.5 (z + z) == z exactly, so the term is dt * (z / 1000) with the cell's precomputed quotient
                                return (oa == ob) ? __dmul_rn(d, s_zqp[oa]) : seg_term(d, s_zp[oa], s_zp[ob]);
                            });
                            const double v = __shfl_sync(FULL, acc, (rank & 3) << 3);
                            if (dirty && (rank >> 2) == t) tn[c] = v;
                        }
                        __syncwarp();
                    }
                }
                tick(2);
                // ======================================================== D: phi of the proposed model (canonical order)
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const int r = phi_ray(c, tid);
                    if (r < R) {
                        const double t = ((dmask[c] >> lane) & 1u) ? tn[c] : s_tstar[r];
                        acc = __dadd_rn(acc, misfit_term(t, tsv[c], sgv[c], noise));
                    }
                }
                acc = warp_sum_canonical(acc);
                if (lane == 0) s_scr[(it & 1) * 4 + warp] = acc;
            } else {  // sigma move (extension): every misfit term changes, t* does not
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const int r = phi_ray(c, tid);
                    if (r < R) acc = __dadd_rn(acc, misfit_term(s_tstar[r], a.tS[r], a.sig[r], pr.zeta));
                }
                acc = warp_sum_canonical(acc);
                if (lane == 0) s_scr[(it & 1) * 4 + warp] = acc;
            }
            __syncthreads();
            {
                const double *sc = s_scr + (it & 1) * 4;
                phin = __dadd_rn(__dadd_rn(__dadd_rn(sc[0], sc[1]), sc[2]), sc[3]);
            }
            if (pm.debug_prior) phin = 1.0;  // MCsub.jl:128-136: the chain samples the prior
            // ============================================================ E: acceptance, evaluated by every thread
            accepted = accept_decision(pr, K, phi, phin, zold, noise, beta, R, pm, sig_zeta);
            tick(3);
            // ============================================================ F: commit / roll back, every warp on its own blocks and rays
            if (act == 2 || act == 4) {  // masked bytes: orphans (old owner = pidx)
                int bi = 0;
#pragma unroll 1
                for (int blk = warp; blk < nBlocks; blk += ST / 32, bi++) {
                    if (bi < 64 && !((blkmask >> bi) & 1ull)) continue;
                    const int w = blk * 32 + lane;
                    const uint32_t mb = (s_mask[w >> 3] >> (sub * 4)) & 0xFu;
                    __syncwarp();
                    if (sub == 0) s_mask[w >> 3] = 0u;
                    if (mb) {
                        const uint32_t ow = s_own32[w];
                        if (!accepted) {
                            const uint32_t m8 = ((mb & 1u) * 0xFFu) | ((mb & 2u) * (0xFF00u >> 1)) | ((mb & 4u) * (0xFF0000u >> 2)) | ((mb & 8u) * (0xFF000000u >> 3));
                            s_own32[w] = (ow & ~m8) | (kk & m8);
                        } else {  // refresh the owner-distance cache (a move also changes it for points that stay with the nucleus)
                            const float4 X = *reinterpret_cast<const float4 *>(a.pxf + 4 * w), Y = *reinterpret_cast<const float4 *>(a.pyf + 4 * w),
                                         Z = *reinterpret_cast<const float4 *>(a.pzf + 4 * w);
                            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if ((mb >> q) & 1u) {
                                    const int o = (ow >> (8 * q)) & 0x7F;  // death: still the old numbering, as are the fl32 nuclei
                                    dcache[4 * w + q] = (o == TG_OWNER_NONE) ? 1e9f : dist2_f32(s_fx[o], s_fy[o], s_fz[o], xs[q], ys[q], zs[q]);
                                }
                        }
                    }
                }
            }
            tick(6);
            if (act == 1 || act == 4) {  // tagged bytes: switch to the new / moved nucleus
                const uint32_t newb = (uint32_t)(act == 1 ? K : pidx) * 0x01010101u;
                int bi = 0;
#pragma unroll 1
                for (int blk = warp; blk < nBlocks; blk += ST / 32, bi++) {
                    if (bi < 64 && !((tagmask >> bi) & 1ull)) continue;
                    const int w = blk * 32 + lane;
                    const uint32_t ow = s_own32[w], t = ow & 0x80808080u;
                    if (t) {
                        const uint32_t m = (t >> 7) * 0xFFu;
                        s_own32[w] = accepted ? ((ow & ~m) | (newb & m)) : (ow & 0x7F7F7F7Fu);
                        if (accepted) {  // switched points: cache their distance to the new nucleus
                            const float4 X = *reinterpret_cast<const float4 *>(a.pxf + 4 * w), Y = *reinterpret_cast<const float4 *>(a.pyf + 4 * w),
                                         Z = *reinterpret_cast<const float4 *>(a.pzf + 4 * w);
                            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if ((t >> (8 * q + 7)) & 1u) dcache[4 * w + q] = dist2_f32(cxf, cyf, czf, xs[q], ys[q], zs[q]);
                        }
                    }
                }
            }
            tick(7);
            if (accepted && act != 5) {  // t* of the touched rays
#pragma unroll
                for (int c = 0; c < NCH; c++)
                    if ((dmask[c] >> lane) & 1u) s_tstar[phi_ray(c, tid)] = tn[c];
            }
            if (act == 2 && accepted) {  // deleteat! renumbering: indices above `kill` shift down (:132-135)
#pragma unroll 1
                for (int blk = warp; blk < nBlocks; blk += ST / 32) {
                    const int w = blk * 32 + lane;
                    const uint32_t ow = s_own32[w];
                    const uint32_t gt = __vcmpgtu4(ow, kk) & ~__vcmpeq4(ow, 0x7F7F7F7Fu);
                    if (gt) s_own32[w] = ow - (gt & 0x01010101u);
                }
                __syncthreads();  // the cache refresh above read the nuclei the delete below shifts
                if (warp == 0) {  // order-preserving delete of the nucleus
                    for (int s0 = 0; s0 < K - 1 - pidx; s0 += 32) {
                        const int i = pidx + s0 + lane;
                        double vx = 0, vy = 0, vz = 0, vt = 0, vq = 0;
                        if (i < K - 1) { vx = s_nx[i + 1]; vy = s_ny[i + 1]; vz = s_nz[i + 1]; vt = s_zeta[i + 1]; vq = s_zqp[i + 1]; }
                        __syncwarp();
                        if (i < K - 1) {
                            s_nx[i] = vx; s_ny[i] = vy; s_nz[i] = vz; s_zeta[i] = vt; s_zp[i] = vt; s_zqp[i] = vq;
                            s_fx[i] = (float)vx; s_fy[i] = (float)vy; s_fz[i] = (float)vz;
                        }
                        __syncwarp();
                    }
                    if (lane == 0) {  // freed slot
                        s_fx[K - 1] = s_fy[K - 1] = s_fz[K - 1] = FINF;
                        s_zp[K - 1] = 0.0; s_zqp[K - 1] = 0.0;
                    }
                }
            }
            if (tid == 0) {
                if (act == 1 && accepted) {  // append!, :85-88
                    s_nx[K] = cx; s_ny[K] = cy; s_nz[K] = cz; s_zeta[K] = pr.zeta;
                    s_fx[K] = cxf; s_fy[K] = cyf; s_fz[K] = czf;
                    s_zp[K] = pr.zeta; s_zqp[K] = s_zqp[ZLUT_TAG];
                } else if (act == 3) {
                    if (accepted) s_zeta[pidx] = pr.zeta;
                    else { s_zp[pidx] = s_zeta[pidx]; s_zqp[pidx] = div1000_exact(s_zeta[pidx]); }
                } else if (act == 2 && !accepted) {
                    s_fx[pidx] = (float)s_nx[pidx];  // un-hide the nucleus that was proposed for deletion
                } else if (act == 4) {
                    if (accepted) { s_nx[pidx] = cx; s_ny[pidx] = cy; s_nz[pidx] = cz; }
                    else { s_fx[pidx] = (float)pr.ox; s_fy[pidx] = (float)pr.oy; s_fz[pidx] = (float)pr.oz; }
                }
            }
            if (accepted) {
                phi = phin;
                if (act == 1) K += 1;
                else if (act == 2) K -= 1;
                else if (act == 5) noise = pr.zeta;
            }
            __syncthreads();  // the committed state is consistent before the next proposal / the history copy
        }
        tick(4);
        // ================================================================ G: bookkeeping, traces, thinning (:275-281)
        int keep = 0;
        if ((double)iter >= pm.burn_in) {
            model_num += 1;
            if (fmod((double)model_num, pm.keep_each) == 0) keep = 1;
        }
        if (tid == 0) {
            if (act >= 1 && act <= 5) {
                s_cnt[act - 1] += 1u;
                if (accepted) s_cnt[5 + act - 1] += 1u;
                if (do_eval) s_cnt[10 + act - 1] += 1u;
            }
            if (a.tr_accept) a.tr_accept[(size_t)chain * a.trace_stride + a.it0 + it] = (int8_t)accepted;
            if (a.tr_phi) a.tr_phi[(size_t)chain * a.trace_stride + a.it0 + it] = phi;
            if (a.tr_K) a.tr_K[(size_t)chain * a.trace_stride + a.it0 + it] = K;
        }
        if (keep && beta == 1.0) {  // tempered replicas (beta < 1, extension) do not contribute to the posterior
            if (n_hist < a.hist_cap) {
                // Only the K valid nuclei of each axis are stored (the record keeps its fixed [4][KC] layout; the rest is never read).
                // The copy may overlap the next iteration's phases A..B: they modify neither the FP64 nuclei nor t*.
                const size_t h = (size_t)chain * a.hist_cap + n_hist;
                double *hc = a.hist_cells + h * 4 * KC;
                for (int i = tid; i < 4 * 32 && K <= 32; i += ST) {
                    const int ax = i >> 5, k = i & 31;
                    if (k < K) hc[ax * KC + k] = s_nx[ax * KC + k];
                }
                if (K > 32)
                    for (int ax = 0; ax < 4; ax++)
                        for (int k = tid; k < K; k += ST) hc[ax * KC + k] = s_nx[ax * KC + k];
                double *hp = a.hist_ptS + h * R;
                for (int i = tid; i < R; i += ST) hp[i] = s_tstar[a.ray_rank[i]];  // caller's ray order; contiguous stores (the history may live in mapped host memory)
                if (tid == 0) {
                    a.hist_K[h] = K; a.hist_phi[h] = phi; a.hist_iter[h] = iter;
                    a.hist_action[h] = act; a.hist_accept[h] = accepted; a.hist_next[h] = 0;
                }
                pending_slot = n_hist;
            }
            n_hist += 1;
        }
        tick(5);
    }

    if (PROF && tid == 0) for (int i = 0; i < 9; i++) a.prof[(size_t)chain * 16 + i] += pt[i];
    // ---- write the chain state back (TMA bulk stores for the arrays)
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
        bulk_store(a.owner + (size_t)chain * a.Ppad, s_owner, (uint32_t)a.Ppad);
        bulk_store(a.tstar + (size_t)chain * a.Rp, s_tstar, (uint32_t)(8 * a.Rp));
        bulk_store(a.cells + (size_t)chain * 4 * KC, s_nx, (uint32_t)(32 * KC));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        a.K[chain] = K; a.phi[chain] = phi; a.noise[chain] = noise;
        a.n_hist[chain] = n_hist; a.model_num[chain] = model_num; a.pending_slot[chain] = pending_slot;
        long long *c = a.counts + (size_t)chain * 15;
        for (int i = 0; i < 15; i++) c[i] += (long long)s_cnt[i];
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

}  // namespace tg
