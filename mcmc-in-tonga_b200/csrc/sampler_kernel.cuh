// sampler_kernel.cuh -- tg_sampler_kernel: the device-resident proposal loop (TD_inversion_function.jl:70-302).
// See sampler.cu for the overview.  Structure of one iteration (one CTA of 4 warps = one chain, state in shared memory):
//   A  warp 0: proposal (Philox or replay), a-priori checks, v_nearest at the new / killed nucleus, zlut
//   B  every warp owns the 128-point blocks b = warp, warp+4, ...:
//        B1 flat pass, 4 points per lane: birth/move compare d(p,new) with d(p,owner) (exact FP64, no FMA);
//           death/move flag the points owned by the killed/moved nucleus ("orphans") in the mask;
//        B2 the warp compacts its orphans into a small queue and rescans them 32 at a time, one orphan per lane,
//           so the K-long nucleus loop runs with full lanes and exists once in the code;
//   C  t* of the touched rays, one thread per (length-sorted) ray, left-to-right;
//   D  canonical phi;  E  alpha + accept (thread 0);  F  commit / roll back;  G  traces, thinning, history.
#pragma once
#include <type_traits>

#include "proposal.cuh"
#include "tonga_internal.cuh"

namespace tg {

struct SamplerArgs {
    // geometry (device order: rays sorted by length)
    const double *px, *py, *pz, *dtT, *tS, *sig;
    const float *pxf, *pyf, *pzf;   // fl32 copies (screening)
    float tol_alpha, tol_beta2;     // screening band
    int exact_only;                 // 1 = skip the FP32 screening (pure FP64 path; used by tests)
    long long *prof;                // PROF instantiation only: [n][16] per-phase clock64 totals of thread 0
    const int32_t *rayid, *ray_off, *ray_orig, *ray_rank;
    int R, Rp, KC, ldT;
    int P, Ppad;
    int n_sm;  // SM count (leader-warp rotation)
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

constexpr int SQ_CAP = 64;  // orphan queue entries per warp

struct SmemLayout {
    size_t o_owner, o_mask, o_tstar, o_tnew, o_dirty, o_nuc, o_nucf, o_zlut, o_scr, o_prop, o_queue, o_bar, total;
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
    L.o_nucf = take(4 * 3 * 128);  // fl32 nuclei, SoA with a fixed stride of 128; unused slots (and 0x7F = none) hold +inf
    L.o_zlut = take(8 * 128);
    L.o_scr = take(8 * 4);
    L.o_prop = take(sizeof(Prop));
    L.o_queue = take((size_t)(ST / 32) * SQ_CAP * (Ppad <= 65536 ? 2 : 4));
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

__device__ __forceinline__ void mark_dirty(uint32_t *dirty, const int32_t *__restrict__ rayid, int p) {
    const int r = rayid[p];
    atomicOr(&dirty[r >> 5], 1u << (r & 31));
}

// FP32 screening of the orphan rescan: best and second-best squared distance over the fl32 nuclei (4 per step; unused
// slots and a killed nucleus hold +inf); returns the winner, or -1 if best and second best are closer than the error band
// (this also covers a winner within the band of the 1e9 "no nucleus" threshold, d2 starts at 1e9) -> exact FP64 rescan.
__device__ __noinline__ int rescan_point_f32(const float *__restrict__ pxf, const float *__restrict__ pyf, const float *__restrict__ pzf,
                                             const float *sf, int K, int p, float tol_alpha, float tol_beta2) {
    const float x = pxf[p], y = pyf[p], z = pzf[p];
    float d1 = 1e9f, d2 = 1e9f;
    int i1 = TG_OWNER_NONE;
#pragma unroll 1
    for (int i = 0; i < K; i += 4) {
        const float4 fx = *reinterpret_cast<const float4 *>(sf + i), fy = *reinterpret_cast<const float4 *>(sf + 128 + i),
                     fz = *reinterpret_cast<const float4 *>(sf + 256 + i);
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
    const float tol = fmaf(tol_alpha, d1 + d2, tol_beta2);
    if (!(d2 - d1 > tol)) return -1;  // ambiguous (or NaN coordinates)
    return i1;
}

// Orphan rescan: nearest nucleus of flat point p among all K nuclei except `skip` (strict <, ascending index).
__device__ __noinline__ int rescan_point(const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
                                         const double *nx, const double *ny, const double *nz, int K, int skip, int p) {
    const double x = px[p], y = py[p], z = pz[p];
    double best = 1e9;
    int bi = TG_OWNER_NONE;
#pragma unroll 2
    for (int i = 0; i < K; i++) {
        const double d = dist2_exact(nx[i], ny[i], nz[i], x, y, z);
        if (d < best && i != skip) { best = d; bi = i; }
    }
    return bi;
}

template <typename QT, bool PROF>
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
    float *s_fx = reinterpret_cast<float *>(smem + L.o_nucf);
    float *s_fy = s_fx + 128, *s_fz = s_fx + 256;
    double *s_zlut = reinterpret_cast<double *>(smem + L.o_zlut);
    double *s_scr = reinterpret_cast<double *>(smem + L.o_scr);
    Prop *s_prop = reinterpret_cast<Prop *>(smem + L.o_prop);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.o_bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // The serial phases (proposal, touched-ray integration, accept, bookkeeping) run on ONE warp per chain.  Warp w of every CTA
    // lives on SM sub-partition w % 4, so the 7 chains of an SM would all queue their serial work on sub-partition 0:
    // rotate the leader warp with the CTA's slot on its SM (CTAs b, b+nSM, b+2nSM, ... share an SM).
    const int lead = 0;  // (rotating the leader with blockIdx.x / n_sm was measured: phase A shrinks, phase B grows by as much)
    const int vtid = (tid - lead * 32) & (ST - 1), vwarp = vtid >> 5;  // virtual ids: the leader is virtual warp 0
    QT *s_queue = reinterpret_cast<QT *>(smem + L.o_queue) + warp * SQ_CAP;
    const int chain = a.perm ? a.perm[blockIdx.x] : (int)blockIdx.x;  // launch order = cost order (tg_order_kernel)
    const int KC = a.KC, R = a.R;
    float *__restrict__ dcache = a.dcache + (size_t)chain * a.Ppad;
    const int nOwnWords = a.Ppad / 4, nMaskWords = a.Ppad / 32, nDirtyWords = (a.Rp + 31) / 32;
    const int nBlocks = a.Ppad / 128;

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
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            s2u(s_bar))
        : "memory");
    __syncthreads();
    {   // fl32 copies of the nuclei; slots >= K hold +inf so that they never win a comparison
        const int K0 = a.K[chain];
        for (int i = tid; i < 3 * 128; i += ST) {
            const int ax = i >> 7, k = i & 127;
            s_fx[i] = (k < K0) ? (float)s_nx[ax * KC + k] : __int_as_float(0x7f800000);
        }
    }
    __syncthreads();

    int K = a.K[chain];
    double phi = a.phi[chain];
    double noise = a.noise[chain];
    const double beta = a.beta[chain];
    int n_hist = a.n_hist[chain];
    long long model_num = a.model_num[chain];
    int pending_slot = a.pending_slot[chain];

    const tonga_params &pm = a.prm;
    const double sig_zeta = pm.zeta_scale * pm.sig / 100;  // TD_inversion_function.jl:22
    const unsigned long long gid = (unsigned long long)(a.chain_id0 + chain);

    long long pt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = PROF ? clock64() : 0;
    auto tick = [&](int ph) {
        if (PROF && tid == 0) { const long long t = clock64(); pt[ph] += t - tprev; tprev = t; }
    };
#pragma unroll 1
    for (long long it = 0; it < a.nIter; it++) {
        const long long iter = a.iter0 + it;
        // ================================================================ A: proposal (warp 0, warp-uniform values)
        if (vwarp == 0) {
            Prop pr;
            const int valid = draw_proposal<0>(pr, a.mode, a.mode == 1 ? a.recs_in + (size_t)chain * a.nIter + it : nullptr, a.seed, iter, gid, s_nx, s_ny,
                                            s_nz, s_zeta, K, noise, pm, sig_zeta, lane);
            const int act = pr.action;
            if (valid && act == 2 && lane == 0) s_fx[pr.idx] = __int_as_float(0x7f800000);  // fl32 screening must not see the killed nucleus
            if (valid && act == 4) {  // from here to the accept decision the nucleus array holds the PROPOSED position
                __syncwarp();
                if (lane == 0) {
                    s_nx[pr.idx] = pr.x; s_ny[pr.idx] = pr.y; s_nz[pr.idx] = pr.z;
                    s_fx[pr.idx] = (float)pr.x; s_fy[pr.idx] = (float)pr.y; s_fz[pr.idx] = (float)pr.z;
                }
            }
            // zlut: owner byte -> zeta under the proposed model (bit 7 set = switches to the implicit new owner)
            if (valid && act != 5) {
                pr.ztag = (act == 1) ? pr.zeta : (act == 4 ? s_zeta[pr.idx] : 0.0);
                for (int o = lane; o < 128; o += 32) {
                    double zv = (o < K) ? s_zeta[o] : 0.0;
                    if (act == 3 && o == pr.idx) zv = pr.zeta;
                    s_zlut[o] = zv;
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
        tick(0);
        pending_slot = -1;
        const int act = s_prop->action;
        const int do_eval = s_prop->do_eval;
        const int pidx = s_prop->idx;
        const uint32_t kk = (uint32_t)pidx * 0x01010101u;
        double phin = phi;
        int accepted = 0;

        if (do_eval) {
            if (act != 5) {
                const double cx = s_prop->x, cy = s_prop->y, cz = s_prop->z;
                const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
                // ======================================================== B1: flat pass over this warp's 128-point blocks
                const float ta = a.tol_alpha, tb = a.tol_beta2;
                unsigned long long blkmask = 0ull;  // warp-iterations whose block holds orphans (bit it; it >= 64 -> always scanned)
                // birth / move: FP32 screening of d(p,new) against d(p,owner); ACT is a compile-time constant so that the birth
                // path carries no move logic.  Branch-free per point; near ties (inside the error band) go to the exact FP64 compare.
                auto b1_switch = [&](auto actc) {
                    constexpr int ACT = decltype(actc)::value;
                    const int mv = (ACT == 4) ? pidx : -1;
                    const float2 ncx = make_float2(-cxf, -cxf), ncy = make_float2(-cyf, -cyf), ncz = make_float2(-czf, -czf);
                    int it = 0;
                    // software pipeline: the next block's coordinates are in flight while the current block is screened
                    float4 nxf = *reinterpret_cast<const float4 *>(a.pxf + 4 * (warp * 32 + lane));
                    float4 nyf = *reinterpret_cast<const float4 *>(a.pyf + 4 * (warp * 32 + lane));
                    float4 nzf = *reinterpret_cast<const float4 *>(a.pzf + 4 * (warp * 32 + lane));
                    int4 nrid = *reinterpret_cast<const int4 *>(a.rayid + 4 * (warp * 32 + lane));
                    float4 ndo = *reinterpret_cast<const float4 *>(dcache + 4 * (warp * 32 + lane));
#pragma unroll 1
                    for (int blk = warp; blk < nBlocks; blk += ST / 32, it++) {
                        const int w = blk * 32 + lane;
                        const uint32_t ow = s_own32[w];
                        uint32_t tags = 0, amb = 0, mbits = 0;
                        const float4 xf = nxf, yf = nyf, zf = nzf;
                        const int4 rid = nrid;
                        const float4 dof = ndo;
                        if (blk + ST / 32 < nBlocks) {
                            const int wn = w + (ST / 32) * 32;
                            nxf = *reinterpret_cast<const float4 *>(a.pxf + 4 * wn);
                            nyf = *reinterpret_cast<const float4 *>(a.pyf + 4 * wn);
                            nzf = *reinterpret_cast<const float4 *>(a.pzf + 4 * wn);
                            nrid = *reinterpret_cast<const int4 *>(a.rayid + 4 * wn);
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
                                const int o = (ow >> (8 * q)) & 0xFF;  // no tags are pending here
                                const float d_o = DO[q];
                                const float diff = DC[q] - d_o;
                                const float tol = fmaf(ta, DC[q] + d_o, tb);
                                bool sw = diff < -tol, am = fabsf(diff) <= tol;
                                if (ACT == 4) {
                                    const bool mine = (o == mv);  // move, type A: owned by the moved nucleus -> rescan in B2
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
                        if (tags) {
                            s_own32[w] = ow | tags;
                            const int RID[4] = {rid.x, rid.y, rid.z, rid.w};
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if ((tags >> (8 * q + 7)) & 1u) atomicOr(&s_dirty[RID[q] >> 5], 1u << (RID[q] & 31));
                        }
                        if (ACT == 4 && __any_sync(0xffffffffu, mbits != 0u)) {  // assemble the block's 4 mask words (8 lanes x 4 bits each)
                            uint32_t nib = mbits << ((lane & 7) * 4);
                            nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
                            nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
                            nib |= __shfl_xor_sync(0xffffffffu, nib, 4);
                            if ((lane & 7) == 0) s_mask[w >> 3] = nib;
                            if (it < 64) blkmask |= 1ull << it;
                        }
                    }
                };
                if (act == 1) b1_switch(std::integral_constant<int, 1>{});
                else if (act == 4) b1_switch(std::integral_constant<int, 4>{});
                else {  // death: flag the orphans; change: mark the rays through the cell
                    int it = 0;
#pragma unroll 1
                    for (int blk = warp; blk < nBlocks; blk += ST / 32, it++) {
                        const int w = blk * 32 + lane;
                        const uint32_t eq = __vcmpeq4(s_own32[w], kk);  // bytes owned by the killed / changed nucleus
                        if (__any_sync(0xffffffffu, eq != 0u)) {
                            if (act == 2) {
                                const uint32_t mbits = (eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u);
                                uint32_t nib = mbits << ((lane & 7) * 4);
                                nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
                                nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
                                nib |= __shfl_xor_sync(0xffffffffu, nib, 4);
                                if ((lane & 7) == 0) s_mask[w >> 3] = nib;
                                if (it < 64) blkmask |= 1ull << it;
                            } else if (eq) {
                                const int4 rid = *reinterpret_cast<const int4 *>(a.rayid + 4 * w);
                                const int RID[4] = {rid.x, rid.y, rid.z, rid.w};
#pragma unroll
                                for (int q = 0; q < 4; q++)
                                    if (((eq >> (8 * q)) & 1u) && (q == 0 || RID[q] != RID[q - 1] || !((eq >> (8 * q - 8)) & 1u)))
                                        atomicOr(&s_dirty[RID[q] >> 5], 1u << (RID[q] & 31));
                            }
                        }
                    }
                }
                tick(8);  // profile slot 8 = B1, slot 1 = B2 (+ barrier)
                // ======================================================== B2: rescan this warp's orphans, 32 at a time
                if (act == 2 || act == 4) {
                    __syncwarp();
                    const int skip = (act == 2) ? pidx : -1;
                    int qn = 0;
                    auto drain = [&](int cnt) {  // lanes < cnt take the top `cnt` queue entries
                        if (lane < cnt) {
                            const int p = (int)s_queue[qn - cnt + lane];
                            const int rid = a.rayid[p];  // in flight during the rescan
                            int bi = a.exact_only ? -1 : rescan_point_f32(a.pxf, a.pyf, a.pzf, s_fx, K, p, ta, tb);
                            if (bi < 0) bi = rescan_point(a.px, a.py, a.pz, s_nx, s_ny, s_nz, K, skip, p);  // exact FP64
                            s_owner[p] = (uint8_t)bi;  // death: old numbering, renumbered on accept
                            if (act == 2 || bi != pidx) atomicOr(&s_dirty[rid >> 5], 1u << (rid & 31));
                        }
                        qn -= cnt;
                        __syncwarp();
                    };
                    int it = 0;
#pragma unroll 1
                    for (int blk = warp; blk < nBlocks; blk += ST / 32, it++) {
                        if (it < 64 && !((blkmask >> it) & 1ull)) continue;
#pragma unroll 1
                        for (int k4 = 0; k4 < 4; k4++) {
                            const int mw = blk * 4 + k4;
                            const uint32_t bits = s_mask[mw];
                            if (bits) {
                                if ((bits >> lane) & 1u) s_queue[qn + __popc(bits & ((1u << lane) - 1u))] = (QT)(mw * 32 + lane);
                                qn += __popc(bits);
                                __syncwarp();
                                if (qn >= 32) drain(32);
                            }
                        }
                    }
                    if (qn > 0) drain(qn);
                }
                __syncthreads();
                tick(1);
                // ======================================================== C: t* of touched rays, one thread per sorted ray
                const double ztag = s_prop->ztag;
                {   // The touched rays are COMPACTED (they are ~20% of all rays): entry e of the compacted, still length-sorted list
                    // goes to thread e, so every warp walks 32 rays of similar length with all lanes busy.  No list is stored:
                    // thread e finds its ray as the e-th set bit of the dirty bitmap (popc prefix + __fns).
                    const int nwords = (R + 31) >> 5;
                    int nd = 0;
                    for (int i = 0; i < nwords; i++) nd += __popc(s_dirty[i]);
#pragma unroll 1
                    for (int e = vtid; e < nd; e += ST) {
                        int cum = 0, wi = 0;
                        uint32_t bits = s_dirty[0];
                        while (cum + __popc(bits) <= e) { cum += __popc(bits); bits = s_dirty[++wi]; }
                        const int r = wi * 32 + (int)__fns(bits, 0, e - cum + 1);
                        const int q0 = a.ray_off[r], n = a.ray_off[r + 1] - q0;
                        s_tnew[r] = ray_tstar_seq<uint8_t>(s_owner, a.dtT, a.ldT, r, q0, n,
                                                           [&](uint8_t o) -> double { return (o & 0x80) ? ztag : s_zlut[o]; });
                    }
                }
                __syncthreads();
                tick(2);
            }
            // ============================================================ D: phi of the proposed model (canonical order)
            const double nz = (act == 5) ? s_prop->zeta : noise;
            phin = phi_canonical_128(R, tid, s_scr, [&](int r) {
                const double t = ((s_dirty[r >> 5] >> (r & 31)) & 1u) ? s_tnew[r] : s_tstar[r];
                return misfit_term(t, a.tS[r], a.sig[r], nz);
            });
            if (pm.debug_prior) phin = 1.0;  // MCsub.jl:128-136: the chain samples the prior
            // ============================================================ E: acceptance (one thread of the leader warp)
            if (vtid == 0)
                s_prop->accept = accept_decision(*s_prop, K, phi, phin, (act == 2 || act == 3) ? s_zeta[pidx] : 0.0, noise, beta, R, pm, sig_zeta);
            __syncthreads();
            tick(3);
            accepted = s_prop->accept;
            // ============================================================ F: commit / roll back
            if (act == 2 || act == 4) {  // masked bytes: overwritten orphans (old owner = pidx)
#pragma unroll 1
                for (int i = tid; i < nMaskWords; i += ST) {
                    uint32_t b = s_mask[i];
                    if (b) {
                        s_mask[i] = 0u;
                        while (b) {
                            const int j = __ffs(b) - 1;
                            b &= b - 1;
                            const int p = 32 * i + j;
                            if (!accepted) {
                                s_owner[p] = (uint8_t)pidx;
                            } else {  // refresh the owner-distance cache (a move also changes it for points that stay with the nucleus)
                                const int o = s_owner[p] & 0x7F;  // death: still the old numbering, as are the fl32 nuclei
                                dcache[p] = (o == TG_OWNER_NONE) ? 1e9f : dist2_f32(s_fx[o], s_fy[o], s_fz[o], a.pxf[p], a.pyf[p], a.pzf[p]);
                            }
                        }
                    }
                }
                __syncthreads();  // move: byte stores above vs word updates below; death: the cache refresh above reads the nuclei the delete below shifts
            }
            tick(6);
            if (act == 1 || act == 4) {  // tagged bytes: switch to the new / moved nucleus
                const uint32_t newb = (uint32_t)(act == 1 ? K : pidx) * 0x01010101u;
#pragma unroll 1
                for (int w = tid; w < nOwnWords; w += ST) {
                    const uint32_t ow = s_own32[w], t = ow & 0x80808080u;
                    if (t) {
                        const uint32_t m = (t >> 7) * 0xFFu;
                        s_own32[w] = accepted ? ((ow & ~m) | (newb & m)) : (ow & 0x7F7F7F7Fu);
                        if (accepted) {  // switched points: cache their distance to the new nucleus
                            const float fx = (float)s_prop->x, fy = (float)s_prop->y, fz = (float)s_prop->z;
#pragma unroll 1
                            for (int q = 0; q < 4; q++)
                                if ((t >> (8 * q + 7)) & 1u) {
                                    const int p = 4 * w + q;
                                    dcache[p] = dist2_f32(fx, fy, fz, a.pxf[p], a.pyf[p], a.pzf[p]);
                                }
                        }
                    }
                }
            }
            tick(7);
            if (act == 2 && accepted) {  // deleteat! renumbering: indices above `kill` shift down (:132-135)
#pragma unroll 1
                for (int w = tid; w < nOwnWords; w += ST) {
                    const uint32_t ow = s_own32[w];
                    const uint32_t gt = __vcmpgtu4(ow, kk) & ~__vcmpeq4(ow, 0x7F7F7F7Fu);
                    if (gt) s_own32[w] = ow - (gt & 0x01010101u);
                }
                if (vwarp == 0) {  // order-preserving delete of the nucleus
                    double vx[4], vy[4], vz[4], vt[4];
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        const int i = pidx + lane + 32 * s;
                        if (i < K - 1) { vx[s] = s_nx[i + 1]; vy[s] = s_ny[i + 1]; vz[s] = s_nz[i + 1]; vt[s] = s_zeta[i + 1]; }
                    }
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        const int i = pidx + lane + 32 * s;
                        if (i < K - 1) {
                            s_nx[i] = vx[s]; s_ny[i] = vy[s]; s_nz[i] = vz[s]; s_zeta[i] = vt[s];
                            s_fx[i] = (float)vx[s]; s_fy[i] = (float)vy[s]; s_fz[i] = (float)vz[s];
                        }
                    }
                    __syncwarp();
                    if (lane == 0) { s_fx[K - 1] = s_fy[K - 1] = s_fz[K - 1] = __int_as_float(0x7f800000); }  // freed slot -> +inf
                }
            }
            if (vtid == 0) {
                if (act == 1 && accepted) {  // append!, :85-88
                    s_nx[K] = s_prop->x; s_ny[K] = s_prop->y; s_nz[K] = s_prop->z; s_zeta[K] = s_prop->zeta;
                    s_fx[K] = (float)s_prop->x; s_fy[K] = (float)s_prop->y; s_fz[K] = (float)s_prop->z;
                } else if (act == 3 && accepted) {
                    s_zeta[pidx] = s_prop->zeta;
                } else if (act == 2 && !accepted) {
                    s_fx[pidx] = (float)s_nx[pidx];  // un-hide the nucleus that was proposed for deletion
                } else if (act == 4 && !accepted) {
                    s_nx[pidx] = s_prop->ox; s_ny[pidx] = s_prop->oy; s_nz[pidx] = s_prop->oz;
                    s_fx[pidx] = (float)s_prop->ox; s_fy[pidx] = (float)s_prop->oy; s_fz[pidx] = (float)s_prop->oz;
                }
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
        tick(4);
        // ================================================================ G: bookkeeping, traces, thinning (:275-281)
        int keep = 0;
        if ((double)iter >= pm.burn_in) {
            model_num += 1;
            if (fmod((double)model_num, pm.keep_each) == 0) keep = 1;
        }
        if (vtid == 0) {
            if (act >= 1 && act <= 5) {
                unsigned long long *c = reinterpret_cast<unsigned long long *>(a.counts) + (size_t)chain * 15 + (act - 1);
                atomicAdd(c, 1ULL);  // RED: fire and forget
                if (accepted) atomicAdd(c + 5, 1ULL);
                if (do_eval) atomicAdd(c + 10, 1ULL);
            }
            if (a.tr_accept) a.tr_accept[(size_t)chain * a.nIter + it] = (int8_t)accepted;
            if (a.tr_phi) a.tr_phi[(size_t)chain * a.nIter + it] = phi;
            if (a.tr_K) a.tr_K[(size_t)chain * a.nIter + it] = K;
        }
        __syncthreads();  // state (nuclei, tstar, owners) consistent before the next proposal / the history copy
        if (keep && beta == 1.0) {  // tempered replicas (beta < 1, extension) do not contribute to the posterior
            if (n_hist < a.hist_cap) {
                const size_t h = (size_t)chain * a.hist_cap + n_hist;
                double *hc = a.hist_cells + h * 4 * KC;
                for (int i = tid; i < 4 * KC; i += ST) hc[i] = s_nx[i];
                double *hp = a.hist_ptS + h * R;
                for (int i = tid; i < R; i += ST) hp[i] = s_tstar[a.ray_rank[i]];  // caller's ray order; contiguous stores (the history may live in mapped host memory)
                if (vtid == 0) {
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
        bulk_store(a.owner + (size_t)chain * a.Ppad, s_owner, (uint32_t)a.Ppad);
        bulk_store(a.tstar + (size_t)chain * a.Rp, s_tstar, (uint32_t)(8 * a.Rp));
        bulk_store(a.cells + (size_t)chain * 4 * KC, s_nx, (uint32_t)(32 * KC));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        a.K[chain] = K; a.phi[chain] = phi; a.noise[chain] = noise;
        a.n_hist[chain] = n_hist; a.model_num[chain] = model_num; a.pending_slot[chain] = pending_slot;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

}  // namespace tg
