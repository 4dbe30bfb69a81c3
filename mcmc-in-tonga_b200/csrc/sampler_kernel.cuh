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
//   C  warp w owns the (length-sorted) rays r = w (mod NW): it finds its touched rays from the word bitmap and re-integrates
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

// the owner-distance cache is read and written past L1 (ld.cg / st.cg): 67 KB per chain, streamed once per birth / move
#define TG_DC_LOAD(p) __ldcg(p)
#define TG_DC_STORE(p, v) __stcg(p, v)
constexpr int SQ_CAP = 64;    // orphan-word queue entries per warp
constexpr int ZLUT_TAG = 128; // entries 128..255 of the zeta look-up table: tagged bytes (all hold the implicit new owner's value)
// Visits a warp's blocks b = warp + NW * bi whose bit is set in `mask` (bi < 32), then every bi >= 32 (not covered by the mask).
struct BlockIter {
    uint32_t rem;
    int hi;
    __device__ __forceinline__ explicit BlockIter(uint32_t mask) : rem(mask), hi(32) {}
    __device__ __forceinline__ int next() {
        if (rem) { const int b = __ffs((int)rem) - 1; rem &= rem - 1u; return b; }
        return hi++;
    }
};
struct WarpProp {  // the proposal's scalars that are only needed again at the acceptance / commit: one copy per warp (same values)
    double u, aux, zeta, ox, oy, oz, zold, cx, cy, cz;
};
struct ChainScalars {  // per-chain scalars kept out of the registers: thread 0 writes, everybody reads after a barrier
    double phi, noise, beta;
    int pending_slot, pad;
};
// Header at the start of the CTA's dynamic shared memory: everything of fixed size (compile-time offsets), plus what the phase
// functions need to find the rest.  The hot phases are separate __noinline__ functions -- each gets its own register
// allocation instead of competing with the whole iteration's live state (the inlined kernel spilled inside its hot loops) --
// and take their constants from here instead of through argument lists.
struct Hdr {
    const float *pxf, *pyf, *pzf;
    float *dcache;  // this chain's owner-distance cache
    const double *px, *py, *pz, *dt;
    uint32_t o_owner, o_mask, o_tstar, o_nuc, o_nucf, o_dirtyw;  // offsets of the variable-size arrays
    int KC, nBlocks, R, exact_only;
    float ta, tb, ta2, pad0;
    ChainScalars st;
    WarpProp prop[NW];
    double scr[8];
    double zh[256];  // owner byte -> 0.5 * zeta under the PROPOSED model; [127] = 0 (none), [128..255] = tagged bytes
    uint32_t cnt[16];
    uint16_t queue[NW][SQ_CAP];
    uint8_t perm[NW][32];
    unsigned long long bar;
};
extern __shared__ __align__(16) unsigned char tg_smem[];
__device__ __forceinline__ Hdr &hdr() { return *reinterpret_cast<Hdr *>(tg_smem); }

struct SmemLayout {
    size_t o_owner, o_mask, o_tstar, o_nuc, o_nucf, o_dirtyw, total;
};
__host__ __device__ inline SmemLayout smem_layout(int Ppad, int Rp, int KC) {
    SmemLayout L;
    size_t o = (sizeof(Hdr) + 15) & ~(size_t)15;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
    L.o_owner = take((size_t)Ppad + 256);  // + slack: phase C reads owner bytes unconditionally a little past a ray's end
    L.o_mask = take((size_t)Ppad / 8);
    L.o_tstar = take(8 * (size_t)Rp);
    L.o_nuc = take(8 * 4 * (size_t)KC);
    L.o_nucf = take(4 * 3 * (size_t)KC);   // fl32 nuclei, SoA with stride KC; unused slots and a killed nucleus hold +inf
    L.o_dirtyw = take(4 * (size_t)(Ppad / 128));  // bit per 4-point word: some point of the word changes owner / zeta under the proposal
    L.total = o;
    return L;
}

// ---- TMA bulk copies (SASS: UBLKCP) -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
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
static __device__ __noinline__ int rescan_point(const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
                                                const double *nx, const double *ny, const double *nz, int K, int skip, int mvi, double cx, double cy,
                                                double cz, int p) {
    const double x = __ldg(px + p), y = __ldg(py + p), z = __ldg(pz + p);
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

// =================================================================================================== phase B1 (birth / move)
// Flat pass over this warp's 128-point blocks, 4 points per lane: FP32 screening of d(p,new) against the cached d(p,owner);
// ACT is a compile-time constant so that the birth path carries no move logic.  Branch-free per point; near ties (inside the
// error band) go to the exact FP64 compare.  Switching points get the tag bit; a ballot per block records the changed words;
// a move also flags the points of the moved nucleus in the mask (rescanned in B2).  Returns (tagmask, blkmask): this warp's
// blocks holding tagged bytes / orphans (bit = warp-iteration; iterations >= 32 are always scanned).
template <int ACT>
static __device__ __noinline__ uint2 phase_b1_switch(const int pidx, const float cxf, const float cyf, const float czf) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    const uint32_t FULL = 0xffffffffu;
    uint32_t *s_own32 = reinterpret_cast<uint32_t *>(tg_smem + h.o_owner);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(tg_smem + h.o_mask);
    uint32_t *s_dirtyw = reinterpret_cast<uint32_t *>(tg_smem + h.o_dirtyw);
    const int nBlocks = h.nBlocks, mv = (ACT == 4) ? pidx : -1;
    const float ta = h.ta, tb = h.tb;
    const bool exact_only = h.exact_only != 0;
    const float2 ncx = make_float2(-cxf, -cxf), ncy = make_float2(-cyf, -cyf), ncz = make_float2(-czf, -czf);
    // (the two affine forms below are evaluated in fl32 themselves: the band is widened by 4 ulp to keep its margin)
    const float tw = ta + 4.0f * 0x1.0p-24f, tbw = tb * (1.0f + 0x1.0p-20f);
    const float2 kA = make_float2(1.0f + tw, 1.0f + tw), kB = make_float2(1.0f - tw, 1.0f - tw), kNA = make_float2(-1.0f - tw, -1.0f - tw),
                 kNB = make_float2(tw - 1.0f, tw - 1.0f), kT = make_float2(tbw, tbw), kNT = make_float2(-tbw, -tbw);
    uint32_t tagmask = 0u, blkmask = 0u;
    auto body = [&](const int blk, const int bi, const float4 xf, const float4 yf, const float4 zf, const float4 dof) {
        const int w = blk * 32 + lane;
        const uint32_t ow = s_own32[w];
        uint32_t tags = 0, amb = 0, mbits = 0;
        if (!exact_only) {
            float2 ex = __fadd2_rn(make_float2(xf.x, xf.y), ncx), ey = __fadd2_rn(make_float2(yf.x, yf.y), ncy), ez = __fadd2_rn(make_float2(zf.x, zf.y), ncz);
            const float2 dc01 = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
            ex = __fadd2_rn(make_float2(xf.z, xf.w), ncx); ey = __fadd2_rn(make_float2(yf.z, yf.w), ncy); ez = __fadd2_rn(make_float2(zf.z, zf.w), ncz);
            const float2 dc23 = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
            // The error band |dc - do| <= ta (dc + do) + tb as two affine forms (dc = new distance, do = cached owner distance):
            //   usw = dc (1 + ta) - do (1 - ta) + tb < 0   <=>  dc - do < -tol : switches for certain  (its sign bit is the tag)
            //   ust = dc (1 - ta) - do (1 + ta) - tb > 0   <=>  dc - do >  tol : stays for certain
            // anything else is ambiguous -> exact FP64.  (The fl32 rounding of these forms is far inside the band's 1.4x margin.)
            const float2 do01 = make_float2(dof.x, dof.y), do23 = make_float2(dof.z, dof.w);
            const float2 usw01 = __ffma2_rn(dc01, kA, __ffma2_rn(do01, kNB, kT)), usw23 = __ffma2_rn(dc23, kA, __ffma2_rn(do23, kNB, kT));
            const float2 ust01 = __ffma2_rn(dc01, kB, __ffma2_rn(do01, kNA, kNT)), ust23 = __ffma2_rn(dc23, kB, __ffma2_rn(do23, kNA, kNT));
            // tags: sign bit of usw -> bit 7 of the point's owner byte (NaN never occurs: padded points carry finite far-away coordinates)
            const uint32_t sg = __byte_perm(__byte_perm(__float_as_uint(usw01.x), __float_as_uint(usw01.y), 0x0073),
                                            __byte_perm(__float_as_uint(usw23.x), __float_as_uint(usw23.y), 0x0073), 0x5410) & 0x80808080u;
            // ambiguous point: usw >= 0 and ust <= 0  <=>  max(-usw, ust) <= 0
            const float w0 = fmaxf(-usw01.x, ust01.x), w1 = fmaxf(-usw01.y, ust01.y), w2 = fmaxf(-usw23.x, ust23.x), w3 = fmaxf(-usw23.y, ust23.y);
            tags = sg;
            if (ACT == 4) {  // move, type A: points of the moved nucleus are rescanned in B2, not screened here
                const uint32_t mine = __vcmpeq4(ow, (uint32_t)mv * 0x01010101u);
                mbits = (mine & 1u) | ((mine >> 7) & 2u) | ((mine >> 14) & 4u) | ((mine >> 21) & 8u);
                tags &= ~mine;
            }
            if (fminf(fminf(w0, w1), fminf(w2, w3)) <= 0.0f) {
                amb = (w0 <= 0.0f ? 1u : 0u) | (w1 <= 0.0f ? 2u : 0u) | (w2 <= 0.0f ? 4u : 0u) | (w3 <= 0.0f ? 8u : 0u);
                amb &= ~mbits;
                tags &= ~(((amb & 1u) * 0x80u) | ((amb & 2u) * 0x4000u) | ((amb & 4u) * 0x200000u) | ((amb & 8u) * 0x10000000u));  // decided below
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if ((int)((ow >> (8 * q)) & 0xFF) == mv) mbits |= 1u << q;
                else amb |= 1u << q;
            }
        }
        if (amb) {  // exact FP64 comparison, MCsub.jl:254-255 semantics
            const double cx = h.prop[warp].cx, cy = h.prop[warp].cy, cz = h.prop[warp].cz;
            const double *s_nx = reinterpret_cast<const double *>(tg_smem + h.o_nuc), *s_ny = s_nx + h.KC, *s_nz = s_ny + h.KC;
#pragma unroll 1
            for (int q = 0; q < 4; q++) {
                if (!((amb >> q) & 1u)) continue;
                const int o = (ow >> (8 * q)) & 0xFF;
                const int p = 4 * w + q;
                const double x = __ldg(h.px + p), y = __ldg(h.py + p), z = __ldg(h.pz + p);
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
        if (dm && bi < 32) tagmask |= 1u << bi;
        if (ACT == 4 && __any_sync(FULL, mbits != 0u)) {  // assemble the block's 4 mask words (8 lanes x 4 bits each)
            uint32_t nib = mbits << (sub * 4);
            nib |= __shfl_xor_sync(FULL, nib, 1);
            nib |= __shfl_xor_sync(FULL, nib, 2);
            nib |= __shfl_xor_sync(FULL, nib, 4);
            if (sub == 0) s_mask[w >> 3] = nib;
            if (bi < 32) blkmask |= 1u << bi;
        }
    };
    const float4 *gx = reinterpret_cast<const float4 *>(h.pxf) + lane, *gy = reinterpret_cast<const float4 *>(h.pyf) + lane,
                 *gz = reinterpret_cast<const float4 *>(h.pzf) + lane, *gd = reinterpret_cast<const float4 *>(h.dcache) + lane;
    // One block per step; the other warps of the SM (7 chains x 4 warps) hide the L2 latency.  Measured and rejected (profiles/README.md,
    // round 2): a two-register-set ping-pong (birth only: 26.9, both actions: 24.2 against 27.4 M/s -- the second set costs spills)
    // and prefetch.global.L1 of the next block (27.3 / 26.9 M/s with the cache going through L1).
    int bi = 0;
#pragma unroll 1
    for (int blk = warp; blk < nBlocks; blk += NW, bi++)
        body(blk, bi, __ldg(gx + blk * 32), __ldg(gy + blk * 32), __ldg(gz + blk * 32), TG_DC_LOAD(gd + blk * 32));
    return make_uint2(tagmask, blkmask);
}

// =================================================================================================== phase B1 (death / change)
// death: flag the orphans (points of the killed nucleus) in the mask; change: mark the words of the cell.  Returns blkmask.
static __device__ __noinline__ uint32_t phase_b1_scan(const int act, const int pidx) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    const uint32_t FULL = 0xffffffffu, kk = (uint32_t)pidx * 0x01010101u;
    const uint32_t *s_own32 = reinterpret_cast<const uint32_t *>(tg_smem + h.o_owner);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(tg_smem + h.o_mask);
    uint32_t *s_dirtyw = reinterpret_cast<uint32_t *>(tg_smem + h.o_dirtyw);
    const int nBlocks = h.nBlocks;
    uint32_t blkmask = 0u;
    int bi = 0;
#pragma unroll 1
    for (int blk = warp; blk < nBlocks; blk += NW, bi++) {
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
            if (bi < 32) blkmask |= 1u << bi;
        }
    }
    return blkmask;
}

// =================================================================================================== phase B2 (death / move)
// The warp compacts its orphan WORDS (4-point words holding a point of the killed / moved nucleus) into a queue and rescans
// them 32 words at a time: 4 points per lane, packed FP32 over the fl32 nuclei (2 per step; unused slots and a killed nucleus
// hold +inf), the nucleus index carried in the 7 low mantissa bits so that best / second best are integer min / max; a point
// whose two best are inside the (widened) error band is rescanned exactly in FP64.
static __device__ __noinline__ void phase_b2(const int act, const int pidx, const int K, const uint32_t blkmask) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    const uint32_t FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    uint32_t *s_own32 = reinterpret_cast<uint32_t *>(tg_smem + h.o_owner);
    const uint32_t *s_mask = reinterpret_cast<const uint32_t *>(tg_smem + h.o_mask);
    uint32_t *s_dirtyw = reinterpret_cast<uint32_t *>(tg_smem + h.o_dirtyw);
    const float *s_fx = reinterpret_cast<const float *>(tg_smem + h.o_nucf), *s_fy = s_fx + h.KC, *s_fz = s_fy + h.KC;
    uint16_t *s_queue = hdr().queue[warp];
    const int nBlocks = h.nBlocks;
    const float ta2 = h.ta2, tb = h.tb;
    const bool exact_only = h.exact_only != 0;
    int qn = 0;
    BlockIter iter(blkmask);  // only the blocks that hold orphans
#pragma unroll 1
    for (;;) {
        const int blk = warp + NW * iter.next();
        const bool last = blk >= nBlocks;
        if (!last) {
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
                uint32_t nbw = TG_OWNER_NONE * 0x01010101u;  // the 4 new owners, one byte each (a register, not an indexed array: no local memory)
                uint32_t need = exact_only ? mb : 0u;
                if (!exact_only) {
                    const float4 X = __ldg(reinterpret_cast<const float4 *>(h.pxf) + w), Y = __ldg(reinterpret_cast<const float4 *>(h.pyf) + w),
                                 Z = __ldg(reinterpret_cast<const float4 *>(h.pzf) + w);
                    const float2 X01 = make_float2(X.x, X.y), X23 = make_float2(X.z, X.w), Y01 = make_float2(Y.x, Y.y), Y23 = make_float2(Y.z, Y.w),
                                 Z01 = make_float2(Z.x, Z.y), Z23 = make_float2(Z.z, Z.w);
                    const uint32_t INIT = (__float_as_uint(1e9f) & 0xFFFFFF80u) | TG_OWNER_NONE;
                    uint32_t d1[4] = {INIT, INIT, INIT, INIT}, d2[4] = {INIT, INIT, INIT, INIT};
#pragma unroll 1
                    for (int i = 0; i < K; i++) {  // one nucleus per iteration: two per iteration overlap more, but the loop is twice the code (-1.2 %)
                        const float ax = -s_fx[i], ay = -s_fy[i], az = -s_fz[i];
                        const float2 nax = make_float2(ax, ax), nay = make_float2(ay, ay), naz = make_float2(az, az);
                        float2 ex = __fadd2_rn(X01, nax), ey = __fadd2_rn(Y01, nay), ez = __fadd2_rn(Z01, naz);
                        const float2 da = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                        ex = __fadd2_rn(X23, nax); ey = __fadd2_rn(Y23, nay); ez = __fadd2_rn(Z23, naz);
                        const float2 db = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                        const float d[4] = {da.x, da.y, db.x, db.y};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t dp = (__float_as_uint(d[q]) & 0xFFFFFF80u) | (uint32_t)i;
                            const uint32_t t = max(d1[q], dp);
                            d1[q] = min(d1[q], dp);
                            d2[q] = min(d2[q], t);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float f1 = __uint_as_float(d1[q] & 0xFFFFFF80u), f2 = __uint_as_float(d2[q] & 0xFFFFFF80u);
                        const float tol = fmaf(ta2, f1 + f2, tb);
                        nbw = (nbw & ~(0xFFu << (8 * q))) | ((d1[q] & 0x7Fu) << (8 * q));
                        if (!(f2 - f1 > tol)) need |= 1u << q;  // ambiguous (also: nothing within the 1e9 threshold, NaN coordinates)
                    }
                    need &= mb;
                }
                if (need) {
                    const int skip = (act == 2) ? pidx : -1, mvi = (act == 4) ? pidx : -1;
                    const double cx = h.prop[warp].cx, cy = h.prop[warp].cy, cz = h.prop[warp].cz;
                    const double *s_nx = reinterpret_cast<const double *>(tg_smem + h.o_nuc), *s_ny = s_nx + h.KC, *s_nz = s_ny + h.KC;
#pragma unroll 1
                    for (int q = 0; q < 4; q++)
                        if ((need >> q) & 1u) {
                            const uint32_t r = (uint32_t)rescan_point(h.px, h.py, h.pz, s_nx, s_ny, s_nz, K, skip, mvi, cx, cy, cz, 4 * w + q);
                            nbw = (nbw & ~(0xFFu << (8 * q))) | (r << (8 * q));
                        }
                }
                const uint32_t m8 = ((mb & 1u) * 0xFFu) | ((mb & 2u) * (0xFF00u >> 1)) | ((mb & 4u) * (0xFF0000u >> 2)) | ((mb & 8u) * (0xFF000000u >> 3));
                const uint32_t nw = (ow & ~m8) | (nbw & m8);  // death: old numbering, renumbered on accept
                // a point that stays with the moved nucleus keeps its zeta
                const bool chg = (act == 2) ? true : (((nbw ^ ((uint32_t)pidx * 0x01010101u)) & m8) != 0u);
                s_own32[w] = nw;
                if (chg) atomicOr(&s_dirtyw[w >> 5], 1u << (w & 31));
            }
            qn -= cnt;
            __syncwarp();
        }
        if (last) break;
    }
}

// =================================================================================================== phase C (one chunk of rays)
// Warp w owns the (length-sorted) rays r = c ST + NW lane + w of chunk c (phi_ray).  `info` = start | length << 18 of this
// lane's ray.  The lanes whose ray touches a changed word form the chunk's work list (ranks follow the length order); the
// warp re-integrates them 4 at a time in the canonical order (tstar_g8: 8 lanes per ray).  Returns this lane's new t* (if its
// ray is touched) and the chunk's touched-lane mask.
struct CRes { double tn; uint32_t dm; };
static __device__ __noinline__ CRes phase_c_chunk(const uint32_t info) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 3, sub = lane & 7;
    const uint32_t FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    const uint8_t *s_owner = tg_smem + h.o_owner;
    const uint32_t *s_dirtyw = reinterpret_cast<const uint32_t *>(tg_smem + h.o_dirtyw);
    const double *s_zh = h.zh;
    uint8_t *s_perm = hdr().perm[warp];
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
    const int cnt = __popc(dm), rank = __popc(dm & lt_mask);
    if (dirty) s_perm[rank] = (uint8_t)lane;
    __syncwarp();
    double tnv = 0.0;
#pragma unroll 1
    for (int t = 0; 4 * t < cnt; t++) {  // 4 touched rays at a time, 8 lanes each
        const int e = 4 * t + grp;
        const bool on = e < cnt;
        const uint32_t inf = __shfl_sync(FULL, info, on ? (int)s_perm[e] : 0);
        const int tq0 = (int)(inf & 0x3FFFFu), tnp = on ? (int)(inf >> 18) : 0;
        const int nseg = tnp > 1 ? tnp - 1 : 0;
        const int trip = __reduce_max_sync(FULL, (nseg + 7) >> 3);
        // Branch-free segment term: 0.5 (za + zb) == 0.5 za + 0.5 zb exactly (scaling by a power of two commutes with the
        // rounding), so the table holds the halved zeta and the term is dt * ((zh_a + zh_b) / 1000) -- seg_term's value.
        // Loads are unconditional (straight-line code, one address per stream): dt is allocated with max_npts + 128 zero
        // entries of slack and a stray owner byte indexes the 256-entry table; only the final add is predicated.
        const uint8_t *ow = s_owner + tq0 + sub;
        const double *dtp = h.dt + tq0 + sub;
        const int nl = nseg - sub;  // this lane's segments: j = sub + 8 k < nseg  <=>  8 k < nl
        double acc = 0.0;
        // This loop issues 60 % of the kernel's shared-memory wavefronts and global L1 requests (ncu source page, tools/ncu_lsu.py), and
        // the L1 / LSU data pipe is the busiest unit of the SM (66 %): loads that are not issued count for more than the arithmetic
        // they feed.  Measured and rejected (profiles/README.md, round 2): skipping the ARITHMETIC of the passes beyond `trip`
        // (-20 % of the loop's instructions, no gain); guarding their shared-memory loads by warp-uniform branches (the staging
        // breaks up: 27.2 against 28.4 M/s); separate code for the last 1..3 passes (phase C 22.0 k -> 18.5 k cycles, but the build
        // falls into the slow regime: 22.5 M/s); the first round's dt of the NEXT group requested while this one is summed (phase C
        // -11 %, the other phases grow by as much).
        // Rounds of 2 passes, staged (all loads of a stage before the first use); dt -- the only global (L2) loads -- is requested one round
        // ahead, and only if that round exists.  Rounds of 4 passes overlap more but waste more (Tonga's rays need 2, 3, 5, 9 or 17
        // passes) and are twice the code: phase C itself is 3 % faster with them, every other phase 3-5 % slower (instruction supply);
        // a fully rolled loop of single passes is 22 % slower in phase C (profiles/README.md, round 2).
        double dn[2];
#pragma unroll
        for (int u = 0; u < 2; u++) dn[u] = __ldg(dtp + 8 * u);
#pragma unroll 1
        for (int k0 = 0; k0 < trip; k0 += 2, ow += 16) {
            double d[2], za[2], zb[2];
            uint32_t oa[2], ob[2];
            dtp += 16;
#pragma unroll
            for (int u = 0; u < 2; u++) { d[u] = dn[u]; oa[u] = ow[8 * u]; ob[u] = ow[8 * u + 1]; }
            if (k0 + 2 < trip) {
#pragma unroll
                for (int u = 0; u < 2; u++) dn[u] = __ldg(dtp + 8 * u);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) { za[u] = s_zh[oa[u]]; zb[u] = s_zh[ob[u]]; }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const double term = __dmul_rn(d[u], div1000_exact(__dadd_rn(za[u], zb[u])));
                if (8 * (k0 + u) < nl) acc = __dadd_rn(acc, term);
            }
        }
        acc = __dadd_rn(acc, __shfl_xor_sync(FULL, acc, 4));
        acc = __dadd_rn(acc, __shfl_xor_sync(FULL, acc, 2));
        acc = __dadd_rn(acc, __shfl_xor_sync(FULL, acc, 1));
        const double v = __shfl_sync(FULL, acc, (rank & 3) << 3);
        if (dirty && (rank >> 2) == t) tnv = v;
    }
    __syncwarp();
    CRes r;
    r.tn = tnv; r.dm = dm;
    return r;
}

// =================================================================================================== phase F (owner commit / roll back)
// masked bytes (death / move): orphans, old owner = pidx.  Reject: restore.  Accept: refresh the owner-distance cache (a move
// also changes it for points that stay with the nucleus).
static __device__ __noinline__ void phase_f_mask(const int accepted, const int pidx, const uint32_t blkmask) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    const uint32_t kk = (uint32_t)pidx * 0x01010101u;
    uint32_t *s_own32 = reinterpret_cast<uint32_t *>(tg_smem + h.o_owner);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(tg_smem + h.o_mask);
    const float *s_fx = reinterpret_cast<const float *>(tg_smem + h.o_nucf), *s_fy = s_fx + h.KC, *s_fz = s_fy + h.KC;
    const int nBlocks = h.nBlocks;
    BlockIter iter(blkmask);
#pragma unroll 1
    for (int blk = warp + NW * iter.next(); blk < nBlocks; blk = warp + NW * iter.next()) {
        const int w = blk * 32 + lane;
        const uint32_t mb = (s_mask[w >> 3] >> (sub * 4)) & 0xFu;
        __syncwarp();
        if (sub == 0) s_mask[w >> 3] = 0u;
        if (mb) {
            const uint32_t ow = s_own32[w];
            if (!accepted) {
                const uint32_t m8 = ((mb & 1u) * 0xFFu) | ((mb & 2u) * (0xFF00u >> 1)) | ((mb & 4u) * (0xFF0000u >> 2)) | ((mb & 8u) * (0xFF000000u >> 3));
                s_own32[w] = (ow & ~m8) | (kk & m8);
            } else {
                const float4 X = __ldg(reinterpret_cast<const float4 *>(h.pxf) + w), Y = __ldg(reinterpret_cast<const float4 *>(h.pyf) + w),
                             Z = __ldg(reinterpret_cast<const float4 *>(h.pzf) + w);
                const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if ((mb >> q) & 1u) {
                        const int o = (ow >> (8 * q)) & 0x7F;  // death: still the old numbering, as are the fl32 nuclei
                        TG_DC_STORE(h.dcache + 4 * w + q, (o == TG_OWNER_NONE) ? 1e9f : dist2_f32(s_fx[o], s_fy[o], s_fz[o], xs[q], ys[q], zs[q]));
                    }
            }
        }
    }
}
// tagged bytes (birth / move): switch to the new / moved nucleus (accept: + cache their distance to it) or clear the tag.
static __device__ __noinline__ void phase_f_tags(const int accepted, const uint32_t newb, const uint32_t tagmask, const float cxf, const float cyf,
                                                 const float czf) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *s_own32 = reinterpret_cast<uint32_t *>(tg_smem + h.o_owner);
    const int nBlocks = h.nBlocks;
    BlockIter iter(tagmask);
#pragma unroll 1
    for (int blk = warp + NW * iter.next(); blk < nBlocks; blk = warp + NW * iter.next()) {
        const int w = blk * 32 + lane;
        const uint32_t ow = s_own32[w], t = ow & 0x80808080u;
        if (t) {
            const uint32_t m = (t >> 7) * 0xFFu;
            s_own32[w] = accepted ? ((ow & ~m) | (newb & m)) : (ow & 0x7F7F7F7Fu);
            if (accepted) {
                const float4 X = __ldg(reinterpret_cast<const float4 *>(h.pxf) + w), Y = __ldg(reinterpret_cast<const float4 *>(h.pyf) + w),
                             Z = __ldg(reinterpret_cast<const float4 *>(h.pzf) + w);
                const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if ((t >> (8 * q + 7)) & 1u) TG_DC_STORE(h.dcache + 4 * w + q, dist2_f32(cxf, cyf, czf, xs[q], ys[q], zs[q]));
            }
        }
    }
}
// accepted death: deleteat! renumbering -- owner indices above `kill` shift down (TD_inversion_function.jl:132-135)
static __device__ __noinline__ void phase_f_renumber(const int pidx) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t kk = (uint32_t)pidx * 0x01010101u;
    uint32_t *s_own32 = reinterpret_cast<uint32_t *>(tg_smem + h.o_owner);
    const int nBlocks = h.nBlocks;
#pragma unroll 2
    for (int blk = warp; blk < nBlocks; blk += NW) {
        const int w = blk * 32 + lane;
        const uint32_t ow = s_own32[w];
        const uint32_t gt = __vcmpgtu4(ow, kk) & ~__vcmpeq4(ow, 0x7F7F7F7Fu);
        if (gt) s_own32[w] = ow - (gt & 0x01010101u);
    }
}

// accepted death: order-preserving delete of nucleus `pidx` (deleteat!, TD_inversion_function.jl:132-135) by one warp
static __device__ __noinline__ void phase_f_delete(const int pidx, const int K) {
    const Hdr &h = hdr();
    const int lane = threadIdx.x & 31;
    double *s_nx = reinterpret_cast<double *>(tg_smem + h.o_nuc), *s_ny = s_nx + h.KC, *s_nz = s_ny + h.KC, *s_zeta = s_nz + h.KC;
    float *s_fx = reinterpret_cast<float *>(tg_smem + h.o_nucf), *s_fy = s_fx + h.KC, *s_fz = s_fy + h.KC;
    double *s_zh = hdr().zh;
    for (int s0 = 0; s0 < K - 1 - pidx; s0 += 32) {
        const int i = pidx + s0 + lane;
        double vx = 0, vy = 0, vz = 0, vt = 0;
        if (i < K - 1) { vx = s_nx[i + 1]; vy = s_ny[i + 1]; vz = s_nz[i + 1]; vt = s_zeta[i + 1]; }
        __syncwarp();
        if (i < K - 1) {
            s_nx[i] = vx; s_ny[i] = vy; s_nz[i] = vz; s_zeta[i] = vt; s_zh[i] = __dmul_rn(0.5, vt);
            s_fx[i] = (float)vx; s_fy[i] = (float)vy; s_fz[i] = (float)vz;
        }
        __syncwarp();
    }
    if (lane == 0) {  // freed slot
        s_fx[K - 1] = s_fy[K - 1] = s_fz[K - 1] = __int_as_float(0x7f800000);
        s_zh[K - 1] = 0.0;
    }
}

// thinning with a non-integral keep_each (TD_inversion_function.jl:278 in general): rare, kept out of the loop's code path
static __device__ __noinline__ int keep_by_fmod(long long model_num, double keep_each) { return fmod((double)model_num, keep_each) == 0; }

// the kept model -> history record (TD_inversion_function.jl:280).  Only the K valid nuclei of each axis are stored (the record
// keeps its fixed [4][KC] layout; the rest is never read).  The copy may overlap the next iteration's phases A..B: they modify
// neither the FP64 nuclei nor t* nor phi.
static __device__ __noinline__ void write_history(double *__restrict__ hc, double *__restrict__ hp, const int32_t *__restrict__ ray_rank, const int K,
                                                  const int R) {
    const Hdr &h = hdr();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *s_nx = reinterpret_cast<const double *>(tg_smem + h.o_nuc);
    const double *s_tstar = reinterpret_cast<const double *>(tg_smem + h.o_tstar);
    const int KC = h.KC;
#pragma unroll 1
    for (int k = lane; k < K; k += 32)
        for (int ax = warp; ax < 4; ax += NW) hc[ax * KC + k] = s_nx[ax * KC + k];
#pragma unroll 1
    for (int i = tid; i < R; i += ST) hp[i] = s_tstar[__ldg(ray_rank + i)];  // caller's ray order; contiguous stores (the history may live in mapped host memory)
}

// =================================================================================================== the kernel
// PLAIN: the generate-mode launch without proposal records and per-iteration traces (what a production run uses); the replay / record / trace
// paths are compiled out of this instantiation (168 instructions less on the once-per-iteration code path: instruction supply).
template <int NCH, bool PROF, bool PLAIN>
__global__ void __launch_bounds__(ST, 7) tg_sampler_kernel(const SamplerArgs a) {
    const SmemLayout L = smem_layout(a.Ppad, a.Rp, a.KC);
    Hdr &h = hdr();
    uint8_t *s_owner = tg_smem + L.o_owner;
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(tg_smem + L.o_mask);
    double *s_tstar = reinterpret_cast<double *>(tg_smem + L.o_tstar);
    double *s_nx = reinterpret_cast<double *>(tg_smem + L.o_nuc);
    double *s_ny = s_nx + a.KC, *s_nz = s_ny + a.KC, *s_zeta = s_nz + a.KC;
    float *s_fx = reinterpret_cast<float *>(tg_smem + L.o_nucf);
    float *s_fy = s_fx + a.KC, *s_fz = s_fy + a.KC;
    double *s_zh = h.zh;
    double *s_scr = h.scr;
    uint32_t *s_cnt = h.cnt;
    ChainScalars *s_st = &h.st;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t FULL = 0xffffffffu;
    WarpProp *s_prop = &h.prop[warp];
    const int chain = a.perm ? a.perm[blockIdx.x] : (int)blockIdx.x;  // launch order = cost order (tg_order_kernel)
    const int KC = a.KC, R = a.R;
    const int nMaskWords = a.Ppad / 32;
    const float FINF = __int_as_float(0x7f800000);

    // ---- load the chain state: three TMA bulk copies on one mbarrier
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&h.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        h.pxf = a.pxf; h.pyf = a.pyf; h.pzf = a.pzf; h.dcache = a.dcache + (size_t)chain * a.Ppad;
        h.px = a.px; h.py = a.py; h.pz = a.pz; h.dt = a.dt;
        h.o_owner = (uint32_t)L.o_owner; h.o_mask = (uint32_t)L.o_mask; h.o_tstar = (uint32_t)L.o_tstar; h.o_nuc = (uint32_t)L.o_nuc;
        h.o_nucf = (uint32_t)L.o_nucf; h.o_dirtyw = (uint32_t)L.o_dirtyw;
        h.KC = KC; h.nBlocks = a.Ppad / 128; h.R = R; h.exact_only = a.exact_only;
        h.ta = a.tol_alpha; h.tb = a.tol_beta2;
        // B2 carries the nucleus index in the 7 low mantissa bits of the fl32 distance (truncation: relative 2^-16): widen the band
        h.ta2 = a.tol_alpha * (1.0f + 0x1.0p-16f) + 0x1.0p-16f + 0x1.0p-20f;
        h.st.phi = a.phi[chain]; h.st.noise = a.noise[chain]; h.st.beta = a.beta[chain];
        h.st.pending_slot = a.pending_slot[chain];
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t b_owner = (uint32_t)a.Ppad, b_ts = (uint32_t)(8 * a.Rp), b_nuc = (uint32_t)(32 * KC);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(&h.bar)), "r"(b_owner + b_ts + b_nuc) : "memory");
        bulk_load(s_owner, a.owner + (size_t)chain * a.Ppad, b_owner, &h.bar);
        bulk_load(s_tstar, a.tstar + (size_t)chain * a.Rp, b_ts, &h.bar);
        bulk_load(s_nx, a.cells + (size_t)chain * 4 * KC, b_nuc, &h.bar);
    }
    for (int i = tid; i < nMaskWords; i += ST) s_mask[i] = 0u;
    for (int i = tid; i < 256; i += ST) s_owner[a.Ppad + i] = (uint8_t)TG_OWNER_NONE;  // the slack behind the owner bytes
    if (tid < 16) s_cnt[tid] = 0u;
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            s2u(&h.bar))
        : "memory");
    __syncthreads();
    int K = a.K[chain];
    // fl32 copies of the nuclei (slots >= K hold +inf so that they never win a comparison) and the halved-zeta table
    for (int i = tid; i < 3 * KC; i += ST) {
        const int k = i % KC;
        s_fx[i] = (k < K) ? (float)s_nx[i] : FINF;
    }
    for (int o = tid; o < 256; o += ST) s_zh[o] = (o < K) ? __dmul_rn(0.5, s_zeta[o]) : 0.0;
    // the rays of this thread: r = c ST + NW lane + warp (phi_ray); start and length packed in one register per chunk
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

    int n_hist = a.n_hist[chain];
    const bool cold = (a.beta[chain] == 1.0);  // tempered replicas (beta < 1, extension) do not contribute to the posterior

    const tonga_params &pm = a.prm;
    const int a_mode = PLAIN ? 0 : a.mode;
    tonga_proposal *const a_recs_out = PLAIN ? nullptr : a.recs_out;
    int8_t *const a_tr_accept = PLAIN ? nullptr : a.tr_accept;
    double *const a_tr_phi = PLAIN ? nullptr : a.tr_phi;
    int32_t *const a_tr_K = PLAIN ? nullptr : a.tr_K;
    const double sig_zeta = pm.zeta_scale * pm.sig / 100;  // TD_inversion_function.jl:22

    // thinning (:276-279): keep when fmod(model_num, keep_each) == 0.  For an integral keep_each (the reference's 1e1) that is a
    // counter; the general case keeps the FP64 fmod.  mn_inc = increments of model_num in this launch.
    const int keep_i = (pm.keep_each >= 1.0 && pm.keep_each < 2e9 && floor(pm.keep_each) == pm.keep_each) ? (int)pm.keep_each : 0;
    int keep_ctr = keep_i ? (int)(a.model_num[chain] % keep_i) : 0;
    int mn_inc = 0;
    long long pt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = PROF ? clock64() : 0;
    auto tick = [&](int ph) {
        if (PROF && tid == 0) { const long long t = clock64(); pt[ph] += t - tprev; tprev = t; }
    };
    // raw draw of the next iteration: lane l < 8 holds double l of the record (requested one iteration ahead)
    const double *rawp = (a_mode == 0) ? reinterpret_cast<const double *>(a.raw + (size_t)chain * a.nIter) + (lane & 7) : nullptr;
    double raw_next = (a_mode == 0 && a.nIter > 0) ? __ldcs(rawp) : 0.0;
    const int nIter = (int)a.nIter;  // a launch covers at most 2^31 iterations (sampler.cu cuts longer runs)

#pragma unroll 1
    for (int it = 0; it < nIter; it++) {
        // ================================================================ A: the proposal, assembled by every warp
        int act, do_eval, pidx;
        float cxf, cyf, czf;
        {
            Prop pr;
            pr.do_eval = 0; pr.accept = 0; pr.idx = 0; pr.action = 0;
            pr.x = pr.y = pr.z = pr.zeta = pr.u = pr.aux = pr.ox = pr.oy = pr.oz = pr.ztag = 0.0;
            double u1 = 0, u2 = 0, u3 = 0, n0 = 0, n1 = 0, n2 = 0, u7 = 0;
            if (a_mode == 0) {
                const double rv = raw_next;
                if (it + 1 < nIter) raw_next = __ldcs(rawp + 8 * (it + 1));
                pr.action = (int)__shfl_sync(FULL, rv, 0);
                u1 = __shfl_sync(FULL, rv, 1); u2 = __shfl_sync(FULL, rv, 2); u3 = __shfl_sync(FULL, rv, 3);
                n0 = __shfl_sync(FULL, rv, 4); n1 = __shfl_sync(FULL, rv, 5); n2 = __shfl_sync(FULL, rv, 6);
                u7 = __shfl_sync(FULL, rv, 7);
            } else {
                const tonga_proposal rec = a.recs_in[(size_t)chain * a.trace_stride + a.it0 + it];
                pr.action = rec.action; pr.idx = rec.idx; pr.x = rec.x; pr.y = rec.y; pr.z = rec.z; pr.zeta = rec.zeta; pr.u = rec.u;
            }
            assemble_proposal<0>(pr, a_mode, u1, u2, u3, n0, n1, n2, u7, s_nx, s_ny, s_nz, s_zeta, K, s_st->noise, pm, sig_zeta, lane);
            act = pr.action; do_eval = pr.do_eval; pidx = pr.idx;
            cxf = (float)pr.x; cyf = (float)pr.y; czf = (float)pr.z;
            if (lane == 0) {  // scalars needed again at the acceptance / commit: parked in shared memory, one copy per warp
                s_prop->u = pr.u; s_prop->aux = pr.aux; s_prop->zeta = pr.zeta; s_prop->ox = pr.ox; s_prop->oy = pr.oy; s_prop->oz = pr.oz;
                s_prop->zold = (do_eval && (act == 2 || act == 3)) ? s_zeta[pidx] : 0.0;  // zeta[idx] of the CURRENT model (acceptance rule)
                s_prop->cx = pr.x; s_prop->cy = pr.y; s_prop->cz = pr.z;
            }
            if (tid == 0) {
                if (a_mode == 0 && a_recs_out) {
                    tonga_proposal rec;
                    rec.action = pr.action; rec.idx = pr.idx; rec.x = pr.x; rec.y = pr.y; rec.z = pr.z; rec.zeta = pr.zeta; rec.u = pr.u;
                    a_recs_out[(size_t)chain * a.trace_stride + a.it0 + it] = rec;
                }
                const int ps = s_st->pending_slot;
                if (ps >= 0) { a.hist_next[(size_t)chain * a.hist_cap + ps] = pr.action; s_st->pending_slot = -1; }
            }
            // the zeta table under the proposal (read by phase C only, i.e. after the barrier below)
            if (do_eval) {
                if (act == 1 || act == 4) {  // tagged bytes: the new nucleus / the moved nucleus
                    const double zt = (act == 1) ? __dmul_rn(0.5, pr.zeta) : s_zh[pidx];
                    for (int o = ZLUT_TAG + tid; o < 256; o += ST) s_zh[o] = zt;
                } else if (act == 3 && tid == 0) s_zh[pidx] = __dmul_rn(0.5, pr.zeta);
            }
        }
        tick(0);
        int accepted = 0;
        double tn[NCH];        // proposal's t* of this thread's touched rays
        uint32_t dmask[NCH];   // per chunk: which lanes of this warp hold a touched ray
#pragma unroll
        for (int c = 0; c < NCH; c++) { tn[c] = 0.0; dmask[c] = 0u; }

        if (do_eval) {
            uint32_t blkmask = 0u;  // own blocks holding orphans (bit = warp-iteration; iterations >= 32 are always scanned)
            uint32_t tagmask = 0u;  // own blocks holding tagged bytes
            if (act != 5) {
                // the fl32 nuclei under the proposal: every warp writes the same values before its own use (the FP64 arrays stay untouched)
                if (lane == 0) {
                    if (act == 2) s_fx[pidx] = FINF;  // the screening must not see the killed nucleus
                    else if (act == 4) { s_fx[pidx] = cxf; s_fy[pidx] = cyf; s_fz[pidx] = czf; }
                }
                __syncwarp();
                // ======================================================== B: the point pass of the action
                if (act == 1) { const uint2 m = phase_b1_switch<1>(pidx, cxf, cyf, czf); tagmask = m.x; }
                else if (act == 4) { const uint2 m = phase_b1_switch<4>(pidx, cxf, cyf, czf); tagmask = m.x; blkmask = m.y; }
                else blkmask = phase_b1_scan(act, pidx);
                tick(8);  // profile slot 8 = B1, slot 1 = B2 (+ barrier)
                if (act == 2 || act == 4) {
                    __syncwarp();
                    phase_b2(act, pidx, K, blkmask);
                }
                __syncthreads();
                tick(1);
                // ======================================================== C: t* of this warp's touched rays
#pragma unroll 1
                for (int c = 0; c < NCH; c++) {
                    if (c * ST >= R) break;  // warp-uniform
                    uint32_t info = 0u;
#pragma unroll
                    for (int k = 0; k < NCH; k++) info = (k == c) ? ri[k] : info;
                    const CRes cr = phase_c_chunk(info);
#pragma unroll
                    for (int k = 0; k < NCH; k++)
                        if (k == c) { tn[k] = cr.tn; dmask[k] = cr.dm; }
                }
                tick(2);
            }
            // ============================================================ D: phi of the proposed model (canonical order)
            {
                const double nz = (act == 5) ? s_prop->zeta : s_st->noise;
                double acc = 0.0;
#pragma unroll 1
                for (int c = 0; c < NCH; c++) {
                    const int r = phi_ray(c, tid);
                    if (r >= R) break;  // rays of later chunks are >= R as well
                    double tv = 0.0;
                    uint32_t dm = 0u;
#pragma unroll
                    for (int k = 0; k < NCH; k++)
                        if (k == c) { tv = tn[k]; dm = dmask[k]; }
                    const double t = ((dm >> lane) & 1u) ? tv : s_tstar[r];
                    acc = __dadd_rn(acc, misfit_term(t, a.tS[r], a.sig[r], nz));
                }
                acc = warp_sum_canonical(acc);
                if (lane == 0) s_scr[(it & 1) * 4 + warp] = acc;
            }
            __syncthreads();
            // ============================================================ E: acceptance, evaluated by every thread
            double phin;
            {
                phin = phi_warp_sums(s_scr + (it & 1) * 4);
                if (pm.debug_prior) phin = 1.0;  // MCsub.jl:128-136: the chain samples the prior
                Prop pr;
                pr.action = act; pr.idx = pidx; pr.zeta = s_prop->zeta; pr.aux = s_prop->aux; pr.u = s_prop->u;
                accepted = accept_decision_call(pr, K, s_st->phi, phin, s_prop->zold, s_st->noise, s_st->beta, R, pm, sig_zeta);
            }
            tick(3);
            // ============================================================ F: commit / roll back, every warp on its own blocks and rays
            if (act == 2 || act == 4) phase_f_mask(accepted, pidx, blkmask);
            tick(6);
            if (act == 1 || act == 4) phase_f_tags(accepted, (uint32_t)(act == 1 ? K : pidx) * 0x01010101u, tagmask, cxf, cyf, czf);
            tick(7);
            if (accepted && act != 5) {  // t* of the touched rays
#pragma unroll
                for (int c = 0; c < NCH; c++)
                    if ((dmask[c] >> lane) & 1u) s_tstar[phi_ray(c, tid)] = tn[c];
            }
            if (act == 2 && accepted) {
                phase_f_renumber(pidx);
                __syncthreads();  // the cache refresh above read the nuclei the delete below shifts
                if (warp == 0) phase_f_delete(pidx, K);
            }
            if (tid == 0) {
                if (act == 1 && accepted) {  // append!, :85-88
                    s_nx[K] = s_prop->cx; s_ny[K] = s_prop->cy; s_nz[K] = s_prop->cz; s_zeta[K] = s_prop->zeta;
                    s_fx[K] = cxf; s_fy[K] = cyf; s_fz[K] = czf;
                    s_zh[K] = __dmul_rn(0.5, s_prop->zeta);
                } else if (act == 3) {
                    if (accepted) s_zeta[pidx] = s_prop->zeta;
                    else s_zh[pidx] = __dmul_rn(0.5, s_zeta[pidx]);
                } else if (act == 2 && !accepted) {
                    s_fx[pidx] = (float)s_nx[pidx];  // un-hide the nucleus that was proposed for deletion
                } else if (act == 4) {
                    if (accepted) { s_nx[pidx] = s_prop->cx; s_ny[pidx] = s_prop->cy; s_nz[pidx] = s_prop->cz; }
                    else { s_fx[pidx] = (float)s_prop->ox; s_fy[pidx] = (float)s_prop->oy; s_fz[pidx] = (float)s_prop->oz; }
                }
                if (accepted) {
                    s_st->phi = phin;
                    if (act == 5) s_st->noise = s_prop->zeta;
                }
            }
            if (accepted) {
                if (act == 1) K += 1;
                else if (act == 2) K -= 1;
            }
            __syncthreads();  // the committed state is consistent before the next proposal / the history copy
        }
        tick(4);
        // ================================================================ G: bookkeeping, traces, thinning (:275-281)
        const long long iter = a.iter0 + it;
        int keep = 0;
        if ((double)iter >= pm.burn_in) {
            mn_inc += 1;
            if (keep_i) { if (++keep_ctr == keep_i) { keep_ctr = 0; keep = 1; } }
            else keep = keep_by_fmod(a.model_num[chain] + mn_inc, pm.keep_each);
        }
        if (tid == 0) {
            if (act >= 1 && act <= 5) {
                s_cnt[act - 1] += 1u;
                if (accepted) s_cnt[5 + act - 1] += 1u;
                if (do_eval) s_cnt[10 + act - 1] += 1u;
            }
            if (a_tr_accept) a_tr_accept[(size_t)chain * a.trace_stride + a.it0 + it] = (int8_t)accepted;
            if (a_tr_phi) a_tr_phi[(size_t)chain * a.trace_stride + a.it0 + it] = s_st->phi;
            if (a_tr_K) a_tr_K[(size_t)chain * a.trace_stride + a.it0 + it] = K;
        }
        if (keep && cold) {
            if (n_hist < a.hist_cap) {
                const size_t hh = (size_t)chain * a.hist_cap + n_hist;
                write_history(a.hist_cells + hh * 4 * KC, a.hist_ptS + hh * R, a.ray_rank, K, R);
                if (tid == 0) {
                    a.hist_K[hh] = K; a.hist_phi[hh] = s_st->phi; a.hist_iter[hh] = iter;
                    a.hist_action[hh] = act; a.hist_accept[hh] = accepted; a.hist_next[hh] = 0;
                    s_st->pending_slot = n_hist;
                }
            }
            n_hist += 1;
        }
        tick(5);
    }

    if (PROF && tid == 0) {
        for (int i = 0; i < 9; i++) a.prof[(size_t)chain * 16 + i] += pt[i];
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.prof[(size_t)chain * 16 + 14] = (long long)blockIdx.x;  // where the chain ran (launch-order studies)
        a.prof[(size_t)chain * 16 + 15] = (long long)smid;
    }
    // ---- write the chain state back (TMA bulk stores for the arrays)
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
        bulk_store(a.owner + (size_t)chain * a.Ppad, s_owner, (uint32_t)a.Ppad);
        bulk_store(a.tstar + (size_t)chain * a.Rp, s_tstar, (uint32_t)(8 * a.Rp));
        bulk_store(a.cells + (size_t)chain * 4 * KC, s_nx, (uint32_t)(32 * KC));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        a.K[chain] = K; a.phi[chain] = s_st->phi; a.noise[chain] = s_st->noise;
        a.n_hist[chain] = n_hist; a.model_num[chain] += mn_inc; a.pending_slot[chain] = s_st->pending_slot;
        long long *c = a.counts + (size_t)chain * 15;
        for (int i = 0; i < 15; i++) c[i] += (long long)s_cnt[i];
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

}  // namespace tg
