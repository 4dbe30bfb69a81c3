// evaluate.cu -- full forward model + misfit, batched over models:
//   evaluate       MCsub.jl:123-185   (ray loop :142-163, misfit :169-175, "likelihood" :179-182)
//   Interpolation  MCsub.jl:306-336   (NaN trim, slice broadcast, nearest nucleus per point)
//   v_nearest      MCsub.jl:247-263   (strict <, lowest index wins, 1e9 start value)
//
// Kernel 1 (tg_eval_kernel): CTA = (ray tile, model).  The model's nuclei are staged in shared memory; phase 1 is a
// flat, fully occupied pass over the tile's points (4 points per thread held in registers while the nucleus loop
// broadcasts each nucleus from shared memory) that produces the owner of every point; phase 2 integrates t* in the
// canonical order (8 lanes per ray, tstar_g8).  Kernel 2 (tg_phi_kernel) reduces the per-ray misfit terms in the canonical
// order.  FP64 throughout, no FMA contraction in the distance (bit-exact owners).
#include <cmath>
#include <cstring>

#include "tonga_internal.cuh"

namespace tg {

constexpr int EVAL_THREADS = 256;
constexpr int EVAL_PPT = 4;  // points per thread per pass
#define TG_NONE16 0xFFFFu

// Screening arithmetic of phase 1 (FP32, any rounding is fine: a rigorous error band decides what is re-done in FP64):
// with c = centre of the nucleus box, p' = p - c, n' = n - c,
//     |p - n|^2 = |p'|^2 + ( |n'|^2 - 2 p'.n' )
// so a nucleus is four fl32 coefficients (-2 n'x, -2 n'y, -2 n'z, |n'|^2) -- ONE 128-bit shared-memory broadcast -- and a
// (point, nucleus) pair costs 3 FMAs, issued as packed FFMA2 over point pairs (1.5 instructions per pair instead of 3 for
// dx^2 + dy^2 + dz^2); |p'|^2 does not depend on the nucleus and is added once, after the selection.  Best and second best are kept as INTEGER min / max of the distance bits with the
// nucleus index (mod 128) in the 7 low mantissa bits (non-negative floats order like their bits; nuclei are visited in
// chunks of 128): 4 ALU instructions per pair instead of a compare and three selects.
// Error band (u = 2^-24; per axis a: A_a = max |p'_a|, |n'_a| over the tile's model, M_a = max |p_a|): every one of the five
// terms is <= 3u off (two input roundings, one operation), the four additions round partial sums bounded by
// S = 4 sum_a A_a^2, and p' inherits the fl32 rounding of p:  |D32 - D| <= u sum_a (28 A_a^2 + 4 A_a (M_a + A_a)) =: B.
// The index bits truncate by 2^-16 relative.  Two distances are separated for certain if
//     d2 - d1 > alpha (d1 + d2) + beta,   alpha = 2^-16 + 2^-19,  beta = 2.5 B
// (2 B for the two values, 25 % margin); anything else -- also a best within the band of the 1e9 "no nucleus" threshold --
// goes to a per-tile queue and is re-scanned in exact FP64 by the whole CTA after the pass (no divergent rescans inside it).
#ifndef EVAL_MIN_CTAS
#define EVAL_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_CTAS)
tg_eval_kernel(const Tile *__restrict__ tiles, int Kcap, const int32_t *__restrict__ Ks, const double *__restrict__ cells,
               const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
               const float *__restrict__ pxf, const float *__restrict__ pyf, const float *__restrict__ pzf, double cenx, double ceny,
               double cenz, float mpx, float mpy, float mpz /* max |p'| per axis over the ray set */, float mux, float muy,
               float muz /* max |p| per axis */, int exact_only, int stage_xyz, const double *__restrict__ dt, const int32_t *__restrict__ ray_off, const int32_t *__restrict__ ray_orig,
               const int32_t *__restrict__ point_orig, int R, int64_t P, int64_t Ppad, int tile_pts,
               double *__restrict__ ptS, int32_t *__restrict__ owners32, uint8_t *__restrict__ owners8, float *__restrict__ dmin32,
               uint16_t *__restrict__ owners16) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int model = blockIdx.y;
    const Tile tile = tiles[blockIdx.x];
    const int K = Ks[model];
    if (K < 0) return;  // wide sampler: this chain has no candidate to evaluate in this iteration
    // shared layout: [nx[Kcap] ny[Kcap] nz[Kcap]] nzeta[Kcap] | mbarrier, counters | float4 coefficients [Kcap + 8] | owner16[tile_pts] | queue16[tile_pts]
    // The FP64 coordinates are only read by the prologue and the (rare) exact re-scan: for large models (stage_xyz == 0) they stay in
    // global memory (L2-resident, uniform addresses) and the CTA keeps 48 KB less at K = 2000 -- 3 instead of 2 CTAs per SM.
    const double *mc = cells + (size_t)model * 4 * Kcap;
    double *s_stage = reinterpret_cast<double *>(smem_raw);
    const int nrows = stage_xyz ? 4 : 1;
    double *s_zeta = s_stage + (nrows - 1) * Kcap;
    const double *s_nx = stage_xyz ? s_stage : mc, *s_ny = s_nx + Kcap, *s_nz = s_ny + Kcap;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_zeta + Kcap);
    int *s_nq = reinterpret_cast<int *>(s_bar + 1);
    float *s_red = reinterpret_cast<float *>(s_nq + 2);  // [3][8] per-warp maxima
    float4 *s_cf = reinterpret_cast<float4 *>(s_red + 24);
    uint16_t *s_owner = reinterpret_cast<uint16_t *>(s_cf + Kcap + 8);
    uint16_t *s_queue = s_owner + tile_pts;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const double *src = mc + (4 - nrows) * Kcap;
    const uint32_t nbytes = (uint32_t)(nrows * Kcap * sizeof(double));
    const bool use_bulk = (nbytes >= 2048u) && ((nbytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    if (use_bulk) {  // one TMA bulk copy of the whole [4][Kcap] block (or of its zeta row)
        if (tid == 0) {
            mbar_init(s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(s_bar, nbytes);
            bulk_g2s(s_stage, src, nbytes, s_bar);
        }
        mbar_wait(s_bar, 0);
    } else {
        for (int i = tid; i < nrows * Kcap; i += EVAL_THREADS) s_stage[i] = src[i];
        __syncthreads();
    }
    // the model's coordinate bound (for the band), then the fl32 coefficients of the nuclei
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int i = tid; i < K; i += EVAL_THREADS) {
        ax = fmaxf(ax, fabsf((float)(s_nx[i] - cenx))); ay = fmaxf(ay, fabsf((float)(s_ny[i] - ceny))); az = fmaxf(az, fabsf((float)(s_nz[i] - cenz)));
    }
    for (int off = 16; off > 0; off >>= 1) {
        ax = fmaxf(ax, __shfl_xor_sync(0xffffffffu, ax, off)); ay = fmaxf(ay, __shfl_xor_sync(0xffffffffu, ay, off)); az = fmaxf(az, __shfl_xor_sync(0xffffffffu, az, off));
    }
    if (lane == 0) { s_red[warp] = ax; s_red[8 + warp] = ay; s_red[16 + warp] = az; }
    if (tid == 0) s_nq[0] = 0;
    __syncthreads();
    float tol_beta, d_off;
    {
        float A[3] = {mpx, mpy, mpz};
        for (int w = 0; w < EVAL_THREADS / 32; w++) { A[0] = fmaxf(A[0], s_red[w]); A[1] = fmaxf(A[1], s_red[8 + w]); A[2] = fmaxf(A[2], s_red[16 + w]); }
        const float Mu[3] = {mux, muy, muz};
        float B = 0.f;
        for (int a_ = 0; a_ < 3; a_++) {
            const float Aa = A[a_] * 1.000001f + 1e-30f;  // (A itself was rounded to fl32)
            B += 28.f * Aa * Aa + 4.f * Aa * (Mu[a_] + Aa);
        }
        B *= 5.9604644775390625e-08f * 1.00001f;
        d_off = 1.5f * B;     // added to every |n'|^2: the computed distances stay >= 0 (non-negative floats order like their bits)
        tol_beta = 2.5f * B + 6.0f * 5.9604644775390625e-08f * d_off;
    }
    for (int i = tid; i < K + 8; i += EVAL_THREADS) {
        if (i < K) {
            const double nx = s_nx[i] - cenx, ny = s_ny[i] - ceny, nz = s_nz[i] - cenz;
            s_cf[i] = make_float4((float)(-2.0 * nx), (float)(-2.0 * ny), (float)(-2.0 * nz), (float)(nx * nx + ny * ny + nz * nz + (double)d_off));
        } else {
            s_cf[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));  // padding nuclei (the loop visits blocks of 8): +inf, never win
        }
    }
    __syncthreads();
    const float tol_alpha = 0x1.0p-16f + 0x1.0p-19f;
    const float fcx = (float)cenx, fcy = (float)ceny, fcz = (float)cenz;

    // ---- phase 1: owners of the tile's points (v_nearest, MCsub.jl:247-263)
    const int npts = tile.p1 - tile.p0;
    for (int base = 0; base < npts; base += EVAL_THREADS * EVAL_PPT) {
        int pidx[EVAL_PPT];
        bool in[EVAL_PPT];
#pragma unroll
        for (int q = 0; q < EVAL_PPT; q++) {
            const int j = base + q * EVAL_THREADS + tid;
            in[q] = j < npts;
            pidx[q] = tile.p0 + (in[q] ? j : 0);
        }
        float D1[EVAL_PPT], D2[EVAL_PPT];
        int I1[EVAL_PPT];
        uint32_t need_exact = exact_only ? (1u << EVAL_PPT) - 1u : 0u;
        float xs[EVAL_PPT], ys[EVAL_PPT], zs[EVAL_PPT];
#pragma unroll
        for (int q = 0; q < EVAL_PPT; q++) { xs[q] = pxf[pidx[q]]; ys[q] = pyf[pidx[q]]; zs[q] = pzf[pidx[q]]; D1[q] = 1e9f; D2[q] = 1e9f; I1[q] = -1; }
        if (!exact_only) {
            const float2 X01 = make_float2(xs[0] - fcx, xs[1] - fcx), X23 = make_float2(xs[2] - fcx, xs[3] - fcx);
            const float2 Y01 = make_float2(ys[0] - fcy, ys[1] - fcy), Y23 = make_float2(ys[2] - fcy, ys[3] - fcy);
            const float2 Z01 = make_float2(zs[0] - fcz, zs[1] - fcz), Z23 = make_float2(zs[2] - fcz, zs[3] - fcz);
            const float2 Q01 = __ffma2_rn(Z01, Z01, __ffma2_rn(Y01, Y01, __fmul2_rn(X01, X01))), Q23 = __ffma2_rn(Z23, Z23, __ffma2_rn(Y23, Y23, __fmul2_rn(X23, X23)));
            // Blocks of 8 nuclei.  Inside a block only the block minimum is kept (one 3-input min per two pairs); across blocks
            // best / second best of the block minima and the id of the best block (5 ALU instructions per block): ~1.1 ALU
            // instructions per pair instead of 3.5 with per-pair (best, second best, index) tracking -- the FMA pipe (2 per pair) is
            // the limiter again.  The winning block is recomputed afterwards (identical arithmetic) for the index and the
            // in-block runner-up.
            float B1[EVAL_PPT], B2[EVAL_PPT];
            int BK[EVAL_PPT];
#pragma unroll
            for (int q = 0; q < EVAL_PPT; q++) { B1[q] = __int_as_float(0x7f800000); B2[q] = __int_as_float(0x7f800000); BK[q] = 0; }
            const int nblk = (K + 7) >> 3;
#pragma unroll 1
            for (int b = 0; b < nblk; b++) {
                const float4 *cf = s_cf + 8 * b;
                float bm[EVAL_PPT];
#pragma unroll
                for (int q = 0; q < EVAL_PPT; q++) bm[q] = __int_as_float(0x7f800000);
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    const float4 cA = cf[i], cB = cf[i + 1];
                    const float2 ca = make_float2(cA.x, cA.x), cbv = make_float2(cA.y, cA.y), cc = make_float2(cA.z, cA.z), ce = make_float2(cA.w, cA.w);
                    const float2 da = __ffma2_rn(X01, ca, __ffma2_rn(Y01, cbv, __ffma2_rn(Z01, cc, ce)));
                    const float2 db = __ffma2_rn(X23, ca, __ffma2_rn(Y23, cbv, __ffma2_rn(Z23, cc, ce)));
                    const float2 ga = make_float2(cB.x, cB.x), gb = make_float2(cB.y, cB.y), gc = make_float2(cB.z, cB.z), ge = make_float2(cB.w, cB.w);
                    const float2 ea = __ffma2_rn(X01, ga, __ffma2_rn(Y01, gb, __ffma2_rn(Z01, gc, ge)));
                    const float2 eb = __ffma2_rn(X23, ga, __ffma2_rn(Y23, gb, __ffma2_rn(Z23, gc, ge)));
                    bm[0] = fminf(fminf(bm[0], da.x), ea.x); bm[1] = fminf(fminf(bm[1], da.y), ea.y);
                    bm[2] = fminf(fminf(bm[2], db.x), eb.x); bm[3] = fminf(fminf(bm[3], db.y), eb.y);
                }
#pragma unroll
                for (int q = 0; q < EVAL_PPT; q++) {
                    const bool lt = bm[q] < B1[q];
                    B2[q] = fminf(B2[q], fmaxf(B1[q], bm[q]));
                    BK[q] = lt ? b : BK[q];
                    B1[q] = fminf(B1[q], bm[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < EVAL_PPT; q++) {  // the winning block again: index of the best, runner-up inside the block
                const float X = (q < 2) ? (q == 0 ? X01.x : X01.y) : (q == 2 ? X23.x : X23.y), Y = (q < 2) ? (q == 0 ? Y01.x : Y01.y) : (q == 2 ? Y23.x : Y23.y),
                            Z = (q < 2) ? (q == 0 ? Z01.x : Z01.y) : (q == 2 ? Z23.x : Z23.y), Q = (q < 2) ? (q == 0 ? Q01.x : Q01.y) : (q == 2 ? Q23.x : Q23.y);
                const float4 *cf = s_cf + 8 * BK[q];
                float d1 = __int_as_float(0x7f800000), d2 = __int_as_float(0x7f800000);
                int i1 = -1;
#pragma unroll 2
                for (int i = 0; i < 8; i++) {
                    const float4 c = cf[i];
                    const float d = __fmaf_rn(X, c.x, __fmaf_rn(Y, c.y, __fmaf_rn(Z, c.z, c.w)));
                    const bool lt = d < d1;
                    d2 = lt ? d1 : fminf(d2, d);
                    i1 = lt ? i : i1;
                    d1 = lt ? d : d1;
                }
                // (d1 == B1 bit for bit: same operations; were it not, d1 / d2 below are still the block's true best / second best)
                // |p'|^2 joins after the selection: x -> fl(x + Q) is monotone, so best / second best of the sums are the sums of the
                // best / second best (values that only tie after the rounding are inside the band anyway -> exact re-scan)
                const float s1 = __fadd_rn(d1, Q), s2 = __fadd_rn(fminf(B2[q], d2), Q);
                if (i1 >= 0 && s1 < 1e9f) { D1[q] = s1; D2[q] = fminf(s2, 1e9f); I1[q] = 8 * BK[q] + i1; }
            }
#pragma unroll
            for (int q = 0; q < EVAL_PPT; q++) {
                const float tol = fmaf(tol_alpha, D1[q] + D2[q], tol_beta);
                if (!(D2[q] - D1[q] > tol)) need_exact |= 1u << q;  // ambiguous (D2 starts at 1e9: also covers the 1e9 threshold; NaN too)
            }
        }
#pragma unroll
        for (int q = 0; q < EVAL_PPT; q++) {
            const int j = base + q * EVAL_THREADS + tid;
            if (!in[q]) continue;
            if ((need_exact >> q) & 1u) { s_queue[atomicAdd(&s_nq[0], 1)] = (uint16_t)j; continue; }
            const int bi = I1[q];
            s_owner[j] = bi < 0 ? (uint16_t)TG_NONE16 : (uint16_t)bi;
            const int64_t p = tile.p0 + j;
            if (owners32) owners32[(size_t)model * P + point_orig[p]] = bi;  // caller's flat order
            if (owners8) owners8[(size_t)model * Ppad + p] = bi < 0 ? (uint8_t)TG_OWNER_NONE : (uint8_t)bi;
            if (owners16) owners16[(size_t)model * Ppad + p] = s_owner[j];  // streamed sampler's chain state
            if (dmin32)  // owner distance cache of the samplers: the direct fl32 form they use themselves
                dmin32[(size_t)model * Ppad + p] = bi < 0 ? 1e9f : dist2_f32((float)s_nx[bi], (float)s_ny[bi], (float)s_nz[bi], xs[q], ys[q], zs[q]);
        }
    }
    __syncthreads();
    // ---- the queue: exact FP64 re-scan (MCsub.jl:252-259: strict <, lowest index wins ties, 1e9 start value)
    // A short queue (the screened mode: a handful of near ties per tile) is re-scanned ONE WARP PER POINT -- lane l takes the nuclei
    // l, l + 32, ..., then the warp picks the smallest distance, lowest index on exact ties -- so that a few points do not keep the
    // whole CTA waiting at the barrier below for one thread's K-long serial loop; a long queue (exact-only mode) one thread per point.
    const int nq = s_nq[0];
    if (nq < 2 * EVAL_THREADS) {
        for (int qi = warp; qi < nq; qi += EVAL_THREADS / 32) {
            const int j = s_queue[qi];
            const int64_t p = tile.p0 + j;
            const double x = px[p], y = py[p], z = pz[p];
            double best = 1e9;  // mdist = 1e9, MCsub.jl:250
            int b = 0x7fffffff;
            for (int i = lane; i < K; i += 32) {
                const double d = dist2_exact(s_nx[i], s_ny[i], s_nz[i], x, y, z);
                if (d < best) { best = d; b = i; }  // ascending i within the lane: strict < keeps the lowest index
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, b, off);
                if (ob < best || (ob == best && oi < b)) { best = ob; b = oi; }
            }
            if (b == 0x7fffffff) b = -1;  // nothing within the 1e9 threshold
            if (lane == 0) {
                s_owner[j] = b < 0 ? (uint16_t)TG_NONE16 : (uint16_t)b;
                if (owners32) owners32[(size_t)model * P + point_orig[p]] = b;
                if (owners8) owners8[(size_t)model * Ppad + p] = b < 0 ? (uint8_t)TG_OWNER_NONE : (uint8_t)b;
                if (owners16) owners16[(size_t)model * Ppad + p] = s_owner[j];
                if (dmin32) dmin32[(size_t)model * Ppad + p] = b < 0 ? 1e9f : (float)best;
            }
        }
    } else
    for (int qi = tid; qi < nq; qi += EVAL_THREADS) {
        const int j = s_queue[qi];
        const int64_t p = tile.p0 + j;
        const double x = px[p], y = py[p], z = pz[p];
        double best = 1e9;  // mdist = 1e9, MCsub.jl:250
        int b = -1;
#pragma unroll 2
        for (int i = 0; i < K; i++) {
            const double d = dist2_exact(s_nx[i], s_ny[i], s_nz[i], x, y, z);
            if (d < best) { best = d; b = i; }  // strict <: lowest index wins ties (:255)
        }
        s_owner[j] = b < 0 ? (uint16_t)TG_NONE16 : (uint16_t)b;
        if (owners32) owners32[(size_t)model * P + point_orig[p]] = b;
        if (owners8) owners8[(size_t)model * Ppad + p] = b < 0 ? (uint8_t)TG_OWNER_NONE : (uint8_t)b;
        if (owners16) owners16[(size_t)model * Ppad + p] = s_owner[j];
        if (dmin32) dmin32[(size_t)model * Ppad + p] = b < 0 ? 1e9f : (float)best;
    }
    __syncthreads();

    // ---- phase 2: t* per ray in the canonical order (tstar_g8: 8 lanes per ray, 4 length-sorted rays per warp), MCsub.jl:147,153/159
    auto zeta_of = [&](uint16_t o) -> double { return o == TG_NONE16 ? 0.0 : s_zeta[o]; };
    const int grp = (threadIdx.x & 31) >> 3, sub = threadIdx.x & 7;
    for (int rb = tile.r0 + 4 * (threadIdx.x >> 5); rb < tile.r1; rb += 4 * (EVAL_THREADS / 32)) {  // warp-uniform
        const int r = rb + grp;
        const bool on = r < tile.r1;
        const int q0 = on ? ray_off[r] : 0, n = on ? ray_off[r + 1] - q0 : 0;
        const int nseg = n > 1 ? n - 1 : 0;
        const int trip = __reduce_max_sync(0xffffffffu, (nseg + 7) >> 3);
        const uint16_t *ow = s_owner + (q0 - tile.p0);
        const double t = tstar_g8(nseg, trip, sub, [&](int j) { return seg_term(dt[q0 + j], zeta_of(ow[j]), zeta_of(ow[j + 1])); });
        if (on && sub == 0) ptS[(size_t)model * R + ray_orig[r]] = t;  // caller's ray order
    }
}

// padded tail of the u8 chain-state owner arrays: NONE
__global__ void tg_owner_pad_kernel(uint8_t *owners8, float *dmin32, int64_t P, int64_t Ppad, int nModels) {
    const int64_t pad = Ppad - P;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < pad * nModels) {
        owners8[(i / pad) * Ppad + P + (i % pad)] = (uint8_t)TG_OWNER_NONE;
        if (dmin32) dmin32[(i / pad) * Ppad + P + (i % pad)] = 1e9f;
    }
}

__global__ void __launch_bounds__(TG_PHI_LANES)
tg_phi_kernel(int R, const double *__restrict__ ptS /* caller's ray order */, const int32_t *__restrict__ ray_orig,
              const double *__restrict__ tS, const double *__restrict__ sig /* sorted ray order */,
              const double *__restrict__ noise, double *__restrict__ phi, int debug_prior) {
    __shared__ double scratch[4];
    const int model = blockIdx.x;
    if (debug_prior) {  // MCsub.jl:128-136: phi = 1, nothing else computed
        if (threadIdx.x == 0) phi[model] = 1.0;
        return;
    }
    const double nz = noise ? noise[model] : 1.0;
    const double *t = ptS + (size_t)model * R;
    const double v = phi_canonical(R, threadIdx.x, scratch, [&](int r) { return misfit_term(t[ray_orig[r]], tS[r], sig[r], nz); });
    if (threadIdx.x == 0) phi[model] = v;
}

int launch_evaluate(tonga_ctx *ctx, int nModels, int Kcap, const int32_t *K_dev, const double *cells_dev,
                    const double *noise_dev, double *ptS_dev, double *phi_dev, int32_t *owners32_dev, uint8_t *owners8_dev, float *dmin32_dev, bool force_geometry, uint16_t *owners16_dev, int exact_only) {
    if (nModels <= 0) return TONGA_OK;
    if (nModels > 65535) return fail(TONGA_ERR_CAPACITY, "evaluate: at most 65535 models per call");
    if ((!ctx->prm.debug_prior || force_geometry) && ctx->n_tiles > 0) {
        const int ex = exact_only < 0 ? ctx->exact_only : exact_only;
        const int stage_xyz = (ex || Kcap <= 1024) ? 1 : 0;  // large models: FP64 coordinates stay in global memory (3 CTAs per SM at K = 2000)
        const size_t smem = sizeof(double) * (stage_xyz ? 4 : 1) * (size_t)Kcap + 8 + 8 + 96 + 16 * ((size_t)Kcap + 8) + 2 * sizeof(uint16_t) * (size_t)ctx->tile_pts;
        if (smem > ctx->smem_optin) return fail(TONGA_ERR_CAPACITY, "evaluate: Kcap too large for shared memory");
        TG_CUDA(cudaFuncSetAttribute(tg_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(ctx->n_tiles, nModels);
        tg_eval_kernel<<<grid, EVAL_THREADS, smem, ctx->stream>>>(ctx->d_tiles, Kcap, K_dev, cells_dev, ctx->d_px, ctx->d_py,
                                                                  ctx->d_pz, ctx->d_pxf, ctx->d_pyf, ctx->d_pzf, ctx->cen[0], ctx->cen[1], ctx->cen[2],
                                                                  ctx->mp_cen[0], ctx->mp_cen[1], ctx->mp_cen[2], ctx->mp_abs[0], ctx->mp_abs[1], ctx->mp_abs[2],
                                                                  ex, stage_xyz, ctx->d_dt, ctx->d_ray_off, ctx->d_ray_orig, ctx->d_point_orig,
                                                                  ctx->R, ctx->P, ctx->Ppad, ctx->tile_pts, ptS_dev,
                                                                  owners32_dev, owners8_dev, dmin32_dev, owners16_dev);
        TG_CUDA(cudaGetLastError());
    }
    if (owners8_dev && ctx->Ppad > ctx->P) {
        const int64_t n = (ctx->Ppad - ctx->P) * nModels;
        tg_owner_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(owners8_dev, dmin32_dev, ctx->P, ctx->Ppad, nModels);
        TG_CUDA(cudaGetLastError());
    }
    if (phi_dev) {
        tg_phi_kernel<<<nModels, TG_PHI_LANES, 0, ctx->stream>>>(ctx->R, ptS_dev, ctx->d_ray_orig, ctx->d_tS, ctx->d_sig, noise_dev, phi_dev,
                                                                ctx->prm.debug_prior);
        TG_CUDA(cudaGetLastError());
    }
    return TONGA_OK;
}

// one thread per query point; nuclei staged in shared memory
__global__ void __launch_bounds__(256)
tg_interp_kernel(int K, const double *__restrict__ cells /* [4][K] */, int n, const double *__restrict__ X, int nY,
                 const double *__restrict__ Y, int nZ, const double *__restrict__ Z, double *__restrict__ zeta_out,
                 int32_t *__restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s = reinterpret_cast<double *>(smem_raw);
    for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) s[i] = cells[i];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double x = X[k], y = (nY == 1) ? Y[0] : Y[k], z = (nZ == 1) ? Z[0] : Z[k];
    double best = 1e9, v = 0.0;
    int bi = -1;
    for (int i = 0; i < K; i++) {
        const double d = dist2_exact(s[i], s[K + i], s[2 * K + i], x, y, z);
        if (d < best) { best = d; bi = i; v = s[3 * K + i]; }
    }
    zeta_out[k] = v;
    if (idx_out) idx_out[k] = bi;
}

}  // namespace tg

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int tonga_evaluate_batch_dev(tonga_ctx *ctx, int32_t nModels, int32_t Kcap, const int32_t *K_dev,
                                        const double *cells_dev, const double *noise_dev, double *ptS_dev, double *phi_dev,
                                        int32_t *owners_dev) {
    if (!ctx || nModels < 0 || Kcap < 1 || !K_dev || !cells_dev || !ptS_dev)
        return tg::fail(TONGA_ERR_ARG, "tonga_evaluate_batch_dev: bad argument (ptS_dev is required)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    return tg::launch_evaluate(ctx, nModels, Kcap, K_dev, cells_dev, noise_dev, ptS_dev, phi_dev, owners_dev, nullptr, nullptr);
}

extern "C" int tonga_evaluate_batch(tonga_ctx *ctx, int32_t nModels, int32_t Kcap, const int32_t *K, const double *cells,
                                    const double *noise, double *ptS, double *phi, int32_t *owners) {
    if (!ctx || nModels < 0 || Kcap < 1 || !K || !cells) return tg::fail(TONGA_ERR_ARG, "tonga_evaluate_batch: bad argument");
    if (nModels == 0) return TONGA_OK;
    for (int i = 0; i < nModels; i++)
        if (K[i] < 0 || K[i] > Kcap) return tg::fail(TONGA_ERR_ARG, "tonga_evaluate_batch: K[i] outside [0, Kcap]");
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    const size_t R = (size_t)ctx->R, P = (size_t)ctx->P, n = (size_t)nModels;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_K = 0, o_cells = o_K + al(4 * n), o_noise = o_cells + al(8 * n * 4 * Kcap), o_pts = o_noise + al(8 * n),
                 o_phi = o_pts + al(8 * n * (R ? R : 1)), o_own = o_phi + al(8 * n), total = o_own + (owners ? al(4 * n * (P ? P : 1)) : 0);
    int rc = tg::ensure_scratch(ctx, total);
    if (rc != TONGA_OK) return rc;
    char *d = (char *)ctx->d_scratch;
    cudaStream_t s = ctx->stream;
    TG_CUDA(cudaMemcpyAsync(d + o_K, K, 4 * n, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(d + o_cells, cells, 8 * n * 4 * Kcap, cudaMemcpyHostToDevice, s));
    if (noise) TG_CUDA(cudaMemcpyAsync(d + o_noise, noise, 8 * n, cudaMemcpyHostToDevice, s));
    rc = tg::launch_evaluate(ctx, nModels, Kcap, (const int32_t *)(d + o_K), (const double *)(d + o_cells),
                             noise ? (const double *)(d + o_noise) : nullptr, (double *)(d + o_pts), (double *)(d + o_phi),
                             owners ? (int32_t *)(d + o_own) : nullptr, nullptr, nullptr);
    if (rc != TONGA_OK) return rc;
    if (ptS && R && !ctx->prm.debug_prior) TG_CUDA(cudaMemcpyAsync(ptS, d + o_pts, 8 * n * R, cudaMemcpyDeviceToHost, s));
    if (phi) TG_CUDA(cudaMemcpyAsync(phi, d + o_phi, 8 * n, cudaMemcpyDeviceToHost, s));
    if (owners && P && !ctx->prm.debug_prior) TG_CUDA(cudaMemcpyAsync(owners, d + o_own, 4 * n * P, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    return TONGA_OK;
}

extern "C" int tonga_misfit(tonga_ctx *ctx, int32_t nModels, const double *ptS, const double *noise, double *phi) {
    if (!ctx || nModels < 0 || !ptS || !phi) return tg::fail(TONGA_ERR_ARG, "tonga_misfit: bad argument");
    if (nModels == 0) return TONGA_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    const size_t R = (size_t)ctx->R, n = (size_t)nModels;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_pts = 0, o_noise = o_pts + al(8 * n * (R ? R : 1)), o_phi = o_noise + al(8 * n), total = o_phi + al(8 * n);
    int rc = tg::ensure_scratch(ctx, total);
    if (rc != TONGA_OK) return rc;
    char *d = (char *)ctx->d_scratch;
    cudaStream_t s = ctx->stream;
    if (R) TG_CUDA(cudaMemcpyAsync(d + o_pts, ptS, 8 * n * R, cudaMemcpyHostToDevice, s));
    if (noise) TG_CUDA(cudaMemcpyAsync(d + o_noise, noise, 8 * n, cudaMemcpyHostToDevice, s));
    tg::tg_phi_kernel<<<nModels, TG_PHI_LANES, 0, s>>>(ctx->R, (const double *)(d + o_pts), ctx->d_ray_orig, ctx->d_tS, ctx->d_sig,
                                                       noise ? (const double *)(d + o_noise) : nullptr, (double *)(d + o_phi), ctx->prm.debug_prior);
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaMemcpyAsync(phi, d + o_phi, 8 * n, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    return TONGA_OK;
}

extern "C" int tonga_evaluate(tonga_ctx *ctx, int32_t K, const double *x, const double *y, const double *z,
                              const double *zeta, double noise, double *ptS_out, double *phi_out, double *like_out,
                              double *loglik_out) {
    if (!ctx || K < 0 || (K > 0 && (!x || !y || !z || !zeta))) return tg::fail(TONGA_ERR_ARG, "tonga_evaluate: bad argument");
    const int Kcap = K > 0 ? K : 1;
    std::vector<double> cells((size_t)4 * Kcap, 0.0);
    if (K > 0) {
        std::memcpy(&cells[0], x, 8 * (size_t)K);
        std::memcpy(&cells[Kcap], y, 8 * (size_t)K);
        std::memcpy(&cells[2 * (size_t)Kcap], z, 8 * (size_t)K);
        std::memcpy(&cells[3 * (size_t)Kcap], zeta, 8 * (size_t)K);
    }
    double phi = 0.0;
    int rc = tonga_evaluate_batch(ctx, 1, Kcap, &K, cells.data(), &noise, ptS_out, &phi, nullptr);
    if (rc != TONGA_OK) return rc;
    if (phi_out) *phi_out = phi;
    if (ctx->prm.debug_prior) {  // MCsub.jl:129-136: likelihood = 1
        if (like_out) *like_out = 1.0;
        if (loglik_out) *loglik_out = 1.0;
        return TONGA_OK;
    }
    // MCsub.jl:179 with allSig scaled by `noise`: sum_k (-log(noise*sig_k*sqrt(2pi))) * R
    const double R = (double)ctx->R;
    if (like_out) *like_out = (noise == 1.0) ? ctx->like_const : (ctx->sum_neglog - R * std::log(noise)) * R;
    if (loglik_out) *loglik_out = (ctx->sum_neglog - R * std::log(noise)) - 0.5 * phi;
    return TONGA_OK;
}

extern "C" int tonga_interpolate(tonga_ctx *ctx, int32_t K, const double *x, const double *y, const double *z,
                                 const double *zeta, int32_t nX, const double *X, int32_t nY, const double *Y, int32_t nZ,
                                 const double *Z, double *zeta_out, int32_t *idx_out, int32_t *npoints_out) {
    if (!ctx || K < 0 || nX < 0 || (K > 0 && (!x || !y || !z || !zeta)) || (nX > 0 && (!X || !Y || !Z || !zeta_out)))
        return tg::fail(TONGA_ERR_ARG, "tonga_interpolate: bad argument");
    int npoints = nX;  // MCsub.jl:312-316
    for (int k = 0; k < nX; k++)
        if (std::isnan(X[k])) { npoints = k; break; }
    if (npoints_out) *npoints_out = npoints;
    if (npoints == 0) return TONGA_OK;
    if ((nY != 1 && nY < npoints) || (nZ != 1 && nZ < npoints))
        return tg::fail(TONGA_ERR_ARG, "tonga_interpolate: Y / Z shorter than the trimmed X (Julia would throw BoundsError)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)npoints, k4 = (size_t)4 * (K > 0 ? K : 1);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t ny = (nY == 1) ? 1 : n, nz = (nZ == 1) ? 1 : n;
    const size_t o_c = 0, o_X = o_c + al(8 * k4), o_Y = o_X + al(8 * n), o_Z = o_Y + al(8 * ny), o_o = o_Z + al(8 * nz),
                 o_i = o_o + al(8 * n), total = o_i + al(4 * n);
    if (8 * k4 > ctx->smem_optin) return tg::fail(TONGA_ERR_CAPACITY, "tonga_interpolate: K too large for shared memory");
    int rc = tg::ensure_scratch(ctx, total);
    if (rc != TONGA_OK) return rc;
    char *d = (char *)ctx->d_scratch;
    cudaStream_t s = ctx->stream;
    if (K > 0) {
        TG_CUDA(cudaMemcpyAsync(d + o_c, x, 8 * (size_t)K, cudaMemcpyHostToDevice, s));
        TG_CUDA(cudaMemcpyAsync(d + o_c + 8 * (size_t)K, y, 8 * (size_t)K, cudaMemcpyHostToDevice, s));
        TG_CUDA(cudaMemcpyAsync(d + o_c + 16 * (size_t)K, z, 8 * (size_t)K, cudaMemcpyHostToDevice, s));
        TG_CUDA(cudaMemcpyAsync(d + o_c + 24 * (size_t)K, zeta, 8 * (size_t)K, cudaMemcpyHostToDevice, s));
    }
    TG_CUDA(cudaMemcpyAsync(d + o_X, X, 8 * n, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(d + o_Y, Y, 8 * ny, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(d + o_Z, Z, 8 * nz, cudaMemcpyHostToDevice, s));
    const size_t smem = 8 * k4;
    TG_CUDA(cudaFuncSetAttribute(tg::tg_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    tg::tg_interp_kernel<<<(unsigned)((n + 255) / 256), 256, smem, s>>>(K, (const double *)(d + o_c), npoints, (const double *)(d + o_X),
                                                                      nY, (const double *)(d + o_Y), nZ, (const double *)(d + o_Z),
                                                                      (double *)(d + o_o), (int32_t *)(d + o_i));
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaMemcpyAsync(zeta_out, d + o_o, 8 * n, cudaMemcpyDeviceToHost, s));
    if (idx_out) TG_CUDA(cudaMemcpyAsync(idx_out, d + o_i, 4 * n, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    return TONGA_OK;
}
