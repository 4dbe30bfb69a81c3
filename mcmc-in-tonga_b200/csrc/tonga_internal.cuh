// tonga_internal.cuh -- shared device helpers and host-side structs of libtonga_b200.so (sm_100a only).
//
// Numerical contract (DESIGN.md "Exactness"):
//   * squared distance exactly as MCsub.jl:254 evaluates it in Julia: fl(fl(dx*dx + dy*dy) + dz*dz) with
//     dx = mx - x; no FMA contraction (explicit __dmul_rn / __dadd_rn), so owners are bit-exact;
//   * per-segment term exactly as MCsub.jl:147,153: (rayl*rayu) * ((0.5*(za+zb)) / 1000); the division is done
//     with a 3-op FMA sequence that returns the correctly rounded quotient (checked against `/` on 3e8 samples);
//   * per-ray misfit term exactly as MCsub.jl:171: ((d*d)*1.0) / (sig*sig);
//   * ONE canonical summation order, shared by every kernel, so "incremental == full" holds bit for bit on the device:
//     t* of a ray is an 8-lane interleaved sum + butterfly (tstar_g8; the reference's own order -- Julia `sum`, SIMD
//     pairwise -- is unspecified); phi is a fixed 128-lane strided sum + butterfly over the rays in length-sorted order
//     (the reference sums k = 1..R sequentially); hence a 1e-9 relative tolerance on t* and phi, none on the owners.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/tonga_b200.h"

// ------------------------------------------------------------------------------------------------ constants
#define TG_OWNER_NONE 0x7Fu  // u8 owner code: no nucleus closer than sqrt(1e9) (v_nearest returns 0.0, MCsub.jl:249-250)
#define TG_OWNER_TAG 0x80u   // u8 owner bit 7: pending "switches to the proposal's implicit new owner"
#define TG_MAX_K_U8 126      // largest nCells representable with u8 owners (indices 0..125, +1 for a birth)
#ifndef TG_PHI_WARPS
#define TG_PHI_WARPS 4       // warps of the canonical phi reduction == warps per chain of the resident sampler
#endif
#define TG_PHI_LANES (32 * TG_PHI_WARPS)  // lanes of the canonical phi reduction
#define TG_PT_TILE 512       // points are padded to a multiple of this (128 threads x 4 points)
#define TG_PAD_COORD 1e18f    // fl32 coordinates of the padding points (P..Ppad): finite and far away, so that the screening arithmetic
                             // never sees a NaN (their FP64 coordinates stay NaN; they are owned by "none" and never switch)
#define TG_PT_SLACK 2048     // extra elements behind the fl32 coordinate arrays / the owner-distance cache: the resident sampler prefetches
                             // up to two block rounds ahead without a bounds check

namespace tg {

// ------------------------------------------------------------------------------------------------ errors
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
#define TG_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return tg::fail(TONGA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

// ------------------------------------------------------------------------------------------------ exact math
__device__ __forceinline__ double dist2_exact(double mx, double my, double mz, double x, double y, double z) {
    // (mx - x)^2 + (my - y)^2 + (mz - z)^2, MCsub.jl:254 -- never contracted into FMAs
    const double dx = __dsub_rn(mx, x), dy = __dsub_rn(my, y), dz = __dsub_rn(mz, z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// FP32 screening distance (any rounding / FMA order is fine: the error band of DESIGN.md 4.2 covers it)
__device__ __forceinline__ float dist2_f32(float mx, float my, float mz, float x, float y, float z) {
    const float dx = mx - x, dy = my - y, dz = mz - z;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

__device__ __forceinline__ double div1000_exact(double a) {
    // correctly rounded a / 1000 (Markstein: y = RN(1/b), q0 = a*y, r = a - b*q0 exactly, q = q0 + r*y)
    const double y = 0.001;
    const double q0 = __dmul_rn(a, y);
    const double r = __fma_rn(-q0, 1000.0, a);
    return __fma_rn(r, y, q0);
}

__device__ __forceinline__ double seg_term(double dt, double za, double zb) {
    // MCsub.jl:147,153:  (rayl*rayu) * ((0.5*(za+zb)) / 1000);  dt = fl(rayl*rayu) is precomputed on the host
    return __dmul_rn(dt, div1000_exact(__dmul_rn(0.5, __dadd_rn(za, zb))));
}

// the general segment term as a real function call (segments that cross a cell boundary are the minority; the resident
// sampler's hot loop keeps only the call site): zlut maps the (clamped) owner byte to zeta under the proposed model
static __device__ __noinline__ double seg_term_mixed(double dt, const double *zlut, int oa, int ob) { return seg_term(dt, zlut[oa], zlut[ob]); }

__device__ __forceinline__ double misfit_term(double pts, double ts, double sig, double noise) {
    // MCsub.jl:171:  ((ptS - tS)^2 * 1.0) / sig^2   with sig = noise * allSig (noise == 1.0 -> reference)
    const double sg = __dmul_rn(noise, sig);
    const double d = __dsub_rn(pts, ts);
    return __ddiv_rn(__dmul_rn(__dmul_rn(d, d), 1.0), __dmul_rn(sg, sg));
}

__device__ __forceinline__ double warp_sum_canonical(double v) {
    // xor butterfly: every lane ends with the same, order-defined sum
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// Canonical t* of one ray (shared by every kernel, so "incremental == full" holds bit for bit): a GROUP OF 8 LANES owns the
// ray; lane s (0..7) adds the terms of segments j = s, s+8, s+16, ... in ascending order into an accumulator that starts
// at 0.0, then the 8 partial sums are combined by the xor butterfly 4, 2, 1 (every lane ends with the same value).  A warp
// integrates 4 rays at a time; `trip` is the warp-uniform number of passes (>= ceil(nseg / 8) of every group in the warp),
// groups without a ray pass nseg = 0.  term(j) is only called for j < nseg.  The reference's own order (Julia `sum` over a
// broadcast vector, MCsub.jl:153/159) is SIMD-pairwise and unspecified; the contract on t* is 1e-9 relative.
// dt is stored flat, in point order: dt[p] = rayl*rayu of the segment from flat point p to p+1 (0.0 at the last point of a ray).
template <typename TermOf>
__device__ __forceinline__ double tstar_g8(int nseg, int trip, int sub, TermOf term) {
    double acc = 0.0;
    for (int k = 0; k < trip; k++) {
        const int j = 8 * k + sub;
        if (j < nseg) acc = __dadd_rn(acc, term(j));
    }
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    return acc;
}
// The same sum with the loads of 8 passes in flight at once: load(j) fetches what term(j, loaded) needs from global memory.
template <typename LoadOf, typename TermOf>
__device__ __forceinline__ double tstar_g8_batched(int nseg, int trip, int sub, LoadOf load, TermOf term) {
    double acc = 0.0;
    for (int k0 = 0; k0 < trip; k0 += 8) {
        double d[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int j = 8 * (k0 + u) + sub;
            d[u] = (j < nseg) ? load(j) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int j = 8 * (k0 + u) + sub;
            if (j < nseg) acc = __dadd_rn(acc, term(j, d[u]));
        }
    }
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    return acc;
}

// Canonical phi over R per-ray terms by a group of exactly TG_PHI_LANES threads (TG_PHI_WARPS warps): thread (warp w, lane l)
// adds the terms of the (length-sorted) rays r = LANES c + WARPS l + w, c = 0, 1, ... in ascending order (so every warp gets
// the same mix of ray lengths -- the resident sampler integrates t* of ray r on warp r % WARPS and keeps it in that thread's
// registers), then xor butterfly inside the warp, then the warp sums left to right.
__device__ __forceinline__ int phi_ray(int c, int t) { return c * TG_PHI_LANES + (t & 31) * TG_PHI_WARPS + (t >> 5); }
__device__ __forceinline__ double phi_warp_sums(const double *scratch) {
    double tot = scratch[0];
#pragma unroll
    for (int w = 1; w < TG_PHI_WARPS; w++) tot = __dadd_rn(tot, scratch[w]);
    return tot;
}
// term(r) must be callable by any thread.  scratch: TG_PHI_WARPS doubles of shared memory.  Result valid in all threads.
template <typename TermOf>
__device__ __forceinline__ double phi_canonical(int R, int t, double *scratch, TermOf term) {
    double acc = 0.0;
    for (int c = 0; c * TG_PHI_LANES < R; c++) {
        const int r = phi_ray(c, t);
        if (r < R) acc = __dadd_rn(acc, term(r));
    }
    acc = warp_sum_canonical(acc);
    if ((t & 31) == 0) scratch[t >> 5] = acc;
    __syncthreads();
    const double tot = phi_warp_sums(scratch);
    __syncthreads();
    return tot;
}

// ------------------------------------------------------------------------------------------------ TMA bulk copies + mbarrier
// Bulk global->shared copy through the TMA unit (cp.async.bulk, SASS UBLKCP) with an mbarrier; used to stage a
// model's nuclei when the tile is large enough to be worth it (>= 2 KB and 16-byte aligned).
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}


// ------------------------------------------------------------------------------------------------ Philox4x32-10
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
        const uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
        uint32_t ka = k0, kb = k1;
#pragma unroll 1  // rolled: the sampler kernel is instruction-cache sensitive (profiles/README.md) and the draw is off the hot loops
        for (int i = 0; i < 10; i++) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, c0, h0, l0);
            mulhilo(0xCD9E8D57u, c2, h1, l1);
            const uint32_t n0 = h1 ^ c1 ^ ka, n1 = l1, n2 = h0 ^ c3 ^ kb, n3 = l0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
__host__ __device__ inline double u53(uint32_t a, uint32_t b) {  // uniform [0,1), 53 bits
    return (double)((((uint64_t)a << 32) | b) >> 11) * 0x1.0p-53;
}
__host__ __device__ inline double u53_open(uint32_t a, uint32_t b) {  // uniform (0,1)
    return ((double)((((uint64_t)a << 32) | b) >> 11) + 0.5) * 0x1.0p-53;
}

// ------------------------------------------------------------------------------------------------ host structs
struct Tile { int r0, r1, p0, p1; };  // rays [r0,r1), flat points [p0,p1)

}  // namespace tg

struct tonga_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    tonga_params prm{};
    int m = 0, R = 0;
    int64_t P = 0, S = 0, Ppad = 0;
    int max_npts = 0;
    // device geometry (SoA, flat point order, padded with NaN coordinates / zero dt)
    double *d_px = nullptr, *d_py = nullptr, *d_pz = nullptr;  // [Ppad]
    float *d_pxf = nullptr, *d_pyf = nullptr, *d_pzf = nullptr; // [Ppad] fl32 copies for the FP32 screening pass
    int exact_only = 0;                                        // 1: full evaluate without FP32 screening (tonga_set_exact_only)
    float tol_alpha = 0.f, tol_beta2 = 0.f;                    // screening band: |dc - do| <= alpha*(dc+do) + beta2 -> exact FP64 recheck
    double cen[3] = {0, 0, 0};                                 // centre of the nucleus box: origin of the full evaluate's screening arithmetic
    float mp_cen[3] = {0, 0, 0}, mp_abs[3] = {0, 0, 0};        // per axis: max |p - cen| and max |p| over the ray points (its error band)
    // Internally rays are SORTED by length (descending); "flat point order" on the device is the CSR order of the sorted
    // rays.  ray_orig / point_orig map back to the caller's order at the API boundary.
    double *d_dt = nullptr;                                    // [Ppad] dt = rayL*rayU of the segment p -> p+1 (0.0 at the last point of a ray and in the padding)
    int32_t *d_rayid = nullptr;                                // [Ppad] sorted ray index of each flat point
    int32_t *d_ray_off = nullptr;                              // [R+1]  CSR offsets over sorted rays
    int32_t *d_ray_orig = nullptr;                             // [R]    sorted ray index -> caller's ray index
    int32_t *d_ray_rank = nullptr;                             // [R]    caller's ray index -> sorted ray index (inverse of ray_orig)
    int32_t *d_point_orig = nullptr;                           // [Ppad] sorted flat point -> caller's flat point (-1 for padding)
    double *d_tS = nullptr, *d_sig = nullptr;                  // [R]    in sorted ray order
    int Rp = 0;                                                // R rounded up to a multiple of 2 (16-byte rows)
    tg::Tile *d_tiles = nullptr;
    int n_tiles = 0;
    int tile_pts = 0;
    std::vector<int32_t> h_ray_off;     // caller's order (tonga_ray_offsets)
    std::vector<int32_t> h_ray_orig;    // sorted -> caller's ray
    std::vector<int32_t> h_point_orig;  // sorted flat point -> caller's flat point
    double like_const = 0.0;  // MCsub.jl:179 for noise = 1
    double sum_neglog = 0.0;  // sum_k -log(allSig_k*sqrt(2pi))
    int sm_count = 0;
    size_t smem_optin = 0;
    std::mutex mu;
    // scratch for host-pointer entry points (grown on demand)
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *h_pinned = nullptr;
    size_t pinned_bytes = 0;
};

namespace tg {
// screening bounds from the per-axis range [lo, hi] of the ray points and the nucleus box of ctx->prm (both creation paths)
void set_screening_bounds(tonga_ctx *ctx, const double lo[3], const double hi[3]);
int ensure_scratch(tonga_ctx *ctx, size_t bytes);
int ensure_pinned(tonga_ctx *ctx, size_t bytes);
// launches (all asynchronous on ctx->stream)
int launch_evaluate(tonga_ctx *ctx, int nModels, int Kcap, const int32_t *K_dev, const double *cells_dev,
                    const double *noise_dev, double *ptS_dev, double *phi_dev, int32_t *owners32_dev,
                    uint8_t *owners8_dev /* [nModels][Ppad] chain-state layout, or NULL */,
                    float *dmin32_dev /* [nModels][Ppad] fl32 squared distance to the owner (sampler cache), or NULL */, bool force_geometry = false,
                    uint16_t *owners16_dev = nullptr /* [nModels][Ppad] streamed-sampler chain state, or NULL */,
                    int exact_only = -1 /* -1: the context's setting (tonga_set_exact_only); 0 / 1: override for this call */);
}  // namespace tg
