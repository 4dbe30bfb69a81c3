// ctx.cu -- geometry context of libtonga_b200.so: the one-time flatten of the reference's DataStruct ray arrays
// (DefStruct.jl:5-30) into device-resident SoA arrays, plus error plumbing.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "tonga_internal.cuh"

namespace tg {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int ensure_scratch(tonga_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return TONGA_OK;
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + (bytes >> 2) + 4096;
    TG_CUDA(cudaMalloc(&ctx->d_scratch, want));
    ctx->scratch_bytes = want;
    return TONGA_OK;
}
int ensure_pinned(tonga_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes) return TONGA_OK;
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr;
    ctx->pinned_bytes = 0;
    size_t want = bytes + (bytes >> 2) + 4096;
    TG_CUDA(cudaMallocHost(&ctx->h_pinned, want));
    ctx->pinned_bytes = want;
    return TONGA_OK;
}
}  // namespace tg

extern "C" const char *tonga_last_error(void) { return tg::g_err.c_str(); }
extern "C" int tonga_version(void) { return 100; }
extern "C" int tonga_device_count(int *n_out) {
    if (!n_out) return tg::fail(TONGA_ERR_ARG, "tonga_device_count: n_out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *n_out = 0;
        return tg::fail(TONGA_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *n_out = n;
    return TONGA_OK;
}

template <typename T>
static int upload(T **dptr, const std::vector<T> &h, cudaStream_t s) {
    TG_CUDA(cudaMalloc((void **)dptr, sizeof(T) * (h.size() ? h.size() : 1)));
    if (h.size()) TG_CUDA(cudaMemcpyAsync(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s));
    return TONGA_OK;
}

extern "C" int tonga_create(tonga_ctx **out, int32_t m, int32_t R, const double *rayX, const double *rayY,
                            const double *rayZ, const double *rayL, const double *rayU, const double *tS,
                            const double *allSig, const tonga_params *params, int32_t device) {
    if (!out) return tg::fail(TONGA_ERR_ARG, "tonga_create: out is NULL");
    *out = nullptr;
    if (m < 1 || R < 0 || !params || (R > 0 && (!rayX || !rayY || !rayZ || !tS || !allSig)) ||
        (R > 0 && m > 1 && (!rayL || !rayU)))
        return tg::fail(TONGA_ERR_ARG, "tonga_create: bad argument");
    if (params->interp_style != 1)
        return tg::fail(TONGA_ERR_ARG, "tonga_create: only interp_style == 1 (nearest) exists; style 2 is broken in the reference (MCsub.jl:332)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return tg::fail(TONGA_ERR_CUDA, "tonga_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return tg::fail(TONGA_ERR_ARG, "tonga_create: bad device index");
    TG_CUDA(cudaSetDevice(device));

    // ---- flatten: MCsub.jl:312-316 (points up to the first NaN of X), :149-161 (segments up to the first NaN of rayL)
    std::vector<int32_t> ray_off(R + 1, 0);
    int max_npts = 0;
    for (int i = 0; i < R; i++) {
        const double *X = rayX + (size_t)i * m;
        int np = m;
        for (int k = 0; k < m; k++)
            if (std::isnan(X[k])) { np = k; break; }
        int nseg = m - 1;
        for (int j = 0; j < m - 1; j++)
            if (std::isnan(rayL[(size_t)i * (m - 1) + j])) { nseg = j; break; }
        const int nz = np > 0 ? np - 1 : 0;
        if (nseg != nz)
            return tg::fail(TONGA_ERR_DATA, "tonga_create: ray " + std::to_string(i) + " has " + std::to_string(np) +
                                                " points but " + std::to_string(nseg) +
                                                " segments (Julia would throw DimensionMismatch at MCsub.jl:153/159)");
        ray_off[i + 1] = ray_off[i] + np;
        if (np > max_npts) max_npts = np;
    }
    const int64_t P = ray_off[R];
    const int64_t Ppad = ((P + TG_PT_TILE - 1) / TG_PT_TILE) * TG_PT_TILE + (P == 0 ? TG_PT_TILE : 0);
    const int Rp = (R + 1) & ~1;
    // device order: rays sorted by length, longest first (stable), so that the 32 rays a warp integrates finish together
    std::vector<int32_t> ray_orig(R);
    for (int i = 0; i < R; i++) ray_orig[i] = i;
    std::stable_sort(ray_orig.begin(), ray_orig.end(), [&](int a, int b) {
        return (ray_off[a + 1] - ray_off[a]) > (ray_off[b + 1] - ray_off[b]);
    });
    std::vector<int32_t> sray_off(R + 1, 0);
    for (int rs = 0; rs < R; rs++) sray_off[rs + 1] = sray_off[rs] + (ray_off[ray_orig[rs] + 1] - ray_off[ray_orig[rs]]);
    const double qnan = std::nan("");
    std::vector<double> px(Ppad, qnan), py(Ppad, qnan), pz(Ppad, qnan);
    std::vector<double> dt(Ppad + max_npts + 128, 0.0);  // flat, point order: dt[p] = segment p -> p+1; zero slack for unconditional batched loads
    std::vector<int32_t> rayid(Ppad, 0), point_orig(Ppad, -1);
    std::vector<double> tS_s(R), sig_s(R);
    int64_t S = 0;
    for (int rs = 0; rs < R; rs++) {
        const int i = ray_orig[rs];
        const int np = ray_off[i + 1] - ray_off[i];
        tS_s[rs] = tS[i];
        sig_s[rs] = allSig[i];
        for (int k = 0; k < np; k++) {
            const size_t src = (size_t)i * m + k, dst = (size_t)sray_off[rs] + k;
            px[dst] = rayX[src];
            py[dst] = rayY[src];
            pz[dst] = rayZ[src];
            rayid[dst] = rs;
            point_orig[dst] = ray_off[i] + k;
            if (k < np - 1) {
                const size_t sg = (size_t)i * (m - 1) + k;
                dt[dst] = rayL[sg] * rayU[sg];  // rayl .* rayu is the first product of MCsub.jl:153 (host: no contraction)
                S++;
            }
        }
    }
    // evaluation tiles: consecutive whole rays with at most tile_pts points
    const int tile_pts = 4096;
    if (max_npts > tile_pts)
        return tg::fail(TONGA_ERR_CAPACITY, "tonga_create: a ray has more than 4096 points");
    std::vector<tg::Tile> tiles;
    {
        int r0 = 0;
        while (r0 < R) {
            int r1 = r0;
            while (r1 < R && sray_off[r1 + 1] - sray_off[r0] <= tile_pts) r1++;
            tiles.push_back({r0, r1, sray_off[r0], sray_off[r1]});
            r0 = r1;
        }
    }

    tonga_ctx *ctx = new tonga_ctx();
    ctx->device = device;
    ctx->prm = *params;
    ctx->m = m;
    ctx->R = R;
    ctx->P = P;
    ctx->S = S;
    ctx->Ppad = Ppad;
    ctx->max_npts = max_npts;
    ctx->h_ray_off = ray_off;
    ctx->h_ray_orig = ray_orig;
    ctx->h_point_orig = point_orig;
    ctx->Rp = Rp;
    ctx->n_tiles = (int)tiles.size();
    ctx->tile_pts = tile_pts;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete ctx;
        return tg::fail(TONGA_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    // MCsub.jl:179: sum(-log.(allSig*sqrt(2*pi)) * length(tS)), sequential on the host (a constant)
    {
        const double PI = 3.141592653589793;
        double lk = 0.0, lg = 0.0;
        for (int k = 0; k < R; k++) {
            lk += (-std::log(allSig[k] * std::sqrt(2 * PI))) * (double)R;
            lg += -std::log(allSig[k] * std::sqrt(2 * PI));
        }
        ctx->like_const = lk;
        ctx->sum_neglog = lg;
    }
    int rc = TONGA_OK;
    auto chk = [&](int r) { if (rc == TONGA_OK && r != TONGA_OK) rc = r; };
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return tg::fail(TONGA_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    chk(upload(&ctx->d_px, px, ctx->stream));
    chk(upload(&ctx->d_py, py, ctx->stream));
    chk(upload(&ctx->d_pz, pz, ctx->stream));
    {   // fl32 copies + the rigorous screening band: with u = 2^-24 and M = max |coordinate| (points and box),
        // |D32 - D| <= 11u*D32 + 7.5u*M^2 for any evaluation order; alpha = 16u and beta = 12u*M^2 leave a 1.4x margin.
        std::vector<float> xf(Ppad + TG_PT_SLACK, 0.f), yf(Ppad + TG_PT_SLACK, 0.f), zf(Ppad + TG_PT_SLACK, 0.f);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int64_t i = 0; i < Ppad; i++) {
            xf[i] = i < P ? (float)px[i] : TG_PAD_COORD; yf[i] = i < P ? (float)py[i] : TG_PAD_COORD; zf[i] = i < P ? (float)pz[i] : TG_PAD_COORD;
            if (i < P) {
                const double v[3] = {px[i], py[i], pz[i]};
                for (int a = 0; a < 3; a++) { if (v[a] < lo[a]) lo[a] = v[a]; if (v[a] > hi[a]) hi[a] = v[a]; }  // NaN never compares
            }
        }
        tg::set_screening_bounds(ctx, lo, hi);
        chk(upload(&ctx->d_pxf, xf, ctx->stream));
        chk(upload(&ctx->d_pyf, yf, ctx->stream));
        chk(upload(&ctx->d_pzf, zf, ctx->stream));
    }
    chk(upload(&ctx->d_dt, dt, ctx->stream));
    chk(upload(&ctx->d_rayid, rayid, ctx->stream));
    chk(upload(&ctx->d_ray_off, sray_off, ctx->stream));
    chk(upload(&ctx->d_ray_orig, ray_orig, ctx->stream));
    {
        std::vector<int32_t> ray_rank(R);
        for (int rs = 0; rs < R; rs++) ray_rank[ray_orig[rs]] = rs;
        chk(upload(&ctx->d_ray_rank, ray_rank, ctx->stream));
    }
    chk(upload(&ctx->d_point_orig, point_orig, ctx->stream));
    chk(upload(&ctx->d_tS, tS_s, ctx->stream));
    chk(upload(&ctx->d_sig, sig_s, ctx->stream));
    chk(upload(&ctx->d_tiles, tiles, ctx->stream));
    if (rc == TONGA_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = tg::fail(TONGA_ERR_CUDA, "tonga_create: upload failed");
    if (rc != TONGA_OK) {
        tonga_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return TONGA_OK;
}

void tg::set_screening_bounds(tonga_ctx *ctx, const double lo[3], const double hi[3]) {
    // fl32 screening band of the samplers: with u = 2^-24 and M = max |coordinate| (points and box),
    // |D32 - D| <= 11u*D32 + 7.5u*M^2 for any evaluation order; alpha = 16u and beta = 12u*M^2 leave a 1.4x margin.
    const tonga_params &pm = ctx->prm;
    const double blo[3] = {pm.xmin, pm.ymin, pm.zmin}, bhi[3] = {pm.xmax, pm.ymax, pm.zmax};
    double M = 0.0;
    for (int a = 0; a < 3; a++) {
        const bool pts = lo[a] <= hi[a];
        for (double v : {pts ? lo[a] : 0.0, pts ? hi[a] : 0.0, blo[a], bhi[a]})
            if (std::fabs(v) > M) M = std::fabs(v);
        ctx->cen[a] = 0.5 * (blo[a] + bhi[a]);
        if (!(std::fabs(ctx->cen[a]) < 1e300)) ctx->cen[a] = 0.0;
        const double mc = pts ? std::max(std::fabs(lo[a] - ctx->cen[a]), std::fabs(hi[a] - ctx->cen[a])) : 0.0;
        const double ma = pts ? std::max(std::fabs(lo[a]), std::fabs(hi[a])) : 0.0;
        ctx->mp_cen[a] = (float)(mc * 1.000001);
        ctx->mp_abs[a] = (float)(ma * 1.000001);
    }
    const double u = 5.9604644775390625e-08;
    ctx->tol_alpha = (float)(16.0 * u);
    ctx->tol_beta2 = (float)(2.0 * 12.0 * u * M * M * 1.0000002);
}

extern "C" void tonga_destroy(tonga_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_px); cudaFree(ctx->d_py); cudaFree(ctx->d_pz); cudaFree(ctx->d_dt);
    cudaFree(ctx->d_pxf); cudaFree(ctx->d_pyf); cudaFree(ctx->d_pzf);
    cudaFree(ctx->d_rayid); cudaFree(ctx->d_ray_off); cudaFree(ctx->d_ray_orig); cudaFree(ctx->d_ray_rank); cudaFree(ctx->d_point_orig);
    cudaFree(ctx->d_tS); cudaFree(ctx->d_sig);
    cudaFree(ctx->d_tiles);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int tonga_info(const tonga_ctx *ctx, int32_t *R, int64_t *P, int64_t *S, int64_t *Ppad) {
    if (!ctx) return tg::fail(TONGA_ERR_ARG, "tonga_info: ctx is NULL");
    if (R) *R = ctx->R;
    if (P) *P = ctx->P;
    if (S) *S = ctx->S;
    if (Ppad) *Ppad = ctx->Ppad;
    return TONGA_OK;
}

extern "C" int tonga_ray_offsets(const tonga_ctx *ctx, int32_t *ray_off) {
    if (!ctx || !ray_off) return tg::fail(TONGA_ERR_ARG, "tonga_ray_offsets: NULL argument");
    std::memcpy(ray_off, ctx->h_ray_off.data(), sizeof(int32_t) * (size_t)(ctx->R + 1));
    return TONGA_OK;
}

extern "C" int tonga_set_exact_only(tonga_ctx *ctx, int32_t exact_only) {
    if (!ctx) return tg::fail(TONGA_ERR_ARG, "tonga_set_exact_only: ctx is NULL");
    ctx->exact_only = exact_only ? 1 : 0;
    return TONGA_OK;
}

extern "C" int tonga_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return tg::fail(TONGA_ERR_ARG, "tonga_host_alloc: NULL");
    *ptr = nullptr;
    TG_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return TONGA_OK;
}
extern "C" int tonga_host_free(void *ptr) {
    if (ptr) TG_CUDA(cudaFreeHost(ptr));
    return TONGA_OK;
}

extern "C" int tonga_synchronize(tonga_ctx *ctx) {
    if (!ctx) return tg::fail(TONGA_ERR_ARG, "tonga_synchronize: ctx is NULL");
    TG_CUDA(cudaSetDevice(ctx->device));
    TG_CUDA(cudaStreamSynchronize(ctx->stream));
    return TONGA_OK;
}

// ------------------------------------------------------------------------------------------------ peak FMA rates
template <typename T>
__global__ void fma_chain_kernel(T *out, int iters) {
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T b = (T)0.999999, c = (T)1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);  // explicit: built with -fmad=false
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <typename T>
static int time_fma(tonga_ctx *ctx, double *tflops) {
    const int threads = 256, blocks = ctx->sm_count * 8, iters = 16384;
    const int rc0 = tg::ensure_scratch(ctx, sizeof(T) * (size_t)threads * blocks);
    if (rc0 != TONGA_OK) return rc0;
    cudaEvent_t e0, e1;
    TG_CUDA(cudaEventCreate(&e0));
    TG_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        TG_CUDA(cudaEventRecord(e0, ctx->stream));
        fma_chain_kernel<T><<<blocks, threads, 0, ctx->stream>>>((T *)ctx->d_scratch, iters);
        TG_CUDA(cudaEventRecord(e1, ctx->stream));
        TG_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        TG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * iters * (double)threads * blocks;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return TONGA_OK;
}

extern "C" int tonga_peak_flops(tonga_ctx *ctx, double *fp64_tflops, double *fp32_tflops) {
    if (!ctx) return tg::fail(TONGA_ERR_ARG, "tonga_peak_flops: ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    double a = 0, b = 0;
    int rc = time_fma<double>(ctx, &a);
    if (rc != TONGA_OK) return rc;
    rc = time_fma<float>(ctx, &b);
    if (rc != TONGA_OK) return rc;
    if (fp64_tflops) *fp64_tflops = a;
    if (fp32_tflops) *fp32_tflops = b;
    return TONGA_OK;
}
