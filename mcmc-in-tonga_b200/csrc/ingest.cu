// ingest.cu -- tonga_create_from_points: the ray preprocessing of load_data_Tonga.jl:66-69 and the flatten of tonga_create on
// the device.  The caller passes the raw NaN-padded m x R point matrices and the slowness at the
// points; segment lengths (rayl), mean slownesses (rayu), their product dt, the length-sorted CSR layout, the fl32 copies and
// the coordinate bound of the screening band are computed by two kernels.  Only R-sized arrays (points per ray, the stable
// length sort, tiles) are handled on the host.  The result is bit-identical to tonga_create fed with
// rayL = sqrt.(dx.^2 + dy.^2 + dz.^2), rayU = 0.5 .* (U[1:end-1] + U[2:end])  (tests: test_device_ingest_matches_host_flatten).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "tonga_internal.cuh"

namespace tg {

// warp per ray: points up to the first NaN of the ray's X column (MCsub.jl:312-316)
__global__ void __launch_bounds__(256) tg_ingest_count_kernel(int m, int R, const double *__restrict__ X, int32_t *__restrict__ npts) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ray >= R) return;
    const double *col = X + (size_t)ray * m;
    int np = m;
    for (int k0 = 0; k0 < m; k0 += 32) {
        const int k = k0 + lane;
        const bool isnan_ = (k < m) && (col[k] != col[k]);
        const unsigned b = __ballot_sync(0xffffffffu, isnan_);
        if (b) { np = k0 + __ffs(b) - 1; break; }
    }
    if (lane == 0) npts[ray] = np;
}

// warp per (length-sorted) ray: scatter into the device layout of tonga_create
__global__ void __launch_bounds__(256)
tg_ingest_scatter_kernel(int m, int R, const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                         const double *__restrict__ U, const int32_t *__restrict__ ray_orig, const int32_t *__restrict__ sray_off,
                         const int32_t *__restrict__ ray_off, double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz,
                         float *__restrict__ pxf, float *__restrict__ pyf, float *__restrict__ pzf, int32_t *__restrict__ rayid,
                         int32_t *__restrict__ point_orig, double *__restrict__ dt, unsigned long long *__restrict__ range_bits /* [6]: ordered bits of lo x,y,z, hi x,y,z */) {
    const int rs = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (rs >= R) return;
    const int i = ray_orig[rs];
    const int q0 = sray_off[rs], np = sray_off[rs + 1] - q0, o0 = ray_off[i];
    const double *x = X + (size_t)i * m, *y = Y + (size_t)i * m, *z = Z + (size_t)i * m, *u = U + (size_t)i * m;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = lane; k < np; k += 32) {
        const double xa = x[k], ya = y[k], za = z[k];
        px[q0 + k] = xa; py[q0 + k] = ya; pz[q0 + k] = za;
        pxf[q0 + k] = (float)xa; pyf[q0 + k] = (float)ya; pzf[q0 + k] = (float)za;
        rayid[q0 + k] = rs;
        point_orig[q0 + k] = o0 + k;
        lo[0] = fmin(lo[0], xa); lo[1] = fmin(lo[1], ya); lo[2] = fmin(lo[2], za);  // fmin / fmax ignore NaN
        hi[0] = fmax(hi[0], xa); hi[1] = fmax(hi[1], ya); hi[2] = fmax(hi[2], za);
        if (k < np - 1) {
            // load_data_Tonga.jl:66-69: rayl = sqrt(dx^2 + dy^2 + dz^2), rayu = 0.5 (U_k + U_k+1); MCsub.jl:153: rayl .* rayu
            const double dx = __dsub_rn(xa, x[k + 1]), dy = __dsub_rn(ya, y[k + 1]), dz = __dsub_rn(za, z[k + 1]);
            const double rayl = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            const double rayu = __dmul_rn(0.5, __dadd_rn(u[k], u[k + 1]));
            dt[q0 + k] = __dmul_rn(rayl, rayu);
        }
    }
    for (int a = 0; a < 3; a++) {
        for (int s = 16; s > 0; s >>= 1) { lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], s)); hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], s)); }
        if (lane == 0) {  // order-preserving map of doubles to unsigned integers
            const unsigned long long bl = (unsigned long long)__double_as_longlong(lo[a]), bh = (unsigned long long)__double_as_longlong(hi[a]);
            atomicMin(range_bits + a, bl ^ ((bl >> 63) ? ~0ull : 0x8000000000000000ull));
            atomicMax(range_bits + 3 + a, bh ^ ((bh >> 63) ? ~0ull : 0x8000000000000000ull));
        }
    }
}

__global__ void tg_ingest_pad_kernel(int64_t P, int64_t Ppad, double *px, double *py, double *pz, float *pxf, float *pyf, float *pzf,
                                     int32_t *rayid, int32_t *point_orig) {
    const int64_t i = P + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Ppad) return;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    px[i] = py[i] = pz[i] = qnan;
    pxf[i] = pyf[i] = pzf[i] = TG_PAD_COORD;
    rayid[i] = 0;
    point_orig[i] = -1;
}

}  // namespace tg

extern "C" int tonga_create_from_points(tonga_ctx **out, int32_t m, int32_t R, const double *rayX, const double *rayY, const double *rayZ,
                                        const double *U, const double *tS, const double *allSig, const tonga_params *params,
                                        int32_t device) {
    if (!out) return tg::fail(TONGA_ERR_ARG, "tonga_create_from_points: out is NULL");
    *out = nullptr;
    if (m < 1 || R < 1 || !params || !rayX || !rayY || !rayZ || !U || !tS || !allSig)
        return tg::fail(TONGA_ERR_ARG, "tonga_create_from_points: bad argument");
    if (params->interp_style != 1)
        return tg::fail(TONGA_ERR_ARG, "tonga_create_from_points: only interp_style == 1 (nearest) exists; style 2 is broken in the reference (MCsub.jl:332)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return tg::fail(TONGA_ERR_CUDA, "tonga_create_from_points: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return tg::fail(TONGA_ERR_ARG, "tonga_create_from_points: bad device index");
    TG_CUDA(cudaSetDevice(device));

    tonga_ctx *ctx = new tonga_ctx();
    struct Guard {
        tonga_ctx *&c;
        void *raw[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        bool armed = true;
        ~Guard() {
            for (void *p : raw) cudaFree(p);
            if (armed && c) { tonga_ctx *t = c; c = nullptr; tonga_destroy(t); }
        }
    } g{ctx};
    ctx->device = device;
    ctx->prm = *params;
    ctx->m = m;
    ctx->R = R;
    cudaDeviceProp prop;
    TG_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    TG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    cudaStream_t s = ctx->stream;

    // ---- raw matrices to the device
    const size_t mR = (size_t)m * R;
    double *dX, *dY, *dZ, *dU;
    int32_t *d_npts;
    TG_CUDA(cudaMalloc(&g.raw[0], 8 * mR)); dX = (double *)g.raw[0];
    TG_CUDA(cudaMalloc(&g.raw[1], 8 * mR)); dY = (double *)g.raw[1];
    TG_CUDA(cudaMalloc(&g.raw[2], 8 * mR)); dZ = (double *)g.raw[2];
    TG_CUDA(cudaMalloc(&g.raw[3], 8 * mR)); dU = (double *)g.raw[3];
    TG_CUDA(cudaMalloc(&g.raw[4], 4 * (size_t)R)); d_npts = (int32_t *)g.raw[4];
    TG_CUDA(cudaMemcpyAsync(dX, rayX, 8 * mR, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(dY, rayY, 8 * mR, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(dZ, rayZ, 8 * mR, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(dU, U, 8 * mR, cudaMemcpyHostToDevice, s));
    const unsigned wgrid = (unsigned)(((size_t)R * 32 + 255) / 256);
    tg::tg_ingest_count_kernel<<<wgrid, 256, 0, s>>>(m, R, dX, d_npts);
    TG_CUDA(cudaGetLastError());
    std::vector<int32_t> npts(R);
    TG_CUDA(cudaMemcpyAsync(npts.data(), d_npts, 4 * (size_t)R, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));

    // ---- R-sized layout on the host (as tonga_create): offsets, stable length sort, tiles
    std::vector<int32_t> ray_off(R + 1, 0);
    int max_npts = 0;
    for (int i = 0; i < R; i++) {
        ray_off[i + 1] = ray_off[i] + npts[i];
        max_npts = std::max(max_npts, (int)npts[i]);
    }
    const int64_t P = ray_off[R];
    if (P < 1) return tg::fail(TONGA_ERR_DATA, "tonga_create_from_points: no ray points");
    const int64_t Ppad = ((P + TG_PT_TILE - 1) / TG_PT_TILE) * TG_PT_TILE;
    std::vector<int32_t> ray_orig(R);
    for (int i = 0; i < R; i++) ray_orig[i] = i;
    std::stable_sort(ray_orig.begin(), ray_orig.end(), [&](int a, int b) { return npts[a] > npts[b]; });
    std::vector<int32_t> sray_off(R + 1, 0);
    int64_t S = 0;
    for (int rs = 0; rs < R; rs++) {
        const int np = npts[ray_orig[rs]];
        sray_off[rs + 1] = sray_off[rs] + np;
        S += np > 0 ? np - 1 : 0;
    }
    const int tile_pts = 4096;
    if (max_npts > tile_pts) return tg::fail(TONGA_ERR_CAPACITY, "tonga_create_from_points: a ray has more than 4096 points");
    std::vector<tg::Tile> tiles;
    for (int r0 = 0; r0 < R;) {
        int r1 = r0;
        while (r1 < R && sray_off[r1 + 1] - sray_off[r0] <= tile_pts) r1++;
        tiles.push_back({r0, r1, sray_off[r0], sray_off[r1]});
        r0 = r1;
    }
    ctx->P = P; ctx->S = S; ctx->Ppad = Ppad; ctx->max_npts = max_npts; ctx->Rp = (R + 1) & ~1;
    ctx->n_tiles = (int)tiles.size(); ctx->tile_pts = tile_pts;
    ctx->h_ray_off = ray_off; ctx->h_ray_orig = ray_orig;
    std::vector<double> tS_s(R), sig_s(R);
    {   // MCsub.jl:179 (a constant), as tonga_create
        const double PI = 3.141592653589793;
        double lk = 0.0, lg = 0.0;
        for (int k = 0; k < R; k++) {
            lk += (-std::log(allSig[k] * std::sqrt(2 * PI))) * (double)R;
            lg += -std::log(allSig[k] * std::sqrt(2 * PI));
        }
        ctx->like_const = lk;
        ctx->sum_neglog = lg;
        for (int rs = 0; rs < R; rs++) { tS_s[rs] = tS[ray_orig[rs]]; sig_s[rs] = allSig[ray_orig[rs]]; }
    }

    // ---- device layout
    auto up = [&](auto **dptr, const auto &h) -> int {
        TG_CUDA(cudaMalloc((void **)dptr, sizeof(h[0]) * h.size()));
        TG_CUDA(cudaMemcpyAsync(*dptr, h.data(), sizeof(h[0]) * h.size(), cudaMemcpyHostToDevice, s));
        return TONGA_OK;
    };
    std::vector<int32_t> ray_rank(R);
    for (int rs = 0; rs < R; rs++) ray_rank[ray_orig[rs]] = rs;
    int rc;
    if ((rc = up(&ctx->d_ray_rank, ray_rank)) || (rc = up(&ctx->d_ray_off, sray_off)) || (rc = up(&ctx->d_ray_orig, ray_orig)) || (rc = up(&ctx->d_tS, tS_s)) ||
        (rc = up(&ctx->d_sig, sig_s)) || (rc = up(&ctx->d_tiles, tiles)))
        return rc;
    int32_t *d_ray_off_orig = nullptr;  // offsets in the caller's ray order (for point_orig); freed below
    TG_CUDA(cudaMalloc((void **)&d_ray_off_orig, 4 * (size_t)(R + 1)));
    struct Free1 { void *p; ~Free1() { cudaFree(p); } } f1{d_ray_off_orig};
    TG_CUDA(cudaMemcpyAsync(d_ray_off_orig, ray_off.data(), 4 * (size_t)(R + 1), cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMalloc((void **)&ctx->d_px, 8 * (size_t)Ppad));
    TG_CUDA(cudaMalloc((void **)&ctx->d_py, 8 * (size_t)Ppad));
    TG_CUDA(cudaMalloc((void **)&ctx->d_pz, 8 * (size_t)Ppad));
    TG_CUDA(cudaMalloc((void **)&ctx->d_pxf, 4 * (size_t)(Ppad + TG_PT_SLACK)));
    TG_CUDA(cudaMemsetAsync(ctx->d_pxf, 0, 4 * (size_t)(Ppad + TG_PT_SLACK), s));
    TG_CUDA(cudaMalloc((void **)&ctx->d_pyf, 4 * (size_t)(Ppad + TG_PT_SLACK)));
    TG_CUDA(cudaMemsetAsync(ctx->d_pyf, 0, 4 * (size_t)(Ppad + TG_PT_SLACK), s));
    TG_CUDA(cudaMalloc((void **)&ctx->d_pzf, 4 * (size_t)(Ppad + TG_PT_SLACK)));
    TG_CUDA(cudaMemsetAsync(ctx->d_pzf, 0, 4 * (size_t)(Ppad + TG_PT_SLACK), s));
    TG_CUDA(cudaMalloc((void **)&ctx->d_rayid, 4 * (size_t)Ppad));
    TG_CUDA(cudaMalloc((void **)&ctx->d_point_orig, 4 * (size_t)Ppad));
    TG_CUDA(cudaMalloc((void **)&ctx->d_dt, 8 * (size_t)(Ppad + max_npts + 128)));  // zero slack for unconditional batched loads
    TG_CUDA(cudaMemsetAsync(ctx->d_dt, 0, 8 * (size_t)(Ppad + max_npts + 128), s));
    int rcs = tg::ensure_scratch(ctx, 48);
    if (rcs != TONGA_OK) return rcs;
    TG_CUDA(cudaMemsetAsync(ctx->d_scratch, 0xFF, 24, s));             // lo: atomicMin from all ones
    TG_CUDA(cudaMemsetAsync((char *)ctx->d_scratch + 24, 0, 24, s));  // hi: atomicMax from zero
    tg::tg_ingest_scatter_kernel<<<wgrid, 256, 0, s>>>(m, R, dX, dY, dZ, dU, ctx->d_ray_orig, ctx->d_ray_off, d_ray_off_orig, ctx->d_px, ctx->d_py,
                                                       ctx->d_pz, ctx->d_pxf, ctx->d_pyf, ctx->d_pzf, ctx->d_rayid, ctx->d_point_orig, ctx->d_dt,
                                                       (unsigned long long *)ctx->d_scratch);
    TG_CUDA(cudaGetLastError());
    if (Ppad > P) {
        tg::tg_ingest_pad_kernel<<<(unsigned)((Ppad - P + 255) / 256), 256, 0, s>>>(P, Ppad, ctx->d_px, ctx->d_py, ctx->d_pz, ctx->d_pxf, ctx->d_pyf,
                                                                                 ctx->d_pzf, ctx->d_rayid, ctx->d_point_orig);
        TG_CUDA(cudaGetLastError());
    }
    ctx->h_point_orig.resize(Ppad);
    unsigned long long rb[6];
    TG_CUDA(cudaMemcpyAsync(ctx->h_point_orig.data(), ctx->d_point_orig, 4 * (size_t)Ppad, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaMemcpyAsync(rb, ctx->d_scratch, 48, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    double lo[3], hi[3];
    for (int a = 0; a < 6; a++) {
        const unsigned long long b = rb[a] ^ ((rb[a] >> 63) ? 0x8000000000000000ull : ~0ull);
        double v;
        std::memcpy(&v, &b, 8);
        (a < 3 ? lo[a] : hi[a - 3]) = v;
    }
    tg::set_screening_bounds(ctx, lo, hi);  // the bands of tonga_create
    g.armed = false;
    *out = ctx;
    return TONGA_OK;
}
