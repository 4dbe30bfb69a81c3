// sampler.cu -- the device-resident RJ-MCMC proposal loop of TD_inversion_function.jl:70-302 for a batch of
// independent chains, with INCREMENTAL Voronoi maintenance instead of a full evaluate() per proposal.
//
// One CTA (128 threads) owns one chain for the whole launch.  The chain's state lives in shared memory:
//   owner8[Ppad]   nearest-nucleus index of every ray point (u8; 0x7F = none; bit 7 = pending tag)
//   tstar[R]       predicted t* per ray (model.ptS)
//   nuclei SoA     x, y, z, zeta [KC] (FP64) + fl32 x, y, z;  zp / zqp[129]: owner byte -> zeta, zeta/1000 under the proposal
//   mask32 / dirtyw  pending-overwrite bits per point / changed bits per 4-point word
// It is loaded once per launch with TMA bulk copies (cp.async.bulk + mbarrier), nIter iterations run without any
// global synchronisation, and it is written back once.  The ray geometry (shared by all chains) is streamed from L2
// with 128-bit loads.  The iteration itself is described in sampler_kernel.cuh; the state-independent random numbers of a
// launch are generated beforehand by tg_pregen_kernel.
// The same canonical t* / phi reductions are used by evaluate.cu, so incremental state == full evaluate bit for bit
// (tonga_chains_verify checks that on the device).
#include <algorithm>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "tonga_internal.cuh"

namespace tg {

constexpr int ST = TG_PHI_LANES;  // threads per chain CTA
constexpr int NW = TG_PHI_WARPS;  // warps per chain CTA
#define TG_SMALL_CHUNKS ((384 + TG_PHI_LANES - 1) / TG_PHI_LANES)  // instantiation for ray sets up to 384 rays (the 381-ray Tonga set)
#define TG_RESIDENT_MAX_CHUNKS 8  // the resident sampler keeps its rays' t* in registers: R <= ST * 8


}  // namespace tg

#include "sampler_kernel.cuh"
#include "wide_kernels.cuh"
#include "stream_cull.cuh"

namespace tg {

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

// CTA -> chain order of the resident sampler.  A chain's cost per iteration falls with its nCells (fewer, larger cells: more
// points change owner, more rays are re-integrated), and the launch lasts as long as its slowest SM.  CTAs are dealt to the
// SMs round-robin in launch order, so launching the chains sorted by K gives every SM one chain of every cost tier.
// (Measured and rejected, profiles/README.md round 2: a launch order from FEEDBACK -- the previous launch's per-chain busy cycles and
// CTA -> SM map, balanced on the host -- gains 1.5-2 % when a launch repeats the previous one exactly, and loses 0.5-9 % for the
// launches of a continuing run, whose cost the previous launch predicts worse than K does.)
// One CTA: counting sort on K (K <= 127 for the resident sampler), O(n).
__global__ void __launch_bounds__(1024) tg_order_kernel(int n, const int32_t *__restrict__ K, int32_t *__restrict__ perm) {
    __shared__ int hist[129], base[129];
    for (int i = threadIdx.x; i < 129; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) atomicAdd(&hist[min(max(K[c], 0), 128)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < 129; i++) { base[i] = acc; acc += hist[i]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) perm[atomicAdd(&base[min(max(K[c], 0), 128)], 1)] = c;  // order inside a bin is irrelevant
}

__global__ void tg_verify_kernel(int n, int64_t P0, int64_t P, int64_t Ppad, int R, int Rp, const int32_t *ray_orig, const uint8_t *own_a,
                                 const uint8_t *own_b, const uint16_t *own16_a, const uint16_t *own16_b, const double *ts_a /* [n][Rp] sorted rays */,
                                 const double *ts_b /* [n][R] caller's ray order */, const double *phi_a, const double *phi_b,
                                 const float *dc_a, const float *dc_b, float tol_alpha, float tol_beta2,
                                 unsigned long long *mism, double *maxd /* [2] as ordered uint64 bits */) {
    const int chain = blockIdx.y;
    unsigned long long local = 0;
    for (int64_t p = P0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x)  // [P0, P): a ray shard's own points
    {
        if (own16_a) local += own16_a[(size_t)chain * Ppad + p] != own16_b[(size_t)chain * Ppad + p];  // streamed sampler
        else local += own_a[(size_t)chain * Ppad + p] != own_b[(size_t)chain * Ppad + p];
        // the sampler's fl32 owner-distance cache must stay within a few ulp of the freshly computed distance
        const float ca = dc_a[(size_t)chain * Ppad + p], cb = dc_b[(size_t)chain * Ppad + p];
        local += !(ca < 0.0f || fabsf(ca - cb) <= tol_alpha * (ca + cb) + tol_beta2);  // stale marker, or inside the screening's error band
    }
    if (local) atomicAdd(mism, local);
    if (blockIdx.x == 0) {
        double dm = 0.0;
        for (int r = threadIdx.x; r < R; r += blockDim.x) {
            const double d = fabs(ts_a[(size_t)chain * Rp + r] - ts_b[(size_t)chain * R + ray_orig[r]]);
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

// ---- posterior rasterisation (plot_model_hist, MCsub.jl:753-825): CTA = (tile of 256 nodes, chain); the chain's kept models
// are visited in order, each staged in shared memory; one thread per node runs v_nearest (exact FP64) and keeps sum, sum^2.
__global__ void __launch_bounds__(256)
tg_raster_kernel(int n_nodes, const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z, int KC, int hist_cap,
                 const int32_t *__restrict__ n_hist, const int32_t *__restrict__ hist_K, const double *__restrict__ hist_cells,
                 double *__restrict__ part /* [nChains][2][n_nodes] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_c = reinterpret_cast<double *>(smem_raw);  // [4][KC]
    const int chain = blockIdx.y;
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = node < n_nodes;
    const double x = on ? X[node] : 0.0, y = on ? Y[node] : 0.0, z = on ? Z[node] : 0.0;
    double s1 = 0.0, s2 = 0.0;
    const int nh = min(n_hist[chain], hist_cap);
    for (int j = 0; j < nh; j++) {
        const size_t h = (size_t)chain * hist_cap + j;
        const double *c = hist_cells + h * 4 * KC;
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * KC; i += blockDim.x) s_c[i] = c[i];
        __syncthreads();
        const int K = hist_K[h];
        double best = 1e9, v = 0.0;
        for (int i = 0; i < K; i++) {
            const double d = dist2_exact(s_c[i], s_c[KC + i], s_c[2 * KC + i], x, y, z);
            if (d < best) { best = d; v = s_c[3 * KC + i]; }
        }
        s1 = __dadd_rn(s1, v);
        s2 = __dadd_rn(s2, __dmul_rn(v, v));
    }
    if (on) {
        part[((size_t)chain * 2 + 0) * n_nodes + node] = s1;
        part[((size_t)chain * 2 + 1) * n_nodes + node] = s2;
    }
}

__global__ void tg_raster_reduce_kernel(int n_nodes, int nChains, const double *__restrict__ part, double *__restrict__ sum, double *__restrict__ sumsq) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    double a = 0.0, b = 0.0;
    for (int c = 0; c < nChains; c++) {  // fixed chain order -> deterministic
        a = __dadd_rn(a, part[((size_t)c * 2 + 0) * n_nodes + node]);
        b = __dadd_rn(b, part[((size_t)c * 2 + 1) * n_nodes + node]);
    }
    sum[node] = a;
    sumsq[node] = b;
}

// ---- parallel tempering (extension): one even / odd swap sweep, one thread per ladder of T replicas (global replica order).
// E = phi/2 + R log(noise); pairs adjacent in temperature order; decisions from Philox4x32-10(seed; step, ladder, rung).
constexpr int TEMPER_MAX_T = 256;
__global__ void tg_temper_swap_kernel(int n_lad, int T, int R, const double *__restrict__ phi, const double *__restrict__ noise, double *__restrict__ beta,
                                      long long step, unsigned long long seed, unsigned long long *__restrict__ stat) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lad) return;
    double b[TEMPER_MAX_T], e[TEMPER_MAX_T];
    int order[TEMPER_MAX_T];
    for (int t = 0; t < T; t++) {
        b[t] = beta[(size_t)l * T + t];
        e[t] = 0.5 * phi[(size_t)l * T + t] + (double)R * log(noise[(size_t)l * T + t]);
    }
    for (int t = 0; t < T; t++) {  // temperature order: coldest (largest beta) first, stable
        int rank = 0;
        for (int s = 0; s < T; s++) rank += (b[s] > b[t]) || (b[s] == b[t] && s < t);
        order[rank] = t;
    }
    const Philox philox{(uint32_t)seed, (uint32_t)(seed >> 32)};
    unsigned long long acc = 0, att = 0;
    for (int t = (int)(step & 1); t + 1 < T; t += 2) {
        const int i = order[t], j = order[t + 1];
        uint32_t w[4];
        philox((uint32_t)step, (uint32_t)((unsigned long long)step >> 32), (uint32_t)l, (uint32_t)t | 0x40000000u, w);  // purpose tag: swap stream
        const double u = u53_open(w[0], w[1]);
        const double loga = (b[i] - b[j]) * (e[i] - e[j]);
        att++;
        if (log(u) < (loga < 0.0 ? loga : 0.0)) {
            const double tmp = b[i]; b[i] = b[j]; b[j] = tmp;
            acc++;
        }
    }
    for (int t = 0; t < T; t++) beta[(size_t)l * T + t] = b[t];
    if (stat) { atomicAdd(&stat[0], acc); atomicAdd(&stat[1], att); }
}

// evaluate's t* (caller's ray order) -> chain state rows (sorted ray order, padded to Rp)
__global__ void tg_copy_tstar_kernel(int n, int R, int Rp, const int32_t *ray_orig, const double *src /* [n][R] */, double *dst /* [n][Rp] */) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < (size_t)n * Rp) {
        const size_t c = i / Rp, r = i % Rp;
        dst[i] = r < (size_t)R ? src[c * R + ray_orig[r]] : 0.0;
    }
}

// chain state -> the caller's layout, packed on the device so that tonga_chains_get_state is a few straight device-to-host copies:
// cells [n][4][KCsrc] -> [n][4][KCdst] with zeros behind the K valid nuclei (both directions: get_state and set_models); t* rows (sorted ray order, padded) -> [n][R] in the caller's order
__global__ void tg_pack_cells_kernel(int n, int KCsrc, int KCdst, const int32_t *K, const double *src, double *dst) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < (size_t)n * 4 * KCdst) {
        const size_t row = i / KCdst, k = i % KCdst;  // row = chain * 4 + axis
        dst[i] = ((int)k < K[row >> 2] && (int)k < KCsrc) ? src[row * KCsrc + k] : 0.0;
    }
}
__global__ void tg_unsort_tstar_kernel(int n, int R, int Rp, const int32_t *ray_orig, const double *src /* [n][Rp] */, double *dst /* [n][R] */) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < (size_t)n * R) {
        const size_t c = i / R, rs = i % R;
        dst[c * R + ray_orig[rs]] = src[c * Rp + rs];
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
    bool host_history = false;  // history arrays live in mapped page-locked host memory: the kernels store the kept models straight
                                // into host memory during the run (posted PCIe writes), no device-to-host copy afterwards
    bool wide = false;      // wide or streamed sampler (wide_kernels.cuh): lock-step launches per iteration, candidate models in global memory
    bool streamed = false;  // streamed sampler: per-point chain state (u16 owner, fl32 owner distance) in HBM, updated incrementally
    int exact_only = 0;
    bool no_order = false;  // TONGA_NO_ORDER=1 in the environment: identity launch order (A/B measurements)
    long long *d_prof = nullptr;  // optional per-phase cycle counters (tonga_chains_profile)
    size_t smem = 0;
    // device state
    int32_t *d_K = nullptr;
    double *d_cells = nullptr, *d_phi = nullptr, *d_noise = nullptr, *d_beta = nullptr, *d_tstar = nullptr;
    uint8_t *d_owner = nullptr;
    float *d_dcache = nullptr, *d_dcache_tmp = nullptr;  // [n][Ppad] fl32 squared distance of every point to its owner
    long long *d_counts = nullptr;
    int32_t *d_pending = nullptr;
    int32_t *d_perm = nullptr;  // resident sampler: launch order
    int32_t *d_n_hist = nullptr;
    long long *d_model_num = nullptr;
    int32_t *d_hist_K = nullptr;
    double *d_hist_cells = nullptr, *d_hist_phi = nullptr, *d_hist_ptS = nullptr;
    long long *d_hist_iter = nullptr;
    int32_t *d_hist_action = nullptr, *d_hist_accept = nullptr, *d_hist_next = nullptr;
    // wide sampler: candidate models
    int32_t *d_Kc = nullptr;
    double *d_cells_c = nullptr;
    tg::Prop *d_props = nullptr;
    float *d_cells_cf = nullptr;
    // streamed sampler
    uint16_t *d_owner16 = nullptr;  // [n][Ppad]
    double *d_tstar_c = nullptr;    // [n][Rp]
    double *d_term_c = nullptr;     // [n][Rp]
    int32_t *d_accept = nullptr;    // [n]
    int32_t *d_active = nullptr;    // [n + 1]: active chain list, then its length
    size_t stream_smem = 0;
    uint8_t *d_tile_changed = nullptr;  // [n][n_stiles]
    tg::Tile *d_stiles = nullptr;   // the streamed sampler's own tiles (whole rays, <= stile_pts points)
    int n_stiles = 0, stile_pts = 0;
    std::vector<tg::Tile> h_stiles;
    // ray sharding of the streamed sampler (tonga_chains_shard_init / _connect)
    int sh_rank = 0, sh_world = 1, sh_tile0 = 0, sh_tile1 = 0;
    bool sh_connected = false;
    unsigned char *d_xch = nullptr;      // own exchange block: flags | err | term[2][n][Rp] | tsc[2][n][Rp]
    size_t xch_bytes = 0, xch_hdr = 0;
    unsigned char *sh_peer[tg::TG_MAX_SHARDS] = {};
    unsigned long long sh_seq = 0;       // exchanges published so far
    int mb_cap = 0;
    size_t mb_cnt_bytes = 0, mb_sec = 0;
    // streamed sampler with spatial culling (stream_cull.cuh)
    bool culled = false;
    int s2_maxn = 0;
    size_t s2_smem = 0;
    float4 *d_sub = nullptr;         // bounding spheres of the 32-point runs
    float4 *d_raysph = nullptr;      // [R] bounding spheres of the rays
    int32_t *d_sub_off = nullptr;    // [R+1]
    float *d_dmax = nullptr;         // [n][Rp]
    double *d_term = nullptr;        // [n][Rp]
    int32_t *d_cand = nullptr, *d_ncand = nullptr, *d_dirty = nullptr, *d_ndirty = nullptr, *d_work_off = nullptr;
    int32_t *d_clist = nullptr, *d_nclist = nullptr, *d_ncommit = nullptr, *d_work_off2 = nullptr;
    // scratch
    double *d_ptS_tmp = nullptr;  // [n][R]  (wide sampler: t* of the candidates)
    double *d_phi_tmp = nullptr;
    uint8_t *d_owner_tmp = nullptr;
    unsigned long long *d_mism = nullptr;
    unsigned long long *d_swapstat = nullptr;  // [2] accepted / attempted tempering swaps
    double *d_maxd = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
};

static inline size_t tg_nz(size_t b) { return b ? b : 1; }
#define TG_ALLOC(ptr, bytes) TG_CUDA(cudaMalloc((void **)&(ptr), tg_nz(bytes)))

static void tonga_chains_destroy_unlocked(tonga_chains *ch);

extern "C" int tonga_chains_create(tonga_ctx *ctx, tonga_chains **out, int32_t nChains, int64_t chain_id0, uint64_t seed,
                                   int32_t hist_cap) {
    return tonga_chains_create_ex(ctx, out, nChains, chain_id0, seed, hist_cap, TONGA_SAMPLER_AUTO);
}

extern "C" int tonga_chains_create_ex(tonga_ctx *ctx, tonga_chains **out, int32_t nChains, int64_t chain_id0, uint64_t seed,
                                      int32_t hist_cap, int32_t sampler) {
    const bool host_history = (sampler & TONGA_HISTORY_ON_HOST) != 0;
    sampler &= ~TONGA_HISTORY_ON_HOST;
    if (!ctx || !out || nChains < 1 || hist_cap < 0 || sampler < TONGA_SAMPLER_AUTO || sampler > TONGA_SAMPLER_STREAMED)
        return tg::fail(TONGA_ERR_ARG, "tonga_chains_create: bad argument");
    *out = nullptr;
    const tonga_params &pm = ctx->prm;
    if (pm.max_cells < 1 || pm.min_cells < 1 || pm.min_cells > pm.max_cells)
        return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: need 1 <= min_cells <= max_cells");
    if (ctx->R < 1 || ctx->P < 1) return tg::fail(TONGA_ERR_ARG, "tonga_chains_create: empty ray set");
    const int KC0 = ((pm.max_cells + 7) / 8) * 8;
    const size_t smem_res = tg::smem_layout((int)ctx->Ppad, ctx->Rp, KC0).total;
    const bool fits = pm.max_cells <= TG_MAX_K_U8 && smem_res <= ctx->smem_optin && ctx->R <= tg::ST * TG_RESIDENT_MAX_CHUNKS && ctx->Ppad < (1 << 18) &&
                      ctx->max_npts < (1 << 13);
    if (sampler == TONGA_SAMPLER_RESIDENT && !fits) {
        if (pm.max_cells > TG_MAX_K_U8)
            return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: the resident sampler needs max_cells <= 126 (u8 owner state)");
        return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: per-chain state (" + std::to_string(smem_res) +
                                                " B) exceeds shared memory, or more than " + std::to_string(tg::ST * TG_RESIDENT_MAX_CHUNKS) +
                                                " rays; the smem-resident sampler handles ray sets up to ~200k points");
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    // streamed sampler: 6 B per ray point and chain of state; AUTO uses it when that fits in 60 % of the free memory
    int stile_pts = 8192;  // tile of the streamed sampler: twice the evaluate tile (half the CTA prologues; 34 KB of shared memory)
    if (const char *e = std::getenv("TONGA_STREAM_TILE")) stile_pts = std::max(1024, std::min(15360, std::atoi(e) & ~31));
    stile_pts = std::max(stile_pts, ctx->max_npts);
    const size_t stream_smem = ((size_t)stile_pts + 8) * 4 + ((size_t)stile_pts + 8) / 8 + 4 + 16;
    const bool stream_ok = pm.max_cells <= 65534 && stream_smem <= ctx->smem_optin;
    if (sampler == TONGA_SAMPLER_STREAMED && !stream_ok)
        return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: the streamed sampler needs max_cells <= 65534 and its tile state in shared memory");
    bool streamed = sampler == TONGA_SAMPLER_STREAMED;
    if (sampler == TONGA_SAMPLER_AUTO && !fits && stream_ok) {
        size_t fr = 0, tot = 0;
        TG_CUDA(cudaMemGetInfo(&fr, &tot));
        streamed = (double)nChains * 6.0 * (double)ctx->Ppad < 0.6 * (double)fr;
    }
    const bool wide = streamed || (sampler == TONGA_SAMPLER_WIDE) || !fits;
    if (wide && nChains > 65535) return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: the wide / streamed samplers run at most 65535 chains per batch");
    std::vector<tg::Tile> stiles;
    if (streamed) {  // consecutive whole (length-sorted) rays with at most stile_pts points
        const int R = ctx->R;
        std::vector<int32_t> so(R + 1, 0);
        for (int rs = 0; rs < R; rs++) so[rs + 1] = so[rs] + (ctx->h_ray_off[ctx->h_ray_orig[rs] + 1] - ctx->h_ray_off[ctx->h_ray_orig[rs]]);
        for (int r0 = 0; r0 < R;) {
            int r1 = r0;
            while (r1 < R && so[r1 + 1] - so[r0] <= stile_pts) r1++;
            stiles.push_back({r0, r1, so[r0], so[r1]});
            r0 = r1;
        }
    }
    if (streamed && (double)stiles.size() * (double)nChains > 2147483647.0)
        return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: tiles x chains exceeds the launch grid of the streamed sampler");
    tonga_chains *ch = new tonga_chains();
    struct Guard {  // frees a half-built batch when an allocation below fails (TG_ALLOC / TG_CUDA return early)
        tonga_chains *&c;
        bool armed = true;
        ~Guard() { if (armed && c) { tonga_chains *t = c; c = nullptr; tonga_chains_destroy_unlocked(t); } }
    } guard{ch};
    ch->ctx = ctx;
    ch->n = nChains;
    ch->KC = ((pm.max_cells + 7) / 8) * 8;
    ch->Rp = ctx->Rp;
    ch->hist_cap = hist_cap;
    ch->chain_id0 = chain_id0;
    ch->seed = seed;
    ch->wide = wide;
    ch->host_history = host_history;
    ch->no_order = std::getenv("TONGA_NO_ORDER") != nullptr;
    ch->streamed = streamed;
    ch->stream_smem = stream_smem;
    ch->stile_pts = stile_pts;
    ch->n_stiles = (int)stiles.size();
    ch->h_stiles = stiles;
    ch->sh_tile1 = ch->n_stiles;
    ch->smem = wide ? 0 : smem_res;
    const size_t n = (size_t)nChains, KC = (size_t)ch->KC, R = (size_t)ctx->R, Rp = (size_t)ch->Rp, Pp = (size_t)ctx->Ppad, H = (size_t)hist_cap;
    TG_ALLOC(ch->d_K, 4 * n);
    TG_ALLOC(ch->d_cells, 8 * n * 4 * KC);
    TG_ALLOC(ch->d_phi, 8 * n);
    TG_ALLOC(ch->d_noise, 8 * n);
    TG_ALLOC(ch->d_beta, 8 * n);
    TG_ALLOC(ch->d_tstar, 8 * n * Rp);
    if (!wide) {
        TG_ALLOC(ch->d_owner, n * Pp);
        TG_ALLOC(ch->d_dcache, 4 * (n * Pp + TG_PT_SLACK));
        TG_CUDA(cudaMemsetAsync(ch->d_dcache, 0, 4 * (n * Pp + TG_PT_SLACK), ctx->stream));
        TG_ALLOC(ch->d_dcache_tmp, 4 * n * Pp);
        TG_ALLOC(ch->d_owner_tmp, n * Pp);
    } else {
        TG_ALLOC(ch->d_Kc, 4 * n);
        TG_ALLOC(ch->d_cells_c, 8 * n * 4 * KC);
        TG_ALLOC(ch->d_props, sizeof(tg::Prop) * n);
        TG_CUDA(cudaMemsetAsync(ch->d_cells_c, 0, 8 * n * 4 * KC, ctx->stream));
    }
    if (streamed) {
        TG_ALLOC(ch->d_owner16, 2 * n * Pp);
        TG_ALLOC(ch->d_dcache, 4 * n * Pp);
        TG_ALLOC(ch->d_tstar_c, 8 * n * Rp);
        TG_ALLOC(ch->d_term_c, 8 * n * Rp);
        TG_ALLOC(ch->d_accept, 4 * n);
        TG_ALLOC(ch->d_stiles, sizeof(tg::Tile) * stiles.size());
        TG_ALLOC(ch->d_tile_changed, n * stiles.size());
        TG_CUDA(cudaMemcpyAsync(ch->d_stiles, stiles.data(), sizeof(tg::Tile) * stiles.size(), cudaMemcpyHostToDevice, ctx->stream));
        TG_CUDA(cudaStreamSynchronize(ctx->stream));  // `stiles` is a local
        TG_ALLOC(ch->d_active, 4 * (n + 1));
        TG_ALLOC(ch->d_cells_cf, 4 * n * 3 * KC);
        TG_CUDA(cudaMemsetAsync(ch->d_owner16, 0xFF, 2 * n * Pp, ctx->stream));  // the padded tail stays "none"
        TG_CUDA(cudaMemsetAsync(ch->d_accept, 0, 4 * n, ctx->stream));
        if (stream_smem > 48 * 1024) {
            TG_CUDA(cudaFuncSetAttribute(tg::tg_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem));
            TG_CUDA(cudaFuncSetAttribute(tg::tg_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem));
        }
        // spatial culling (stream_cull.cuh): one warp per candidate ray; needs a ray's owners in the warp's shared memory
        ch->s2_maxn = ((ctx->max_npts + 8 + 31) / 32) * 32;
        ch->s2_smem = (size_t)tg::S2_WARPS * ((tg::s2_warp_smem(ch->s2_maxn) + 15) & ~(size_t)15);
        const char *ce = std::getenv("TONGA_STREAM_CULL");
        ch->culled = !(ce && ce[0] == '0') && ch->s2_smem <= std::min<size_t>(ctx->smem_optin, 56 * 1024);
        if (ch->culled) {
            const int R = ctx->R;
            std::vector<int32_t> so(R + 1, 0);  // 32-point runs per (sorted) ray
            for (int rs = 0; rs < R; rs++) {
                const int npt = ctx->h_ray_off[ctx->h_ray_orig[rs] + 1] - ctx->h_ray_off[ctx->h_ray_orig[rs]];
                so[rs + 1] = so[rs] + (npt + 31) / 32;
            }
            TG_ALLOC(ch->d_sub_off, 4 * (size_t)(R + 1));
            TG_ALLOC(ch->d_sub, sizeof(float4) * (size_t)std::max(so[R], 1));
            TG_ALLOC(ch->d_raysph, sizeof(float4) * (size_t)R);
            TG_CUDA(cudaMemcpyAsync(ch->d_sub_off, so.data(), 4 * (size_t)(R + 1), cudaMemcpyHostToDevice, ctx->stream));
            tg::tg_sub_spheres_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(R, ctx->d_ray_off, ch->d_sub_off, ctx->d_px, ctx->d_py, ctx->d_pz, ch->d_sub, ch->d_raysph);
            TG_CUDA(cudaGetLastError());
            TG_CUDA(cudaStreamSynchronize(ctx->stream));  // `so` is a local
            TG_ALLOC(ch->d_dmax, 4 * n * Rp);
            TG_ALLOC(ch->d_term, 8 * n * Rp);
            TG_ALLOC(ch->d_cand, 4 * n * R);
            TG_ALLOC(ch->d_dirty, 4 * n * R);
            TG_ALLOC(ch->d_clist, 4 * n * R);
            TG_ALLOC(ch->d_nclist, 4 * n);
            TG_ALLOC(ch->d_ncommit, 4 * n);
            TG_ALLOC(ch->d_work_off2, 4 * (n + 1));
            TG_CUDA(cudaMemsetAsync(ch->d_nclist, 0, 4 * n, ctx->stream));
            TG_CUDA(cudaMemsetAsync(ch->d_ncommit, 0, 4 * n, ctx->stream));
            TG_ALLOC(ch->d_ncand, 4 * n);
            TG_ALLOC(ch->d_ndirty, 4 * n);
            TG_ALLOC(ch->d_work_off, 4 * (n + 1));
            TG_CUDA(cudaMemsetAsync(ch->d_ncand, 0, 4 * n, ctx->stream));
            TG_CUDA(cudaMemsetAsync(ch->d_ndirty, 0, 4 * n, ctx->stream));
            if (ch->s2_smem > 48 * 1024) {
                TG_CUDA(cudaFuncSetAttribute(tg::tg_stream2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->s2_smem));
                TG_CUDA(cudaFuncSetAttribute(tg::tg_stream2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->s2_smem));
            }
        }
    }
    TG_ALLOC(ch->d_counts, 8 * n * 15);
    TG_ALLOC(ch->d_pending, 4 * n);
    TG_ALLOC(ch->d_perm, 4 * n);
    TG_ALLOC(ch->d_n_hist, 4 * n);
    TG_ALLOC(ch->d_model_num, 8 * n);
#define TG_HIST_ALLOC(ptr, bytes)                                                                                              \
    do {                                                                                                                       \
        if (host_history) TG_CUDA(cudaHostAlloc((void **)&(ptr), tg_nz(bytes), cudaHostAllocMapped | cudaHostAllocPortable));   \
        else TG_ALLOC(ptr, bytes);                                                                                             \
    } while (0)
    TG_HIST_ALLOC(ch->d_hist_K, 4 * n * H);
    TG_HIST_ALLOC(ch->d_hist_cells, 8 * n * H * 4 * KC);
    TG_HIST_ALLOC(ch->d_hist_phi, 8 * n * H);
    TG_HIST_ALLOC(ch->d_hist_ptS, 8 * n * H * R);
    TG_HIST_ALLOC(ch->d_hist_iter, 8 * n * H);
    TG_HIST_ALLOC(ch->d_hist_action, 4 * n * H);
    TG_HIST_ALLOC(ch->d_hist_accept, 4 * n * H);
    TG_HIST_ALLOC(ch->d_hist_next, 4 * n * H);
#undef TG_HIST_ALLOC
    TG_ALLOC(ch->d_ptS_tmp, 8 * n * R);
    TG_ALLOC(ch->d_phi_tmp, 8 * n);
    TG_ALLOC(ch->d_mism, 8);
    TG_ALLOC(ch->d_swapstat, 16);
    TG_CUDA(cudaMemsetAsync(ch->d_swapstat, 0, 16, ctx->stream));
    TG_ALLOC(ch->d_maxd, 16);
    cudaStream_t s = ctx->stream;
    TG_CUDA(cudaMemsetAsync(ch->d_counts, 0, 8 * n * 15, s));
    TG_CUDA(cudaMemsetAsync(ch->d_pending, 0xFF, 4 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_n_hist, 0, 4 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_model_num, 0, 8 * n, s));
    TG_CUDA(cudaMemsetAsync(ch->d_cells, 0, 8 * n * 4 * KC, s));
    // unused history slots read as zeros (get_history copies whole arrays)
    TG_CUDA(cudaMemsetAsync(ch->d_hist_K, 0, tg_nz(4 * n * H), s));
    TG_CUDA(cudaMemsetAsync(ch->d_hist_cells, 0, tg_nz(8 * n * H * 4 * KC), s));
    TG_CUDA(cudaMemsetAsync(ch->d_hist_phi, 0, tg_nz(8 * n * H), s));
    TG_CUDA(cudaMemsetAsync(ch->d_hist_ptS, 0, tg_nz(8 * n * H * R), s));
    TG_CUDA(cudaMemsetAsync(ch->d_hist_iter, 0, tg_nz(8 * n * H), s));
    TG_CUDA(cudaMemsetAsync(ch->d_hist_action, 0, tg_nz(4 * n * H), s));
    TG_CUDA(cudaMemsetAsync(ch->d_hist_accept, 0, tg_nz(4 * n * H), s));
    TG_CUDA(cudaMemsetAsync(ch->d_hist_next, 0, tg_nz(4 * n * H), s));
    {
        std::vector<double> ones(n, 1.0);
        TG_CUDA(cudaMemcpyAsync(ch->d_beta, ones.data(), 8 * n, cudaMemcpyHostToDevice, s));
        TG_CUDA(cudaMemcpyAsync(ch->d_noise, ones.data(), 8 * n, cudaMemcpyHostToDevice, s));
        TG_CUDA(cudaStreamSynchronize(s));
    }
    if (!wide) {
        TG_CUDA(cudaFuncSetAttribute(tg::tg_sampler_kernel<TG_SMALL_CHUNKS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->smem));
        TG_CUDA(cudaFuncSetAttribute(tg::tg_sampler_kernel<TG_SMALL_CHUNKS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->smem));
        TG_CUDA(cudaFuncSetAttribute(tg::tg_sampler_kernel<TG_SMALL_CHUNKS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->smem));
        TG_CUDA(cudaFuncSetAttribute(tg::tg_sampler_kernel<TG_RESIDENT_MAX_CHUNKS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->smem));
        TG_CUDA(cudaFuncSetAttribute(tg::tg_sampler_kernel<TG_RESIDENT_MAX_CHUNKS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->smem));
        TG_CUDA(cudaFuncSetAttribute(tg::tg_sampler_kernel<TG_RESIDENT_MAX_CHUNKS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch->smem));
    }
    if (8 * 4 * KC > 48 * 1024) {
        if (8 * 4 * KC > ctx->smem_optin) return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_create: max_cells too large for shared memory");
        TG_CUDA(cudaFuncSetAttribute(tg::tg_raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * 4 * KC)));
    }
    TG_CUDA(cudaEventCreate(&ch->ev0));
    TG_CUDA(cudaEventCreate(&ch->ev1));
    guard.armed = false;
    *out = ch;
    return TONGA_OK;
}

static void tonga_chains_destroy_unlocked(tonga_chains *ch);

extern "C" void tonga_chains_destroy(tonga_chains *ch) {
    if (!ch) return;
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    tonga_chains_destroy_unlocked(ch);
}

static void tonga_chains_destroy_unlocked(tonga_chains *ch) {
    if (!ch) return;
    cudaSetDevice(ch->ctx->device);
    cudaStreamSynchronize(ch->ctx->stream);
    void *ptrs[] = {ch->d_K, ch->d_cells, ch->d_phi, ch->d_noise, ch->d_beta, ch->d_tstar, ch->d_owner, ch->d_dcache, ch->d_dcache_tmp, ch->d_counts, ch->d_pending,
                    ch->d_n_hist, ch->d_model_num, ch->d_ptS_tmp, ch->d_phi_tmp, ch->d_owner_tmp, ch->d_mism,
                    ch->d_maxd, ch->d_swapstat, ch->d_perm, ch->d_Kc, ch->d_cells_c, ch->d_props, ch->d_owner16, ch->d_tstar_c, ch->d_accept, ch->d_cells_cf, ch->d_term_c, ch->d_active, ch->d_stiles, ch->d_tile_changed};
    for (void *p : ptrs) cudaFree(p);
    void *hist[] = {ch->d_hist_K, ch->d_hist_cells, ch->d_hist_phi, ch->d_hist_ptS, ch->d_hist_iter, ch->d_hist_action, ch->d_hist_accept, ch->d_hist_next};
    for (void *p : hist) {
        if (ch->host_history) cudaFreeHost(p);
        else cudaFree(p);
    }
    if (ch->d_prof) cudaFree(ch->d_prof);
    if (ch->d_xch) cudaFree(ch->d_xch);
    void *cull[] = {ch->d_raysph, ch->d_sub, ch->d_sub_off, ch->d_dmax, ch->d_term, ch->d_cand, ch->d_ncand, ch->d_dirty, ch->d_ndirty, ch->d_work_off, ch->d_clist, ch->d_nclist, ch->d_ncommit, ch->d_work_off2};
    for (void *p : cull) cudaFree(p);
    if (ch->ev0) cudaEventDestroy(ch->ev0);
    if (ch->ev1) cudaEventDestroy(ch->ev1);
    delete ch;
}

// full evaluate of the current device-resident models -> owners, t*, phi of the chain state
static int establish_state(tonga_chains *ch) {
    tonga_ctx *ctx = ch->ctx;
    int rc = tg::launch_evaluate(ctx, ch->n, ch->KC, ch->d_K, ch->d_cells, ch->d_noise, ch->d_ptS_tmp, ch->d_phi, nullptr, ch->d_owner, ch->d_dcache,
                                 /*force_geometry=*/true, ch->d_owner16);  // debug_prior: phi = 1 but the chain state (owners, t*) is still well defined
    if (rc != TONGA_OK) return rc;
    const size_t tot = (size_t)ch->n * ch->Rp;
    tg::tg_copy_tstar_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ch->n, ctx->R, ch->Rp, ctx->d_ray_orig, ch->d_ptS_tmp, ch->d_tstar);
    TG_CUDA(cudaGetLastError());
    if (ch->culled) {  // culling state of the streamed sampler: dmax per ray, misfit terms; tstar_c = t*, term_c = term
        tg::tg_stream_init_kernel<<<dim3((unsigned)((ctx->R + 7) / 8), (unsigned)ch->n), 256, 0, ctx->stream>>>(ctx->R, ch->Rp, ctx->Ppad, ctx->d_ray_off, ch->d_dcache, ch->d_tstar,
                                                                                                      ctx->d_tS, ctx->d_sig, ch->d_noise, ch->d_dmax, ch->d_term,
                                                                                                      ch->d_tstar_c, ch->d_term_c);
        TG_CUDA(cudaGetLastError());
    }
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
    // the caller's [n][4][Kcap] block goes to the device as it is (at PCIe speed from page-locked memory) and is brought into the chains'
    // [n][4][KC] layout there, zeros behind the K valid nuclei
    cudaStream_t s = ctx->stream;
    TG_CUDA(cudaStreamSynchronize(s));  // (the scratch buffer may still be read by an earlier call)
    {
        int rc0 = tg::ensure_scratch(ctx, 8 * n * 4 * (size_t)Kcap);
        if (rc0 != TONGA_OK) return rc0;
    }
    TG_CUDA(cudaMemcpyAsync(ch->d_K, K, 4 * n, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(ctx->d_scratch, cells, 8 * n * 4 * (size_t)Kcap, cudaMemcpyHostToDevice, s));
    {
        const size_t tot = n * 4 * KC;
        tg::tg_pack_cells_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(ch->n, Kcap, ch->KC, ch->d_K, (const double *)ctx->d_scratch, ch->d_cells);
        TG_CUDA(cudaGetLastError());
    }
    std::vector<double> nz(n, 1.0);
    if (noise) std::memcpy(nz.data(), noise, 8 * n);
    TG_CUDA(cudaMemcpyAsync(ch->d_noise, nz.data(), 8 * n, cudaMemcpyHostToDevice, s));
    int rc = establish_state(ch);
    if (rc != TONGA_OK) return rc;
    TG_CUDA(cudaStreamSynchronize(s));
    return TONGA_OK;
}

extern "C" int tonga_chains_set_exact_only(tonga_chains *ch, int32_t exact_only) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_exact_only: NULL");
    ch->exact_only = exact_only ? 1 : 0;
    return TONGA_OK;
}

extern "C" int tonga_chains_profile(tonga_chains *ch, int32_t enable, int64_t *cycles /* [nChains][16] or NULL */) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_profile: NULL");
    if (ch->wide) return tg::fail(TONGA_ERR_STATE, "tonga_chains_profile: per-phase counters exist only in the resident sampler");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    TG_CUDA(cudaStreamSynchronize(ch->ctx->stream));
    const size_t bytes = 8 * (size_t)ch->n * 16;
    if (cycles && ch->d_prof) TG_CUDA(cudaMemcpy(cycles, ch->d_prof, bytes, cudaMemcpyDeviceToHost));
    if (enable && !ch->d_prof) {
        TG_CUDA(cudaMalloc((void **)&ch->d_prof, bytes));
        TG_CUDA(cudaMemset(ch->d_prof, 0, bytes));
    } else if (enable) {
        TG_CUDA(cudaMemset(ch->d_prof, 0, bytes));
    } else if (ch->d_prof) {
        cudaFree(ch->d_prof);
        ch->d_prof = nullptr;
    }
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

extern "C" int tonga_chains_get_beta(tonga_chains *ch, double *beta) {
    if (!ch || !beta) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_beta: NULL");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    TG_CUDA(cudaMemcpyAsync(beta, ch->d_beta, 8 * (size_t)ch->n, cudaMemcpyDeviceToHost, ch->ctx->stream));
    TG_CUDA(cudaStreamSynchronize(ch->ctx->stream));
    return TONGA_OK;
}

extern "C" int tonga_chains_temper_swap(tonga_chains *ch, int32_t n_all, const double *phi_all, const double *noise_all, double *beta_all,
                                        int64_t offset, int32_t ladder_size, int64_t step, uint64_t seed) {
    if (!ch || ladder_size < 1 || ladder_size > tg::TEMPER_MAX_T || step < 0)
        return tg::fail(TONGA_ERR_ARG, "tonga_chains_temper_swap: bad argument (1 <= ladder_size <= 256)");
    const bool own = !phi_all && !noise_all && !beta_all;
    if (!own && (!phi_all || !noise_all || !beta_all)) return tg::fail(TONGA_ERR_ARG, "tonga_chains_temper_swap: pass all three arrays or none");
    if (own) { n_all = ch->n; offset = 0; }
    if (n_all < 1 || n_all % ladder_size != 0 || offset < 0 || offset + ch->n > n_all)
        return tg::fail(TONGA_ERR_ARG, "tonga_chains_temper_swap: n_all must be a multiple of ladder_size and hold this batch's slice");
    if (!ch->have_models) return tg::fail(TONGA_ERR_STATE, "tonga_chains_temper_swap: no models yet");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    const int n_lad = n_all / ladder_size;
    tg::tg_temper_swap_kernel<<<(n_lad + 63) / 64, 64, 0, ctx->stream>>>(n_lad, ladder_size, ctx->R, own ? ch->d_phi : phi_all, own ? ch->d_noise : noise_all,
                                                                        own ? ch->d_beta : beta_all, step, seed, ch->d_swapstat);
    TG_CUDA(cudaGetLastError());
    if (!own) TG_CUDA(cudaMemcpyAsync(ch->d_beta, beta_all + offset, 8 * (size_t)ch->n, cudaMemcpyDeviceToDevice, ctx->stream));
    return TONGA_OK;
}

extern "C" int tonga_chains_temper_stats(tonga_chains *ch, int64_t *accepted, int64_t *attempted, int32_t reset) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_temper_stats: NULL");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    unsigned long long st[2] = {0, 0};
    TG_CUDA(cudaMemcpyAsync(st, ch->d_swapstat, 16, cudaMemcpyDeviceToHost, ch->ctx->stream));
    if (reset) TG_CUDA(cudaMemsetAsync(ch->d_swapstat, 0, 16, ch->ctx->stream));
    TG_CUDA(cudaStreamSynchronize(ch->ctx->stream));
    if (accepted) *accepted = (int64_t)st[0];
    if (attempted) *attempted = (int64_t)st[1];
    return TONGA_OK;
}

extern "C" int tonga_chains_scalar_ptrs(tonga_chains *ch, void **phi, void **noise, void **beta) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_scalar_ptrs: NULL");
    if (phi) *phi = ch->d_phi;
    if (noise) *noise = ch->d_noise;
    if (beta) *beta = ch->d_beta;
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
    // resident sampler, mode 0: raw draws of one piece of the launch (64 B per chain-iteration, at most 256 MB)
    int64_t raw_iters = std::max<int64_t>(1, std::min<int64_t>(nIter, (int64_t)((256u << 20) / (64 * n))));
    if (const char *e = std::getenv("TONGA_RAW_ITERS")) raw_iters = std::max<int64_t>(1, std::min<int64_t>(raw_iters, std::atoll(e)));  // tests: force the cut
    const size_t o_rec = 0, o_acc = o_rec + (recs ? al(sizeof(tonga_proposal) * N) : 0), o_phi = o_acc + (tr_accept ? al(N) : 0),
                 o_K = o_phi + (tr_phi ? al(8 * N) : 0), o_raw = o_K + (tr_K ? al(4 * N) : 0),
                 total = o_raw + ((!ch->wide && mode == 0) ? al(64 * n * (size_t)raw_iters) : 0);
    int rc = tg::ensure_scratch(ctx, total);
    if (rc != TONGA_OK) return rc;
    char *d = (char *)ctx->d_scratch;
    if (mode == 1) TG_CUDA(cudaMemcpyAsync(d + o_rec, recs, sizeof(tonga_proposal) * N, cudaMemcpyHostToDevice, s));

    const tonga_proposal *recs_in = (mode == 1) ? (const tonga_proposal *)(d + o_rec) : nullptr;
    tonga_proposal *recs_out = (mode == 0 && recs) ? (tonga_proposal *)(d + o_rec) : nullptr;
    int8_t *d_tr_accept = tr_accept ? (int8_t *)(d + o_acc) : nullptr;
    double *d_tr_phi = tr_phi ? (double *)(d + o_phi) : nullptr;
    int32_t *d_tr_K = tr_K ? (int32_t *)(d + o_K) : nullptr;
    TG_CUDA(cudaEventRecord(ch->ev0, s));
    if (ch->wide) {
        tg::WideArgs w{};
        w.prm = ctx->prm; w.R = ctx->R; w.Rp = ch->Rp; w.KC = ch->KC; w.mode = mode; w.hist_cap = ch->hist_cap; w.nIter = nIter;
        w.seed = ch->seed; w.chain_id0 = ch->chain_id0; w.ray_orig = ctx->d_ray_orig; w.ray_rank = ctx->d_ray_rank; w.tS = ctx->d_tS; w.sig = ctx->d_sig;
        w.K = ch->d_K; w.cells = ch->d_cells; w.phi = ch->d_phi; w.noise = ch->d_noise; w.beta = ch->d_beta; w.tstar = ch->d_tstar;
        w.counts = ch->d_counts; w.pending_slot = ch->d_pending;
        w.Kc = ch->d_Kc; w.cells_c = ch->d_cells_c; w.ptS_c = ch->d_ptS_tmp; w.props = ch->d_props;
        w.recs_in = recs_in; w.recs_out = recs_out; w.tr_accept = d_tr_accept; w.tr_phi = d_tr_phi; w.tr_K = d_tr_K;
        w.n_hist = ch->d_n_hist; w.model_num = ch->d_model_num; w.hist_K = ch->d_hist_K; w.hist_cells = ch->d_hist_cells;
        w.hist_phi = ch->d_hist_phi; w.hist_ptS = ch->d_hist_ptS; w.hist_iter = ch->d_hist_iter; w.hist_action = ch->d_hist_action;
        w.hist_accept = ch->d_hist_accept; w.hist_next = ch->d_hist_next;
        w.streamed = ch->streamed ? 1 : 0; w.tstar_c = ch->d_tstar_c; w.accept_flag = ch->d_accept; w.cells_cf = ch->d_cells_cf; w.term_c = ch->d_term_c;
        w.active = ch->streamed ? ch->d_active : nullptr; w.n_active = ch->streamed ? ch->d_active + ch->n : nullptr;
        tg::StreamArgs sa{};
        sa.tiles = ch->d_stiles; sa.pxf = ctx->d_pxf; sa.pyf = ctx->d_pyf; sa.pzf = ctx->d_pzf; sa.px = ctx->d_px; sa.py = ctx->d_py; sa.pz = ctx->d_pz;
        sa.dt = ctx->d_dt; sa.ray_off = ctx->d_ray_off; sa.tol_alpha = ctx->tol_alpha; sa.tol_beta2 = ctx->tol_beta2;
        sa.exact_only = (ch->exact_only || ctx->exact_only) ? 1 : 0; sa.KC = ch->KC; sa.Rp = ch->Rp; sa.tile_pts = ch->stile_pts;
        sa.Ppad = ctx->Ppad; sa.props = ch->d_props; sa.Kc = ch->d_Kc; sa.cells_c = ch->d_cells_c; sa.cells_cf = ch->d_cells_cf; sa.n_chains = ch->n; sa.owner = ch->d_owner16; sa.dcache = ch->d_dcache;
        sa.tstar = ch->d_tstar; sa.tstar_c = ch->d_tstar_c; sa.accept_flag = ch->d_accept;
        sa.tile_changed = ch->d_tile_changed; sa.active = ch->d_active; sa.n_active = ch->d_active + ch->n; sa.n_tiles = ch->n_stiles;
        sa.term_c = ch->d_term_c; sa.tS = ctx->d_tS; sa.sig = ctx->d_sig; sa.noise = ch->d_noise;
        const bool sharded = ch->streamed && ch->sh_world > 1;
        if (sharded && !ch->sh_connected) return tg::fail(TONGA_ERR_STATE, "tonga_chains_run: ray-sharded batch is not connected (tonga_chains_shard_connect)");
        tg::ShardArgs sh{};
        sh.rank = 0; sh.world = 1; sh.tile0 = 0; sh.tile1 = ch->n_stiles;
        const size_t xbuf = 8 * n * (size_t)ch->Rp;  // one (term or t*) buffer of one parity
        if (sharded) {
            sh.rank = ch->sh_rank; sh.world = ch->sh_world; sh.tile0 = ch->sh_tile0; sh.tile1 = ch->sh_tile1;
            sh.flags = (unsigned long long *)ch->d_xch; sh.err = (int *)(ch->d_xch + 128 * (size_t)ch->sh_world);
            double tmo = 30e3;
            if (const char *e = std::getenv("TONGA_SHARD_TIMEOUT_MS")) tmo = std::max(1.0, std::atof(e));
            sh.timeout_ns = (unsigned long long)(tmo * 1e6);
            for (int g = 0; g < sh.world; g++) sh.peer_base[g] = ch->sh_peer[g];
            TG_CUDA(cudaMemsetAsync(sh.err, 0, 4, s));
            if (ch->culled) {
                sh.mb = 1; sh.mb_cap = ch->mb_cap; sh.mb_n = ch->n; sh.ndirty = ch->d_ndirty;
                sh.mb_hdr = ch->xch_hdr; sh.mb_sec = ch->mb_sec; sh.mb_cnt_bytes = ch->mb_cnt_bytes;
            }
        }
        const bool culled = ch->streamed && ch->culled;
        tg::CullArgs ca{};
        if (culled) {
            ca.sub = ch->d_sub; ca.raysph = ch->d_raysph; ca.sub_off = ch->d_sub_off; ca.dmax = ch->d_dmax; ca.term = ch->d_term;
            ca.cand = ch->d_cand; ca.ncand = ch->d_ncand; ca.clist = ch->d_clist; ca.nclist = ch->d_nclist; ca.ncommit = ch->d_ncommit; ca.work_off2 = ch->d_work_off2; ca.dirty = ch->d_dirty; ca.ndirty = ch->d_ndirty;
            ca.work_off = ch->d_work_off; ca.R = ctx->R; ca.ray0 = 0; ca.ray1 = ctx->R; ca.maxn = ch->s2_maxn; ca.p0 = 0; ca.p1 = ctx->Ppad;
            if (sharded) {
                int32_t r0 = 0, r1 = 0;
                int64_t q0 = 0, q1 = 0;
                tonga_chains_shard_info(ch, nullptr, nullptr, &r0, &r1, &q0, &q1);
                ca.ray0 = r0; ca.ray1 = r1; ca.p0 = q0; ca.p1 = q1;
            }
            w.nclist = ch->d_nclist; w.ncommit = ch->d_ncommit; w.culled = 1; w.ncand = ch->d_ncand; w.ndirty = ch->d_ndirty; w.dirty = ch->d_dirty; w.term = ch->d_term;
        }
        const dim3 cgrid((unsigned)std::max(1, (ca.ray1 - ca.ray0 + tg::CULL_THREADS - 1) / tg::CULL_THREADS), (unsigned)ch->n);
        const unsigned s2grid = (unsigned)(ctx->sm_count > 0 ? ctx->sm_count : 148) * (unsigned)S2_MIN_CTAS;
        const dim3 sgrid((unsigned)((size_t)(sh.tile1 - sh.tile0) * (size_t)((ch->n + tg::STREAM_GROUP - 1) / tg::STREAM_GROUP)));
        const int exact = (ch->exact_only || ctx->exact_only) ? 1 : 0;
        for (int64_t it = 0; it < nIter; it++) {
            w.it = it; w.iter = ch->iter_done + 1 + it;
            if (sharded) sh.seq = ++ch->sh_seq;  // this iteration's exchange
            if (sharded && !culled) {  // tile kernels: parity buffers of (t*, term) for ALL rays in every rank's block
                const size_t par = (size_t)(sh.seq & 1ull);
                for (int g = 0; g < sh.world; g++) {
                    sh.peer_term[g] = (double *)(ch->sh_peer[g] + ch->xch_hdr + par * xbuf);
                    sh.peer_tsc[g] = (double *)(ch->sh_peer[g] + ch->xch_hdr + (2 + par) * xbuf);
                }
                w.term_c = sa.term_c = sh.peer_term[sh.rank];
                w.tstar_c = sa.tstar_c = sh.peer_tsc[sh.rank];
            }
            w.sh = sh; sa.sh = sh;
            if (ch->streamed) TG_CUDA(cudaMemsetAsync(ch->d_active + ch->n, 0, 4, s));
            tg::tg_wide_propose_kernel<<<ch->n, tg::WIDE_PROPOSE_THREADS, 0, s>>>(w);
            if (culled) {
                ca.s = sa;
                tg::tg_cull_kernel<<<cgrid, tg::CULL_THREADS, 0, s>>>(ca);
                tg::tg_cull_prefix_kernel<<<1, 1024, 0, s>>>(ch->d_active, ch->d_active + ch->n, ch->d_ncand, ch->d_work_off);
                tg::tg_stream2_kernel<false><<<s2grid, tg::S2_THREADS, ch->s2_smem, s>>>(ca);
                if (sharded) tg::tg_shard_signal_kernel<<<1, 256, 0, s>>>(sh);
            } else if (ch->streamed) {
                if (sgrid.x > 0) tg::tg_stream_kernel<false><<<sgrid, tg::STREAM_THREADS, ch->stream_smem, s>>>(sa);
                if (sharded) tg::tg_shard_signal_kernel<<<1, 32, 0, s>>>(sh);
            } else if (!ctx->prm.debug_prior) {
                rc = tg::launch_evaluate(ctx, ch->n, ch->KC, ch->d_Kc, ch->d_cells_c, nullptr, ch->d_ptS_tmp, nullptr, nullptr, nullptr, nullptr, false, nullptr, exact);
                if (rc != TONGA_OK) return rc;
            }
            tg::tg_wide_accept_kernel<<<ch->n, TG_PHI_LANES, 0, s>>>(w);
            if (culled) {
                tg::tg_cull_prefix_kernel<<<1, 1024, 0, s>>>(ch->d_active, ch->d_active + ch->n, ch->d_ncommit, ch->d_work_off2);
                tg::tg_stream2_kernel<true><<<s2grid, tg::S2_THREADS, ch->s2_smem, s>>>(ca);
                tg::tg_renumber_kernel<<<dim3(128, (unsigned)ch->n), 256, 0, s>>>(ca);
            } else if (ch->streamed && sgrid.x > 0) tg::tg_stream_kernel<true><<<sgrid, tg::STREAM_THREADS, ch->stream_smem, s>>>(sa);
        }
        TG_CUDA(cudaGetLastError());
    } else {
    tg::SamplerArgs a{};
    a.px = ctx->d_px; a.py = ctx->d_py; a.pz = ctx->d_pz; a.dt = ctx->d_dt; a.tS = ctx->d_tS; a.sig = ctx->d_sig;
    a.pxf = ctx->d_pxf; a.pyf = ctx->d_pyf; a.pzf = ctx->d_pzf; a.tol_alpha = ctx->tol_alpha; a.tol_beta2 = ctx->tol_beta2;
    a.exact_only = (ch->exact_only || ctx->exact_only) ? 1 : 0;
    a.prof = ch->d_prof;
    a.ray_off = ctx->d_ray_off; a.ray_rank = ctx->d_ray_rank;
    a.R = ctx->R; a.Rp = ch->Rp; a.KC = ch->KC; a.P = (int)ctx->P; a.Ppad = (int)ctx->Ppad;
    a.prm = ctx->prm;
    a.K = ch->d_K; a.cells = ch->d_cells; a.phi = ch->d_phi; a.noise = ch->d_noise; a.beta = ch->d_beta;
    a.owner = ch->d_owner; a.dcache = ch->d_dcache; a.tstar = ch->d_tstar; a.counts = ch->d_counts; a.pending_slot = ch->d_pending;
    a.mode = mode; a.trace_stride = nIter;
    a.recs_in = recs_in; a.recs_out = recs_out; a.tr_accept = d_tr_accept; a.tr_phi = d_tr_phi; a.tr_K = d_tr_K;
    a.chain_id0 = ch->chain_id0;
    a.hist_cap = ch->hist_cap; a.n_hist = ch->d_n_hist; a.model_num = ch->d_model_num;
    a.hist_K = ch->d_hist_K; a.hist_cells = ch->d_hist_cells; a.hist_phi = ch->d_hist_phi; a.hist_ptS = ch->d_hist_ptS;
    a.hist_iter = ch->d_hist_iter; a.hist_action = ch->d_hist_action; a.hist_accept = ch->d_hist_accept; a.hist_next = ch->d_hist_next;
    a.raw = (const tg::RawDraw *)(d + o_raw);
    // the launch is cut into pieces of at most raw_iters iterations (the raw draws of a piece are generated just before it)
    for (int64_t it0 = 0; it0 < nIter; it0 += raw_iters) {
        const int64_t ni = std::min<int64_t>(raw_iters, nIter - it0);
        a.it0 = it0; a.nIter = ni; a.iter0 = ch->iter_done + 1 + it0;
        if (mode == 0) {
            const long long tot = (long long)ch->n * ni;
            tg::tg_pregen_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(ch->n, ni, a.iter0, ch->seed, ch->chain_id0, ctx->prm.n_actions, (tg::RawDraw *)(d + o_raw));
        }
        a.perm = nullptr;
        if (!ch->no_order) {
            tg::tg_order_kernel<<<1, 1024, 0, s>>>(ch->n, ch->d_K, ch->d_perm);
            a.perm = ch->d_perm;
        }
        const bool small = ctx->R <= tg::ST * TG_SMALL_CHUNKS;
        const bool plain = a.mode == 0 && !a.recs_out && !a.tr_accept && !a.tr_phi && !a.tr_K;  // no replay, records or traces: the lean instantiation
        if (ch->d_prof) {  // instrumented instantiation (tonga_chains_profile)
            if (small) tg::tg_sampler_kernel<TG_SMALL_CHUNKS, true, false><<<ch->n, tg::ST, ch->smem, s>>>(a);
            else tg::tg_sampler_kernel<TG_RESIDENT_MAX_CHUNKS, true, false><<<ch->n, tg::ST, ch->smem, s>>>(a);
        } else if (plain) {
            if (small) tg::tg_sampler_kernel<TG_SMALL_CHUNKS, false, true><<<ch->n, tg::ST, ch->smem, s>>>(a);
            else tg::tg_sampler_kernel<TG_RESIDENT_MAX_CHUNKS, false, true><<<ch->n, tg::ST, ch->smem, s>>>(a);
        } else {
            if (small) tg::tg_sampler_kernel<TG_SMALL_CHUNKS, false, false><<<ch->n, tg::ST, ch->smem, s>>>(a);
            else tg::tg_sampler_kernel<TG_RESIDENT_MAX_CHUNKS, false, false><<<ch->n, tg::ST, ch->smem, s>>>(a);
        }
        TG_CUDA(cudaGetLastError());
    }
    }
    TG_CUDA(cudaEventRecord(ch->ev1, s));
    ch->iter_done += nIter;
    if (mode == 0 && recs) TG_CUDA(cudaMemcpyAsync(recs, d + o_rec, sizeof(tonga_proposal) * N, cudaMemcpyDeviceToHost, s));
    if (tr_accept) TG_CUDA(cudaMemcpyAsync(tr_accept, d + o_acc, N, cudaMemcpyDeviceToHost, s));
    if (tr_phi) TG_CUDA(cudaMemcpyAsync(tr_phi, d + o_phi, 8 * N, cudaMemcpyDeviceToHost, s));
    if (tr_K) TG_CUDA(cudaMemcpyAsync(tr_K, d + o_K, 4 * N, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    TG_CUDA(cudaEventElapsedTime(&ch->last_ms, ch->ev0, ch->ev1));
    if (ch->streamed && ch->sh_world > 1) {
        int err = 0;
        TG_CUDA(cudaMemcpy(&err, ch->d_xch + 128 * (size_t)ch->sh_world, 4, cudaMemcpyDeviceToHost));
        if (err) return tg::fail(TONGA_ERR_PEER, "tonga_chains_run: a ray-shard peer did not publish its exchange in time (the ranks must call run with the same arguments); the batch's state is undefined");
    }
    return TONGA_OK;
}

// ---- ray sharding of the streamed sampler -------------------------------------------------------------------------------
extern "C" int tonga_chains_shard_init(tonga_chains *ch, int32_t rank, int32_t world, void **xch_base, uint64_t *xch_bytes) {
    if (!ch || world < 1 || world > tg::TG_MAX_SHARDS || rank < 0 || rank >= world) return tg::fail(TONGA_ERR_ARG, "tonga_chains_shard_init: bad argument (1 <= world <= 16)");
    if (!ch->streamed) return tg::fail(TONGA_ERR_STATE, "tonga_chains_shard_init: ray sharding needs the STREAMED sampler");
    if (ch->d_xch) return tg::fail(TONGA_ERR_STATE, "tonga_chains_shard_init: already initialised");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    // own tiles: a contiguous range, balanced by ray points (tiles hold whole rays; every tile is <= stile_pts points)
    int t0 = 0, t1 = 0;
    tonga_shard_range(ch->n_stiles, rank, world, &t0, &t1);
    ch->sh_rank = rank; ch->sh_world = world; ch->sh_tile0 = t0; ch->sh_tile1 = t1;
    ch->xch_hdr = (128 * (size_t)(world + 1) + 255) & ~(size_t)255;
    ch->xch_bytes = ch->xch_hdr + 4 * 8 * (size_t)ch->n * (size_t)ch->Rp;
    if (ch->culled) {  // mailbox protocol: only the dirty rays travel; capacity = the largest own ray count of any rank
        int cap = 1;
        for (int g = 0; g < world; g++) {
            int a0 = 0, a1 = 0;
            tonga_shard_range(ch->n_stiles, g, world, &a0, &a1);
            if (a1 > a0) cap = std::max(cap, ch->h_stiles[a1 - 1].r1 - ch->h_stiles[a0].r0);
        }
        ch->mb_cap = cap;
        ch->mb_cnt_bytes = (4 * (size_t)ch->n + 255) & ~(size_t)255;
        ch->mb_sec = ch->mb_cnt_bytes + 20 * (size_t)ch->n * (size_t)cap;
        ch->mb_sec = (ch->mb_sec + 255) & ~(size_t)255;
        ch->xch_bytes = ch->xch_hdr + 2 * (size_t)world * ch->mb_sec;
    }
    TG_CUDA(cudaMalloc((void **)&ch->d_xch, ch->xch_bytes));
    TG_CUDA(cudaMemset(ch->d_xch, 0, ch->xch_bytes));
    ch->sh_peer[rank] = ch->d_xch;
    ch->sh_connected = (world == 1);
    if (xch_base) *xch_base = ch->d_xch;
    if (xch_bytes) *xch_bytes = ch->xch_bytes;
    return TONGA_OK;
}

extern "C" int tonga_shard_range(int32_t n_tiles, int32_t rank, int32_t world, int32_t *tile0, int32_t *tile1) {
    if (n_tiles < 0 || world < 1 || rank < 0 || rank >= world || !tile0 || !tile1) return tg::fail(TONGA_ERR_ARG, "tonga_shard_range: bad argument");
    *tile0 = (int32_t)(((int64_t)n_tiles * rank) / world);
    *tile1 = (int32_t)(((int64_t)n_tiles * (rank + 1)) / world);
    return TONGA_OK;
}

extern "C" int tonga_chains_shard_connect(tonga_chains *ch, void *const *peer_bases) {
    if (!ch || !peer_bases) return tg::fail(TONGA_ERR_ARG, "tonga_chains_shard_connect: NULL");
    if (!ch->d_xch) return tg::fail(TONGA_ERR_STATE, "tonga_chains_shard_connect: call tonga_chains_shard_init first");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    for (int g = 0; g < ch->sh_world; g++) {
        if (g == ch->sh_rank) continue;
        if (!peer_bases[g]) return tg::fail(TONGA_ERR_ARG, "tonga_chains_shard_connect: NULL peer block");
        ch->sh_peer[g] = (unsigned char *)peer_bases[g];
    }
    ch->sh_connected = true;
    return TONGA_OK;
}

extern "C" int tonga_chains_shard_info(const tonga_chains *ch, int32_t *rank, int32_t *world, int32_t *ray0, int32_t *ray1, int64_t *point0, int64_t *point1) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_shard_info: NULL");
    if (rank) *rank = ch->sh_rank;
    if (world) *world = ch->sh_world;
    const bool any = ch->streamed && ch->sh_tile1 > ch->sh_tile0;
    const int r0 = !ch->streamed ? 0 : (any ? ch->h_stiles[ch->sh_tile0].r0 : 0), r1 = !ch->streamed ? ch->ctx->R : (any ? ch->h_stiles[ch->sh_tile1 - 1].r1 : 0);
    const int64_t p0 = !ch->streamed ? 0 : (any ? ch->h_stiles[ch->sh_tile0].p0 : 0), p1 = !ch->streamed ? ch->ctx->P : (any ? ch->h_stiles[ch->sh_tile1 - 1].p1 : 0);
    if (ray0) *ray0 = r0;
    if (ray1) *ray1 = r1;
    if (point0) *point0 = p0;
    if (point1) *point1 = p1;
    return TONGA_OK;
}

// CUDA IPC helpers for hosts that have no other way to exchange device pointers between the per-GPU processes
extern "C" int tonga_ipc_export(const void *dev_ptr, unsigned char *handle /* [64] */) {
    if (!dev_ptr || !handle) return tg::fail(TONGA_ERR_ARG, "tonga_ipc_export: NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    TG_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    std::memcpy(handle, &h, 64);
    return TONGA_OK;
}
extern "C" int tonga_ipc_open(int32_t device, const unsigned char *handle /* [64] */, void **dev_ptr) {
    if (!handle || !dev_ptr) return tg::fail(TONGA_ERR_ARG, "tonga_ipc_open: NULL");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    TG_CUDA(cudaSetDevice(device));
    TG_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TONGA_OK;
}
extern "C" int tonga_ipc_close(int32_t device, void *dev_ptr) {
    if (!dev_ptr) return TONGA_OK;
    TG_CUDA(cudaSetDevice(device));
    TG_CUDA(cudaIpcCloseMemHandle(dev_ptr));
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
    const size_t n = (size_t)ch->n, R = (size_t)ctx->R, P = (size_t)ctx->P, Pp = (size_t)ctx->Ppad;
    // cells and t* are brought into the caller's layout on the device (scratch), then everything is copied straight into the caller's
    // buffers on the stream -- at PCIe speed when they are page-locked (tonga_host_alloc) -- with one synchronisation at the end
    cudaStream_t s = ctx->stream;
    std::vector<int32_t> hK(n);
    TG_CUDA(cudaMemcpyAsync(hK.data(), ch->d_K, 4 * n, cudaMemcpyDeviceToHost, s));
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_c = 0, o_t = o_c + (cells ? al(8 * n * 4 * (size_t)(Kcap > 0 ? Kcap : 1)) : 0), total = o_t + (ptS ? al(8 * n * (R ? R : 1)) : 0);
    if (cells && Kcap < 1) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_state: Kcap < 1");
    if (total) {
        int rc = tg::ensure_scratch(ctx, total);
        if (rc != TONGA_OK) return rc;
    }
    char *sc = (char *)ctx->d_scratch;
    if (cells) {
        const size_t tot = n * 4 * (size_t)Kcap;
        tg::tg_pack_cells_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(ch->n, ch->KC, Kcap, ch->d_K, ch->d_cells, (double *)(sc + o_c));
        TG_CUDA(cudaGetLastError());
        TG_CUDA(cudaMemcpyAsync(cells, sc + o_c, 8 * tot, cudaMemcpyDeviceToHost, s));
    }
    if (phi) TG_CUDA(cudaMemcpyAsync(phi, ch->d_phi, 8 * n, cudaMemcpyDeviceToHost, s));
    if (noise) TG_CUDA(cudaMemcpyAsync(noise, ch->d_noise, 8 * n, cudaMemcpyDeviceToHost, s));
    if (ptS && R) {  // device rows are in sorted ray order
        const size_t tot = n * R;
        tg::tg_unsort_tstar_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(ch->n, ctx->R, ch->Rp, ctx->d_ray_orig, ch->d_tstar, (double *)(sc + o_t));
        TG_CUDA(cudaGetLastError());
        TG_CUDA(cudaMemcpyAsync(ptS, sc + o_t, 8 * tot, cudaMemcpyDeviceToHost, s));
    }
    TG_CUDA(cudaStreamSynchronize(s));
    if (K) std::memcpy(K, hK.data(), 4 * n);
    if (cells)
        for (size_t i = 0; i < n; i++)
            if (hK[i] > Kcap) return tg::fail(TONGA_ERR_CAPACITY, "tonga_chains_get_state: Kcap smaller than a chain's nCells");
    if (owners && ch->streamed) {
        std::vector<uint16_t> ho(n * Pp);
        TG_CUDA(cudaMemcpy(ho.data(), ch->d_owner16, 2 * n * Pp, cudaMemcpyDeviceToHost));
        int64_t sp0 = 0, sp1 = (int64_t)P;
        if (ch->sh_world > 1) tonga_chains_shard_info(ch, nullptr, nullptr, nullptr, nullptr, &sp0, &sp1);
        for (size_t i = 0; i < n; i++)
            for (size_t p = 0; p < P; p++) {
                const uint16_t o = ho[i * Pp + p];
                const bool mine = (int64_t)p >= sp0 && (int64_t)p < sp1;  // ray-sharded: -2 = a point of another rank's shard
                owners[i * P + (size_t)ctx->h_point_orig[p]] = !mine ? -2 : ((o == 0xFFFFu) ? -1 : (int32_t)o);
            }
    } else if (owners && ch->wide) {  // no resident owner state: one forward model of the current models (caller's point order)
        int rc = tg::ensure_scratch(ctx, 4 * n * P);
        if (rc != TONGA_OK) return rc;
        rc = tg::launch_evaluate(ctx, ch->n, ch->KC, ch->d_K, ch->d_cells, ch->d_noise, ch->d_ptS_tmp, nullptr, (int32_t *)ctx->d_scratch, nullptr, nullptr, true);
        if (rc != TONGA_OK) return rc;
        TG_CUDA(cudaMemcpyAsync(owners, ctx->d_scratch, 4 * n * P, cudaMemcpyDeviceToHost, ctx->stream));
        TG_CUDA(cudaStreamSynchronize(ctx->stream));
    } else if (owners) {
        std::vector<uint8_t> ho(n * Pp);
        TG_CUDA(cudaMemcpy(ho.data(), ch->d_owner, n * Pp, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++)
            for (size_t p = 0; p < P; p++) {  // device order: sorted rays
                const uint8_t o = ho[i * Pp + p];
                owners[i * P + (size_t)ctx->h_point_orig[p]] = (o == TG_OWNER_NONE) ? -1 : (int32_t)o;
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
    if (hist_K) TG_CUDA(cudaMemcpy(hist_K, ch->d_hist_K, 4 * nh, cudaMemcpyDefault));
    if (hist_cells) {
        if (Kcap == (int)KC) {
            TG_CUDA(cudaMemcpy(hist_cells, ch->d_hist_cells, 8 * nh * 4 * KC, cudaMemcpyDefault));
        } else {
            if (Kcap < 1) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_history: Kcap < 1");
            std::vector<double> hc(nh * 4 * KC);
            TG_CUDA(cudaMemcpy(hc.data(), ch->d_hist_cells, 8 * nh * 4 * KC, cudaMemcpyDefault));
            const size_t kc = std::min((size_t)Kcap, KC);
            for (size_t i = 0; i < nh * 4; i++) std::memcpy(hist_cells + i * (size_t)Kcap, &hc[i * KC], 8 * kc);
        }
    }
    if (hist_phi) TG_CUDA(cudaMemcpy(hist_phi, ch->d_hist_phi, 8 * nh, cudaMemcpyDefault));
    if (hist_ptS) TG_CUDA(cudaMemcpy(hist_ptS, ch->d_hist_ptS, 8 * nh * R, cudaMemcpyDefault));
    if (hist_iter) TG_CUDA(cudaMemcpy(hist_iter, ch->d_hist_iter, 8 * nh, cudaMemcpyDefault));
    if (hist_action) TG_CUDA(cudaMemcpy(hist_action, ch->d_hist_action, 4 * nh, cudaMemcpyDefault));
    if (hist_accept) TG_CUDA(cudaMemcpy(hist_accept, ch->d_hist_accept, 4 * nh, cudaMemcpyDefault));
    if (hist_next_action) TG_CUDA(cudaMemcpy(hist_next_action, ch->d_hist_next, 4 * nh, cudaMemcpyDefault));
    return TONGA_OK;
}


// ---- checkpoint / resume (TD_inversion_function.jl:40-67, 282-294): the model part is tonga_chains_get_state / set_models; these
// entry points carry the rest of a chain's progress so that a resumed batch continues bit-identically.
extern "C" int tonga_chains_get_progress(tonga_chains *ch, int64_t *iter_done, int64_t *model_num, int32_t *pending_slot) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_get_progress: NULL");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    TG_CUDA(cudaStreamSynchronize(ch->ctx->stream));
    const size_t n = (size_t)ch->n;
    if (iter_done) *iter_done = ch->iter_done;
    if (model_num) TG_CUDA(cudaMemcpy(model_num, ch->d_model_num, 8 * n, cudaMemcpyDeviceToHost));
    if (pending_slot) TG_CUDA(cudaMemcpy(pending_slot, ch->d_pending, 4 * n, cudaMemcpyDeviceToHost));
    return TONGA_OK;
}

extern "C" int tonga_chains_set_progress(tonga_chains *ch, int64_t iter_done, const int64_t *model_num, const int32_t *pending_slot,
                                         const int64_t *counts) {
    if (!ch || iter_done < 0) return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_progress: bad argument");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    TG_CUDA(cudaStreamSynchronize(ch->ctx->stream));
    const size_t n = (size_t)ch->n;
    if (pending_slot)
        for (size_t i = 0; i < n; i++)
            if (pending_slot[i] < -1 || pending_slot[i] >= ch->hist_cap)
                return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_progress: pending_slot outside [-1, hist_cap)");
    ch->iter_done = iter_done;
    if (model_num) TG_CUDA(cudaMemcpy(ch->d_model_num, model_num, 8 * n, cudaMemcpyHostToDevice));
    if (pending_slot) TG_CUDA(cudaMemcpy(ch->d_pending, pending_slot, 4 * n, cudaMemcpyHostToDevice));
    if (counts) TG_CUDA(cudaMemcpy(ch->d_counts, counts, 8 * n * 15, cudaMemcpyHostToDevice));
    return TONGA_OK;
}

extern "C" int tonga_chains_set_history(tonga_chains *ch, int32_t Kcap, const int32_t *n_hist, const int32_t *hist_K, const double *hist_cells,
                                        const double *hist_phi, const double *hist_ptS, const int64_t *hist_iter, const int32_t *hist_action,
                                        const int32_t *hist_accept, const int32_t *hist_next_action) {
    if (!ch || !n_hist) return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_history: bad argument");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    TG_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t n = (size_t)ch->n, KC = (size_t)ch->KC, R = (size_t)ctx->R, H = (size_t)ch->hist_cap, nh = n * H;
    for (size_t i = 0; i < n; i++)
        if (n_hist[i] < 0) return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_history: negative n_hist");
    if (nh > 0 && hist_cells && Kcap != (int)KC) return tg::fail(TONGA_ERR_ARG, "tonga_chains_set_history: Kcap must equal tonga_chains_kcap()");
    TG_CUDA(cudaMemcpy(ch->d_n_hist, n_hist, 4 * n, cudaMemcpyHostToDevice));
    if (nh == 0) return TONGA_OK;
    if (hist_K) TG_CUDA(cudaMemcpy(ch->d_hist_K, hist_K, 4 * nh, cudaMemcpyDefault));
    if (hist_cells) TG_CUDA(cudaMemcpy(ch->d_hist_cells, hist_cells, 8 * nh * 4 * KC, cudaMemcpyDefault));
    if (hist_phi) TG_CUDA(cudaMemcpy(ch->d_hist_phi, hist_phi, 8 * nh, cudaMemcpyDefault));
    if (hist_ptS) TG_CUDA(cudaMemcpy(ch->d_hist_ptS, hist_ptS, 8 * nh * R, cudaMemcpyDefault));
    if (hist_iter) TG_CUDA(cudaMemcpy(ch->d_hist_iter, hist_iter, 8 * nh, cudaMemcpyDefault));
    if (hist_action) TG_CUDA(cudaMemcpy(ch->d_hist_action, hist_action, 4 * nh, cudaMemcpyDefault));
    if (hist_accept) TG_CUDA(cudaMemcpy(ch->d_hist_accept, hist_accept, 4 * nh, cudaMemcpyDefault));
    if (hist_next_action) TG_CUDA(cudaMemcpy(ch->d_hist_next, hist_next_action, 4 * nh, cudaMemcpyDefault));
    return TONGA_OK;
}

extern "C" int tonga_chains_verify(tonga_chains *ch, int64_t *owner_mismatch, double *max_dphi, double *max_dts) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_verify: NULL");
    if (!ch->have_models) return tg::fail(TONGA_ERR_STATE, "tonga_chains_verify: no models yet");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t nPp = (size_t)ch->n * (size_t)ctx->Ppad;
    uint16_t *own16_tmp = nullptr;
    float *dc_tmp = ch->d_dcache_tmp;
    struct Tmp { void *a = nullptr, *b = nullptr; ~Tmp() { cudaFree(a); cudaFree(b); } } tmp;  // streamed: the reference copy lives only during verify
    if (ch->streamed) {
        TG_CUDA(cudaMalloc(&tmp.a, 2 * nPp));
        TG_CUDA(cudaMalloc(&tmp.b, 4 * nPp));
        own16_tmp = (uint16_t *)tmp.a;
        dc_tmp = (float *)tmp.b;
    }
    int rc = tg::launch_evaluate(ctx, ch->n, ch->KC, ch->d_K, ch->d_cells, ch->d_noise, ch->d_ptS_tmp, ch->d_phi_tmp, nullptr, ch->d_owner_tmp, dc_tmp, true, own16_tmp);
    if (rc != TONGA_OK) return rc;
    TG_CUDA(cudaMemsetAsync(ch->d_mism, 0, 8, s));
    TG_CUDA(cudaMemsetAsync(ch->d_maxd, 0, 16, s));
    dim3 grid(ch->streamed ? 64 : 8, ch->n);
    int64_t vp0 = 0, vp1 = (ch->wide && !ch->streamed) ? 0 : ctx->P;
    if (ch->streamed && ch->sh_world > 1) tonga_chains_shard_info(ch, nullptr, nullptr, nullptr, nullptr, &vp0, &vp1);  // the per-point state of the other shards' points is not maintained here
    tg::tg_verify_kernel<<<grid, 256, 0, s>>>(ch->n, vp0, vp1, ctx->Ppad, ctx->R, ch->Rp, ctx->d_ray_orig, ch->d_owner, ch->d_owner_tmp,
                                              ch->d_owner16, own16_tmp, ch->d_tstar, ch->d_ptS_tmp,
                                              ch->d_phi, ch->d_phi_tmp, ch->d_dcache, dc_tmp, ctx->tol_alpha, ctx->tol_beta2, ch->d_mism, ch->d_maxd);
    TG_CUDA(cudaGetLastError());
    if (ch->culled) {  // the culling state must be consistent with the chain state (violations count as mismatches)
        int32_t vr0 = 0, vr1 = ctx->R;
        if (ch->sh_world > 1) tonga_chains_shard_info(ch, nullptr, nullptr, &vr0, &vr1, nullptr, nullptr);
        tg::tg_stream_check_kernel<<<dim3((unsigned)((ctx->R + 7) / 8), (unsigned)ch->n), 256, 0, s>>>(ctx->R, ch->Rp, ctx->Ppad, vr0, vr1, ctx->d_ray_off, ch->d_dcache, ch->d_tstar,
                                                                                             ctx->d_tS, ctx->d_sig, ch->d_noise, ch->d_dmax, ch->d_term, ch->d_tstar_c,
                                                                                             ch->d_term_c, ch->d_mism);
        TG_CUDA(cudaGetLastError());
    }
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

extern "C" int tonga_chains_raster(tonga_chains *ch, int32_t n_nodes, const double *X, const double *Y, const double *Z, double *sum_out,
                                   double *sumsq_out, int64_t *count_out) {
    if (!ch || n_nodes < 0 || (n_nodes > 0 && (!X || !Y || !Z || !sum_out || !sumsq_out)))
        return tg::fail(TONGA_ERR_ARG, "tonga_chains_raster: bad argument");
    tonga_ctx *ctx = ch->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    TG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t n = (size_t)ch->n, nn = (size_t)n_nodes;
    std::vector<int32_t> nh(n);
    TG_CUDA(cudaStreamSynchronize(s));
    TG_CUDA(cudaMemcpy(nh.data(), ch->d_n_hist, 4 * n, cudaMemcpyDeviceToHost));
    long long cnt = 0;
    for (size_t i = 0; i < n; i++) cnt += std::min<long long>(nh[i], ch->hist_cap);
    if (count_out) *count_out = cnt;
    if (n_nodes == 0) return TONGA_OK;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_X = 0, o_Y = o_X + al(8 * nn), o_Z = o_Y + al(8 * nn), o_part = o_Z + al(8 * nn), o_sum = o_part + al(8 * n * 2 * nn),
                 o_sq = o_sum + al(8 * nn), total = o_sq + al(8 * nn);
    int rc = tg::ensure_scratch(ctx, total);
    if (rc != TONGA_OK) return rc;
    char *d = (char *)ctx->d_scratch;
    TG_CUDA(cudaMemcpyAsync(d + o_X, X, 8 * nn, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(d + o_Y, Y, 8 * nn, cudaMemcpyHostToDevice, s));
    TG_CUDA(cudaMemcpyAsync(d + o_Z, Z, 8 * nn, cudaMemcpyHostToDevice, s));
    dim3 grid((unsigned)((nn + 255) / 256), (unsigned)ch->n);
    tg::tg_raster_kernel<<<grid, 256, 8 * 4 * (size_t)ch->KC, s>>>(n_nodes, (const double *)(d + o_X), (const double *)(d + o_Y),
                                                                  (const double *)(d + o_Z), ch->KC, ch->hist_cap, ch->d_n_hist, ch->d_hist_K,
                                                                  ch->d_hist_cells, (double *)(d + o_part));
    TG_CUDA(cudaGetLastError());
    tg::tg_raster_reduce_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(n_nodes, ch->n, (const double *)(d + o_part), (double *)(d + o_sum),
                                                                          (double *)(d + o_sq));
    TG_CUDA(cudaGetLastError());
    TG_CUDA(cudaMemcpyAsync(sum_out, d + o_sum, 8 * nn, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaMemcpyAsync(sumsq_out, d + o_sq, 8 * nn, cudaMemcpyDeviceToHost, s));
    TG_CUDA(cudaStreamSynchronize(s));
    return TONGA_OK;
}

extern "C" int tonga_chains_history_host(tonga_chains *ch, void **hist_K, void **hist_cells, void **hist_phi, void **hist_ptS, void **hist_iter,
                                         void **hist_action, void **hist_accept, void **hist_next_action) {
    if (!ch) return tg::fail(TONGA_ERR_ARG, "tonga_chains_history_host: NULL");
    if (!ch->host_history) return tg::fail(TONGA_ERR_STATE, "tonga_chains_history_host: this batch keeps its history in device memory (create it with TONGA_HISTORY_ON_HOST)");
    std::lock_guard<std::mutex> lk(ch->ctx->mu);
    TG_CUDA(cudaSetDevice(ch->ctx->device));
    TG_CUDA(cudaStreamSynchronize(ch->ctx->stream));  // every store of the finished runs has reached host memory
    if (hist_K) *hist_K = ch->d_hist_K;
    if (hist_cells) *hist_cells = ch->d_hist_cells;
    if (hist_phi) *hist_phi = ch->d_hist_phi;
    if (hist_ptS) *hist_ptS = ch->d_hist_ptS;
    if (hist_iter) *hist_iter = ch->d_hist_iter;
    if (hist_action) *hist_action = ch->d_hist_action;
    if (hist_accept) *hist_accept = ch->d_hist_accept;
    if (hist_next_action) *hist_next_action = ch->d_hist_next;
    return TONGA_OK;
}

extern "C" int tonga_chains_kcap(const tonga_chains *ch) { return ch ? ch->KC : 0; }

extern "C" int tonga_chains_sampler(const tonga_chains *ch) { return !ch ? 0 : (ch->streamed ? TONGA_SAMPLER_STREAMED : (ch->wide ? TONGA_SAMPLER_WIDE : TONGA_SAMPLER_RESIDENT)); }

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
