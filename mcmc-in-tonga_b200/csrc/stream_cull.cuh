// stream_cull.cuh -- the streamed sampler's point pass with spatial culling (BASELINE config 3: 100k rays, 2e7 points, up to
// 2000 nuclei).  The tile kernel of wide_kernels.cuh streams every point of every active chain per proposal (18 B per point);
// but a birth / death / change / move only concerns the rays that come close to the nucleus: a point p can switch to a new
// position c only if |p - c|^2 < d(p, owner(p)), and a point owned by nucleus k lies within sqrt(d(p, owner(p))) of k.
// Kept per chain and ray: dmax = an upper bound of the cached squared owner distance over the ray's points (4 B per ray, L2-
// resident for the active chains); kept per ray set: a bounding sphere for every run of 32 consecutive ray points.  Per proposal:
//   tg_cull_kernel          one thread per (active chain, ray): the ray is a CANDIDATE if one of its spheres comes closer to the
//                           proposal's position(s) than sqrt(dmax) (conservative: generous rounding margins) -> per-chain lists;
//   tg_cull_prefix_kernel   exclusive prefix over the active chains' list lengths -> one flat work list;
//   tg_stream2_kernel<0>    candidate pass, ONE WARP PER CANDIDATE RAY (persistent warps striding the flat list): the same
//                           classification as the tile kernel (fl32 screening against the cached owner distance, exact FP64 inside
//                           the error band, orphans rescanned over the candidate model), candidate owners in the warp's shared
//                           memory; a ray whose owners / zeta change is re-integrated (canonical tstar_g8 order), its (t*, misfit
//                           term) go to tstar_c / term_c and the ray joins the chain's DIRTY list.  Read-only on the chain state;
//   tg_wide_accept_kernel   sums term_c in the canonical order as before; then walks the dirty list: accept -> t*, term of the dirty
//                           rays become the chain's; reject -> tstar_c / term_c are restored (invariant: they equal the chain's
//                           t* / term outside an iteration);
//   tg_stream2_kernel<1>    commit pass over the same work list for the accepted chains: owners, owner-distance cache, new dmax;
//   tg_renumber_kernel      accepted deaths: deleteat! renumbering of the owners above the killed index (the only full sweep left).
// Results are bit-identical to the tile kernel / the resident sampler (same classification, same canonical sums).
#pragma once
#include "wide_kernels.cuh"

namespace tg {

struct CullArgs {
    StreamArgs s;            // geometry, candidate models, chain state (tiles unused)
    const float4 *sub;       // bounding sphere (centre, radius) of every 32-point run, rays in sorted order
    const float4 *raysph;    // [R] bounding sphere of the whole ray (first, coarse test)
    const int32_t *sub_off;  // [R+1]
    float *dmax;             // [n][Rp]
    double *term;            // [n][Rp] misfit term of the chain's current t* (what term_c is restored to)
    int32_t *cand;           // [n][R] candidate rays of this proposal
    int32_t *ncand;          // [n]
    int32_t *clist;          // [n][R] rays in which the candidate pass found something to commit (the commit pass's work list)
    int32_t *nclist;         // [n] its length; the accept kernel turns it into ncommit (0 for rejected chains / a change)
    int32_t *ncommit;        // [n]
    int32_t *work_off2;      // [n+1] prefix of ncommit over the active list
    int32_t *dirty;          // [n][R] rays whose t* differs under the candidate model
    int32_t *ndirty;         // [n]
    int32_t *work_off;       // [n+1] prefix of ncand over the active list
    int R, ray0, ray1;       // all rays; own rays (ray sharding)
    int maxn;                // shared-memory capacity per warp in points (>= max points per ray + 8)
    long long p0, p1;        // own points (renumber sweep)
};

constexpr int CULL_THREADS = 256;
constexpr int S2_THREADS = 256;
constexpr int S2_WARPS = S2_THREADS / 32;
#ifndef S2_MIN_CTAS
#define S2_MIN_CTAS 4
#endif
constexpr int S2_LIST = 256;  // nuclei near the killed / moved nucleus that an orphan is rescanned against (else: all)
__host__ __device__ inline size_t s2_warp_smem(int maxn) {  // term64[maxn] | owner16[maxn] | queue16[maxn] | list16[S2_LIST] | chg32[maxn/32 + 1] | cnt[2]
    return (size_t)maxn * 12 + (size_t)S2_LIST * 2 + (size_t)(maxn / 32 + 1) * 4 + 8;
}

// ---- static geometry: bounding spheres of the 32-point runs (one thread per ray) ---------------------------------------------
__global__ void tg_sub_spheres_kernel(int R, const int32_t *__restrict__ ray_off, const int32_t *__restrict__ sub_off, const double *__restrict__ px,
                                      const double *__restrict__ py, const double *__restrict__ pz, float4 *__restrict__ sub, float4 *__restrict__ raysph) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int q0 = ray_off[r], n = ray_off[r + 1] - q0;
    for (int s0 = -1, s = sub_off[r]; s0 < n; s0 += 32) {  // first round (s0 = -1): the whole ray
        const bool whole = s0 < 0;
        if (whole) s0 = 0;
        const int m = whole ? n : min(32, n - s0);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int j = 0; j < m; j++) {
            const double v[3] = {px[q0 + s0 + j], py[q0 + s0 + j], pz[q0 + s0 + j]};
            for (int a = 0; a < 3; a++) { lo[a] = fmin(lo[a], v[a]); hi[a] = fmax(hi[a], v[a]); }
        }
        const double c[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
        double rad2 = 0.0;
        for (int j = 0; j < m; j++) {
            const double dx = px[q0 + s0 + j] - c[0], dy = py[q0 + s0 + j] - c[1], dz = pz[q0 + s0 + j] - c[2];
            rad2 = fmax(rad2, dx * dx + dy * dy + dz * dz);
        }
        // rounded outwards: the fl32 centre is off by <= 1 ulp of the largest coordinate, the radius by 1 ulp
        const double cm = fmax(fmax(fabs(c[0]), fabs(c[1])), fabs(c[2]));
        const float radf = (float)(sqrt(rad2) * (1.0 + 1e-6) + 4e-7 * cm + 1e-30);
        const float4 sp = make_float4((float)c[0], (float)c[1], (float)c[2], radf);
        if (whole) { raysph[r] = sp; s0 = -32; }
        else sub[s++] = sp;
    }
}

// ---- (re)establish the culling state of a batch from its chain state: dmax, term; tstar_c = t*, term_c = term --------------
__global__ void __launch_bounds__(256) tg_stream_init_kernel(int R, int Rp, long long Ppad, const int32_t *__restrict__ ray_off, const float *__restrict__ dcache,
                                                             const double *__restrict__ tstar, const double *__restrict__ tS, const double *__restrict__ sig,
                                                             const double *__restrict__ noise, float *__restrict__ dmax, double *__restrict__ term,
                                                             double *__restrict__ tstar_c, double *__restrict__ term_c) {
    const int chain = blockIdx.y, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= R) return;
    const int q0 = ray_off[r], n = ray_off[r + 1] - q0;
    float m = 0.0f;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, dcache[(size_t)chain * Ppad + q0 + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        const size_t i = (size_t)chain * Rp + r;
        dmax[i] = m;
        const double t = tstar[i], v = misfit_term(t, tS[r], sig[r], noise[chain]);
        term[i] = v; term_c[i] = v; tstar_c[i] = t;
    }
}

// ---- cull: which rays can the proposal concern at all? ----------------------------------------------------------------------
// A ray is skipped only if every one of its spheres is provably farther from the position than sqrt(dmax): lower bound of the
// distance = |c - centre| - radius, shrunk by 1e-5 relative and 1e-2 absolute (fl32 arithmetic errors are ~1e-7 relative), against
// dmax grown by 1e-4 relative (the cache holds fl32 roundings of the exact squared distances, ~1e-7 relative).
__device__ __forceinline__ bool sphere_near(const float4 s, float cx, float cy, float cz, float dmax_hi) {
    const float dx = s.x - cx, dy = s.y - cy, dz = s.z - cz;
    const float dist = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    const float lb = fmaxf(0.0f, (dist - s.w) * 0.99999f - 0.01f);
    return !(lb * lb > dmax_hi);  // (NaN -> near)
}
__global__ void __launch_bounds__(CULL_THREADS) tg_cull_kernel(const CullArgs a) {
    const int slot = blockIdx.y;
    if (slot >= *a.s.n_active) return;
    const int chain = a.s.active[slot];
    const Prop pr = a.s.props[chain];
    const int act = pr.action, lane = threadIdx.x & 31;
    const int r = a.ray0 + blockIdx.x * CULL_THREADS + threadIdx.x;
    bool cand = false;
    if (r < a.ray1) {
        const bool has_new = (act == 1 || act == 4), has_old = (act == 2 || act == 3 || act == 4);
        const float nx = (float)pr.x, ny = (float)pr.y, nz = (float)pr.z;
        const float ox = (float)pr.ox, oy = (float)pr.oy, oz = (float)pr.oz;  // killed / changed nucleus, old position of a moved one
        const float dm = a.dmax[(size_t)chain * a.s.Rp + r] * 1.0001f;
        const float4 rs = __ldg(a.raysph + r);
        const bool near_new = has_new && sphere_near(rs, nx, ny, nz, dm), near_old = has_old && sphere_near(rs, ox, oy, oz, dm);
        if (near_new || near_old) {
            for (int s = a.sub_off[r]; s < a.sub_off[r + 1] && !cand; s++) {
                const float4 sp = __ldg(a.sub + s);
                cand = (near_new && sphere_near(sp, nx, ny, nz, dm)) || (near_old && sphere_near(sp, ox, oy, oz, dm));
            }
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, cand);
    if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.ncand + chain, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) a.cand[(size_t)chain * a.R + base + __popc(m & ((1u << lane) - 1u))] = r;
    }
}

// exclusive prefix of the candidate counts over the active list (one CTA): work_off[0..n_active]
__global__ void __launch_bounds__(1024) tg_cull_prefix_kernel(const int32_t *__restrict__ active, const int32_t *__restrict__ n_active,
                                                              const int32_t *__restrict__ ncand, int32_t *__restrict__ work_off) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, na = *n_active;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < na; b += 1024) {
        const int i = b + tid;
        const int v = i < na ? ncand[active[i]] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            wsum[lane] = w;
        }
        __syncthreads();
        const int incl = x + (warp ? wsum[warp - 1] : 0) + carry;
        if (i < na) work_off[i] = incl - v;
        __syncthreads();
        if (tid == 1023) carry = incl;
        __syncthreads();
    }
    if (tid == 0) work_off[na] = carry;
}

// ---- one warp per candidate ray -------------------------------------------------------------------------------------------
template <bool COMMIT>
__global__ void __launch_bounds__(S2_THREADS, S2_MIN_CTAS) tg_stream2_kernel(const CullArgs ca) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StreamArgs &a = ca.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t FULL = 0xffffffffu;
    const int n_active = *a.n_active;
    const int32_t *woff = COMMIT ? ca.work_off2 : ca.work_off;
    const int32_t *wlist = COMMIT ? ca.clist : ca.cand;
    const int total = woff[n_active];
    unsigned char *wbase = smem_raw + (size_t)warp * ((s2_warp_smem(ca.maxn) + 15) & ~(size_t)15);
    double *s_term = reinterpret_cast<double *>(wbase);
    uint16_t *s_owner = reinterpret_cast<uint16_t *>(s_term + ca.maxn);
    uint16_t *s_queue = s_owner + ca.maxn;
    uint16_t *s_list = s_queue + ca.maxn;
    uint32_t *s_chg = reinterpret_cast<uint32_t *>(s_list + S2_LIST);
    int *s_cnt = reinterpret_cast<int *>(s_chg + ca.maxn / 32 + 1);
    const float ta = a.tol_alpha, tb = a.tol_beta2;

    constexpr int CH = 8;  // items per chunk: lanes 0..7 fetch the chunk's metadata (slot, ray, offsets) side by side
    for (int c0 = (blockIdx.x * S2_WARPS + warp) * CH; c0 < total; c0 += gridDim.x * S2_WARPS * CH) {
        int m_chain = 0, m_ci = 0, m_r = 0, m_q0 = 0, m_n = 0;
        {
            const int item = c0 + (lane & (CH - 1));
            if (item < total) {
                // item -> (active slot, index in the chain's candidate list): largest slot with work_off[slot] <= item
                int lo_s = 0, hi_s = n_active;
                while (hi_s - lo_s > 1) {
                    const int mid = (lo_s + hi_s) >> 1;
                    if (__ldg(woff + mid) <= item) lo_s = mid; else hi_s = mid;
                }
                m_chain = a.active[lo_s];
                m_ci = item - __ldg(woff + lo_s);
                m_r = wlist[(size_t)m_chain * ca.R + m_ci];
                m_q0 = a.ray_off[m_r];
                m_n = a.ray_off[m_r + 1] - m_q0;
            }
        }
#pragma unroll 1
    for (int ii = 0; ii < CH && c0 + ii < total; ii++) {
        const int chain = __shfl_sync(FULL, m_chain, ii), r = __shfl_sync(FULL, m_r, ii);
        const int q0 = __shfl_sync(FULL, m_q0, ii), n = __shfl_sync(FULL, m_n, ii);
        const Prop pr = a.props[chain];
        const int act = pr.action;  // (COMMIT: the work list holds accepted births / deaths / moves only)
        const int Kn = a.Kc[chain];
        const double *cc = a.cells_c + (size_t)chain * 4 * a.KC;
        const float *cf = a.cells_cf + (size_t)chain * 3 * a.KC;
        uint16_t *own = a.owner + (size_t)chain * a.Ppad;
        float *dc = a.dcache + (size_t)chain * a.Ppad;
        const int idx = pr.idx;
        const double cx = pr.x, cy = pr.y, cz = pr.z;
        const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
        const int newo = (act == 1) ? Kn - 1 : idx;
        const long long p0a = (long long)(q0 & ~3);
        const int lo = (int)(q0 - p0a), hi = lo + n;  // valid relative range [lo, hi)
        const int ngroups = (hi + 3) >> 2;
        __syncwarp();
        if (!COMMIT) for (int i = lane; i < (hi >> 5) + 1; i += 32) s_chg[i] = 0u;
        if (lane == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
        __syncwarp();
        float dmx = 0.0f;  // COMMIT: largest cached owner distance of the ray after the commit

        auto phase1 = [&](auto actc) {
            constexpr int ACT = decltype(actc)::value;
            for (int g = lane; g < ngroups; g += 32) {
                const long long p = p0a + 4 * g;
                const ushort4 ov = *reinterpret_cast<const ushort4 *>(own + p);
                const uint32_t o[4] = {ov.x, ov.y, ov.z, ov.w};
                uint32_t no[4] = {ov.x, ov.y, ov.z, ov.w};
                uint32_t valid = 0xFu;
                if (4 * g < lo) valid &= 0xFu << (lo - 4 * g);
                if (4 * g + 4 > hi) valid &= 0xFu >> (4 * g + 4 - hi);
                uint32_t chg = 0, wr = 0, orph = 0;
                float dnew[4] = {0.f, 0.f, 0.f, 0.f}, dfin[4] = {0.f, 0.f, 0.f, 0.f};
                if (ACT == 1 || ACT == 4) {
                    const float4 dv = *reinterpret_cast<const float4 *>(dc + p);
                    const float4 xv = *reinterpret_cast<const float4 *>(a.pxf + p), yv = *reinterpret_cast<const float4 *>(a.pyf + p),
                                 zv = *reinterpret_cast<const float4 *>(a.pzf + p);
                    const float d_o[4] = {dv.x, dv.y, dv.z, dv.w}, x[4] = {xv.x, xv.y, xv.z, xv.w}, y[4] = {yv.x, yv.y, yv.z, yv.w}, z[4] = {zv.x, zv.y, zv.z, zv.w};
                    uint32_t amb = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float d_c = dist2_f32(cxf, cyf, czf, x[q], y[q], z[q]);
                        dnew[q] = d_c;
                        const float diff = d_c - d_o[q], tol = fmaf(ta, d_c + d_o[q], tb);
                        const bool mine = (ACT == 4) && ((int)o[q] == idx);
                        orph |= mine ? (1u << q) : 0u;
                        chg |= (!mine && diff < -tol) ? (1u << q) : 0u;
                        amb |= (!mine && (a.exact_only || !(fabsf(diff) > tol))) ? (1u << q) : 0u;
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
                            const bool sw = (de_c < de_o) || (ACT == 4 && de_c == de_o && idx < (int)oo && oo != TG_NONE16S);
                            chg = sw ? (chg | (1u << q)) : (chg & ~(1u << q));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        no[q] = ((chg >> q) & 1u) ? (uint32_t)newo : o[q];
                        dfin[q] = ((chg >> q) & 1u) ? dnew[q] : d_o[q];
                    }
                    wr = chg;
                } else if (ACT == 2) {
#pragma unroll
                    for (int q = 0; q < 4; q++) orph |= ((int)o[q] == idx) ? (1u << q) : 0u;
                    orph &= valid;
                    if (COMMIT) {  // the unchanged points' distances still count for the ray's new dmax
                        const float4 dv = *reinterpret_cast<const float4 *>(dc + p);
                        dfin[0] = dv.x; dfin[1] = dv.y; dfin[2] = dv.z; dfin[3] = dv.w;
                    } else {       // candidate owners in the NEW numbering (deleteat!), as the candidate model has them
#pragma unroll
                        for (int q = 0; q < 4; q++) no[q] = (o[q] != TG_NONE16S && (int)o[q] > idx) ? o[q] - 1 : o[q];
                    }
                } else {  // change: owners stay, the rays through the cell are re-integrated
#pragma unroll
                    for (int q = 0; q < 4; q++) chg |= ((int)o[q] == idx) ? (1u << q) : 0u;
                    chg &= valid;
                }
                if ((ACT == 2 || ACT == 4) && orph) {
#pragma unroll 1
                    for (int q = 0; q < 4; q++)
                        if ((orph >> q) & 1u) s_queue[atomicAdd(&s_cnt[0], 1)] = (uint16_t)(4 * g + q);
                }
                if (!COMMIT) {
                    *reinterpret_cast<ushort4 *>(s_owner + 4 * g) = make_ushort4((uint16_t)no[0], (uint16_t)no[1], (uint16_t)no[2], (uint16_t)no[3]);
                    if (chg) atomicOr(&s_chg[g >> 3], chg << ((4 * g) & 31));
                } else {
                    if (wr) {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if ((wr >> q) & 1u) { own[p + q] = (uint16_t)no[q]; dc[p + q] = dnew[q]; }  // (only births / moves set wr)
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (((valid & ~orph) >> q) & 1u) dmx = fmaxf(dmx, dfin[q]);
                }
            }
        };
        if (act == 1) phase1(std::integral_constant<int, 1>{});
        else if (act == 4) phase1(std::integral_constant<int, 4>{});
        else if (act == 2) phase1(std::integral_constant<int, 2>{});
        else phase1(std::integral_constant<int, 3>{});
        __syncwarp();
        // ---- orphans (death / move): nearest nucleus of the candidate model, ONE LANE PER ORPHAN.  The new owner of a point of
        // the killed / moved nucleus k is a Voronoi neighbour of k: with r = |p - k| <= sqrt(dmax of the ray) and delta = distance
        // from k to its nearest other nucleus j0, |p - j*| <= |p - j0| <= r + delta, hence |k - j*| <= 2 r + delta.  The warp lists
        // the nuclei inside that radius (grown by 0.1 % + 1 km, far more than the screening band) once per ray -- ascending, so the
        // lowest index still wins exact ties -- and every orphan scans the list only; a moved nucleus is always listed.
        const int nq = s_cnt[0];
        if (nq > 0) {
            const float kx = (float)pr.ox, ky = (float)pr.oy, kz = (float)pr.oz;  // k before the proposal (the commit pass runs after the model was committed)
            const float rmax = sqrtf(ca.dmax[(size_t)chain * a.Rp + r] * 1.0001f);
            float dl2 = 3.0e38f;
            for (int j = lane; j < Kn; j += 32) {
                const float d = dist2_f32(cf[j], cf[a.KC + j], cf[2 * a.KC + j], kx, ky, kz);
                if (!(act == 4 && j == idx)) dl2 = fminf(dl2, d);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dl2 = fminf(dl2, __shfl_xor_sync(FULL, dl2, o));
            const float rho = (2.0f * rmax + sqrtf(dl2)) * 1.001f + 1.0f, rho2 = rho * rho;
            int nl = 0;
            for (int j0 = 0; j0 < Kn && nl <= S2_LIST; j0 += 32) {
                const int j = j0 + lane;
                const bool in = j < Kn && ((act == 4 && j == idx) || !(dist2_f32(cf[j], cf[a.KC + j], cf[2 * a.KC + j], kx, ky, kz) > rho2));
                const uint32_t m = __ballot_sync(FULL, in);
                const int pos = nl + __popc(m & ((1u << lane) - 1u));
                if (in && pos < S2_LIST) s_list[pos] = (uint16_t)j;
                nl += __popc(m);
            }
            const bool all = nl > S2_LIST || !(rho2 < 3.0e38f);  // too many (or no bound): scan every nucleus
            const int nn = all ? Kn : nl;
            __syncwarp();
            for (int e0 = 0; e0 < nq; e0 += 32) {
                const int e = e0 + lane;
                const bool on = e < nq;
                const int j = on ? (int)s_queue[e] : 0;
                const long long p = p0a + j;
                int bi = -2;
                float dbest = 1e9f;
                if (on && !a.exact_only) {
                    const float x = a.pxf[p], y = a.pyf[p], z = a.pzf[p];
                    float d1 = 1e9f, d2 = 1e9f;
                    int i1 = -1;
                    for (int i = 0; i < nn; i++) {
                        const int jj = all ? i : (int)s_list[i];
                        const float d = dist2_f32(__ldg(cf + jj), __ldg(cf + a.KC + jj), __ldg(cf + 2 * a.KC + jj), x, y, z);
                        const bool lt = d < d1;
                        d2 = lt ? d1 : fminf(d2, d);
                        i1 = lt ? jj : i1;
                        d1 = lt ? d : d1;
                    }
                    dbest = d1;
                    const float tol = fmaf(ta, d1 + d2, tb);
                    if (d2 - d1 > tol) bi = i1;
                }
                if (on && bi == -2) {  // exact FP64; ascending indices, strict <: the lowest index wins ties (MCsub.jl:255)
                    const double x = a.px[p], y = a.py[p], z = a.pz[p];
                    double best = 1e9;
                    int b = -1;
                    for (int i = 0; i < nn; i++) {
                        const int jj = all ? i : (int)s_list[i];
                        const double d = dist2_exact(cc[jj], cc[a.KC + jj], cc[2 * a.KC + jj], x, y, z);
                        if (d < best) { best = d; b = jj; }
                    }
                    bi = b;
                    dbest = (float)best;
                }
                if (on) {
                    if (COMMIT) {
                        dmx = fmaxf(dmx, bi < 0 ? 1e9f : dbest);
                        // a death's owners stay in the OLD numbering here: tg_renumber_kernel shifts every owner above the killed index afterwards
                        const int bo = (act == 2 && bi >= idx) ? bi + 1 : bi;
                        own[p] = bi < 0 ? (uint16_t)TG_NONE16S : (uint16_t)bo;
                        dc[p] = bi < 0 ? 1e9f : dbest;
                    } else {
                        s_owner[j] = bi < 0 ? (uint16_t)TG_NONE16S : (uint16_t)bi;
                        if (act == 2 || bi != idx) atomicOr(&s_chg[j >> 5], 1u << (j & 31));  // a point that stays with the moved nucleus keeps its zeta
                    }
                }
            }
        }
        __syncwarp();
        if (COMMIT) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dmx = fmaxf(dmx, __shfl_xor_sync(FULL, dmx, o));
            if (lane == 0) ca.dmax[(size_t)chain * a.Rp + r] = dmx;
            continue;
        }
        // ---- did anything change?  (a move's orphans always commit: their owner distance changes)
        bool anyc = false;
        for (int i = lane; i < (hi >> 5) + 1; i += 32) anyc |= (s_chg[i] != 0u);
        const bool dirty = __any_sync(FULL, anyc);
        if (lane == 0 && act != 3 && (dirty || nq > 0)) ca.clist[(size_t)chain * ca.R + atomicAdd(ca.nclist + chain, 1)] = r;
        if (!dirty) continue;
        // ---- t* of the ray under the candidate model (canonical order: lanes 0..7)
        const double *zc = cc + 3 * (size_t)a.KC;
        const int nseg = n > 1 ? n - 1 : 0;
        const bool on = lane < 8;
        const int trip = (nseg + 7) >> 3;
        const uint16_t *ow = s_owner + lo;
        // the segment terms by ALL lanes (loads of several rounds in flight), then the ordered canonical sum from shared memory
#pragma unroll 4
        for (int j = lane; j < nseg; j += 32) {
            const uint16_t oa = ow[j], ob = ow[j + 1];
            s_term[j] = seg_term(__ldg(a.dt + q0 + j), oa == TG_NONE16S ? 0.0 : __ldg(zc + oa), ob == TG_NONE16S ? 0.0 : __ldg(zc + ob));
        }
        __syncwarp();
        const double t = tstar_g8(on ? nseg : 0, trip, lane & 7, [&](int j) { return s_term[j]; });
        if (lane == 0) {
            const size_t i = (size_t)chain * a.Rp + r;
            const double term = misfit_term(t, a.tS[r], a.sig[r], a.noise[chain]);
            a.tstar_c[i] = t; a.term_c[i] = term;
            const int di = atomicAdd(ca.ndirty + chain, 1);
            ca.dirty[(size_t)chain * ca.R + di] = r;
            for (int w = 0; w < a.sh.world; w++) {  // ray-sharded: the record goes into this rank's section of every other block (P2P stores)
                if (w == a.sh.rank) continue;
                unsigned char *sec = mb_section(a.sh, w, a.sh.rank);
                const size_t k = (size_t)chain * a.sh.mb_cap + di;
                mb_ray(a.sh, sec)[k] = r; mb_t(a.sh, sec)[k] = t; mb_term(a.sh, sec)[k] = term;
            }
        }
    }
    }
}

// ---- accepted deaths: deleteat! renumbering (TD_inversion_function.jl:132-135), the one remaining sweep over all points -----
__global__ void __launch_bounds__(256) tg_renumber_kernel(const CullArgs ca) {
    const StreamArgs &a = ca.s;
    const int chain = blockIdx.y;
    if (!a.accept_flag[chain]) return;
    const Prop pr = a.props[chain];
    if (pr.action != 2 || !pr.do_eval) return;
    const uint32_t idx = (uint32_t)pr.idx;
    uint16_t *own = a.owner + (size_t)chain * a.Ppad;
    const long long g0 = ca.p0 >> 3, g1 = (ca.p1 + 7) >> 3;  // groups of 8 owners (16 B); the shards' ranges never share a written element below
    for (long long g = g0 + blockIdx.x * 256ll + threadIdx.x; g < g1; g += (long long)gridDim.x * 256) {
        uint4 v = *reinterpret_cast<const uint4 *>(own + 8 * g);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t a0 = w[k] & 0xFFFFu, a1 = w[k] >> 16;
            if (a0 != TG_NONE16S && a0 > idx) { a0--; any = true; }
            if (a1 != TG_NONE16S && a1 > idx) { a1--; any = true; }
            w[k] = a0 | (a1 << 16);
        }
        if (any) *reinterpret_cast<uint4 *>(own + 8 * g) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- verify helper: the culling state is consistent with the chain state (counts violations) -------------------------------
__global__ void __launch_bounds__(256) tg_stream_check_kernel(int R, int Rp, long long Ppad, int ray0, int ray1, const int32_t *__restrict__ ray_off,
                                                              const float *__restrict__ dcache, const double *__restrict__ tstar, const double *__restrict__ tS,
                                                              const double *__restrict__ sig, const double *__restrict__ noise, const float *__restrict__ dmax,
                                                              const double *__restrict__ term, const double *__restrict__ tstar_c, const double *__restrict__ term_c,
                                                              unsigned long long *mism) {
    const int chain = blockIdx.y, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= R) return;
    const size_t i = (size_t)chain * Rp + r;
    unsigned long long bad = 0;
    if (r >= ray0 && r < ray1) {  // dmax bounds the owner-distance cache (own rays)
        const int q0 = ray_off[r], n = ray_off[r + 1] - q0;
        float m = 0.0f;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, dcache[(size_t)chain * Ppad + q0 + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        bad += !(dmax[i] >= m);
    }
    if (lane == 0) {
        const double t = tstar[i], v = misfit_term(t, tS[r], sig[r], noise[chain]);
        bad += (__double_as_longlong(term[i]) != __double_as_longlong(v)) + (__double_as_longlong(term_c[i]) != __double_as_longlong(v)) +
               (__double_as_longlong(tstar_c[i]) != __double_as_longlong(t));
        if (bad) atomicAdd(mism, bad);
    }
}

}  // namespace tg
