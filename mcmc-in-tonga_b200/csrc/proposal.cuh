// proposal.cuh -- the proposal draw / a-priori checks and the acceptance rule of TD_inversion_function.jl:72-272, shared by the
// shared-memory-resident sampler (sampler_kernel.cuh) and the wide sampler (wide_sampler.cu) so that both follow the
// reference identically.
#pragma once
#include "tonga_internal.cuh"

namespace tg {

struct Prop {  // proposal of the current iteration
    int action, idx, do_eval, accept;
    double x, y, z, zeta, u;
    double aux;            // birth: czeta (:81); death: zetanew (:146)
    double ox, oy, oz;     // move: the old position (resident sampler: the nucleus array holds the proposed one during B..E)
    double ztag;           // zeta of the implicit new owner of tagged bytes (birth: zetanew; move: zeta[idx])
};

// ---- warp-cooperative v_nearest (MCsub.jl:247-263) over K nuclei, skipping index `skip` --------------------------
// SPACE (0 = shared-memory nuclei, 1 = global-memory nuclei) only separates the instantiations: each is called with
// pointers of one address space, so the compiler specialises the loads (LDS vs LDG) instead of emitting generic ones.
template <int SPACE>
__device__ __noinline__ int warp_nearest(const double *nx, const double *ny, const double *nz, int K, int skip, double x,
                                         double y, double z, int lane) {
    double best = 1e9;
    int bi = 0x7fffffff;
#pragma unroll 1  // (code size: the kernel is bound by instruction supply; K <= 32 makes this a single trip anyway)
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

// Standard normal deviate from one uniform u in [0,1) by inversion: Wichura's algorithm AS 241 (PPND16, relative accuracy
// 1e-16): a ratio of two degree-7 polynomials for |u - 1/2| <= 0.425 (85 % of the draws, no transcendental function), and
// the same in r = sqrt(-log(tail probability)) outside.  It replaced a Box-Muller pair (log + sqrt + sincospi on the serial
// proposal path of the sampler: 3.5 k of 97 k cycles per iteration).  Inlined on purpose: as a separate function the kernel's
// instruction-cache hit rate fell from 87 % to 80 % (profiles/README.md, "instruction cache").  The reference draws `rand(Normal(mu, sigma))`
// (TD_inversion_function.jl:82,188,226-228) from Julia's own generator, whose stream is not reproducible anyway;
// parity is established by replaying recorded proposals.
__device__ __forceinline__ double poly7(const double (&c)[8], double r) {
    double v = c[7];
#pragma unroll
    for (int k = 6; k >= 0; k--) v = fma(v, r, c[k]);
    return v;
}
__device__ __forceinline__ double normal_icdf(double u) {
    const double q = u - 0.5;
    if (fabs(q) <= 0.425) {
        const double a[8] = {3.3871328727963666080e0, 1.3314166789178437745e+2, 1.9715909503065514427e+3, 1.3731693765509461125e+4,
                             4.5921953931549871457e+4, 6.7265770927008700853e+4, 3.3430575583588128105e+4, 2.5090809287301226727e+3};
        const double b[8] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
                             2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4, 5.2264952788528545610e+3};
        const double r = 0.180625 - q * q;
        return q * poly7(a, r) / poly7(b, r);
    }
    const double tail = (q < 0.0) ? u + 0x1.0p-54 : 1.0 - u;  // in (0, 0.075]; never 0
    double r = sqrt(-log(tail));
    double v;
    if (r <= 5.0) {
        const double c[8] = {1.42343711074968357734e0, 4.63033784615654529590e0, 5.76949722146069140550e0, 3.64784832476320460504e0,
                             1.27045825245236838258e0, 2.41780725177450611770e-1, 2.27238449892691845833e-2, 7.74545014278341407640e-4};
        const double d[8] = {1.0, 2.05319162663775882187e0, 1.67638483018380384940e0, 6.89767334985100004550e-1,
                             1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4, 1.05075007164441684324e-9};
        r -= 1.6;
        v = poly7(c, r) / poly7(d, r);
    } else {
        const double e[8] = {6.65790464350110377720e0, 5.46378491116411436990e0, 1.78482653991729133580e0, 2.96560571828504891230e-1,
                             2.65321895265761230930e-2, 1.24266094738807843860e-3, 2.71155556874348757815e-5, 2.01033439929228813265e-7};
        const double f[8] = {1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
                             7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7, 2.04426310338993978564e-15};
        r -= 5.0;
        v = poly7(e, r) / poly7(f, r);
    }
    return q < 0.0 ? -v : v;
}

// IEEE division / natural logarithm as real function calls: the acceptance rule below has ~20 division sites over its prior
// variants and is executed once per iteration; inlined they were 1.5 k SASS instructions of a kernel that is bound by
// instruction supply (profiles/README.md).  Same results as the inlined operators.
static __device__ __noinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
static __device__ __noinline__ double log_fn(double a) { return log(a); }

// The state-independent part of one iteration's random numbers ("raw draw"): Philox4x32-10 keyed by `seed`, counter =
// (iteration, global chain id, slot 0..3) -> 8 uniforms uu[0..7]; action from uu[0] (rand(1:4), TD_inversion_function.jl:72),
// uu[1..3] position / index uniforms, three standard normals by inversion of uu[4..6], uu[7] the accept uniform.
// The resident sampler pre-generates these for a whole launch (tg_pregen_kernel, one thread per chain-iteration); the
// wide / streamed samplers compute them with the lanes of a warp (draw_proposal).  Both give the same values.
struct RawDraw {
    double act, u1, u2, u3, n0, n1, n2, u7;
};
static_assert(sizeof(RawDraw) == 64, "RawDraw is read as 8 doubles, one per lane");

__device__ __forceinline__ int action_of(double u0, int n_actions) {
    const int nact = n_actions >= 4 ? n_actions : 4;
    const int act0 = 1 + (int)floor(u0 * nact);  // rand(1:4), TD_inversion_function.jl:72
    return act0 > nact ? nact : act0;
}

__device__ __forceinline__ void raw_draw_thread(RawDraw &rd, unsigned long long seed, long long iter, unsigned long long gid, int n_actions) {
    const Philox philox{(uint32_t)seed, (uint32_t)(seed >> 32)};
    double uu[8];
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        uint32_t w[4];
        philox((uint32_t)iter, (uint32_t)((unsigned long long)iter >> 32), (uint32_t)gid, (uint32_t)s | ((uint32_t)(gid >> 32) << 8), w);
        uu[2 * s] = u53(w[0], w[1]);
        uu[2 * s + 1] = u53(w[2], w[3]);
    }
    rd.act = (double)action_of(uu[0], n_actions);
    rd.u1 = uu[1]; rd.u2 = uu[2]; rd.u3 = uu[3];
    rd.n0 = normal_icdf(uu[4]); rd.n1 = normal_icdf(uu[5]); rd.n2 = normal_icdf(uu[6]);
    rd.u7 = uu[7];
}

// Warp-collective (all 32 lanes call it with identical scalars): builds the proposal of one iteration from the raw draw
// (mode 0; pr.action preset) or from a recorded proposal (mode 1; pr.action / idx / x / y / z / zeta / u preset), applies
// the a-priori checks of the reference and evaluates v_nearest at the new / killed nucleus.  sig_zeta = zeta_scale * sig / 100
// (TD_inversion_function.jl:22, hoisted by the caller).  Returns `valid` (1 = the candidate must be evaluated).
template <int SPACE>
__device__ __forceinline__ int assemble_proposal(Prop &pr, int mode, double u1, double u2, double u3, double n0, double n1, double n2, double u7,
                                                 const double *nx, const double *ny, const double *nz, const double *nzeta, int K, double noise,
                                                 const tonga_params &pm, double sig_zeta, int lane) {
    const int act = pr.action;
    const double s100 = div_rn(pm.sig, 100);
    int valid = 0;
    if (act == 1) {  // ---- birth :76-125
        if (K < pm.max_cells) {
            if (mode == 0) {
                pr.x = u1 * (pm.xmax - pm.xmin) + pm.xmin;  // :78
                pr.y = u2 * (pm.ymax - pm.ymin) + pm.ymin;  // :79
                pr.z = u3 * (pm.zmax - pm.zmin) + pm.zmin;  // :80
            }
            const int ci = warp_nearest<SPACE>(nx, ny, nz, K, -1, pr.x, pr.y, pr.z, lane);  // :81
            const double czeta = ci < 0 ? 0.0 : nzeta[ci];
            pr.aux = czeta;
            if (mode == 0) {
                pr.zeta = czeta + sig_zeta * n0;  // :82
                pr.u = u7;                        // :121
            }
            if (pm.prior == 1) valid = (pr.zeta > 0 && pr.zeta < pm.zeta_scale);  // :92
            else if (pm.prior == 2) valid = 1;
            else valid = (pr.zeta > 0);  // :111
        }
    } else if (act == 2) {  // ---- death :126-181
        if (K > pm.min_cells) {
            if (mode == 0) {
                const int k = (int)floor(u1 * K);  // :128
                pr.idx = k >= K ? K - 1 : k;
                pr.u = u7;  // :176
            }
            const int kill = pr.idx;
            if (kill >= 0 && kill < K) {
                const int zi = warp_nearest<SPACE>(nx, ny, nz, K, kill, nx[kill], ny[kill], nz[kill], lane);  // :146
                pr.aux = zi < 0 ? 0.0 : nzeta[zi];
                valid = (pm.prior == 3) ? (pr.aux > 0) : 1;  // :165
            }
        }
    } else if (act == 3) {  // ---- change :183-218
        if (mode == 0) {
            const int k = (int)floor(u1 * K);  // :184
            pr.idx = k >= K ? K - 1 : k;
            pr.zeta = nzeta[pr.idx] + sig_zeta * n0;  // :188
            pr.u = u7;                                // :214
        }
        if (pr.idx >= 0 && pr.idx < K) {
            if (pm.prior == 1) valid = (pr.zeta > 0 && pr.zeta < pm.zeta_scale);  // :195
            else if (pm.prior == 2) valid = 1;
            else valid = (pr.zeta > 0);  // :206 (alpha = 0 otherwise)
            // the reference evaluates first (:191) but discards the result when invalid
        }
    } else if (act == 4) {  // ---- move :220-251
        if (K > 0) {
            if (mode == 0) {
                const int k = (int)floor(u1 * K);  // :222
                pr.idx = k >= K ? K - 1 : k;
                pr.x = nx[pr.idx] + (s100 * (pm.xmax - pm.xmin)) * n0;  // :30,:226
                pr.y = ny[pr.idx] + (s100 * (pm.ymax - pm.ymin)) * n1;  // :31,:227
                pr.z = nz[pr.idx] + (s100 * (pm.zmax - pm.zmin)) * n2;  // :32,:228
                pr.u = u7;                                                         // :247
            }
            if (pr.idx >= 0 && pr.idx < K) {
                valid = (pr.x >= pm.xmin && pr.x <= pm.xmax && pr.y >= pm.ymin && pr.y <= pm.ymax && pr.z >= pm.zmin &&
                         pr.z <= pm.zmax);  // :230-232
                if (valid) { pr.ox = nx[pr.idx]; pr.oy = ny[pr.idx]; pr.oz = nz[pr.idx]; }
            }
        }
    } else if (act == 5) {  // ---- sigma :252-272 (dead code in the reference; extension)
        if (mode == 0) {
            pr.zeta = noise + div_rn(pm.max_sig * pm.sig, 100) * n0;  // :23,:254
            pr.u = u7;
        }
        valid = (pr.zeta > 0 && pr.zeta < pm.max_sig);  // :257
    }
    pr.do_eval = valid;
    return valid;
}

// Warp-collective: raw draw with the lanes of the warp (mode 0) or a recorded proposal (mode 1), then assemble_proposal.
// pr.action / idx / x / y / z / zeta / u / aux are set on every lane.
template <int SPACE>
__device__ __forceinline__ int draw_proposal(Prop &pr, int mode, const tonga_proposal *rec_in, unsigned long long seed, long long iter,
                                             unsigned long long gid, const double *nx, const double *ny, const double *nz,
                                             const double *nzeta, int K, double noise, const tonga_params &pm, double sig_zeta, int lane) {
    pr.do_eval = 0; pr.accept = 0; pr.idx = 0;
    pr.x = pr.y = pr.z = pr.zeta = pr.u = pr.aux = pr.ox = pr.oy = pr.oz = pr.ztag = 0.0;
    double uu[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nrm[3] = {0, 0, 0};
    if (mode == 0) {
        const Philox philox{(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t w[4] = {0, 0, 0, 0};
        if (lane < 4) philox((uint32_t)iter, (uint32_t)((unsigned long long)iter >> 32), (uint32_t)gid, (uint32_t)lane | ((uint32_t)(gid >> 32) << 8), w);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t w0 = __shfl_sync(0xffffffffu, w[0], s), w1 = __shfl_sync(0xffffffffu, w[1], s);
            const uint32_t w2 = __shfl_sync(0xffffffffu, w[2], s), w3 = __shfl_sync(0xffffffffu, w[3], s);
            uu[2 * s] = u53(w0, w1);
            uu[2 * s + 1] = u53(w2, w3);
        }
        pr.action = action_of(uu[0], pm.n_actions);
        // the (up to) three Gaussian steps of the proposal: one inversion, lanes 0..2 in parallel
        const double nl = normal_icdf(lane == 0 ? uu[4] : (lane == 1 ? uu[5] : uu[6]));
        nrm[0] = __shfl_sync(0xffffffffu, nl, 0);
        nrm[1] = __shfl_sync(0xffffffffu, nl, 1);
        nrm[2] = __shfl_sync(0xffffffffu, nl, 2);
    } else {
        const tonga_proposal rec = *rec_in;
        pr.action = rec.action; pr.idx = rec.idx; pr.x = rec.x; pr.y = rec.y; pr.z = rec.z; pr.zeta = rec.zeta; pr.u = rec.u;
    }
    return assemble_proposal<SPACE>(pr, mode, uu[1], uu[2], uu[3], nrm[0], nrm[1], nrm[2], uu[7], nx, ny, nz, nzeta, K, noise, pm, sig_zeta, lane);
}

// The acceptance rule, TD_inversion_function.jl:96-97,107-108,113-114 (birth), :151-152,160-162,166-168 (death), :196,202-203,207-208
// (change), :241-242 (move), :264-267 (sigma, extension), with the reference's operation order: every variant has the form
// (ratio * constant) * exp(argument) or exp(argument), so the branches only build the three pieces and ONE exp follows.
// K = nCells of the CURRENT model, zeta_idx = current zeta[idx] (death / change), beta = inverse temperature (1 = reference),
// R = number of data.
__device__ __forceinline__ int accept_decision(const Prop &pr, int K, double phi, double phin, double zeta_idx, double noise, double beta,
                                               int R, const tonga_params &pm, double sig_zeta) {
    const double SQ2PI = 2.5066282746310002;  // sqrt(2 * 3.141592653589793), correctly rounded
    const int act = pr.action;
    const double K0 = (double)K, zs = pm.zeta_scale;
    const double zn = pr.zeta, aux = pr.aux, u = pr.u;
    const double dphi2 = beta * ((phin - phi) / 2);
    if (act == 5) {  // sigma: log form :264-267
        double la = beta * (log_fn(div_rn(noise, zn)) * (double)R) - dphi2;  // beta = 1: TD_inversion_function.jl:264
        la = (la != la) ? la : (la < 0.0 ? la : 0.0);
        return log_fn(u) <= la;
    }
    double pre = 1.0, arg = -dphi2;  // change (prior 1) :196 and move :241-242: alpha = exp(-dphi2)
    if (act == 1) {
        const double g = div_rn((aux - zn) * (aux - zn), 2 * (sig_zeta * sig_zeta));
        const double ratio = div_rn(K0, K0 + 1);
        if (pm.prior == 1) { pre = ratio * div_rn(sig_zeta * SQ2PI, zs); arg = g - dphi2; }                                   // :96-97
        else if (pm.prior == 2) { pre = ratio * div_rn(sig_zeta, zs); arg = -div_rn(zn * zn, zs * zs) + g - dphi2; }          // :107-108
        else { pre = ratio * div_rn(SQ2PI * sig_zeta, zs); arg = -div_rn(zn, zs) + g - dphi2; }                               // :113-114
    } else if (act == 2) {
        const double zk = zeta_idx;
        const double g = div_rn((zk - aux) * (zk - aux), 2 * (sig_zeta * sig_zeta));
        const double ratio = div_rn(K0, K0 - 1);
        if (pm.prior == 1) { pre = ratio * div_rn(zs, sig_zeta * SQ2PI); arg = -g - dphi2; }                                  // :151-152
        else if (pm.prior == 2) { pre = ratio * div_rn(zs, sig_zeta); arg = div_rn(zk * zk, 2 * (zs * zs)) - g - dphi2; }     // :160-162
        else { pre = ratio * div_rn(zs, SQ2PI * sig_zeta); arg = div_rn(zk, zs) - g - dphi2; }                                // :166-168
    } else if (act == 3) {
        const double zo = zeta_idx;
        if (pm.prior == 2) arg = div_rn(zo * zo - zn * zn, 2 * (zs * zs)) - dphi2;  // :202-203
        else if (pm.prior == 3) arg = div_rn(zo - zn, zs) - dphi2;                 // :207-208
    }
    const double e = exp(arg);
    const double alpha = (act <= 2) ? pre * e : e;  // (ratio * constant) * exp(...), as the reference associates it
    return u < jl_min1(alpha);
}

// the same as a real function call (the resident sampler keeps its per-iteration code path short: instruction supply)
static __device__ __noinline__ int accept_decision_call(const Prop &pr, int K, double phi, double phin, double zeta_idx, double noise, double beta,
                                                        int R, const tonga_params &pm, double sig_zeta) {
    return accept_decision(pr, K, phi, phin, zeta_idx, noise, beta, R, pm, sig_zeta);
}

}  // namespace tg
