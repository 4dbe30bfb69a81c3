/*
 * tonga_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See tonga_oracle.h.
 *
 * PARITY: the reference (pure Julia) cannot run here and ships no golden vectors for the forward model, so owners and t* are
 * "parity unpinned" (pinned only by the independent NumPy twin, KATs and the prior-recovery test).  orc_misfit (phi, likelihood)
 * IS pinned by reference-produced values: model.jld's 100 stored phi are reproduced bit for bit (tests/test_oracle.py).
 *
 * Compile: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared  (oracle/Makefile)
 * All file:line citations are relative to /root/reference.
 */
#include "tonga_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.141592653589793 /* Float64(pi) */

/* Julia's min(a, b) propagates NaN (C's fmin does not). Used for `min([1 alpha]...)`. */
static double jl_min(double a, double b) {
    if (isnan(a) || isnan(b)) return NAN;
    return a < b ? a : b;
}

/* ------------------------------------------------------------------ RNG (oracle's own) */
static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void orc_rng_seed(orc_rng *g, uint64_t seed) {
    for (int i = 0; i < 4; i++) g->s[i] = splitmix64(&seed);
    g->have_spare = 0;
    g->spare = 0.0;
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t xoshiro_next(orc_rng *g) {
    uint64_t *s = g->s;
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
}
double orc_rand(orc_rng *g) { return (double)(xoshiro_next(g) >> 11) * 0x1.0p-53; }
double orc_randn(orc_rng *g) {
    if (g->have_spare) { g->have_spare = 0; return g->spare; }
    double u, v, s;
    do {
        u = 2.0 * orc_rand(g) - 1.0;
        v = 2.0 * orc_rand(g) - 1.0;
        s = u * u + v * v;
    } while (s >= 1.0 || s == 0.0);
    s = sqrt(-2.0 * log(s) / s);
    g->spare = v * s;
    g->have_spare = 1;
    return u * s;
}

/* ------------------------------------------------------------------ v_nearest, MCsub.jl:247-263
 *   v = 0.0; mdist = 1e9                                   (:249-250)
 *   for i: distance = (mx[i]-x)^2 + (my[i]-y)^2 + (mz[i]-z)^2   (:254)   [x^2 == x*x; (a+b)+c; no FMA]
 *          if distance < mdist: mdist = distance; v = mv[i]      (:255-258) [strict <: lowest index wins ties]
 * idx_out (not in the reference) = 0-based index of the winning nucleus, -1 if every distance >= 1e9. */
double orc_v_nearest(double x, double y, double z, int K, const double *mx, const double *my,
                     const double *mz, const double *mv, int32_t *idx_out) {
    double v = 0.0;
    double mdist = 1e9;
    int32_t best = -1;
    for (int i = 0; i < K; i++) {
        double dx = mx[i] - x, dy = my[i] - y, dz = mz[i] - z;
        double distance = dx * dx + dy * dy + dz * dz;
        if (distance < mdist) {
            mdist = distance;
            v = mv[i];
            best = i;
        }
    }
    if (idx_out) *idx_out = best;
    return v;
}

/* ------------------------------------------------------------------ Interpolation, MCsub.jl:306-336
 *   npoints = (index of first NaN in X) - 1, else length(X)          (:312-316)
 *   length(Y)==1 / length(Z)==1 -> broadcast to npoints (slices)     (:317-322)
 *   interp_style == 1: zeta[k] = v_nearest(X[k],Y[k],Z[k], cells)    (:326-327)
 * Returns npoints. */
int orc_interpolation(int K, const double *mx, const double *my, const double *mz, const double *mv,
                      int nX, const double *X, int nY, const double *Y, int nZ, const double *Z,
                      double *zeta_out, int32_t *idx_out) {
    int npoints = nX;
    for (int k = 0; k < nX; k++)
        if (isnan(X[k])) { npoints = k; break; }
    for (int k = 0; k < npoints; k++) {
        double yy = (nY == 1) ? Y[0] : Y[k];
        double zz = (nZ == 1) ? Z[0] : Z[k];
        int32_t id;
        zeta_out[k] = orc_v_nearest(X[k], yy, zz, K, mx, my, mz, mv, &id);
        if (idx_out) idx_out[k] = id;
    }
    return npoints;
}

/* ------------------------------------------------------------------ misfit and "likelihood", MCsub.jl:169-182
 * PINNED by the reference's own output: model.jld stores (ptS, tS, phi, likelihood) of 100 models of a 487-ray run with
 * allSig = 0.2; this function reproduces all 100 phi and the likelihood constant BIT FOR BIT (tests/test_oracle.py,
 * tests/golden/model_jld.npz).
 * :169-172  C += (ptS - tS)[k].^2 .* 1.0 / allSig[k][1]^2     -- ((d*d)*1.0)/(sig*sig), k = 1..n sequential.
 * Extension: allSig scaled by the hierarchical noise factor (1.0 reproduces the reference exactly). */
void orc_misfit(int n, const double *ptS, const double *tS, const double *allSig, double noise, double *phi,
                double *likelihood, double *loglik_gauss) {
    double C = 0.0;
    double lg = 0.0, lk = 0.0;
    for (int k = 0; k < n; k++) {
        double sg = noise * allSig[k];
        double df = ptS[k] - tS[k];
        C += ((df * df) * 1.0) / (sg * sg);
        /* :179 likelihood = sum(-log.(allSig * sqrt(2*pi)) * length(tS)); the `- sum(0.5 ...)` of :180 is a
         * separate, discarded statement -> a model-independent constant. */
        lk += (-log(sg * sqrt(2 * ORC_PI))) * (double)n;
        lg += -log(sg * sqrt(2 * ORC_PI));
    }
    if (phi) *phi = C;               /* :173 */
    if (likelihood) *likelihood = lk; /* :182 */
    if (loglik_gauss) *loglik_gauss = lg - 0.5 * C; /* what :179-180 evidently intended; not used by the sampler */
}

/* ------------------------------------------------------------------ evaluate, MCsub.jl:123-185
 * Returns valid (=1 always, :128,:184) or a negative error code for inputs on which the Julia code would
 * throw (segment count != npoints-1 -> DimensionMismatch in the broadcast at :153/:159). */
int orc_evaluate(const orc_params *p, const orc_data *d, orc_model *mdl, int32_t *owners,
                 double *loglik_gauss) {
    const int m = d->m, n = d->R;
    mdl->phi = 1.0;        /* :130-131 */
    mdl->likelihood = 1.0; /* :129,:132 */
    if (p->debug_prior == 1) return 1; /* :134-136 */

    double *zeta0 = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    int32_t *idx0 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    if (owners)
        for (long i = 0; i < (long)m * n; i++) owners[i] = -1;

    for (int i = 0; i < n; i++) { /* :142 */
        const double *X = d->rayX + (size_t)i * m, *Y = d->rayY + (size_t)i * m, *Z = d->rayZ + (size_t)i * m;
        int npoints = orc_interpolation(mdl->K, mdl->x, mdl->y, mdl->z, mdl->zeta, m, X, m, Y, m, Z, zeta0, idx0); /* :143 */
        if (owners)
            for (int k = 0; k < npoints; k++) owners[(size_t)i * m + k] = idx0[k];
        /* :149-161  number of valid segments = first NaN in rayL[:,i] - 1, else the full column */
        const double *rayl = d->rayL + (size_t)i * (m - 1), *rayu = d->rayU + (size_t)i * (m - 1);
        int nseg = m - 1;
        for (int j = 0; j < m - 1; j++)
            if (isnan(rayl[j])) { nseg = j; break; }
        int nzeta = npoints > 0 ? npoints - 1 : 0; /* length(rayzeta), :146-147 */
        if (nseg != nzeta) { free(zeta0); free(idx0); return -2; }
        /* :147 rayzeta = 0.5 .* (zeta0[1:end-1] + zeta0[2:end])
         * :153/:159 ptS[i] = sum(rayl .* rayu .* (rayzeta ./ 1000))  -- (rayl*rayu)*(rayzeta/1000), summed left to right */
        double s = 0.0;
        for (int j = 0; j < nseg; j++) {
            double rayzeta = 0.5 * (zeta0[j] + zeta0[j + 1]);
            s += (rayl[j] * rayu[j]) * (rayzeta / 1000);
        }
        mdl->ptS[i] = s;
    }
    free(zeta0);
    free(idx0);

    orc_misfit(n, mdl->ptS, d->tS, d->allSig, mdl->noise, &mdl->phi, &mdl->likelihood, loglik_gauss); /* :169-182 */
    return 1;
}

/* ------------------------------------------------------------------ build_starting, MCsub.jl:76-121 */
int orc_build_starting(const orc_params *p, const orc_data *d, orc_rng *g, orc_model *mdl) {
    /* :86-87 nCells = floor(exp(rand*log(max_cells/min_cells) + log(min_cells))) */
    double nC = floor(exp(orc_rand(g) * log((double)p->max_cells / (double)p->min_cells) + log((double)p->min_cells)));
    int K = (int)nC;
    if (K > mdl->cap) return -3;
    mdl->K = K;
    for (int i = 0; i < K; i++) mdl->x[i] = p->xmin + (p->xmax - p->xmin) * orc_rand(g); /* :92 */
    for (int i = 0; i < K; i++) mdl->y[i] = p->ymin + (p->ymax - p->ymin) * orc_rand(g); /* :93 */
    for (int i = 0; i < K; i++) mdl->z[i] = p->zmin + (p->zmax - p->zmin) * orc_rand(g); /* :94 */
    if (p->prior == 1) {
        for (int i = 0; i < K; i++) mdl->zeta[i] = orc_rand(g) * p->zeta_scale; /* :100 */
    } else if (p->prior == 2) {
        for (int i = 0; i < K; i++) mdl->zeta[i] = 0.0 + p->zeta_scale * orc_randn(g); /* :105 */
    } else {
        for (int i = 0; i < K; i++) mdl->zeta[i] = -log(orc_rand(g)) * p->zeta_scale; /* :108 */
    }
    mdl->phi = -1; mdl->likelihood = -1; mdl->action = -1; mdl->accept = -1; /* :111-115 */
    int v = orc_evaluate(p, d, mdl, NULL, NULL); /* :118 */
    if (v < 0) return v;
    return 1; /* :119 */
}

/* ------------------------------------------------------------------ proposal loop, TD_inversion_function.jl:70-274 */
static void model_copy(orc_model *dst, const orc_model *src, int R) {
    dst->K = src->K;
    memcpy(dst->x, src->x, sizeof(double) * (size_t)src->K);
    memcpy(dst->y, src->y, sizeof(double) * (size_t)src->K);
    memcpy(dst->z, src->z, sizeof(double) * (size_t)src->K);
    memcpy(dst->zeta, src->zeta, sizeof(double) * (size_t)src->K);
    dst->phi = src->phi;
    dst->likelihood = src->likelihood;
    memcpy(dst->ptS, src->ptS, sizeof(double) * (size_t)R);
    dst->action = src->action;
    dst->accept = src->accept;
    dst->noise = src->noise;
}

static int model_alloc(orc_model *mm, int cap, int R) {
    mm->cap = cap;
    mm->x = (double *)calloc((size_t)cap, sizeof(double));
    mm->y = (double *)calloc((size_t)cap, sizeof(double));
    mm->z = (double *)calloc((size_t)cap, sizeof(double));
    mm->zeta = (double *)calloc((size_t)cap, sizeof(double));
    mm->ptS = (double *)calloc((size_t)(R > 0 ? R : 1), sizeof(double));
    return (mm->x && mm->y && mm->z && mm->zeta && mm->ptS) ? 0 : -1;
}
static void model_free(orc_model *mm) { free(mm->x); free(mm->y); free(mm->z); free(mm->zeta); free(mm->ptS); }

int orc_chain_run(const orc_params *p, const orc_data *d, orc_model *model, int64_t iter0, int64_t nIter,
                  int mode, orc_proposal *recs, orc_rng *g,
                  int8_t *tr_accept, double *tr_phi, int32_t *tr_K,
                  int32_t hist_cap, int32_t *n_hist, int64_t *model_num,
                  int32_t *hist_K, double *hist_cells, double *hist_phi, double *hist_ptS,
                  int64_t *hist_iter, int32_t *hist_action, int32_t *hist_accept) {
    const int R = d->R;
    const int cap = model->cap;
    /* :22-23 */
    const double sig_zeta = p->zeta_scale * p->sig / 100;
    const double sig_sig = p->max_sig * p->sig / 100;
    /* :30-32 */
    const double xr = (p->sig / 100) * (p->xmax - p->xmin);
    const double yr = (p->sig / 100) * (p->ymax - p->ymin);
    const double zr = (p->sig / 100) * (p->zmax - p->zmin);
    const int nact = p->reserved >= 4 ? p->reserved : 4; /* :72 rand(1:4); 5 enables the sigma move (extension) */
    const double ndata = (double)R;                      /* :29 n = length(dataStruct.dataX) */

    orc_model mn;
    if (model_alloc(&mn, cap, R)) return -1;
    int rc = 0;

    for (int64_t it = 0; it < nIter; it++) {
        const int64_t iter = iter0 + it;
        orc_proposal rec;
        memset(&rec, 0, sizeof rec);
        if (mode == 0) rec = recs[it];
        else rec.action = 1 + (int)floor(orc_rand(g) * nact); /* :72 */
        const int action = rec.action;
        model->action = action; /* :73 */
        model->accept = 0;      /* :74 */
        double alpha = 0.0;     /* loop-carried in the reference (:69) but always reassigned before it matters */
        int valid = 0;
        const double K0 = (double)model->K;

        if (action == 1) { /* ---- birth :76-125 */
            if (model->K < p->max_cells) { /* :77 */
                if (mode == 1) {
                    rec.x = orc_rand(g) * (p->xmax - p->xmin) + p->xmin; /* :78 */
                    rec.y = orc_rand(g) * (p->ymax - p->ymin) + p->ymin; /* :79 */
                    rec.z = orc_rand(g) * (p->zmax - p->zmin) + p->zmin; /* :80 */
                }
                double czeta = orc_v_nearest(rec.x, rec.y, rec.z, model->K, model->x, model->y, model->z, model->zeta, NULL); /* :81 */
                if (mode == 1) rec.zeta = czeta + sig_zeta * orc_randn(g); /* :82 */
                const double zetanew = rec.zeta;
                if (model->K + 1 > cap) { rc = -3; break; }
                model_copy(&mn, model, R); /* :84 */
                mn.x[mn.K] = rec.x; mn.y[mn.K] = rec.y; mn.z[mn.K] = rec.z; mn.zeta[mn.K] = zetanew; /* :85-88 */
                mn.K += 1; /* :89 */
                if (p->prior == 1) { /* :91-104 */
                    if (zetanew > 0 && zetanew < p->zeta_scale) {
                        valid = orc_evaluate(p, d, &mn, NULL, NULL); /* :93 */
                        alpha = ((K0) / (K0 + 1)) * ((sig_zeta * sqrt(2 * ORC_PI)) / (p->zeta_scale)) *
                                exp(((czeta - zetanew) * (czeta - zetanew)) / (2 * (sig_zeta * sig_zeta)) -
                                    (mn.phi - model->phi) / 2); /* :96-97 */
                        alpha = jl_min(1, alpha); /* :100 */
                    } else {
                        valid = 0; /* :103 */
                    }
                } else if (p->prior == 2) { /* :105-109 */
                    valid = orc_evaluate(p, d, &mn, NULL, NULL);
                    alpha = ((K0) / (K0 + 1)) * (sig_zeta / p->zeta_scale) *
                            exp(-(zetanew * zetanew) / (p->zeta_scale * p->zeta_scale) +
                                ((czeta - zetanew) * (czeta - zetanew)) / (2 * (sig_zeta * sig_zeta)) - (mn.phi - model->phi) / 2);
                    alpha = jl_min(1, alpha);
                } else { /* :110-119 */
                    if (zetanew > 0) {
                        valid = orc_evaluate(p, d, &mn, NULL, NULL);
                        alpha = ((K0) / (K0 + 1)) * (sqrt(2 * ORC_PI) * sig_zeta / p->zeta_scale) *
                                exp(-zetanew / p->zeta_scale + ((czeta - zetanew) * (czeta - zetanew)) / (2 * (sig_zeta * sig_zeta)) -
                                    (mn.phi - model->phi) / 2);
                        alpha = jl_min(1, alpha);
                    } else {
                        valid = 0;
                    }
                }
                if (valid < 0) { rc = valid; break; }
                if (mode == 1) rec.u = orc_rand(g); /* :121 -- drawn even when valid == 0 */
                if (rec.u < alpha && valid == 1) { model_copy(model, &mn, R); model->accept = 1; } /* :121-124 */
            }
        } else if (action == 2) { /* ---- death :126-181 */
            if (model->K > p->min_cells) { /* :127 */
                if (mode == 1) rec.idx = (int)floor(orc_rand(g) * model->K); /* :128 rand(1:nCells), 0-based here */
                const int kill = rec.idx;
                if (kill < 0 || kill >= model->K) { rc = -4; break; }
                model_copy(&mn, model, R); /* :130 */
                for (int i = kill; i < mn.K - 1; i++) { /* :132-135 deleteat! (order preserving) */
                    mn.x[i] = mn.x[i + 1]; mn.y[i] = mn.y[i + 1]; mn.z[i] = mn.z[i + 1]; mn.zeta[i] = mn.zeta[i + 1];
                }
                mn.K -= 1; /* :136 */
                valid = orc_evaluate(p, d, &mn, NULL, NULL); /* :141 */
                if (valid < 0) { rc = valid; break; }
                const double zetanew = orc_v_nearest(model->x[kill], model->y[kill], model->z[kill], mn.K, mn.x, mn.y, mn.z, mn.zeta, NULL); /* :146 */
                const double zk = model->zeta[kill];
                if (p->prior == 1) { /* :150-156 */
                    alpha = ((K0) / (K0 - 1)) * ((p->zeta_scale) / (sig_zeta * sqrt(2 * ORC_PI))) *
                            exp(-((zk - zetanew) * (zk - zetanew)) / (2 * (sig_zeta * sig_zeta)) - (mn.phi - model->phi) / 2);
                    alpha = jl_min(1, alpha);
                } else if (p->prior == 2) { /* :157-163 (the second evaluate at :159 is idempotent) */
                    alpha = ((K0) / (K0 - 1)) * (p->zeta_scale / sig_zeta) *
                            exp((zk * zk) / (2 * (p->zeta_scale * p->zeta_scale)) - ((zk - zetanew) * (zk - zetanew)) / (2 * (sig_zeta * sig_zeta)) -
                                (mn.phi - model->phi) / 2);
                    alpha = jl_min(1, alpha);
                } else { /* :164-173 */
                    if (zetanew > 0) {
                        alpha = ((K0) / (K0 - 1)) * (p->zeta_scale / (sqrt(2 * ORC_PI) * sig_zeta)) *
                                exp(zk / p->zeta_scale - ((zk - zetanew) * (zk - zetanew)) / (2 * (sig_zeta * sig_zeta)) -
                                    (mn.phi - model->phi) / 2);
                        alpha = jl_min(1, alpha);
                    } else {
                        valid = 0;
                    }
                }
                if (mode == 1) rec.u = orc_rand(g); /* :176 */
                if (rec.u < alpha && valid == 1) { model_copy(model, &mn, R); model->accept = 1; } /* :176-180 */
            }
        } else if (action == 3) { /* ---- change zeta :183-218 */
            if (mode == 1) rec.idx = (int)floor(orc_rand(g) * model->K); /* :184 */
            const int ch = rec.idx;
            if (ch < 0 || ch >= model->K) { rc = -4; break; }
            model_copy(&mn, model, R); /* :186 */
            if (mode == 1) rec.zeta = model->zeta[ch] + sig_zeta * orc_randn(g); /* :188 */
            mn.zeta[ch] = rec.zeta; /* :189 */
            valid = orc_evaluate(p, d, &mn, NULL, NULL); /* :191 (before the bounds check) */
            if (valid < 0) { rc = valid; break; }
            if (p->prior == 1) { /* :194-200 */
                if (mn.zeta[ch] > 0 && mn.zeta[ch] < p->zeta_scale) {
                    alpha = exp(-(mn.phi - model->phi) / 2);
                    alpha = jl_min(1, alpha);
                } else {
                    valid = 0;
                }
            } else if (p->prior == 2) { /* :201-204 */
                alpha = exp((model->zeta[ch] * model->zeta[ch] - mn.zeta[ch] * mn.zeta[ch]) / (2 * (p->zeta_scale * p->zeta_scale)) -
                            (mn.phi - model->phi) / 2);
                alpha = jl_min(1, alpha);
            } else { /* :205-212 */
                if (mn.zeta[ch] > 0) {
                    alpha = exp((model->zeta[ch] - mn.zeta[ch]) / p->zeta_scale - (mn.phi - model->phi) / 2);
                    alpha = jl_min(1, alpha);
                } else {
                    alpha = 0;
                }
            }
            if (mode == 1) rec.u = orc_rand(g); /* :214 */
            if (rec.u < alpha && valid == 1) { model_copy(model, &mn, R); model->accept = 1; } /* :214-218 */
        } else if (action == 4) { /* ---- move :220-251 */
            if (model->K > 0) { /* :221 */
                if (mode == 1) rec.idx = (int)floor(orc_rand(g) * model->K); /* :222 */
                const int mv = rec.idx;
                if (mv < 0 || mv >= model->K) { rc = -4; break; }
                model_copy(&mn, model, R); /* :224 */
                if (mode == 1) {
                    rec.x = model->x[mv] + xr * orc_randn(g); /* :226 */
                    rec.y = model->y[mv] + yr * orc_randn(g); /* :227 */
                    rec.z = model->z[mv] + zr * orc_randn(g); /* :228 */
                }
                if (rec.x >= p->xmin && rec.x <= p->xmax && rec.y >= p->ymin && rec.y <= p->ymax &&
                    rec.z >= p->zmin && rec.z <= p->zmax) { /* :230-232 */
                    mn.x[mv] = rec.x; mn.y[mv] = rec.y; mn.z[mv] = rec.z; /* :234-236 */
                    valid = orc_evaluate(p, d, &mn, NULL, NULL); /* :238 */
                    if (valid < 0) { rc = valid; break; }
                    alpha = exp(-(mn.phi - model->phi) / 2); /* :241 */
                    alpha = jl_min(1, alpha);
                } else {
                    valid = 0; /* :244 */
                }
                if (mode == 1) rec.u = orc_rand(g); /* :247 */
                if (rec.u < alpha && valid == 1) { model_copy(model, &mn, R); model->accept = 1; } /* :247-250 */
            }
        } else if (action == 5) { /* ---- change sigma :252-272 -- DEAD CODE in the reference.
             * Extension spec: the per-datum sigmas are scaled by a common noise factor lambda (model->noise);
             * sig_n ~ N(lambda, sig_sig) :254; require 0 < sig_n < max_sig :257; re-evaluate :261;
             * log(alpha) = n*log(lambda/sig_n) - (phi' - phi)/2, clipped at 0 :264-265; accept iff log(u) <= alpha :267.
             * The accept uniform is drawn only inside the bounds check (:267 sits inside the `if` of :257). */
            if (mode == 1) rec.zeta = model->noise + sig_sig * orc_randn(g); /* :254 */
            const double sig_n = rec.zeta;
            if (sig_n > 0 && sig_n < p->max_sig) { /* :257 */
                model_copy(&mn, model, R);
                mn.noise = sig_n; /* :259-260 */
                valid = orc_evaluate(p, d, &mn, NULL, NULL); /* :261 */
                if (valid < 0) { rc = valid; break; }
                alpha = log(model->noise / sig_n) * ndata - (mn.phi - model->phi) / 2; /* :264 */
                alpha = jl_min(log(1.0), alpha); /* :265 */
                if (mode == 1) rec.u = orc_rand(g);
                if (log(rec.u) <= alpha && valid == 1) { model_copy(model, &mn, R); model->accept = 1; } /* :267-271 */
            }
        } else {
            rc = -5;
            break;
        }

        if (mode == 1 && recs) recs[it] = rec;
        if (tr_accept) tr_accept[it] = (int8_t)model->accept;
        if (tr_phi) tr_phi[it] = model->phi;
        if (tr_K) tr_K[it] = model->K;

        /* thinning :275-281 (checkpoint writes :282-294 are out of scope) */
        if ((double)iter >= p->burn_in) {
            *model_num += 1;
            if (fmod((double)*model_num, p->keep_each) == 0) {
                if (*n_hist < hist_cap) {
                    const int h = *n_hist;
                    if (hist_K) hist_K[h] = model->K;
                    if (hist_cells) {
                        double *c = hist_cells + (size_t)h * 4 * cap;
                        memcpy(c, model->x, sizeof(double) * (size_t)model->K);
                        memcpy(c + cap, model->y, sizeof(double) * (size_t)model->K);
                        memcpy(c + 2 * cap, model->z, sizeof(double) * (size_t)model->K);
                        memcpy(c + 3 * cap, model->zeta, sizeof(double) * (size_t)model->K);
                    }
                    if (hist_phi) hist_phi[h] = model->phi;
                    if (hist_ptS) memcpy(hist_ptS + (size_t)h * R, model->ptS, sizeof(double) * (size_t)R);
                    if (hist_iter) hist_iter[h] = iter;
                    if (hist_action) hist_action[h] = model->action;
                    if (hist_accept) hist_accept[h] = model->accept;
                }
                *n_hist += 1; /* counts every kept model even beyond capacity */
            }
        }
    }
    model_free(&mn);
    return rc;
}

/* ------------------------------------------------------------------ ray preprocessing, load_data_Tonga.jl:66-69
 *   rayl = sqrt.((x[1:end-1,:]-x[2:end,:]).^2 + (y..).^2 + (z..).^2)     (:66-68)
 *   rayu = 0.5 .* (U[1:end-1,:] + U[2:end,:])                            (:69)
 * NaN wherever an endpoint is padding (NaN arithmetic). */
void orc_ray_lengths(int m, int R, const double *x, const double *y, const double *z, const double *U,
                     double *rayL, double *rayU) {
    for (int i = 0; i < R; i++)
        for (int j = 0; j < m - 1; j++) {
            size_t a = (size_t)i * m + j, o = (size_t)i * (m - 1) + j;
            double dx = x[a] - x[a + 1], dy = y[a] - y[a + 1], dz = z[a] - z[a + 1];
            rayL[o] = sqrt(dx * dx + dy * dy + dz * dz);
            rayU[o] = 0.5 * (U[a] + U[a + 1]);
        }
}
