/*
 * tonga_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the hot path of Geronimorz/MCMC-in-Tonga: v_nearest / Interpolation /
 * evaluate / build_starting (MCsub.jl) and the proposal loop of TD_inversion_function
 * (TD_inversion_function.jl).  Every function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libtonga_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference is pure Julia and cannot run in this image (no julia), it has no
 * tests / golden vectors / KATs, and its only shipped output (model.jld) belongs to an unshipped
 * 487-ray data set.  This restatement is therefore pinned only by
 * (a) an independent NumPy twin (oracle/oracle_np.py) written from the same cited lines and
 * (b) the invariants model.jld does hold (tests/test_oracle.py).
 *
 * Arithmetic: IEEE binary64 throughout, compiled with -ffp-contract=off and without -ffast-math, so
 * `a*a + b*b + c*c` is evaluated exactly as Julia evaluates it (no FMA contraction, left to right).
 * Known, documented deviation: Julia's `sum` over a broadcast vector is a SIMD/pairwise reduction
 * whose order is unspecified; the oracle sums left to right.  t* and phi are therefore defined to
 * ~1e-15 relative by the reference itself; the parity tolerance is 1e-9 relative (BASELINE.json).
 */
#ifndef TONGA_ORACLE_H
#define TONGA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Hot-path subset of `struct parameters` (define_TDstructure.jl:1-44) + the box that
 * TD_inversion_function.jl:27-32 / MCsub.jl:92-94 derive from min/max of xVec,yVec,zVec. */
typedef struct {
    double xmin, xmax, ymin, ymax, zmin, zmax; /* min(xVec...), max(xVec...), ...            */
    double sig;                                /* percent (define_TDstructure.jl:8)            */
    double zeta_scale;                         /* define_TDstructure.jl:9                      */
    double max_sig;                            /* define_TDstructure.jl:12                     */
    double n_iter, burn_in, keep_each;         /* Float64 in the reference (:24-27)            */
    int32_t min_cells, max_cells;              /* :10-11                                       */
    int32_t prior;                             /* 1 uniform, 2 normal, 3 exponential (:15)     */
    int32_t debug_prior;                       /* :3                                           */
    int32_t interp_style;                      /* 1 nearest (only working style)    */
    int32_t reserved;
} orc_params;

/* `DataStruct` fields the hot path reads (DefStruct.jl:5-30), in the reference's own layout:
 * column-major m x R (a column is one ray), NaN tail padding. */
typedef struct {
    int32_t m, R;
    const double *rayX, *rayY, *rayZ; /* m x R     */
    const double *rayL, *rayU;        /* (m-1) x R */
    const double *tS, *allSig;        /* R         */
} orc_data;

/* One RJ-MCMC proposal as the reference draws it (TD_inversion_function.jl:72-251), recorded so the
 * same stream can drive the oracle and the device (reference seeds are wall-clock).
 *   action 1 birth : (x,y,z) = new nucleus (:78-80), zeta = zetanew (:82), u = accept uniform (:121)
 *   action 2 death : idx = kill, 0-based (:128), u (:176)
 *   action 3 change: idx (:184), zeta = zetan (:188), u (:214)
 *   action 4 move  : idx (:222), (x,y,z) = proposed position (:226-228), u (:247)
 *   action 5 sigma : zeta = proposed noise scale, u   (extension; dead code in the reference :252-272)
 */
typedef struct {
    int32_t action;
    int32_t idx;
    double x, y, z;
    double zeta;
    double u;
} orc_proposal;

/* `mutable struct Model` (DefStruct.jl:32-48) with fixed-capacity arrays. */
typedef struct {
    int32_t K;   /* nCells (Float64 in the reference) */
    int32_t cap; /* capacity of the arrays below      */
    double *x, *y, *z, *zeta;
    double phi;
    double likelihood;
    double *ptS; /* R */
    int32_t action, accept;
    double noise; /* hierarchical noise scale lambda (extension); 1.0 = reference behaviour */
} orc_model;

typedef struct { uint64_t s[4]; int have_spare; double spare; } orc_rng;

void   orc_rng_seed(orc_rng *g, uint64_t seed);
double orc_rand(orc_rng *g);   /* uniform [0,1)  */
double orc_randn(orc_rng *g);  /* standard normal */

double orc_v_nearest(double x, double y, double z, int K, const double *mx, const double *my,
                     const double *mz, const double *mv, int32_t *idx_out);

int orc_interpolation(int K, const double *mx, const double *my, const double *mz, const double *mv,
                      int nX, const double *X, int nY, const double *Y, int nZ, const double *Z,
                      double *zeta_out, int32_t *idx_out);

/* misfit phi + the constant "likelihood" (+ the Gaussian log-likelihood evidently intended), MCsub.jl:169-182; outputs may be NULL */
void orc_misfit(int n, const double *ptS, const double *tS, const double *allSig, double noise, double *phi,
                double *likelihood, double *loglik_gauss);

int orc_evaluate(const orc_params *p, const orc_data *d, orc_model *mdl, int32_t *owners /* m*R or NULL */,
                 double *loglik_gauss /* or NULL */);

int orc_build_starting(const orc_params *p, const orc_data *d, orc_rng *g, orc_model *mdl);

/* Run iterations iter0 .. iter0+nIter-1 (1-based `iter`, as in TD_inversion_function.jl:70).
 *   mode 0: replay `recs[0..nIter)`;  mode 1: draw proposals from `g` and write them to `recs` (if non-NULL).
 * Per-iteration traces (each nIter long, any may be NULL): accepted flag, phi after the iteration, K after.
 * Thinning (TD_inversion_function.jl:275-281): kept models are appended to the hist_* arrays (capacity
 * hist_cap records; hist_cells is [hist_cap][4][mdl->cap], hist_ptS is [hist_cap][R]); *n_hist and
 * *model_num carry over between calls.  Returns 0, or <0 on error. */
int orc_chain_run(const orc_params *p, const orc_data *d, orc_model *mdl, int64_t iter0, int64_t nIter,
                  int mode, orc_proposal *recs, orc_rng *g,
                  int8_t *tr_accept, double *tr_phi, int32_t *tr_K,
                  int32_t hist_cap, int32_t *n_hist, int64_t *model_num,
                  int32_t *hist_K, double *hist_cells, double *hist_phi, double *hist_ptS,
                  int64_t *hist_iter, int32_t *hist_action, int32_t *hist_accept);

/* tonga_oracle_mt.c: the chain farm (main_inversion.jl:15), nChains independent chains on nThreads host threads; the second
 * form starts from the caller's models (cells0[nChains][4][Kcap0], x / y / z / zeta rows) instead of build_starting */
int orc_chain_farm(const orc_params *p, const orc_data *d, int nChains, int64_t nIter, int nThreads,
                   uint64_t seed, double *phi_out, int32_t *K_out, int64_t *acc_out);
int orc_chain_farm_from(const orc_params *p, const orc_data *d, int nChains, int64_t nIter, int nThreads, uint64_t seed,
                        const int32_t *K0, const double *cells0, int Kcap0, double *phi_out, int32_t *K_out, int64_t *acc_out);

/* ray preprocessing, load_data_Tonga.jl:66-69 */
void orc_ray_lengths(int m, int R, const double *x, const double *y, const double *z, const double *U,
                     double *rayL, double *rayU);

#ifdef __cplusplus
}
#endif
#endif
