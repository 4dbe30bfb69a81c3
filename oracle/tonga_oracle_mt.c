/*
 * tonga_oracle_mt.c -- CPU ORACLE (test infrastructure, NOT the product): the reference's chain farm.
 *
 * main_inversion.jl:15 runs `pmap(x -> TD_inversion_function(TD_parameters, dataStruct, x), 1:n_chains)`:
 * independent chains, one per worker process.  Here: one pthread per worker, chains handed out round-robin.
 * Used only by bench.py's cpu_baseline / `--impl reference` legs to time the reference algorithm on the
 * host cores of the GPU box (julia is not installed).
 */
#include "tonga_oracle.h"

#include <pthread.h>
#include <stdlib.h>

typedef struct {
    const orc_params *p;
    const orc_data *d;
    int first, stride, nChains;
    int64_t nIter;
    uint64_t seed;
    int cap;
    double *phi_out;   /* nChains */
    int32_t *K_out;    /* nChains */
    int64_t *acc_out;  /* nChains */
    int rc;
    /* optional start models (orc_chain_farm_from): K0[nChains], cells0[nChains][4][Kcap0]; NULL -> build_starting */
    const int32_t *K0;
    const double *cells0;
    int Kcap0;
} worker_arg;

static void *worker(void *vp) {
    worker_arg *a = (worker_arg *)vp;
    const int R = a->d->R;
    orc_model m;
    m.cap = a->cap;
    m.x = (double *)calloc((size_t)a->cap, sizeof(double));
    m.y = (double *)calloc((size_t)a->cap, sizeof(double));
    m.z = (double *)calloc((size_t)a->cap, sizeof(double));
    m.zeta = (double *)calloc((size_t)a->cap, sizeof(double));
    m.ptS = (double *)calloc((size_t)R, sizeof(double));
    int8_t *acc = (int8_t *)malloc((size_t)a->nIter);
    a->rc = 0;
    for (int c = a->first; c < a->nChains; c += a->stride) {
        orc_rng g;
        orc_rng_seed(&g, a->seed + (uint64_t)c);
        m.noise = 1.0;
        int rc;
        if (a->K0) { /* the caller's start model + the evaluate that build_starting ends with (MCsub.jl:118) */
            const int k = a->K0[c];
            if (k < 1 || k > a->cap || k > a->Kcap0) { a->rc = -1; break; }
            const double *c0 = a->cells0 + (size_t)c * 4 * (size_t)a->Kcap0;
            for (int i = 0; i < k; i++) {
                m.x[i] = c0[i]; m.y[i] = c0[a->Kcap0 + i]; m.z[i] = c0[2 * (size_t)a->Kcap0 + i]; m.zeta[i] = c0[3 * (size_t)a->Kcap0 + i];
            }
            m.K = k;
            rc = orc_evaluate(a->p, a->d, &m, NULL, NULL);
        } else {
            rc = orc_build_starting(a->p, a->d, &g, &m); /* TD_inversion_function.jl:43-45 */
        }
        if (rc < 0) { a->rc = rc; break; }
        int32_t nh = 0;
        int64_t mnum = 0;
        rc = orc_chain_run(a->p, a->d, &m, 1, a->nIter, 1, NULL, &g, acc, NULL, NULL, 0, &nh, &mnum,
                           NULL, NULL, NULL, NULL, NULL, NULL, NULL);
        if (rc < 0) { a->rc = rc; break; }
        int64_t s = 0;
        for (int64_t i = 0; i < a->nIter; i++) s += acc[i];
        if (a->phi_out) a->phi_out[c] = m.phi;
        if (a->K_out) a->K_out[c] = m.K;
        if (a->acc_out) a->acc_out[c] = s;
    }
    free(acc);
    free(m.x); free(m.y); free(m.z); free(m.zeta); free(m.ptS);
    return NULL;
}

static int farm(const orc_params *p, const orc_data *d, int nChains, int64_t nIter, int nThreads, uint64_t seed, double *phi_out,
                int32_t *K_out, int64_t *acc_out, const int32_t *K0, const double *cells0, int Kcap0);

/* Run nChains independent chains of nIter iterations each on nThreads host threads. */
int orc_chain_farm(const orc_params *p, const orc_data *d, int nChains, int64_t nIter, int nThreads,
                   uint64_t seed, double *phi_out, int32_t *K_out, int64_t *acc_out) {
    return farm(p, d, nChains, nIter, nThreads, seed, phi_out, K_out, acc_out, NULL, NULL, 0);
}

/* The same farm started from the caller's models (K0[nChains], cells0[nChains][4][Kcap0], x / y / z / zeta rows) instead of
 * build_starting: bench.py times both arms on the same chain states. */
int orc_chain_farm_from(const orc_params *p, const orc_data *d, int nChains, int64_t nIter, int nThreads, uint64_t seed,
                        const int32_t *K0, const double *cells0, int Kcap0, double *phi_out, int32_t *K_out, int64_t *acc_out) {
    if (!K0 || !cells0 || Kcap0 < 1) return -1;
    return farm(p, d, nChains, nIter, nThreads, seed, phi_out, K_out, acc_out, K0, cells0, Kcap0);
}

static int farm(const orc_params *p, const orc_data *d, int nChains, int64_t nIter, int nThreads, uint64_t seed, double *phi_out,
                int32_t *K_out, int64_t *acc_out, const int32_t *K0, const double *cells0, int Kcap0) {
    if (nThreads < 1) nThreads = 1;
    if (nThreads > nChains) nThreads = nChains;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nThreads);
    worker_arg *args = (worker_arg *)malloc(sizeof(worker_arg) * (size_t)nThreads);
    for (int t = 0; t < nThreads; t++) {
        args[t] = (worker_arg){p, d, t, nThreads, nChains, nIter, seed, p->max_cells + 1, phi_out, K_out, acc_out, 0, K0, cells0, Kcap0};
        pthread_create(&th[t], NULL, worker, &args[t]);
    }
    int rc = 0;
    for (int t = 0; t < nThreads; t++) {
        pthread_join(th[t], NULL);
        if (args[t].rc < 0) rc = args[t].rc;
    }
    free(th);
    free(args);
    return rc;
}
