/*
 * tonga_b200.h -- C ABI of libtonga_b200.so: the B200 (sm_100a) implementation of the per-iteration forward
 * model + likelihood + proposal loop of Geronimorz/MCMC-in-Tonga's reversible-jump t* tomography.
 *
 * The reference has NO FFI today: its boundary is a set of Julia top-level functions made global by
 * `@everywhere include(...)` (main_inversion.jl:4-9).  A drop-in replacement is a .jl file included after
 * MCsub.jl / TD_inversion_function.jl that re-defines those methods and forwards to `ccall` on the entry
 * points below (mcmc-in-tonga_b200/julia/TongaB200.jl; binding shown in INTEGRATION.md).  Each entry point cites
 * the reference interface it replaces (file:line relative to the reference repo).
 *
 * Conventions
 *   - plain C types only; every host pointer is caller-owned and is NOT retained after the call returns
 *     (the library copies at create time); outputs go to caller-allocated buffers.
 *   - return value: 0 = TONGA_OK, negative = error class; message via tonga_last_error() (thread local).
 *     No C++ exception crosses the ABI.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with TONGA_ERR_CUDA.
 *   - nucleus / ray / point indices are 0-based at the ABI (the Julia shim converts).
 *   - "cells" arrays are SoA per model: [model][4][Kcap] doubles = x[Kcap], y[Kcap], z[Kcap], zeta[Kcap].
 *   - flat point order: ray 0's valid points, then ray 1's, ... (CSR; NaN padding of the reference layout dropped).
 */
#ifndef TONGA_B200_H
#define TONGA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TONGA_OK 0
#define TONGA_ERR_ARG (-1)      /* bad argument                                                       */
#define TONGA_ERR_CUDA (-2)     /* CUDA runtime error / no device                                     */
#define TONGA_ERR_CAPACITY (-3) /* model / shared-memory / history capacity exceeded                  */
#define TONGA_ERR_STATE (-4)    /* call order (e.g. run before models were set)                       */
#define TONGA_ERR_DATA (-5)     /* ray data on which the Julia code would throw (DimensionMismatch)   */
#define TONGA_ERR_PEER (-6)     /* ray-sharded run: a peer rank did not publish its exchange in time  */

typedef struct tonga_ctx tonga_ctx;       /* device-resident ray geometry + observations (one per GPU)        */
typedef struct tonga_chains tonga_chains; /* a batch of independent RJ-MCMC chains resident on that GPU       */

/* Hot-path subset of `struct parameters` (define_TDstructure.jl:1-44) plus the nucleus box, i.e. min/max of
 * dataStruct.xVec / yVec / zVec as TD_inversion_function.jl:30-32,78-80,230-232 and MCsub.jl:92-94 use them. */
typedef struct {
    double xmin, xmax, ymin, ymax, zmin, zmax;
    double sig;                        /* percent, define_TDstructure.jl:8                                     */
    double zeta_scale;                 /* :9                                                                   */
    double max_sig;                    /* :12 (only the sigma move, an extension, reads it)                    */
    double n_iter, burn_in, keep_each; /* :24-26 (Float64 in the reference)                                    */
    int32_t min_cells, max_cells;      /* :10-11                                                               */
    int32_t prior;                     /* :15  1 uniform (default), 2 normal, 3 exponential                    */
    int32_t debug_prior;               /* :3   1 = sample the prior: evaluate() returns phi = 1 (MCsub.jl:134) */
    int32_t interp_style;              /* :13  must be 1 (nearest); style 2 is broken in the reference (F7)    */
    int32_t n_actions;                 /* 4 = reference (`rand(1:4)`, TD_inversion_function.jl:72); 5 adds the sigma move */
} tonga_params;

/* One proposal of TD_inversion_function.jl:72-251 in recorded form (replay parity; reference seeds are wall-clock,
 * TD_inversion_function.jl:13).  action: 1 birth (x,y,z,zeta=zetanew,u) :76-125 | 2 death (idx=kill,u) :126-181 |
 * 3 change (idx,zeta=zetan,u) :183-218 | 4 move (idx,x,y,z,u) :220-251 | 5 sigma (zeta=sig_n,u) :252-272 (extension). */
typedef struct {
    int32_t action;
    int32_t idx;
    double x, y, z;
    double zeta;
    double u;
} tonga_proposal;

const char *tonga_last_error(void);
int tonga_version(void);
int tonga_device_count(int *n_out);

/* ---- geometry context: replaces the DataStruct fields evaluate() reads (DefStruct.jl:5-30; MCsub.jl:138-171).
 * rayX/rayY/rayZ: m x R column-major (a column is one ray), NaN tail padding; rayL/rayU: (m-1) x R
 * (load_data_Tonga.jl:66-69); tS, allSig: R.  Flattened once into device SoA arrays (points, dt = rayL*rayU). */
int tonga_create(tonga_ctx **out, int32_t m, int32_t R, const double *rayX, const double *rayY, const double *rayZ,
                 const double *rayL, const double *rayU, const double *tS, const double *allSig,
                 const tonga_params *params, int32_t device);
/* The same context built from the raw point matrices: rayL = sqrt.(dx.^2 + dy.^2 + dz.^2) and rayU = 0.5 .* (U[k] + U[k+1])
 * (load_data_Tonga.jl:66-69) are computed on the device together with the flatten, for ray sets where the host-side
 * preprocessing and flatten of tonga_create take seconds (BASELINE config 3: 2e7 points).  U[m x R] = slowness at the ray
 * points (column-major like rayX; NaN padding).  Bit-identical to tonga_create fed with those rayL / rayU. */
int tonga_create_from_points(tonga_ctx **out, int32_t m, int32_t R, const double *rayX, const double *rayY, const double *rayZ,
                             const double *U, const double *tS, const double *allSig, const tonga_params *params, int32_t device);
/* Frees the context.  Every tonga_chains batch created on it must have been destroyed first (a batch refers to its context;
 * the Python mirror's Context.close() and the Julia shim's `finally` blocks take care of the order). */
void tonga_destroy(tonga_ctx *ctx);
/* R, number of valid points P, number of segments S, padded point count used on the device */
int tonga_info(const tonga_ctx *ctx, int32_t *R, int64_t *P, int64_t *S, int64_t *Ppad);
/* ray_off[R+1]: CSR offsets of the flat point order (for interpreting `owners`) */
int tonga_ray_offsets(const tonga_ctx *ctx, int32_t *ray_off);
int tonga_synchronize(tonga_ctx *ctx);
/* Distance arithmetic of the full evaluate.  0 (default): FP32 screening (best / second-best nucleus per point) with an exact
 * FP64 re-scan of every point whose two best candidates fall inside a rigorous rounding-error band -- owners stay bit-exact;
 * 1: every distance in exact FP64 (same results, slower). */
int tonga_set_exact_only(tonga_ctx *ctx, int32_t exact_only);

/* ---- evaluate(model, dataStruct, TD_parameters), MCsub.jl:123-185.
 * ptS_out[R] (model.ptS), *phi_out (model.phi), *like_out (model.likelihood: the model-independent constant of
 * MCsub.jl:179), *loglik_out (the Gaussian log-likelihood :179-180 evidently intended); any may be NULL.
 * noise = hierarchical noise scale multiplying allSig (1.0 = reference). */
int tonga_evaluate(tonga_ctx *ctx, int32_t K, const double *x, const double *y, const double *z, const double *zeta,
                   double noise, double *ptS_out, double *phi_out, double *like_out, double *loglik_out);

/* The misfit alone, MCsub.jl:169-173: phi[i] = sum_k ((ptS[i][k] - tS[k])^2 * 1.0) / (noise[i] * allSig[k])^2 for nModels given t*
 * vectors ptS[nModels][R] (caller's ray order; noise may be NULL = 1.0), computed by the kernel every evaluate / proposal uses
 * (tg_phi_kernel, canonical summation order).  Lets the reference's own stored (ptS, phi) pairs (model.jld) be checked directly. */
int tonga_misfit(tonga_ctx *ctx, int32_t nModels, const double *ptS, const double *noise, double *phi);

/* Batched evaluate over nModels independent models (replaces nModels calls of MCsub.jl:123-185).
 * K[nModels]; cells[nModels][4][Kcap]; noise[nModels] or NULL; outputs (any may be NULL): ptS[nModels][R],
 * phi[nModels], owners[nModels][P] (0-based nearest nucleus per flat point, -1 if none within sqrt(1e9)). */
int tonga_evaluate_batch(tonga_ctx *ctx, int32_t nModels, int32_t Kcap, const int32_t *K, const double *cells,
                         const double *noise, double *ptS, double *phi, int32_t *owners);
/* Same with DEVICE pointers (inputs already resident in HBM); asynchronous on the context's stream.
 * owners_dev may be NULL. */
int tonga_evaluate_batch_dev(tonga_ctx *ctx, int32_t nModels, int32_t Kcap, const int32_t *K_dev,
                             const double *cells_dev, const double *noise_dev, double *ptS_dev, double *phi_dev,
                             int32_t *owners_dev);

/* ---- Interpolation(TD_parameters, model, X, Y, Z), MCsub.jl:306-336, and v_nearest, MCsub.jl:247-263.
 * X[nX] is trimmed at its first NaN (:312-316); nY / nZ may be 1 (slice broadcast, :317-322).
 * zeta_out[nX], idx_out[nX] (may be NULL); *npoints_out = number of values written. */
int tonga_interpolate(tonga_ctx *ctx, int32_t K, const double *x, const double *y, const double *z,
                      const double *zeta, int32_t nX, const double *X, int32_t nY, const double *Y, int32_t nZ,
                      const double *Z, double *zeta_out, int32_t *idx_out, int32_t *npoints_out);

/* ---- chain batch: replaces `pmap(x -> TD_inversion_function(TD_parameters, dataStruct, x), 1:n_chains)`
 * (main_inversion.jl:15) for nChains chains resident on this GPU.  chain_id0 = global id of the first chain
 * (device Philox streams are keyed by (seed, global chain id) so results do not depend on the GPU count).
 * hist_cap = kept models per chain the history can hold (TD_inversion_function.jl:25 num_models_per_chain). */
int tonga_chains_create(tonga_ctx *ctx, tonga_chains **out, int32_t nChains, int64_t chain_id0, uint64_t seed,
                        int32_t hist_cap);
/* Two samplers implement the same loop and give bit-identical chains:
 *   RESIDENT  one CTA per chain, the whole chain state (a nucleus index per ray point, t* per ray) in shared memory and
 *             updated incrementally; needs max_cells <= 126 and a ray set of up to ~200k points (the 381-ray Tonga set
 *             uses 32 KB).  The fast path (BASELINE.json configs 1, 2, 4, 5).
 *   WIDE      no per-point chain state: every proposal runs the full batched forward model (as the reference does,
 *             MCsub.jl:123-185) on the candidate models of all chains; any ray set and max_cells up to ~5000
 *             (BASELINE.json config 3: 100k rays, 2000 nuclei).  At most 65535 chains per batch.
 *   STREAMED  per-point chain state (u16 nucleus index + fl32 squared distance per ray point, 6 B) in HBM, updated
 *             incrementally: every proposal is one streaming pass over the points (18 B per point) plus a second one when it
 *             is accepted, instead of the O(points x cells) forward model.  max_cells <= 65534; memory 6 B x points x chains.
 * AUTO = RESIDENT when it fits, else STREAMED when its state fits in device memory, else WIDE; tonga_chains_create is
 * create_ex(AUTO). */
#define TONGA_SAMPLER_AUTO 0
#define TONGA_SAMPLER_RESIDENT 1
#define TONGA_SAMPLER_WIDE 2
#define TONGA_SAMPLER_STREAMED 3
/* OR-ed into `sampler`: keep the history (kept models) in mapped page-locked HOST memory.  The kernels then store every kept
 * model straight into host memory while they run (contiguous posted writes over PCIe), so no device-to-host copy of the
 * history follows a run; tonga_chains_history_host returns the arrays (same layouts as tonga_chains_get_history with
 * Kcap = tonga_chains_kcap(); valid until the batch is destroyed, stable after any call that synchronises). */
#define TONGA_HISTORY_ON_HOST 0x100
int tonga_chains_create_ex(tonga_ctx *ctx, tonga_chains **out, int32_t nChains, int64_t chain_id0, uint64_t seed,
                           int32_t hist_cap, int32_t sampler);
int tonga_chains_history_host(tonga_chains *ch, void **hist_K, void **hist_cells, void **hist_phi, void **hist_ptS, void **hist_iter,
                              void **hist_action, void **hist_accept, void **hist_next_action);
/* TONGA_SAMPLER_RESIDENT, _WIDE or _STREAMED (0 for NULL) */
int tonga_chains_sampler(const tonga_chains *ch);
void tonga_chains_destroy(tonga_chains *ch);

/* Start models.  build_starting (MCsub.jl:76-121) on the device ... */
int tonga_chains_build_starting(tonga_chains *ch);
/* ... or from the host (e.g. a checkpoint, TD_inversion_function.jl:56): K[nChains], cells[nChains][4][Kcap],
 * noise[nChains] or NULL.  Both run a full evaluate to establish owners, t*, phi. */
int tonga_chains_set_models(tonga_chains *ch, int32_t Kcap, const int32_t *K, const double *cells, const double *noise);
/* Distance arithmetic of the incremental update.  0 (default): FP32 screening of every nearest-nucleus comparison with an
 * exact FP64 recheck of all comparisons that fall inside a rigorous rounding-error band (cell indices stay bit-exact);
 * 1: every comparison in exact FP64 (same results, slower; used by the tests to prove the screening changes nothing). */
int tonga_chains_set_exact_only(tonga_chains *ch, int32_t exact_only);
/* per-chain inverse temperature beta (parallel-tempering extension; 1.0 = reference; NULL resets to 1): the misfit term
 * of every acceptance ratio is scaled by beta, and only chains at beta == 1 append to the history. */
int tonga_chains_set_beta(tonga_chains *ch, const double *beta);
int tonga_chains_get_beta(tonga_chains *ch, double *beta /* [nChains] */);
/* Parallel tempering (extension, BASELINE config 5; the reference has none -- its spec here follows the dead sigma branch
 * TD_inversion_function.jl:252-272 for the noise term): one even / odd sweep of swap attempts between adjacent rungs of every
 * ladder of `ladder_size` consecutive replicas, decided ON THE DEVICE from the replicas' current phi / noise.  Replicas keep
 * their state; what is exchanged is the inverse temperature beta.  E = phi/2 + R log(noise); rungs i, j (adjacent in
 * temperature order, parity = step & 1) swap iff log u < min(0, (beta_i - beta_j)(E_i - E_j)), u drawn from Philox4x32-10
 * keyed by `seed` with counter (step, ladder, rung) -- a pure function of its inputs, so every rank of a multi-GPU ladder takes
 * the same decisions from the same all-gathered scalars.
 *   phi_all / noise_all / beta_all: DEVICE arrays over all n_all replicas of the ladders in global replica order (e.g. filled
 *   by an NCCL all-gather of every rank's tonga_chains_scalar_ptrs arrays); NULL = this batch's own arrays (n_all = nChains,
 *   offset = 0).  beta_all is updated in place and this batch's slice [offset, offset + nChains) becomes its beta.
 *   swap statistics accumulate on the device (tonga_chains_temper_stats).  Asynchronous on the context's stream. */
int tonga_chains_temper_swap(tonga_chains *ch, int32_t n_all, const double *phi_all, const double *noise_all, double *beta_all,
                             int64_t offset, int32_t ladder_size, int64_t step, uint64_t seed);
int tonga_chains_temper_stats(tonga_chains *ch, int64_t *accepted, int64_t *attempted, int32_t reset);
/* device pointers of the per-chain scalars phi / noise / beta ([nChains] doubles each) for zero-copy collectives */
int tonga_chains_scalar_ptrs(tonga_chains *ch, void **phi, void **noise, void **beta);

/* ---- ray sharding of a STREAMED batch over the GPUs of one node (BASELINE config 3; SURVEY 8e: the seam is evaluate's ray
 * loop, MCsub.jl:142, and the one exchange per proposal is the sum over rays of the misfit, MCsub.jl:169-172).
 * Rank `rank` of `world` (one process per GPU, every rank holding the same batch: same ray set, nChains, chain_id0, seed,
 * start models) maintains the per-point state of a contiguous range of the sampler's ray tiles only (tonga_shard_range); the
 * models, phi, t*, counters and the history are replicated: every rank draws the same proposals and takes the same decisions.
 * Per proposal the candidate pass of a rank writes (t*, misfit term) of its rays into the exchange block of EVERY rank over
 * peer memory (NVLink P2P stores issued by the kernel itself), raises its sequence flag there, and the accept kernel waits
 * for all flags before it sums the terms in the canonical order -- so a sharded run is bit-identical to the unsharded one.
 *   shard_init     allocates this rank's exchange block (a separate cudaMalloc, exportable with tonga_ipc_export) and returns it;
 *   shard_connect  peer_bases[world]: device pointers, valid on THIS rank's device, of every rank's block (own entry ignored):
 *                  tonga_ipc_open of the peers' handles, or plain pointers when the shards live in one process;
 *   then build_starting / set_models / run / get_state / get_history as usual, called by all ranks with the same arguments
 *   (tonga_chains_run returns TONGA_ERR_PEER if a peer fails to show up within TONGA_SHARD_TIMEOUT_MS, default 30000).
 * get_state's owners are -2 outside the rank's own points; verify checks the own points (and all of t*, phi). */
int tonga_chains_shard_init(tonga_chains *ch, int32_t rank, int32_t world, void **xch_base, uint64_t *xch_bytes);
int tonga_chains_shard_connect(tonga_chains *ch, void *const *peer_bases);
int tonga_chains_shard_info(const tonga_chains *ch, int32_t *rank, int32_t *world, int32_t *ray0, int32_t *ray1, int64_t *point0,
                            int64_t *point1); /* own rays / points in the library's (length-sorted) order */
/* own tiles [tile0, tile1) of rank `rank`: contiguous, disjoint, covering 0..n_tiles (host arithmetic only) */
int tonga_shard_range(int32_t n_tiles, int32_t rank, int32_t world, int32_t *tile0, int32_t *tile1);
/* CUDA IPC for hosts without another way to pass device pointers between the per-GPU processes: handle = 64 bytes */
int tonga_ipc_export(const void *dev_ptr, unsigned char *handle);
int tonga_ipc_open(int32_t device, const unsigned char *handle, void **dev_ptr);
int tonga_ipc_close(int32_t device, void *dev_ptr);

/* The proposal loop, TD_inversion_function.jl:70-302, nIter iterations for every chain, entirely on the device
 * (incremental Voronoi update, t* re-integration of touched rays, misfit, alpha, accept/reject, thinning).
 *   mode 0: device Philox proposals.          recs (host, [nChains][nIter]) receives them if non-NULL.
 *   mode 1: replay recs (host, [nChains][nIter]).
 * Optional host traces [nChains][nIter]: accept flag, phi after the iteration, nCells after the iteration. */
int tonga_chains_run(tonga_chains *ch, int64_t nIter, int32_t mode, tonga_proposal *recs, int8_t *tr_accept,
                     double *tr_phi, int32_t *tr_K);

/* Current models: K[nChains], cells[nChains][4][Kcap], phi, ptS[nChains][R], noise, owners[nChains][P]; any NULL. */
int tonga_chains_get_state(tonga_chains *ch, int32_t Kcap, int32_t *K, double *cells, double *phi, double *ptS,
                           double *noise, int32_t *owners);
/* iteration counter, and per chain and action (birth, death, change, move, sigma): counts[nChains][3][5] =
 * proposals drawn / accepted / evaluated (i.e. not rejected a priori: at the nCells limits, zeta or position out of bounds) */
int tonga_chains_get_stats(tonga_chains *ch, int64_t *iter, int64_t *counts);
/* Forget history, thinning counters, statistics and the iteration counter (models and streams stay); the next
 * tonga_chains_run starts again at iter = 1, as a fresh TD_inversion_function call would. */
int tonga_chains_reset(tonga_chains *ch);
/* Device time (CUDA events on the library's stream) of the sampler kernel of the last tonga_chains_run, in ms. */
int tonga_chains_last_kernel_ms(tonga_chains *ch, float *ms);
/* Thinned history (model_hist, TD_inversion_function.jl:275-281,304): n_hist[nChains] models kept so far; arrays are
 * [nChains][hist_cap][...]: hist_K, hist_cells[..][4][Kcap], hist_phi, hist_ptS[..][R], hist_iter, hist_action,
 * hist_accept, hist_next_action (the action of the following iteration, which the reference's aliasing writes into the
 * stored model: SURVEY 5.4).  Any may be NULL. */
int tonga_chains_get_history(tonga_chains *ch, int32_t Kcap, int32_t *n_hist, int32_t *hist_K, double *hist_cells,
                             double *hist_phi, double *hist_ptS, int64_t *hist_iter, int32_t *hist_action,
                             int32_t *hist_accept, int32_t *hist_next_action);
/* Posterior rasterisation, the compute part of plot_model_hist (MCsub.jl:753-825): for every kept model of every chain
 * evaluate v_nearest (MCsub.jl:247-263, exact FP64) at n_nodes points (X, Y, Z: e.g. xVec x zVec at y = ySlice, :766-768, or
 * xVec x yVec at z = zSlice, :800-802) and accumulate per node: sum_out[n_nodes] = sum of zeta, sumsq_out[n_nodes] = sum of zeta^2,
 * *count_out = number of models.  Deterministic: models are summed in (chain, kept index) order.  mean = sum/count;
 * std (Julia `std`, :775) = sqrt((sumsq - sum^2/count)/(count-1)); mask where std > 5 (:777-781).  Multi-GPU: add the
 * three outputs across ranks (they are plain sums). */
int tonga_chains_raster(tonga_chains *ch, int32_t n_nodes, const double *X, const double *Y, const double *Z, double *sum_out,
                        double *sumsq_out, int64_t *count_out);
/* ---- checkpoint / resume (TD_inversion_function.jl:40-67 load, :282-294 save).  The model part of a checkpoint is
 * tonga_chains_get_state / tonga_chains_set_models; these carry the rest: the iteration count (= the Philox counter), the
 * thinning phase model_num[nChains] (:276), the history slot awaiting its next action (-1 = none), the counters
 * ([nChains][3][5], may be NULL) and the kept models (arrays exactly as tonga_chains_get_history returns them, Kcap =
 * tonga_chains_kcap()).  A batch restored with set_models + set_progress + set_history continues bit-identically. */
int tonga_chains_get_progress(tonga_chains *ch, int64_t *iter_done, int64_t *model_num, int32_t *pending_slot);
int tonga_chains_set_progress(tonga_chains *ch, int64_t iter_done, const int64_t *model_num, const int32_t *pending_slot,
                              const int64_t *counts);
int tonga_chains_set_history(tonga_chains *ch, int32_t Kcap, const int32_t *n_hist, const int32_t *hist_K, const double *hist_cells,
                             const double *hist_phi, const double *hist_ptS, const int64_t *hist_iter, const int32_t *hist_action,
                             const int32_t *hist_accept, const int32_t *hist_next_action);

/* Consistency check on the device: re-run the full evaluate for every chain's current model and compare it with the
 * incrementally maintained state.  owner_mismatch = number of points whose owner differs; max_dphi / max_dts =
 * largest |difference| in phi / t* (the two paths share their reduction order, so both should be exactly 0). */
int tonga_chains_verify(tonga_chains *ch, int64_t *owner_mismatch, double *max_dphi, double *max_dts);
/* Device pointers of the packed history / state for zero-copy consumers (e.g. torch.distributed gathers).
 * Layouts as in tonga_chains_get_history with Kcap = tonga_chains_kcap(). */
int tonga_chains_kcap(const tonga_chains *ch);
int tonga_chains_device_ptrs(tonga_chains *ch, void **n_hist, void **hist_K, void **hist_cells, void **hist_phi,
                             void **hist_ptS, void **state_K, void **state_cells, void **state_phi);

/* Developer aid: per-phase clock64 totals of the sampler kernel (thread 0 of every chain): cycles[nChains][16] = proposal (A),
 * point pass (B), t* (C), phi + accept (D+E), commit tail (F4), bookkeeping (G), F1 masks, F2 tags, F3 renumber, 7 unused.  enable != 0 (re)starts counting,
 * enable == 0 stops it; cycles (may be NULL) receives the totals accumulated so far. */
int tonga_chains_profile(tonga_chains *ch, int32_t enable, int64_t *cycles);

/* Page-locked host buffers for callers that want full-speed host<->device copies (optional: every entry point also
 * accepts ordinary pageable memory). */
int tonga_host_alloc(void **ptr, uint64_t bytes);
int tonga_host_free(void *ptr);

/* ---- measurement helpers (bench.py): peak FP64 / FP32 FMA rate of this device in TFLOP/s (FMA = 2 flop). */
int tonga_peak_flops(tonga_ctx *ctx, double *fp64_tflops, double *fp32_tflops);

#ifdef __cplusplus
}
#endif
#endif
