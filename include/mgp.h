/* mgp.h -- C ABI of libmgp_b200.so: the B200-native Gaussian-process surrogate hot path of
 * MIP-EGO (wangronin/MIP-EGO, mipego 2.0.0).
 *
 * The reference is pure Python and has no FFI for this path: the boundary it exposes is the
 * duck-typed protocol between its BO loop / acquisition optimisers and
 * `GaussianProcess` + `InfillCriteria`.  Each entry point below names
 * the reference code it replaces (paths relative to the reference root).  The Python classes
 * in mip-ego_b200/ bind these with ctypes and mirror the reference's class surface.
 *
 * Conventions
 *  - every function returns 0 on success and a negative MGP_E* code on failure;
 *    mgp_last_error(h) returns a human-readable message for the last failure on that handle;
 *  - no exceptions cross the ABI; all pointers are caller-owned; host arrays are C-contiguous
 *    row-major IEEE float64 (int64 for indices) and read-only unless named out_*;
 *  - a handle is bound to one CUDA device and owns all device memory it allocates; it is not
 *    thread-safe (one handle per Python object; the caller serialises calls);
 *  - functions without the _dev suffix take HOST pointers and are synchronous: host<->device
 *    copies happen inside the call.  *_dev functions take DEVICE pointers (on the handle's
 *    device) plus a cudaStream_t (as void*); they enqueue work on that stream and return
 *    without synchronising.  Calls on different streams are ordered by the library itself: every
 *    entry point makes its stream wait for the previous user of the handle's workspaces and of the
 *    fitted model (an event recorded when that call's work was enqueued), so e.g. mgp_acq_dev on a
 *    user stream followed by mgp_acq or mgp_finalize_fit is safe without an explicit
 *    synchronisation.  What stays the caller's job: the lifetime of its own device buffers;
 *  - all arithmetic is float64.  There is no CPU fallback: without a CUDA device every compute
 *    entry point fails with MGP_ECUDA.
 */
#ifndef MGP_H_
#define MGP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGP_ABI_VERSION 4

/* error codes */
#define MGP_OK 0
#define MGP_EINVAL (-1)   /* bad argument (shape mismatch, NULL, unknown enum) */
#define MGP_ECUDA (-2)    /* CUDA runtime failure (message has the CUDA error string) */
#define MGP_ESTATE (-3)   /* call order: predict before fit / set_state, etc. */
#define MGP_ENOMEM (-4)   /* device allocation failed */
#define MGP_ENOTPD (-5)   /* correlation matrix not positive definite (finalize_fit only) */
#define MGP_ENCCL (-6)    /* NCCL failure */

/* correlation functions: mipego/GaussianProcess/kernel.py */
#define MGP_CORR_SE 0        /* squared_exponential  kernel.py:277-317 */
#define MGP_CORR_MATERN32 1  /* matern, nu = 1.5     kernel.py:147-185 */
#define MGP_CORR_ABSEXP 2    /* absolute_exponential kernel.py:235-274 */

/* trend (prior mean): mipego/GaussianProcess/trend.py:69-88 (constant_trend) */
#define MGP_TREND_CONST_FIXED 0      /* beta given  -> simple kriging   (gpr.py:693-695) */
#define MGP_TREND_CONST_ESTIMATED 1  /* beta = None -> ordinary kriging (gpr.py:686-692) */
/* linear_trend f(x) = [1, x_1..x_d], p = d + 1 basis functions (trend.py:90-110) */
#define MGP_TREND_LINEAR_FIXED 2      /* beta (p values) given */
#define MGP_TREND_LINEAR_ESTIMATED 3  /* beta = None -> universal kriging */

/* likelihood / estimation mode: gpr.py:246-251 */
#define MGP_MODE_NOISELESS 0  /* par = theta             gpr.py:814-835 */
#define MGP_MODE_NOISY 1      /* par = [theta, sigma2]   gpr.py:855-874 */
#define MGP_MODE_NOISE_ESTIM 2  /* par = [theta, alpha], R = alpha R0 + (1 - alpha) I   gpr.py:837-853 */
/* OR-ed into the mode: log_likelihood_restricted (REML), gpr.py:699-796.
 * par = [theta, sigma2] (NOISELESS: noise 0; NOISY: the given noise_var) or
 * [theta, sigma2, noise_var] (NOISE_ESTIM). */
#define MGP_LIK_RESTRICTED 16

/* acquisition functions: mipego/InfillCriteria.py */
#define MGP_ACQ_EI 0    /* EI        InfillCriteria.py:113-148 */
#define MGP_ACQ_PI 1    /* EpsilonPI InfillCriteria.py:150-187 (param = epsilon; 0 = plain PI) */
#define MGP_ACQ_UCB 2   /* UCB       InfillCriteria.py:75-111  (param = alpha) */
#define MGP_ACQ_MGFI 3  /* MGFI      InfillCriteria.py:196-252 (param = t, capped at 22.36) */

typedef struct mgp_handle mgp_t;

/* ---- lifetime ------------------------------------------------------------------------ */
int mgp_abi_version(void);
int mgp_create(int device, mgp_t **out_h);
int mgp_destroy(mgp_t *h);
const char *mgp_last_error(const mgp_t *h);

/* ---- training data: GaussianProcess._check_data, gpr.py:263-288 ------------------------
 * Uploads X (n x d) and y (n).  Returns MGP_EINVAL with "duplicate rows" if two rows of X
 * coincide (the reference raises Exception, gpr.py:275-277).  `beta`: the given trend
 * coefficients for the *_FIXED trends (1 value for the constant trend, d + 1 for the linear one),
 * ignored (may be NULL) otherwise. */
int mgp_set_train(mgp_t *h, int n, int d, const double *X, const double *y, int corr, int trend,
                  const double *beta);

/* ---- batched marginal likelihood: GaussianProcess.log_likelihood_concentrated,
 * gpr.py:798-946 (with correlation_matrix :658-668, _compute_aux_var :676-697,
 * corr_grad_theta :622-656) evaluated for B parameter vectors at once.
 * params: B x n_par (n_par = d for NOISELESS, d+1 for NOISY and NOISE_ESTIM; with
 * MGP_LIK_RESTRICTED d+1, d+1, d+2).  out_llf: B.  out_grad: B x n_par
 * (d llf / d par, NOT d/d log10 par -- SURVEY Q5) or NULL when eval_grad == 0.
 * out_status[b]: 0 ok; k > 0: Cholesky failed at pivot k (llf = -inf, grad = 0,
 * gpr.py:820-826); -1: llf > 0 rejected (llf = -inf; grad = 0 for the concentrated likelihood,
 * gpr.py:889-891; the restricted likelihood still returns its gradient, gpr.py:745-783). */
int mgp_nll_batch(mgp_t *h, int B, const double *params, int n_par, int mode, double noise_var,
                  int eval_grad, double *out_llf, double *out_grad, int *out_status);

/* ---- harvest the fitted state for the chosen parameters: the final likelihood call of
 * _optimize_hyperparameter (gpr.py:1087-1101) + compute_beta_gamma (gpr.py:670-674).
 * Builds on the device: L = chol(R), L^-1, rho, Yt, Ft, G, sigma2, beta, gamma, and the packed
 * operands of the posterior kernels.  out_llf (may be NULL) receives the log-likelihood.
 * Returns MGP_ENOTPD if R is not positive definite. */
int mgp_finalize_fit(mgp_t *h, const double *par, int n_par, int mode, double noise_var,
                     double *out_llf);

/* ---- incremental training: GaussianProcess.update (gpr.py:387-390, "TODO: implement incremental
 * training") for k rows APPENDED to the training set, at the hyper-parameters of the last
 * mgp_finalize_fit.  L, L^-1 and the packed posterior operands are extended by a rank-k border
 * update in O(n^2 k) (three GEMMs against the kept leading block + a k x k factorisation) instead of
 * the 2/3 n^3 rebuild; rho, sigma2, beta, gamma and the likelihood are recomputed for the new y.
 * X_add: k x d, y_add: k.  Returns MGP_ESTATE when the incremental path does not apply (model
 * adopted with mgp_set_state, non-constant trend, fewer than 512 kept rows or more than a quarter of
 * them appended): the caller then rebuilds with mgp_set_train + mgp_finalize_fit.  MGP_EINVAL
 * "duplicate rows" / MGP_ENOTPD leave the previous model intact. */
int mgp_append_train(mgp_t *h, int k, const double *X_add, const double *y_add, double *out_llf);

/* ---- the `env` harvest of a likelihood call WITHOUT adopting the state: what
 * log_likelihood_concentrated / _restricted(par, env) fill in (gpr.py:876-887, 785-794).  Runs the
 * pipeline for one parameter vector in scratch memory and copies out L (n x n row-major, upper = 0),
 * rho, Yt (n each), Ft (n x p; meaningful for estimated trends) and
 * out_scalars[5] = {sigma2, beta_0, G, llf, noise_var} (G: constant trend only).  Any out pointer
 * may be NULL.  The fitted model of the handle (mgp_finalize_fit / mgp_set_state) is left untouched,
 * like the reference, whose env harvest has no side effects.  MGP_ENOTPD if R is not positive
 * definite (nothing is written then). */
int mgp_nll_env(mgp_t *h, const double *par, int n_par, int mode, double noise_var, double *out_L,
                double *out_rho, double *out_Yt, double *out_Ft, double *out_scalars);

/* ---- state exchange (pickling; accepting an externally fitted model).
 * get: any out pointer may be NULL.  out_L: n x n row-major lower triangle (upper = 0);
 * out_theta: d; out_gamma, out_rho, out_Yt: n; out_Ft: n x p row-major (p = 1 or d + 1);
 * out_scalars[5] = {sigma2, beta_0, G, llf, noise_var} (G: constant trend only); all p trend
 * coefficients are returned by mgp_get_trend
 * (Ft and G are meaningful only for MGP_TREND_CONST_ESTIMATED).
 * set: theta (d), L (n x n row-major, lower), gamma (n), sigma2, beta; everything else
 * (L^-1, Ft, G, packed operands) is rebuilt on the device.  Requires mgp_set_train first. */
int mgp_get_state(mgp_t *h, double *out_theta, double *out_L, double *out_gamma, double *out_rho,
                  double *out_Yt, double *out_Ft, double *out_scalars);
int mgp_set_state(mgp_t *h, const double *theta, const double *L, const double *gamma,
                  double sigma2, const double *beta /* p values */);
/* Trend coefficients beta (p values; mean.beta after compute_beta_gamma, gpr.py:670-674). */
int mgp_get_trend(mgp_t *h, double *out_beta);

/* ---- posterior: GaussianProcess.predict, gpr.py:392-483.  Xs: m x d.  out_mean: m.
 * out_mse: m or NULL.  MSE = |sigma2 (1 - |L^-1 r|^2 + u^2)| (gpr.py:471-474). */
int mgp_predict(mgp_t *h, int64_t m, const double *Xs, double *out_mean, double *out_mse);

/* ---- posterior gradient: GaussianProcess.gradient + corr_dx, gpr.py:511-617, applied to each
 * of the m rows of Xs.  out_mean_dx, out_mse_dx: m x d. */
int mgp_gradient(mgp_t *h, int64_t m, const double *Xs, double *out_mean_dx, double *out_mse_dx);

/* ---- fused posterior + acquisition: InfillCriteria.{EI,EpsilonPI,UCB,MGFI}.__call__
 * (InfillCriteria.py:75-252) applied element-wise to each row of Xs, for n_sets parameter
 * values sharing one posterior evaluation (ParallelBO, BayesOpt.py:56-101).
 * set_params: n_sets values (epsilon | alpha | t; ignored for EI).  plugin: user-facing value
 * (it is negated when minimize == 0, InfillCriteria.py:64-73).  out_val: n_sets x m.
 * out_dx: n_sets x m x d or NULL (return_dx). */
int mgp_acq(mgp_t *h, int64_t m, const double *Xs, int kind, int n_sets, const double *set_params,
            double plugin, int minimize, double *out_val, double *out_dx);

/* Same, returning only the k best (largest) candidates per parameter set: what
 * argmax_restart (optimizer/utils.py:8-87) extracts from a population.
 * out_val: n_sets x k (descending); out_idx: n_sets x k row indices into Xs. */
int mgp_acq_topk(mgp_t *h, int64_t m, const double *Xs, int kind, int n_sets,
                 const double *set_params, double plugin, int minimize, int k, double *out_val,
                 int64_t *out_idx);

/* ---- device-resident variants (inputs/outputs already in HBM; no host copies, no sync).
 * d_* are device pointers on the handle's device; set_params stays a host pointer.
 * d_out_mean / d_out_mse may be NULL in mgp_acq_dev; d_out_dx may be NULL. */
int mgp_predict_dev(mgp_t *h, int64_t m, const double *d_Xs, double *d_out_mean, double *d_out_mse,
                    void *stream);
int mgp_acq_dev(mgp_t *h, int64_t m, const double *d_Xs, int kind, int n_sets,
                const double *set_params, double plugin, int minimize, double *d_out_val,
                double *d_out_dx, void *stream);
int mgp_acq_topk_dev(mgp_t *h, int64_t m, const double *d_Xs, int kind, int n_sets,
                     const double *set_params, double plugin, int minimize, int k,
                     double *d_out_val, int64_t *d_out_idx, void *stream);
int mgp_nll_batch_dev(mgp_t *h, int B, const double *d_params, int n_par, int mode,
                      double noise_var, int eval_grad, double *d_out_llf, double *d_out_grad,
                      int *d_out_status, void *stream);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch.  The
 * reference has no collective on this path; its only parallelism is joblib processes over the
 * n_point acquisition sub-problems (BayesOpt.py:94-97), each of which unpickles the whole model.
 * Here candidates and MLE restarts are sharded over ranks and two exchanges remain:
 *   - the fitted model goes once per fit from the root's HBM to every rank's HBM (no host bounce,
 *     no re-factorisation: every rank holds the bit-identical model, so results do not depend on
 *     the number of shards);
 *   - the per-rank k best (value, global index) pairs are all-gathered and merged on the device in
 *     the single-GPU order (value descending, NaN last, ties -> smaller index), identically on
 *     every rank.
 * NCCL is bound at run time with dlopen (the copy the process already loaded, e.g. PyTorch's,
 * else libnccl.so.2, else $MGP_NCCL_LIB): the library has no link-time dependency on it.
 * mgp_comm_unique_id: rank 0 creates the 128-byte ncclUniqueId; the caller ships it to the other
 * ranks by any means (torch.distributed, MPI, a file).  mgp_comm_init: collective over all ranks.
 * mgp_bcast_model: collective; every rank has called mgp_set_train with the same X, y (checked:
 * n, d, kernel, trend and checksums of X and y must agree, else EVERY rank returns MGP_EINVAL);
 * the root must be fitted.  mgp_allgather_topk(_dev): val / idx are this rank's n_sets x k best
 * (idx < 0 = empty slot), idx_offset = global index of this rank's row 0; out_* receive the n_sets
 * x k global best.  mgp_allgather_f64_dev: plain all-gather of `count` doubles per rank (rows of
 * sharded likelihood evaluations). */
int mgp_comm_unique_id(char *out_id /* 128 bytes */);
int mgp_comm_init(mgp_t *h, int rank, int world, const char *id /* 128 bytes */);
int mgp_comm_destroy(mgp_t *h);
const char *mgp_comm_info(const mgp_t *h);
int mgp_bcast_model(mgp_t *h, int root);
int mgp_allgather_topk(mgp_t *h, int n_sets, int k, const double *val, const int64_t *idx,
                       int64_t idx_offset, double *out_val, int64_t *out_idx);
int mgp_allgather_topk_dev(mgp_t *h, int n_sets, int k, const double *d_val, const int64_t *d_idx,
                           int64_t idx_offset, double *d_out_val, int64_t *d_out_idx, void *stream);
int mgp_allgather_f64_dev(mgp_t *h, const double *d_send, int64_t count, double *d_recv, void *stream);

/* ---- tuning / introspection ------------------------------------------------------------
 * mgp_set_option: "chunk" = candidates per pass of the posterior pipeline (multiple of 128, 0 =
 * default); "isotropic" = 1: ONE theta shared by all features (kernel.py:312-317) -- every parameter
 * vector then has 1 instead of d leading theta entries, and the likelihood gradient follows the
 * reference's indexing of its derivative tensor (entry 0 = derivative w.r.t. the first feature's
 * theta; set it before the first likelihood call); "small_path" = 0 disables the matrix-vector path
 * for <= 8 candidates; "nll_groups", "gemm_cta_cap": scheduling of batched likelihoods (0 = automatic);
 * "trtri_rowwise" = -1 (automatic: n <= 2048), 0, 1: form L^-1 block row by block row on a second stream
 * while the Cholesky factorisation runs instead of by recursive doubling afterwards;
 * "nll_graphs" = 0 issues every batched likelihood launch by launch (default 1: the launch sequence of
 * a batch shape is captured in a CUDA graph at its second use and replayed with one cudaGraphLaunch;
 * not used on the legacy default stream);
 * "time_kernels" = 1 brackets every posterior kernel launch with CUDA events on the
 * launching stream (resets the timers).
 * mgp_get_counter: "launches" = kernels launched by this handle so far; "npad", "chunk",
 * "sm_count", "small_path", "isotropic", "n", "appends", "graph_replays", "nll_graphs" (captured shapes), "rank", "world", "bcast_bytes" (payload of the last
 * mgp_bcast_model); with time_kernels: "<k>_ns" / "<k>_calls" for k in rgen, trmm, epilogue, topk
 * (accumulated device time in nanoseconds; synchronises on the recorded events). */
int mgp_set_option(mgp_t *h, const char *name, int64_t value);
int64_t mgp_get_counter(const mgp_t *h, const char *name);

#ifdef __cplusplus
}
#endif
#endif /* MGP_H_ */
