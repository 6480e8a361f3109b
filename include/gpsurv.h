/*
 * gpsurv.h — C ABI of libgpsurv.so: the B200 (sm_100a) FP64 implementation of the GP
 * training / prediction hot path of kllloyd/GPSurvival.
 *
 * The reference has no FFI layer (it is 58 plain R scripts); this ABI sits *underneath*
 * the unchanged R function signatures.  Each entry point names the reference code it
 * replaces (paths relative to /root/reference/code/).  See INTEGRATION.md for the
 * `.Call` glue and SURVEY.md §8b for the contract.
 *
 * Conventions
 *   - every function returns a gps_status (0 = GPS_OK); text via gps_last_error().
 *   - all matrices are FP64, column-major (R layout), leading dimension = rows.
 *   - host pointers are read during the call only; the library never retains them and
 *     never allocates memory the caller must free (except the opaque handle).
 *   - `events` is 0/1 as double (R numeric); 0 = right-censored.
 *   - hyper-parameter vector `par` (ForOptimisationLog.R:9-31, GPRegression.R:22-25):
 *       [log sn2, log sf2, log ell_1..log ell_p, (w_1..w_d, b  — Linear mean only),
 *        (log sc2 — GPS2 'noiseCorrLearned' only, last)]   with p = 1 (SqExp) or d (ARD).
 *   - no CPU fallback: without an sm_100 device gps_create fails with GPS_ERR_CUDA.
 *   - a handle is not thread-safe; use one handle per host thread per device.
 */
#ifndef GPSURV_H
#define GPSURV_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gps_handle_s* gps_handle;

typedef enum {
  GPS_OK = 0,
  GPS_ERR_ARG = 1,          /* bad argument (NULL, size, NaN/Inf in par) */
  GPS_ERR_CUDA = 2,         /* CUDA runtime failure or no sm_100 device */
  GPS_ERR_NOT_PD = 3,       /* Ky not positive definite; gps_last_pivot() = failing column (1-based) */
  GPS_ERR_UNSUPPORTED = 4,  /* RF / InformedARD covariance, as-written-ARD gradient */
  GPS_ERR_STATE = 5         /* call order (e.g. predict before finalize) */
} gps_status;

typedef enum {            /* covFuncForm (+ extraParam = maternParam for 'Matern', CovFunc.R:31-37) */
  GPS_COV_SQEXP = 0, GPS_COV_ARD = 1,
  GPS_COV_MATERN1 = 2,    /* 'Matern', extraParam 1: sf2 exp(-r)                                   */
  GPS_COV_MATERN3 = 3,    /* 'Matern', extraParam 3: sf2 (1 + sqrt3 r) exp(-sqrt3 r)              */
  GPS_COV_MATERN5 = 4     /* 'Matern', extraParam 5: sf2 (1 + sqrt5 r + 5 r^2/3) exp(-sqrt5 r)    */
} gps_cov_form;           /* Matern gradients are derived here (dK/dlog ell = -sf2 f'(r)/r * r^2): the reference has none */
typedef enum { GPS_MEAN_ZERO = 0, GPS_MEAN_LINEAR = 1 } gps_mean_form;     /* meanFuncForm */
typedef enum {
  GPS_NC_NONE = 0,      /* noiseCorr FALSE / 'none'      (GPR, GPS1) */
  GPS_NC_LEARNED = 1,   /* 'noiseCorrLearned'            (GPS2)      */
  GPS_NC_VEC = 2        /* 'noiseCorrVec'                (GPS3)      */
} gps_noise_corr;
typedef enum {
  GPS_ARD_INTENDED = 0,   /* column k divided by ell_k (north_star item 1) */
  GPS_ARD_AS_WRITTEN = 1  /* R recycling of CovFunc.R:25,27: element (i,j) / ell[(i+j*n) mod p] */
} gps_ard_mode;

const char* gps_version(void);
const char* gps_last_error(gps_handle h);      /* h may be NULL: last error of the calling thread's stateless call */
int gps_last_pivot(gps_handle h);              /* failing pivot of the last GPS_ERR_NOT_PD (1-based), else 0 */
/* Per-fit failing pivot of the last evaluation (0 = factorised), `batch` ints.  When gps_nlml /
 * gps_nlml_grad return GPS_ERR_NOT_PD their outputs are still filled: +Inf (NaN gradient) for the
 * fits that failed, valid values for the others — so an optimiser can reject just that point, where
 * the reference would fall back to Matrix::nearPD (ForOptimisationLog.R:43-51). */
int gps_batch_pivots(gps_handle h, int* pivots_out);

/* ForOptimisationLog.R:43-51: when chol(Ky) fails the reference factorises Matrix::nearPD(Ky)$mat instead and carries on.
 * The library does the same ON THE DEVICE (Higham's alternating projections with Matrix::nearPD's defaults over a Jacobi
 * eigen-solver, csrc/nearpd.cuh; n <= 4096) for every fit of the batch whose factorisation failed, in gps_nlml and — so that
 * optim's fn and gr agree on which points are acceptable — in gps_nlml_grad.  enable = 0 turns the repair off: such fits
 * then report GPS_ERR_NOT_PD / +Inf as described above.  Default: on.  gps_regress_finalize and gps_predict never repair
 * (GPRegression.R:65-69 and GPPredict.R:43 call plain chol).  gps_near_pd_used: per fit, 1 when the last evaluation's value
 * came from a repaired matrix. */
int gps_set_near_pd(gps_handle h, int enable);
int gps_near_pd_used(gps_handle h, int* used_out);

/* ---- lifetime ----------------------------------------------------------------------- */
/* One handle = one fit (one trainingTestStructure) on one device.  `batch` > 1 makes it a
 * batch of `batch` independent fits of identical shape evaluated by the same launches
 * (restarts / CV folds — SURVEY.md §8e); every array argument below then carries the
 * batch as the slowest dimension. */
int gps_create(int device, gps_handle* out);
int gps_create_batch(int device, int batch, gps_handle* out);
int gps_destroy(gps_handle h);

/* Options that the reference passes as strings in parameterStructure (GPRegression.R:13-20). */
int gps_set_options(gps_handle h, int cov_form, int mean_form, int noise_corr, int ard_mode);

/* ---- data ---------------------------------------------------------------------------- */
/* trainingData (n x d), trainingTargets (n), events (n) — GPRegression.R:7-11.  X is constant
 * over all evaluations of a fit and is uploaded once. */
int gps_set_data(gps_handle h, const double* X, int n, int d, const double* y, const double* events);
/* The same, with the fits of the batch given as ROW-INDEX VIEWS of one pooled data set (CV folds, restarts of one
 * data set — the reference's runSyntheticExp1.R:208-213 / CV loops call ApplyGP once per fold on row subsets): Xpool is
 * npool x d (column-major), ypool / eventspool length npool, rows[batch][n] the 0-based pool rows of every fit.  The
 * pool is uploaded once; the per-fit copies are gathered on the device. */
int gps_set_data_indexed(gps_handle h, const double* Xpool, int npool, int d, const double* ypool, const double* eventspool,
                         const int* rows, int n);
/* trainingTargetsLearned changes every GPS round (ApplyGP.R:201). */
int gps_set_targets(gps_handle h, const double* y);
/* logHypNoiseCorrection for GPS3: one value per censored row, in row order (ApplyGP.R:149,209;
 * ForOptimisationLog.R:36-39).  len must equal the number of censored rows.  (For GPS2 the
 * correction travels inside `par`.) */
int gps_set_noise_corr(gps_handle h, const double* log_sc2, int len);

/* ---- CovFunc.R:1-54 (SqExp / ARD) ------------------------------------------------------ */
/* K (n1 x n2) = sf2 * exp(-0.5 * ||x1_i/ell - x2_j/ell||^2).  X2 == NULL means X2 = X1
 * (the identical(x1,x2) branch, CovFunc.R:24-25).  Stateless: `h` only supplies the device
 * and scratch memory (one grow-only arena, no allocation per call); h may be NULL — the calling thread's default handle
 * on the current device is then used (created on first use).  gps_adjust accepts NULL the same way. */
int gps_cov(gps_handle h, const double* X1, int n1, const double* X2, int n2, int d,
            double sf2, const double* ell, int p, int cov_form, int ard_mode, double* K_out);

/* ---- ForOptimisationLog.R:1-68 --------------------------------------------------------- */
/* Negative log marginal likelihood at `par` (imposePriors = FALSE).  nlml_out: `batch` values;
 * par: batch x len. */
int gps_nlml(gps_handle h, const double* par, int len, double* nlml_out);

/* ---- OptimisationGradientLog.R:1-94 ---------------------------------------------------- */
/* Gradient w.r.t. `par`, same order as `par` (:88-89).  Reuses the factorisation of the last
 * gps_nlml when `par`, targets and noise correction are unchanged (optim calls fn and gr
 * separately — SURVEY.md §3.3).  nlml_out may be NULL.
 * SqExp: the reference branch (:50-57) restated; ARD: the analytic contraction of north_star item 4 (ard_mode =
 * INTENDED); Matern: derived (the reference has none, :6).  NOT offered: the reference's as-written ARD branch
 * (:58-69) — it is not the derivative of its own likelihood (unscaled dist(X), exp(+r/4/ell^3), no chain-rule
 * factors), an optimiser fed with it does not descend; with ard_mode = AS_WRITTEN and d > 1 the call returns
 * GPS_ERR_UNSUPPORTED and says so.  The oracle keeps a restatement of it for the record. */
int gps_nlml_grad(gps_handle h, const double* par, int len, double* nlml_out, double* grad_out);

/* ---- GPRegression.R:60-82 (everything after optim) ------------------------------------- */
/* Recompute the model at `par` and keep it resident for gps_predict.  Optional outputs
 * (NULL to skip; K and L are n x n, 8n^2 bytes each): K_out noise-free covariance (:61),
 * L_out lower Cholesky factor of Ky (:65-69), alpha_out (:70), mean_out meanTraining (:60),
 * logmarlik_out positive-sign log marginal likelihood (:71). */
int gps_regress_finalize(gps_handle h, const double* par, int len, double* K_out, double* L_out,
                         double* alpha_out, double* mean_out, double* logmarlik_out);

/* ---- GPPredict.R:1-60 ------------------------------------------------------------------- */
/* Predict at testData (m x d) with the hyper-parameters `par`.  If the resident model was
 * built from the same par / targets / noise correction it is reused; otherwise (GPS2/GPS3 with
 * a new correction or new targets, GPPredict.R:36-46) it is rebuilt first.  With
 * `reuse_model` != 0 the resident model is used as is, whatever the current targets are —
 * that is what the reference does for GPR / GPS1, where L and alpha come from
 * regressionStructure (GPPredict.R:12-13).  Outputs: funcMeanPred, varFunction, varData (m each). */
int gps_predict(gps_handle h, const double* par, int len, int reuse_model, const double* Xstar, int m,
                double* mu_out, double* varf_out, double* vard_out);

/* GPPredict at TRAINING rows of each fit: rows[batch][m], 0-based indices into the fit's own trainingData.  This is what
 * ApplyGP asks for every GPS round (testData = the censored training rows, ApplyGP.R:165-171) and for the training-set
 * predictions (:265-282); only the indices cross the bus. */
int gps_predict_rows(gps_handle h, const double* par, int len, int reuse_model, const int* rows, int m,
                     double* mu_out, double* varf_out, double* vard_out);

/* ---- AdjustTrainingSurvivalMeanVariance.R:1-37 ------------------------------------------ */
/* Lower-truncated-normal moments for c censored samples: a = censoring time, mu/s = predictive
 * mean and the value the reference passes as `sigma` (ApplyGP.R:173,190 pass the variance).
 * `1 - pnorm` is computed exactly as the reference does, including the 1e-12 branch (:28-31). */
int gps_adjust(gps_handle h, const double* a, const double* mu, const double* s, int c,
               double* mean_out, double* var_out);

/* ---- CalculateMetrics.R:18,21 ---------------------------------------------------------------- */
/* Concordance index and RMSE of m test predictions per fit (arrays [batch][m]): c.index with survcomp::concordance.index's
 * semantics — a pair is usable when the shorter time is an event, concordant when the HIGHER score has the shorter time
 * (x is a risk score there; the reference passes predicted log survival times as is), pairs tied in x excluded
 * (outx = TRUE); rmse = sqrt(sum((time - pred)^2) / m).  The rnorm jitter CalculateMetrics.R:12-16 applies to tied
 * predictions is not restated: pairs_out ([batch][3], optional) reports concordant / discordant / tied pair counts.
 * survAUC::predErr (the Brier score of :23) stays on the host. */
int gps_metrics(gps_handle h, const double* pred, const double* time, const double* event, int m, double* cindex_out,
                double* rmse_out, long long* pairs_out);

/* ---- MakeSyntheticData.R:41,148 (synthetic targets) ---------------------------------- */
/* out = t(chol(Ky)) %*% v = L v, with Ky = K(par) + exp(par[0]) I built from the handle's data and options
 * (par[0] plays the role of the jitter the reference gets from nearPD when K is numerically singular).
 * v, out: batch x n.  The caller finishes the recipe: y = exp(mean + out + sqrt(sn2_gen) * d2). */
int gps_chol_matvec(gps_handle h, const double* par, int len, const double* v, double* out);

/* The recipe itself on the device: MakeSyntheticData.R:40-41,145-149 ('likeRealData': x ~ N(0,1), K = CovFunc(x, x) SqExp,
 * y = exp(t(chol(K)) %*% d1 + sqrt(sn2) d2)) and CensorData.R:36-48 ('NormalLoopSample': samples picked uniformly among the
 * uncensored ones are censored at |N(cens_mean, cens_sd)| when that is shorter, until n_censor are) for `batch` data sets.
 * Random stream: Philox4x32-10, counter = (block, stream, data set, 0), key = seed (R's wall-clock-seeded Mersenne-Twister
 * cannot be reproduced).  `jitter` is added to K's diagonal before the Cholesky (GPS_ERR_NOT_PD: raise it; the reference
 * calls nearPD there).  Outputs: X_out [batch][d][n] column-major, y_out / events_out [batch][n] after censoring (time scale),
 * y_uncensored_out (optional), n_censored_out (optional, per data set).  The handle's data / options are replaced. */
int gps_synth_generate(gps_handle h, unsigned long long seed, int n, int d, double sn2, double sf2, double ell, double jitter,
                       int n_censor, double cens_mean, double cens_sd, double* X_out, double* y_out, double* events_out,
                       double* y_uncensored_out, int* n_censored_out);

/* ---- whole fits: ApplyGP.R:136-282 with stats::optim restated in C++ ---------- */
/* Runs, for every fit of the batch in lock-step, what ApplyGP does after its target transforms: for a
 * survival model the GPS outer loop  while((change > tolerance & count < max_count) | count <= 5)
 * { GPRegression (optim with `maxit`), GPPredict on the censored rows, AdjustTrainingSurvivalMeanVariance,
 * target / noise-correction update }  (ApplyGP.R:162-214), otherwise one GPRegression (:224-232); then
 * the test-set and training-set predictions (:237-282).  Options (covariance, mean, noiseCorr) come
 * from gps_set_options; y_log are the log (and, for real data, median-shifted) targets of ApplyGP.R:77-96.
 * method: 0 = 'Nelder-Mead' (R's nmmin), 1 = 'BFGS' (R's vmmin, dense state kept on the device),
 * 2 = L-BFGS (not an R method: vmmin's line search and stopping rules over a 10-pair two-loop recursion; points, gradients,
 * directions and phases stay on the device, a tick costs 16 bytes of host traffic, line-search trials are fn-only).
 * par_start: batch x len start vectors as GPRegression.R:22-25 assembles them (restarts may differ per fit;
 * for GPS3 the start noise correction is rep(par_start[0], nCensored), ApplyGP.R:149).
 * A point whose Ky is not positive definite is a rejected point (+Inf) for the optimiser; a fit that cannot be
 * finalised at all is reported through gps_fit_batch_status and does not abort the batch.
 * Outputs (any may be NULL): par_out batch x len = logHypChosen flattened; objective_out, logmarlik_out,
 * rounds_out: batch; targets_learned_out, train_mean: batch x n; test_*: batch x m;
 * evaluations_out: number of batched evaluations used. */
int gps_fit_batch(gps_handle h, const double* X, int n, int d, const double* y_log, const double* events,
                  const double* par_start, int len, int method, int maxit, double tolerance, int max_count, int survival,
                  const double* Xtest, int m, double* par_out, double* objective_out, double* logmarlik_out,
                  int* rounds_out, double* targets_learned_out, double* test_mean, double* test_varf, double* test_vard,
                  double* train_mean, long long* evaluations_out);

/* Per-fit status of the last gps_fit_batch (`batch` ints): GPS_OK, or GPS_ERR_NOT_PD for a fit whose Ky could not be
 * factorised at the point its optimiser returned (e.g. a non-finite start value) — that fit's numeric outputs are NaN and
 * it was dropped from the remaining rounds; the other fits of the batch are unaffected (the reference, too, would lose
 * only that one ApplyGP call). */
int gps_fit_batch_status(gps_handle h, int* status_out);

/* ---- several GPUs from one host process (the restart x fold loops of runSyntheticExp1.R:208-213) ---------
 * `total` independent fits of identical shape are split over `devices[0..ndev)` (LPT by estimated cost 3 n^2 d + n^3,
 * which for identical shapes is an even split), one host thread and one handle per device, sub-batches sized to the
 * device's free memory, NO collective and no peer traffic: results are scattered back into the caller's arrays at the
 * fits' positions.  Options as gps_set_options; every other argument as gps_fit_batch with `total` as the batch;
 * status_out[total] as gps_fit_batch_status; evaluations_out = the largest per-device number of batched evaluations.
 * A device may be listed more than once (two host threads sharing one GPU).
 * Between calls ONE idle handle per device stays alive (creating a handle, its buffers and its graphs costs as much as 64
 * C4-shaped fits); it is reused when the next call's per-device batch matches, replaced otherwise, and freed by
 * gps_multi_release(). */
int gps_fit_batch_multi(const int* devices, int ndev, int total, int cov_form, int mean_form, int noise_corr, int ard_mode,
                        const double* X, int n, int d, const double* y_log, const double* events, const double* par_start, int len,
                        int method, int maxit, double tolerance, int max_count, int survival, const double* Xtest, int m,
                        double* par_out, double* objective_out, double* logmarlik_out, int* rounds_out, double* targets_learned_out,
                        double* test_mean, double* test_varf, double* test_vard, double* train_mean, int* status_out,
                        long long* evaluations_out);
/* The same for CV folds / restarts of ONE data set: the pool (npool x d) is uploaded once per device and fit t trains on
 * pool rows train_rows[t][0..n) and is tested on test_rows[t][0..m) — no per-fit copy of X crosses the bus. */
int gps_fit_folds_multi(const int* devices, int ndev, int total, int cov_form, int mean_form, int noise_corr, int ard_mode,
                        const double* Xpool, int npool, int d, const double* y_log_pool, const double* events_pool,
                        const int* train_rows, int n, const int* test_rows, int m, const double* par_start, int len, int method,
                        int maxit, double tolerance, int max_count, int survival, double* par_out, double* objective_out,
                        double* logmarlik_out, int* rounds_out, double* targets_learned_out, double* test_mean, double* test_varf,
                        double* test_vard, double* train_mean, int* status_out, long long* evaluations_out);
/* The partitioner itself (pure host code): part_out[i] in [0, nparts) for jobs with the given costs, longest first. */
int gps_partition_lpt(const double* cost, int count, int nparts, int* part_out);
/* Free the idle handles the multi-device entry points keep between calls (device memory back to the caller). */
int gps_multi_release(void);

/* Identity of the model gps_regress_finalize left resident (0 = none): bumped by every finalize.  A drop-in GPPredict keeps
 * the token GPRegression returned and refuses to predict GPR / GPS1 from a different resident model (GPPredict.R:12-13 uses
 * the L / alpha of the regressionStructure it is given). */
unsigned long long gps_model_token(gps_handle h);
/* Shape of the resident data (0 before gps_set_data): n, d, batch, largest censored-row count.  NULL outputs are skipped. */
int gps_get_shape(gps_handle h, int* n_out, int* d_out, int* batch_out, int* ncens_out);

/* ---- introspection used by bench.py / tests ------------------------------------------- */
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
long long gps_launch_count(gps_handle h);
/* Device time in ms of the last gps_nlml / gps_nlml_grad / finalize / predict (CUDA events
 * around the device work of that call, excluding host<->device copies). */
double gps_last_device_ms(gps_handle h);
/* Enable (1) / disable (0) CUDA-graph replay of evaluations (default on). */
int gps_set_graphs(gps_handle h, int enable);
/* Time `reps` back-to-back device-only repetitions of the last evaluation kind
 * (0 = nlml, 1 = nlml+grad) at the last `par`, inputs resident; returns ms per repetition. */
int gps_time_eval(gps_handle h, int kind, int reps, double* ms_out);
/* Stage timers: fills ms[0..7] = {prescale+gram, cholesky, inverse, kinv, wk+grad gemm, rest, 0, total}
 * for one nlml+grad evaluation at the last par, each stage timed with CUDA events. */
int gps_time_stages(gps_handle h, double* ms_out);
/* Raw FP64 GEMM through the library's DMMA kernel: C(MxN) = A(MxK) * B(NxK)' — for the
 * peak-fraction measurement beside cublasDgemm.  Device-resident synthetic operands. */
int gps_bench_gemm(gps_handle h, int M, int N, int K, int reps, double* ms_out);
/* Developer probe: clock64() stamps (cycles since kernel start) at the phase boundaries of the fused
 * Cholesky panel kernel, run once on block column 0 of the current Ky. */
int gps_debug_panel_clocks(gps_handle h, long long* out, int n_out);
/* Developer probe (GPS_DAG_TRACE=1): per-task time stamps of the last run of the one-kernel Cholesky (plan 0) or
 * Cholesky + triangular inverse (plan 1): 4 values per task {picked, ready, done (globaltimer ns),
 * sm | leaf << 8 | K/16 << 9 | fit << 24 | segment << 32}.  Returns the task count (-1: no trace). */
int gps_debug_dag_trace(gps_handle h, int plan, long long* out, long long cap);

#ifdef __cplusplus
}
#endif
#endif /* GPSURV_H */
