/*
 * gpsurv_glue.c — .Call glue between R and libgpsurv.so (include/gpsurv.h).
 *
 * NOT COMPILED OR RUN IN THE BUILD IMAGE: R, Rscript, libR.so and the R headers are absent
 * there and on the GPU box.  The file is deliberately marshalling only — every
 * formula lives behind the C ABI, which is what the test-suite exercises through ctypes.
 *
 * Build (on a machine with R):   R CMD SHLIB gpsurv_glue.c -L../gpsurvival_b200 -lgpsurv -o gpsurv_glue.so
 * Load  (top of the drop-in ApplyGP.R, next to its source() block, ApplyGP.R:18-24):
 *        dyn.load('gpsurv_glue.so')
 *
 * Conventions: R numeric matrices are column-major REAL() buffers, exactly the layout the
 * library expects; inputs are only read during the call; outputs are allocated here
 * (PROTECT/UNPROTECT) and filled by the library.  The library never throws and never calls
 * Rf_error; status codes are turned into R conditions *here*, after the C call returned.
 * GPS_ERR_NOT_PD is raised as a classed condition text ("gps_not_pd: ...") so that the
 * drop-in ForOptimisationLog.R can keep the reference's nearPD fallback (:43-51).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "../include/gpsurv.h"

static gps_handle get_handle(SEXP ptr) {
  gps_handle h = (gps_handle)R_ExternalPtrAddr(ptr);
  if (!h) Rf_error("gpsurv: handle has been destroyed");
  return h;
}

static void handle_finalizer(SEXP ptr) {
  gps_handle h = (gps_handle)R_ExternalPtrAddr(ptr);
  if (h) { gps_destroy(h); R_ClearExternalPtr(ptr); }
}

static void check_status(int rc, gps_handle h) {
  if (rc == GPS_OK) return;
  if (rc == GPS_ERR_NOT_PD) Rf_error("gps_not_pd: %s", gps_last_error(h));
  Rf_error("gpsurv error %d: %s", rc, gps_last_error(h));
}

/* .Call('gps_r_create', device) -> external pointer */
SEXP gps_r_create(SEXP device) {
  gps_handle h = NULL;
  int rc = gps_create(Rf_asInteger(device), &h);
  if (rc != GPS_OK) Rf_error("gpsurv error %d: %s", rc, gps_last_error(NULL));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* The shapes R hands over are checked HERE (a mismatched testData or an integer matrix must raise an R error, not read past
 * the end of a buffer): every numeric input must be REALSXP, lengths must match the resident data (gps_get_shape). */
static void need_real(SEXP x, const char* what) {
  if (!Rf_isReal(x)) Rf_error("gpsurv: %s must be a numeric (double) vector or matrix", what);
}
static void resident_shape(gps_handle h, int* n, int* d, int* ncens) {
  int batch = 0;
  if (gps_get_shape(h, n, d, &batch, ncens) != GPS_OK || *n == 0) Rf_error("gpsurv: no training data bound (gps_r_set_data first)");
}

/* covFuncForm: 0 SqExp, 1 ARD, 2/3/4 Matern with extraParam 1/3/5 (CovFunc.R:31-37); meanFuncForm: 0 Zero, 1 Linear;
 * noiseCorr: 0 none, 1 Learned, 2 Vec; ard_mode: 0 intended, 1 as written */
SEXP gps_r_set_options(SEXP ptr, SEXP cov, SEXP mean, SEXP nc, SEXP ard_mode) {
  gps_handle h = get_handle(ptr);
  check_status(gps_set_options(h, Rf_asInteger(cov), Rf_asInteger(mean), Rf_asInteger(nc), Rf_asInteger(ard_mode)), h);
  return R_NilValue;
}

/* trainingData (n x d numeric matrix), trainingTargets (n), events (n, 0/1 numeric) */
SEXP gps_r_set_data(SEXP ptr, SEXP X, SEXP y, SEXP events) {
  gps_handle h = get_handle(ptr);
  if (!Rf_isReal(X) || !Rf_isMatrix(X) || !Rf_isReal(y) || !Rf_isReal(events)) Rf_error("gpsurv: numeric inputs expected");
  int n = Rf_nrows(X), d = Rf_ncols(X);
  if (XLENGTH(y) != n || XLENGTH(events) != n) Rf_error("gpsurv: targets/events length mismatch");
  check_status(gps_set_data(h, REAL(X), n, d, REAL(y), REAL(events)), h);
  return R_NilValue;
}

SEXP gps_r_set_targets(SEXP ptr, SEXP y) {
  gps_handle h = get_handle(ptr);
  int n, d, nc;
  resident_shape(h, &n, &d, &nc);
  need_real(y, "trainingTargets");
  if (XLENGTH(y) != n) Rf_error("gpsurv: trainingTargets has length %d, the bound data has %d rows", (int)XLENGTH(y), n);
  check_status(gps_set_targets(h, REAL(y)), h);
  return R_NilValue;
}

SEXP gps_r_set_noise_corr(SEXP ptr, SEXP log_sc2) {
  gps_handle h = get_handle(ptr);
  need_real(log_sc2, "logHypNoiseCorrection");
  check_status(gps_set_noise_corr(h, REAL(log_sc2), (int)XLENGTH(log_sc2)), h);
  return R_NilValue;
}

/* CovFunc: x2 may be NULL (identical(x1, x2)) */
SEXP gps_r_cov(SEXP ptr, SEXP x1, SEXP x2, SEXP sf2, SEXP ell, SEXP cov, SEXP ard_mode) {
  gps_handle h = Rf_isNull(ptr) ? NULL : get_handle(ptr);      /* CovFunc needs no fit: NULL = the library's default handle */
  need_real(x1, "x1");
  need_real(ell, "lengthHyp");
  if (!Rf_isMatrix(x1)) Rf_error("gpsurv: x1 must be a matrix");
  int n1 = Rf_nrows(x1), d = Rf_ncols(x1);
  if (!Rf_isNull(x2)) {
    need_real(x2, "x2");
    if (!Rf_isMatrix(x2) || Rf_ncols(x2) != d) Rf_error("gpsurv: x2 must be a matrix with the %d columns of x1", d);
  }
  if (XLENGTH(ell) != 1 && XLENGTH(ell) != d && Rf_asInteger(ard_mode) == 0) Rf_error("gpsurv: lengthHyp needs 1 or %d entries", d);
  int n2 = Rf_isNull(x2) ? n1 : Rf_nrows(x2);
  SEXP K = PROTECT(Rf_allocMatrix(REALSXP, n1, n2));
  int rc = gps_cov(h, REAL(x1), n1, Rf_isNull(x2) ? NULL : REAL(x2), n2, d, Rf_asReal(sf2), REAL(ell),
                   (int)XLENGTH(ell), Rf_asInteger(cov), Rf_asInteger(ard_mode), REAL(K));
  UNPROTECT(1);
  check_status(rc, h);
  return K;
}

/* ForOptimisationLog -> 1 x 1 matrix (ForOptimisationLog.R:67) */
SEXP gps_r_nlml(SEXP ptr, SEXP par) {
  gps_handle h = get_handle(ptr);
  need_real(par, "logHyp");
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, 1, 1));
  int rc = gps_nlml(h, REAL(par), (int)XLENGTH(par), REAL(out));
  UNPROTECT(1);
  check_status(rc, h);
  return out;
}

/* OptimisationGradientLog -> k x 1 matrix (OptimisationGradientLog.R:93) */
SEXP gps_r_nlml_grad(SEXP ptr, SEXP par) {
  gps_handle h = get_handle(ptr);
  need_real(par, "logHyp");
  int len = (int)XLENGTH(par);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, len, 1));
  int rc = gps_nlml_grad(h, REAL(par), len, NULL, REAL(out));
  UNPROTECT(1);
  check_status(rc, h);
  return out;
}

/* GPRegression.R:60-82 -> list(K, L, alpha, meanTraining, logMarLik, gpsToken); want_matrices = FALSE skips K and L.
 * gpsToken identifies the model now resident in the handle (gps_model_token): GPPredict checks it. */
SEXP gps_r_finalize(SEXP ptr, SEXP par, SEXP n_, SEXP want_matrices) {
  gps_handle h = get_handle(ptr);
  int n = Rf_asInteger(n_), want = Rf_asLogical(want_matrices);
  int nres, d, nc;
  resident_shape(h, &nres, &d, &nc);
  need_real(par, "logHyp");
  if (n != nres) Rf_error("gpsurv: nTraining = %d but the bound data has %d rows", n, nres);
  SEXP K = PROTECT(want ? Rf_allocMatrix(REALSXP, n, n) : R_NilValue);
  SEXP L = PROTECT(want ? Rf_allocMatrix(REALSXP, n, n) : R_NilValue);
  SEXP alpha = PROTECT(Rf_allocMatrix(REALSXP, n, 1));
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, n, 1));
  SEXP lml = PROTECT(Rf_allocMatrix(REALSXP, 1, 1));
  int rc = gps_regress_finalize(h, REAL(par), (int)XLENGTH(par), want ? REAL(K) : NULL, want ? REAL(L) : NULL,
                                REAL(alpha), REAL(mean), REAL(lml));
  SEXP tok = PROTECT(Rf_allocVector(REALSXP, 1));
  REAL(tok)[0] = (double)gps_model_token(h);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 6));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, 6));
  const char* nm[6] = {"K", "L", "alpha", "meanTraining", "logMarLik", "gpsToken"};
  SEXP vals[6] = {K, L, alpha, mean, lml, tok};
  for (int i = 0; i < 6; i++) { SET_VECTOR_ELT(out, i, vals[i]); SET_STRING_ELT(names, i, Rf_mkChar(nm[i])); }
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(8);
  check_status(rc, h);
  return out;
}

/* GPPredict.R:49-56 -> list(funcMeanPred (m x 1), varFunction, varData) */
SEXP gps_r_predict(SEXP ptr, SEXP par, SEXP reuse_model, SEXP Xstar) {
  gps_handle h = get_handle(ptr);
  int n, d, nc;
  resident_shape(h, &n, &d, &nc);
  need_real(par, "logHyp");
  need_real(Xstar, "testData");
  if (!Rf_isMatrix(Xstar) || Rf_ncols(Xstar) != d) Rf_error("gpsurv: testData must be a matrix with the %d columns of trainingData", d);
  int m = Rf_nrows(Xstar);
  SEXP mu = PROTECT(Rf_allocMatrix(REALSXP, m, 1));
  SEXP vf = PROTECT(Rf_allocVector(REALSXP, m));
  SEXP vd = PROTECT(Rf_allocVector(REALSXP, m));
  int rc = gps_predict(h, REAL(par), (int)XLENGTH(par), Rf_asLogical(reuse_model), REAL(Xstar), m, REAL(mu), REAL(vf),
                       REAL(vd));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, 3));
  SET_VECTOR_ELT(out, 0, mu); SET_VECTOR_ELT(out, 1, vf); SET_VECTOR_ELT(out, 2, vd);
  SET_STRING_ELT(names, 0, Rf_mkChar("funcMeanPred")); SET_STRING_ELT(names, 1, Rf_mkChar("varFunction"));
  SET_STRING_ELT(names, 2, Rf_mkChar("varData"));
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(5);
  check_status(rc, h);
  return out;
}

/* AdjustTrainingSurvivalMeanVariance, vectorised over the censored rows -> list(mu, sigma) */
SEXP gps_r_adjust(SEXP ptr, SEXP a, SEXP mu, SEXP s) {
  gps_handle h = Rf_isNull(ptr) ? NULL : get_handle(ptr);
  need_real(a, "a");
  need_real(mu, "mu");
  need_real(s, "sigma");
  int c = (int)XLENGTH(a);
  if (XLENGTH(mu) != c || XLENGTH(s) != c) Rf_error("gpsurv: a, mu and sigma must have the same length");
  SEXP em = PROTECT(Rf_allocVector(REALSXP, c));
  SEXP ev = PROTECT(Rf_allocVector(REALSXP, c));
  int rc = gps_adjust(h, REAL(a), REAL(mu), REAL(s), c, REAL(em), REAL(ev));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(out, 0, em); SET_VECTOR_ELT(out, 1, ev);
  SET_STRING_ELT(names, 0, Rf_mkChar("mu")); SET_STRING_ELT(names, 1, Rf_mkChar("sigma"));
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(4);
  check_status(rc, h);
  return out;
}

/* token of the model resident in the handle (0 = none) */
SEXP gps_r_model_token(SEXP ptr) {
  gps_handle h = get_handle(ptr);
  SEXP tok = PROTECT(Rf_allocVector(REALSXP, 1));
  REAL(tok)[0] = (double)gps_model_token(h);
  UNPROTECT(1);
  return tok;
}

/* nearPD repair of non-positive-definite covariance matrices on the device (default on) */
SEXP gps_r_set_near_pd(SEXP ptr, SEXP enable) {
  gps_handle h = get_handle(ptr);
  check_status(gps_set_near_pd(h, Rf_asLogical(enable)), h);
  return R_NilValue;
}

/* CalculateMetrics.R:18,21 -> list(c.index, rmse) for one set of test predictions */
SEXP gps_r_metrics(SEXP ptr, SEXP pred, SEXP time, SEXP event) {
  gps_handle h = get_handle(ptr);
  need_real(pred, "testPredictions");
  need_real(time, "testTargets");
  need_real(event, "testEvents");
  int m = (int)XLENGTH(pred);
  if (XLENGTH(time) != m || XLENGTH(event) != m) Rf_error("gpsurv: predictions, targets and events must have the same length");
  SEXP ci = PROTECT(Rf_allocVector(REALSXP, 1));
  SEXP rm = PROTECT(Rf_allocVector(REALSXP, 1));
  int rc = gps_metrics(h, REAL(pred), REAL(time), REAL(event), m, REAL(ci), REAL(rm), NULL);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(out, 0, ci); SET_VECTOR_ELT(out, 1, rm);
  SET_STRING_ELT(names, 0, Rf_mkChar("c.index")); SET_STRING_ELT(names, 1, Rf_mkChar("rmse"));
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(4);
  check_status(rc, h);
  return out;
}

/* The restart x fold loops of runSyntheticExp1.R:208-213 over every GPU of the box from ONE R process: `total` GPS fits as
 * row-index views of one pooled data set (gps_fit_folds_multi).  devices: integer vector; trainRows / testRows: integer
 * matrices n x total / m x total of 1-based pool rows (R indexing; shifted here); opts = c(cov, mean, noiseCorr, ardMode);
 * ctrl = c(method, maxit, maxCount, survival); parStart: len x total.  Returns list(par, objective, logMarLik, count,
 * funcMeanPred (m x total), varData, trainingSetPredictions (n x total), status). */
SEXP gps_r_fit_folds_multi(SEXP devices, SEXP Xpool, SEXP ypool, SEXP evpool, SEXP trainRows, SEXP testRows, SEXP parStart,
                           SEXP opts, SEXP ctrl, SEXP tolerance) {
  need_real(Xpool, "pooled data");
  need_real(ypool, "pooled log targets");
  need_real(evpool, "pooled events");
  need_real(parStart, "start hyper-parameters");
  if (!Rf_isMatrix(Xpool) || !Rf_isMatrix(trainRows) || !Rf_isMatrix(testRows) || !Rf_isMatrix(parStart))
    Rf_error("gpsurv: matrices expected");
  if (TYPEOF(devices) != INTSXP || TYPEOF(trainRows) != INTSXP || TYPEOF(testRows) != INTSXP || TYPEOF(opts) != INTSXP ||
      TYPEOF(ctrl) != INTSXP || LENGTH(opts) != 4 || LENGTH(ctrl) != 4)
    Rf_error("gpsurv: integer devices / row indices / opts(4) / ctrl(4) expected");
  int npool = Rf_nrows(Xpool), d = Rf_ncols(Xpool), n = Rf_nrows(trainRows), total = Rf_ncols(trainRows), m = Rf_nrows(testRows),
      len = Rf_nrows(parStart);
  if (XLENGTH(ypool) != npool || XLENGTH(evpool) != npool || Rf_ncols(testRows) != total || Rf_ncols(parStart) != total)
    Rf_error("gpsurv: inconsistent sizes");
  int* tr = (int*)R_alloc((size_t)n * total, sizeof(int));
  int* te = (int*)R_alloc((size_t)m * total, sizeof(int));
  for (R_xlen_t i = 0; i < (R_xlen_t)n * total; i++) tr[i] = INTEGER(trainRows)[i] - 1;
  for (R_xlen_t i = 0; i < (R_xlen_t)m * total; i++) te[i] = INTEGER(testRows)[i] - 1;
  SEXP par = PROTECT(Rf_allocMatrix(REALSXP, len, total));
  SEXP obj = PROTECT(Rf_allocVector(REALSXP, total));
  SEXP lml = PROTECT(Rf_allocVector(REALSXP, total));
  SEXP cnt = PROTECT(Rf_allocVector(INTSXP, total));
  SEXP tm = PROTECT(Rf_allocMatrix(REALSXP, m, total));
  SEXP tvd = PROTECT(Rf_allocMatrix(REALSXP, m, total));
  SEXP trm = PROTECT(Rf_allocMatrix(REALSXP, n, total));
  SEXP st = PROTECT(Rf_allocVector(INTSXP, total));
  const int* o = INTEGER(opts);
  const int* c = INTEGER(ctrl);
  int rc = gps_fit_folds_multi(INTEGER(devices), LENGTH(devices), total, o[0], o[1], o[2], o[3], REAL(Xpool), npool, d, REAL(ypool),
                               REAL(evpool), tr, n, te, m, REAL(parStart), len, c[0], c[1], Rf_asReal(tolerance), c[2], c[3], REAL(par),
                               REAL(obj), REAL(lml), INTEGER(cnt), NULL, REAL(tm), NULL, REAL(tvd), REAL(trm), INTEGER(st), NULL);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 8));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, 8));
  const char* nm[8] = {"par", "objective", "logMarLik", "count", "funcMeanPred", "varData", "trainingSetPredictions", "status"};
  SEXP vals[8] = {par, obj, lml, cnt, tm, tvd, trm, st};
  for (int i = 0; i < 8; i++) { SET_VECTOR_ELT(out, i, vals[i]); SET_STRING_ELT(names, i, Rf_mkChar(nm[i])); }
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(10);
  check_status(rc, NULL);
  return out;
}

/* Free the idle handles gps_fit_folds_multi keeps between calls (one per device): device memory back to the session. */
SEXP gps_r_multi_release(void) {
  check_status(gps_multi_release(), NULL);
  return R_NilValue;
}

static const R_CallMethodDef call_methods[] = {
  {"gps_r_create", (DL_FUNC)&gps_r_create, 1},
  {"gps_r_set_options", (DL_FUNC)&gps_r_set_options, 5},
  {"gps_r_set_data", (DL_FUNC)&gps_r_set_data, 4},
  {"gps_r_set_targets", (DL_FUNC)&gps_r_set_targets, 2},
  {"gps_r_set_noise_corr", (DL_FUNC)&gps_r_set_noise_corr, 2},
  {"gps_r_cov", (DL_FUNC)&gps_r_cov, 7},
  {"gps_r_nlml", (DL_FUNC)&gps_r_nlml, 2},
  {"gps_r_nlml_grad", (DL_FUNC)&gps_r_nlml_grad, 2},
  {"gps_r_finalize", (DL_FUNC)&gps_r_finalize, 4},
  {"gps_r_predict", (DL_FUNC)&gps_r_predict, 4},
  {"gps_r_adjust", (DL_FUNC)&gps_r_adjust, 4},
  {"gps_r_model_token", (DL_FUNC)&gps_r_model_token, 1},
  {"gps_r_set_near_pd", (DL_FUNC)&gps_r_set_near_pd, 2},
  {"gps_r_metrics", (DL_FUNC)&gps_r_metrics, 4},
  {"gps_r_fit_folds_multi", (DL_FUNC)&gps_r_fit_folds_multi, 10},
  {"gps_r_multi_release", (DL_FUNC)&gps_r_multi_release, 0},
  {NULL, NULL, 0}
};

void R_init_gpsurv_glue(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
