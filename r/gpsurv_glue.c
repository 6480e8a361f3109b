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

/* covFuncForm: 0 SqExp, 1 ARD; meanFuncForm: 0 Zero, 1 Linear; noiseCorr: 0 none, 1 Learned, 2 Vec */
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
  check_status(gps_set_targets(h, REAL(y)), h);
  return R_NilValue;
}

SEXP gps_r_set_noise_corr(SEXP ptr, SEXP log_sc2) {
  gps_handle h = get_handle(ptr);
  check_status(gps_set_noise_corr(h, REAL(log_sc2), (int)XLENGTH(log_sc2)), h);
  return R_NilValue;
}

/* CovFunc: x2 may be NULL (identical(x1, x2)) */
SEXP gps_r_cov(SEXP ptr, SEXP x1, SEXP x2, SEXP sf2, SEXP ell, SEXP cov, SEXP ard_mode) {
  gps_handle h = get_handle(ptr);
  int n1 = Rf_nrows(x1), d = Rf_ncols(x1);
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
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, 1, 1));
  int rc = gps_nlml(h, REAL(par), (int)XLENGTH(par), REAL(out));
  UNPROTECT(1);
  check_status(rc, h);
  return out;
}

/* OptimisationGradientLog -> k x 1 matrix (OptimisationGradientLog.R:93) */
SEXP gps_r_nlml_grad(SEXP ptr, SEXP par) {
  gps_handle h = get_handle(ptr);
  int len = (int)XLENGTH(par);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, len, 1));
  int rc = gps_nlml_grad(h, REAL(par), len, NULL, REAL(out));
  UNPROTECT(1);
  check_status(rc, h);
  return out;
}

/* GPRegression.R:60-82 -> list(K, L, alpha, meanTraining, logMarLik); want_matrices = FALSE skips K and L */
SEXP gps_r_finalize(SEXP ptr, SEXP par, SEXP n_, SEXP want_matrices) {
  gps_handle h = get_handle(ptr);
  int n = Rf_asInteger(n_), want = Rf_asLogical(want_matrices);
  SEXP K = PROTECT(want ? Rf_allocMatrix(REALSXP, n, n) : R_NilValue);
  SEXP L = PROTECT(want ? Rf_allocMatrix(REALSXP, n, n) : R_NilValue);
  SEXP alpha = PROTECT(Rf_allocMatrix(REALSXP, n, 1));
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, n, 1));
  SEXP lml = PROTECT(Rf_allocMatrix(REALSXP, 1, 1));
  int rc = gps_regress_finalize(h, REAL(par), (int)XLENGTH(par), want ? REAL(K) : NULL, want ? REAL(L) : NULL,
                                REAL(alpha), REAL(mean), REAL(lml));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, 5));
  const char* nm[5] = {"K", "L", "alpha", "meanTraining", "logMarLik"};
  SEXP vals[5] = {K, L, alpha, mean, lml};
  for (int i = 0; i < 5; i++) { SET_VECTOR_ELT(out, i, vals[i]); SET_STRING_ELT(names, i, Rf_mkChar(nm[i])); }
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(7);
  check_status(rc, h);
  return out;
}

/* GPPredict.R:49-56 -> list(funcMeanPred (m x 1), varFunction, varData) */
SEXP gps_r_predict(SEXP ptr, SEXP par, SEXP reuse_model, SEXP Xstar) {
  gps_handle h = get_handle(ptr);
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
  gps_handle h = get_handle(ptr);
  int c = (int)XLENGTH(a);
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
  {NULL, NULL, 0}
};

void R_init_gpsurv_glue(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
