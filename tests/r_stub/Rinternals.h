/* TEST INFRASTRUCTURE — a minimal stand-in for R's C API (R is not installed in the build image).
 * Only the declarations r/gpsurv_glue.c uses, with the signatures of R >= 3.x (Rinternals.h), so that the glue can be
 * type-checked against include/gpsurv.h (`gcc -fsyntax-only`).  Nothing here is ever linked or run. */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
extern SEXP R_NilValue;
extern SEXP R_NamesSymbol;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
R_xlen_t XLENGTH(SEXP);
int LENGTH(SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isMatrix(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int TYPEOF(SEXP);
char* R_alloc(size_t, int);
SEXP Rf_mkChar(const char*);
SEXP Rf_mkString(const char*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
void Rf_warning(const char*, ...);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean onexit);
#endif
