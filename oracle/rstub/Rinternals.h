/* oracle/rstub/Rinternals.h -- TEST INFRASTRUCTURE (see R.h). */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include "R.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct rstub_sexp *SEXP;
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define CHARSXP 9
#define RAWSXP 24
#define VECSXP 19
#define EXTPTRSXP 22
#define SYMSXP 1
#include <limits.h>
#define NA_INTEGER INT_MIN
#define NA_LOGICAL INT_MIN
typedef int Rboolean;
typedef ptrdiff_t R_xlen_t;
extern SEXP R_NilValue;
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_protect(SEXP s);
void Rf_unprotect(int n);
double *REAL(SEXP s);
int *INTEGER(SEXP s);
int *LOGICAL(SEXP s);
int LENGTH(SEXP s);
int Rf_length(SEXP s);
int Rf_nrows(SEXP s);
int Rf_ncols(SEXP s);
int Rf_isLogical(SEXP s);
int Rf_isNumeric(SEXP s);
int Rf_isNull(SEXP s);
SEXP STRING_ELT(SEXP s, R_xlen_t i);
const char *CHAR(SEXP s);
SEXP Rf_install(const char *name);
int Rf_isReal(SEXP s);
int Rf_isMatrix(SEXP s);
int Rf_isInteger(SEXP s);
int Rf_asInteger(SEXP s);
int Rf_asLogical(SEXP s);
double Rf_asReal(SEXP s);
SEXP Rf_GetOption1(SEXP tag);
R_xlen_t XLENGTH(SEXP s);
int TYPEOF(SEXP s);
unsigned char *RAW(SEXP s);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
char *R_alloc(size_t n, int size);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, void (*fun)(SEXP), Rboolean onexit);
#ifndef FALSE
#define FALSE 0
#define TRUE 1
#endif
SEXP Rf_lang2(SEXP a, SEXP b);
SEXP Rf_eval(SEXP expr, SEXP rho);
void Rf_defineVar(SEXP sym, SEXP val, SEXP rho);
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define length(s) Rf_length(s)
#define nrows Rf_nrows
#define ncols Rf_ncols
#define isLogical Rf_isLogical
#define isNumeric Rf_isNumeric
#define isNull Rf_isNull
#define install Rf_install
#define lang2 Rf_lang2
#define eval Rf_eval
#define defineVar Rf_defineVar
#ifdef __cplusplus
}
#endif
#endif
