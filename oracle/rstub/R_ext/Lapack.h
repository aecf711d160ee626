/* oracle/rstub/R_ext/Lapack.h -- TEST INFRASTRUCTURE (see R.h).  dpotri_ is
 * provided by oracle/rminc_oracle.c's LAPACK dtrti2+dlauu2 restatement. */
#ifndef RSTUB_LAPACK_H
#define RSTUB_LAPACK_H
#define FCONE
#ifdef __cplusplus
extern "C" {
#endif
void dpotri_(const char *uplo, const int *n, double *a, const int *lda, int *info);
#ifdef __cplusplus
}
#endif
#endif
