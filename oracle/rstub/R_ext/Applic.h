/* oracle/rstub/R_ext/Applic.h -- TEST INFRASTRUCTURE (see R.h).  dqrls_ is
 * provided by oracle/rminc_oracle.c's restatement of R's src/appl/dqrls.f. */
#ifndef RSTUB_APPLIC_H
#define RSTUB_APPLIC_H
#ifdef __cplusplus
extern "C" {
#endif
void dqrls_(double *x, int *n, int *p, double *y, int *ny, double *tol,
            double *b, double *rsd, double *qty, int *k, int *jpvt,
            double *qraux, double *work);
#ifdef __cplusplus
}
#endif
#endif
