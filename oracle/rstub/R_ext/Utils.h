/* oracle/rstub/R_ext/Utils.h -- TEST INFRASTRUCTURE (see R.h). */
#ifndef RSTUB_UTILS_H
#define RSTUB_UTILS_H
#ifdef __cplusplus
extern "C" {
#endif
void rsort_with_index(double *x, int *indx, int n);
void R_CheckUserInterrupt(void);
#ifdef __cplusplus
}
#endif
#endif
