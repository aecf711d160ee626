/* oracle/rstub/R.h -- TEST INFRASTRUCTURE.  Minimal stand-in for R's C API so
 * that the reference's own src/modelling_functions.c and src/vertex_modelling.c
 * can be compiled where they lie (R itself is absent from this image).  Only
 * the entry points those two files use are provided; see rstub.c. */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifdef __cplusplus
extern "C" {
#endif
void Rprintf(const char *fmt, ...);
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
#define error Rf_error
#define F77_CALL(x) x##_
#define F77_NAME(x) x##_
#ifdef __cplusplus
}
#endif
#endif
