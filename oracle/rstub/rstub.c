/*
 * oracle/rstub/rstub.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Just enough of R's C API and of libminc2 to run the reference's own
 * src/modelling_functions.c (voxel_lm, minc2_model) and src/vertex_modelling.c
 * (vertex_lm_loop), compiled from /root/reference where they lie, on in-memory
 * data.  Plus the plain-C harness entry points (ref_*) that tests call through
 * ctypes, and (last section) the rest of the R API that the product's .Call shim
 * (rminc_b200/csrc/rshim.c) uses together with a .Call dispatcher over the routine
 * tables R_registerRoutines receives, so that tests can EXECUTE the shim: mini-SEXPs
 * in, SEXP out, Rf_error -> longjmp, R's collect-the-unprotected contract included.
 * -DRSTUB_NO_REFERENCE leaves out everything that needs the reference's sources.
 * LINPACK dqrls_ and LAPACK dpotri_ (third-party, not in the
 * reference tree) are forwarded to oracle/rminc_oracle.c's restatements.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include "Rinternals.h"
#include "R_ext/Utils.h"
#include "R_ext/Rdynload.h"
#include "minc2.h"

/* ------------------------------------------------------------ mini SEXPs */
struct rstub_sexp {
    SEXPTYPE type;
    R_xlen_t len;
    int nrow, ncol; /* ncol < 0: no dim attribute */
    int protect_count;
    int permanent;
    void *data;
    struct rstub_sexp *next_zero; /* list of objects with protect_count == 0 */
    int in_zero;
    void (*finalizer)(struct rstub_sexp *); /* external pointers */
};

static struct rstub_sexp nil_obj = {NILSXP, 0, 0, -1, 0, 1, NULL, NULL, 0, NULL};
SEXP R_NilValue = &nil_obj;

static SEXP zero_list = NULL;
#define PSTACK_MAX 10000
static SEXP pstack[PSTACK_MAX];
static int pstack_top = 0;
static jmp_buf err_jmp;
static int err_armed = 0;
static char err_msg[512];
static int alloc_fill_on = 0;      /* model "whatever a fresh R allocation holds" */
static double alloc_fill_value = 0.0;

static void zero_push(SEXP s)
{
    if (s->in_zero || s->permanent) return;
    s->in_zero = 1;
    s->next_zero = zero_list;
    zero_list = s;
}

/* R may collect any unprotected object at the next allocation; so do we. */
static void sweep(void)
{
    SEXP s = zero_list;
    zero_list = NULL;
    while (s) {
        SEXP nx = s->next_zero;
        s->in_zero = 0;
        if (s->protect_count == 0 && !s->permanent) {
            if (s->finalizer) s->finalizer(s);          /* R runs a C finalizer when it collects the pointer */
            if (s->type != EXTPTRSXP) free(s->data);
            free(s);
        }
        s = nx;
    }
}

static size_t elt_size(SEXPTYPE t)
{
    switch (t) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case STRSXP: case VECSXP: return sizeof(SEXP);
    case CHARSXP: return 1;
    default: return 1;
    }
}

static SEXP alloc_obj(SEXPTYPE type, R_xlen_t n, int permanent)
{
    sweep();
    SEXP s = (SEXP)calloc(1, sizeof(struct rstub_sexp));
    s->type = type;
    s->len = n;
    s->nrow = (int)n;
    s->ncol = -1;
    s->permanent = permanent;
    s->data = malloc((size_t)(n > 0 ? n : 1) * elt_size(type) + (type == CHARSXP ? 1 : 0));
    if (alloc_fill_on && type == REALSXP)
        for (R_xlen_t i = 0; i < n; ++i) ((double *)s->data)[i] = alloc_fill_value;
    zero_push(s);
    return s;
}

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n) { return alloc_obj(type, n, 0); }

SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol)
{
    SEXP s = alloc_obj(type, (R_xlen_t)nrow * ncol, 0);
    s->nrow = nrow;
    s->ncol = ncol;
    return s;
}

SEXP Rf_protect(SEXP s)
{
    if (pstack_top >= PSTACK_MAX) Rf_error("protect(): protection stack overflow");
    pstack[pstack_top++] = s;
    s->protect_count++;
    return s;
}

void Rf_unprotect(int n)
{
    while (n-- > 0 && pstack_top > 0) {
        SEXP s = pstack[--pstack_top];
        if (--s->protect_count == 0) zero_push(s);
    }
}

double *REAL(SEXP s) { return (double *)s->data; }
int *INTEGER(SEXP s) { return (int *)s->data; }
int *LOGICAL(SEXP s) { return (int *)s->data; }
int LENGTH(SEXP s) { return (int)s->len; }
int Rf_length(SEXP s) { return (int)s->len; }
int Rf_nrows(SEXP s) { return s->ncol < 0 ? (int)s->len : s->nrow; }
int Rf_ncols(SEXP s) { return s->ncol < 0 ? 1 : s->ncol; }
int Rf_isLogical(SEXP s) { return s->type == LGLSXP; }
int Rf_isNumeric(SEXP s) { return s->type == REALSXP || s->type == INTSXP || s->type == LGLSXP; }
int Rf_isNull(SEXP s) { return s->type == NILSXP; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
const char *CHAR(SEXP s) { return (const char *)s->data; }

SEXP Rf_lang2(SEXP a, SEXP b) { (void)a; (void)b; return R_NilValue; }
SEXP Rf_eval(SEXP e, SEXP rho) { (void)e; (void)rho; Rf_error("rstub: eval() is not available"); }
void Rf_defineVar(SEXP a, SEXP b, SEXP c) { (void)a; (void)b; (void)c; }
void R_CheckUserInterrupt(void) {}

void Rprintf(const char *fmt, ...) { (void)fmt; } /* the reference's tests sink this output too */

void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof err_msg, fmt, ap);
    va_end(ap);
    if (err_armed) longjmp(err_jmp, 1);
    fprintf(stderr, "rstub: uncaught R error: %s\n", err_msg);
    abort();
}

void rsort_with_index(double *x, int *indx, int n)
{ /* insertion sort; only the (out-of-scope) wilcoxon method uses it */
    for (int i = 1; i < n; ++i) {
        double v = x[i]; int iv = indx[i]; int j = i - 1;
        while (j >= 0 && x[j] > v) { x[j + 1] = x[j]; indx[j + 1] = indx[j]; --j; }
        x[j + 1] = v; indx[j + 1] = iv;
    }
}

/* -------------------------------------------- third-party numerical kernels */
void orc_dqrls(double *x, int n, int p, const double *y, int ny, double tol,
               double *b, double *rsd, double *qty, int *k, int *jpvt,
               double *qraux, double *work);
int orc_dpotri_upper(int p, double *a, int lda);

void dqrls_(double *x, int *n, int *p, double *y, int *ny, double *tol, double *b,
            double *rsd, double *qty, int *k, int *jpvt, double *qraux, double *work)
{
    /* R's dqrls takes and returns a 1-based jpvt; voxel_lm passes 0-based
     * values (pivot[j] = j) and reads them back as 0-based, which works because
     * dqrdc2 only permutes the entries. */
    orc_dqrls(x, *n, *p, y, *ny, *tol, b, rsd, qty, k, jpvt, qraux, work);
}

void dpotri_(const char *uplo, const int *n, double *a, const int *lda, int *info)
{
    if (uplo[0] != 'U' && uplo[0] != 'u') { *info = -1; return; }
    *info = orc_dpotri_upper(*n, a, *lda);
}

/* ------------------------------------------------- in-memory "MINC volumes" */
struct rstub_volume {
    char name[256];
    const double *data;
    misize_t sizes[3];
};
struct rstub_dim { misize_t size; };
#define MAX_VOLS 4096
static struct rstub_volume vols[MAX_VOLS];
static struct rstub_dim vol_dims[MAX_VOLS][3];
static int n_vols = 0;

void rstub_clear_volumes(void) { n_vols = 0; }

int rstub_register_volume(const char *name, const double *data, const int64_t sizes[3])
{
    if (n_vols >= MAX_VOLS) return -1;
    struct rstub_volume *v = &vols[n_vols];
    snprintf(v->name, sizeof v->name, "%s", name);
    v->data = data;
    for (int d = 0; d < 3; ++d) { v->sizes[d] = (misize_t)sizes[d]; vol_dims[n_vols][d].size = v->sizes[d]; }
    return n_vols++;
}

int miopen_volume(const char *filename, int mode, mihandle_t *volume)
{
    (void)mode;
    for (int i = 0; i < n_vols; ++i)
        if (strcmp(vols[i].name, filename) == 0) { *volume = &vols[i]; return MI_NOERROR; }
    return MI_ERROR;
}
int miclose_volume(mihandle_t volume) { (void)volume; return MI_NOERROR; }
int miget_volume_dimensions(mihandle_t volume, int cls, int attr, int order,
                            int array_length, midimhandle_t dimensions[])
{
    (void)cls; (void)attr; (void)order;
    int idx = (int)(volume - vols);
    for (int d = 0; d < array_length && d < 3; ++d) dimensions[d] = &vol_dims[idx][d];
    return 3;
}
int miget_dimension_sizes(const midimhandle_t dimensions[], misize_t array_length, misize_t sizes[])
{
    for (misize_t d = 0; d < array_length; ++d) sizes[d] = dimensions[d]->size;
    return MI_NOERROR;
}
int miget_real_value_hyperslab(mihandle_t volume, int buffer_data_type, const misize_t start[],
                               const misize_t count[], void *buffer)
{
    (void)buffer_data_type;
    double *out = (double *)buffer;
    const misize_t *s = volume->sizes;
    size_t o = 0;
    for (misize_t a = 0; a < count[0]; ++a)
        for (misize_t b = 0; b < count[1]; ++b)
            for (misize_t c = 0; c < count[2]; ++c)
                out[o++] = volume->data[((start[0] + a) * s[1] + (start[1] + b)) * s[2] + (start[2] + c)];
    return MI_NOERROR;
}

/* ------------------------------------------------------- objects owned by the harness */
static SEXP perm_real(const double *src, R_xlen_t n, int nrow, int ncol)
{
    SEXP s = alloc_obj(REALSXP, n, 1);
    if (ncol >= 0) { s->nrow = nrow; s->ncol = ncol; }
    if (src) memcpy(s->data, src, (size_t)n * sizeof(double));
    return s;
}
static SEXP perm_logical_na_matrix(void)
{ /* `matrix()` in R: a 1 x 1 logical NA */
    SEXP s = alloc_obj(LGLSXP, 1, 1);
    s->nrow = 1; s->ncol = 1;
    ((int *)s->data)[0] = INT32_MIN;
    return s;
}
static SEXP perm_string(const char *const *strs, int n)
{
    SEXP s = alloc_obj(STRSXP, n, 1);
    for (int i = 0; i < n; ++i) {
        size_t L = strlen(strs[i]);
        SEXP c = alloc_obj(CHARSXP, (R_xlen_t)L, 1);
        memcpy(c->data, strs[i], L + 1);
        ((SEXP *)s->data)[i] = c;
    }
    return s;
}
static void perm_free(SEXP s)
{
    if (!s || s == R_NilValue) return;
    if (s->type == STRSXP)
        for (R_xlen_t i = 0; i < s->len; ++i) perm_free(((SEXP *)s->data)[i]);
    free(s->data);
    free(s);
}
static void reset_after_call(void)
{
    while (pstack_top > 0) { SEXP s = pstack[--pstack_top]; if (--s->protect_count == 0) zero_push(s); }
    sweep();
}

#ifndef RSTUB_NO_REFERENCE
/* ------------------------------------------------------- harness entry points */
SEXP voxel_lm(SEXP Sy, SEXP Sx, int n, int p, double *coefficients, double *residuals,
              double *effects, double *work, double *qraux, double *v, int *pivot,
              double *se, double *t);
SEXP vertex_lm_loop(SEXP data_left, SEXP data_right, SEXP mmatrix);
SEXP vertex_anova_loop(SEXP data, SEXP Sx, SEXP asgn);
SEXP vertex_wlm_loop(SEXP data_left, SEXP data_right, SEXP mmatrix, SEXP ws);
SEXP per_voxel_anova(SEXP filenames, SEXP Sx, SEXP asgn, SEXP have_mask, SEXP mask, SEXP mask_lower_value,
                     SEXP mask_upper_value);
SEXP minc2_model(SEXP filenames, SEXP filenames_right, SEXP mmatrix, SEXP asgn, SEXP have_mask,
                 SEXP mask, SEXP mask_lower_value, SEXP mask_upper_value, SEXP rho,
                 SEXP nresults, SEXP method);

/* one voxel through the reference's voxel_lm.  stats = [F, t.., R2, logLik]. */
int ref_voxel_lm(const double *y, const double *X, int n, int p, double *coef, double *stats,
                 int *pivot, char *err, size_t errlen)
{
    SEXP Sy = perm_real(y, n, 0, -1), Sx = perm_real(X, (R_xlen_t)n * p, n, p);
    double *resid = malloc(sizeof(double) * n), *eff = malloc(sizeof(double) * n);
    double *work = malloc(sizeof(double) * 2 * p), *qraux = malloc(sizeof(double) * p);
    double *v = malloc(sizeof(double) * p * p), *se = malloc(sizeof(double) * p), *t = malloc(sizeof(double) * p);
    int rc = 0;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        SEXP out = voxel_lm(Sy, Sx, n, p, coef, resid, eff, work, qraux, v, pivot, se, t);
        memcpy(stats, REAL(out), sizeof(double) * (p + 3));
    } else {
        rc = 1;
        if (err && errlen) snprintf(err, errlen, "%s", err_msg);
    }
    err_armed = 0;
    reset_after_call();
    perm_free(Sy); perm_free(Sx);
    free(resid); free(eff); free(work); free(qraux); free(v); free(se); free(t);
    return rc;
}

/* the reference's vertex_lm_loop on an in-memory V x n matrix (ld == V). */
int ref_vertex_lm_loop(const double *Y, int64_t V, int n, const double *X, int p, double *out,
                       char *err, size_t errlen)
{
    /* Y is borrowed, not copied: build the SEXP around the caller's buffer */
    SEXP left = alloc_obj(REALSXP, 0, 1);
    free(left->data);
    left->data = (void *)Y; left->len = (R_xlen_t)V * n; left->nrow = (int)V; left->ncol = n;
    SEXP right = perm_logical_na_matrix();
    SEXP mm = perm_real(X, (R_xlen_t)n * p, n, p);
    int rc = 0;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        SEXP res = vertex_lm_loop(left, right, mm);
        memcpy(out, REAL(res), sizeof(double) * (size_t)V * (2 * p + 3));
    } else {
        rc = 1;
        if (err && errlen) snprintf(err, errlen, "%s", err_msg);
    }
    err_armed = 0;
    reset_after_call();
    left->data = NULL; perm_free(left); perm_free(right); perm_free(mm);
    return rc;
}

/* the reference's minc2_model(method = "lm") over registered in-memory volumes.
 * Y is V x n column-major: column i is subject i's volume in file order.
 * mask may be NULL.  out must hold V*(2p+3) doubles and is pre-filled with
 * fill (the reference never writes columns 1.. of masked rows; whatever the
 * fresh R allocation held stays there -- pass 0.0 to model fresh zero pages, or
 * NaN to see exactly which cells the reference wrote). */
int ref_minc2_model_lm(const double *Y, const int64_t sizes[3], int n, const double *X, int p,
                       const double *mask, double lo, double hi, double *out, double fill,
                       char *err, size_t errlen)
{
    int64_t V = sizes[0] * sizes[1] * sizes[2];
    rstub_clear_volumes();
    char **names = malloc(sizeof(char *) * n);
    for (int i = 0; i < n; ++i) {
        names[i] = malloc(32);
        snprintf(names[i], 32, "subject_%05d.mnc", i);
        rstub_register_volume(names[i], Y + (size_t)i * V, sizes);
    }
    const char *mask_name = "mask.mnc";
    if (mask) rstub_register_volume(mask_name, mask, sizes);
    SEXP files = perm_string((const char *const *)names, n);
    SEXP right = perm_logical_na_matrix();
    SEXP mm = perm_real(X, (R_xlen_t)n * p, n, p);
    double hm = mask ? 1.0 : 0.0;
    SEXP have_mask = perm_real(&hm, 1, 0, -1);
    SEXP maskf = perm_string(&mask_name, mask ? 1 : 0);
    SEXP slo = perm_real(&lo, 1, 0, -1), shi = perm_real(&hi, 1, 0, -1);
    const char *m = "lm";
    SEXP method = perm_string(&m, 1);
    int rc = 0;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        alloc_fill_on = 1; alloc_fill_value = fill;
        SEXP res = minc2_model(files, right, mm, R_NilValue, have_mask, maskf, slo, shi,
                               R_NilValue, R_NilValue, method);
        memcpy(out, REAL(res), sizeof(double) * (size_t)V * (2 * p + 3));
    } else {
        rc = 1;
        if (err && errlen) snprintf(err, errlen, "%s", err_msg);
    }
    err_armed = 0;
    alloc_fill_on = 0;
    reset_after_call();
    perm_free(files); perm_free(right); perm_free(mm); perm_free(have_mask); perm_free(maskf);
    perm_free(slo); perm_free(shi); perm_free(method);
    for (int i = 0; i < n; ++i) free(names[i]);
    free(names);
    rstub_clear_volumes();
    return rc;
}

static SEXP perm_int(const int32_t *src, R_xlen_t n)
{
    SEXP s = alloc_obj(INTSXP, n, 1);
    memcpy(s->data, src, (size_t)n * sizeof(int));
    return s;
}

/* the reference's vertex_anova_loop on an in-memory V x n matrix; out is V x (max(asgn)). */
int ref_vertex_anova_loop(const double *Y, int64_t V, int n, const double *X, int p, const int32_t *asgn,
                          double *out, char *err, size_t errlen)
{
    int nterms = 0;
    for (int i = 0; i < p; ++i) if (asgn[i] > nterms) nterms = asgn[i];
    SEXP data = alloc_obj(REALSXP, 0, 1);
    free(data->data);
    data->data = (void *)Y; data->len = (R_xlen_t)V * n; data->nrow = (int)V; data->ncol = n;
    SEXP mm = perm_real(X, (R_xlen_t)n * p, n, p);
    SEXP as = perm_int(asgn, p);
    int rc = 0;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        SEXP res = vertex_anova_loop(data, mm, as);
        memcpy(out, REAL(res), sizeof(double) * (size_t)V * nterms);
    } else {
        rc = 1;
        if (err && errlen) snprintf(err, errlen, "%s", err_msg);
    }
    err_armed = 0;
    reset_after_call();
    data->data = NULL; perm_free(data); perm_free(mm); perm_free(as);
    return rc;
}

/* the reference's per_voxel_anova (slice_loop driver) over registered in-memory volumes. */
int ref_per_voxel_anova(const double *Y, const int64_t sizes[3], int n, const double *X, int p, const int32_t *asgn,
                        const double *mask, double lo, double hi, double *out, char *err, size_t errlen)
{
    int64_t V = sizes[0] * sizes[1] * sizes[2];
    int nterms = 0;
    for (int i = 0; i < p; ++i) if (asgn[i] > nterms) nterms = asgn[i];
    rstub_clear_volumes();
    char **names = malloc(sizeof(char *) * n);
    for (int i = 0; i < n; ++i) {
        names[i] = malloc(32);
        snprintf(names[i], 32, "subject_%05d.mnc", i);
        rstub_register_volume(names[i], Y + (size_t)i * V, sizes);
    }
    const char *mask_name = "mask.mnc";
    if (mask) rstub_register_volume(mask_name, mask, sizes);
    SEXP files = perm_string((const char *const *)names, n);
    SEXP mm = perm_real(X, (R_xlen_t)n * p, n, p);
    SEXP as = perm_int(asgn, p);
    double hm = mask ? 1.0 : 0.0;
    SEXP have_mask = perm_real(&hm, 1, 0, -1);
    SEXP maskf = perm_string(&mask_name, mask ? 1 : 0);
    SEXP slo = perm_real(&lo, 1, 0, -1), shi = perm_real(&hi, 1, 0, -1);
    int rc = 0;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        SEXP res = per_voxel_anova(files, mm, as, have_mask, maskf, slo, shi);
        memcpy(out, REAL(res), sizeof(double) * (size_t)V * nterms);
    } else {
        rc = 1;
        if (err && errlen) snprintf(err, errlen, "%s", err_msg);
    }
    err_armed = 0;
    reset_after_call();
    perm_free(files); perm_free(mm); perm_free(as); perm_free(have_mask); perm_free(maskf); perm_free(slo); perm_free(shi);
    for (int i = 0; i < n; ++i) free(names[i]);
    free(names);
    rstub_clear_volumes();
    return rc;
}

/* the reference's vertex_wlm_loop (anatLm(weights=)) on an in-memory V x n matrix. */
int ref_vertex_wlm_loop(const double *Y, int64_t V, int n, const double *X, int p, const double *w, double *out,
                        char *err, size_t errlen)
{
    SEXP left = alloc_obj(REALSXP, 0, 1);
    free(left->data);
    left->data = (void *)Y; left->len = (R_xlen_t)V * n; left->nrow = (int)V; left->ncol = n;
    SEXP right = perm_logical_na_matrix();
    SEXP mm = perm_real(X, (R_xlen_t)n * p, n, p);
    SEXP ws = perm_real(w, n, 0, -1);
    int rc = 0;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        SEXP res = vertex_wlm_loop(left, right, mm, ws);
        memcpy(out, REAL(res), sizeof(double) * (size_t)V * (2 * p + 3));
    } else {
        rc = 1;
        if (err && errlen) snprintf(err, errlen, "%s", err_msg);
    }
    err_armed = 0;
    reset_after_call();
    left->data = NULL; perm_free(left); perm_free(right); perm_free(mm); perm_free(ws);
    return rc;
}

/* the reference's vertex_lm_loop with a voxel-wise covariate: data_right = Z (V x n); Xs may be NULL
 * (mmatrix = logical NA -> design [1 | z]). */
int ref_vertex_lm_loop_cov(const double *Y, const double *Z, int64_t V, int n, const double *Xs, int ps, double *out,
                           char *err, size_t errlen)
{
    const int p = Xs ? ps + 1 : 2;
    SEXP left = alloc_obj(REALSXP, 0, 1), right = alloc_obj(REALSXP, 0, 1);
    free(left->data); free(right->data);
    left->data = (void *)Y; left->len = (R_xlen_t)V * n; left->nrow = (int)V; left->ncol = n;
    right->data = (void *)Z; right->len = (R_xlen_t)V * n; right->nrow = (int)V; right->ncol = n;
    SEXP mm = Xs ? perm_real(Xs, (R_xlen_t)n * ps, n, ps) : perm_logical_na_matrix();
    int rc = 0;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        SEXP res = vertex_lm_loop(left, right, mm);
        memcpy(out, REAL(res), sizeof(double) * (size_t)V * (2 * p + 3));
    } else {
        rc = 1;
        if (err && errlen) snprintf(err, errlen, "%s", err_msg);
    }
    err_armed = 0;
    reset_after_call();
    left->data = NULL; right->data = NULL;
    perm_free(left); perm_free(right); perm_free(mm);
    return rc;
}
#endif /* RSTUB_NO_REFERENCE */

/* ================================================================================================================
 * The rest of the R API used by the product's .Call shim (rminc_b200/csrc/rshim.c), and a .Call dispatcher.
 *
 * The shim is compiled UNCHANGED against these stand-ins (make rshim / make ref) and executed by tests/: arguments are
 * built as mini-SEXPs, the routine is looked up by name in the table the shim hands to R_registerRoutines (the analogue
 * of src/RMINC_init.c:57-88), the argument count is checked as R's .Call checks it, Rf_error longjmps back here, and
 * every unprotected object is collected at the next allocation -- so a missing PROTECT in the shim shows up as garbage.
 * ================================================================================================================ */
int Rf_isReal(SEXP s) { return s->type == REALSXP; }
int Rf_isInteger(SEXP s) { return s->type == INTSXP; }
int Rf_isMatrix(SEXP s) { return s->ncol >= 0; }
int TYPEOF(SEXP s) { return (int)s->type; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
unsigned char *RAW(SEXP s) { return (unsigned char *)s->data; }

int Rf_asInteger(SEXP s)
{
    if (s->len < 1) return NA_INTEGER;
    switch (s->type) {
    case INTSXP: case LGLSXP: return ((int *)s->data)[0];
    case REALSXP: { double d = ((double *)s->data)[0]; return (d != d || d >= 2147483648.0 || d <= -2147483649.0) ? NA_INTEGER : (int)d; }
    default: return NA_INTEGER;
    }
}
int Rf_asLogical(SEXP s)
{
    if (s->len < 1) return NA_LOGICAL;
    switch (s->type) {
    case LGLSXP: return ((int *)s->data)[0];
    case INTSXP: { int v = ((int *)s->data)[0]; return v == NA_INTEGER ? NA_LOGICAL : v != 0; }
    case REALSXP: { double d = ((double *)s->data)[0]; return d != d ? NA_LOGICAL : d != 0.0; }
    default: return NA_LOGICAL;
    }
}
double Rf_asReal(SEXP s)
{
    if (s->len < 1) return NAN;
    switch (s->type) {
    case REALSXP: return ((double *)s->data)[0];
    case INTSXP: case LGLSXP: { int v = ((int *)s->data)[0]; return v == NA_INTEGER ? NAN : (double)v; }
    default: return NAN;
    }
}

SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v)
{
    ((SEXP *)x->data)[i] = v;
    v->protect_count++;          /* reachable from the list: survives as long as the list does (never collected here) */
    return v;
}

/* R_alloc: transient storage R frees at the end of the .Call */
static void *ralloc_list[4096];
static int ralloc_top = 0;
char *R_alloc(size_t n, int size)
{
    if (ralloc_top >= 4096) Rf_error("R_alloc: too many transient allocations");
    void *p = calloc(n > 0 ? n : 1, (size_t)size);
    if (!p) Rf_error("R_alloc: cannot allocate");
    ralloc_list[ralloc_top++] = p;
    return (char *)p;
}
static void ralloc_reset(void)
{
    while (ralloc_top > 0) free(ralloc_list[--ralloc_top]);
}

/* symbols and options(): Rf_install interns the name, Rf_GetOption1 reads a small table tests can set */
#define MAX_SYMS 64
static SEXP sym_table[MAX_SYMS];
static int n_syms = 0;
SEXP Rf_install(const char *name)
{
    for (int i = 0; i < n_syms; ++i)
        if (strcmp((const char *)sym_table[i]->data, name) == 0) return sym_table[i];
    if (n_syms >= MAX_SYMS) Rf_error("rstub: symbol table full");
    size_t L = strlen(name);
    SEXP s = alloc_obj(CHARSXP, (R_xlen_t)L, 1);
    s->type = SYMSXP;
    memcpy(s->data, name, L + 1);
    return sym_table[n_syms++] = s;
}
static struct { SEXP sym, value; } opt_table[MAX_SYMS];
static int n_opts = 0;
SEXP Rf_GetOption1(SEXP tag)
{
    for (int i = 0; i < n_opts; ++i)
        if (opt_table[i].sym == tag) return opt_table[i].value;
    return R_NilValue;
}
/* options(name = value) for an integer value; value == NA_INTEGER removes the option */
void rstub_set_option_int(const char *name, int value)
{
    SEXP sym = Rf_install(name);
    int slot = -1;
    for (int i = 0; i < n_opts; ++i)
        if (opt_table[i].sym == sym) slot = i;
    if (value == NA_INTEGER) {
        if (slot >= 0) opt_table[slot] = opt_table[--n_opts];
        return;
    }
    if (slot < 0) slot = n_opts++;
    SEXP v = alloc_obj(INTSXP, 1, 1);
    ((int *)v->data)[0] = value;
    opt_table[slot].sym = sym;
    opt_table[slot].value = v;
}

/* external pointers: data is the address itself */
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag; (void)prot;
    SEXP s = alloc_obj(EXTPTRSXP, 0, 0);
    free(s->data);
    s->data = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { if (s->type == EXTPTRSXP) s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, void (*fun)(SEXP), Rboolean onexit) { (void)onexit; s->finalizer = fun; }

/* routine tables */
#define MAX_TABLES 4
static const R_CallMethodDef *call_tables[MAX_TABLES];
static int n_tables = 0;
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)info; (void)c; (void)f; (void)e;
    for (int i = 0; i < n_tables; ++i)
        if (call_tables[i] == call) return 1;
    if (call && n_tables < MAX_TABLES) call_tables[n_tables++] = call;
    return 1;
}
int R_useDynamicSymbols(DllInfo *info, int value) { (void)info; (void)value; return 1; }

#ifndef RSTUB_NO_REFERENCE
/* the reference's own registration lines for the routines compiled into this library (src/RMINC_init.c:57-88) */
static const R_CallMethodDef ref_call_methods[] = {
    {"vertex_lm_loop", (DL_FUNC)&vertex_lm_loop, 3},
    {"vertex_anova_loop", (DL_FUNC)&vertex_anova_loop, 3},
    {"vertex_wlm_loop", (DL_FUNC)&vertex_wlm_loop, 4},
    {"per_voxel_anova", (DL_FUNC)&per_voxel_anova, 7},
    {"minc2_model", (DL_FUNC)&minc2_model, 11},
    {NULL, NULL, 0}};
#endif

void R_init_rmincgpu(DllInfo *dll) __attribute__((weak));     /* present when rshim.c is linked in */

/* dyn.load(): runs the shim's R_init_rmincgpu (and registers the reference's routines where they are compiled in).
 * Returns the number of routines that can be .Call'ed. */
int rstub_dyn_load(void)
{
    if (R_init_rmincgpu) R_init_rmincgpu(NULL);
#ifndef RSTUB_NO_REFERENCE
    R_registerRoutines(NULL, NULL, ref_call_methods, NULL, NULL);
#endif
    int count = 0;
    for (int t = 0; t < n_tables; ++t)
        for (const R_CallMethodDef *m = call_tables[t]; m->name; ++m) ++count;
    return count;
}
/* name and arity of the i-th registered routine (NULL past the end) */
const char *rstub_routine(int i, int *nargs)
{
    for (int t = 0; t < n_tables; ++t)
        for (const R_CallMethodDef *m = call_tables[t]; m->name; ++m)
            if (i-- == 0) { if (nargs) *nargs = m->numArgs; return m->name; }
    return NULL;
}

typedef SEXP (*rcall_fn)();
/* .Call(name, args...).  0 = ok (*result is kept alive until rstub_release), 1 = the routine raised an R error (message
 * in err), 2 = no such routine, 3 = wrong number of arguments (R's own checks for a registered routine). */
int rstub_dotcall(const char *name, int nargs, SEXP *a, SEXP *result, char *err, size_t errlen)
{
    const R_CallMethodDef *hit = NULL;
    for (int t = 0; t < n_tables && !hit; ++t)
        for (const R_CallMethodDef *m = call_tables[t]; m->name; ++m)
            if (strcmp(m->name, name) == 0) { hit = m; break; }
    if (!hit) { snprintf(err, errlen, "C symbol name \"%s\" not in load table", name); return 2; }
    if (hit->numArgs != nargs) {
        snprintf(err, errlen, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, hit->numArgs, name);
        return 3;
    }
    rcall_fn f = (rcall_fn)hit->fun;
    int rc = 0;
    SEXP out = R_NilValue;
    err_armed = 1;
    if (setjmp(err_jmp) == 0) {
        switch (nargs) {
        case 0: out = f(); break;
        case 1: out = f(a[0]); break;
        case 2: out = f(a[0], a[1]); break;
        case 3: out = f(a[0], a[1], a[2]); break;
        case 4: out = f(a[0], a[1], a[2], a[3]); break;
        case 5: out = f(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: out = f(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        case 8: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
        case 9: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
        case 10: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
        case 11: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]); break;
        case 12: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]); break;
        case 13: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]); break;
        case 14: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13]); break;
        case 15: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14]); break;
        case 16: out = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]); break;
        default: snprintf(err, errlen, "rstub: %d arguments not supported", nargs); rc = 3;
        }
        if (rc == 0 && out && out != R_NilValue) out->protect_count++;      /* the value returned to "R" stays alive */
    } else {
        rc = 1;
        snprintf(err, errlen, "%s", err_msg);
        out = R_NilValue;
    }
    err_armed = 0;
    reset_after_call();       /* the protection stack unwinds, unprotected objects are collected */
    ralloc_reset();
    if (result) *result = out;
    return rc;
}

/* ---- objects built by the tests (permanent until rstub_release) */
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_new_real(const double *src, R_xlen_t n, int nrow, int ncol) { return perm_real(src, n, nrow, ncol); }
SEXP rstub_new_int(const int *src, R_xlen_t n, int nrow, int ncol, int logical)
{
    SEXP s = alloc_obj(logical ? LGLSXP : INTSXP, n, 1);
    if (ncol >= 0) { s->nrow = nrow; s->ncol = ncol; }
    if (src) memcpy(s->data, src, (size_t)n * sizeof(int));
    return s;
}
SEXP rstub_new_raw(const unsigned char *src, R_xlen_t n)
{
    SEXP s = alloc_obj(RAWSXP, n, 1);
    if (src) memcpy(s->data, src, (size_t)n);
    return s;
}
SEXP rstub_new_strings(const char *const *strs, int n) { return perm_string(strs, n); }
SEXP rstub_new_list(R_xlen_t n)
{
    SEXP s = alloc_obj(VECSXP, n, 1);
    for (R_xlen_t i = 0; i < n; ++i) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
void rstub_list_set(SEXP list, R_xlen_t i, SEXP v) { ((SEXP *)list->data)[i] = v; }
int rstub_type(SEXP s) { return (int)s->type; }
R_xlen_t rstub_length(SEXP s) { return s->len; }
int rstub_nrow(SEXP s) { return s->nrow; }
int rstub_ncol(SEXP s) { return s->ncol; }
void *rstub_data(SEXP s) { return s->data; }
SEXP rstub_list_get(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
/* drops an object built by rstub_new_* or returned by rstub_dotcall (a list releases its elements; an external
 * pointer runs its finalizer, as R's collector would) */
void rstub_release(SEXP s)
{
    if (!s || s == R_NilValue) return;
    if (s->type == VECSXP)
        for (R_xlen_t i = 0; i < s->len; ++i) {
            SEXP e = ((SEXP *)s->data)[i];
            if (e && e != R_NilValue) { if (e->protect_count > 0) e->protect_count--; rstub_release(e); }
        }
    if (s->type == STRSXP) { perm_free(s); return; }
    if (s->finalizer) s->finalizer(s);
    if (s->in_zero) {           /* still on the collector's list: let the next sweep free it */
        s->protect_count = 0;
        s->permanent = 0;
        s->finalizer = NULL;
        return;
    }
    if (s->type != EXTPTRSXP) free(s->data);
    free(s);
}
