/*
 * oracle/rstub/rstub.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Just enough of R's C API and of libminc2 to run the reference's own
 * src/modelling_functions.c (voxel_lm, minc2_model) and src/vertex_modelling.c
 * (vertex_lm_loop), compiled from /root/reference where they lie, on in-memory
 * data.  Plus the plain-C harness entry points (ref_*) that tests call through
 * ctypes.  LINPACK dqrls_ and LAPACK dpotri_ (third-party, not in the
 * reference tree) are forwarded to oracle/rminc_oracle.c's restatements.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include "Rinternals.h"
#include "R_ext/Utils.h"
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
};

static struct rstub_sexp nil_obj = {NILSXP, 0, 0, -1, 0, 1, NULL, NULL, 0};
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
            free(s->data);
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
    case STRSXP: return sizeof(SEXP);
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
SEXP Rf_install(const char *name) { (void)name; return R_NilValue; }
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
