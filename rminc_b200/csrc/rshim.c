/*
 * rshim.c -- the thin .Call layer between R and the C ABI of include/rminc_b200.h.
 *
 * NOT compiled into librminc_b200.so (R's headers are not part of the CUDA build); the RMINC
 * maintainer adds this one file to RMINC's src/ and links -lrminc_b200.
 * It is kept free of logic: argument unpacking, allocMatrix for the result, rc -> Rf_error.
 * The R API is touched only on the calling thread, never after a failure inside the library
 * has left anything allocated (the C ABI frees everything before it returns).
 *
 * Entry points registered exactly like the reference's (src/RMINC_init.c:57-82):
 *   rmincgpu_vertex_lm_loop  same signature as vertex_lm_loop (src/vertex_modelling.c:8), so
 *                            vertexLm / anatLm only change the routine name
 *   rmincgpu_minc2_model     minc2_model(method = "lm") with the reference's own arguments: the files are read here
 *   rmincgpu_per_voxel_anova per_voxel_anova with the reference's own arguments: the files are read here
 *   rmincgpu_lm              minc2_model(method = "lm") with the volumes already staged
 *   rmincgpu_lm_cov          the same with filenames_right (a voxel-wise covariate) staged as a second matrix
 *   rmincgpu_lm_stored       the same on volumes in their stored type (uint16 + real ranges, float32)
 *   rmincgpu_randomize       the mincRandomize loop (R/minc_voxel_statistics.R:1465-1497)
 *   rmincgpu_randomize_stored  the same on volumes in their stored type
 *   rmincgpu_graph_tfce      graph_tfce_wqu / graph_tfce (src/vertex_tfce.cpp:92-157), every column of a matrix at once
 *   rmincgpu_randomize_tfce  vertexTFCE.vertexLm's permutation loop (R/minc_vertex_statistics.R:892-948)
 *   rmincgpu_volume_tfce, rmincgpu_randomize_volume_tfce   mincTFCE on volumes and mincTFCE.mincLm's permutation loop
 *                            (R/minc_voxel_statistics.R:1186-1439) without the external TFCE program
 *   rmincgpu_fdr             mincFDR's p-value + p.adjust(., "fdr") + threshold step for one column (R/minc_FDR.R:385-505)
 *   rmincgpu_anova           per_voxel_anova / vertex_anova_loop (src/minc_anova.c:142, src/vertex_modelling.c:178)
 *   rmincgpu_dataset_*       the resident dataset: one upload, then mincLm / mincAnova / mincRandomize / mincFDR on the
 *                            shards (an external pointer with a finalizer; R/minc_voxel_statistics.R:911-912, :1457-1477)
 *
 * Single-GPU routines run on device getOption("RMINC_GPU_DEVICE", 0); the sharded ones take the GPU count as an argument
 * (options(RMINC_GPU_N) on the R side).
 */
#include <limits.h>
#include <stdint.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "rminc_b200.h"

static void check(int rc, const char *msg)
{
    if (rc != RMG_OK) Rf_error("%s", msg);   /* e.g. "element (4, 4) is zero, so the inverse cannot be computed" */
}

/* device of the single-GPU routines: options(RMINC_GPU_DEVICE = d), default 0 */
static int gpu_device(void)
{
    SEXP o = Rf_GetOption1(Rf_install("RMINC_GPU_DEVICE"));
    if (Rf_isNull(o)) return 0;
    const int d = Rf_asInteger(o);
    if (d == NA_INTEGER || d < 0) Rf_error("options(RMINC_GPU_DEVICE) must be a non-negative device number");
    return d;
}

/* an R matrix has int dimensions */
static int matrix_rows(R_xlen_t V)
{
    if (V > INT_MAX) Rf_error("%.0f voxels do not fit the rows of an R matrix (at most %d)", (double)V, INT_MAX);
    return (int)V;
}

static int flag_arg(SEXP s, const char *what)
{
    const int v = Rf_asLogical(s);
    if (v == NA_LOGICAL) Rf_error("%s must be TRUE or FALSE", what);
    return v;
}

static int count_arg(SEXP s, const char *what)
{
    const int v = Rf_asInteger(s);
    if (v == NA_INTEGER) Rf_error("%s must not be NA", what);
    return v;
}

/* 1-based R indices -> 0-based, in R_alloc'd memory (freed by R at the end of the .Call) */
static int *zero_based(SEXP x, R_xlen_t count, const char *what)
{
    int *out = (int *)R_alloc((size_t)(count > 0 ? count : 1), sizeof(int));
    for (R_xlen_t i = 0; i < count; ++i) {
        if (INTEGER(x)[i] == NA_INTEGER) Rf_error("%s must not contain NA", what);
        out[i] = INTEGER(x)[i] - 1;
    }
    return out;
}

/* Arguments of the stored-type routines.  The C ABI reads ceil(V / slice_len) x n doubles from scale and offset: a table
 * with fewer rows, a slice length < 1 or NA would be an out-of-bounds host read there, so it is an R error here. */
static int64_t stored_slice_len(int dt, SEXP scale, SEXP offset, int n, R_xlen_t V, SEXP slice_len)
{
    if (dt != RMG_STORED_U16 && dt != RMG_STORED_F32) Rf_error("dtype must be 1 (uint16) or 2 (float32)");
    if (dt != RMG_STORED_U16) return V > 0 ? (int64_t)V : 1;
    const double sl = Rf_asReal(slice_len);
    if (!(sl >= 1.0) || sl != (double)(int64_t)sl) Rf_error("slice_len must be a whole number >= 1");
    if (!Rf_isReal(scale) || !Rf_isReal(offset) || Rf_ncols(scale) != n || Rf_ncols(offset) != n)
        Rf_error("scale and offset must be nslices x n double matrices");
    const int64_t nslices = V > 0 ? ((int64_t)V + (int64_t)sl - 1) / (int64_t)sl : 0;
    if ((int64_t)Rf_nrows(scale) != nslices || (int64_t)Rf_nrows(offset) != nslices)
        Rf_error("scale and offset must have ceiling(V / slice_len) = %.0f rows (got %d and %d)", (double)nslices,
                 Rf_nrows(scale), Rf_nrows(offset));
    return (int64_t)sl;
}

static const double *mask_or_null(SEXP mask, R_xlen_t V)
{
    if (Rf_isNull(mask)) return NULL;
    if (!Rf_isReal(mask) || XLENGTH(mask) != V) Rf_error("Mask Volume input volume dimension mismatch.");
    return REAL(mask);
}

/* vertex_lm_loop with data_right a V x n matrix: the design of vertex v is [mmatrix | data_right[v, ]], or
 * [1 | data_right[v, ]] when mmatrix is the logical NA matrix (src/vertex_modelling.c:33-55, :112-140). */
static SEXP vertex_lm_loop_cov(SEXP data_left, SEXP data_right, SEXP mmatrix)
{
    if (!Rf_isReal(data_left) || !Rf_isReal(data_right)) Rf_error("data_left and data_right must be double matrices");
    const R_xlen_t V = Rf_nrows(data_left);
    const int n = Rf_ncols(data_left);
    if (Rf_nrows(data_right) != V || Rf_ncols(data_right) != n) Rf_error("data_right must have the shape of data_left");
    const int has_static = !Rf_isLogical(mmatrix);
    if (has_static && (!Rf_isReal(mmatrix) || Rf_nrows(mmatrix) != n)) Rf_error("model matrix does not match the data");
    const int ps = has_static ? Rf_ncols(mmatrix) : 0, p = has_static ? ps + 1 : 2;
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), 2 * p + 3));
    char err[512] = "";
    int rc = rmg_vertex_lm_cov(REAL(data_left), REAL(data_right), V, V > 0 ? V : 1, n, has_static ? REAL(mmatrix) : NULL, ps,
                               REAL(out), V > 0 ? V : 1, gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

SEXP rmincgpu_vertex_lm_loop(SEXP data_left, SEXP data_right, SEXP mmatrix)
{
    if (!Rf_isLogical(data_right)) return vertex_lm_loop_cov(data_left, data_right, mmatrix);
    if (!Rf_isReal(data_left) || !Rf_isReal(mmatrix)) Rf_error("data_left and mmatrix must be double matrices");
    const R_xlen_t V = Rf_nrows(data_left);
    const int n = Rf_ncols(data_left), p = Rf_ncols(mmatrix);
    if (Rf_nrows(mmatrix) != n) Rf_error("model matrix has %d rows, data has %d subjects", Rf_nrows(mmatrix), n);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), 2 * p + 3));
    char err[512] = "";
    int rc = rmg_vertex_lm(REAL(data_left), V, V > 0 ? V : 1, n, REAL(mmatrix), p, REAL(out), V > 0 ? V : 1, gpu_device(),
                           err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* vertex_wlm_loop (src/weighted_least_squares.c:168): same arity as the reference routine. */
SEXP rmincgpu_vertex_wlm_loop(SEXP data_left, SEXP data_right, SEXP mmatrix, SEXP ws)
{
    if (!Rf_isLogical(data_right))
        Rf_error("voxel-wise covariates (a file term on the right-hand side) are not on the GPU path");
    if (!Rf_isReal(data_left) || !Rf_isReal(mmatrix) || !Rf_isReal(ws)) Rf_error("data_left, mmatrix and ws must be double");
    const R_xlen_t V = Rf_nrows(data_left);
    const int n = Rf_ncols(data_left), p = Rf_ncols(mmatrix);
    if (Rf_nrows(mmatrix) != n || LENGTH(ws) != n) Rf_error("Incorrect number of weights supplied, need one per observation");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), 2 * p + 3));
    char err[512] = "";
    int rc = rmg_vertex_wlm(REAL(data_left), V, V > 0 ? V : 1, n, REAL(mmatrix), p, REAL(ws), REAL(out), V > 0 ? V : 1,
                            gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

SEXP rmincgpu_lm(SEXP Y, SEXP mmatrix, SEXP mask, SEXP mask_lower, SEXP mask_upper, SEXP ngpu)
{
    if (!Rf_isReal(Y) || !Rf_isReal(mmatrix)) Rf_error("Y and mmatrix must be double matrices");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y), p = Rf_ncols(mmatrix);
    if (Rf_nrows(mmatrix) != n) Rf_error("model matrix has %d rows, data has %d subjects", Rf_nrows(mmatrix), n);
    const double *m = mask_or_null(mask, V);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), 2 * p + 3));
    char err[512] = "";
    const int G = count_arg(ngpu, "ngpu");
    int rc = G > 1 ? rmg_lm_fit_multi(REAL(Y), V, V > 0 ? V : 1, n, REAL(mmatrix), p, m, Rf_asReal(mask_lower),
                                      Rf_asReal(mask_upper), REAL(out), V > 0 ? V : 1, G, err, sizeof err)
                   : rmg_lm_fit(REAL(Y), V, V > 0 ? V : 1, n, REAL(mmatrix), p, m, Rf_asReal(mask_lower),
                                Rf_asReal(mask_upper), REAL(out), V > 0 ? V : 1, gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* minc2_model(method = "lm") with filenames_right: Z staged like Y; mmatrix may be the logical NA matrix. */
SEXP rmincgpu_lm_cov(SEXP Y, SEXP Z, SEXP mmatrix, SEXP mask, SEXP mask_lower, SEXP mask_upper)
{
    if (!Rf_isReal(Y) || !Rf_isReal(Z)) Rf_error("Y and Z must be double matrices");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y);
    if (Rf_nrows(Z) != V || Rf_ncols(Z) != n) Rf_error("the right-hand-side volumes must match the left-hand-side volumes");
    const int has_static = !Rf_isLogical(mmatrix);
    if (has_static && (!Rf_isReal(mmatrix) || Rf_nrows(mmatrix) != n)) Rf_error("model matrix does not match the data");
    const int ps = has_static ? Rf_ncols(mmatrix) : 0, p = has_static ? ps + 1 : 2;
    const double *m = mask_or_null(mask, V);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), 2 * p + 3));
    char err[512] = "";
    int rc = rmg_lm_fit_cov(REAL(Y), REAL(Z), V, V > 0 ? V : 1, n, has_static ? REAL(mmatrix) : NULL, ps, m,
                            Rf_asReal(mask_lower), Rf_asReal(mask_upper), REAL(out), V > 0 ? V : 1, gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* mincLm on volumes in their stored type.  U is a RAWSXP holding V x n samples (uint16 or float32, column-major,
 * filled by a staging routine that calls miget_voxel_value_hyperslab with the file's own type -- INTEGRATION.md);
 * scale / offset are nslices x n double matrices from image-min / image-max and the valid range. */
SEXP rmincgpu_lm_stored(SEXP U, SEXP dtype, SEXP nvoxels, SEXP scale, SEXP offset, SEXP voxel_min, SEXP slice_len,
                        SEXP mmatrix, SEXP mask, SEXP mask_lower, SEXP mask_upper)
{
    if (TYPEOF(U) != RAWSXP || !Rf_isReal(mmatrix)) Rf_error("U must be a raw vector and mmatrix a double matrix");
    const int dt = count_arg(dtype, "dtype");
    if (dt != RMG_STORED_U16 && dt != RMG_STORED_F32) Rf_error("dtype must be 1 (uint16) or 2 (float32)");
    const size_t es = dt == RMG_STORED_U16 ? 2 : 4;
    const R_xlen_t V = (R_xlen_t)Rf_asReal(nvoxels);
    const int n = Rf_nrows(mmatrix), p = Rf_ncols(mmatrix);
    if (V < 0 || (R_xlen_t)(XLENGTH(U) / es) != V * n) Rf_error("U holds %.0f samples, expected %.0f", (double)(XLENGTH(U) / es), (double)V * n);
    const int scaled = dt == RMG_STORED_U16;
    const int64_t sl = stored_slice_len(dt, scale, offset, n, V, slice_len);
    const double *m = mask_or_null(mask, V);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), 2 * p + 3));
    char err[512] = "";
    int rc = rmg_lm_fit_stored(RAW(U), dt, V, V > 0 ? V : 1, n, scaled ? REAL(scale) : NULL, scaled ? REAL(offset) : NULL,
                               Rf_asReal(voxel_min), sl, REAL(mmatrix), p, m, Rf_asReal(mask_lower), Rf_asReal(mask_upper),
                               REAL(out), V > 0 ? V : 1, gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

SEXP rmincgpu_randomize(SEXP Y, SEXP mmatrix, SEXP mask, SEXP mask_lower, SEXP mask_upper, SEXP idx,
                        SEXP columns, SEXP two_sided, SEXP ngpu)
{
    if (!Rf_isReal(Y) || !Rf_isReal(mmatrix) || !Rf_isInteger(idx) || !Rf_isInteger(columns))
        Rf_error("Y/mmatrix must be double, idx/columns integer");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y), p = Rf_ncols(mmatrix), R = Rf_ncols(idx), nc = LENGTH(columns);
    if (Rf_nrows(idx) != n) Rf_error("idx must have one row per subject");
    const double *m = mask_or_null(mask, V);
    /* R is 1-based: idx holds sample(seq_len(n)), columns are column numbers of the mincLm matrix */
    int *idx0 = zero_based(idx, (R_xlen_t)n * R, "idx");
    int *col0 = zero_based(columns, nc, "columns");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, nc));
    char err[512] = "";
    const int G = count_arg(ngpu, "ngpu");
    int rc = G > 1 ? rmg_randomize_multi(REAL(Y), V, V > 0 ? V : 1, n, REAL(mmatrix), p, m, Rf_asReal(mask_lower),
                                         Rf_asReal(mask_upper), idx0, R, col0, nc, flag_arg(two_sided, "two_sided"),
                                         REAL(out), G, err, sizeof err)
                   : rmg_randomize(REAL(Y), V, V > 0 ? V : 1, n, REAL(mmatrix), p, m, Rf_asReal(mask_lower),
                                   Rf_asReal(mask_upper), idx0, R, col0, nc, flag_arg(two_sided, "two_sided"), REAL(out),
                                   gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* The mincRandomize loop of a model with a voxel-wise covariate: Z (V x n, the right-hand-side volumes) is re-sampled with
 * the rows of the static model matrix (NULL: the intercept), R/minc_voxel_statistics.R:1468-1477. */
SEXP rmincgpu_randomize_cov(SEXP Y, SEXP Z, SEXP mmatrix, SEXP mask, SEXP mask_lower, SEXP mask_upper, SEXP idx, SEXP columns,
                            SEXP two_sided, SEXP ngpu)
{
    if (!Rf_isReal(Y) || !Rf_isReal(Z) || !Rf_isInteger(idx) || !Rf_isInteger(columns))
        Rf_error("Y/Z must be double, idx/columns integer");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y), R = Rf_ncols(idx), nc = LENGTH(columns);
    if (Rf_nrows(Z) != V || Rf_ncols(Z) != n) Rf_error("the covariate volumes must match the left-hand-side volumes");
    const int has_static = !Rf_isNull(mmatrix);
    if (has_static && (!Rf_isReal(mmatrix) || Rf_nrows(mmatrix) != n)) Rf_error("mmatrix must be a double matrix with one row per subject");
    const int ps = has_static ? Rf_ncols(mmatrix) : 0;
    if (Rf_nrows(idx) != n) Rf_error("idx must have one row per subject");
    const double *m = mask_or_null(mask, V);
    int *idx0 = zero_based(idx, (R_xlen_t)n * R, "idx");
    int *col0 = zero_based(columns, nc, "columns");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, nc));
    char err[512] = "";
    const int G = count_arg(ngpu, "ngpu");
    const double *xs = has_static ? REAL(mmatrix) : NULL;
    int rc = G > 1 ? rmg_randomize_cov_multi(REAL(Y), REAL(Z), V, V > 0 ? V : 1, n, xs, ps, m, Rf_asReal(mask_lower),
                                             Rf_asReal(mask_upper), idx0, R, col0, nc, flag_arg(two_sided, "two_sided"),
                                             REAL(out), G, err, sizeof err)
                   : rmg_randomize_cov(REAL(Y), REAL(Z), V, V > 0 ? V : 1, n, xs, ps, m, Rf_asReal(mask_lower),
                                       Rf_asReal(mask_upper), idx0, R, col0, nc, flag_arg(two_sided, "two_sided"), REAL(out),
                                       gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* The mincRandomize loop on volumes in their stored type (U, dtype .. slice_len as rmincgpu_lm_stored). */
SEXP rmincgpu_randomize_stored(SEXP U, SEXP dtype, SEXP nvoxels, SEXP scale, SEXP offset, SEXP voxel_min, SEXP slice_len,
                               SEXP mmatrix, SEXP mask, SEXP mask_lower, SEXP mask_upper, SEXP idx, SEXP columns,
                               SEXP two_sided, SEXP ngpu)
{
    if (TYPEOF(U) != RAWSXP || !Rf_isReal(mmatrix) || !Rf_isInteger(idx) || !Rf_isInteger(columns))
        Rf_error("U must be a raw vector, mmatrix double, idx/columns integer");
    const int dt = count_arg(dtype, "dtype");
    if (dt != RMG_STORED_U16 && dt != RMG_STORED_F32) Rf_error("dtype must be 1 (uint16) or 2 (float32)");
    const size_t es = dt == RMG_STORED_U16 ? 2 : 4;
    const R_xlen_t V = (R_xlen_t)Rf_asReal(nvoxels);
    const int n = Rf_nrows(mmatrix), p = Rf_ncols(mmatrix), R = Rf_ncols(idx), nc = LENGTH(columns);
    if (V < 0 || (R_xlen_t)(XLENGTH(U) / es) != V * n) Rf_error("U holds %.0f samples, expected %.0f", (double)(XLENGTH(U) / es), (double)V * n);
    if (Rf_nrows(idx) != n) Rf_error("idx must have one row per subject");
    const int scaled = dt == RMG_STORED_U16;
    const int64_t sl = stored_slice_len(dt, scale, offset, n, V, slice_len);
    const double *m = mask_or_null(mask, V);
    int *idx0 = zero_based(idx, (R_xlen_t)n * R, "idx");
    int *col0 = zero_based(columns, nc, "columns");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, nc));
    char err[512] = "";
    const int G = count_arg(ngpu, "ngpu");
    const double *sc = scaled ? REAL(scale) : NULL, *of = scaled ? REAL(offset) : NULL;
    int rc = G > 1 ? rmg_randomize_stored_multi(RAW(U), dt, V, V > 0 ? V : 1, n, sc, of, Rf_asReal(voxel_min),
                                                sl, REAL(mmatrix), p, m, Rf_asReal(mask_lower),
                                                Rf_asReal(mask_upper), idx0, R, col0, nc, flag_arg(two_sided, "two_sided"), REAL(out),
                                                G, err, sizeof err)
                   : rmg_randomize_stored(RAW(U), dt, V, V > 0 ? V : 1, n, sc, of, Rf_asReal(voxel_min),
                                          sl, REAL(mmatrix), p, m, Rf_asReal(mask_lower),
                                          Rf_asReal(mask_upper), idx0, R, col0, nc, flag_arg(two_sided, "two_sided"), REAL(out),
                                          gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* per_voxel_anova / vertex_anova_loop replacement: V x max(asgn) F-statistics. */
SEXP rmincgpu_anova(SEXP Y, SEXP mmatrix, SEXP asgn, SEXP mask, SEXP mask_lower, SEXP mask_upper)
{
    if (!Rf_isReal(Y) || !Rf_isReal(mmatrix) || !Rf_isInteger(asgn)) Rf_error("Y/mmatrix must be double, asgn integer");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y), p = Rf_ncols(mmatrix);
    if (Rf_nrows(mmatrix) != n || LENGTH(asgn) != p) Rf_error("model matrix / assign do not match the data");
    int nterms = 0;
    for (int i = 0; i < p; ++i) if (INTEGER(asgn)[i] > nterms) nterms = INTEGER(asgn)[i];
    const double *m = mask_or_null(mask, V);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), nterms));
    char err[512] = "";
    int rc = rmg_anova_fit(REAL(Y), V, V > 0 ? V : 1, n, REAL(mmatrix), p, INTEGER(asgn), m, Rf_asReal(mask_lower),
                           Rf_asReal(mask_upper), REAL(out), V > 0 ? V : 1, gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* mincFDR's inner step for one column (R/minc_FDR.R:385-505): list(qvalue = <V doubles>, thresholds = <5 doubles>).
 * stat_type: 1 = t (two-sided, pt2), 2 = F (upper tail); df = c(df1) or c(df1, df2); mask NULL or V doubles (> 0.5). */
SEXP rmincgpu_fdr(SEXP stat, SEXP stat_type, SEXP df, SEXP mask)
{
    if (!Rf_isReal(stat) || !Rf_isReal(df)) Rf_error("stat and df must be double vectors");
    const R_xlen_t V = XLENGTH(stat);
    const double *m = mask_or_null(mask, V);
    SEXP q = PROTECT(Rf_allocVector(REALSXP, V));
    SEXP th = PROTECT(Rf_allocVector(REALSXP, 5));
    char err[512] = "";
    int rc = rmg_fdr(REAL(stat), V, Rf_asInteger(stat_type), REAL(df)[0], LENGTH(df) > 1 ? REAL(df)[1] : 0.0, m, REAL(q),
                     REAL(th), gpu_device(), err, sizeof err);
    if (rc != RMG_OK) {
        UNPROTECT(2);
        check(rc, err);
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, q);
    SET_VECTOR_ELT(out, 1, th);
    UNPROTECT(3);
    return out;
}

/* The same with mincFDR's `method` (R/minc_FDR.R:429-438): 0 = "FDR" / "p.adjust", 1 = "BY", 2 = "pFDR" / "fastqvalue",
 * 3 = "qvalue"; pi0 (methods 2 and 3) is what the R side gets by smoothing rmincgpu_fdr_pi0's fractions with
 * smooth.spline, exactly as fast.qvalue does (R/minc_misc.R:437-451). */
SEXP rmincgpu_fdr_method(SEXP stat, SEXP stat_type, SEXP df, SEXP mask, SEXP method, SEXP pi0)
{
    if (!Rf_isReal(stat) || !Rf_isReal(df)) Rf_error("stat and df must be double vectors");
    const R_xlen_t V = XLENGTH(stat);
    const double *m = mask_or_null(mask, V);
    const int meth = Rf_asInteger(method);
    if (meth == NA_INTEGER) Rf_error("method must not be NA");
    SEXP q = PROTECT(Rf_allocVector(REALSXP, V));
    SEXP th = PROTECT(Rf_allocVector(REALSXP, 5));
    char err[512] = "";
    int rc = rmg_fdr_method(REAL(stat), V, Rf_asInteger(stat_type), REAL(df)[0], LENGTH(df) > 1 ? REAL(df)[1] : 0.0, m, REAL(q),
                            REAL(th), meth, Rf_asReal(pi0), gpu_device(), err, sizeof err);
    if (rc != RMG_OK) {
        UNPROTECT(2);
        check(rc, err);
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, q);
    SET_VECTOR_ELT(out, 1, th);
    UNPROTECT(3);
    return out;
}

/* mean(p >= lambda) for every lambda over the voxels with mask > 0.5 (R/minc_misc.R:439-441) */
SEXP rmincgpu_fdr_pi0(SEXP stat, SEXP stat_type, SEXP df, SEXP mask, SEXP lambda)
{
    if (!Rf_isReal(stat) || !Rf_isReal(df) || !Rf_isReal(lambda)) Rf_error("stat, df and lambda must be double vectors");
    const R_xlen_t V = XLENGTH(stat);
    const double *m = mask_or_null(mask, V);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, XLENGTH(lambda)));
    char err[512] = "";
    int rc = rmg_fdr_pi0_fractions(REAL(stat), V, Rf_asInteger(stat_type), REAL(df)[0], LENGTH(df) > 1 ? REAL(df)[1] : 0.0, m,
                                   REAL(lambda), (int)XLENGTH(lambda), REAL(out), gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* Adjacency list (R list of integer vectors, 0-based as vertexTFCE documents) -> CSR in R_alloc'd memory. */
static void adjacency_to_csr(SEXP adjacencies, int **ptr_out, int **idx_out)
{
    const R_xlen_t V = XLENGTH(adjacencies);
    int *ptr = (int *)R_alloc((size_t)V + 1, sizeof(int));
    ptr[0] = 0;
    for (R_xlen_t v = 0; v < V; ++v) ptr[v + 1] = ptr[v] + LENGTH(VECTOR_ELT(adjacencies, v));
    int *idx = (int *)R_alloc((size_t)(ptr[V] > 0 ? ptr[V] : 1), sizeof(int));
    for (R_xlen_t v = 0; v < V; ++v) {
        SEXP a = VECTOR_ELT(adjacencies, v);
        for (int e = 0; e < LENGTH(a); ++e)
            idx[ptr[v] + e] = Rf_isInteger(a) ? INTEGER(a)[e] : (int)REAL(a)[e];
    }
    *ptr_out = ptr;
    *idx_out = idx;
}

/* graph_tfce_wqu / graph_tfce (src/vertex_tfce.cpp:92-157) for every column of `map`: same arguments as the Rcpp
 * exports plus side (0 both, 1 positive, 2 negative) so that one routine serves vertexTFCE.numeric's three branches. */
SEXP rmincgpu_graph_tfce(SEXP map, SEXP adjacencies, SEXP E, SEXP H, SEXP nsteps, SEXP weights, SEXP side)
{
    if (!Rf_isReal(map) || !Rf_isReal(weights)) Rf_error("map and weights must be double");
    const int ismat = Rf_ncols(map) > 1 || Rf_nrows(map) != LENGTH(map);
    const R_xlen_t V = ismat ? Rf_nrows(map) : XLENGTH(map);
    const int nmaps = ismat ? Rf_ncols(map) : 1;
    if (XLENGTH(adjacencies) != V) Rf_error("Mismatch between adjacency list and data");
    int *ptr, *idx;
    adjacency_to_csr(adjacencies, &ptr, &idx);
    SEXP out = PROTECT(ismat ? Rf_allocMatrix(REALSXP, matrix_rows(V), nmaps) : Rf_allocVector(REALSXP, V));
    char err[512] = "";
    int rc = rmg_graph_tfce(REAL(map), V, nmaps, ptr, idx, XLENGTH(weights) == V ? REAL(weights) : NULL, Rf_asReal(E),
                            Rf_asReal(H), Rf_asInteger(nsteps), Rf_asInteger(side), REAL(out), gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* vertexTFCE.vertexLm's mincRandomize_core(post_proc = vertexTFCE.matrix) loop in one call. */
SEXP rmincgpu_randomize_tfce(SEXP Y, SEXP mmatrix, SEXP idx, SEXP columns, SEXP two_sided, SEXP adjacencies, SEXP weights,
                             SEXP E, SEXP H, SEXP nsteps, SEXP side)
{
    if (!Rf_isReal(Y) || !Rf_isReal(mmatrix) || !Rf_isInteger(idx) || !Rf_isInteger(columns))
        Rf_error("Y/mmatrix must be double, idx/columns integer");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y), p = Rf_ncols(mmatrix), R = Rf_ncols(idx), nc = LENGTH(columns);
    if (Rf_nrows(idx) != n || XLENGTH(adjacencies) != V) Rf_error("idx / adjacency list do not match the data");
    int *idx0 = zero_based(idx, (R_xlen_t)n * R, "idx");
    int *col0 = zero_based(columns, nc, "columns");
    int *ptr, *aidx;
    adjacency_to_csr(adjacencies, &ptr, &aidx);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, nc));
    char err[512] = "";
    int rc = rmg_randomize_tfce(REAL(Y), V, V > 0 ? V : 1, n, REAL(mmatrix), p, idx0, R, col0, nc, flag_arg(two_sided, "two_sided"), ptr,
                                aidx, Rf_isReal(weights) && XLENGTH(weights) == V ? REAL(weights) : NULL, Rf_asReal(E),
                                Rf_asReal(H), Rf_asInteger(nsteps), Rf_asInteger(side), REAL(out), gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* mincTFCE.mincSingleDim / .matrix without the trip through a temporary file and the external TFCE program
 * (R/minc_voxel_statistics.R:1186-1283): map is a vector or a V x nmaps matrix, sizes = minc.dimensions.sizes(likeVolume)
 * (file order), voxel_volume = prod(abs(minc.separation.sizes(likeVolume))). */
SEXP rmincgpu_volume_tfce(SEXP map, SEXP sizes, SEXP connectivity, SEXP voxel_volume, SEXP d, SEXP E, SEXP H, SEXP side)
{
    if (!Rf_isReal(map)) Rf_error("map must be double");
    if (LENGTH(sizes) != 3) Rf_error("sizes must have three entries");
    const int ismat = Rf_isMatrix(map);
    const R_xlen_t V = ismat ? Rf_nrows(map) : XLENGTH(map);
    const int nmaps = ismat ? Rf_ncols(map) : 1;
    int64_t dims[3];
    for (int i = 0; i < 3; ++i) dims[i] = (int64_t)(Rf_isReal(sizes) ? REAL(sizes)[i] : INTEGER(sizes)[i]);
    if (dims[0] * dims[1] * dims[2] != V) Rf_error("sizes do not match the data");
    SEXP out = PROTECT(ismat ? Rf_allocMatrix(REALSXP, matrix_rows(V), nmaps) : Rf_allocVector(REALSXP, V));
    char err[512] = "";
    int rc = rmg_volume_tfce(REAL(map), dims, nmaps, Rf_asInteger(connectivity), Rf_asReal(voxel_volume), Rf_asReal(d),
                             Rf_asReal(E), Rf_asReal(H), Rf_asInteger(side), REAL(out), gpu_device(), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* mincTFCE.mincLm's mincRandomize_core(post_proc = mincTFCE.matrix) loop in one call (:1382-1439). */
SEXP rmincgpu_randomize_volume_tfce(SEXP Y, SEXP mmatrix, SEXP mask, SEXP mask_lower, SEXP mask_upper, SEXP idx, SEXP columns,
                                    SEXP two_sided, SEXP sizes, SEXP connectivity, SEXP voxel_volume, SEXP d, SEXP E, SEXP H,
                                    SEXP side, SEXP ngpu)
{
    if (!Rf_isReal(Y) || !Rf_isReal(mmatrix) || !Rf_isInteger(idx) || !Rf_isInteger(columns))
        Rf_error("Y/mmatrix must be double, idx/columns integer");
    if (LENGTH(sizes) != 3) Rf_error("sizes must have three entries");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y), p = Rf_ncols(mmatrix), R = Rf_ncols(idx), nc = LENGTH(columns);
    if (Rf_nrows(idx) != n) Rf_error("idx must have one row per subject");
    const double *m = mask_or_null(mask, V);
    int64_t dims[3];
    for (int i = 0; i < 3; ++i) dims[i] = (int64_t)(Rf_isReal(sizes) ? REAL(sizes)[i] : INTEGER(sizes)[i]);
    int *idx0 = zero_based(idx, (R_xlen_t)n * R, "idx");
    int *col0 = zero_based(columns, nc, "columns");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, nc));
    char err[512] = "";
    int rc = rmg_randomize_volume_tfce(REAL(Y), V, V > 0 ? V : 1, n, REAL(mmatrix), p, m, Rf_asReal(mask_lower),
                                       Rf_asReal(mask_upper), idx0, R, col0, nc, flag_arg(two_sided, "two_sided"), dims,
                                       Rf_asInteger(connectivity), Rf_asReal(voxel_volume), Rf_asReal(d), Rf_asReal(E),
                                       Rf_asReal(H), Rf_asInteger(side), REAL(out), count_arg(ngpu, "ngpu"), err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* ---- the resident dataset: Y is uploaded once (sharded over ngpu devices) and stays in HBM behind an external pointer;
 *      mincLm, mincAnova, mincRandomize and mincFDR then run on the resident shards (rmg_dataset_*).  The pointer's
 *      finalizer frees the device memory when R collects the handle; rmincgpu_dataset_free does it at once. */
static void dataset_finalizer(SEXP h)
{
    rmg_dataset *ds = (rmg_dataset *)R_ExternalPtrAddr(h);
    if (ds) rmg_dataset_free(ds);
    R_ClearExternalPtr(h);
}

static rmg_dataset *dataset_of(SEXP h)
{
    if (TYPEOF(h) != EXTPTRSXP) Rf_error("not a GPU dataset handle");
    rmg_dataset *ds = (rmg_dataset *)R_ExternalPtrAddr(h);
    if (!ds) Rf_error("the GPU dataset has been freed");
    return ds;
}

SEXP rmincgpu_dataset_create(SEXP Y, SEXP mask, SEXP mask_lower, SEXP mask_upper, SEXP ngpu)
{
    if (!Rf_isReal(Y) || !Rf_isMatrix(Y)) Rf_error("Y must be a double matrix (voxels x subjects)");
    const R_xlen_t V = Rf_nrows(Y);
    const int n = Rf_ncols(Y);
    const double *m = mask_or_null(mask, V);
    rmg_dataset *ds = NULL;
    char err[512] = "";
    int rc = rmg_dataset_create(REAL(Y), V, V > 0 ? V : 1, n, m, Rf_asReal(mask_lower), Rf_asReal(mask_upper),
                                count_arg(ngpu, "ngpu"), &ds, err, sizeof err);
    /* Y lives on R's heap, which the collector may move or free after this call returns: wait for the upload */
    if (rc == RMG_OK) rc = rmg_dataset_sync(ds, err, sizeof err);
    if (rc != RMG_OK) {
        if (ds) rmg_dataset_free(ds);
        check(rc, err);
    }
    SEXP h = PROTECT(R_MakeExternalPtr(ds, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(h, dataset_finalizer, TRUE);
    UNPROTECT(1);
    return h;
}

SEXP rmincgpu_dataset_free(SEXP h)
{
    if (TYPEOF(h) != EXTPTRSXP) Rf_error("not a GPU dataset handle");
    dataset_finalizer(h);
    return R_NilValue;
}

SEXP rmincgpu_dataset_lm(SEXP h, SEXP mmatrix)
{
    rmg_dataset *ds = dataset_of(h);
    if (!Rf_isReal(mmatrix)) Rf_error("mmatrix must be a double matrix");
    const R_xlen_t V = (R_xlen_t)rmg_dataset_voxels(ds);
    const int n = rmg_dataset_subjects(ds), p = Rf_ncols(mmatrix);
    if (Rf_nrows(mmatrix) != n) Rf_error("model matrix has %d rows, data has %d subjects", Rf_nrows(mmatrix), n);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), 2 * p + 3));
    char err[512] = "";
    int rc = rmg_dataset_lm_fit(ds, REAL(mmatrix), p, REAL(out), V > 0 ? V : 1, err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

SEXP rmincgpu_dataset_anova(SEXP h, SEXP mmatrix, SEXP asgn)
{
    rmg_dataset *ds = dataset_of(h);
    if (!Rf_isReal(mmatrix) || !Rf_isInteger(asgn)) Rf_error("mmatrix must be double, asgn integer");
    const R_xlen_t V = (R_xlen_t)rmg_dataset_voxels(ds);
    const int n = rmg_dataset_subjects(ds), p = Rf_ncols(mmatrix);
    if (Rf_nrows(mmatrix) != n || LENGTH(asgn) != p) Rf_error("model matrix / assign do not match the data");
    int nterms = 0;
    for (int i = 0; i < p; ++i) if (INTEGER(asgn)[i] > nterms) nterms = INTEGER(asgn)[i];
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows(V), nterms));
    char err[512] = "";
    int rc = rmg_dataset_anova(ds, REAL(mmatrix), p, INTEGER(asgn), REAL(out), V > 0 ? V : 1, err, sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

SEXP rmincgpu_dataset_randomize(SEXP h, SEXP mmatrix, SEXP idx, SEXP columns, SEXP two_sided)
{
    rmg_dataset *ds = dataset_of(h);
    if (!Rf_isReal(mmatrix) || !Rf_isInteger(idx) || !Rf_isInteger(columns)) Rf_error("mmatrix must be double, idx/columns integer");
    const int n = rmg_dataset_subjects(ds), p = Rf_ncols(mmatrix), R = Rf_ncols(idx), nc = LENGTH(columns);
    if (Rf_nrows(mmatrix) != n || Rf_nrows(idx) != n) Rf_error("model matrix / idx must have one row per subject");
    int *idx0 = zero_based(idx, (R_xlen_t)n * R, "idx");
    int *col0 = zero_based(columns, nc, "columns");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, nc));
    char err[512] = "";
    int rc = rmg_dataset_randomize(ds, REAL(mmatrix), p, idx0, R, col0, nc, flag_arg(two_sided, "two_sided"), REAL(out), err,
                                   sizeof err);
    UNPROTECT(1);
    check(rc, err);
    return out;
}

/* mincFDR's step for column `column` (1-based) of the result of the last fit on the handle */
SEXP rmincgpu_dataset_fdr(SEXP h, SEXP column, SEXP stat_type, SEXP df)
{
    rmg_dataset *ds = dataset_of(h);
    if (!Rf_isReal(df) || LENGTH(df) < 1) Rf_error("df must be a double vector");
    const R_xlen_t V = (R_xlen_t)rmg_dataset_voxels(ds);
    SEXP q = PROTECT(Rf_allocVector(REALSXP, V));
    SEXP th = PROTECT(Rf_allocVector(REALSXP, 5));
    char err[512] = "";
    int rc = rmg_dataset_fdr(ds, count_arg(column, "column") - 1, count_arg(stat_type, "stat_type"), REAL(df)[0],
                             LENGTH(df) > 1 ? REAL(df)[1] : 0.0, REAL(q), REAL(th), err, sizeof err);
    if (rc != RMG_OK) {
        UNPROTECT(2);
        check(rc, err);
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, q);
    SET_VECTOR_ELT(out, 1, th);
    UNPROTECT(3);
    return out;
}

/* The other methods of mincFDR on the resident column: method 0 = FDR / p.adjust, 1 = BY, 2 = pFDR / fastqvalue, 3 = qvalue
 * (R/minc_FDR.R:429-438), pi0 for the last two. */
SEXP rmincgpu_dataset_fdr_method(SEXP h, SEXP column, SEXP stat_type, SEXP df, SEXP method, SEXP pi0)
{
    rmg_dataset *ds = dataset_of(h);
    if (!Rf_isReal(df) || LENGTH(df) < 1) Rf_error("df must be a double vector");
    const R_xlen_t V = (R_xlen_t)rmg_dataset_voxels(ds);
    SEXP q = PROTECT(Rf_allocVector(REALSXP, V));
    SEXP th = PROTECT(Rf_allocVector(REALSXP, 5));
    char err[512] = "";
    int rc = rmg_dataset_fdr_method(ds, count_arg(column, "column") - 1, count_arg(stat_type, "stat_type"), REAL(df)[0],
                                    LENGTH(df) > 1 ? REAL(df)[1] : 0.0, REAL(q), REAL(th), count_arg(method, "method"),
                                    Rf_asReal(pi0), err, sizeof err);
    if (rc != RMG_OK) {
        UNPROTECT(2);
        check(rc, err);
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, q);
    SET_VECTOR_ELT(out, 1, th);
    UNPROTECT(3);
    return out;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * minc2_model itself, method "lm", with the reference's own eleven arguments (src/modelling_functions.c:743-745; R caller
 * mincLm_c_wrapper, R/minc_voxel_statistics.R:695-712): the .Call a maintainer can swap without staging anything in R.
 * The files are decoded by the library's MINC 2 reader on all host cores straight into page-locked memory -- uint16 /
 * float32 studies in their stored type (rmg_lm_fit_stored), other types as the real doubles (rmg_lm_fit), a study with
 * filenames_right as two double matrices (rmg_lm_fit_cov) -- and fitted on options(RMINC_GPU_N) GPUs (default 1).
 * asgn, rho and nresults belong to the other methods of minc2_model, which stay on the reference's CPU path. */
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static int gpu_count_option(void)
{
    SEXP o = Rf_GetOption1(Rf_install("RMINC_GPU_N"));
    if (Rf_isNull(o)) return 1;
    const int g = Rf_asInteger(o);
    if (g == NA_INTEGER || g < 1) Rf_error("options(RMINC_GPU_N) must be a positive number of GPUs");
    return g;
}

/* Page-locking fresh memory runs at ~1.4 GB/s (measured: 0.33 s for 453 MB) -- for a one-off 7 GB study that is 5 s, ten
 * times what the library's pinned ring costs a pageable matrix.  So: page-locked (pooled by the library, free for later calls
 * of the same size) up to 2 GiB or when options(RMINC_GPU_PINNED = TRUE) asks for it, pageable otherwise. */
static void *staging_alloc(size_t bytes, int *pinned)
{
    SEXP o = Rf_GetOption1(Rf_install("RMINC_GPU_PINNED"));
    const int want = (!Rf_isNull(o) && Rf_asLogical(o) == 1) || bytes <= ((size_t)2 << 30);
    void *p = want ? rmg_host_alloc(bytes ? bytes : 1) : NULL;
    *pinned = p != NULL;
    return p ? p : malloc(bytes ? bytes : 1);
}

static void staging_free(void *p, int pinned)
{
    if (!p) return;
    if (pinned) rmg_host_free(p);
    else free(p);
}

SEXP rmincgpu_minc2_model(SEXP filenames, SEXP filenames_right, SEXP mmatrix, SEXP asgn, SEXP have_mask, SEXP mask,
                          SEXP mask_lower, SEXP mask_upper, SEXP rho, SEXP nresults, SEXP method)
{
    (void)asgn; (void)rho; (void)nresults;
    if (TYPEOF(method) != STRSXP || Rf_length(method) < 1 || strcmp(CHAR(STRING_ELT(method, 0)), "lm") != 0)
        Rf_error("rmincgpu_minc2_model runs method \"lm\"; the other methods of minc2_model stay on the CPU path");
    if (TYPEOF(filenames) != STRSXP || Rf_length(filenames) < 1) Rf_error("filenames must be a character vector");
    const int n = Rf_length(filenames);
    const int cov = TYPEOF(filenames_right) == STRSXP;
    if (cov && Rf_length(filenames_right) != n) Rf_error("filenames_right must name one volume per subject");
    const int has_static = !Rf_isLogical(mmatrix);
    if (has_static && (!Rf_isReal(mmatrix) || Rf_nrows(mmatrix) != n)) Rf_error("model matrix does not match the number of files");
    if (!cov && !has_static) Rf_error("mmatrix must be a double model matrix");
    const int ps = has_static ? Rf_ncols(mmatrix) : 0;
    const int p = cov ? (has_static ? ps + 1 : 2) : ps;
    const int masked = Rf_asReal(have_mask) == 1.0;
    if (masked && (TYPEOF(mask) != STRSXP || Rf_length(mask) < 1)) Rf_error("mask must be a file name");
    const int ngpu = gpu_count_option(), device = gpu_device();
    const double lo = Rf_asReal(mask_lower), hi = Rf_asReal(mask_upper);
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (ncpu < 1) ncpu = 1;

    char err[512] = "";
    rmg_minc2_volume *first = NULL;
    rmg_minc2_info info;
    int rc = rmg_minc2_open(CHAR(STRING_ELT(filenames, 0)), &first, err, sizeof err);
    check(rc, err);                                           /* "Error opening input file: <name>." like :809 */
    rmg_minc2_get_info(first, &info);
    rmg_minc2_close(first);
    const int64_t V = info.nvoxels;
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows((R_xlen_t)V), 2 * p + 3));
    const char **paths = (const char **)R_alloc((size_t)n, sizeof(char *));
    const char **paths_right = (const char **)R_alloc((size_t)n, sizeof(char *));
    for (int j = 0; j < n; ++j) {
        paths[j] = CHAR(STRING_ELT(filenames, j));
        paths_right[j] = cov ? CHAR(STRING_ELT(filenames_right, j)) : NULL;
    }
    double *m = NULL;
    if (masked) {                                             /* the mask as doubles, whatever its stored type */
        rmg_minc2_volume *mv = NULL;
        rmg_minc2_info mi;
        rc = rmg_minc2_open(CHAR(STRING_ELT(mask, 0)), &mv, err, sizeof err);
        if (rc == RMG_OK) {
            rmg_minc2_get_info(mv, &mi);
            if (mi.nvoxels != V) {
                rmg_minc2_close(mv);
                UNPROTECT(1);
                Rf_error("Mask Volume input volume dimension mismatch.");
            }
            m = (double *)R_alloc((size_t)(V > 0 ? V : 1), sizeof(double));
            rc = rmg_minc2_read_real(mv, m, V, (int)ncpu, err, sizeof err);
            rmg_minc2_close(mv);
        }
        if (rc != RMG_OK) {
            UNPROTECT(1);
            check(rc, err);
        }
    }
    const int stored = !cov && (info.dtype == RMG_MINC2_U16 || info.dtype == RMG_MINC2_F32);
    const int64_t ld = V > 0 ? V : 1;
    int pin_a = 0, pin_b = 0;
    void *A = NULL, *B = NULL;
    double *rmin = NULL, *rmax = NULL;
    if (stored) {
        const int dt = info.dtype == RMG_MINC2_U16 ? RMG_STORED_U16 : RMG_STORED_F32;
        const size_t es = dt == RMG_STORED_U16 ? 2 : 4;
        const int64_t nsl = info.nslices > 0 ? info.nslices : 1;
        A = staging_alloc((size_t)ld * n * es, &pin_a);
        rmin = (double *)malloc(sizeof(double) * (size_t)nsl * n);
        rmax = (double *)malloc(sizeof(double) * (size_t)nsl * n);
        double vr[2] = {0.0, 1.0};
        if (!A || !rmin || !rmax) { rc = RMG_ERR_INVALID; strcpy(err, "out of host memory while staging the volumes"); }
        else rc = rmg_minc2_read_table(paths, n, A, ld, info.dtype, rmin, rmax, nsl, vr, (int)ncpu, err, sizeof err);
        if (rc == RMG_OK && dt == RMG_STORED_U16) {           /* image-min / image-max -> scale / offset, in place */
            for (int64_t i = 0; i < nsl * n; ++i) rmax[i] = (rmax[i] - rmin[i]) / (vr[1] - vr[0]);
        }
        if (rc == RMG_OK)
            rc = ngpu > 1 ? rmg_lm_fit_stored_multi(A, dt, V, ld, n, rmax, rmin, vr[0], nsl > 1 ? info.slice_len : ld, REAL(mmatrix),
                                                    p, m, lo, hi, REAL(out), ld, ngpu, err, sizeof err)
                          : rmg_lm_fit_stored(A, dt, V, ld, n, rmax, rmin, vr[0], nsl > 1 ? info.slice_len : ld, REAL(mmatrix), p, m,
                                              lo, hi, REAL(out), ld, device, err, sizeof err);
    } else {
        A = staging_alloc((size_t)ld * n * sizeof(double), &pin_a);
        if (cov) B = staging_alloc((size_t)ld * n * sizeof(double), &pin_b);
        if (!A || (cov && !B)) { rc = RMG_ERR_INVALID; strcpy(err, "out of host memory while staging the volumes"); }
        else rc = rmg_minc2_read_table(paths, n, A, ld, RMG_MINC2_F64, NULL, NULL, 1, NULL, (int)ncpu, err, sizeof err);
        if (rc == RMG_OK && cov) rc = rmg_minc2_read_table(paths_right, n, B, ld, RMG_MINC2_F64, NULL, NULL, 1, NULL, (int)ncpu, err, sizeof err);
        if (rc == RMG_OK) {
            if (cov)
                rc = ngpu > 1 ? rmg_lm_fit_cov_multi((const double *)A, (const double *)B, V, ld, n, has_static ? REAL(mmatrix) : NULL,
                                                     ps, m, lo, hi, REAL(out), ld, ngpu, err, sizeof err)
                              : rmg_lm_fit_cov((const double *)A, (const double *)B, V, ld, n, has_static ? REAL(mmatrix) : NULL, ps, m,
                                               lo, hi, REAL(out), ld, device, err, sizeof err);
            else
                rc = ngpu > 1 ? rmg_lm_fit_multi((const double *)A, V, ld, n, REAL(mmatrix), p, m, lo, hi, REAL(out), ld, ngpu, err, sizeof err)
                              : rmg_lm_fit((const double *)A, V, ld, n, REAL(mmatrix), p, m, lo, hi, REAL(out), ld, device, err, sizeof err);
        }
    }
    staging_free(A, pin_a);
    staging_free(B, pin_b);
    free(rmin);
    free(rmax);
    UNPROTECT(1);
    check(rc, err);
    return out;
}


/* per_voxel_anova with the reference's own seven arguments (src/minc_anova.c:142-143; R caller mincAnova,
 * R/minc_voxel_statistics.R:583-600): the files are read here, as the real doubles, and the F-statistic per term comes from
 * rmg_anova_fit.  have_mask is 1 or 0 like minc2_model's. */
SEXP rmincgpu_per_voxel_anova(SEXP filenames, SEXP Sx, SEXP asgn, SEXP have_mask, SEXP mask, SEXP mask_lower, SEXP mask_upper)
{
    if (TYPEOF(filenames) != STRSXP || Rf_length(filenames) < 1) Rf_error("filenames must be a character vector");
    const int n = Rf_length(filenames);
    if (!Rf_isReal(Sx) || !Rf_isInteger(asgn) || Rf_nrows(Sx) != n || LENGTH(asgn) != Rf_ncols(Sx))
        Rf_error("model matrix / assign do not match the files");
    const int p = Rf_ncols(Sx);
    int nterms = 0;
    for (int i = 0; i < p; ++i) if (INTEGER(asgn)[i] > nterms) nterms = INTEGER(asgn)[i];
    const int masked = Rf_asReal(have_mask) == 1.0;
    if (masked && (TYPEOF(mask) != STRSXP || Rf_length(mask) < 1)) Rf_error("mask must be a file name");
    const int device = gpu_device();
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (ncpu < 1) ncpu = 1;
    char err[512] = "";
    rmg_minc2_volume *first = NULL;
    rmg_minc2_info info;
    int rc = rmg_minc2_open(CHAR(STRING_ELT(filenames, 0)), &first, err, sizeof err);
    check(rc, err);
    rmg_minc2_get_info(first, &info);
    rmg_minc2_close(first);
    const int64_t V = info.nvoxels, ld = V > 0 ? V : 1;
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, matrix_rows((R_xlen_t)V), nterms));
    const char **paths = (const char **)R_alloc((size_t)n, sizeof(char *));
    for (int j = 0; j < n; ++j) paths[j] = CHAR(STRING_ELT(filenames, j));
    double *m = NULL;
    if (masked) {
        rmg_minc2_volume *mv = NULL;
        rmg_minc2_info mi;
        rc = rmg_minc2_open(CHAR(STRING_ELT(mask, 0)), &mv, err, sizeof err);
        if (rc == RMG_OK) {
            rmg_minc2_get_info(mv, &mi);
            if (mi.nvoxels != V) {
                rmg_minc2_close(mv);
                UNPROTECT(1);
                Rf_error("Mask Volume input volume dimension mismatch.");
            }
            m = (double *)R_alloc((size_t)ld, sizeof(double));
            rc = rmg_minc2_read_real(mv, m, V, (int)ncpu, err, sizeof err);
            rmg_minc2_close(mv);
        }
        if (rc != RMG_OK) {
            UNPROTECT(1);
            check(rc, err);
        }
    }
    int pinned = 0;
    void *A = staging_alloc((size_t)ld * n * sizeof(double), &pinned);
    if (!A) { rc = RMG_ERR_INVALID; strcpy(err, "out of host memory while staging the volumes"); }
    else rc = rmg_minc2_read_table(paths, n, A, ld, RMG_MINC2_F64, NULL, NULL, 1, NULL, (int)ncpu, err, sizeof err);
    if (rc == RMG_OK && nterms > 0)
        rc = rmg_anova_fit((const double *)A, V, ld, n, REAL(Sx), p, INTEGER(asgn), m, Rf_asReal(mask_lower), Rf_asReal(mask_upper),
                           REAL(out), ld, device, err, sizeof err);
    staging_free(A, pinned);
    UNPROTECT(1);
    check(rc, err);
    return out;
}


/* kernel launches issued by the library since it was loaded: lets a test assert that a call really ran on the GPU */
SEXP rmincgpu_launch_count(void)
{
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 1));
    REAL(out)[0] = (double)rmg_launch_count();
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef callMethods[] = {
    {"rmincgpu_vertex_lm_loop", (DL_FUNC)&rmincgpu_vertex_lm_loop, 3},
    {"rmincgpu_vertex_wlm_loop", (DL_FUNC)&rmincgpu_vertex_wlm_loop, 4},
    {"rmincgpu_lm", (DL_FUNC)&rmincgpu_lm, 6},
    {"rmincgpu_lm_cov", (DL_FUNC)&rmincgpu_lm_cov, 6},
    {"rmincgpu_lm_stored", (DL_FUNC)&rmincgpu_lm_stored, 11},
    {"rmincgpu_randomize", (DL_FUNC)&rmincgpu_randomize, 9},
    {"rmincgpu_randomize_cov", (DL_FUNC)&rmincgpu_randomize_cov, 10},
    {"rmincgpu_randomize_stored", (DL_FUNC)&rmincgpu_randomize_stored, 15},
    {"rmincgpu_anova", (DL_FUNC)&rmincgpu_anova, 6},
    {"rmincgpu_fdr", (DL_FUNC)&rmincgpu_fdr, 4},
    {"rmincgpu_fdr_method", (DL_FUNC)&rmincgpu_fdr_method, 6},
    {"rmincgpu_fdr_pi0", (DL_FUNC)&rmincgpu_fdr_pi0, 5},
    {"rmincgpu_graph_tfce", (DL_FUNC)&rmincgpu_graph_tfce, 7},
    {"rmincgpu_randomize_tfce", (DL_FUNC)&rmincgpu_randomize_tfce, 11},
    {"rmincgpu_volume_tfce", (DL_FUNC)&rmincgpu_volume_tfce, 8},
    {"rmincgpu_randomize_volume_tfce", (DL_FUNC)&rmincgpu_randomize_volume_tfce, 16},
    {"rmincgpu_dataset_create", (DL_FUNC)&rmincgpu_dataset_create, 5},
    {"rmincgpu_dataset_free", (DL_FUNC)&rmincgpu_dataset_free, 1},
    {"rmincgpu_dataset_lm", (DL_FUNC)&rmincgpu_dataset_lm, 2},
    {"rmincgpu_dataset_anova", (DL_FUNC)&rmincgpu_dataset_anova, 3},
    {"rmincgpu_dataset_randomize", (DL_FUNC)&rmincgpu_dataset_randomize, 5},
    {"rmincgpu_dataset_fdr", (DL_FUNC)&rmincgpu_dataset_fdr, 4},
    {"rmincgpu_dataset_fdr_method", (DL_FUNC)&rmincgpu_dataset_fdr_method, 6},
    {"rmincgpu_minc2_model", (DL_FUNC)&rmincgpu_minc2_model, 11},
    {"rmincgpu_per_voxel_anova", (DL_FUNC)&rmincgpu_per_voxel_anova, 7},
    {"rmincgpu_launch_count", (DL_FUNC)&rmincgpu_launch_count, 0},
    {NULL, NULL, 0}};

void R_init_rmincgpu(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, callMethods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
