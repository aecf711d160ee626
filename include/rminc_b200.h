/*
 * rminc_b200.h -- C ABI of the B200-native voxel-wise OLS hot path of RMINC.
 *
 * Plain C, plain pointers and sizes; no R, torch or CUDA types.  Every entry
 * point is what RMINC's own .Call boundary for this path would bind
 * (reference: src/RMINC_init.c:57-82).  All matrices are column-major (R
 * layout).  There is NO CPU fallback: without a CUDA device every compute
 * entry point returns RMG_ERR_NO_DEVICE.
 *
 * Result matrix layout (reference: src/modelling_functions.c:1159-1174,
 * src/vertex_modelling.c:144-157): V rows x (2p+3) columns,
 *   col 0 = F-statistic, col 1 = R-squared, cols 2..p+1 = beta (pivoted order,
 *   zeros after the numerical rank), cols p+2..2p+1 = t-statistics,
 *   col 2p+2 = logLik.
 * Rows outside the mask are +0.0 in every column (reference:
 * R/minc_voxel_statistics.R:898-905, src/minc_anova.c:270-275).
 *
 * Errors: functions return RMG_OK or a code below and write a NUL-terminated
 * message into err (if err != NULL and errlen > 0).  Nothing is thrown or
 * longjmp'd across the boundary; all device and pinned memory allocated inside
 * a call is released before it returns.
 */
#ifndef RMINC_B200_H
#define RMINC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMG_OK 0
#define RMG_ERR_INVALID 1    /* bad argument (NULL pointer, n <= p, p > RMG_MAX_P ...) */
#define RMG_ERR_SINGULAR 2   /* dpotri info > 0: "element (i, i) is zero, so the inverse cannot be
                                computed" -- the reference aborts the whole call with an R error
                                (src/modelling_functions.c:551-557) */
#define RMG_ERR_CUDA 3       /* a CUDA runtime/driver call failed */
#define RMG_ERR_NO_DEVICE 4  /* no CUDA device / device index out of range */
#define RMG_ERR_IO 5         /* a volume file cannot be opened or is not what it should be (rmg_minc2_*) */

#define RMG_MAX_P 32         /* largest supported number of design columns */

/* Library / environment probes (never touch the GPU beyond a device count). */
int rmg_version(void);
int rmg_device_count(void);
int rmg_max_p(void);

/*
 * Host-side factorisation of the shared design X (n x p), done ONCE per call
 * instead of once per voxel.  Replaces the per-voxel F77 dqrls (LINPACK dqrdc2,
 * tol 1e-7, limited column pivoting) and LAPACK dpotri of
 * src/modelling_functions.c:489-495 and :549.
 *   pivot[p]   0-based column permutation, as voxel_lm's pivot[] after dqrls
 *   rank       numerical rank k
 *   Q[n*p]     first k columns: the thin orthonormal factor of X[:, pivot[:k]]
 *   Rinv[p*p]  leading k x k block: inverse of the k x k triangle R_k
 *   cdiag[p]   diag((R'R)^-1) of the full p x p triangle, as dpotri leaves it
 * Returns RMG_ERR_SINGULAR (and *dpotri_info = i, 1-based) when R[i,i] == 0.
 * Any output pointer may be NULL.
 */
int rmg_design_factor(const double *X, int n, int p, int32_t *pivot, int *rank,
                      double *Q, double *Rinv, double *cdiag, int *dpotri_info,
                      char *err, size_t errlen);

/*
 * vertexLm / anatLm.  Drop-in for
 *   SEXP vertex_lm_loop(SEXP data_left, SEXP data_right, SEXP mmatrix)
 * (src/vertex_modelling.c:8; R callers R/minc_vertex_statistics.R:408-414,
 * R/minc_anatomy.R:1144-1150) with a static design (data_right = logical NA).
 *   Y    V x n host matrix, leading dimension ldY >= V   (data_left)
 *   X    n x p host model matrix                          (mmatrix)
 *   out  V x (2p+3) host matrix, leading dimension ldOut >= V
 *   device  CUDA device ordinal
 */
int rmg_vertex_lm(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                  double *out, int64_t ldOut, int device, char *err, size_t errlen);

/*
 * anatLm(weights = w).  Drop-in for
 *   SEXP vertex_wlm_loop(SEXP data_left, SEXP data_right, SEXP mmatrix, SEXP ws)
 * (src/weighted_least_squares.c:168; R caller R/minc_anatomy.R:1135-1142): weighted least squares,
 * per-voxel body voxel_wlm (:14-166).  weights: n values, finite and > 0.  Same result layout.
 */
int rmg_vertex_wlm(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                   const double *weights, double *out, int64_t ldOut, int device, char *err, size_t errlen);

/*
 * mincLm.  Drop-in for minc2_model(..., method = "lm") (src/modelling_functions.c:743-745;
 * R caller mincLm_c_wrapper, R/minc_voxel_statistics.R:695-712) once the R layer has staged
 * the n volumes as the columns of Y (row order = file dimension order, :1059) and the mask
 * volume as a double vector.
 *   mask        NULL (have_mask = 0) or V doubles
 *   mask_lo/hi  voxel v is fitted iff mask[v] > mask_lo - 0.5 && mask[v] < mask_hi + 0.5
 *               (:1063-1066; R defaults 1 and 99999999, R/minc_voxel_statistics.R:781-787)
 * Y may be pageable or pinned host memory.
 */
int rmg_lm_fit(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
               const double *mask, double mask_lo, double mask_hi,
               double *out, int64_t ldOut, int device, char *err, size_t errlen);

/*
 * Same fit with every buffer already resident on `device` (dY, dmask, dout are device
 * pointers).  The work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the
 * legacy default stream) and the call returns without synchronising.
 */
int rmg_lm_fit_device(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p,
                      const double *dmask, double mask_lo, double mask_hi,
                      double *dout, int64_t ldOut, int device, void *stream,
                      char *err, size_t errlen);

/*
 * Voxel-wise covariate: a MINC file (or vertex file) term on the right-hand side of the formula, e.g.
 * mincLm(jacobians ~ group + otherVolume).  The design of voxel v is [Xs | z_v] (p = ps + 1 columns), or
 * [1 | z_v] (p = 2) when there is no static part.  Drop-in for minc2_model(filenames, filenames_right, mmatrix, ...)
 * with filenames_right a file list (src/modelling_functions.c:848-854, :1118-1144; R caller
 * R/minc_voxel_statistics.R:695-712 with parseLmFormula's data.matrix.right, R/minc_interface.R:1266-1300) and for
 *   SEXP vertex_lm_loop(SEXP data_left, SEXP data_right, SEXP mmatrix)
 * with data_right a V x n matrix (src/vertex_modelling.c:33-55, :112-140).
 *   Z     V x n host matrix, same leading dimension as Y          (data_right / filenames_right staged like Y)
 *   Xs    n x ps static model matrix; NULL or ps = 0 means "no static part" (mmatrix = logical NA)
 *   out   V x (2p+3), p = ps + 1 (2 without a static part); the covariate's beta / t are the LAST of each block
 * Per-voxel pivoting as dqrdc2 decides it for the last column: a covariate that is (numerically) in the span of
 * Xs at some voxel gives rank p-1 there: beta = 0 and t = 0 in its slot, the rest is the fit on Xs alone.  A
 * covariate that is exactly zero for every subject at a fitted voxel fails the whole call with RMG_ERR_SINGULAR,
 * as the reference's dpotri check does.  Xs must have full column rank (RMG_ERR_INVALID otherwise), ps <= 16.
 */
int rmg_lm_fit_cov(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                   const double *mask, double mask_lo, double mask_hi,
                   double *out, int64_t ldOut, int device, char *err, size_t errlen);
int rmg_vertex_lm_cov(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                      double *out, int64_t ldOut, int device, char *err, size_t errlen);
/* Device-resident form, enqueued on `stream` without synchronising.  dstatus (nullable) is a device int32 the
 * caller zeroes beforehand: it is raised to p when a fitted voxel's covariate is all zero (those rows are NaN). */
int rmg_lm_fit_cov_device(const double *dY, const double *dZ, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                          const double *dmask, double mask_lo, double mask_hi, double *dout, int64_t ldOut,
                          int32_t *dstatus, int device, void *stream, char *err, size_t errlen);

/*
 * mincLm on volumes kept in their STORED type.  minc2_model reads every slab with
 * miget_real_value_hyperslab(..., MI_TYPE_DOUBLE, ...) (src/modelling_functions.c:1037-1055; mincGetVolume:
 * src/minc_reader.c:294-331): libminc (BIC-MNI/libminc, pinned at b01ba22 by inst/tools/ch_minc_depends.m4:85)
 * expands the stored voxels to doubles on the CPU.  Here the R layer hands over the stored samples themselves
 * (miget_voxel_value_hyperslab with the file's own type) plus the real-range tables, 2 or 4 bytes per sample
 * over PCIe and HBM instead of 8, and the kernel forms
 *     y = ((double)u - voxel_min) * scale[slice, subject] + offset[slice, subject]
 * in registers with three correctly rounded FP64 operations (no fused multiply-add), i.e. the same doubles the
 * CPU expansion yields; the fit is then rmg_lm_fit's.
 *   U          V x n samples, leading dimension ldU (in samples); dtype RMG_STORED_U16 or RMG_STORED_F32
 *   scale, offset   [nslices x n] doubles, slice fastest, nslices = ceil(V / slice_len); from MINC's per-slice
 *              image-min / image-max and the file's valid range:
 *              scale = (image_max - image_min) / (valid_max - valid_min), offset = image_min.
 *              Global scaling: slice_len = V (one row).  Ignored (may be NULL) for RMG_STORED_F32, which
 *              converts exactly.
 *   voxel_min  valid_min of the stored type; must be integral
 *   slice_len  voxels per slice of the slowest file dimension (sizes[1] * sizes[2])
 */
#define RMG_STORED_U16 1
#define RMG_STORED_F32 2
int rmg_lm_fit_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n,
                      const double *scale, const double *offset, double voxel_min, int64_t slice_len,
                      const double *X, int p, const double *mask, double mask_lo, double mask_hi,
                      double *out, int64_t ldOut, int device, char *err, size_t errlen);
/* Device-resident form (dU, dscale, doffset, dmask, dout on `device`), enqueued on `stream`.  dU may be a voxel
 * shard of a larger volume: v_offset is the global index of its first voxel (slice = (v_offset + v) / slice_len) and
 * nslices the number of rows of dscale / doffset (<= 0: ceil((v_offset + V) / slice_len)). */
int rmg_lm_fit_stored_device(const void *dU, int dtype, int64_t V, int64_t ldU, int n,
                             const double *dscale, const double *doffset, double voxel_min, int64_t slice_len,
                             int64_t nslices, int64_t v_offset, const double *X, int p, const double *dmask, double mask_lo, double mask_hi,
                             double *dout, int64_t ldOut, int device, void *stream, char *err, size_t errlen);

/*
 * mincAnova / vertexAnova / anatAnova.  Drop-in for per_voxel_anova (src/minc_anova.c:142-304;
 * R caller R/minc_voxel_statistics.R:557-566) and vertex_anova_loop (src/vertex_modelling.c:178;
 * R callers R/minc_vertex_statistics.R:314-320, R/minc_anatomy.R anatAnova), whose per-voxel body is
 * voxel_anova (src/minc_anova.c:4-80): sequential F-statistics, one per model term.
 *   asgn[p]   model.matrix's "assign" attribute (term index of every design column, 0 = intercept)
 *   out       V x max(asgn) matrix: F_t = (sum of effects^2 over the term's columns / df_t) / (RSS / (n-p))
 * Rows outside the mask are 0.0 (src/minc_anova.c:270-275).  The model matrix must have full
 * column rank (RMG_ERR_INVALID otherwise).
 */
int rmg_anova_fit(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const int32_t *asgn,
                  const double *mask, double mask_lo, double mask_hi,
                  double *out, int64_t ldOut, int device, char *err, size_t errlen);
int rmg_anova_fit_device(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p,
                         const int32_t *asgn, const double *dmask, double mask_lo, double mask_hi,
                         double *dout, int64_t ldOut, int device, void *stream, char *err, size_t errlen);

/*
 * mincFDR / vertexFDR / anatFDR for one result column (mincFDR.mincMultiDim, R/minc_FDR.R:254-520).
 *   stat        V statistics (a tvalue- or F-statistic column of the fit)
 *   stat_type   RMG_STAT_T: p = pt2(t, df1) = 2 pt(-|t|, df1) (two-sided, :392-397)
 *               RMG_STAT_F: p = pf(F, df1, df2, lower.tail = FALSE) (:398-412)
 *   mask        NULL or V doubles; voxels with mask > 0.5 take part (:394)
 *   qvalues     V doubles: p.adjust(p, "fdr") (Benjamini-Hochberg) over the participating voxels, NaN where the
 *               statistic is NaN (such voxels do not count), 1 outside the mask (:369, :505)
 *   thresholds  5 doubles (nullable): the statistic at q <= 0.01, 0.05, 0.10, 0.15, 0.20, i.e.
 *               qt(max p / 2, df1, lower.tail = FALSE) or qf(max p, df1, df2, lower.tail = FALSE) of the largest
 *               accepted p-value; NaN when no voxel passes (:444-475)
 * pt / pf live in libR (nmath), not in the reference tree: restated through the regularised incomplete beta function.
 */
#define RMG_STAT_T 1
#define RMG_STAT_F 2
int rmg_fdr(const double *stat, int64_t V, int stat_type, double df1, double df2, const double *mask,
            double *qvalues, double *thresholds, int device, char *err, size_t errlen);
/* dstat / dmask / dq on `device`, work enqueued on `stream`; synchronises the stream before it returns (the
 * thresholds are host values). */
int rmg_fdr_device(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask,
                   double *dq, double *thresholds, int device, void *stream, char *err, size_t errlen);
/*
 * The other q-value methods of mincFDR (R/minc_FDR.R:429-438).  Same arguments as rmg_fdr plus
 *   method   RMG_FDR_BH      "FDR" / "p.adjust": p.adjust(p, "fdr")                                  (:432-434)
 *            RMG_FDR_BY      "BY": p.adjust(p, "BY") = pmin(1, cummin(c n / i p)), c = sum(1 / (1:n))   (:435-437)
 *            RMG_FDR_STOREY  "pFDR" / "fastqvalue": fast.qvalue (R/minc_misc.R:366-560): pi0 m p / #{p' <= p},
 *                            then the reference's qvalue_min (src/modelling_functions.c:15-46) exactly as written: a
 *                            minimum with the NEXT-ranked raw value (not a running minimum); the largest p keeps its
 *                            raw value unless the raw value of the element after it in the masked vector exceeds 1
 *                            (the reference's 1-based index used 0-based), then it becomes 1
 *            RMG_FDR_QVALUE  "qvalue": qvalue::qvalue (third party, Suggests: of the reference's DESCRIPTION; its
 *                            published form since 2.0): pi0 pmin(1, cummin(m / i p))
 *   pi0      for _STOREY / _QVALUE: the estimated proportion of true nulls, in (0, 1] (else RMG_ERR_INVALID with the
 *            reference's message).  Its default estimator is stats::smooth.spline (R, third party) over
 *            mean(p >= lambda) / (1 - lambda); rmg_fdr_pi0_fractions supplies the means, the caller smooths.
 * _STOREY and _QVALUE fail with RMG_ERR_INVALID when a selected voxel has a NaN statistic (the reference's
 * `if (min(p) < 0 || max(p) > 1)` is R's "missing value where TRUE/FALSE needed").  At most 2^31 - 1 voxels per call.
 */
#define RMG_FDR_BH 0
#define RMG_FDR_BY 1
#define RMG_FDR_STOREY 2
#define RMG_FDR_QVALUE 3
int rmg_fdr_method(const double *stat, int64_t V, int stat_type, double df1, double df2, const double *mask,
                   double *qvalues, double *thresholds, int method, double pi0, int device, char *err, size_t errlen);
int rmg_fdr_method_device(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask,
                          double *dq, double *thresholds, int method, double pi0, int device, void *stream, char *err,
                          size_t errlen);
/* fractions[i] = mean(p >= lambda[i]) over the participating voxels, 1 <= nlambda <= 64, lambda in [0, 1)
 * (R/minc_misc.R:439-441); lambda and fractions are host arrays in both forms. */
int rmg_fdr_pi0_fractions(const double *stat, int64_t V, int stat_type, double df1, double df2, const double *mask,
                          const double *lambda, int nlambda, double *fractions, int device, char *err, size_t errlen);
int rmg_fdr_pi0_fractions_device(const double *dstat, int64_t V, int stat_type, double df1, double df2,
                                 const double *dmask, const double *lambda, int nlambda, double *fractions, int device,
                                 void *stream, char *err, size_t errlen);
/* p-value of one statistic, evaluated on the host by the same code the device runs. */
double rmg_pvalue(double stat, int stat_type, double df1, double df2);

/*
 * mincRandomize.  Replaces the R loop of R/minc_voxel_statistics.R:1465-1497, which re-runs
 * the whole mincLm call once per permutation.
 *   idx      n x R, 0-based row indices drawn by the caller (R's sample()); permutation r
 *            refits with X_pi = X[idx[, r], ] and the same Y and mask
 *   cols     0-based columns of the V x (2p+3) result whose maxima are wanted (default in R:
 *            the tvalue- columns)
 *   two_sided  non-zero: abs() before the maximum (alternative = "two.sided")
 *   dist     R x ncols host matrix (column-major): per permutation, the maximum over ALL V
 *            rows of the statistic with non-finite values and masked rows counted as 0
 * A permuted design whose R factor has an exactly-zero diagonal fails the whole call with
 * RMG_ERR_SINGULAR, as the reference's refit would.
 */
int rmg_randomize(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                  const double *mask, double mask_lo, double mask_hi,
                  const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                  double *dist, int device, char *err, size_t errlen);
/*
 * The same permutation test on volumes in their stored type (arguments U .. slice_len as rmg_lm_fit_stored): the
 * samples cross the host link as stored (2 or 4 bytes each) and are expanded once on the device into the doubles
 * libminc's reader hands minc2_model (src/modelling_functions.c:1037-1055), which stay resident for all R
 * permutations; the reference re-reads and re-converts every file in every permutation
 * (R/minc_voxel_statistics.R:1477).  Both host forms upload Y in voxel blocks and start the permutations of a block as
 * soon as it has landed; pageable input is staged through pinned buffers by all host threads.
 */
int rmg_randomize_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n,
                         const double *scale, const double *offset, double voxel_min, int64_t slice_len,
                         const double *X, int p, const double *mask, double mask_lo, double mask_hi,
                         const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                         double *dist, int device, char *err, size_t errlen);

/*
 * vertexTFCE: threshold-free cluster enhancement on an adjacency graph.  Drop-in for graph_tfce_wqu / graph_tfce
 * (src/vertex_tfce.cpp:92-157; R callers vertexTFCE.numeric / vertexTFCE.matrix, R/minc_vertex_statistics.R:738-887),
 * batched over maps.
 *   maps      V x nmaps (column-major: one map per column), out likewise
 *   adj_ptr / adj_idx   CSR form of the 0-based adjacency list vertexTFCE takes (adj_ptr[V+1], adj_idx[adj_ptr[V]])
 *                       The list must be symmetric (u in adj(v) <=> v in adj(u)), as an undirected surface graph is;
 *                       a one-directional edge is RMG_ERR_INVALID (the reference would follow it from the higher-valued
 *                       end only, src/vertex_tfce.cpp:104-118)
 *   weights   V vertex areas, NULL = ones (R/minc_vertex_statistics.R:752-758)
 *   E, H, nsteps   extent / height exponents and number of threshold steps (defaults 0.5, 2, 100); nsteps <= 253
 *   side      0 "both" (tfce(x) - tfce(-x)), 1 "positive", 2 "negative" (-tfce(-x))   (:788-796)
 */
int rmg_graph_tfce(const double *maps, int64_t V, int nmaps, const int32_t *adj_ptr, const int32_t *adj_idx,
                   const double *weights, double E, double H, int nsteps, int side, double *out, int device,
                   char *err, size_t errlen);
/*
 * vertexTFCE.vertexLm's permutation test (R/minc_vertex_statistics.R:892-948 -> mincRandomize_core with
 * post_proc = vertexTFCE.matrix): per permutation refit, non-finite -> 0, graph TFCE of every column in `cols`,
 * abs when two_sided, column maximum.  dist: R x ncols.  No mask (vertexLm has none).
 */
int rmg_randomize_tfce(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                       const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                       const int32_t *adj_ptr, const int32_t *adj_idx, const double *weights,
                       double E, double H, int nsteps, int side, double *dist, int device, char *err, size_t errlen);
/* The same over n_gpus devices of the box: the R permutations are split into contiguous blocks, one host thread per GPU,
 * no exchange (replaces mincRandomize_core's groupingVector(R, n_groups) + mclapply, R/minc_voxel_statistics.R:1503-1509). */
int rmg_randomize_tfce_multi(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                             const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                             const int32_t *adj_ptr, const int32_t *adj_idx, const double *weights,
                             double E, double H, int nsteps, int side, double *dist, int n_gpus, char *err, size_t errlen);

/*
 * RMINC:::neighbour_list(x, y, z, n) (src/vertex_tfce.cpp:178-214): the adjacency of a regular grid in CSR form, vertex
 * = k*y*x + j*x + i (first argument fastest), n = 6, 18 or 26.  Host only (no GPU needed).  Call with adj_ptr = adj_idx
 * = NULL to get the number of entries in *nnz, then with adj_ptr[x*y*z + 1] and adj_idx[capacity >= nnz].
 */
int rmg_neighbour_list(int x, int y, int z, int n, int32_t *adj_ptr, int32_t *adj_idx, int64_t capacity, int64_t *nnz,
                       char *err, size_t errlen);
/*
 * RMINC:::mesh_area(vertices, triangles) (src/vertex_tfce.cpp:221-257; R/RcppExports.R:52): the per-vertex areas that
 * vertexTFCE uses as weights when the surface is a bic_obj and no weights are given (R/minc_vertex_statistics.R:752-754).
 *   vertices   3 x nvertices column-major (bic_obj$vertex_matrix), triangles 3 x ntriangles, 1-based vertex numbers as
 *              doubles (bic_obj$triangle_matrix);  areas: nvertices doubles.
 * A literal restatement, including the reference's first cross-product component (d1y*d2z - d1z*d2x).  Host only.
 */
int rmg_mesh_area(const double *vertices, int64_t nvertices, const double *triangles, int64_t ntriangles, double *areas,
                  char *err, size_t errlen);

/*
 * mincTFCE on volumes.  The reference writes the map to a file and shells out to the external `TFCE` program of
 * minc-stuffs (mincTFCE.mincSingleDim, R/minc_voxel_statistics.R:1186-1243), which is NOT in the reference tree; the
 * enhancement here is the tree's own algorithm (graph_tfce, src/vertex_tfce.cpp:92-157) run on the voxel grid, with the
 * arguments mincTFCE passes on: height step d (instead of nsteps: thresholds d, 2d, ... up to the map maximum), E, H
 * and the side.  Parity with the external program is therefore unpinned.
 *   maps      V x nmaps, V = dims[0]*dims[1]*dims[2], row = v0*dims[1]*dims[2] + v1*dims[2] + v2 (mincLm's row order)
 *   connectivity  6 (faces), 18 or 26;  voxel_volume  the extent of one voxel (cluster extent = voxels * volume)
 *   side      0 "both", 1 "positive", 2 "negative"
 * At most 253 thresholds per map (RMG_ERR_INVALID otherwise: increase d).
 */
int rmg_volume_tfce(const double *maps, const int64_t *dims, int nmaps, int connectivity, double voxel_volume,
                    double d, double E, double H, int side, double *out, int device, char *err, size_t errlen);
/*
 * mincTFCE.mincLm's permutation test (R/minc_voxel_statistics.R:1382-1439 -> mincRandomize_core with
 * post_proc = mincTFCE.matrix): per permutation refit (masked voxels 0), non-finite -> 0, volume TFCE of every column in
 * `cols`, abs when two_sided, column maximum.  dist: R x ncols.  n_gpus > 1 splits the R permutations into contiguous
 * blocks over devices 0..n_gpus-1 (one host thread per GPU, no exchange); otherwise device 0.
 */
int rmg_randomize_volume_tfce(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                              const double *mask, double mask_lo, double mask_hi,
                              const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                              const int64_t *dims, int connectivity, double voxel_volume,
                              double d, double E, double H, int side, double *dist, int n_gpus, char *err, size_t errlen);

/*
 * Device-resident variant: dY/dmask on `device`; ddist is a device R x ncols matrix that
 * receives this device's maxima over ITS voxels (voxel shards on several GPUs are combined by
 * the caller with an element-wise max, e.g. ncclAllReduce(ncclMax)).  Enqueued on `stream`.
 */
int rmg_randomize_device(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p,
                         const double *dmask, double mask_lo, double mask_hi,
                         const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                         double *ddist, int device, void *stream, char *err, size_t errlen);

/*
 * Single-process multi-GPU forms (R is one process): contiguous voxel shards over devices
 * 0..n_gpus-1, one host thread per GPU, replacing create_parallel_mask + mclapply
 * (R/minc_voxel_statistics.R:848-906, :1503-1509).  n_gpus <= 0 means all visible devices.
 */
int rmg_lm_fit_multi(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                     const double *mask, double mask_lo, double mask_hi,
                     double *out, int64_t ldOut, int n_gpus, char *err, size_t errlen);
int rmg_lm_fit_cov_multi(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                         const double *mask, double mask_lo, double mask_hi,
                         double *out, int64_t ldOut, int n_gpus, char *err, size_t errlen);
int rmg_lm_fit_stored_multi(const void *U, int dtype, int64_t V, int64_t ldU, int n,
                            const double *scale, const double *offset, double voxel_min, int64_t slice_len,
                            const double *X, int p, const double *mask, double mask_lo, double mask_hi,
                            double *out, int64_t ldOut, int n_gpus, char *err, size_t errlen);
/*
 * mincRandomize on a model with a voxel-wise covariate (y ~ static terms + a column of files).  The reference's loop
 * re-evaluates the whole mincLm call with the predictor rows of the data frame re-sampled -- the covariate files move
 * with the static columns (R/minc_voxel_statistics.R:1468-1477; src/modelling_functions.c:848-854, 1118-1144) -- so
 * permutation r fits y_v on [Xs | z_v][idx[, r], ] at every voxel.
 *   Y, Z     V x n, same leading dimension ldY (as rmg_lm_fit_cov); Xs n x ps static part (NULL / 0: the intercept)
 *   cols     0-based columns of the 2 (ps + 1) + 3 column result (the covariate is the last coefficient)
 *   dist     R x ncols column-major maxima, non-finite -> 0, abs when two_sided (as rmg_randomize)
 * Errors: a rank-deficient permuted static part is RMG_ERR_INVALID (as rmg_lm_fit_cov); an all-zero (re-sampled)
 * covariate at a fitted voxel is the reference's dpotri error, RMG_ERR_SINGULAR.
 * One streaming pass over Y and Z per permutation; _multi shards the voxels over n_gpus devices (host max-merge).
 */
int rmg_randomize_cov(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                      const double *mask, double mask_lo, double mask_hi,
                      const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                      double *dist, int device, char *err, size_t errlen);
int rmg_randomize_cov_multi(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                            const double *mask, double mask_lo, double mask_hi,
                            const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                            double *dist, int n_gpus, char *err, size_t errlen);
int rmg_randomize_multi(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                        const double *mask, double mask_lo, double mask_hi,
                        const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                        double *dist, int n_gpus, char *err, size_t errlen);
int rmg_randomize_stored_multi(const void *U, int dtype, int64_t V, int64_t ldU, int n,
                               const double *scale, const double *offset, double voxel_min, int64_t slice_len,
                               const double *X, int p, const double *mask, double mask_lo, double mask_hi,
                               const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                               double *dist, int n_gpus, char *err, size_t errlen);

/*
 * The resident dataset.  mincLm stores `data` and the call so that mincRandomize, mincAnova and mincFDR can be run on
 * the same files later (R/minc_voxel_statistics.R:911-912, :1296-1297, :1457-1477); the reference re-reads every file
 * each time.  Here the staged subjects x voxels matrix is uploaded ONCE, sharded over n_gpus devices of the box in
 * contiguous voxel ranges inside this one process (create_parallel_mask + mclapply, :848-906), and stays in HBM behind
 * an opaque handle until rmg_dataset_free.  Every routine below runs on the resident shards; the only exchange of the
 * path is the element-wise max of the per-shard R x ncols permutation maxima.
 *   rmg_dataset_create          Y V x n host doubles (pinned or pageable), mask / mask_lo / mask_hi as rmg_lm_fit;
 *                               n_gpus <= 0: every visible device.  The upload is queued in voxel blocks and the call
 *                               returns while it is still in flight (for pinned Y); consumers wait block by block.
 *   rmg_dataset_create_stored   the volumes in their stored type (arguments as rmg_lm_fit_stored); the samples cross the
 *                               host link as stored and are expanded once on the device into the resident doubles.
 *   rmg_dataset_sync            waits until the whole matrix is resident.
 *   rmg_dataset_lm_fit / _anova the fit of rmg_lm_fit / rmg_anova_fit; out (host, V x ncol, may be NULL) receives the
 *                               result, which ALSO stays resident for rmg_dataset_fdr.
 *   rmg_dataset_randomize       rmg_randomize on the resident matrix (the per-voxel sums every permutation shares are
 *                               kept on the handle, so a second test pays no pass over Y for them).
 *   rmg_dataset_fdr             rmg_fdr on column `column` of the resident result of the last fit, with the dataset's mask.
 */
typedef struct rmg_dataset rmg_dataset;
int rmg_dataset_create(const double *Y, int64_t V, int64_t ldY, int n, const double *mask, double mask_lo, double mask_hi,
                       int n_gpus, rmg_dataset **ds, char *err, size_t errlen);
int rmg_dataset_create_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale, const double *offset,
                              double voxel_min, int64_t slice_len, const double *mask, double mask_lo, double mask_hi,
                              int n_gpus, rmg_dataset **ds, char *err, size_t errlen);
void rmg_dataset_free(rmg_dataset *ds);
int64_t rmg_dataset_voxels(const rmg_dataset *ds);
int rmg_dataset_subjects(const rmg_dataset *ds);
int rmg_dataset_gpus(const rmg_dataset *ds);
int rmg_dataset_sync(rmg_dataset *ds, char *err, size_t errlen);
int rmg_dataset_lm_fit(rmg_dataset *ds, const double *X, int p, double *out, int64_t ldOut, char *err, size_t errlen);
int rmg_dataset_anova(rmg_dataset *ds, const double *X, int p, const int32_t *asgn, double *out, int64_t ldOut,
                      char *err, size_t errlen);
int rmg_dataset_randomize(rmg_dataset *ds, const double *X, int p, const int32_t *idx, int R, const int32_t *cols, int ncols,
                          int two_sided, double *dist, char *err, size_t errlen);
int rmg_dataset_fdr(rmg_dataset *ds, int column, int stat_type, double df1, double df2, double *qvalues, double *thresholds,
                    char *err, size_t errlen);
/* The other methods of mincFDR (R/minc_FDR.R:429-438; RMG_FDR_* and pi0 as rmg_fdr_method) and fast.qvalue's pi0 grid
 * (mean(p >= lambda[i]), as rmg_fdr_pi0_fractions) on a column of the resident result. */
int rmg_dataset_fdr_method(rmg_dataset *ds, int column, int stat_type, double df1, double df2, double *qvalues, double *thresholds,
                           int method, double pi0, char *err, size_t errlen);
int rmg_dataset_fdr_pi0_fractions(rmg_dataset *ds, int column, int stat_type, double df1, double df2, const double *lambda,
                                  int nlambda, double *fractions, char *err, size_t errlen);
/* Pinned host staging buffers (the ring of pageable uploads, the packed permutation operand) are kept across calls;
 * this returns every idle one to the OS. */
void rmg_release_host_buffers(void);
/* Page-locked host memory with its pages interleaved over the NUMA nodes the process may allocate on: where a host
 * layer should stage Y (mincTable's matrix, R/minc_interface.R:1031-1077) so that every GPU's copy engine reads it
 * at the link rate.  NULL without a CUDA device or when the allocation fails. */
void *rmg_host_alloc(size_t bytes);
void rmg_host_free(void *ptr);
/* Concurrent pinned host -> device copy rate of the box: device g copies bytes_per_gpu bytes starting at
 * host + g * bytes_per_gpu, all n_gpus devices at once, `repeats` times.  gbs[g] (nullable): GB/s seen by device g;
 * *aggregate_gbs (nullable): all bytes / wall time.  The ceiling end-to-end numbers are reported against. */
int rmg_h2d_probe(const void *host, size_t bytes_per_gpu, int n_gpus, int repeats, double *gbs, double *aggregate_gbs,
                  char *err, size_t errlen);

/*
 * MINC 2 volumes without libminc / libhdf5 (host only; no GPU involved).  Replaces, for the files in front of the hot
 * path, miopen_volume + miget_volume_dimensions + miget_dimension_sizes + miget_real_value_hyperslab as minc2_model and
 * mincGetVolume use them (src/modelling_functions.c:806-829, 1037-1055; src/minc_reader.c:294-331;
 * R/minc_interface.R:1031-1077) -- and hands out the voxels in their STORED type with the per-slice real ranges, which is
 * what rmg_lm_fit_stored / rmg_randomize_stored take.  The HDF5 container is parsed here (rminc_b200/csrc/minc2_reader.cpp);
 * chunks are inflated by n_threads host threads directly into the caller's buffer.
 *   real value = ((double)stored - valid_range[0]) * (max_s - min_s) / (valid_range[1] - valid_range[0]) + min_s
 * with s the range entry of the voxel (entry = voxel index / slice_len); float volumes are taken as they are.
 * Errors: RMG_ERR_IO with the reference's "Error opening input file: <path>." for a missing file, or a description of
 * what in the file is unsupported / damaged.
 */
#define RMG_MINC2_MAX_DIMS 5
#define RMG_MINC2_U8 1
#define RMG_MINC2_I8 2
#define RMG_MINC2_U16 3      /* == the stored-type code rmg_lm_fit_stored takes for uint16 voxels is RMG_STORED_U16 */
#define RMG_MINC2_I16 4
#define RMG_MINC2_U32 5
#define RMG_MINC2_I32 6
#define RMG_MINC2_F32 7
#define RMG_MINC2_F64 8
typedef struct rmg_minc2_volume rmg_minc2_volume;
typedef struct {
    int32_t ndims;                                  /* of /minc-2.0/image/0/image, slowest dimension first (file order) */
    int32_t dtype;                                  /* RMG_MINC2_* */
    int32_t elem_size;                              /* bytes per stored voxel */
    int32_t has_real_range;                         /* image-min / image-max present */
    int64_t sizes[RMG_MINC2_MAX_DIMS];
    int64_t nvoxels;
    int64_t nslices;                                /* entries of image-min / image-max (1 = one range for the volume) */
    int64_t slice_len;                              /* voxels per range entry = nvoxels / nslices */
    double valid_range[2];
    double starts[RMG_MINC2_MAX_DIMS];
    double steps[RMG_MINC2_MAX_DIMS];
    double dircos[RMG_MINC2_MAX_DIMS][3];
    char dimnames[RMG_MINC2_MAX_DIMS][16];          /* "zspace", "yspace", "xspace", ... from the image's dimorder */
} rmg_minc2_info;
int rmg_minc2_open(const char *path, rmg_minc2_volume **vol, char *err, size_t errlen);
int rmg_minc2_get_info(const rmg_minc2_volume *vol, rmg_minc2_info *info);
/* image_min / image_max: info.nslices doubles each */
int rmg_minc2_read_ranges(const rmg_minc2_volume *vol, double *image_min, double *image_max, char *err, size_t errlen);
/* the whole image in its stored type (little-endian), file order; dst_bytes >= nvoxels * elem_size */
int rmg_minc2_read_stored(const rmg_minc2_volume *vol, void *dst, int64_t dst_bytes, int n_threads, char *err, size_t errlen);
/* the doubles miget_real_value_hyperslab(MI_TYPE_DOUBLE) yields for the whole image; count >= nvoxels */
int rmg_minc2_read_real(const rmg_minc2_volume *vol, double *dst, int64_t count, int n_threads, char *err, size_t errlen);
void rmg_minc2_close(rmg_minc2_volume *vol);
/*
 * mincTable in one call (R/minc_interface.R:1031-1077): column j of U (leading dimension ldU >= voxels per volume) receives
 * file paths[j], files decoded in parallel by n_threads host threads.  dtype RMG_MINC2_U16 / RMG_MINC2_F32: the stored voxels
 * (every file must have that type) and, for uint16, image_min / image_max [nslices x n, column-major] (a file with a single
 * range fills its column with it) -- the operands of rmg_lm_fit_stored after scale = (max - min) / (valid_range[1] -
 * valid_range[0]), offset = min; dtype RMG_MINC2_F64: the real values, any stored type.  U may be pinned memory
 * (rmg_host_alloc).  valid_range (nullable): the files' common valid range.
 */
int rmg_minc2_read_table(const char *const *paths, int n, void *U, int64_t ldU, int dtype, double *image_min, double *image_max,
                         int64_t nslices, double *valid_range, int n_threads, char *err, size_t errlen);
/* The container underneath, for header look-ups (mincinfo-style) and for checking the parser against files written by the
 * HDF5 library: a dataset by its path below the root, raw elements in file order (dst nullable: shape only; dims holds up to 8
 * entries; type_class 1 = unsigned, 2 = signed integer, 3 = float, 4 = string, 0 = other), and one attribute of an object
 * (numbers for numeric attributes, text for fixed-length strings; *count = how many values the attribute holds). */
int rmg_h5_read_dataset(const char *path, const char *dataset, void *dst, int64_t dst_bytes, int32_t *rank, int64_t *dims,
                        int32_t *elem_size, int32_t *type_class, int n_threads, char *err, size_t errlen);
int rmg_h5_read_attribute(const char *path, const char *object, const char *attribute, double *numbers, int64_t max_numbers,
                          int64_t *count, char *text, size_t textlen, char *err, size_t errlen);
/* One zlib stream (RFC 1950: what HDF5's deflate filter stores per chunk) -> dst; returns the bytes written or -1 (malformed or
 * truncated stream, wrong Adler-32, dst too small).  use_zlib = 0: the reader's own decoder (csrc/inflate_fast.hpp: 64-bit bit
 * buffer, two-level tables, three literals per refill -- 1.4-1.5x zlib on volume data); 1: zlib's uncompress, its twin and back-up. */
int64_t rmg_zlib_uncompress(const void *src, int64_t src_bytes, void *dst, int64_t dst_bytes, int use_zlib);

/* Timing of the most recent rmg_*_device / host call on this thread, for bench.py:
 * milliseconds spent in the dominant kernel(s), measured with CUDA events on the stream the
 * kernels were launched on (host entry points only; 0 if unavailable). */
double rmg_last_kernel_ms(void);
/* Which fit kernel the most recent fit on this thread launched: 1 = TMA-staged ring,
 * 2 = direct streaming loads (+ split-n), 3 = voxel-wise covariate kernel,
 * 4 = stored-type kernel (direct loads), 5 = stored-type kernel (TMA ring), 0 = none yet. */
int rmg_last_fit_variant(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t rmg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RMINC_B200_H */
