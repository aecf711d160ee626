/*
 * rminc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of RMINC's voxel-wise ordinary
 * least-squares hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; nothing under
 * rminc_b200/ imports, links or executes it.
 *
 * Parity pin: the reference ships no golden vectors for this path (its tests
 * download data at run time and compare row 1 with base-R lm()).  This file is
 * pinned three ways (tests/test_oracle.py):
 *   (1) against the reference's OWN voxel_lm / vertex_lm_loop / minc2_model
 *       compiled from /root/reference/src into oracle/_ref (see oracle/Makefile;
 *       R's C API, libminc and LINPACK dqrls are absent from the image and are
 *       stood in by oracle/rstub -- the arithmetic in voxel_lm is the
 *       reference's own object code),
 *   (2) against published base-R `summary(lm())` / `logLik` outputs
 *       (tests/golden/r_lm_known_answers.json) -- the same oracle the
 *       reference's testthat files use,
 *   (3) against an independent numpy/scipy closed-form OLS.
 *
 * Third-party arithmetic restated here (sources are NOT in /root/reference):
 *   R >= 4.0 (CI pins R 4.4, Dockerfile:1) src/appl/dqrls.f, dqrdc2.f, dqrsl.f
 *   (R's modified LINPACK) and reference BLAS dnrm2/ddot/daxpy/dscal;
 *   LAPACK 3.x dpotri = dtrtri (dtrti2) + dlauum (dlauu2), unblocked because
 *   p <= 64.  Call sites: src/modelling_functions.c:494-495, :549.
 *
 * Next-row restatements in this file and their pins:
 *   orc_anova_loop / orc_vertex_wlm_loop / orc_lm_loop_cov: bit for bit against the reference's object code
 *       (vertex_anova_loop, per_voxel_anova, vertex_wlm_loop, vertex_lm_loop with data_right) and golden vectors.
 *   orc_expand_u16 / orc_expand_f32 (stored voxels -> doubles): libminc (BIC-MNI/libminc @ b01ba22,
 *       inst/tools/ch_minc_depends.m4:85) is a third-party dependency absent from the image; its published
 *       real-range conversion is restated -- PARITY UNPINNED against libminc's object code.
 *   mincFDR (oracle.py: p_adjust_bh, fdr): R's p.adjust restated from its published source; pt / pf come from
 *       scipy -- PARITY UNPINNED against R's nmath beyond textbook critical values.
 *
 * All matrices are column-major, as in R.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ BLAS-1 */

/* Reference-BLAS dnrm2 (scaled sum of squares, the 1993 Sven Hammarling form
 * shipped with R's bundled BLAS). */
double orc_dnrm2(int n, const double *x, int incx)
{
    if (n < 1 || incx < 1) return 0.0;
    if (n == 1) return fabs(x[0]);
    double scale = 0.0, ssq = 1.0;
    for (int ix = 0; ix < n * incx; ix += incx) {
        if (x[ix] != 0.0) {
            double absxi = fabs(x[ix]);
            if (scale < absxi) {
                double r = scale / absxi;
                ssq = 1.0 + ssq * r * r;
                scale = absxi;
            } else {
                double r = absxi / scale;
                ssq += r * r;
            }
        }
    }
    return scale * sqrt(ssq);
}

static double orc_ddot(int n, const double *x, const double *y)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

static void orc_daxpy(int n, double a, const double *x, double *y)
{
    if (a == 0.0) return;
    for (int i = 0; i < n; ++i) y[i] += a * x[i];
}

static void orc_dscal(int n, double a, double *x)
{
    for (int i = 0; i < n; ++i) x[i] *= a;
}

/* --------------------------------------------------------------- dqrdc2.f */
/* R's Householder QR with limited column pivoting (columns whose norm falls
 * below tol * original norm are cycled to the end).  1-based Fortran indices
 * are kept in comments; arrays are 0-based here.  work is p x 2. */
#define X_(i, j) x[(size_t)(i) + (size_t)(j) * (size_t)ldx]
void orc_dqrdc2(double *x, int ldx, int n, int p, double tol, int *k_out,
                double *qraux, int *jpvt, double *work)
{
    double *work1 = work, *work2 = work + p;
    for (int j = 0; j < p; ++j) {
        qraux[j] = orc_dnrm2(n, &X_(0, j), 1);
        work1[j] = qraux[j];
        work2[j] = qraux[j];
        if (work2[j] == 0.0) work2[j] = 1.0;
    }
    int lup = n < p ? n : p;
    int k = p + 1; /* 1-based, as in the Fortran */
    for (int l = 1; l <= lup; ++l) {
        /* cycle negligible columns l..p to the end */
        while (l < k && qraux[l - 1] < work2[l - 1] * tol) {
            int lp1 = l + 1;
            for (int i = 0; i < n; ++i) {
                double t = X_(i, l - 1);
                for (int j = lp1; j <= p; ++j) X_(i, j - 2) = X_(i, j - 1);
                X_(i, p - 1) = t;
            }
            int ii = jpvt[l - 1];
            double t = qraux[l - 1], tt = work1[l - 1], ttt = work2[l - 1];
            for (int j = lp1; j <= p; ++j) {
                jpvt[j - 2] = jpvt[j - 1];
                qraux[j - 2] = qraux[j - 1];
                work1[j - 2] = work1[j - 1];
                work2[j - 2] = work2[j - 1];
            }
            jpvt[p - 1] = ii;
            qraux[p - 1] = t;
            work1[p - 1] = tt;
            work2[p - 1] = ttt;
            k -= 1;
        }
        if (l == n) continue;
        /* Householder transformation for column l */
        double nrmxl = orc_dnrm2(n - l + 1, &X_(l - 1, l - 1), 1);
        if (nrmxl == 0.0) continue;
        if (X_(l - 1, l - 1) != 0.0) nrmxl = copysign(fabs(nrmxl), X_(l - 1, l - 1));
        orc_dscal(n - l + 1, 1.0 / nrmxl, &X_(l - 1, l - 1));
        X_(l - 1, l - 1) = 1.0 + X_(l - 1, l - 1);
        /* apply to remaining columns, updating the norms */
        for (int j = l + 1; j <= p; ++j) {
            double t = -orc_ddot(n - l + 1, &X_(l - 1, l - 1), &X_(l - 1, j - 1)) / X_(l - 1, l - 1);
            orc_daxpy(n - l + 1, t, &X_(l - 1, l - 1), &X_(l - 1, j - 1));
            if (qraux[j - 1] == 0.0) continue;
            double r = fabs(X_(l - 1, j - 1)) / qraux[j - 1];
            double tt = 1.0 - r * r;
            if (tt < 0.0) tt = 0.0;
            t = tt;
            if (fabs(t) < 1e-6) {
                qraux[j - 1] = orc_dnrm2(n - l, &X_(l, j - 1), 1);
                work1[j - 1] = qraux[j - 1];
            } else {
                qraux[j - 1] = qraux[j - 1] * sqrt(t);
            }
        }
        /* save the transformation */
        qraux[l - 1] = X_(l - 1, l - 1);
        X_(l - 1, l - 1) = -nrmxl;
    }
    k = k - 1;
    if (k > n) k = n;
    *k_out = k;
}

/* ---------------------------------------------------------------- dqrsl.f */
/* Only the job = 1110 branch dqrls uses: qty = Q'y, b, rsd.  qy/xb unused.
 * Returns info (0, or j if R[j,j] == 0 during back-substitution). */
int orc_dqrsl_1110(double *x, int ldx, int n, int k, const double *qraux,
                   const double *y, double *qty, double *b, double *rsd)
{
    int info = 0;
    int ju = k < n - 1 ? k : n - 1;
    if (ju == 0) { /* special action when n == 1 */
        qty[0] = y[0];
        if (X_(0, 0) != 0.0) b[0] = y[0] / X_(0, 0);
        else info = 1;
        rsd[0] = 0.0;
        return info;
    }
    memcpy(qty, y, (size_t)n * sizeof(double));
    /* compute trans(q)*y */
    for (int j = 1; j <= ju; ++j) {
        if (qraux[j - 1] == 0.0) continue;
        double temp = X_(j - 1, j - 1);
        X_(j - 1, j - 1) = qraux[j - 1];
        double t = -orc_ddot(n - j + 1, &X_(j - 1, j - 1), &qty[j - 1]) / X_(j - 1, j - 1);
        orc_daxpy(n - j + 1, t, &X_(j - 1, j - 1), &qty[j - 1]);
        X_(j - 1, j - 1) = temp;
    }
    /* set up b and rsd */
    memcpy(b, qty, (size_t)k * sizeof(double));
    if (k < n) memcpy(rsd + k, qty + k, (size_t)(n - k) * sizeof(double));
    for (int i = 0; i < k; ++i) rsd[i] = 0.0;
    /* compute b: back-substitution on the leading k x k triangle */
    for (int jj = 1; jj <= k; ++jj) {
        int j = k - jj + 1;
        if (X_(j - 1, j - 1) == 0.0) { info = j; break; }
        b[j - 1] = b[j - 1] / X_(j - 1, j - 1);
        if (j != 1) {
            double t = -b[j - 1];
            orc_daxpy(j - 1, t, &X_(0, j - 1), b);
        }
    }
    /* compute rsd = Q * [0; qty(k+1:n)] */
    for (int jj = 1; jj <= ju; ++jj) {
        int j = ju - jj + 1;
        if (qraux[j - 1] == 0.0) continue;
        double temp = X_(j - 1, j - 1);
        X_(j - 1, j - 1) = qraux[j - 1];
        double t = -orc_ddot(n - j + 1, &X_(j - 1, j - 1), &rsd[j - 1]) / X_(j - 1, j - 1);
        orc_daxpy(n - j + 1, t, &X_(j - 1, j - 1), &rsd[j - 1]);
        X_(j - 1, j - 1) = temp;
    }
    return info;
}

/* ---------------------------------------------------------------- dqrls.f */
void orc_dqrls(double *x, int n, int p, const double *y, int ny, double tol,
               double *b, double *rsd, double *qty, int *k, int *jpvt,
               double *qraux, double *work)
{
    orc_dqrdc2(x, n, n, p, tol, k, qraux, jpvt, work);
    if (*k > 0) {
        for (int jj = 0; jj < ny; ++jj)
            orc_dqrsl_1110(x, n, n, *k, qraux, y + (size_t)jj * n, qty + (size_t)jj * n,
                           b + (size_t)jj * p, rsd + (size_t)jj * n);
    } else {
        for (int i = 0; i < n; ++i)
            for (int jj = 0; jj < ny; ++jj) rsd[i + (size_t)jj * n] = y[i + (size_t)jj * n];
    }
    for (int j = *k; j < p; ++j)
        for (int jj = 0; jj < ny; ++jj) b[j + (size_t)jj * p] = 0.0;
}

/* ------------------------------------------------- dpotri("Upper", p, a, lda) */
/* dtrtri('U','N') [singularity check + dtrti2] then dlauum('U') [dlauu2].
 * Returns LAPACK info: >0 = first exactly-zero diagonal (1-based). */
#define A_(i, j) a[(size_t)(i) + (size_t)(j) * (size_t)lda]
int orc_dpotri_upper(int p, double *a, int lda)
{
    for (int i = 0; i < p; ++i)
        if (A_(i, i) == 0.0) return i + 1;
    /* dtrti2: invert the upper triangle in place */
    for (int j = 0; j < p; ++j) {
        A_(j, j) = 1.0 / A_(j, j);
        double ajj = -A_(j, j);
        /* x := U(0:j,0:j) * x with x = a(0:j, j): dtrmv upper, no-trans, non-unit */
        for (int c = 0; c < j; ++c) {
            double temp = A_(c, j);
            if (temp != 0.0) {
                for (int r = 0; r < c; ++r) A_(r, j) += temp * A_(r, c);
                A_(c, j) = temp * A_(c, c);
            }
        }
        for (int r = 0; r < j; ++r) A_(r, j) *= ajj;
    }
    /* dlauu2: U * U' into the upper triangle */
    for (int i = 0; i < p; ++i) {
        double aii = A_(i, i);
        if (i < p - 1) {
            double s = 0.0;
            for (int c = i; c < p; ++c) s += A_(i, c) * A_(i, c);
            A_(i, i) = s;
            /* a(0:i, i) = aii * a(0:i, i) + A(0:i, i+1:p) * a(i, i+1:p)' */
            for (int r = 0; r < i; ++r) A_(r, i) *= aii;
            for (int c = i + 1; c < p; ++c) {
                double temp = A_(i, c);
                if (temp != 0.0)
                    for (int r = 0; r < i; ++r) A_(r, i) += temp * A_(r, c);
            }
        } else {
            for (int r = 0; r <= i; ++r) A_(r, i) *= aii;
        }
    }
    return 0;
}

/* ------------------------------------------------------------- voxel_lm */
/* src/modelling_functions.c:448-576.  y[n], X[n x p].  Writes
 *   coef[p]      (pivoted order, quirk Q2: the "unpivot" at :507-515 is a no-op)
 *   stats[p+3] = [F, t_0..t_{p-1}, R2, logLik]            (:542, :566-573)
 *   pivot[p], *rank.
 * Returns dpotri's info (the reference raises an R error when it is != 0,
 * :551-557).  Scratch is allocated per call, as the reference allocates per
 * voxel (:471, :476). */
static int orc_voxel_lm_full(const double *y, const double *X, int n, int p, double *coef,
                             double *stats, int *pivot, int *rank, double *resid, double *effects)
{
    double *x = (double *)malloc((size_t)n * p * sizeof(double));
    double *qraux = (double *)malloc((size_t)p * sizeof(double));
    double *work = (double *)malloc((size_t)2 * p * sizeof(double));
    double *se = (double *)malloc((size_t)p * sizeof(double));
    /* :476-482 copy X (row loop outside, as written) */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) x[i + (size_t)j * n] = X[i + (size_t)j * n];
    double tol = 1e-07;
    for (int j = 0; j < p; ++j) pivot[j] = j;
    orc_dqrls(x, n, p, y, 1, tol, coef, resid, effects, rank, pivot, qraux, work);

    /* :497-515 "unpivot": only copies coef[j] onto itself where pivot[j]==j */
    double rss = 0, mss = 0, sum_fitted = 0;
    for (int i = 0; i < n; ++i) {
        rss += pow(resid[i], 2);
        sum_fitted += (y[i] - resid[i]);
    }
    double logLik = -.5 * n * (log(2 * M_PI) + 1 - log(n) + log(rss));
    double mean_fitted = sum_fitted / n;
    for (int i = 0; i < n; ++i) mss += pow((y[i] - resid[i]) - mean_fitted, 2);
    int rdf = n - p;
    double resvar = rss / rdf;
    stats[0] = (mss / (p - 1)) / resvar;

    int info = orc_dpotri_upper(p, x, n);
    if (info == 0) {
        for (int i = 0; i < p; ++i) se[pivot[i]] = sqrt(x[(size_t)i * n + i] * resvar);
        for (int i = 0; i < p; ++i) stats[i + 1] = coef[i] / se[i];
        stats[p + 1] = mss / (mss + rss);
        stats[p + 2] = logLik;
    }
    free(x); free(qraux); free(work); free(se);
    return info;
}

int orc_voxel_lm(const double *y, const double *X, int n, int p, double *coef,
                 double *stats, int *pivot, int *rank)
{
    double *resid = (double *)malloc((size_t)n * sizeof(double));
    double *effects = (double *)malloc((size_t)n * sizeof(double));
    int info = orc_voxel_lm_full(y, X, n, p, coef, stats, pivot, rank, resid, effects);
    free(resid); free(effects);
    return info;
}

/* ------------------------------------------------------------ voxel_anova */
/* src/minc_anova.c:4-80.  asgn[p] is model.matrix's "assign" attribute (0 = intercept).
 * f[maxasgn-1]: F per term = (sum of effects_i^2 over the term's columns / df_term) / (ssr / (n-p)).
 * effects = the first p entries of Q'y as dqrls leaves them (indexed by PIVOTED column position
 * while asgn is indexed by original column -- the reference does not un-pivot). */
int orc_voxel_anova(const double *y, const double *X, int n, int p, const int32_t *asgn, double *f)
{
    double *resid = (double *)malloc((size_t)n * sizeof(double));
    double *effects = (double *)malloc((size_t)n * sizeof(double));
    double *coef = (double *)malloc((size_t)p * sizeof(double));
    double *stats = (double *)malloc((size_t)(p + 3) * sizeof(double));
    int *pivot = (int *)malloc((size_t)p * sizeof(int));
    int rank, maxasgn = 0;
    for (int i = 0; i < p; ++i) if (asgn[i] > maxasgn) maxasgn = asgn[i];
    maxasgn++;
    double *ss = (double *)calloc((size_t)maxasgn, sizeof(double));
    int *df = (int *)calloc((size_t)maxasgn, sizeof(int));
    int info = orc_voxel_lm_full(y, X, n, p, coef, stats, pivot, &rank, resid, effects);
    if (info == 0) {
        int dfr = n - p;
        for (int i = 0; i < p; ++i) {
            ss[asgn[i]] += pow(effects[i], 2);
            df[asgn[i]]++;
        }
        double ssr = 0.0;
        for (int i = 0; i < n; ++i) ssr += pow(resid[i], 2);
        for (int i = 1; i < maxasgn; ++i) f[i - 1] = (ss[i] / df[i]) / (ssr / dfr);
    }
    free(resid); free(effects); free(coef); free(stats); free(pivot); free(ss); free(df);
    return info;
}

/* vertex_anova_loop (src/vertex_modelling.c:178-282) and per_voxel_anova (src/minc_anova.c:142-304):
 * rows outside the mask are 0.0 in every column (:270-275).  out is V x nterms, ld ldOut. */
int orc_anova_loop(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p,
                   const int32_t *asgn, const double *mask, double lo, double hi, double *out, int64_t ldOut)
{
    int maxasgn = 0, rc = 0;
    for (int i = 0; i < p; ++i) if (asgn[i] > maxasgn) maxasgn = asgn[i];
    const int nterms = maxasgn;
    double *ybuf = (double *)malloc((size_t)n * sizeof(double));
    double *f = (double *)malloc((size_t)(nterms > 0 ? nterms : 1) * sizeof(double));
    for (int64_t i = 0; i < V && rc == 0; ++i) {
        if (mask && !(mask[i] > lo - 0.5 && mask[i] < hi + .5)) {
            for (int k = 0; k < nterms; ++k) out[i + k * ldOut] = 0.0;
            continue;
        }
        for (int j = 0; j < n; ++j) ybuf[j] = Y[i + ldY * j];
        rc = orc_voxel_anova(ybuf, X, n, p, asgn, f);
        if (rc) break;
        for (int k = 0; k < nterms; ++k) out[i + k * ldOut] = f[k];
    }
    free(ybuf); free(f);
    return rc;
}

/* ------------------------------------------------------------- voxel_wlm */
/* src/weighted_least_squares.c:14-166.  Rows of X and y are scaled by sqrt(w); residuals are
 * un-scaled again; wrss = sum w res^2; mss about the WEIGHTED mean of the fitted values; logLik
 * gains 0.5 * sum(log w).  stats = [F, t.., R2, logLik]. */
int orc_voxel_wlm(const double *y, const double *X, const double *w, int n, int p, double *coef,
                  double *stats, int *pivot, int *rank)
{
    double *x = (double *)malloc((size_t)n * p * sizeof(double));
    double *yw = (double *)malloc((size_t)n * sizeof(double));
    double *xws = (double *)malloc((size_t)n * sizeof(double));
    double *resid = (double *)malloc((size_t)n * sizeof(double));
    double *effects = (double *)malloc((size_t)n * sizeof(double));
    double *fitted = (double *)malloc((size_t)n * sizeof(double));
    double *qraux = (double *)malloc((size_t)p * sizeof(double));
    double *work = (double *)malloc((size_t)2 * p * sizeof(double));
    double *se = (double *)malloc((size_t)p * sizeof(double));
    double total_log_weight = 0;
    for (int i = 0; i < n; ++i) {
        total_log_weight += log(w[i]);
        xws[i] = sqrt(w[i]);
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < p; ++j) x[i + (size_t)j * n] = X[i + (size_t)j * n] * xws[i];
    for (int i = 0; i < n; ++i) yw[i] = y[i] * xws[i];
    for (int j = 0; j < p; ++j) pivot[j] = j;
    orc_dqrls(x, n, p, yw, 1, 1e-07, coef, resid, effects, rank, pivot, qraux, work);
    double rss = 0, wrss = 0, sumw = 0, mss = 0;
    for (int i = 0; i < n; ++i) {
        sumw += w[i];
        resid[i] /= xws[i];
        double res2 = pow(resid[i], 2);
        rss += res2;
        wrss += w[i] * res2;
        fitted[i] = y[i] - resid[i];
    }
    double logLik = .5 * (total_log_weight - n * (log(2 * M_PI) + 1 - log(n) + log(wrss)));
    double m = 0;
    for (int i = 0; i < n; ++i) m += (w[i] * fitted[i]) / sumw;
    for (int i = 0; i < n; ++i) mss += w[i] * pow(fitted[i] - m, 2);
    int rdf = n - p;
    double resvar = wrss / rdf;
    stats[0] = (mss / (p - 1)) / resvar;
    int info = orc_dpotri_upper(p, x, n);
    if (info == 0) {
        for (int i = 0; i < p; ++i) se[pivot[i]] = sqrt(x[(size_t)i * n + i] * resvar);
        for (int i = 0; i < p; ++i) stats[i + 1] = coef[i] / se[i];
        stats[p + 1] = mss / (mss + wrss);
        stats[p + 2] = logLik;
    }
    free(x); free(yw); free(xws); free(resid); free(effects); free(fitted); free(qraux); free(work); free(se);
    (void)rss;
    return info;
}

/* vertex_wlm_loop (src/weighted_least_squares.c:168-335), static design. */
int orc_vertex_wlm_loop(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *w,
                        double *out, int64_t ldOut)
{
    double *ybuf = (double *)malloc((size_t)n * sizeof(double));
    double *coef = (double *)malloc((size_t)p * sizeof(double));
    double *stats = (double *)malloc((size_t)(p + 3) * sizeof(double));
    int *pivot = (int *)malloc((size_t)p * sizeof(int));
    int rank, rc = 0;
    for (int64_t i = 0; i < V && rc == 0; ++i) {
        for (int j = 0; j < n; ++j) ybuf[j] = Y[i + ldY * j];
        rc = orc_voxel_wlm(ybuf, X, w, n, p, coef, stats, pivot, &rank);
        if (rc) break;
        out[i] = stats[0];
        out[i + 1 * ldOut] = stats[p + 1];
        for (int k = 2; k < p + 2; ++k) out[i + k * ldOut] = coef[k - 2];
        for (int k = 1; k < p + 1; ++k) out[i + (k + p + 1) * ldOut] = stats[k];
        out[i + (2 * p + 2) * ldOut] = stats[p + 2];
    }
    free(ybuf); free(coef); free(stats); free(pivot);
    return rc;
}

/* --------------------------------------------------------- vertex_lm_loop */
/* src/vertex_modelling.c:104-158.  Y is V x n column-major with leading
 * dimension ldY (the reference has ldY == V); out is V x (2p+3) with leading
 * dimension ldOut, columns [F, R2, beta x p, t x p, logLik].  Returns 0 or the
 * dpotri info of the first failing row (the reference aborts the whole call). */
int orc_vertex_lm_loop(const double *Y, int64_t V, int64_t ldY, int n,
                       const double *X, int p, double *out, int64_t ldOut)
{
    double *ybuf = (double *)malloc((size_t)n * sizeof(double));
    double *xbuf = (double *)malloc((size_t)n * p * sizeof(double));
    double *coef = (double *)malloc((size_t)p * sizeof(double));
    double *stats = (double *)malloc((size_t)(p + 3) * sizeof(double));
    int *pivot = (int *)malloc((size_t)p * sizeof(int));
    int rank, rc = 0;
    for (int64_t i = 0; i < V && rc == 0; ++i) {
        for (int j = 0; j < n; ++j) ybuf[j] = Y[i + ldY * j];           /* :106-110 */
        for (int j = 0; j < n * p; ++j) xbuf[j] = X[j];                 /* :128-131 */
        rc = orc_voxel_lm(ybuf, xbuf, n, p, coef, stats, pivot, &rank); /* :141 */
        if (rc) break;
        out[i] = stats[0];                                              /* :144 */
        out[i + 1 * ldOut] = stats[p + 1];                              /* :146 */
        for (int k = 2; k < p + 2; ++k) out[i + k * ldOut] = coef[k - 2];
        for (int k = 1; k < p + 1; ++k) out[i + (k + p + 1) * ldOut] = stats[k];
        out[i + (2 * p + 2) * ldOut] = stats[p + 2];                    /* :157 */
    }
    free(ybuf); free(xbuf); free(coef); free(stats); free(pivot);
    return rc;
}

/* ----------------------------------- vertex_lm_loop / minc2_model with a voxel-wise covariate */
/* src/vertex_modelling.c:112-140, src/modelling_functions.c:1118-1144: with a file term on the right-hand
 * side the design of voxel i is [Xs | z_i] (p = ps + 1), or [1 | z_i] (p = 2) when there is no static part
 * (Xs == NULL, ps == 0).  Z is V x n like Y.  mask may be NULL; masked rows are 0. */
int orc_lm_loop_cov(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                    const double *mask, double lo, double hi, double *out, int64_t ldOut)
{
    const int p = Xs ? ps + 1 : 2;
    double *ybuf = (double *)malloc((size_t)n * sizeof(double));
    double *xbuf = (double *)malloc((size_t)n * p * sizeof(double));
    double *coef = (double *)malloc((size_t)p * sizeof(double));
    double *stats = (double *)malloc((size_t)(p + 3) * sizeof(double));
    int *pivot = (int *)malloc((size_t)p * sizeof(int));
    int rank, rc = 0;
    for (int64_t i = 0; i < V && rc == 0; ++i) {
        if (mask && !(mask[i] > lo - 0.5 && mask[i] < hi + 0.5)) {
            for (int k = 0; k < 2 * p + 3; ++k) out[i + k * ldOut] = 0.0;
            continue;
        }
        for (int j = 0; j < n; ++j) ybuf[j] = Y[i + ldY * j];
        if (!Xs) {
            for (int j = 0; j < n; ++j) xbuf[j] = 1.0;
            for (int j = 0; j < n; ++j) xbuf[j + n] = Z[i + ldY * j];
        } else {
            for (int j = 0; j < n * ps; ++j) xbuf[j] = Xs[j];
            for (int j = 0; j < n; ++j) xbuf[j + (size_t)n * ps] = Z[i + ldY * j];
        }
        rc = orc_voxel_lm(ybuf, xbuf, n, p, coef, stats, pivot, &rank);
        if (rc) break;
        out[i] = stats[0];
        out[i + 1 * ldOut] = stats[p + 1];
        for (int k = 2; k < p + 2; ++k) out[i + k * ldOut] = coef[k - 2];
        for (int k = 1; k < p + 1; ++k) out[i + (k + p + 1) * ldOut] = stats[k];
        out[i + (2 * p + 2) * ldOut] = stats[p + 2];
    }
    free(ybuf); free(xbuf); free(coef); free(stats); free(pivot);
    return rc;
}

/* ------------------------------------------------------- minc2_model, "lm" */
/* src/modelling_functions.c:1012-1175 with the files already read: vol[i] is
 * subject i's volume (sizes[0]*sizes[1]*sizes[2] doubles in file dimension
 * order == one column of Y).  mask may be NULL (have_mask == 0).  Mask rule
 * :1063-1066: m > lo-0.5 && m < hi+0.5.  Output row = v0*s1*s2 + v1*s2 + v2
 * (:1059).  Masked rows are +0.0 in EVERY column: the sequential reference
 * zeroes only column 0 (:982) and leaves the rest of a fresh allocation
 * untouched; the parallel path (R/minc_voxel_statistics.R:898-905) and
 * per_voxel_anova (src/minc_anova.c:270-275) write explicit zeros.  This
 * oracle normalises to all-zero rows. */
int orc_minc2_model_lm(const double *Y, const int64_t sizes[3], int n,
                       const double *X, int p, const double *mask, double lo,
                       double hi, double *out)
{
    int64_t V = sizes[0] * sizes[1] * sizes[2];
    int64_t slice = sizes[1] * sizes[2];
    int ncol = 2 * p + 3;
    for (int64_t i = 0; i < V * ncol; ++i) out[i] = 0.0;
    double *ybuf = (double *)malloc((size_t)n * sizeof(double));
    double *xbuf = (double *)malloc((size_t)n * p * sizeof(double));
    double *coef = (double *)malloc((size_t)p * sizeof(double));
    double *stats = (double *)malloc((size_t)(p + 3) * sizeof(double));
    int *pivot = (int *)malloc((size_t)p * sizeof(int));
    int rank, rc = 0;
    for (int64_t v0 = 0; v0 < sizes[0] && rc == 0; ++v0) {
        const double *mask_buffer = mask ? mask + v0 * slice : NULL;
        int read_cur_slab = 1;
        if (mask) { /* :1028-1034 */
            read_cur_slab = 0;
            for (int64_t i = 0; i < slice; ++i)
                if (mask_buffer[i] > lo - 0.5 && mask_buffer[i] < hi + 0.5) read_cur_slab = 1;
        }
        if (!read_cur_slab) continue;
        for (int64_t v1 = 0; v1 < sizes[1] && rc == 0; ++v1) {
            for (int64_t v2 = 0; v2 < sizes[2]; ++v2) {
                int64_t output_index = v0 * slice + v1 * sizes[2] + v2;
                int64_t buffer_index = sizes[2] * v1 + v2;
                if (mask && !(mask_buffer[buffer_index] > lo - 0.5 &&
                              mask_buffer[buffer_index] < hi + 0.5))
                    continue;
                for (int i = 0; i < n; ++i) ybuf[i] = Y[output_index + V * i];
                for (int j = 0; j < n * p; ++j) xbuf[j] = X[j];
                rc = orc_voxel_lm(ybuf, xbuf, n, p, coef, stats, pivot, &rank);
                if (rc) break;
                out[output_index] = stats[0];
                out[output_index + 1 * V] = stats[p + 1];
                for (int k = 2; k < p + 2; ++k) out[output_index + k * V] = coef[k - 2];
                for (int k = 1; k < p + 1; ++k) out[output_index + (k + p + 1) * V] = stats[k];
                out[output_index + (2 * p + 2) * V] = stats[p + 2];
            }
        }
    }
    free(ybuf); free(xbuf); free(coef); free(stats); free(pivot);
    return rc;
}

/* ----------------------------------------------------------- mincRandomize */
/* R/minc_voxel_statistics.R:1465-1497.  For each of R index vectors (idx is
 * n x R, 0-based, drawn by the caller so RNG semantics stay the caller's):
 * permute the predictor rows (X_pi = X[idx,]; all predictor columns move
 * together, :1470-1474), refit every voxel (update(lmod, data=shuf) re-runs
 * the original masked call, :1477), set non-finite to 0 (:1478), abs() when
 * two-sided (:1484-1486), column max over ALL V rows incl. masked zeros
 * (:1488).  cols are 0-based column indices into the V x (2p+3) matrix.
 * dist is R x ncols column-major.  scratch must hold V*(2p+3) doubles. */
int orc_randomize(const double *Y, int64_t V, int n, const double *X, int p,
                  const double *mask, double lo, double hi, const int32_t *idx,
                  int R, const int32_t *cols, int ncols, int two_sided,
                  double *dist, double *scratch)
{
    double *Xp = (double *)malloc((size_t)n * p * sizeof(double));
    int64_t sizes[3] = {1, 1, V};
    int rc = 0;
    for (int r = 0; r < R && rc == 0; ++r) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < p; ++j)
                Xp[i + (size_t)j * n] = X[idx[i + (size_t)r * n] + (size_t)j * n];
        rc = orc_minc2_model_lm(Y, sizes, n, Xp, p, mask, lo, hi, scratch);
        if (rc) break;
        for (int c = 0; c < ncols; ++c) {
            const double *col = scratch + (size_t)cols[c] * V;
            double best = -INFINITY;
            for (int64_t v = 0; v < V; ++v) {
                double s = col[v];
                if (!isfinite(s)) s = 0.0;
                if (two_sided) s = fabs(s);
                if (s > best) best = s;
            }
            dist[r + (size_t)c * R] = best;
        }
    }
    free(Xp);
    return rc;
}

/* design-only probe used by tests: pivot, rank, R (p x p upper, ld p),
 * c = diag((R'R)^-1) as dpotri leaves it, and info. */
int orc_design_probe(const double *X, int n, int p, int *pivot, int *rank,
                     double *Rout, double *cdiag)
{
    double *x = (double *)malloc((size_t)n * p * sizeof(double));
    double *y = (double *)calloc((size_t)n, sizeof(double));
    double *b = (double *)malloc((size_t)p * sizeof(double));
    double *rsd = (double *)malloc((size_t)n * sizeof(double));
    double *qty = (double *)malloc((size_t)n * sizeof(double));
    double *qraux = (double *)malloc((size_t)p * sizeof(double));
    double *work = (double *)malloc((size_t)2 * p * sizeof(double));
    memcpy(x, X, (size_t)n * p * sizeof(double));
    for (int j = 0; j < p; ++j) pivot[j] = j;
    orc_dqrls(x, n, p, y, 1, 1e-07, b, rsd, qty, rank, pivot, qraux, work);
    for (int j = 0; j < p; ++j)
        for (int i = 0; i < p; ++i) Rout[i + (size_t)j * p] = (i <= j && i < n) ? x[i + (size_t)j * n] : 0.0;
    int info = orc_dpotri_upper(p, x, n);
    for (int i = 0; i < p; ++i) cdiag[i] = x[(size_t)i * n + i];
    free(x); free(y); free(b); free(rsd); free(qty); free(qraux); free(work);
    return info;
}

/* ------------------------------------------------------------------ stored type -> doubles */
/* What the reference's read loop hands to voxel_lm (src/modelling_functions.c:1037-1055): libminc's
 * miget_real_value_hyperslab(MI_TYPE_DOUBLE) expands the stored voxels with the file's valid range and the
 * per-slice real range (image-min / image-max).  libminc (BIC-MNI/libminc @ b01ba22,
 * inst/tools/ch_minc_depends.m4:85) is absent from the image, so its published conversion is restated:
 *     real = (voxel - valid_min) * (slice_max - slice_min) / (valid_max - valid_min) + slice_min
 * with scale[slice + nslices*j] = (slice_max - slice_min) / (valid_max - valid_min) and
 * offset[slice + nslices*j] = slice_min precomputed per (slice, subject).  Plain C doubles: one subtraction, one
 * multiplication, one addition, each rounded (this file is compiled without FMA contraction).
 * U is V x n column-major (ld ldU); out V x n (ld ldOut); slice = v / slice_len. */
void orc_expand_u16(const uint16_t *U, int64_t V, int64_t ldU, int n, const double *scale, const double *offset,
                    double voxel_min, int64_t slice_len, double *out, int64_t ldOut)
{
    const int64_t nslices = (V + slice_len - 1) / slice_len;
    for (int j = 0; j < n; ++j)
        for (int64_t v = 0; v < V; ++v) {
            const int64_t s = v / slice_len;
            volatile double t = (double)U[v + ldU * j] - voxel_min;
            volatile double m = t * scale[s + nslices * j];
            out[v + ldOut * j] = m + offset[s + nslices * j];
        }
}

void orc_expand_f32(const float *U, int64_t V, int64_t ldU, int n, double *out, int64_t ldOut)
{
    for (int j = 0; j < n; ++j)
        for (int64_t v = 0; v < V; ++v) out[v + ldOut * j] = (double)U[v + ldU * j];
}

/* ------------------------------------------------------------------ graph TFCE */
/* graph_tfce_wqu / graph_tfce (src/vertex_tfce.cpp:92-157): threshold-free cluster enhancement on an arbitrary
 * adjacency graph.  Vertices are visited in descending order of the map; for thresholds t = hmax, hmax - dh, ...
 * while t >= 0 (dh = hmax / nsteps, t decremented in floating point exactly like the reference's loop) the vertices
 * with map >= t are united with their neighbours in [t, map[v]] by weighted quick union (smaller tree under larger,
 * path halving, cluster mass carried at the root), and every such vertex receives t^H * mass(cluster)^E * dh.
 * side: 0 both (tfce(map) - tfce(-map)), 1 positive, 2 negative (-tfce(-map)); R/minc_vertex_statistics.R:788-796.
 * adj_ptr / adj_idx: CSR adjacency, 0-based. */
typedef struct { const double *v; } orc_sortctx;
static const double *orc_sort_vals;
static int orc_cmp_desc(const void *a, const void *b)
{
    const double x = orc_sort_vals[*(const int32_t *)a], y = orc_sort_vals[*(const int32_t *)b];
    if (x > y) return -1;
    if (x < y) return 1;
    return (*(const int32_t *)a > *(const int32_t *)b) - (*(const int32_t *)a < *(const int32_t *)b);
}

static int32_t orc_wqu_root(int32_t *id, int32_t i)
{
    while (i != id[i]) {
        id[i] = id[id[i]];
        i = id[i];
    }
    return i;
}

static void orc_tfce_one_side(const double *map, int64_t V, const int32_t *adj_ptr, const int32_t *adj_idx,
                              const double *weights, double E, double H, int nsteps, double *tfce)
{
    int32_t *order = (int32_t *)malloc((size_t)V * sizeof(int32_t));
    int32_t *id = (int32_t *)malloc((size_t)V * sizeof(int32_t));
    int32_t *sz = (int32_t *)malloc((size_t)V * sizeof(int32_t));
    double *mass = (double *)malloc((size_t)V * sizeof(double));
    for (int64_t v = 0; v < V; ++v) {
        order[v] = (int32_t)v;
        id[v] = (int32_t)v;
        sz[v] = 1;
        mass[v] = weights[v];
        tfce[v] = 0.0;
    }
    orc_sort_vals = map;
    qsort(order, (size_t)V, sizeof(int32_t), orc_cmp_desc);
    const double hmin = 0.0, hmax = V > 0 ? map[order[0]] : 0.0;
    const double dh = (hmax - hmin) / (double)nsteps;
    if (hmax > hmin) {
        for (double t = hmax; t >= hmin; t -= dh) {
            int64_t nvisited = 0;
            for (int64_t s = 0; s < V && map[order[s]] >= t; ++s) {
                const int32_t ind = order[s];
                ++nvisited;
                for (int32_t e = adj_ptr[ind]; e < adj_ptr[ind + 1]; ++e) {
                    const int32_t nb = adj_idx[e];
                    if (map[nb] >= t && map[nb] <= map[ind]) {
                        const int32_t i = orc_wqu_root(id, nb), j = orc_wqu_root(id, ind);
                        if (i != j) {
                            if (sz[i] < sz[j]) { id[i] = j; sz[j] += sz[i]; mass[j] += mass[i]; }
                            else { id[j] = i; sz[i] += sz[j]; mass[i] += mass[j]; }
                        }
                    }
                }
            }
            const double height_mul = pow(t, H);
            for (int64_t s = 0; s < nvisited; ++s) {
                int32_t r = order[s];
                while (r != id[r]) r = id[r];
                tfce[order[s]] += height_mul * pow(mass[r], E) * dh;
            }
        }
    }
    free(order); free(id); free(sz); free(mass);
}

void orc_graph_tfce(const double *map, int64_t V, const int32_t *adj_ptr, const int32_t *adj_idx, const double *weights,
                    double E, double H, int nsteps, int side, double *out)
{
    if (side == 1) {
        orc_tfce_one_side(map, V, adj_ptr, adj_idx, weights, E, H, nsteps, out);
        return;
    }
    double *neg = (double *)malloc((size_t)(V > 0 ? V : 1) * sizeof(double));
    double *tmp = (double *)malloc((size_t)(V > 0 ? V : 1) * sizeof(double));
    for (int64_t v = 0; v < V; ++v) neg[v] = -map[v];
    orc_tfce_one_side(neg, V, adj_ptr, adj_idx, weights, E, H, nsteps, tmp);
    if (side == 2) {
        for (int64_t v = 0; v < V; ++v) out[v] = -tmp[v];
    } else {
        orc_tfce_one_side(map, V, adj_ptr, adj_idx, weights, E, H, nsteps, out);
        for (int64_t v = 0; v < V; ++v) out[v] -= tmp[v];
    }
    free(neg); free(tmp);
}
