// design.cpp -- see design.hpp.  Host C++17, no dependencies.
//
// The pivoting / rank decisions must be the ones R's dqrdc2 (called through dqrls at
// src/modelling_functions.c:494 with tol = 1e-7) takes, because they define which
// coefficient lands in which output column and when the whole call errors out
// (SURVEY.md App. A.2-A.3).  So the Householder reduction below keeps LINPACK's
// conventions: v scaled so that v[0] = 1 + |x_ll|/|x_l|, H = I - v v'/v[0],
// R[l,l] = -sign(x_ll)|x_l|, running column-norm downdates with the 1e-6 refresh
// rule, and negligible columns cycled to the back (never swapped).
#include "design.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>

namespace rmg {
namespace {

// Overflow-safe 2-norm, one scaled pass (the classic BLAS-1 formulation).
double nrm2(const double *x, int n)
{
    if (n < 1) return 0.0;
    if (n == 1) return std::fabs(x[0]);
    double scale = 0.0, ssq = 1.0;
    for (int i = 0; i < n; ++i) {
        if (x[i] == 0.0) continue;
        const double a = std::fabs(x[i]);
        if (scale < a) {
            const double r = scale / a;
            ssq = 1.0 + ssq * r * r;
            scale = a;
        } else {
            const double r = a / scale;
            ssq += r * r;
        }
    }
    return scale * std::sqrt(ssq);
}

double dot(const double *a, const double *b, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

void axpy(double t, const double *v, double *y, int n)
{
    if (t == 0.0) return;
    for (int i = 0; i < n; ++i) y[i] += t * v[i];
}

struct Householder {
    std::vector<double> a;      // n x p, reduced in place: R above, reflector tails below
    std::vector<double> vhead;  // p: v[0] of each reflector (0 = identity)
    std::vector<int32_t> perm;  // p
    int rank = 0;
};

// Rotate column l to the last position, shifting l+1..p-1 one place left.
template <class T>
void rotate_to_back(std::vector<T> &v, int l)
{
    std::rotate(v.begin() + l, v.begin() + l + 1, v.end());
}

Householder reduce(const double *X, int n, int p, double tol)
{
    Householder H;
    H.a.assign(X, X + static_cast<size_t>(n) * p);
    H.vhead.assign(p, 0.0);
    H.perm.resize(p);
    for (int j = 0; j < p; ++j) H.perm[j] = j;
    auto col = [&](int j) { return H.a.data() + static_cast<size_t>(j) * n; };

    std::vector<double> live(p), refreshed(p), original(p);
    for (int j = 0; j < p; ++j) {
        live[j] = refreshed[j] = nrm2(col(j), n);
        original[j] = live[j] == 0.0 ? 1.0 : live[j];
    }
    int kept = p;  // columns [0, kept) are still candidates
    const int steps = std::min(n, p);
    for (int l = 0; l < steps; ++l) {
        while (l < kept && live[l] < original[l] * tol) {
            // physically move column l behind everything else
            std::vector<double> tmp(col(l), col(l) + n);
            std::copy(col(l + 1), col(p - 1) + n, col(l));
            std::copy(tmp.begin(), tmp.end(), col(p - 1));
            rotate_to_back(H.perm, l);
            rotate_to_back(live, l);
            rotate_to_back(refreshed, l);
            rotate_to_back(original, l);
            --kept;
        }
        if (l == n - 1) break;
        double *v = col(l) + l;
        const int m = n - l;
        double nrm = nrm2(v, m);
        if (nrm == 0.0) continue;
        if (v[0] != 0.0) nrm = std::copysign(std::fabs(nrm), v[0]);
        const double inv = 1.0 / nrm;
        for (int i = 0; i < m; ++i) v[i] *= inv;
        v[0] = 1.0 + v[0];
        for (int j = l + 1; j < p; ++j) {
            double *c = col(j) + l;
            const double t = -dot(v, c, m) / v[0];
            axpy(t, v, c, m);
            if (live[j] == 0.0) continue;
            const double r = std::fabs(c[0]) / live[j];
            const double shrink = std::max(1.0 - r * r, 0.0);
            if (std::fabs(shrink) < 1e-6) {
                live[j] = nrm2(c + 1, m - 1);
                refreshed[j] = live[j];
            } else {
                live[j] *= std::sqrt(shrink);
            }
        }
        H.vhead[l] = v[0];
        v[0] = -nrm;
    }
    H.rank = std::min(kept, n);
    return H;
}

// y <- H_l y for reflector l (acts on rows l..n-1).
void apply_reflector(const Householder &H, int n, int l, double *y)
{
    const double v0 = H.vhead[l];
    if (v0 == 0.0) return;
    const double *tail = H.a.data() + static_cast<size_t>(l) * n + l;  // tail[0] holds R[l,l]
    const int m = n - l;
    double s = v0 * y[l];
    for (int i = 1; i < m; ++i) s += tail[i] * y[l + i];
    const double t = -s / v0;
    y[l] += t * v0;
    for (int i = 1; i < m; ++i) y[l + i] += t * tail[i];
}

}  // namespace

std::string singular_message(int info)
{
    char buf[128];
    std::snprintf(buf, sizeof buf, "element (%d, %d) is zero, so the inverse cannot be computed", info, info);
    return buf;
}

int factor_design(const double *X_in, int n, int p, Design &D, const double *w)
{
    D = Design{};
    D.n = n;
    D.p = p;
    D.sumw = n;
    std::vector<double> Xw;
    const double *X = X_in;
    if (w) {   // :41-53: xws = sqrt(w); new_x = X * xws (row-wise)
        D.rowscale.resize(n);
        D.sumw = 0.0;
        double tlw = 0.0;
        for (int i = 0; i < n; ++i) {
            tlw += std::log(w[i]);
            D.rowscale[i] = std::sqrt(w[i]);
            D.sumw += w[i];
        }
        D.loglik_add = 0.5 * tlw;
        Xw.resize(static_cast<size_t>(n) * p);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < p; ++j) Xw[i + static_cast<size_t>(j) * n] = X_in[i + static_cast<size_t>(j) * n] * D.rowscale[i];
        X = Xw.data();
    }
    const Householder H = reduce(X, n, p, 1e-7);
    const int k = H.rank;
    D.k = k;
    D.pivot = H.perm;

    // R: the p x p upper triangle of the reduced matrix (rows beyond n do not exist).
    D.R.assign(static_cast<size_t>(p) * p, 0.0);
    for (int j = 0; j < p; ++j)
        for (int i = 0; i <= j && i < n; ++i) D.R[i + static_cast<size_t>(j) * p] = H.a[i + static_cast<size_t>(j) * n];

    // Thin Q: columns 0..k-1 of H_0 H_1 ... applied to the identity.
    D.Q.assign(static_cast<size_t>(n) * p, 0.0);
    const int nref = std::min(k, n - 1);
    for (int c = 0; c < k; ++c) {
        double *q = D.Q.data() + static_cast<size_t>(c) * n;
        q[c] = 1.0;
        for (int l = nref - 1; l >= 0; --l) apply_reflector(H, n, l, q);
    }

    // q1 = Q'u and the part of u outside span(Q), u = 1 (or sqrt(w) for weighted fits: the image of
    // the constant vector in the scaled problem).
    D.q1.assign(p, 0.0);
    for (int c = 0; c < k; ++c) {
        const double *q = D.Q.data() + static_cast<size_t>(c) * n;
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += q[i] * (w ? D.rowscale[i] : 1.0);
        D.q1[c] = s;
    }
    double gap = 0.0;
    for (int i = 0; i < n; ++i) {
        double r = w ? D.rowscale[i] : 1.0;
        for (int c = 0; c < k; ++c) r -= D.Q[i + static_cast<size_t>(c) * n] * D.q1[c];
        gap += r * r;
    }
    D.has_intercept = gap < 1e-26 * D.sumw;
    D.gap = D.has_intercept ? 0.0 : gap;

    // Inverse of the leading k x k triangle (used for beta = R_k^-1 Q_k'y).
    D.Rinv.assign(static_cast<size_t>(p) * p, 0.0);
    auto Rk = [&](int i, int j) { return D.R[i + static_cast<size_t>(j) * p]; };
    auto Ri = [&](int i, int j) -> double & { return D.Rinv[i + static_cast<size_t>(j) * p]; };
    bool rk_ok = true;
    for (int j = 0; j < k; ++j)
        if (Rk(j, j) == 0.0) rk_ok = false;
    if (rk_ok) {
        for (int j = 0; j < k; ++j) {
            Ri(j, j) = 1.0 / Rk(j, j);
            for (int i = j - 1; i >= 0; --i) {
                double s = 0.0;
                for (int m = i + 1; m <= j; ++m) s += Rk(i, m) * Ri(m, j);
                Ri(i, j) = -s / Rk(i, i);
            }
        }
    }

    // dpotri on the FULL p x p triangle, as voxel_lm calls it (:549): first the
    // singularity check of dtrtri, then U^-1 column by column, then diag(U^-1 U^-T).
    D.cdiag.assign(p, 0.0);
    D.ct.assign(p, 0.0);
    D.dpotri_info = 0;
    for (int i = 0; i < p; ++i)
        if (Rk(i, i) == 0.0) {
            D.dpotri_info = i + 1;
            return D.dpotri_info;
        }
    std::vector<double> U(D.R);  // p x p
    auto Uij = [&](int i, int j) -> double & { return U[i + static_cast<size_t>(j) * p]; };
    for (int j = 0; j < p; ++j) {
        Uij(j, j) = 1.0 / Uij(j, j);
        const double ajj = -Uij(j, j);
        for (int c = 0; c < j; ++c) {  // x <- Uinv[0:j,0:j] * x, x = column j above the diagonal
            const double t = Uij(c, j);
            if (t != 0.0) {
                for (int r = 0; r < c; ++r) Uij(r, j) += t * Uij(r, c);
                Uij(c, j) = t * Uij(c, c);
            }
        }
        for (int r = 0; r < j; ++r) Uij(r, j) *= ajj;
    }
    for (int i = 0; i < p; ++i) {
        double s = 0.0;
        if (i < p - 1) {
            for (int c = i; c < p; ++c) s += Uij(i, c) * Uij(i, c);
        } else {
            s = Uij(i, i) * Uij(i, i);
        }
        D.cdiag[i] = s;
    }
    for (int i = 0; i < p; ++i) D.ct[D.pivot[i]] = D.cdiag[i];
    return 0;
}

}  // namespace rmg
