// design.hpp -- host-side factorisation of the shared model matrix (piece K0).
//
// The reference re-factorises the SAME design X for every voxel
// (src/modelling_functions.c:476-495 copy X + F77 dqrls; :549 dpotri).  Everything
// that depends only on X is computed here once per call: column pivoting and
// numerical rank exactly as R's dqrdc2 decides them (tol 1e-7), the thin
// orthonormal factor Q_k, R_k^-1, diag((R'R)^-1) as dpotri leaves it, the
// dpotri error contract, and the constants the device epilogue needs.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rmg {

constexpr int kMaxP = 32;

struct Design {
    int n = 0, p = 0, k = 0;          // rows, columns, numerical rank
    std::vector<int32_t> pivot;       // p, 0-based; column pivot[i] of X sits at position i
    std::vector<double> Q;            // n x p column-major, columns >= k are zero
    std::vector<double> R;            // p x p column-major upper triangle (pivoted order)
    std::vector<double> Rinv;         // p x p column-major; leading k x k = inverse of R[:k,:k]
    std::vector<double> cdiag;        // p: diag((R'R)^-1) over the full p x p triangle
    std::vector<double> ct;           // p: ct[m] = cdiag[i] with pivot[i] == m   (se "unpivot",
                                      //    src/modelling_functions.c:560-569)
    std::vector<double> q1;           // p: Q' 1 (zeros after k)
    double gap = 0.0;                 // n - |q1|^2 = |1 - Q Q'1|^2; exactly 0 when 1 is in span(X)
    bool has_intercept = false;       // 1 in span(X): per-voxel shift by the first sample is exact
    int dpotri_info = 0;              // > 0: R[info,info] == 0 -> the reference raises an R error
    // weighted least squares (voxel_wlm, src/weighted_least_squares.c:14-166); empty = unweighted
    std::vector<double> rowscale;     // n: sqrt(w_j); every sample is multiplied by it before accumulation
    double sumw = 0.0;                // sum of the weights (n when unweighted)
    double loglik_add = 0.0;          // 0.5 * sum(log w_j)
};

// Returns 0, or dpotri_info (> 0) when the inverse cannot be formed.  Throws nothing.
// w == nullptr: ordinary least squares.  Otherwise rows of X are scaled by sqrt(w) first and the
// "constant vector" of the mean/intercept logic becomes sqrt(w) itself.
int factor_design(const double *X, int n, int p, Design &D, const double *w = nullptr);

// The reference's message for dpotri info > 0 (src/modelling_functions.c:554).
std::string singular_message(int info);

}  // namespace rmg
