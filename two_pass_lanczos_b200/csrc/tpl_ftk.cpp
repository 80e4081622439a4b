// tpl_ftk.cpp -- host f(T_k) e_1 solvers on the k x k Lanczos tridiagonal (component H3).
//
// "f(T_k): the host computes f(T_k)e1 on the small tridiagonal, unchanged from the reference": these are the
// closures the reference's binaries and tests define around faer, restated without faer:
//   tpl_ftk_inv    T_k y = e1 by tridiagonal partial-pivot LU  (sp_lu at src/bin/tradeoff.rs:245-258,
//                  src/bin/stability.rs:161-170; dense partial_piv_lu at tests/correctness.rs:171-179)
//   tpl_ftk_exp    y = Q exp(L) Q^T e1 via symmetric tridiagonal EVD (self_adjoint_eigen at
//                  src/bin/stability.rs:175-193, tests/correctness.rs:214-241)
//   tpl_ftk_square y = T_k (T_k e1)                              (tests/correctness.rs:290-299)
// All have the tpl_ftk_solver signature, so they plug straight into tpl_lanczos / tpl_lanczos_two_pass.
#include <cmath>
#include <cstring>
#include <vector>

#include "tpl_internal.h"

namespace {

// Gaussian elimination with partial pivoting on a tridiagonal system (one right-hand side).
// dl/du: sub/super-diagonal (n-1), d: diagonal (n); all overwritten.  Returns false when singular.
bool tridiag_solve(size_t n, std::vector<double>& dl, std::vector<double>& d, std::vector<double>& du,
                   std::vector<double>& b) {
  if (n == 0) return true;
  std::vector<double> du2(n > 2 ? n - 2 : 0, 0.0);
  for (size_t i = 0; i + 1 < n; ++i) {
    if (std::fabs(d[i]) >= std::fabs(dl[i])) {
      if (d[i] == 0.0) return false;
      const double f = dl[i] / d[i];
      d[i + 1] -= f * du[i];
      b[i + 1] -= f * b[i];
      if (i + 2 < n) du2[i] = 0.0;
    } else {  // swap rows i and i+1
      const double f = d[i] / dl[i];
      d[i] = dl[i];
      const double t = d[i + 1];
      d[i + 1] = du[i] - f * t;
      if (i + 2 < n) {
        du2[i] = du[i + 1];
        du[i + 1] = -f * du[i + 1];
      }
      du[i] = t;
      std::swap(b[i], b[i + 1]);
      b[i + 1] -= f * b[i];
    }
  }
  if (d[n - 1] == 0.0) return false;
  b[n - 1] /= d[n - 1];
  if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
  if (n > 2)
    for (size_t i = n - 2; i-- > 0;) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / d[i];
  return true;
}

// Implicit-shift QL iteration for a symmetric tridiagonal matrix with eigenvector accumulation.
// d (n): diagonal in, eigenvalues out.  e (n): sub-diagonal in e[0..n-2].  z: n*n, eigenvector j is the
// contiguous block z[j*n .. j*n+n) (identity on entry).  Returns false if an eigenvalue needs > 60 sweeps.
bool tridiag_eigh(size_t n, std::vector<double>& d, std::vector<double>& e, std::vector<double>& z) {
  if (n == 0) return true;
  e[n - 1] = 0.0;
  for (size_t l = 0; l < n; ++l) {
    int iter = 0;
    size_t m;
    do {
      for (m = l; m + 1 < n; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        size_t i = m;
        bool underflow = false;
        while (i-- > l) {
          double f = s * e[i];
          const double b = c * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            underflow = true;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          double* zi = &z[i * n];
          double* zi1 = &z[(i + 1) * n];
          for (size_t k = 0; k < n; ++k) {
            f = zi1[k];
            zi1[k] = s * zi[k] + c * f;
            zi[k] = c * zi[k] - s * f;
          }
        }
        if (underflow) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

}  // namespace

extern "C" {

int tpl_ftk_inv(const double* alphas, size_t na, const double* betas, size_t nb, double* y, size_t* y_len,
                void*) {
  if (na == 0) {
    if (y_len) *y_len = 0;
    return 0;
  }
  if (nb + 1 < na) return tpl::fail(TPL_ERR_SOLVER, "tpl_ftk_inv: betas shorter than alphas - 1");
  std::vector<double> dl(betas, betas + (na - 1)), du(dl), d(alphas, alphas + na), b(na, 0.0);
  b[0] = 1.0;
  if (!tridiag_solve(na, dl, d, du, b)) return tpl::fail(TPL_ERR_SOLVER, "tpl_ftk_inv: T_k is singular");
  std::memcpy(y, b.data(), na * sizeof(double));
  *y_len = na;
  return 0;
}

// residual norms of ALL Lanczos iterates from one pass-1 decomposition, without touching a vector.
// x_j = ||b|| V_j T_j^{-1} e_1 has the residual b - A x_j = -||b|| beta_j (e_j^T T_j^{-1} e_1) v_{j+1}.  Its norm is read off a
// progressive Givens QR of the (j+1) x j tridiagonal (Paige-Saunders): with rotations G_i = [[c_i, s_i], [-s_i, c_i]]
// annihilating beta_i, ||r_j|| = ||b|| |s_1 ... s_j| / |c_j|  (c_j = 0: T_j is singular, the iterate does not exist -> inf).
// Stable for indefinite T (KKT systems), unlike the pivot-free LU recurrence.
int tpl_ftk_inv_residuals(const double* alphas, size_t na, const double* betas, size_t nb, double b_norm, double* res) {
  if ((na && !alphas) || (nb && !betas) || (na && !res)) return tpl::fail(TPL_ERR_PANIC, "null argument");
  double c_prev = 1.0, s_prev = 0.0, c_prev2 = 1.0;  // G_{j-1}, and the cosine of G_{j-2}
  double phi = b_norm;                               // ||b|| |s_1 ... s_{j-1}|
  for (size_t j = 0; j < na; ++j) {
    if (j >= nb) {  // beta_j unknown: nothing can be said from here on
      for (size_t i = j; i < na; ++i) res[i] = NAN;
      break;
    }
    const double sup = j ? c_prev2 * betas[j - 1] : 0.0;        // row j-1 of column j after G_{j-2}
    const double gamma = c_prev * alphas[j] - s_prev * sup;     // row j after G_{j-1}
    const double rho = std::hypot(gamma, betas[j]);
    if (rho == 0.0) {  // breakdown with a singular T_j: invariant subspace, no iterate
      for (size_t i = j; i < na; ++i) res[i] = NAN;
      break;
    }
    const double c = gamma / rho, s = betas[j] / rho;
    res[j] = c == 0.0 ? INFINITY : phi * std::fabs(s) / std::fabs(c);
    phi *= std::fabs(s);
    c_prev2 = c_prev;
    c_prev = c;
    s_prev = s;
  }
  return 0;
}

int tpl_ftk_exp(const double* alphas, size_t na, const double* betas, size_t nb, double* y, size_t* y_len,
                void*) {
  if (na == 0) {
    if (y_len) *y_len = 0;
    return 0;
  }
  if (nb + 1 < na) return tpl::fail(TPL_ERR_SOLVER, "tpl_ftk_exp: betas shorter than alphas - 1");
  std::vector<double> d(alphas, alphas + na), e(na, 0.0), z(na * na, 0.0);
  for (size_t i = 0; i + 1 < na; ++i) e[i] = betas[i];
  for (size_t i = 0; i < na; ++i) z[i * na + i] = 1.0;
  if (!tridiag_eigh(na, d, e, z))
    return tpl::fail(TPL_ERR_EVD, "A numerical error occurred during the eigendecomposition of T_k: NoConvergence");
  // y = sum_i q_i * exp(lambda_i) * q_i[0]
  for (size_t k = 0; k < na; ++k) y[k] = 0.0;
  for (size_t i = 0; i < na; ++i) {
    const double* q = &z[i * na];
    const double w = std::exp(d[i]) * q[0];
    for (size_t k = 0; k < na; ++k) y[k] += w * q[k];
  }
  *y_len = na;
  return 0;
}

int tpl_ftk_square(const double* alphas, size_t na, const double* betas, size_t nb, double* y, size_t* y_len,
                   void*) {
  if (na == 0) {
    if (y_len) *y_len = 0;
    return 0;
  }
  if (nb + 1 < na) return tpl::fail(TPL_ERR_SOLVER, "tpl_ftk_square: betas shorter than alphas - 1");
  for (size_t k = 0; k < na; ++k) y[k] = 0.0;
  // c = T e1 = (alpha_0, beta_0, 0, ...);  y = T c
  const double c0 = alphas[0], c1 = na > 1 ? betas[0] : 0.0;
  y[0] = alphas[0] * c0 + (na > 1 ? betas[0] * c1 : 0.0);
  if (na > 1) y[1] = betas[0] * c0 + alphas[1] * c1;
  if (na > 2) y[2] = betas[1] * c1;
  *y_len = na;
  return 0;
}

}  // extern "C"
