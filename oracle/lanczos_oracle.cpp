// =====================================================================================
// oracle/lanczos_oracle.cpp  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// A single-threaded CPU restatement of the reference's two-pass / one-pass Lanczos path
// (lukefleed/two-pass-lanczos), used ONLY by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py as the checker and the timed CPU arm.
// Nothing under two_pass_lanczos_b200/ may link, import or call this file.
//
// The reference itself (Rust + faer 0.22.6, Cargo.lock:398-401) cannot be compiled in this
// image (no cargo/rustc, faer is an un-vendored registry crate).  Control flow and
// per-element operation order are restated from the in-tree Rust sources cited at each
// function; only the leaf arithmetic that lives inside faer (summation order of the sparse
// matvec, of dot products and of norm_l2) is a documented choice:
//   * sparse matvec  : CSC scatter, columns ascending, y[i] += a*x[j] (mul and add rounded
//                      separately; build with -ffp-contract=off)
//   * dot / norm_l2  : 32 interleaved partial sums combined pairwise (SIMD-loop shape), norm = sqrt(sum of squares)
// PARITY PINNING: the oracle is pinned (a) by every known-answer / analytic test the reference
// holds for this path (src/algorithms/mod.rs:385-428, tests/correctness.rs:165-325,
// src/lib.rs:35-84, src/error.rs:69-129) and (b) by OUTPUTS OF THE REFERENCE ITSELF: fed with
// the restated StdRng::seed_from_u64(42) right-hand side it reproduces every row of the
// reference's results/accuracy_*.csv (src/bin/stability.rs) to 1e-14 ... 1e-10 relative on the
// printed errors (ill-conditioned tails: 3e-4) and follows results/orthogonality_*.csv
// (src/bin/orthogonality.rs) within the chaotic-amplification envelope -- see
// tests/test_oracle_reference_kats.py and tests/golden/published_curves.json.
// What stays open is only the LAST-BIT summation order inside faer's un-vendored kernels (the
// reference ships no golden alpha/beta/x vectors).
//
// All `a - c*b` updates are two roundings (sub(a, mul(c,b))), normalisation multiplies by
// the rounded reciprocal -- src/algorithms/mod.rs:183-198,277-278,312-315 and
// src/algorithms/lanczos_two_pass.rs:186-198,248-249,289-299.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

enum : int {
  ORC_OK = 0,
  ORC_BREAKDOWN = 1,          // error.rs:24-27 (never constructed by the reference)
  ORC_DIMENSION_MISMATCH = 2, // error.rs:29-35 (never constructed)
  ORC_INPUT_ERROR = 3,        // error.rs:37-38
  ORC_PARAMETER_MISMATCH = 4, // error.rs:40-45
  ORC_EVD_ERROR = 5,          // error.rs:47-48 (never constructed)
  ORC_SOLVER_ERROR = 6,       // error.rs:50-51
  ORC_PANIC = 7,              // k == 0: Vec::with_capacity(k-1) underflow, lanczos_two_pass.rs:76
  ORC_IO = 101,
  ORC_PARSE_INT = 102,
  ORC_PARSE_FLOAT = 103,
  ORC_PROBLEM_LINE_MISSING = 104,
  ORC_UNEXPECTED_EOF = 105,
  ORC_ARC_COUNT_MISMATCH = 106,
  ORC_SPARSE_CONSTRUCTION = 107,
  ORC_INVALID_NODE_INDEX = 108,
  ORC_MALFORMED_ARC_LINE = 109, // reference panics (parts[1]/parts[2] out of bounds, data_loader.rs:118-119)
};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// ---- sparse column matrix exactly as faer::sparse::SparseColMat<usize,f64> holds it -------
struct Csc {
  size_t nrows = 0, ncols = 0;
  std::vector<uint64_t> colptr;  // ncols+1
  std::vector<uint64_t> rowidx;  // nnz, ascending within a column
  std::vector<double> val;
};

struct Trip {
  uint64_t r, c;
  double v;
};

// SparseColMat::try_new_from_triplets: bounds-checked, sorted by (col,row), duplicates summed.
bool csc_from_triplets(size_t nrows, size_t ncols, std::vector<Trip>& t, Csc& out) {
  for (const Trip& e : t)
    if (e.r >= nrows || e.c >= ncols) return false;
  std::stable_sort(t.begin(), t.end(), [](const Trip& a, const Trip& b) {
    return a.c != b.c ? a.c < b.c : a.r < b.r;
  });
  out.nrows = nrows;
  out.ncols = ncols;
  out.colptr.assign(ncols + 1, 0);
  out.rowidx.clear();
  out.val.clear();
  size_t i = 0;
  while (i < t.size()) {
    size_t j = i + 1;
    double s = t[i].v;
    while (j < t.size() && t[j].c == t[i].c && t[j].r == t[i].r) s += t[j++].v;
    out.rowidx.push_back(t[i].r);
    out.val.push_back(s);
    out.colptr[t[i].c + 1]++;
    i = j;
  }
  for (size_t c = 0; c < ncols; ++c) out.colptr[c + 1] += out.colptr[c];
  return true;
}

// faer `SparseColMatRef as LinOp::apply` (call sites mod.rs:177, lanczos_two_pass.rs:186).
template <class R>
void csc_matvec(const Csc& a, const R* x, R* y) {
  for (size_t i = 0; i < a.nrows; ++i) y[i] = R(0);
  for (size_t j = 0; j < a.ncols; ++j) {
    const R xj = x[j];
    for (uint64_t p = a.colptr[j]; p < a.colptr[j + 1]; ++p) {
      R prod = R(a.val[p]) * xj;
      y[a.rowidx[p]] = y[a.rowidx[p]] + prod;
    }
  }
}

template <class R>
R dot(const R* a, const R* b, size_t n) {
  // 32 independent partial sums (the shape of a 4-way unrolled 8-lane SIMD loop), combined pairwise, then a
  // scalar tail.  faer's reduction kernels are not in the tree; the ACCURACY CLASS of this order is pinned by the
  // reference's published orthogonality curves (results/orthogonality_*.csv): with it the oracle's ||I - V^T V||_F
  // follows them within a factor 2 from k = 20 to k = 1000, a strict left-to-right sum loses orthogonality
  // 10-15 x faster (tests/test_oracle_reference_kats.py::test_published_orthogonality_curves).
  constexpr size_t L = 32;
  R acc[L];
  for (size_t l = 0; l < L; ++l) acc[l] = R(0);
  size_t i = 0;
  for (; i + L <= n; i += L)
    for (size_t l = 0; l < L; ++l) {
      R p = a[i + l] * b[i + l];
      acc[l] = acc[l] + p;
    }
  for (size_t w = L / 2; w >= 1; w /= 2)
    for (size_t l = 0; l < w; ++l) acc[l] = acc[l] + acc[l + w];
  R s = acc[0];
  for (; i < n; ++i) {
    R p = a[i] * b[i];
    s = s + p;
  }
  return s;
}
template <class R>
R norm_l2(const R* a, size_t n) {
  return std::sqrt(dot(a, a, n));
}

template <class R>
R breakdown_tol() {  // mod.rs:140-143 : 1000 * f64::EPSILON (an f64 constant even for wider R)
  return R(std::numeric_limits<double>::epsilon() * 1000.0);
}

// ---- lanczos_recurrence_step, mod.rs:167-212 ------------------------------------------------
// returns true when beta > tol (Some(beta)), false on breakdown (None)
template <class R>
bool recurrence_step(const Csc& a, R* w, const R* v_curr, const R* v_prev, R beta_prev, R& alpha,
                     R& beta) {
  const size_t n = a.nrows;
  csc_matvec(a, v_curr, w);                                       // :177
  for (size_t i = 0; i < n; ++i) { R t = beta_prev * v_prev[i]; w[i] = w[i] - t; }  // :183-186
  alpha = dot(v_curr, w, n);                                      // :191
  for (size_t i = 0; i < n; ++i) { R t = alpha * v_curr[i]; w[i] = w[i] - t; }      // :195-198
  beta = norm_l2(w, n);                                           // :202
  return !(beta <= breakdown_tol<R>());                           // :206-211
}

// ---- LanczosIteration, mod.rs:230-341 ---------------------------------------------------------
template <class R>
struct Iteration {
  const Csc* a;
  std::vector<R> v_prev, v_curr, work;
  R beta_prev = R(0);
  size_t k = 0, max_k = 0;

  int init(const Csc& op, const R* b, size_t maxk, R b_norm) {   // mod.rs:261-289
    if (b_norm <= breakdown_tol<R>())
      return fail(ORC_INPUT_ERROR, "Invalid input parameter: Input vector `b` must not be a zero vector.");
    a = &op;
    const size_t n = op.nrows;
    R inv = R(1) / b_norm;
    v_prev.assign(n, R(0));
    v_curr.resize(n);
    for (size_t i = 0; i < n; ++i) v_curr[i] = b[i] * inv;
    work.assign(n, R(0));
    max_k = maxk;
    return ORC_OK;
  }
  // next_step, mod.rs:292-340.  returns false when k >= max_k
  bool next(R& alpha, R& beta) {
    if (k >= max_k) return false;
    R bt;
    bool ok = recurrence_step(*a, work.data(), v_curr.data(), v_prev.data(), beta_prev, alpha, bt);
    k += 1;
    if (ok) {
      R inv = R(1) / bt;                                           // :312
      for (R& wi : work) wi = wi * inv;                            // :313-315
      std::swap(v_prev, v_curr);                                   // :322
      std::swap(v_curr, work);                                     // :323
      beta_prev = bt;
      beta = bt;
    } else {
      beta = R(0);                                                 // :331-338 (no rotation)
    }
    return true;
  }
};

typedef int (*orc_ftk_fn)(const double* alphas, size_t na, const double* betas, size_t nb,
                          double* y, size_t* y_len, void* user);
typedef int (*orc_step_cb)(size_t steps, const double* v, size_t ld, const double* alphas,
                           const double* betas, void* user);

// ---- lanczos_pass_one, lanczos_two_pass.rs:65-110 ---------------------------------------------
template <class R>
int pass_one(const Csc& a, const R* b, size_t k, std::vector<R>& alphas, std::vector<R>& betas,
             size_t& steps, R& b_norm) {
  if (k == 0) return fail(ORC_PANIC, "capacity overflow (k == 0: Vec::with_capacity(k - 1))");
  b_norm = norm_l2(b, a.nrows);                                    // :74
  alphas.clear();
  betas.clear();
  Iteration<R> it;
  if (int rc = it.init(a, b, k, b_norm)) return rc;
  steps = 0;
  for (size_t i = 0; i < k; ++i) {                                 // :84
    R al, be;
    if (!it.next(al, be)) break;
    alphas.push_back(al);
    steps += 1;
    if (be <= breakdown_tol<R>()) break;                           // :90-93
    if (i < k - 1) betas.push_back(be);                            // :96-98
  }
  return ORC_OK;
}

// ---- lanczos_standard, lanczos.rs:55-156 ------------------------------------------------------
// V is n x k column-major with leading dimension n; columns >= steps are left zero.
template <class R>
int standard(const Csc& a, const R* b, size_t k, R* V, std::vector<R>& alphas,
             std::vector<R>& betas, size_t& steps, R& b_norm, orc_step_cb cb, void* user) {
  if (k == 0) return fail(ORC_PANIC, "capacity overflow (k == 0: Vec::with_capacity(k - 1))");
  const size_t n = a.nrows;
  b_norm = norm_l2(b, n);                                          // :65
  std::fill(V, V + n * k, R(0));                                   // :70
  alphas.clear();
  betas.clear();
  Iteration<R> it;
  if (int rc = it.init(a, b, k, b_norm)) return rc;
  std::copy(it.v_curr.begin(), it.v_curr.end(), V);                // :81-82
  steps = 0;
  for (size_t i = 0; i < k; ++i) {
    R al, be;
    if (!it.next(al, be)) break;
    alphas.push_back(al);
    steps += 1;
    if (cb) {                                                      // :93-106
      std::vector<double> ad(alphas.begin(), alphas.end()), bd(betas.begin(), betas.end());
      std::vector<double> vd;
      const double* vp;
      if (sizeof(R) == sizeof(double)) {
        vp = reinterpret_cast<const double*>(V);
      } else {
        vd.assign(V, V + n * steps);
        vp = vd.data();
      }
      if (!cb(steps, vp, n, ad.data(), bd.data(), user)) break;
    }
    if (be <= breakdown_tol<R>()) break;                           // :108-111
    if (i < k - 1) {                                               // :117-123
      betas.push_back(be);
      std::copy(it.v_curr.begin(), it.v_curr.end(), V + n * (i + 1));
    }
  }
  return ORC_OK;
}

// ---- lanczos_pass_two_impl, lanczos_two_pass.rs:206-312 -----------------------------------------
template <class R>
int pass_two(const Csc& a, const R* b, const R* alphas, const R* betas, size_t steps, R b_norm,
             const R* y, size_t y_len, R* x, R* V /*nullable, n x steps*/) {
  const size_t n = a.nrows;
  if (steps != y_len) {                                            // :220-227
    char buf[160];
    snprintf(buf, sizeof buf, "Parameter mismatch: `y_k` expects size %zu, but got %zu.", steps, y_len);
    return fail(ORC_PARAMETER_MISMATCH, buf);
  }
  if (b_norm <= breakdown_tol<R>())                                // :229-235
    return fail(ORC_INPUT_ERROR,
                "Invalid input parameter: The initial vector `b` must not be a zero vector.");
  if (steps == 0) {                                                // :237-244
    std::fill(x, x + n, R(0));
    return ORC_OK;
  }
  std::vector<R> v_prev(n, R(0)), v_curr(n), work(n, R(0));
  R inv = R(1) / b_norm;
  for (size_t i = 0; i < n; ++i) v_curr[i] = b[i] * inv;           // :248-249
  for (size_t i = 0; i < n; ++i) x[i] = v_curr[i] * y[0];          // :252
  if (V) std::copy(v_curr.begin(), v_curr.end(), V);               // :254-258
  for (size_t j = 0; j + 1 < steps; ++j) {                         // :266
    R alpha_j = alphas[j], beta_j = betas[j];
    R beta_prev = j == 0 ? R(0) : betas[j - 1];
    csc_matvec(a, v_curr.data(), work.data());                     // :186
    for (size_t i = 0; i < n; ++i) { R t = beta_prev * v_prev[i]; work[i] = work[i] - t; }  // :190-192
    for (size_t i = 0; i < n; ++i) { R t = alpha_j * v_curr[i]; work[i] = work[i] - t; }    // :196-198
    R ib = R(1) / beta_j;                                          // :289
    for (size_t i = 0; i < n; ++i) work[i] = work[i] * ib;         // :290-292
    R c = y[j + 1];
    for (size_t i = 0; i < n; ++i) { R t = c * work[i]; x[i] = x[i] + t; }                  // :296-299
    std::swap(v_prev, v_curr);                                     // :302
    std::swap(v_curr, work);                                       // :303
    if (V) std::copy(v_curr.begin(), v_curr.end(), V + n * (j + 1));  // :306-308
  }
  return ORC_OK;
}

int call_ftk(orc_ftk_fn f, void* user, const std::vector<double>& al, const std::vector<double>& be,
             size_t steps, std::vector<double>& y) {
  // closure call + shape validation, solvers.rs:71-87 / :155-165
  y.assign(steps + 64, 0.0);
  size_t ylen = steps;
  int rc = f(al.data(), al.size(), be.data(), be.size(), y.data(), &ylen, user);
  if (rc != 0) {
    char buf[96];
    snprintf(buf, sizeof buf, "The user-provided f(T_k) solver failed: callback returned %d", rc);
    return fail(ORC_SOLVER_ERROR, buf);
  }
  if (ylen != steps) {
    char buf[160];
    snprintf(buf, sizeof buf, "Parameter mismatch: `y_k_prime` expects size %zu, but got %zu.", steps, ylen);
    return fail(ORC_PARAMETER_MISMATCH, buf);
  }
  y.resize(steps);
  return ORC_OK;
}

// ---- strict Rust-style scalar parsing -------------------------------------------------------
bool parse_usize(const std::string& s, uint64_t& out) {  // <usize as FromStr>
  size_t i = 0;
  if (s.empty()) return false;
  if (s[0] == '+') i = 1;
  if (i >= s.size()) return false;
  uint64_t v = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    uint64_t d = uint64_t(s[i] - '0');
    if (v > (UINT64_MAX - d) / 10) return false;
    v = v * 10 + d;
  }
  out = v;
  return true;
}

bool ieq(const std::string& s, size_t pos, const char* lit) {
  size_t L = strlen(lit);
  if (s.size() - pos != L) return false;
  for (size_t i = 0; i < L; ++i)
    if (tolower((unsigned char)s[pos + i]) != lit[i]) return false;
  return true;
}

bool parse_f64(const std::string& s, double& out) {  // <f64 as FromStr> (core::num::dec2flt grammar)
  if (s.empty()) return false;
  size_t i = 0;
  bool neg = false;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
  if (i >= s.size()) return false;
  if (ieq(s, i, "inf") || ieq(s, i, "infinity")) { out = neg ? -INFINITY : INFINITY; return true; }
  if (ieq(s, i, "nan")) { out = NAN; return true; }
  size_t nd = 0, j = i;
  while (j < s.size() && isdigit((unsigned char)s[j])) { ++j; ++nd; }
  if (j < s.size() && s[j] == '.') {
    ++j;
    while (j < s.size() && isdigit((unsigned char)s[j])) { ++j; ++nd; }
  }
  if (nd == 0) return false;
  if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
    ++j;
    if (j < s.size() && (s[j] == '+' || s[j] == '-')) ++j;
    size_t ne = 0;
    while (j < s.size() && isdigit((unsigned char)s[j])) { ++j; ++ne; }
    if (ne == 0) return false;
  }
  if (j != s.size()) return false;
  out = strtod(s.c_str(), nullptr);  // glibc strtod is correctly rounded, as is Rust's dec2flt
  return true;
}

// core::str::from_utf8: shortest-form encodings of scalar values only (no C0/C1 lead bytes, no overlong 3- and 4-byte
// forms, no surrogates U+D800..DFFF, nothing above U+10FFFF)
bool valid_utf8(const std::string& s) {
  size_t i = 0, n = s.size();
  while (i < n) {
    unsigned char c = s[i];
    if (c < 0x80) { ++i; continue; }
    size_t len = (c >= 0xC2 && c <= 0xDF) ? 2 : (c >> 4) == 14 ? 3 : (c >= 0xF0 && c <= 0xF4) ? 4 : 0;
    if (!len || i + len > n) return false;
    for (size_t k = 1; k < len; ++k)
      if ((((unsigned char)s[i + k]) >> 6) != 2) return false;
    const unsigned char c1 = s[i + 1];
    if ((c == 0xE0 && c1 < 0xA0) || (c == 0xED && c1 >= 0xA0) || (c == 0xF0 && c1 < 0x90) || (c == 0xF4 && c1 >= 0x90))
      return false;
    i += len;
  }
  return true;
}

// BufRead::lines(): split on '\n'; a '\r' is stripped only as part of a "\r\n" terminator (a last line without '\n'
// keeps its trailing '\r')
bool next_line(std::ifstream& f, std::string& line) {
  if (!std::getline(f, line)) return false;
  const bool terminated = !f.eof();  // getline sets eofbit when it ran into the end without finding the delimiter
  if (terminated && !line.empty() && line.back() == '\r') line.pop_back();
  return true;
}

void split_ws(const std::string& line, std::vector<std::string>& parts) {
  parts.clear();
  size_t i = 0, n = line.size();
  while (i < n) {
    while (i < n && isspace((unsigned char)line[i])) ++i;
    size_t j = i;
    while (j < n && !isspace((unsigned char)line[j])) ++j;
    if (j > i) parts.emplace_back(line, i, j - i);
    i = j;
  }
}

// ---- parse_dmx, data_loader.rs:68-156 ------------------------------------------------------------
int parse_dmx(const char* path, size_t& num_nodes, size_t& num_arcs, Csc& e) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(ORC_IO, std::string("I/O error: cannot open ") + path);
  num_nodes = num_arcs = 0;
  std::vector<Trip> trips;
  uint64_t arc_counter = 0;
  bool found = false;
  std::string line;
  std::vector<std::string> parts;
  while (next_line(f, line)) {
    if (!valid_utf8(line)) return fail(ORC_IO, "I/O error: stream did not contain valid UTF-8");
    split_ws(line, parts);
    if (parts.empty()) continue;
    if (parts[0] == "c") continue;
    if (parts[0] == "p") {                                          // :94-102
      if (parts.size() >= 4 && parts[1] == "min") {
        uint64_t a, b;
        if (!parse_usize(parts[2], a))
          return fail(ORC_PARSE_INT, "Parse error: Failed to parse integer from '" + parts[2] + "'");
        if (!parse_usize(parts[3], b))
          return fail(ORC_PARSE_INT, "Parse error: Failed to parse integer from '" + parts[3] + "'");
        num_nodes = a;
        num_arcs = b;
        found = true;
      } else {
        return fail(ORC_PROBLEM_LINE_MISSING,
                    "Format error: The 'p min' problem line was not found or was malformed.");
      }
    } else if (parts[0] == "a") {                                   // :104-134
      if (parts.size() < 3)
        return fail(ORC_MALFORMED_ARC_LINE, "panic: index out of bounds on a malformed 'a' line");
      uint64_t uv[2];
      for (int q = 0; q < 2; ++q) {
        uint64_t val;
        if (!parse_usize(parts[1 + q], val))
          return fail(ORC_PARSE_INT, "Parse error: Failed to parse integer from '" + parts[1 + q] + "'");
        if (val == 0)
          return fail(ORC_INVALID_NODE_INDEX, "Format error: Invalid node index '" + parts[1 + q] +
                                                  "'. DIMACS format requires 1-based positive integers.");
        uv[q] = val - 1;
      }
      trips.push_back({uv[0], arc_counter, 1.0});
      trips.push_back({uv[1], arc_counter, -1.0});
      arc_counter += 1;
    }
  }
  if (!found)
    return fail(ORC_PROBLEM_LINE_MISSING,
                "Format error: The 'p min' problem line was not found or was malformed.");
  if (!csc_from_triplets(num_nodes, num_arcs, trips, e))            // :152-153
    return fail(ORC_SPARSE_CONSTRUCTION,
                "Internal error: Failed to construct the sparse matrix from triplets.");
  return ORC_OK;
}

// ---- parse_qfc, data_loader.rs:166-198 -----------------------------------------------------------
int parse_qfc(const char* path, size_t expected_arcs, std::vector<double>& q) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(ORC_IO, std::string("I/O error: cannot open ") + path);
  std::string line;
  if (!next_line(f, line)) return fail(ORC_UNEXPECTED_EOF, "Format error: Unexpected end of file while reading data.");
  uint64_t m;
  if (!valid_utf8(line)) return fail(ORC_IO, "I/O error: stream did not contain valid UTF-8");
  if (!parse_usize(line, m)) return fail(ORC_PARSE_INT, "Parse error: Failed to parse integer from 'm'");
  if (m != expected_arcs) {
    char buf[160];
    snprintf(buf, sizeof buf, "Dimension mismatch: qfc file specifies %llu arcs, but dmx file has %zu.",
             (unsigned long long)m, expected_arcs);
    return fail(ORC_ARC_COUNT_MISMATCH, buf);
  }
  // lines.skip(m): Skip drops items without inspecting them (I/O errors inside skipped
  // lines are discarded by Iterator::skip's nth()).
  for (size_t i = 0; i < expected_arcs; ++i)
    if (!next_line(f, line)) break;
  q.clear();
  for (size_t i = 0; i < expected_arcs; ++i) {                      // .take(m), no length check
    if (!next_line(f, line)) break;
    if (!valid_utf8(line)) return fail(ORC_IO, "I/O error: stream did not contain valid UTF-8");
    double c;
    if (!parse_f64(line, c))
      return fail(ORC_PARSE_FLOAT, "Parse error: Failed to parse float from '" + line + "'");
    q.push_back(c);
  }
  return ORC_OK;
}

// ---- load_kkt_system, data_loader.rs:211-259 -------------------------------------------------------
int load_kkt(const char* dmx, const char* qfc, Csc& a, size_t& nodes, size_t& arcs) {
  Csc e;
  if (int rc = parse_dmx(dmx, nodes, arcs, e)) return rc;
  std::vector<double> q;
  if (int rc = parse_qfc(qfc, arcs, q)) return rc;
  const size_t n = nodes + arcs;
  std::vector<Trip> t;
  for (size_t i = 0; i < q.size(); ++i) t.push_back({i, i, q[i]});  // :226-232
  for (size_t c = 0; c < e.ncols; ++c)                               // triplet_iter over merged E
    for (uint64_t p = e.colptr[c]; p < e.colptr[c + 1]; ++p) {
      t.push_back({e.rowidx[p] + arcs, c, e.val[p]});               // :236-240
      t.push_back({c, e.rowidx[p] + arcs, e.val[p]});               // :242-246
    }
  if (!csc_from_triplets(n, n, t, a))
    return fail(ORC_SPARSE_CONSTRUCTION,
                "Internal error: Failed to construct the sparse matrix from triplets.");
  return ORC_OK;
}

template <class R>
std::vector<R> widen(const double* p, size_t n) {
  return std::vector<R>(p, p + n);
}
template <class R>
void narrow(const std::vector<R>& v, double* out) {
  for (size_t i = 0; i < v.size(); ++i) out[i] = double(v[i]);
}

}  // namespace

// =========================================== C interface ===========================================
extern "C" {

typedef struct orc_csc orc_csc;
struct orc_csc { Csc m; };

const char* orc_last_error(void) { return g_err.c_str(); }

int orc_csc_new(size_t nrows, size_t ncols, size_t ntrip, const uint64_t* rows, const uint64_t* cols,
                const double* vals, orc_csc** out) {
  std::vector<Trip> t(ntrip);
  for (size_t i = 0; i < ntrip; ++i) t[i] = {rows[i], cols[i], vals[i]};
  orc_csc* h = new orc_csc;
  if (!csc_from_triplets(nrows, ncols, t, h->m)) {
    delete h;
    return fail(ORC_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  }
  *out = h;
  return ORC_OK;
}
void orc_csc_free(orc_csc* h) { delete h; }
size_t orc_csc_nrows(const orc_csc* h) { return h->m.nrows; }
size_t orc_csc_nnz(const orc_csc* h) { return h->m.val.size(); }
void orc_csc_export(const orc_csc* h, uint64_t* colptr, uint64_t* rowidx, double* val) {
  std::copy(h->m.colptr.begin(), h->m.colptr.end(), colptr);
  std::copy(h->m.rowidx.begin(), h->m.rowidx.end(), rowidx);
  std::copy(h->m.val.begin(), h->m.val.end(), val);
}

int orc_load_kkt(const char* dmx, const char* qfc, orc_csc** out, size_t* nodes, size_t* arcs) {
  orc_csc* h = new orc_csc;
  int rc = load_kkt(dmx, qfc, h->m, *nodes, *arcs);
  if (rc) { delete h; return rc; }
  *out = h;
  return ORC_OK;
}

void orc_matvec(const orc_csc* a, const double* x, double* y) { csc_matvec(a->m, x, y); }

// alphas must hold k entries, betas k-1 (k>=1) entries.
int orc_pass_one(const orc_csc* a, const double* b, size_t k, double* alphas, double* betas,
                 size_t* steps, double* b_norm) {
  std::vector<double> al, be;
  int rc = pass_one<double>(a->m, b, k, al, be, *steps, *b_norm);
  if (rc) return rc;
  std::copy(al.begin(), al.end(), alphas);
  std::copy(be.begin(), be.end(), betas);
  return ORC_OK;
}

int orc_pass_two(const orc_csc* a, const double* b, const double* alphas, const double* betas,
                 size_t steps, double b_norm, const double* y, size_t y_len, double* x, double* V) {
  return pass_two<double>(a->m, b, alphas, betas, steps, b_norm, y, y_len, x, V);
}

int orc_standard(const orc_csc* a, const double* b, size_t k, double* V, double* alphas,
                 double* betas, size_t* steps, double* b_norm, orc_step_cb cb, void* user) {
  std::vector<double> al, be;
  int rc = standard<double>(a->m, b, k, V, al, be, *steps, *b_norm, cb, user);
  if (rc) return rc;
  std::copy(al.begin(), al.end(), alphas);
  std::copy(be.begin(), be.end(), betas);
  return ORC_OK;
}

// solvers::lanczos, solvers.rs:46-107
int orc_lanczos(const orc_csc* a, const double* b, size_t k, orc_ftk_fn f, void* user, double* x) {
  const size_t n = a->m.nrows;
  if (k == 0) return fail(ORC_PANIC, "capacity overflow (k == 0: Vec::with_capacity(k - 1))");
  std::vector<double> V(n * k), al, be, y;
  size_t steps;
  double bn;
  if (int rc = standard<double>(a->m, b, k, V.data(), al, be, steps, bn, nullptr, nullptr)) return rc;
  if (steps == 0) { std::fill(x, x + n, 0.0); return ORC_OK; }     // :65-67
  if (int rc = call_ftk(f, user, al, be, steps, y)) return rc;
  // matmul(x, Replace, V, y', alpha = b_norm), solvers.rs:96-104 : x = b_norm * (V y')
  for (size_t i = 0; i < n; ++i) x[i] = 0.0;
  for (size_t j = 0; j < steps; ++j) {
    const double* col = V.data() + n * j;
    for (size_t i = 0; i < n; ++i) { double t = col[i] * y[j]; x[i] = x[i] + t; }
  }
  for (size_t i = 0; i < n; ++i) x[i] = bn * x[i];
  return ORC_OK;
}

// solvers::lanczos_two_pass, solvers.rs:133-175
int orc_lanczos_two_pass(const orc_csc* a, const double* b, size_t k, orc_ftk_fn f, void* user,
                         double* x) {
  const size_t n = a->m.nrows;
  std::vector<double> al, be, y;
  size_t steps;
  double bn;
  if (int rc = pass_one<double>(a->m, b, k, al, be, steps, bn)) return rc;
  if (steps == 0) { std::fill(x, x + n, 0.0); return ORC_OK; }     // :150-152
  if (int rc = call_ftk(f, user, al, be, steps, y)) return rc;
  for (double& yi : y) yi = yi * bn;                               // :169
  return pass_two<double>(a->m, b, al.data(), be.data(), steps, bn, y.data(), y.size(), x, nullptr);
}

// ---- extended-precision (long double, 64-bit mantissa on x86) variants for error-ball arguments ----
int orc_pass_one_ld(const orc_csc* a, const double* b, size_t k, double* alphas, double* betas,
                    size_t* steps, double* b_norm) {
  std::vector<long double> bl = widen<long double>(b, a->m.nrows), al, be;
  long double bn;
  int rc = pass_one<long double>(a->m, bl.data(), k, al, be, *steps, bn);
  if (rc) return rc;
  narrow(al, alphas);
  narrow(be, betas);
  *b_norm = double(bn);
  return ORC_OK;
}

int orc_lanczos_two_pass_ld(const orc_csc* a, const double* b, size_t k, orc_ftk_fn f, void* user,
                            double* x) {
  const size_t n = a->m.nrows;
  std::vector<long double> bl = widen<long double>(b, n), al, be;
  size_t steps;
  long double bn;
  if (int rc = pass_one<long double>(a->m, bl.data(), k, al, be, steps, bn)) return rc;
  if (steps == 0) { std::fill(x, x + n, 0.0); return ORC_OK; }
  std::vector<double> ad(al.begin(), al.end()), bd(be.begin(), be.end()), y;
  if (int rc = call_ftk(f, user, ad, bd, steps, y)) return rc;
  std::vector<long double> yl(y.begin(), y.end()), xl(n);
  for (auto& yi : yl) yi = yi * bn;
  int rc = pass_two<long double>(a->m, bl.data(), al.data(), be.data(), steps, bn, yl.data(), yl.size(),
                                 xl.data(), nullptr);
  if (rc) return rc;
  narrow(xl, x);
  return ORC_OK;
}

}  // extern "C"
