// tpl_loader.cpp -- host-side `.dmx` / `.qfc` KKT loader (component H2).
//
// Replaces utils::data_loader::{parse_dmx, parse_qfc, load_kkt_system}
// (src/utils/data_loader.rs:68-156, 166-198, 211-259) with the SAME observable semantics:
//   * `.dmx`: lines are whitespace-tokenised; first token "c" -> skipped, "p" -> must be
//     `p min <nodes> <arcs>` (else ProblemLineMissing), "a" -> tokens 1,2 are 1-based tail/head
//     (0 -> InvalidDimacsNodeIndex, non-integer -> ParseInt), anything else ignored; the j-th `a`
//     line is arc j; +1 at the tail row, -1 at the head row; duplicates are summed (a self-loop
//     becomes one explicit 0); out-of-range indices -> SparseMatrixConstructionError.
//   * `.qfc`: line 0 is m (no trimming), the next m LINES are skipped, then up to m lines are read as
//     one f64 each -- with NO check that m values arrived, so qfcgen's 3-line layout yields an empty D
//     (SURVEY C2).  Reproduced on purpose: this is a drop-in.
//   * A = [[D, E^T], [E, 0]], n = nodes + arcs, arcs first (data_loader.rs:222-248).
// The whole file is read once and scanned in place (the reference goes line by line through a
// BufReader); a 500k-arc pair loads in tens of milliseconds.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "tpl_internal.h"

namespace {

using tpl::fail;

bool read_file(const char* path, std::string& out, std::string& why) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    why = strerror(errno);
    return false;
  }
  char buf[1 << 16];
  size_t got;
  out.clear();
  while ((got = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
  bool ok = !ferror(f);
  if (!ok) why = strerror(errno);
  fclose(f);
  return ok;
}

// iterates `BufRead::lines()`: split at '\n', drop one trailing '\r'
struct LineCursor {
  const char* p;
  const char* end;
  bool next(const char*& b, const char*& e) {
    if (p >= end) return false;
    const char* nl = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
    b = p;
    e = nl ? nl : end;
    p = nl ? nl + 1 : end;
    if (e > b && e[-1] == '\r') --e;
    return true;
  }
};

bool utf8_ok(const char* b, const char* e) {
  const unsigned char* s = reinterpret_cast<const unsigned char*>(b);
  size_t n = size_t(e - b), i = 0;
  while (i < n) {
    unsigned char c = s[i];
    if (c < 0x80) { ++i; continue; }
    size_t len = (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 0;
    if (!len || i + len > n) return false;
    for (size_t k = 1; k < len; ++k)
      if ((s[i + k] >> 6) != 0x2) return false;
    i += len;
  }
  return true;
}

inline bool is_ws(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// <usize as FromStr>: optional '+', decimal digits, overflow is an error
bool rust_usize(const char* b, const char* e, uint64_t& out) {
  if (b < e && *b == '+') ++b;
  if (b >= e) return false;
  uint64_t v = 0;
  for (; b < e; ++b) {
    if (*b < '0' || *b > '9') return false;
    uint64_t dgt = uint64_t(*b - '0');
    if (v > (UINT64_MAX - dgt) / 10) return false;
    v = v * 10 + dgt;
  }
  out = v;
  return true;
}

bool lit_ci(const char* b, const char* e, const char* lit) {
  size_t n = strlen(lit);
  if (size_t(e - b) != n) return false;
  for (size_t i = 0; i < n; ++i)
    if ((b[i] | 0x20) != lit[i]) return false;
  return true;
}

// <f64 as FromStr>: [+-]? ( inf | infinity | nan | digits [. digits] | . digits ) [ (e|E) [+-]? digits ]
bool rust_f64(const char* b, const char* e, double& out) {
  const char* s = b;
  if (s >= e) return false;
  bool neg = false;
  if (*s == '+' || *s == '-') { neg = *s == '-'; ++s; }
  if (s >= e) return false;
  if (lit_ci(s, e, "inf") || lit_ci(s, e, "infinity")) { out = neg ? -INFINITY : INFINITY; return true; }
  if (lit_ci(s, e, "nan")) { out = NAN; return true; }
  size_t digits = 0;
  while (s < e && *s >= '0' && *s <= '9') { ++s; ++digits; }
  if (s < e && *s == '.') {
    ++s;
    while (s < e && *s >= '0' && *s <= '9') { ++s; ++digits; }
  }
  if (!digits) return false;
  if (s < e && (*s == 'e' || *s == 'E')) {
    ++s;
    if (s < e && (*s == '+' || *s == '-')) ++s;
    size_t ed = 0;
    while (s < e && *s >= '0' && *s <= '9') { ++s; ++ed; }
    if (!ed) return false;
  }
  if (s != e) return false;
  std::string tmp(b, e);
  out = strtod(tmp.c_str(), nullptr);  // correctly rounded, like core::num::dec2flt
  return true;
}

struct Trip {
  uint64_t r, c;
  double v;
};

// SparseColMat::try_new_from_triplets semantics: bounds check, (col,row) order, duplicates summed.
bool build_csc(size_t nrows, size_t ncols, std::vector<Trip>& t, std::vector<uint64_t>& colptr,
               std::vector<uint64_t>& rowidx, std::vector<double>& val) {
  for (const Trip& x : t)
    if (x.r >= nrows || x.c >= ncols) return false;
  std::stable_sort(t.begin(), t.end(),
                   [](const Trip& a, const Trip& b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
  colptr.assign(ncols + 1, 0);
  rowidx.clear();
  val.clear();
  rowidx.reserve(t.size());
  val.reserve(t.size());
  for (size_t i = 0; i < t.size();) {
    size_t j = i + 1;
    double s = t[i].v;
    while (j < t.size() && t[j].c == t[i].c && t[j].r == t[i].r) s += t[j++].v;
    rowidx.push_back(t[i].r);
    val.push_back(s);
    ++colptr[t[i].c + 1];
    i = j;
  }
  for (size_t c = 0; c < ncols; ++c) colptr[c + 1] += colptr[c];
  return true;
}

int parse_dmx(const char* path, tpl_kkt& k, std::vector<uint64_t>& tails, std::vector<uint64_t>& heads) {
  std::string text, why;
  if (!read_file(path, text, why)) return fail(TPL_ERR_IO, "I/O error: %s", why.c_str());
  LineCursor cur{text.data(), text.data() + text.size()};
  const char *b, *e;
  bool found = false;
  std::vector<std::pair<const char*, const char*>> tok;
  while (cur.next(b, e)) {
    if (!utf8_ok(b, e)) return fail(TPL_ERR_IO, "I/O error: stream did not contain valid UTF-8");
    tok.clear();
    for (const char* s = b; s < e;) {
      while (s < e && is_ws(*s)) ++s;
      const char* t = s;
      while (t < e && !is_ws(*t)) ++t;
      if (t > s) tok.emplace_back(s, t);
      s = t;
    }
    if (tok.empty()) continue;
    const size_t l0 = size_t(tok[0].second - tok[0].first);
    const char c0 = l0 == 1 ? *tok[0].first : '\0';
    if (c0 == 'c') continue;
    if (c0 == 'p') {
      if (tok.size() >= 4 && size_t(tok[1].second - tok[1].first) == 3 && !memcmp(tok[1].first, "min", 3)) {
        uint64_t nn, na;
        if (!rust_usize(tok[2].first, tok[2].second, nn))
          return fail(TPL_ERR_PARSE_INT, "Parse error: Failed to parse integer from '%s'",
                      std::string(tok[2].first, tok[2].second).c_str());
        if (!rust_usize(tok[3].first, tok[3].second, na))
          return fail(TPL_ERR_PARSE_INT, "Parse error: Failed to parse integer from '%s'",
                      std::string(tok[3].first, tok[3].second).c_str());
        k.num_nodes = nn;
        k.num_arcs = na;
        found = true;
      } else {
        return fail(TPL_ERR_PROBLEM_LINE_MISSING,
                    "Format error: The 'p min' problem line was not found or was malformed.");
      }
    } else if (c0 == 'a') {
      if (tok.size() < 3)
        return fail(TPL_ERR_MALFORMED_ARC_LINE,
                    "Format error: arc line has fewer than 3 fields (the reference panics here).");
      uint64_t uv[2];
      for (int q = 0; q < 2; ++q) {
        uint64_t v;
        std::string s(tok[1 + q].first, tok[1 + q].second);
        if (!rust_usize(tok[1 + q].first, tok[1 + q].second, v))
          return fail(TPL_ERR_PARSE_INT, "Parse error: Failed to parse integer from '%s'", s.c_str());
        if (v == 0)
          return fail(TPL_ERR_INVALID_NODE_INDEX,
                      "Format error: Invalid node index '%s'. DIMACS format requires 1-based positive integers.",
                      s.c_str());
        uv[q] = v - 1;
      }
      tails.push_back(uv[0]);
      heads.push_back(uv[1]);
    }
  }
  if (!found)
    return fail(TPL_ERR_PROBLEM_LINE_MISSING,
                "Format error: The 'p min' problem line was not found or was malformed.");
  std::vector<Trip> t;
  t.reserve(2 * tails.size());
  for (size_t j = 0; j < tails.size(); ++j) {
    t.push_back({tails[j], j, 1.0});
    t.push_back({heads[j], j, -1.0});
  }
  if (!build_csc(k.num_nodes, k.num_arcs, t, k.e_colptr, k.e_rowidx, k.e_val))
    return fail(TPL_ERR_SPARSE_CONSTRUCTION,
                "Internal error: Failed to construct the sparse matrix from triplets.");
  return TPL_OK;
}

int parse_qfc(const char* path, size_t expected_arcs, std::vector<double>& costs) {
  std::string text, why;
  if (!read_file(path, text, why)) return fail(TPL_ERR_IO, "I/O error: %s", why.c_str());
  LineCursor cur{text.data(), text.data() + text.size()};
  const char *b, *e;
  if (!cur.next(b, e))
    return fail(TPL_ERR_UNEXPECTED_EOF, "Format error: Unexpected end of file while reading data.");
  if (!utf8_ok(b, e)) return fail(TPL_ERR_IO, "I/O error: stream did not contain valid UTF-8");
  uint64_t m;
  if (!rust_usize(b, e, m)) return fail(TPL_ERR_PARSE_INT, "Parse error: Failed to parse integer from 'm'");
  if (m != expected_arcs)
    return fail(TPL_ERR_ARC_COUNT_MISMATCH,
                "Dimension mismatch: qfc file specifies %llu arcs, but dmx file has %zu.",
                (unsigned long long)m, expected_arcs);
  for (size_t i = 0; i < expected_arcs; ++i)  // lines.skip(m): contents never inspected
    if (!cur.next(b, e)) break;
  costs.clear();
  for (size_t i = 0; i < expected_arcs; ++i) {  // .take(m) with no length check
    if (!cur.next(b, e)) break;
    if (!utf8_ok(b, e)) return fail(TPL_ERR_IO, "I/O error: stream did not contain valid UTF-8");
    double c;
    if (!rust_f64(b, e, c))
      return fail(TPL_ERR_PARSE_FLOAT, "Parse error: Failed to parse float from '%s'", std::string(b, e).c_str());
    costs.push_back(c);
  }
  return TPL_OK;
}

}  // namespace

extern "C" {

int tpl_load_kkt(const char* dmx_path, const char* qfc_path, tpl_kkt** out) {
  tpl::clear_error();
  if (!dmx_path || !qfc_path || !out) return fail(TPL_ERR_PANIC, "null argument");
  tpl_kkt* k = new tpl_kkt;
  std::vector<uint64_t> tails, heads;
  int rc = parse_dmx(dmx_path, *k, tails, heads);
  if (!rc) rc = parse_qfc(qfc_path, k->num_arcs, k->costs);
  if (rc) {
    delete k;
    return rc;
  }
  const size_t m = k->num_arcs, p = k->num_nodes, n = m + p;
  std::vector<Trip> t;
  t.reserve(k->costs.size() + 2 * k->e_val.size());
  for (size_t i = 0; i < k->costs.size(); ++i) t.push_back({i, i, k->costs[i]});
  for (size_t c = 0; c < m; ++c)
    for (uint64_t q = k->e_colptr[c]; q < k->e_colptr[c + 1]; ++q) {
      t.push_back({k->e_rowidx[q] + m, c, k->e_val[q]});
      t.push_back({c, k->e_rowidx[q] + m, k->e_val[q]});
    }
  if (!build_csc(n, n, t, k->colptr, k->rowidx, k->val)) {
    delete k;
    return fail(TPL_ERR_SPARSE_CONSTRUCTION,
                "Internal error: Failed to construct the sparse matrix from triplets.");
  }
  // incidence view: exact iff every arc column of E came from exactly one `a` line (self-loops are fine:
  // their merged explicit 0 contributes nothing and the incidence kernels skip them)
  k->regular = tails.size() == m && m <= 0x7fffffffu && p <= 0x7fffffffu;
  k->tail.assign(m, 0);
  k->head.assign(m, 0);
  k->d.assign(m, 0.0);
  for (size_t i = 0; i < k->costs.size(); ++i) k->d[i] = k->costs[i];
  for (size_t j = 0; j < std::min(m, tails.size()); ++j) {
    k->tail[j] = uint32_t(tails[j]);
    k->head[j] = uint32_t(heads[j]);
  }
  *out = k;
  return TPL_OK;
}

void tpl_kkt_free(tpl_kkt* kkt) { delete kkt; }
size_t tpl_kkt_num_nodes(const tpl_kkt* kkt) { return kkt->num_nodes; }
size_t tpl_kkt_num_arcs(const tpl_kkt* kkt) { return kkt->num_arcs; }
size_t tpl_kkt_num_costs(const tpl_kkt* kkt) { return kkt->costs.size(); }
size_t tpl_kkt_nnz(const tpl_kkt* kkt) { return kkt->val.size(); }

int tpl_kkt_csc(const tpl_kkt* kkt, size_t* n, size_t* nnz, const uint64_t** colptr, const uint64_t** rowidx,
                const double** val) {
  if (n) *n = kkt->num_nodes + kkt->num_arcs;
  if (nnz) *nnz = kkt->val.size();
  if (colptr) *colptr = kkt->colptr.data();
  if (rowidx) *rowidx = kkt->rowidx.data();
  if (val) *val = kkt->val.data();
  return TPL_OK;
}

int tpl_kkt_incidence(const tpl_kkt* kkt, const uint32_t** tail, const uint32_t** head, const double** d,
                      size_t* d_len, int* regular) {
  if (tail) *tail = kkt->tail.data();
  if (head) *head = kkt->head.data();
  if (d) *d = kkt->d.data();
  if (d_len) *d_len = kkt->costs.size();
  if (regular) *regular = kkt->regular ? 1 : 0;
  return TPL_OK;
}

}  // extern "C"
