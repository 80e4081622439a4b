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
// BufReader).
//
// the binary instance container (`tpl_write_kkt_binary` / `tpl_load_kkt_binary`).  A text pair of a
// 50M-arc instance is ~1.5 GB and parse-bound; the container holds the incidence view as raw little-endian arrays
// (12 bytes per arc + 8 per cost) behind a 64-byte header with a checksum and loads at file-read speed.  The CSC of
// KKTSystem.a is then built on first use, directly from the arc list in O(nnz) (no triplet sort), entry for entry what
// the triplet path produces (tests/test_loader_cpu.py compares the two).
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
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
  out.clear();
  if (fseek(f, 0, SEEK_END) == 0) {  // regular file: one read into a buffer of the right size
    const long size = ftell(f);
    if (size > 0) out.reserve(size_t(size));
    rewind(f);
  }
  std::vector<char> buf(1 << 20);
  size_t got;
  while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) out.append(buf.data(), got);
  bool ok = !ferror(f);
  if (!ok) why = strerror(errno);
  fclose(f);
  return ok;
}

// iterates `BufRead::lines()`: split at '\n'; a '\r' goes only as part of a "\r\n" terminator (a last line that has no
// '\n' keeps its trailing '\r', which then fails `parse::<usize>()` / `parse::<f64>()` as in the reference)
struct LineCursor {
  const char* p;
  const char* end;
  bool next(const char*& b, const char*& e) {
    if (p >= end) return false;
    const char* nl = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
    b = p;
    e = nl ? nl : end;
    p = nl ? nl + 1 : end;
    if (nl && e > b && e[-1] == '\r') --e;
    return true;
  }
};

// core::str::from_utf8: shortest-form encodings of scalar values only (C0/C1 and F5..FF never start a character, no
// overlong 3- / 4-byte forms, no surrogates, nothing above U+10FFFF)
bool utf8_ok(const char* b, const char* e) {
  const unsigned char* s = reinterpret_cast<const unsigned char*>(b);
  size_t n = size_t(e - b), i = 0;
  while (i < n) {
    unsigned char c = s[i];
    if (c < 0x80) { ++i; continue; }
    size_t len = (c >= 0xC2 && c <= 0xDF) ? 2 : (c >> 4) == 0xE ? 3 : (c >= 0xF0 && c <= 0xF4) ? 4 : 0;
    if (!len || i + len > n) return false;
    for (size_t k = 1; k < len; ++k)
      if ((s[i + k] >> 6) != 0x2) return false;
    const unsigned char c1 = s[i + 1];
    if ((c == 0xE0 && c1 < 0xA0) || (c == 0xED && c1 >= 0xA0) || (c == 0xF0 && c1 < 0x90) || (c == 0xF4 && c1 >= 0x90))
      return false;
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
  // correctly rounded, like core::num::dec2flt; the syntax was validated above (from_chars takes no leading '+')
  const char* first = (*b == '+') ? b + 1 : b;
  const std::from_chars_result r = std::from_chars(first, e, out, std::chars_format::general);
  if (r.ec == std::errc::result_out_of_range) {  // dec2flt saturates: overflow -> inf, underflow -> 0
    std::string tmp(b, e);
    out = strtod(tmp.c_str(), nullptr);
  }
  return true;
}

// Fast path of a cost line: only the characters of a plain decimal ([0-9 . e E -]) and std::from_chars consumes the whole
// line without a range error.  Under that alphabet from_chars and <f64 as FromStr> accept exactly the same strings (both
// follow `-? (digits [. digits*] | . digits) ([eE] -? digits)?`), and both round correctly.
inline bool fast_f64_line(const char* b, const char* e, double& out) {
  if (b >= e) return false;
  for (const char* s = b; s < e; ++s) {
    const char ch = *s;
    if (!((ch >= '0' && ch <= '9') || ch == '.' || ch == 'e' || ch == 'E' || ch == '-')) return false;
  }
  const std::from_chars_result r = std::from_chars(b, e, out, std::chars_format::general);
  return r.ec == std::errc() && r.ptr == e;
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

// Fast path for the overwhelmingly common line `a <tail> <head> [ignored tokens]`, pure ASCII, both indices plain positive
// decimals that fit: appends the arc and returns true.  Returns false WITHOUT side effects for every other line (other
// kinds, malformed or zero indices, overflow, non-ASCII bytes anywhere), which then takes the general path and gets the
// reference's exact error.
inline bool fast_arc_line(const char* b, const char* e, std::vector<uint64_t>& tails, std::vector<uint64_t>& heads) {
  const char* s = b;
  while (s < e && is_ws(*s)) ++s;
  if (s + 1 >= e || *s != 'a' || !is_ws(s[1])) return false;
  ++s;
  uint64_t uv[2];
  for (int q = 0; q < 2; ++q) {
    while (s < e && is_ws(*s)) ++s;
    if (s < e && *s == '+') ++s;
    const char* t = s;
    uint64_t v = 0;
    while (t < e && *t >= '0' && *t <= '9') {
      if (t - s >= 18) return false;  // (cannot overflow below 19 digits; longer ones go the general way)
      v = v * 10 + uint64_t(*t - '0');
      ++t;
    }
    if (t == s || v == 0 || (t < e && !is_ws(*t))) return false;
    uv[q] = v - 1;
    s = t;
  }
  for (const char* t = s; t < e; ++t)
    if (static_cast<unsigned char>(*t) >= 0x80) return false;  // UTF-8 validity of the ignored rest is the general path's job
  tails.push_back(uv[0]);
  heads.push_back(uv[1]);
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
    if (fast_arc_line(b, e, tails, heads)) continue;  // the plain `a <tail> <head> ...` line; anything else: general path
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
        if (!rust_usize(tok[1 + q].first, tok[1 + q].second, v))
          return fail(TPL_ERR_PARSE_INT, "Parse error: Failed to parse integer from '%s'",
                      std::string(tok[1 + q].first, tok[1 + q].second).c_str());
        if (v == 0)
          return fail(TPL_ERR_INVALID_NODE_INDEX,
                      "Format error: Invalid node index '%s'. DIMACS format requires 1-based positive integers.",
                      std::string(tok[1 + q].first, tok[1 + q].second).c_str());
        uv[q] = v - 1;
      }
      tails.push_back(uv[0]);
      heads.push_back(uv[1]);
    }
  }
  if (!found)
    return fail(TPL_ERR_PROBLEM_LINE_MISSING,
                "Format error: The 'p min' problem line was not found or was malformed.");
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
    double c;
    if (fast_f64_line(b, e, c)) {  // plain decimal spelling; everything else is validated the long way
      costs.push_back(c);
      continue;
    }
    if (!utf8_ok(b, e)) return fail(TPL_ERR_IO, "I/O error: stream did not contain valid UTF-8");
    if (!rust_f64(b, e, c))
      return fail(TPL_ERR_PARSE_FLOAT, "Parse error: Failed to parse float from '%s'", std::string(b, e).c_str());
    costs.push_back(c);
  }
  return TPL_OK;
}


// ---------------------------------------------------------------- binary container
constexpr char kMagic[8] = {'T', 'P', 'L', 'K', 'K', 'T', '1', '\n'};
struct BinHeader {  // 64 bytes, little endian
  char magic[8];
  uint32_t version;  // 1
  uint32_t flags;    // reserved, 0
  uint64_t nodes, arcs, n_costs;
  uint64_t checksum;  // word_hash over the payload as written
  uint64_t reserved[2];
};
static_assert(sizeof(BinHeader) == 64, "header layout");

// 4-lane multiply-xorshift hash over 8-byte words (a 1.2 GB payload hashes in a fraction of its read time)
uint64_t word_hash(const unsigned char* p, size_t bytes, uint64_t seed) {
  uint64_t h[4] = {seed ^ 0x9E3779B97F4A7C15ull, seed ^ 0xC2B2AE3D27D4EB4Full, seed ^ 0x165667B19E3779F9ull,
                   seed ^ 0x27D4EB2F165667C5ull};
  size_t i = 0;
  for (; i + 32 <= bytes; i += 32) {
    uint64_t w[4];
    memcpy(w, p + i, 32);
    for (int q = 0; q < 4; ++q) {
      h[q] = (h[q] ^ w[q]) * 0xFF51AFD7ED558CCDull;
      h[q] ^= h[q] >> 29;
    }
  }
  uint64_t t = bytes;
  for (; i < bytes; ++i) t = (t ^ p[i]) * 0x100000001B3ull;
  for (int q = 0; q < 4; ++q) {
    t = (t ^ h[q]) * 0xC4CEB9FE1A85EC53ull;
    t ^= t >> 32;
  }
  return t;
}

size_t pad8(size_t bytes) { return (bytes + 7) & ~size_t(7); }

// E and A = [[D,E^T],[E,0]] straight from the arc list, entry for entry what build_csc makes of the loader's triplets:
// column j < m holds D_jj (when the .qfc gave it), then the tail / head rows in ascending order (a self-loop's +1 and -1
// merge into one explicit 0); column m+i holds the arcs at node i in ascending arc order.
void csc_from_incidence(tpl_kkt& k) {
  const size_t m = k.num_arcs, p = k.num_nodes, n = m + p, nd = k.costs.size();
  k.e_colptr.assign(m + 1, 0);
  k.e_rowidx.clear();
  k.e_val.clear();
  k.e_rowidx.reserve(2 * m);
  k.e_val.reserve(2 * m);
  std::vector<uint64_t> deg(p + 1, 0);
  for (size_t j = 0; j < m; ++j) {
    const uint64_t t = k.tail[j], h = k.head[j];
    if (t == h) {
      k.e_rowidx.push_back(t);
      k.e_val.push_back(1.0 + -1.0);
      ++deg[t + 1];
    } else {
      const bool tf = t < h;
      k.e_rowidx.push_back(tf ? t : h);
      k.e_val.push_back(tf ? 1.0 : -1.0);
      k.e_rowidx.push_back(tf ? h : t);
      k.e_val.push_back(tf ? -1.0 : 1.0);
      ++deg[t + 1];
      ++deg[h + 1];
    }
    k.e_colptr[j + 1] = k.e_rowidx.size();
  }
  const size_t nnz = nd + 2 * k.e_rowidx.size();
  k.colptr.assign(n + 1, 0);
  k.rowidx.assign(nnz, 0);
  k.val.assign(nnz, 0.0);
  size_t q = 0;
  for (size_t j = 0; j < m; ++j) {
    if (j < nd) {
      k.rowidx[q] = j;
      k.val[q++] = k.costs[j];
    }
    for (uint64_t e = k.e_colptr[j]; e < k.e_colptr[j + 1]; ++e) {
      k.rowidx[q] = k.e_rowidx[e] + m;
      k.val[q++] = k.e_val[e];
    }
    k.colptr[j + 1] = q;
  }
  for (size_t i = 0; i < p; ++i) {
    deg[i + 1] += deg[i];
    k.colptr[m + i + 1] = q + deg[i + 1];
  }
  std::vector<uint64_t> fill(deg.begin(), deg.end() - 1);
  for (size_t j = 0; j < m; ++j)
    for (uint64_t e = k.e_colptr[j]; e < k.e_colptr[j + 1]; ++e) {
      const size_t at = q + fill[k.e_rowidx[e]]++;
      k.rowidx[at] = j;
      k.val[at] = k.e_val[e];
    }
}

std::mutex g_csc_mutex;

}  // namespace

namespace tpl {
// the CSC views of a container-loaded system are built on first use
void kkt_ensure_csc(const tpl_kkt* kkt) {
  std::lock_guard<std::mutex> lock(g_csc_mutex);
  if (!kkt->lazy_csc) return;
  tpl_kkt& k = *const_cast<tpl_kkt*>(kkt);
  csc_from_incidence(k);
  k.lazy_csc = false;
}
}  // namespace tpl

extern "C" {

int tpl_load_kkt(const char* dmx_path, const char* qfc_path, tpl_kkt** out) {
  tpl::clear_error();
  if (!dmx_path || !qfc_path || !out) return fail(TPL_ERR_PANIC, "null argument");
  tpl_kkt* k = new tpl_kkt;
  std::vector<uint64_t> tails, heads;
  int rc = parse_dmx(dmx_path, *k, tails, heads);
  const size_t m = k->num_arcs, p = k->num_nodes, n = m + p;
  if (!rc) {  // E is assembled inside parse_dmx in the reference (data_loader.rs:139-155): its bounds error comes first
    bool in_range = tails.size() <= m;  // the j-th `a` line is column j of E
    for (size_t j = 0; j < tails.size() && in_range; ++j) in_range = tails[j] < p && heads[j] < p;
    if (!in_range)
      rc = fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  }
  if (!rc) rc = parse_qfc(qfc_path, k->num_arcs, k->costs);
  if (rc) {
    delete k;
    return rc;
  }
  // incidence view: exact iff every arc column of E came from exactly one `a` line (self-loops are fine:
  // their merged explicit 0 contributes nothing and the incidence kernels skip them)
  k->regular = tails.size() == m && m <= 0x7fffffffu && p <= 0x7fffffffu;
  k->tail.assign(m, 0);
  k->head.assign(m, 0);
  k->d.assign(m, 0.0);
  for (size_t i = 0; i < k->costs.size(); ++i) k->d[i] = k->costs[i];
  for (size_t j = 0; j < tails.size(); ++j) {
    k->tail[j] = uint32_t(tails[j]);
    k->head[j] = uint32_t(heads[j]);
  }
  if (k->regular) {
    k->lazy_csc = true;  // plain arc list: the CSC views come straight from it, on first use (csc_from_incidence)
  } else {
    // fewer `a` lines than announced (only a debug_assert in the reference, data_loader.rs:145-148): the general
    // triplet route of SparseColMat::try_new_from_triplets
    std::vector<Trip> t;
    t.reserve(2 * tails.size());
    for (size_t j = 0; j < tails.size(); ++j) {
      t.push_back({tails[j], j, 1.0});
      t.push_back({heads[j], j, -1.0});
    }
    bool ok = build_csc(p, m, t, k->e_colptr, k->e_rowidx, k->e_val);
    t.clear();
    t.reserve(k->costs.size() + 2 * k->e_val.size());
    for (size_t i = 0; i < k->costs.size(); ++i) t.push_back({i, i, k->costs[i]});
    for (size_t c = 0; ok && c < m; ++c)
      for (uint64_t q = k->e_colptr[c]; q < k->e_colptr[c + 1]; ++q) {
        t.push_back({k->e_rowidx[q] + m, c, k->e_val[q]});
        t.push_back({c, k->e_rowidx[q] + m, k->e_val[q]});
      }
    ok = ok && build_csc(n, n, t, k->colptr, k->rowidx, k->val);
    if (!ok) {
      delete k;
      return fail(TPL_ERR_SPARSE_CONSTRUCTION,
                  "Internal error: Failed to construct the sparse matrix from triplets.");
    }
  }
  *out = k;
  return TPL_OK;
}


int tpl_write_kkt_binary(const char* path, size_t nodes, size_t arcs, const uint32_t* tail, const uint32_t* head,
                         const double* costs, size_t n_costs) {
  tpl::clear_error();
  if (!path || (arcs && (!tail || !head)) || (n_costs && !costs)) return fail(TPL_ERR_PANIC, "null argument");
  if (n_costs > arcs) return fail(TPL_ERR_ARC_COUNT_MISMATCH, "Dimension mismatch: %zu costs for %zu arcs.", n_costs, arcs);
  if (arcs > 0x7fffffffu || nodes > 0x7fffffffu)
    return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: the container holds at most 2^31-1 arcs and nodes.");
  for (size_t j = 0; j < arcs; ++j)
    if (tail[j] >= nodes || head[j] >= nodes)
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  const size_t idx_bytes = pad8(4 * arcs);
  BinHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kMagic, 8);
  h.version = 1;
  h.nodes = nodes;
  h.arcs = arcs;
  h.n_costs = n_costs;
  std::vector<unsigned char> idx(2 * idx_bytes, 0);
  if (arcs) {
    memcpy(idx.data(), tail, 4 * arcs);
    memcpy(idx.data() + idx_bytes, head, 4 * arcs);
  }
  uint64_t c = word_hash(idx.data(), idx.size(), nodes * 0x10001ull + arcs);
  c = word_hash(reinterpret_cast<const unsigned char*>(costs), 8 * n_costs, c);
  h.checksum = c;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(TPL_ERR_IO, "I/O error: %s", strerror(errno));
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && (idx.empty() || fwrite(idx.data(), 1, idx.size(), f) == idx.size()) &&
            (!n_costs || fwrite(costs, 8, n_costs, f) == n_costs);
  ok = (fclose(f) == 0) && ok;
  if (!ok) return fail(TPL_ERR_IO, "I/O error: short write to '%s'", path);
  return TPL_OK;
}

int tpl_kkt_save_binary(const tpl_kkt* kkt, const char* path) {
  tpl::clear_error();
  if (!kkt || !path) return fail(TPL_ERR_PANIC, "null argument");
  if (!kkt->regular)
    return fail(TPL_ERR_SPARSE_CONSTRUCTION,
                "Internal error: the instance is not a plain arc list; the container stores the incidence view only.");
  return tpl_write_kkt_binary(path, kkt->num_nodes, kkt->num_arcs, kkt->tail.data(), kkt->head.data(),
                              kkt->costs.data(), kkt->costs.size());
}

int tpl_load_kkt_binary(const char* path, tpl_kkt** out) {
  tpl::clear_error();
  if (!path || !out) return fail(TPL_ERR_PANIC, "null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(TPL_ERR_IO, "I/O error: %s", strerror(errno));
  BinHeader h;
  if (fread(&h, sizeof h, 1, f) != 1) {
    fclose(f);
    return fail(TPL_ERR_UNEXPECTED_EOF, "Format error: Unexpected end of file while reading data.");
  }
  if (memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) {
    fclose(f);
    return fail(TPL_ERR_PROBLEM_LINE_MISSING, "Format error: not a TPLKKT1 container (bad magic or version).");
  }
  if (h.arcs > 0x7fffffffu || h.nodes > 0x7fffffffu || h.n_costs > h.arcs) {
    fclose(f);
    return fail(TPL_ERR_ARC_COUNT_MISMATCH, "Dimension mismatch: container header holds %llu arcs, %llu nodes, %llu costs.",
                (unsigned long long)h.arcs, (unsigned long long)h.nodes, (unsigned long long)h.n_costs);
  }
  const size_t m = h.arcs, idx_bytes = pad8(4 * m);
  {  // the file must be exactly as long as the header says, checked before anything of that size is allocated
    const long at = ftell(f);
    long size = -1;
    if (at >= 0 && fseek(f, 0, SEEK_END) == 0) size = ftell(f);
    const unsigned long long want = sizeof(BinHeader) + 2ull * idx_bytes + 8ull * h.n_costs;
    if (at < 0 || size < 0 || fseek(f, at, SEEK_SET) != 0) {
      fclose(f);
      return fail(TPL_ERR_IO, "I/O error: cannot determine the size of '%s'", path);
    }
    if ((unsigned long long)size != want) {
      fclose(f);
      return (unsigned long long)size < want
                 ? fail(TPL_ERR_UNEXPECTED_EOF, "Format error: Unexpected end of file while reading data.")
                 : fail(TPL_ERR_ARC_COUNT_MISMATCH, "Dimension mismatch: container is longer than its header says.");
    }
  }
  tpl_kkt* k = new tpl_kkt;
  k->num_nodes = h.nodes;
  k->num_arcs = m;
  std::vector<unsigned char> idx(2 * idx_bytes);
  k->costs.resize(h.n_costs);
  bool ok = (idx.empty() || fread(idx.data(), 1, idx.size(), f) == idx.size()) &&
            (!h.n_costs || fread(k->costs.data(), 8, h.n_costs, f) == h.n_costs);
  const bool trailing = ok && fgetc(f) != EOF;
  fclose(f);
  if (!ok || trailing) {
    delete k;
    return ok ? fail(TPL_ERR_ARC_COUNT_MISMATCH, "Dimension mismatch: container is longer than its header says.")
              : fail(TPL_ERR_UNEXPECTED_EOF, "Format error: Unexpected end of file while reading data.");
  }
  uint64_t c = word_hash(idx.data(), idx.size(), h.nodes * 0x10001ull + h.arcs);
  c = word_hash(reinterpret_cast<const unsigned char*>(k->costs.data()), 8 * k->costs.size(), c);
  if (c != h.checksum) {
    delete k;
    return fail(TPL_ERR_IO, "I/O error: container checksum mismatch (corrupt file)");
  }
  k->tail.resize(m);
  k->head.resize(m);
  if (m) {
    memcpy(k->tail.data(), idx.data(), 4 * m);
    memcpy(k->head.data(), idx.data() + idx_bytes, 4 * m);
  }
  for (size_t j = 0; j < m; ++j)
    if (k->tail[j] >= h.nodes || k->head[j] >= h.nodes) {
      delete k;
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
    }
  k->d.assign(m, 0.0);
  std::copy(k->costs.begin(), k->costs.end(), k->d.begin());
  k->regular = true;
  k->lazy_csc = true;
  *out = k;
  return TPL_OK;
}

void tpl_kkt_free(tpl_kkt* kkt) { delete kkt; }
size_t tpl_kkt_num_nodes(const tpl_kkt* kkt) { return kkt->num_nodes; }
size_t tpl_kkt_num_arcs(const tpl_kkt* kkt) { return kkt->num_arcs; }
size_t tpl_kkt_num_costs(const tpl_kkt* kkt) { return kkt->costs.size(); }
size_t tpl_kkt_nnz(const tpl_kkt* kkt) {
  tpl::kkt_ensure_csc(kkt);
  return kkt->val.size();
}

int tpl_kkt_csc(const tpl_kkt* kkt, size_t* n, size_t* nnz, const uint64_t** colptr, const uint64_t** rowidx,
                const double** val) {
  tpl::kkt_ensure_csc(kkt);
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
