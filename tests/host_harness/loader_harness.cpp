// Host-only harness for the sanitizer test (tests/test_host_sanitizers_cpu.py): links tpl_loader.cpp and tpl_ftk.cpp with
// stubs for the error slot (which lives in tpl_engine.cu) and drives them over well-formed, malformed and corrupt inputs.
// Built with -fsanitize=address,undefined; any report makes the process exit non-zero.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tpl_internal.h"
#include "tplanczos.h"

namespace tpl {
static std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
void clear_error() { g_err.clear(); }
int fail_parameter_mismatch(const char*, size_t, size_t) { return 4; }
int fail_input(const char*) { return 3; }
int fail_solver(const char*) { return 6; }
}  // namespace tpl

static std::string write_file(const std::string& dir, const char* name, const std::string& data) {
  std::string p = dir + "/" + name;
  FILE* f = fopen(p.c_str(), "wb");
  fwrite(data.data(), 1, data.size(), f);
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string dir = argv[1];
  int checks = 0;
  // 1. a real text pair -> CSC on demand -> container -> back
  tpl_kkt* k = nullptr;
  if (tpl_load_kkt(argv[2], argv[3], &k)) return 3;
  size_t n = 0, nnz = 0;
  const uint64_t *cp, *ri;
  const double* va;
  tpl_kkt_csc(k, &n, &nnz, &cp, &ri, &va);
  if (cp[n] != nnz) return 4;
  const std::string bin = dir + "/h.tplkkt";
  if (tpl_kkt_save_binary(k, bin.c_str())) return 5;
  tpl_kkt* k2 = nullptr;
  if (tpl_load_kkt_binary(bin.c_str(), &k2)) return 6;
  if (tpl_kkt_nnz(k2) != nnz) return 7;
  tpl_kkt_free(k2);
  // 2. every truncation of the container and every single-byte corruption of its first 96 bytes is refused or loads cleanly
  std::string raw;
  {
    FILE* f = fopen(bin.c_str(), "rb");
    char buf[4096];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) raw.append(buf, got);
    fclose(f);
  }
  for (size_t cut = 0; cut < raw.size(); cut += (cut < 200 ? 1 : 997)) {
    const std::string p = write_file(dir, "t.tplkkt", raw.substr(0, cut));
    tpl_kkt* t = nullptr;
    if (tpl_load_kkt_binary(p.c_str(), &t) == 0) return 8;  // a shorter file can never be valid
    ++checks;
  }
  for (size_t at = 0; at < 96 && at < raw.size(); ++at)
    for (int bit : {0, 3, 7}) {
      std::string bad = raw;
      bad[at] = char(bad[at] ^ (1 << bit));
      const std::string p = write_file(dir, "c.tplkkt", bad);
      tpl_kkt* t = nullptr;
      if (tpl_load_kkt_binary(p.c_str(), &t) == 0) {
        tpl_kkt_nnz(t);  // (flipping a reserved header byte is harmless: the file still loads, consistently)
        tpl_kkt_free(t);
      }
      ++checks;
    }
  // 3. malformed text
  const char* dmx_cases[] = {"", "p min 3 2\na 1\n", "p min 3 2\na 1 2\na 2 9\n", "p min 2 1\na 1 2\na 2 1\n", "a 1 2\n",
                             "p min 18446744073709551616 1\n", "p min 3 2\na 1 2 \xff\n", "p min 0 0\n"};
  for (const char* d : dmx_cases) {
    const std::string dp = write_file(dir, "m.dmx", d), qp = write_file(dir, "m.qfc", "2\n0\n0\n1e400\n-1e-400\n");
    tpl_kkt* t = nullptr;
    if (tpl_load_kkt(dp.c_str(), qp.c_str(), &t) == 0) {
      tpl_kkt_nnz(t);
      tpl_kkt_free(t);
    }
    ++checks;
  }
  // 4. f(T_k) solvers and residual estimates on degenerate inputs
  std::vector<double> al = {0.0, 1.0, -1.0, 2.0}, be = {1.0, 0.0, 1e-300, 1e300}, y(4), res(4);
  size_t ylen = 0;
  tpl_ftk_inv(al.data(), 4, be.data(), 3, y.data(), &ylen, nullptr);
  tpl_ftk_exp(al.data(), 4, be.data(), 3, y.data(), &ylen, nullptr);
  tpl_ftk_square(al.data(), 4, be.data(), 3, y.data(), &ylen, nullptr);
  tpl_ftk_inv_residuals(al.data(), 4, be.data(), 4, 1.0, res.data());
  tpl_ftk_inv_residuals(al.data(), 4, be.data(), 1, 1.0, res.data());
  tpl_ftk_inv_residuals(nullptr, 0, nullptr, 0, 1.0, nullptr);
  tpl_kkt_free(k);
  printf("harness ok: %d checks\n", checks);
  return 0;
}
