// tpl_internal.h -- shared host-side declarations of libtplanczos (not part of the public ABI).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/tplanczos.h"

namespace tpl {

// thread-local last-error slot; returns `code` so call sites can `return fail(...)`.
int fail(int code, const char* fmt, ...);
void clear_error();

// Display strings of the reference's error kinds (src/error.rs:23-57).
int fail_parameter_mismatch(const char* param_name, size_t expected, size_t actual);
int fail_input(const char* msg);
int fail_solver(const char* msg);

constexpr double kBreakdownTol = 2.220446049250313e-16 * 1000.0;  // 1000 * f64::EPSILON, mod.rs:140-143

}  // namespace tpl

// KKTSystem (src/utils/data_loader.rs:51-58) as loaded on the host.
struct tpl_kkt {
  size_t num_nodes = 0, num_arcs = 0;
  std::vector<double> costs;  // quadratic costs the .qfc really provided (<= num_arcs entries)
  // E (nodes x arcs) after faer's triplet merge, CSC with ascending rows per column
  std::vector<uint64_t> e_colptr, e_rowidx;
  std::vector<double> e_val;
  // A = [[D,E^T],[E,0]] CSC
  std::vector<uint64_t> colptr, rowidx;
  std::vector<double> val;
  // incidence view
  std::vector<uint32_t> tail, head;
  std::vector<double> d;  // zero-padded to num_arcs
  bool regular = false;
  bool lazy_csc = false;  // container-loaded: the CSC views above are built on first use (kkt_ensure_csc)
};

namespace tpl {
void kkt_ensure_csc(const tpl_kkt* kkt);
}
