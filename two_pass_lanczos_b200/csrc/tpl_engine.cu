// tpl_engine.cu -- host driver + C ABI of the B200-native two-pass Lanczos engine (component H1).
//
// Mirrors the control flow of the reference's src/solvers.rs and src/algorithms/{lanczos,lanczos_two_pass}.rs
// around the persistent kernels of tpl_kernels.cuh.  No CPU fallback: every compute entry point needs a
// CUDA device and fails with TPL_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tpl_internal.h"
#include "tpl_blocks.cuh"
#include "tpl_blocks_host.h"
#include "tpl_build.cuh"
#include "tpl_cells_host.h"
#include "tpl_csr.cuh"
#include "tpl_dense.cuh"
#include "tpl_kernels.cuh"
#include "tpl_sharded.cuh"
#include "tpl_tiles.cuh"
#include "tpl_tiles_host.h"

// ============================================================================ errors
namespace tpl {
static thread_local std::string g_err;

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
int fail_parameter_mismatch(const char* param_name, size_t expected, size_t actual) {
  // src/error.rs:40 "Parameter mismatch: `{param_name}` expects size {expected}, but got {actual}."
  return fail(TPL_ERR_PARAMETER_MISMATCH, "Parameter mismatch: `%s` expects size %zu, but got %zu.", param_name,
              expected, actual);
}
int fail_input(const char* msg) { return fail(TPL_ERR_INPUT, "Invalid input parameter: %s", msg); }  // error.rs:37
int fail_solver(const char* msg) {
  return fail(TPL_ERR_SOLVER, "The user-provided f(T_k) solver failed: %s", msg);  // error.rs:50
}
}  // namespace tpl

using tpl::fail;

#define CUDA_TRY(expr)                                                                                \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(TPL_ERR_CUDA, "CUDA error: %s (%s) at %s:%d", cudaGetErrorString(e_), #expr, __FILE__, \
                  __LINE__);                                                                          \
  } while (0)

extern "C" const char* tpl_last_error_message(void) { return tpl::g_err.c_str(); }
extern "C" const char* tpl_version(void) { return "tplanczos 0.1 (sm_100a)"; }

// ============================================================================ NCCL (bound at run time)
// The communicator of the arc-partitioned multi-GPU mode is NCCL over NVLink.  libnccl.so.2 is resolved with dlopen
// the first time a sharded operator is made, so that the library itself loads (and every single-GPU entry point
// works) on hosts without NCCL, and so that a process that already carries an NCCL (e.g. the one bundled with
// torch.distributed) shares that copy instead of loading a second one.
namespace nccl {
struct UniqueId {
  char internal[128];
};
typedef struct ncclComm* Comm;
typedef int Result;
constexpr int kSum = 0, kFloat64 = 8;
typedef Result (*GetUniqueIdFn)(UniqueId*);
typedef Result (*CommInitRankFn)(Comm*, int, UniqueId, int);
typedef Result (*CommDestroyFn)(Comm);
typedef Result (*AllReduceFn)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(Result);
struct Api {
  GetUniqueIdFn GetUniqueId = nullptr;
  CommInitRankFn CommInitRank = nullptr;
  CommDestroyFn CommDestroy = nullptr;
  AllReduceFn AllReduce = nullptr;
  GetErrorStringFn GetErrorString = nullptr;
  bool ok = false;
  std::string why;
};
static Api& api() {
  static Api a = [] {
    Api r;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) {
      r.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown error");
      return r;
    }
    r.GetUniqueId = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
    r.CommInitRank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
    r.CommDestroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
    r.AllReduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
    r.GetErrorString = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
    r.ok = r.GetUniqueId && r.CommInitRank && r.CommDestroy && r.AllReduce && r.GetErrorString;
    if (!r.ok) r.why = "libnccl.so.2 lacks a required symbol";
    return r;
  }();
  return a;
}
}  // namespace nccl

#define NCCL_TRY(expr)                                                                                       \
  do {                                                                                                       \
    nccl::Result r_ = (expr);                                                                                \
    if (r_ != 0)                                                                                             \
      return fail(TPL_ERR_COMM, "NCCL error: %s (%s) at %s:%d", nccl::api().GetErrorString(r_), #expr, __FILE__, \
                  __LINE__);                                                                                 \
  } while (0)

// ============================================================================ handle
namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

constexpr size_t kHeaderDoubles = 8;  // State (48 B) padded to 64 B in the coefficient block

}  // namespace

struct tpl_op {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int format = 0;  // 1 = CSR, 2 = incidence, 3 = dense symmetric
  uint32_t n = 0;
  int G = 0;  // CTAs of the persistent grid (= SM count)
  tpl::IncidenceOp inc{};
  tpl::SellOp sell{};  // generic sparse operator: SELL-32 slices + long-row segments (tpl_csr.cuh)
  tpl::DenseOp dense{};
  std::vector<std::pair<void*, size_t>> allocs;
  size_t device_bytes = 0;
  uint64_t matrix_bytes = 0;
  size_t smem_bytes = 0;
  // workspace
  double* buf[3] = {nullptr, nullptr, nullptr};
  double* b_d = nullptr;
  double* x_d = nullptr;
  double* coef_d = nullptr;  // [header | alphas cap | betas cap | y cap]
  size_t coef_cap = 0;
  uint4* slots = nullptr;      // grid-sync lines [2][G][kSlotAtoms]
  bool resident_ok = false;    // the per-CTA slice of the incidence operator fits in shared memory
  tpl::ResidentOp res{};
  size_t smem_res1 = 0, smem_res2 = 0;
  bool cells_ok = false;       // 2-D cell partition fits in shared memory (tpl_cells.cuh)
  tpl::CellOp cell{};
  size_t smem_cell1 = 0, smem_cell2 = 0;
  void* cell_xchg = nullptr;   // inbox | gather | all-reduce atoms, cleared before every pass
  size_t cell_xchg_bytes = 0;
  bool tiled_ok = false;       // streaming kernels with tiled node sums are usable (shared-memory budget)
  tpl::TileOp tile{};
  void* tile_block = nullptr;  // exchange buffers of the tiled kernels (partials | node values | slots)
  void* blk_block = nullptr;   // exchange buffers of the blocked kernels
  void* fab_block = nullptr;   // sharded handle: the exchange block of the family that runs fused (one allocation = one IPC handle)
  std::vector<void*> fab_peers;  // peer blocks opened through CUDA IPC
  bool fab_connected = false;  // world > 1 and every peer's block is mapped: the persistent tiled kernels span all ranks
  unsigned fab_epoch = 0;      // barrier epoch the next fused pass starts from (never reset: peers write into our slots)
  size_t smem_tile1 = 0, smem_tile2 = 0;
  size_t blk_lent_words = 0;   // words of the blocked tile lists (blk.tl.lent)
  bool device_built = false;   // node lists and blocked layout were constructed on the device (tpl_build.cuh)
  bool blocked_ok = false;     // blocked streaming kernels (tpl_blocks.cuh): 2-D node-block partition, cell-order vectors
  tpl::BlockOp blk{};
  size_t smem_blk1 = 0, smem_blk2 = 0, smem_blk2v = 0;
  uint32_t blk_max_cell = 0;   // arcs of the largest cell
  double* bbuf[3] = {nullptr, nullptr, nullptr};  // the three rotating vectors in cell order: [Mpad arcs | p nodes]
  double* h_pin = nullptr;  // pinned mirror of coef_d
  double* V_int = nullptr;
  size_t V_int_elems = 0;
  double* sweep_d = nullptr;    // k-sweep: staging block for up to kSweepChunk solutions when the caller's X is on the host
  size_t sweep_elems = 0;
  double* sweep_y_d = nullptr;  // k-sweep: coefficient matrix Y (kmax x chunk)
  size_t sweep_y_elems = 0;
  cudaEvent_t ev[6] = {};
  bool timed[3] = {false, false, false};
  uint64_t launches = 0;
  int mode = 0;
  // arc-partitioned multi-GPU mode (world > 1): this handle holds rank `rank`'s arc block and a node replica
  int rank = 0, world = 1;
  nccl::Comm comm = nullptr;
  // REPLICATED execution of a sharded handle (world > 1, the whole operator fits the on-chip cell kernels): `inner` is an
  // unsharded handle over ALL arcs on this rank's GPU.  A solve all-reduces the ranks' arc slices of b into the full vector,
  // runs the single-GPU kernels redundantly on every rank (no per-step communication at all) and hands back this rank's
  // slice.  A Lanczos step of such a job is a few microseconds of on-chip work; sharding it costs two cross-GPU barriers per
  // step (measured: 19.7 ms on 2 GPUs against 4.5 ms on one).
  tpl_op* inner = nullptr;
  size_t rep_lo = 0, rep_hi = 0, rep_m = 0;  // this rank's arc range and the global arc count
  double* rep_in = nullptr;                  // [m + p] full input vector
  double* rep_out = nullptr;                 // [m + p] full output vector
  double* rep_V = nullptr;                   // full basis / solution block of the last call that needed one
  size_t rep_V_elems = 0;
  double* red_d = nullptr;   // [2][p + 1] node sums + alpha partial (double-buffered for pass 2)
  double* red2_d = nullptr;  // [1]

  tpl::State* st_d() const { return reinterpret_cast<tpl::State*>(coef_d); }
  double* alphas_d() const { return coef_d + kHeaderDoubles; }
  double* betas_d() const { return coef_d + kHeaderDoubles + coef_cap; }
  double* y_d() const { return coef_d + kHeaderDoubles + 2 * coef_cap; }
  unsigned long long* trace_d = nullptr;
  int trace_steps = 0;
  tpl::GridSync gs() const { return tpl::GridSync{slots, tpl::Trace{trace_d, trace_steps}, -1, 4}; }
};

namespace {

template <class T>
int dev_alloc(tpl_op* op, T** out, size_t count) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CUDA_TRY(cudaMalloc(&p, bytes));
  op->allocs.emplace_back(p, bytes);
  op->device_bytes += bytes;
  *out = static_cast<T*>(p);
  return TPL_OK;
}
template <class T>
int dev_upload(tpl_op* op, const T** out, const std::vector<T>& host) {
  T* p = nullptr;
  if (int rc = dev_alloc(op, &p, host.size())) return rc;
  if (!host.empty()) CUDA_TRY(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = p;
  return TPL_OK;
}
// a device buffer allocated elsewhere (tpl_build.cuh) becomes the handle's
template <class T>
void dev_adopt(tpl_op* op, T* p, size_t count) {
  if (!p) return;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  op->allocs.emplace_back((void*)p, bytes);
  op->device_bytes += bytes;
}
int dev_free(tpl_op* op, void* p) {
  if (!p) return TPL_OK;
  auto it = std::find_if(op->allocs.begin(), op->allocs.end(), [p](const std::pair<void*, size_t>& a) { return a.first == p; });
  if (it != op->allocs.end()) {
    op->device_bytes -= it->second;
    op->allocs.erase(it);
  }
  CUDA_TRY(cudaFree(p));
  return TPL_OK;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------- long-row (segment) format
constexpr size_t kSegSmemMax = 96 * 1024;  // most shared memory a CTA spends on segment sums
struct HostLongRows {
  std::vector<uint32_t> row, seg_ptr, ent_ptr, ent_idx, cta_ptr;
  std::vector<double> ent_val;
  uint32_t max_segs = 0, max_ents = 0, max_rows = 0;
};

// row_ent[q]..row_ent[q+1] is the entry range of long row q inside ent_idx/ent_val (already filled).
void build_segments(HostLongRows& h, const std::vector<uint64_t>& row_ent, int G) {
  const size_t nlong = h.row.size();
  uint64_t L = 256, longest = 0;
  for (size_t q = 0; q < nlong; ++q) longest = std::max<uint64_t>(longest, row_ent[q + 1] - row_ent[q]);
  for (;;) {  // longer segments until a CTA's share of segment sums is small -- as far as that helps: a row is at least one segment
    uint64_t segs = 0;
    for (size_t q = 0; q < nlong; ++q) segs += (row_ent[q + 1] - row_ent[q] + L - 1) / L;
    if (segs <= (uint64_t)G * 3072 || L >= longest) break;
    L *= 2;
  }
  h.seg_ptr.assign(nlong + 1, 0);
  h.ent_ptr.clear();
  for (size_t q = 0; q < nlong; ++q) {
    h.seg_ptr[q] = (uint32_t)h.ent_ptr.size();
    for (uint64_t e = row_ent[q]; e < row_ent[q + 1]; e += L) h.ent_ptr.push_back((uint32_t)e);
  }
  h.seg_ptr[nlong] = (uint32_t)h.ent_ptr.size();
  h.ent_ptr.push_back((uint32_t)row_ent[nlong]);
  // Fix segment ends: a segment ends where the next one starts, except the last of a row which ends at
  // the row end.  Because rows are stored back to back, ent_ptr[s+1] is correct in both cases.
  const uint32_t nseg = h.seg_ptr[nlong];
  const uint32_t target = std::max<uint32_t>(1, (nseg + G - 1) / G);
  h.cta_ptr.assign(G + 1, 0);
  for (int c = 0; c <= G; ++c) {
    const uint64_t want = (uint64_t)c * target;
    size_t q = std::lower_bound(h.seg_ptr.begin(), h.seg_ptr.begin() + nlong, want,
                                [](uint32_t a, uint64_t b) { return (uint64_t)a < b; }) -
               h.seg_ptr.begin();
    h.cta_ptr[c] = (uint32_t)q;
  }
  h.cta_ptr[0] = 0;
  h.cta_ptr[G] = (uint32_t)nlong;
  // rows without segments at the tail must still be owned: spread every row, by count, if there are no segments
  if (nseg == 0)
    for (int c = 0; c <= G; ++c) h.cta_ptr[c] = (uint32_t)std::min<uint64_t>(nlong, ((uint64_t)nlong * c + G - 1) / G);
  h.max_segs = h.max_ents = h.max_rows = 0;
  for (int c = 0; c < G; ++c) {
    const uint32_t q0 = h.cta_ptr[c], q1 = h.cta_ptr[c + 1];
    h.max_segs = std::max(h.max_segs, h.seg_ptr[q1] - h.seg_ptr[q0]);
    h.max_ents = std::max(h.max_ents, h.ent_ptr[h.seg_ptr[q1]] - h.ent_ptr[h.seg_ptr[q0]]);
    h.max_rows = std::max(h.max_rows, q1 - q0);
  }
}

int upload_long_rows(tpl_op* op, const HostLongRows& h, tpl::LongRows& d, const uint32_t* ent_idx_dev = nullptr) {
  d.nlong = (uint32_t)h.row.size();
  d.max_segs = h.max_segs;
  d.max_ents = h.max_ents;
  d.max_rows = h.max_rows;
  if (int rc = dev_upload(op, &d.row, h.row)) return rc;
  if (int rc = dev_upload(op, &d.seg_ptr, h.seg_ptr)) return rc;
  if (int rc = dev_upload(op, &d.ent_ptr, h.ent_ptr)) return rc;
  if (ent_idx_dev) {
    d.ent_idx = ent_idx_dev;  // built on the device (already adopted by the handle)
  } else if (int rc = dev_upload(op, &d.ent_idx, h.ent_idx)) {
    return rc;
  }
  if (int rc = dev_upload(op, &d.cta_ptr, h.cta_ptr)) return rc;
  d.ent_val = nullptr;
  if (!h.ent_val.empty())
    if (int rc = dev_upload(op, &d.ent_val, h.ent_val)) return rc;
  // Operators with very many rows in this format (the node rows of a sparse graph with millions of nodes): a CTA's segment
  // sums do not fit in shared memory; they go through a scratch array in HBM instead.
  d.seg_scratch = nullptr;
  if ((size_t)h.max_segs * sizeof(double) > kSegSmemMax) {
    if (int rc = dev_alloc(op, &d.seg_scratch, h.seg_ptr.empty() ? 1 : (size_t)h.seg_ptr.back())) return rc;
    d.max_segs = 1;
  }
  return TPL_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the KERNEL (per device), shared by every handle of the
// process: it is only ever raised, so that a handle created later with a smaller shared-memory footprint cannot lower
// the cap under an earlier handle whose cooperative launches need more (a launch may always ask for less than the cap).
template <class K>
int set_smem(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> high;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = high[{reinterpret_cast<const void*>(kernel), dev}];
  if (bytes <= cur) return TPL_OK;
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  cur = bytes;
  return TPL_OK;
}

int open_device(tpl_op* op, int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(TPL_ERR_CUDA, "CUDA error: no usable CUDA device (%s); libtplanczos has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= count) return fail(TPL_ERR_CUDA, "CUDA error: device %d out of range (%d devices)", device, count);
  op->device = device;
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (!prop.cooperativeLaunch) return fail(TPL_ERR_CUDA, "CUDA error: device lacks cooperative launch");
  op->G = std::min<int>(prop.multiProcessorCount, (int)tpl::kMaxGridCtas);  // one CTA per SM (grid_sync gathers at most kMaxGridCtas payloads)
  CUDA_TRY(cudaStreamCreateWithFlags(&op->stream, cudaStreamNonBlocking));
  for (auto& ev : op->ev) CUDA_TRY(cudaEventCreate(&ev));
  return TPL_OK;
}

int ensure_coef(tpl_op* op, size_t k) {
  if (k <= op->coef_cap && op->coef_d) return TPL_OK;
  size_t cap = std::max<size_t>(k, 64);
  if (op->coef_d) {
    if (int rc = dev_free(op, op->coef_d)) return rc;
    op->coef_d = nullptr;
  }
  if (op->h_pin) {
    cudaFreeHost(op->h_pin);
    op->h_pin = nullptr;
  }
  const size_t doubles = kHeaderDoubles + 3 * cap;
  if (int rc = dev_alloc(op, &op->coef_d, doubles)) return rc;
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&op->h_pin), doubles * sizeof(double)));
  op->coef_cap = cap;
  return TPL_OK;
}

// Exchange block of one kernel family (partial node sums | node values | barrier slots) for `world` ranks; every family
// has a local one (world = 1: the single-GPU kernels are the fused path with one rank).  The fused multi-GPU passes run on
// ONE family per handle -- the blocked streaming kernels when their layout exists, else the tiled ones -- whose block is
// then re-allocated for `world` ranks and exported (op->fab_block).
tpl::Fabric& active_fabric(tpl_op* op) { return op->blocked_ok ? op->blk.tl.fab : op->tile.fab; }
void fabric_layout(const tpl_op* op, bool blocked, const tpl::Fabric& f, size_t& part, size_t& nodes, size_t& slot_off) {
  part = blocked ? 2 * (size_t)f.Bp * f.world * (op->blk.GC + op->blk.GR) : 2 * (size_t)f.Gtot * f.Bp;
  nodes = 2 * (size_t)op->inc.p + 2;
  slot_off = ((part + nodes) * sizeof(double) + 127) / 128 * 128;  // barrier lines: 128-byte aligned
}
int alloc_fabric(tpl_op* op, bool blocked, int rank, int world, void** block_slot) {
  const size_t p = op->inc.p;
  const size_t G = blocked ? (size_t)op->blk.GR * op->blk.GC : (size_t)op->G;  // CTAs of the kernels that use it
  tpl::Fabric& f = blocked ? op->blk.tl.fab : op->tile.fab;
  if (*block_slot) {
    if (int rc = dev_free(op, *block_slot)) return rc;
    *block_slot = nullptr;
  }
  const uint32_t R = (uint32_t)std::max<size_t>(1, (p + (size_t)world * G - 1) / ((size_t)world * G));
  (blocked ? op->blk.tl.R : op->tile.R) = R;
  f = tpl::Fabric{};
  f.rank = rank;
  f.world = world;
  f.Gtot = (uint32_t)(world * G);
  f.Bp = (uint32_t)(G * R);
  size_t part, nodes, slot_off;
  fabric_layout(op, blocked, f, part, nodes, slot_off);
  const size_t bytes = slot_off + 2 * (size_t)f.Gtot * tpl::kSlotAtoms * sizeof(uint4);
  char* block = nullptr;
  if (int rc = dev_alloc(op, &block, bytes)) return rc;
  CUDA_TRY(cudaMemset(block, 0, bytes));
  *block_slot = block;
  f.partials[rank] = reinterpret_cast<double*>(block);
  f.nodebuf[rank] = f.partials[rank] + part;
  f.slots[rank] = reinterpret_cast<uint4*>(block + slot_off);
  return TPL_OK;
}
int setup_local_fabric(tpl_op* op, bool blocked) { return alloc_fabric(op, blocked, 0, 1, blocked ? &op->blk_block : &op->tile_block); }
// sharded handle: the active family's block for `world` ranks (the other family is switched off by the caller)
int setup_fabric(tpl_op* op, int rank, int world) {
  void** slot = op->blocked_ok ? &op->blk_block : &op->tile_block;
  if (int rc = alloc_fabric(op, op->blocked_ok, rank, world, slot)) return rc;
  op->fab_block = *slot;
  return TPL_OK;
}

int finish_setup(tpl_op* op) {
  const size_t n = op->n;
  for (auto& b : op->buf)
    if (int rc = dev_alloc(op, &b, n)) return rc;
  if (int rc = dev_alloc(op, &op->b_d, n)) return rc;
  if (int rc = dev_alloc(op, &op->x_d, n)) return rc;
  if (int rc = dev_alloc(op, &op->slots, 2 * (size_t)op->G * tpl::kSlotAtoms)) return rc;
  CUDA_TRY(cudaMemset(op->slots, 0, sizeof(uint4) * 2 * op->G * tpl::kSlotAtoms));
  if (int rc = ensure_coef(op, 1024)) return rc;
  // opt in to the dynamic shared memory the kernels need and check the grid is co-resident
  const size_t smem = op->smem_bytes;
  int per_sm = 0;
  if (op->format == 2) {
    if (int rc = set_smem(tpl::pass1_kernel<tpl::IncidenceOp, false>, smem)) return rc;
    if (int rc = set_smem(tpl::pass1_kernel<tpl::IncidenceOp, true>, smem)) return rc;
    if (int rc = set_smem(tpl::pass2_kernel<tpl::IncidenceOp, false>, smem)) return rc;
    if (int rc = set_smem(tpl::pass2_kernel<tpl::IncidenceOp, true>, smem)) return rc;
    if (int rc = set_smem(tpl::apply_kernel<tpl::IncidenceOp>, smem)) return rc;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tpl::pass1_kernel<tpl::IncidenceOp, true>,
                                                           tpl::kBlock, smem));
    // resident shape: the CTA's slice of the operator and of the vectors stays in shared memory
    if (op->resident_ok) {
      const uint32_t A = (op->inc.m + op->G - 1) / op->G;
      op->smem_res1 = tpl::resident_smem_bytes(op->inc.p, A, op->G, op->res.R, op->res.max_long, false);
      op->smem_res2 = tpl::resident_smem_bytes(op->inc.p, A, op->G, op->res.R, op->res.max_long, true);
      int max_optin = 0;
      CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, op->device));
      op->resident_ok = op->smem_res2 + 2048 <= (size_t)max_optin;  // + the kernels' static shared memory (CtaShared: 1.5 KB)
    }
    if (op->resident_ok) {
      if (int rc = set_smem(tpl::pass1_resident_kernel<false>, op->smem_res1)) return rc;
      if (int rc = set_smem(tpl::pass1_resident_kernel<true>, op->smem_res1)) return rc;
      if (int rc = set_smem(tpl::pass2_resident_kernel<false>, op->smem_res2)) return rc;
      if (int rc = set_smem(tpl::pass2_resident_kernel<true>, op->smem_res2)) return rc;
      int r1 = 0, r2 = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r1, tpl::pass1_resident_kernel<true>, tpl::kBlock,
                                                             op->smem_res1));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r2, tpl::pass2_resident_kernel<true>, tpl::kBlock,
                                                             op->smem_res2));
      if (r1 < 1 || r2 < 1) op->resident_ok = false;
    }
    if (op->cells_ok) {
      op->smem_cell1 = tpl::cell_smem_bytes(op->cell, false);
      op->smem_cell2 = tpl::cell_smem_bytes(op->cell, true);
      if (int rc = set_smem(tpl::pass1_cell_kernel<false>, op->smem_cell1)) return rc;
      if (int rc = set_smem(tpl::pass1_cell_kernel<true>, op->smem_cell1)) return rc;
      if (int rc = set_smem(tpl::pass2_cell_kernel<false>, op->smem_cell2)) return rc;
      if (int rc = set_smem(tpl::pass2_cell_kernel<true>, op->smem_cell2)) return rc;
      int c1 = 0, c2 = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c1, tpl::pass1_cell_kernel<true>, tpl::kBlock, op->smem_cell1));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c2, tpl::pass2_cell_kernel<true>, tpl::kBlock, op->smem_cell2));
      if (c1 < 1 || c2 < 1) op->cells_ok = false;
    }
    if (op->tiled_ok) {
      if (int rc = set_smem(tpl::pass1_tiled_kernel<false>, op->smem_tile1)) return rc;
      if (int rc = set_smem(tpl::pass1_tiled_kernel<true>, op->smem_tile1)) return rc;
      if (int rc = set_smem(tpl::pass2_tiled_kernel<false>, op->smem_tile2)) return rc;
      if (int rc = set_smem(tpl::pass2_tiled_kernel<true>, op->smem_tile2)) return rc;
      int t1 = 0, t2 = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t1, tpl::pass1_tiled_kernel<true>, tpl::kBlock, op->smem_tile1));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t2, tpl::pass2_tiled_kernel<true>, tpl::kBlock, op->smem_tile2));
      if (t1 < 1 || t2 < 1) op->tiled_ok = false;
    }
    if (op->blocked_ok) {
      if (int rc = set_smem(tpl::pass1_blocked_kernel<false>, op->smem_blk1)) return rc;
      if (int rc = set_smem(tpl::pass1_blocked_kernel<true>, op->smem_blk1)) return rc;
      if (int rc = set_smem(tpl::pass2_blocked_kernel<false>, op->smem_blk2)) return rc;
      if (int rc = set_smem(tpl::pass2_blocked_kernel<true>, op->smem_blk2v)) return rc;
      int b1 = 0, b2 = 0, b3 = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, tpl::pass1_blocked_kernel<true>, tpl::kBlock, op->smem_blk1));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b2, tpl::pass2_blocked_kernel<false>, tpl::kBlock, op->smem_blk2));
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b3, tpl::pass2_blocked_kernel<true>, tpl::kBlock, op->smem_blk2v));
      if (b1 < 1 || b2 < 1 || b3 < 1) op->blocked_ok = false;
    }
  } else if (op->format == 3) {
    if (int rc = set_smem(tpl::pass1_dense_kernel<false>, smem)) return rc;
    if (int rc = set_smem(tpl::pass1_dense_kernel<true>, smem)) return rc;
    if (int rc = set_smem(tpl::pass2_dense_kernel<false>, smem)) return rc;
    if (int rc = set_smem(tpl::pass2_dense_kernel<true>, smem)) return rc;
    if (int rc = set_smem(tpl::apply_dense_kernel, smem)) return rc;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tpl::pass1_dense_kernel<true>, tpl::kBlock, smem));
  } else {
    if (int rc = set_smem(tpl::pass1_csr_kernel<false>, smem)) return rc;
    if (int rc = set_smem(tpl::pass1_csr_kernel<true>, smem)) return rc;
    if (int rc = set_smem(tpl::pass2_csr_kernel<false>, smem)) return rc;
    if (int rc = set_smem(tpl::pass2_csr_kernel<true>, smem)) return rc;
    if (int rc = set_smem(tpl::apply_csr_kernel, smem)) return rc;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tpl::pass1_csr_kernel<true>, tpl::kBlock, smem));
  }
  if (per_sm < 1) return fail(TPL_ERR_CUDA, "CUDA error: persistent kernel does not fit on an SM (smem %zu B)", smem);
  return TPL_OK;
}

constexpr size_t kSmemBudget = 200 * 1024;

}  // namespace

// ============================================================================ operator construction
extern "C" {

void tpl_op_free(tpl_op* op) {
  if (!op) return;
  DeviceGuard g(op->device);
  if (op->stream) cudaStreamSynchronize(op->stream);
  if (op->inner) tpl_op_free(op->inner);
  if (op->comm) nccl::api().CommDestroy(op->comm);
  for (void* peer : op->fab_peers) cudaIpcCloseMemHandle(peer);
  for (auto& a : op->allocs) cudaFree(a.first);
  if (op->h_pin) cudaFreeHost(op->h_pin);
  for (auto& ev : op->ev)
    if (ev) cudaEventDestroy(ev);
  if (op->own_stream && op->stream) cudaStreamDestroy(op->stream);
  delete op;
}

int tpl_op_from_csc(size_t n, const uint64_t* colptr, const uint64_t* rowidx, const double* val, int device,
                    tpl_op** out) {
  tpl::clear_error();
  if (!colptr || !out || (n && colptr[n] && (!rowidx || !val))) return fail(TPL_ERR_PANIC, "null argument");
  const uint64_t nnz = colptr[n];
  if (n == 0 || n >= 0x7fffffffull || nnz >= 0xffffffffull)
    return fail(TPL_ERR_DIMENSION_MISMATCH, "Dimension mismatch: operator has %zu columns but vector has %zu rows.", n, n);
  for (size_t j = 0; j < n; ++j)
    if (colptr[j] > colptr[j + 1])
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  for (uint64_t q = 0; q < nnz; ++q)
    if (rowidx[q] >= n)
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  tpl_op* op = new tpl_op;
  int rc = open_device(op, device);
  if (rc) {
    delete op;
    return rc;
  }
  op->format = 1;
  op->n = (uint32_t)n;
  // CSC -> CSR (column indices ascending within every row, i.e. the reference's accumulation order)
  std::vector<uint32_t> row_ptr(n + 1, 0), col(nnz);
  std::vector<double> v(nnz);
  for (uint64_t q = 0; q < nnz; ++q) ++row_ptr[rowidx[q] + 1];
  for (size_t i = 0; i < n; ++i) row_ptr[i + 1] += row_ptr[i];
  {
    std::vector<uint32_t> fill(row_ptr.begin(), row_ptr.end() - 1);
    for (size_t j = 0; j < n; ++j)
      for (uint64_t q = colptr[j]; q < colptr[j + 1]; ++q) {
        const uint32_t dst = fill[rowidx[q]]++;
        col[dst] = (uint32_t)j;
        v[dst] = val[q];
      }
  }
  const uint32_t long_thresh = 96;
  HostLongRows h;
  std::vector<uint64_t> row_ent{0};
  for (size_t i = 0; i < n; ++i) {
    const uint32_t len = row_ptr[i + 1] - row_ptr[i];
    if (len > long_thresh) {
      h.row.push_back((uint32_t)i);
      h.ent_idx.insert(h.ent_idx.end(), col.begin() + row_ptr[i], col.begin() + row_ptr[i + 1]);
      h.ent_val.insert(h.ent_val.end(), v.begin() + row_ptr[i], v.begin() + row_ptr[i + 1]);
      row_ent.push_back(h.ent_idx.size());
    }
  }
  build_segments(h, row_ent, op->G);
  // SELL-32 slices of the short rows (tpl_csr.cuh): entry e of row r at sptr[r / 32] + 32 e + r % 32
  {
    const size_t nslice = (n + 31) / 32;
    std::vector<uint32_t> sptr(nslice + 1, 0);
    std::vector<uint8_t> rlen(n);
    static_assert(96 < 0xff, "row lengths of short rows fit a byte");
    for (size_t sl = 0; sl < nslice; ++sl) {
      uint32_t width = 0;
      for (size_t i = sl * 32; i < std::min(n, sl * 32 + 32); ++i) {
        const uint32_t len = row_ptr[i + 1] - row_ptr[i];
        rlen[i] = len > long_thresh ? 0xff : (uint8_t)len;
        if (len <= long_thresh) width = std::max(width, len);
      }
      sptr[sl + 1] = sptr[sl] + 32 * width;
    }
    std::vector<uint32_t> scol(sptr[nslice]);
    std::vector<double> sval(sptr[nslice], 0.0);
    for (size_t sl = 0; sl < nslice; ++sl) {
      const uint32_t width = (sptr[sl + 1] - sptr[sl]) / 32;
      for (uint32_t ln = 0; ln < 32; ++ln) {
        const size_t i = sl * 32 + ln;
        const uint32_t len = i < n && rlen[i] != 0xff ? rlen[i] : 0;
        for (uint32_t e = 0; e < width; ++e) {
          const size_t w = (size_t)sptr[sl] + 32 * e + ln;
          scol[w] = e < len ? col[row_ptr[i] + e] : (uint32_t)std::min(i, n - 1);
          if (e < len) sval[w] = v[row_ptr[i] + e];
        }
      }
    }
    op->sell.n = (uint32_t)n;
    op->sell.nslice = (uint32_t)nslice;
    rc = dev_upload(op, &op->sell.sptr, sptr);
    if (!rc) rc = dev_upload(op, &op->sell.scol, scol);
    if (!rc) rc = dev_upload(op, &op->sell.sval, sval);
    if (!rc) rc = dev_upload(op, &op->sell.rlen, rlen);
  }
  if (!rc) rc = upload_long_rows(op, h, op->sell.lr);
  op->matrix_bytes = 12ull * nnz + 4ull * (n + 1);
  op->smem_bytes = sizeof(double) * std::max<size_t>(op->sell.lr.max_segs, 1);
  if (!rc) rc = finish_setup(op);
  if (rc) {
    std::string keep = tpl::g_err;
    tpl_op_free(op);
    tpl::g_err = keep;
    return rc;
  }
  *out = op;
  return TPL_OK;
}

namespace {
int dense_operator(size_t n, const double* a, size_t lda, bool cplx, int device, tpl_op** out);
}
int tpl_op_from_dense(size_t n, const double* a, size_t lda, int device, tpl_op** out) {
  return dense_operator(n, a, lda, false, device, out);
}
int tpl_op_from_dense_hermitian(size_t n, const double* a, size_t lda, int device, tpl_op** out) {
  return dense_operator(n, a, lda, true, device, out);
}
int tpl_op_is_complex(const tpl_op* op) { return op && op->format == 3 && op->dense.cplx ? 1 : 0; }
int tpl_op_from_diagonal(size_t n, const double* diag, int device, tpl_op** out) {
  tpl::clear_error();
  if (!out || (n && !diag)) return fail(TPL_ERR_PANIC, "null argument");
  std::vector<uint64_t> colptr(n + 1), rowidx(n);
  for (size_t i = 0; i < n; ++i) colptr[i] = rowidx[i] = i;
  colptr[n] = n;
  return tpl_op_from_csc(n, colptr.data(), rowidx.data(), diag, device, out);
}
namespace {
// n = rows of the matrix (complex rows when cplx); the handle's vectors have n (2 n when cplx) doubles
int dense_operator(size_t n, const double* a, size_t lda, bool cplx, int device, tpl_op** out) {
  tpl::clear_error();
  if (!out || (n && !a)) return fail(TPL_ERR_PANIC, "null argument");
  const size_t w = cplx ? 2 : 1;  // doubles per entry
  if (n == 0 || n * w >= 0x7fffffffull || lda < n)
    return fail(TPL_ERR_DIMENSION_MISMATCH, "Dimension mismatch: operator has %zu columns but vector has %zu rows.", n, lda);
  tpl_op* op = new tpl_op;
  int rc = open_device(op, device);
  if (rc) {
    delete op;
    return rc;
  }
  op->format = 3;
  op->n = (uint32_t)(n * w);
  double* ad = nullptr;
  rc = dev_alloc(op, &ad, n * n * w);
  if (!rc && cudaMemcpy2D(ad, n * w * sizeof(double), a, lda * w * sizeof(double), n * w * sizeof(double), n, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = fail(TPL_ERR_CUDA, "CUDA error: copying the dense operator to the device failed");
  op->dense.n = (uint32_t)(n * w);
  op->dense.cplx = cplx ? 1u : 0u;
  op->dense.lda = n;
  op->dense.a = ad;
  op->dense.stage = n * w * sizeof(double) <= kSmemBudget ? 1u : 0u;
  op->smem_bytes = op->dense.stage ? n * w * sizeof(double) : 0;
  op->matrix_bytes = 8ull * w * n * n;
  if (!rc) rc = finish_setup(op);
  if (rc) {
    std::string keep = tpl::g_err;
    tpl_op_free(op);
    tpl::g_err = keep;
    return rc;
  }
  *out = op;
  return TPL_OK;
}
}  // namespace

namespace {
constexpr size_t kDeviceBuildArcs = 1u << 20;  // from here on the tables are built on the device
// TPL_BUILD_TIMING=1: wall-clock laps of operator construction on stderr
struct BuildLaps {
  bool on = std::getenv("TPL_BUILD_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "from_kkt %-36s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};
}  // namespace

int tpl_op_from_kkt(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len,
                    int device, tpl_op** out) {
  tpl::clear_error();
  BuildLaps laps;
  if (!out || (m && (!tail || !head)) || (d_len && !d)) return fail(TPL_ERR_PANIC, "null argument");
  if (d_len > m) return fail(TPL_ERR_PARAMETER_MISMATCH, "Parameter mismatch: `d` expects size %zu, but got %zu.", m, d_len);
  const size_t n = m + p;
  if (n == 0 || n >= 0x7fffffffull)
    return fail(TPL_ERR_DIMENSION_MISMATCH, "Dimension mismatch: operator has %zu columns but vector has %zu rows.", n, n);
  for (size_t j = 0; j < m; ++j)
    if (tail[j] >= p || head[j] >= p)
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  laps.lap("validate");
  tpl_op* op = new tpl_op;
  int rc = open_device(op, device);
  if (rc) {
    delete op;
    return rc;
  }
  laps.lap("open device");
  op->format = 2;
  op->n = (uint32_t)n;
  // Large instances build their tables on the device (tpl_build.cuh); the host builders below construct the small ones and
  // check the device's work (tpl_op_layout_check).  TPL_HOST_BUILD=1 / TPL_DEVICE_BUILD=1 force either way.
  bool device_build = m >= kDeviceBuildArcs;
  if (std::getenv("TPL_HOST_BUILD")) device_build = false;
  if (std::getenv("TPL_DEVICE_BUILD")) device_build = true;
  op->device_built = device_build;
  op->inc.m = (uint32_t)m;
  op->inc.p = (uint32_t)p;
  {  // d (zero-padded to m), tail, head straight from the caller's arrays
    double* dd = nullptr;
    uint32_t *t = nullptr, *hd = nullptr;
    rc = dev_alloc(op, &dd, m);
    if (!rc) rc = dev_alloc(op, &t, m);
    if (!rc) rc = dev_alloc(op, &hd, m);
    if (!rc && m) {
      CUDA_TRY(cudaMemsetAsync(dd, 0, m * sizeof(double), op->stream));
      if (d_len) CUDA_TRY(cudaMemcpyAsync(dd, d, d_len * sizeof(double), cudaMemcpyHostToDevice, op->stream));
      CUDA_TRY(cudaMemcpyAsync(t, tail, m * sizeof(uint32_t), cudaMemcpyHostToDevice, op->stream));
      CUDA_TRY(cudaMemcpyAsync(hd, head, m * sizeof(uint32_t), cudaMemcpyHostToDevice, op->stream));
      CUDA_TRY(cudaStreamSynchronize(op->stream));
    }
    op->inc.d = dd;
    op->inc.tail = t;
    op->inc.head = hd;
  }
  laps.lap("upload d, tail, head");
  // node -> arc lists, ascending arc index inside each node (tail: +x_j, head: -x_j; self-loops cancel)
  HostLongRows h;
  std::vector<uint64_t> row_ent(p + 1, 0);
  uint32_t* ent_idx_dev = nullptr;
  // the vector workspace of the blocked kernels (three Lanczos vectors and x in cell order, at most 128 padding slots per
  // cell) doubles as the builders' scratch
  void* work[4] = {nullptr, nullptr, nullptr, nullptr};
  const size_t work_doubles = m + (size_t)op->G * tpl::kBStage + p;
  if (device_build)
    for (void*& w : work) {
      double* q = nullptr;
      if (!rc) rc = dev_alloc(op, &q, work_doubles);
      w = q;
    }
  if (device_build && !rc) {
    if (const int e = tpl::build_node_lists_device(m, p, op->inc.tail, op->inc.head, op->stream, work, &ent_idx_dev, row_ent)) {
      tpl_op_free(op);
      return fail(TPL_ERR_CUDA, "CUDA error while building the node lists on the device: %s", cudaGetErrorString((cudaError_t)e));
    }
    dev_adopt(op, ent_idx_dev, (size_t)row_ent[p]);
  } else {
    for (size_t j = 0; j < m; ++j)
      if (tail[j] != head[j]) {
        ++row_ent[tail[j] + 1];
        ++row_ent[head[j] + 1];
      }
    for (size_t u = 0; u < p; ++u) row_ent[u + 1] += row_ent[u];
    h.ent_idx.resize(row_ent[p]);
    std::vector<uint64_t> fill(row_ent.begin(), row_ent.end() - 1);
    for (size_t j = 0; j < m; ++j)
      if (tail[j] != head[j]) {
        h.ent_idx[fill[tail[j]]++] = (uint32_t)j;
        h.ent_idx[fill[head[j]]++] = (uint32_t)j | tpl::kSignBit;
      }
  }
  h.row.resize(p);
  for (size_t u = 0; u < p; ++u) h.row[u] = (uint32_t)(m + u);
  build_segments(h, row_ent, op->G);
  if (!rc) rc = upload_long_rows(op, h, op->inc.lr, ent_idx_dev);
  laps.lap("node -> arc lists");
  // resident shape: per-CTA node -> local-arc lists (tail: +, head: - with bit 15), ascending local arc index
  {
    const int G = op->G;
    const size_t A = (m + G - 1) / G;
    const uint32_t R = (uint32_t)((p + G - 1) / G);
    op->res.R = std::max<uint32_t>(R, 1);
    if (A <= 0x7fff && !rc) {
      std::vector<uint32_t> nl_ptr((size_t)G * (p + 1), 0), nl_long_ptr(G + 1, 0), nl_long;
      std::vector<uint16_t> nl_ent;
      nl_ent.reserve(2 * m);
      std::vector<uint32_t> cnt(p + 1);
      uint32_t max_long = 0;
      for (int c = 0; c < G; ++c) {
        const size_t lo = std::min(m, A * (size_t)c), hi = std::min(m, lo + A);
        std::fill(cnt.begin(), cnt.end(), 0u);
        for (size_t j = lo; j < hi; ++j)
          if (tail[j] != head[j]) {
            ++cnt[tail[j] + 1];
            ++cnt[head[j] + 1];
          }
        uint32_t* lp = nl_ptr.data() + (size_t)c * (p + 1);
        const uint32_t base = (uint32_t)nl_ent.size();
        lp[0] = base;
        for (size_t u = 0; u < p; ++u) lp[u + 1] = lp[u] + cnt[u + 1];
        nl_ent.resize(lp[p]);
        std::vector<uint32_t> fill(lp, lp + p);
        for (size_t j = lo; j < hi; ++j)
          if (tail[j] != head[j]) {
            nl_ent[fill[tail[j]]++] = (uint16_t)(j - lo);
            nl_ent[fill[head[j]]++] = (uint16_t)((j - lo) | 0x8000u);
          }
        nl_long_ptr[c] = (uint32_t)nl_long.size();
        for (size_t u = 0; u < p; ++u)
          if (lp[u + 1] - lp[u] > tpl::kLongList) nl_long.push_back((uint32_t)u);
        max_long = std::max<uint32_t>(max_long, (uint32_t)nl_long.size() - nl_long_ptr[c]);
      }
      nl_long_ptr[G] = (uint32_t)nl_long.size();
      op->res.max_long = std::max<uint32_t>(max_long, 1);
      rc = dev_upload(op, &op->res.nl_ptr, nl_ptr);
      if (!rc) rc = dev_upload(op, &op->res.nl_ent, nl_ent);
      if (!rc) rc = dev_upload(op, &op->res.nl_long_ptr, nl_long_ptr);
      if (!rc) rc = dev_upload(op, &op->res.nl_long, nl_long);
      if (!rc) rc = dev_alloc(op, &op->res.partials, 2 * (size_t)G * G * op->res.R);
      if (!rc) rc = dev_alloc(op, &op->res.nodebuf, 2 * p);
      op->resident_ok = !rc;
    }
  }
  laps.lap("resident lists");
  // cell shape: 2-D partition of the arcs, per-cell jagged-diagonal tail lists and head lists, tagged-atom exchange buffers
  if (!rc) {
    int max_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, op->device));
    tpl::HostCells hc;
    tpl::build_cells(m, p, tail, head, op->G, (size_t)max_optin, hc);
    if (hc.ok) {
      tpl::CellOp& co = op->cell;
      co = tpl::cells_probe(hc);
      rc = dev_upload(op, &co.hdr, hc.hdr);
      if (!rc) rc = dev_upload(op, &co.gidx, hc.gidx);
      if (!rc) rc = dev_upload(op, &co.lth, hc.lth);
      if (!rc) rc = dev_upload(op, &co.lines, hc.lines);
      if (!rc) rc = dev_upload(op, &co.tmap, hc.tmap);
      if (!rc) rc = dev_upload(op, &co.push, hc.push);
      if (!rc) rc = dev_upload(op, &co.walk, hc.walk);
      if (!rc) rc = dev_upload(op, &co.ent4, hc.ent4);
      if (!rc) rc = dev_upload(op, &co.slot_base, hc.slot_base);
      const size_t atoms = 2 * ((size_t)co.inbox_atoms + (size_t)co.L * tpl::kLine + (size_t)co.Gc * tpl::kLine * tpl::kArCopies);
      uint4* xchg = nullptr;
      if (!rc) rc = dev_alloc(op, &xchg, atoms);
      if (!rc) {
        op->cell_xchg = xchg;
        op->cell_xchg_bytes = atoms * sizeof(uint4);
        co.inbox = xchg;
        co.gather = xchg + 2 * (size_t)co.inbox_atoms;
        co.ar = co.gather + 2 * (size_t)co.L * tpl::kLine;
        op->cells_ok = true;
      }
    }
  }
  laps.lap("cell tables");
  // blocked streaming shape (tpl_blocks.cuh): node blocks, cell-order operator, tile lists over local node ids
  if (!rc && m >= 1 && p >= 1) {
    int max_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, op->device));
    tpl::HostBlocks hb_host;
    tpl::DeviceBlocks db;
    if (device_build) {
      if (const int e = tpl::build_blocks_device(m, p, op->inc.tail, op->inc.head, op->inc.d, op->G, (size_t)max_optin, op->stream, work, db)) {
        tpl_op_free(op);
        return fail(TPL_ERR_CUDA, "CUDA error while building the blocked layout on the device: %s", cudaGetErrorString((cudaError_t)e));
      }
    } else {
      tpl::build_blocks(m, p, tail, head, d, d_len, op->G, (size_t)max_optin, hb_host);
    }
    const tpl::HostBlocks& hb = device_build ? db.meta : hb_host;
    laps.lap(device_build ? "blocked layout (device)" : "blocked layout (host)");
    if (hb.ok) {
      tpl::BlockOp& bo = op->blk;
      bo = tpl::BlockOp{};
      bo.GR = hb.GR; bo.GC = hb.GC; bo.PT = hb.PT; bo.PH = hb.PH; bo.Mpad = hb.Mpad; bo.m = (uint32_t)m;
      bo.ring1 = hb.ring1; bo.ring2 = hb.ring2; bo.ring2v = hb.ring2v; bo.lblk = hb.lblk; bo.nl = hb.nl; bo.ntb = hb.ntb;
      if (const char* e = std::getenv("TPL_BLOCK_DBG")) bo.dbg = (uint32_t)std::atoi(e);  // timing experiments (tpl_blocks.cuh)
      bo.tl.T = hb.T;
      bo.tl.ntile = hb.ntile;
      rc = dev_upload(op, &bo.cell_off, hb.cell_off);
      if (!rc) rc = dev_upload(op, &bo.tbs, hb.tbs);
      if (!rc) rc = dev_upload(op, &bo.hbs, hb.hbs);
      if (!rc) rc = dev_upload(op, &bo.tbn, hb.tbn);
      if (!rc) rc = dev_upload(op, &bo.hbn, hb.hbn);
      if (device_build) {
        dev_adopt(op, db.d, hb.Mpad);
        dev_adopt(op, db.th, hb.Mpad);
        dev_adopt(op, db.gidx, hb.Mpad);
        dev_adopt(op, db.pdesc, hb.Mpad / tpl::kBStage);
        dev_adopt(op, db.thdr, (size_t)hb.GR * hb.GC * hb.ntile);
        dev_adopt(op, db.lent, db.lent_words);
        bo.d = db.d;
        bo.th = db.th;
        bo.gidx = db.gidx;
        bo.pdesc = db.pdesc;
        bo.tl.thdr = db.thdr;
        bo.tl.lent = db.lent;
        bo.tl.piece = nullptr;
        op->blk_lent_words = db.lent_words;
      } else {
        if (!rc) rc = dev_upload(op, &bo.d, hb.d);
        if (!rc) rc = dev_upload(op, &bo.th, hb.th);
        if (!rc) rc = dev_upload(op, &bo.gidx, hb.gidx);
        if (!rc) rc = dev_upload(op, &bo.pdesc, hb.pdesc);
        if (!rc) rc = dev_upload(op, &bo.tl.thdr, hb.thdr);
        if (!rc) rc = dev_upload(op, &bo.tl.lent, hb.lent);
        if (!rc) rc = dev_upload(op, &bo.tl.piece, hb.piece);
        op->blk_lent_words = hb.lent.size();
      }
      if (device_build) {  // the builders' scratch becomes the vector workspace
        bo.xc = static_cast<double*>(work[3]);
        for (int q = 0; q < 3; ++q) op->bbuf[q] = static_cast<double*>(work[q]);
        for (void*& w : work) w = nullptr;
      } else {
        if (!rc) rc = dev_alloc(op, &bo.xc, (size_t)hb.Mpad);
        for (auto& bb : op->bbuf)
          if (!rc) rc = dev_alloc(op, &bb, (size_t)hb.Mpad + p);
      }
      const uint32_t PL = hb.PT + hb.PH;
      op->smem_blk1 = tpl::block_smem_bytes(PL, hb.T, (int)hb.ring1, hb.lblk, hb.nl, hb.ntb, false, false, hb.ntile);
      op->smem_blk2 = tpl::block_smem_bytes(PL, hb.T, (int)hb.ring2, hb.lblk, hb.nl, hb.ntb, true, false, hb.ntile);
      op->smem_blk2v = tpl::block_smem_bytes(PL, hb.T, (int)hb.ring2v, hb.lblk, hb.nl, hb.ntb, true, true, hb.ntile);
      op->blocked_ok = !rc;
      for (size_t q = 0; q + 1 < hb.cell_off.size(); ++q) op->blk_max_cell = std::max(op->blk_max_cell, hb.cell_off[q + 1] - hb.cell_off[q]);
      if (!rc) rc = setup_local_fabric(op, true);
    }
  }
  for (void* w : work)
    if (w) dev_free(op, w);  // no blocked layout: the scratch is not needed
  laps.lap("blocked layout (upload)");
  // tiled streaming shape: per-CTA, per-tile entry lists; T = the largest tile (multiple of the stream batch, at most
  // 16384) for which pass 2's layout (node segment + accumulators + tile) fits in shared memory
  if (!rc && p >= 1 && p < (1u << 17) && !(op->blocked_ok && op->blk_max_cell >= 4 * 10240)) {
    int max_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, op->device));
    const long budget = (long)max_optin - 3072 - (long)(2 * p + 2 * tpl::kMaxPieces) * 8;  // static shared memory of the kernels + margin
    const uint32_t step = tpl::kUnrollB * tpl::kStreamThreads;  // a tile is a whole number of the largest stream batch
    uint32_t T = budget > 0 ? (uint32_t)std::min<long>(8192 / step * step, budget / 16 / step * step) : 0;  // two tile buffers
    const size_t A = (m + op->G - 1) / op->G;
    if (T >= step) {
      T = (uint32_t)std::min<size_t>(T, std::max<size_t>(step, (A + step - 1) / step * step));
      tpl::HostTiles ht;
      tpl::build_tiles(m, p, tail, head, op->G, T, ht);
      if (ht.T) {
        op->tile.T = ht.T;
        op->tile.ntile = ht.ntile;
        rc = dev_upload(op, &op->tile.thdr, ht.thdr);
        if (!rc) rc = dev_upload(op, &op->tile.lent, ht.lent);
        if (!rc) rc = dev_upload(op, &op->tile.piece, ht.piece);
        if (!rc) rc = setup_local_fabric(op, false);
        op->smem_tile1 = tpl::tile_smem_bytes((uint32_t)p, T, false);
        op->smem_tile2 = tpl::tile_smem_bytes((uint32_t)p, T, true);
        op->tiled_ok = !rc;
      }
    }
  }
  laps.lap("tiled lists");
  const size_t seg_bytes = sizeof(double) * std::max<size_t>(op->inc.lr.max_segs, 1);
  op->inc.stage_nodes = (p * sizeof(double) + seg_bytes <= kSmemBudget) ? 1 : 0;
  op->smem_bytes = seg_bytes + (op->inc.stage_nodes ? p * sizeof(double) : 0);
  op->matrix_bytes = 24ull * m + 4ull * p;  // (d, tail, head) + node->arc lists + list pointers
  if (!rc) rc = finish_setup(op);
  laps.lap("finish setup");
  if (rc) {
    std::string keep = tpl::g_err;
    tpl_op_free(op);
    tpl::g_err = keep;
    return rc;
  }
  *out = op;
  return TPL_OK;
}

int tpl_op_from_kkt_system(const tpl_kkt* kkt, int format, int device, tpl_op** out) {
  tpl::clear_error();
  if (!kkt || !out) return fail(TPL_ERR_PANIC, "null argument");
  if (format == 2 && !kkt->regular)
    return fail(TPL_ERR_SPARSE_CONSTRUCTION,
                "Internal error: the instance is not a plain arc list; use the generic CSR operator.");
  if (format == 2 || (format == 0 && kkt->regular))
    return tpl_op_from_kkt(kkt->num_arcs, kkt->num_nodes, kkt->tail.data(), kkt->head.data(), kkt->d.data(),
                           kkt->costs.size(), device, out);
  tpl::kkt_ensure_csc(kkt);
  return tpl_op_from_csc(kkt->num_arcs + kkt->num_nodes, kkt->colptr.data(), kkt->rowidx.data(), kkt->val.data(),
                         device, out);
}

size_t tpl_op_nrows(const tpl_op* op) { return op ? op->n : 0; }
int tpl_op_check_len(const tpl_op* op, size_t len) {
  if (!op) return fail(TPL_ERR_PANIC, "null argument");
  if (len != op->n)  // src/error.rs:29-35
    return fail(TPL_ERR_DIMENSION_MISMATCH, "Dimension mismatch: operator has %zu columns but vector has %zu rows.",
                (size_t)op->n, len);
  return TPL_OK;
}
int tpl_op_format(const tpl_op* op) { return op ? op->format : 0; }
int tpl_op_device(const tpl_op* op) { return op ? op->device : -1; }
uint64_t tpl_op_kernel_launches(const tpl_op* op) { return op ? op->launches + (op->inner ? op->inner->launches : 0) : 0; }
uint64_t tpl_op_matrix_bytes(const tpl_op* op) { return op ? op->matrix_bytes : 0; }
uint64_t tpl_op_device_bytes(const tpl_op* op) { return op ? op->device_bytes : 0; }

int tpl_op_set_stream(tpl_op* op, void* cuda_stream) {
  if (!op) return fail(TPL_ERR_PANIC, "null argument");
  DeviceGuard g(op->device);
  // (the inner handle of a replicated operator shares this handle's stream: it moves first, while the old stream still exists)
  if (op->inner)
    if (int rc = tpl_op_set_stream(op->inner, cuda_stream)) return rc;
  CUDA_TRY(cudaStreamSynchronize(op->stream));
  if (op->own_stream) CUDA_TRY(cudaStreamDestroy(op->stream));
  op->stream = static_cast<cudaStream_t>(cuda_stream);
  op->own_stream = false;
  return TPL_OK;
}

int tpl_op_trace_enable(tpl_op* op, size_t max_steps) {
  if (!op) return fail(TPL_ERR_PANIC, "null argument");
  DeviceGuard g(op->device);
  if (op->trace_d) {
    if (int rc = dev_free(op, op->trace_d)) return rc;
    op->trace_d = nullptr;
    op->trace_steps = 0;
  }
  if (max_steps == 0) return TPL_OK;
  const size_t words = (size_t)op->G * max_steps * tpl::kTraceMarks;
  if (int rc = dev_alloc(op, &op->trace_d, words)) return rc;
  CUDA_TRY(cudaMemset(op->trace_d, 0, words * sizeof(unsigned long long)));
  op->trace_steps = (int)max_steps;
  return TPL_OK;
}

int tpl_op_trace_read(tpl_op* op, uint64_t* out, size_t capacity, size_t* ctas, size_t* steps, size_t* marks) {
  if (!op || !ctas || !steps || !marks) return fail(TPL_ERR_PANIC, "null argument");
  DeviceGuard g(op->device);
  *ctas = (size_t)op->G;
  *steps = (size_t)op->trace_steps;
  *marks = (size_t)tpl::kTraceMarks;
  const size_t words = *ctas * *steps * *marks;
  if (!out) return TPL_OK;  // size query
  if (capacity < words) return tpl::fail_parameter_mismatch("out", words, capacity);
  if (!words) return TPL_OK;
  CUDA_TRY(cudaStreamSynchronize(op->stream));
  CUDA_TRY(cudaMemcpy(out, op->trace_d, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemset(op->trace_d, 0, words * sizeof(unsigned long long)));
  return TPL_OK;
}

int tpl_tiles_plan(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int ctas, uint32_t tile_arcs, int threads,
                   uint64_t stats[8]) {
  tpl::clear_error();
  if (!tail || !head || !stats) return fail(TPL_ERR_PANIC, "null argument");
  if (ctas < 1 || tile_arcs < 1 || tile_arcs + tpl::kMaxPieces > 16384 || p >= (1u << 17))
    return fail(TPL_ERR_PANIC, "tile plan: ctas >= 1, 1 <= tile_arcs <= %u and nodes < 2^17 are required", 16384 - tpl::kMaxPieces);
  for (size_t j = 0; j < m; ++j)
    if (tail[j] >= p || head[j] >= p)
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  tpl::HostTiles ht;
  tpl::build_tiles(m, p, tail, head, ctas, tile_arcs, ht, threads);
  std::fill(stats, stats + 8, 0ull);
  if (!ht.T) return fail(TPL_ERR_PANIC, "tile plan: more than 2^32 list entries");
  stats[0] = (uint64_t)tpl::check_tiles(m, p, tail, head, ctas, ht);
  stats[1] = ht.ntile;
  stats[2] = ht.lent.size();
  stats[3] = ht.piece.size();
  uint64_t pads = 0, lmax = 0, hsh = 1469598103934665603ull;
  for (uint32_t e : ht.lent) {
    pads += e == tpl::kEntPad;
    hsh = (hsh ^ e) * 1099511628211ull;
  }
  for (uint32_t e : ht.piece) hsh = (hsh ^ e) * 1099511628211ull;
  for (const uint4& x : ht.thdr) {
    lmax = std::max<uint64_t>(lmax, x.y);
    hsh = (hsh ^ x.x) * 1099511628211ull;
    hsh = (hsh ^ x.y) * 1099511628211ull;
    hsh = (hsh ^ x.z) * 1099511628211ull;
    hsh = (hsh ^ x.w) * 1099511628211ull;
  }
  stats[4] = pads;
  stats[5] = lmax;
  stats[6] = hsh;
  stats[7] = tpl::kFoldThreads;
  return TPL_OK;
}

int tpl_op_layout_check(tpl_op* op, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len, uint64_t stats[8]) {
  tpl::clear_error();
  if (!op || !stats || !tail || !head || (d_len && !d)) return fail(TPL_ERR_PANIC, "null argument");
  if (op->format != 2 || op->comm) return fail(TPL_ERR_PANIC, "tpl_op_layout_check needs an unsharded KKT handle");
  CUDA_TRY(cudaSetDevice(op->device));
  CUDA_TRY(cudaStreamSynchronize(op->stream));
  std::fill(stats, stats + 8, 0ull);
  const size_t m = op->inc.m, p = op->inc.p;
  stats[0] = op->device_built;
  stats[1] = op->blocked_ok;
  auto fetch = [&](auto& vec, const auto* dev, size_t count) -> int {
    vec.resize(count);
    if (count) CUDA_TRY(cudaMemcpy(vec.data(), dev, count * sizeof(vec[0]), cudaMemcpyDeviceToHost));
    return TPL_OK;
  };
  if (op->blocked_ok) {
    const tpl::BlockOp& bo = op->blk;
    tpl::HostBlocks h;
    h.GR = bo.GR; h.GC = bo.GC; h.PT = bo.PT; h.PH = bo.PH; h.Mpad = bo.Mpad; h.T = bo.tl.T; h.ntile = bo.tl.ntile;
    h.ring1 = bo.ring1; h.ring2 = bo.ring2; h.ring2v = bo.ring2v; h.lblk = bo.lblk; h.nl = bo.nl; h.ntb = bo.ntb;
    const size_t Gc = (size_t)bo.GR * bo.GC;
    if (int rc = fetch(h.cell_off, bo.cell_off, Gc + 1)) return rc;
    if (int rc = fetch(h.tbs, bo.tbs, bo.GR + 1)) return rc;
    if (int rc = fetch(h.hbs, bo.hbs, bo.GC + 1)) return rc;
    if (int rc = fetch(h.tbn, bo.tbn, h.tbs[bo.GR])) return rc;
    if (int rc = fetch(h.hbn, bo.hbn, h.hbs[bo.GC])) return rc;
    if (int rc = fetch(h.d, bo.d, bo.Mpad)) return rc;
    if (int rc = fetch(h.th, bo.th, bo.Mpad)) return rc;
    if (int rc = fetch(h.gidx, bo.gidx, bo.Mpad)) return rc;
    if (int rc = fetch(h.pdesc, bo.pdesc, bo.Mpad / tpl::kBStage)) return rc;
    if (int rc = fetch(h.thdr, bo.tl.thdr, Gc * bo.tl.ntile)) return rc;
    if (int rc = fetch(h.lent, bo.tl.lent, op->blk_lent_words)) return rc;
    h.ok = true;
    stats[2] = (uint64_t)tpl::check_blocks(m, p, tail, head, d, d_len, h);
    uint64_t hsh = 1469598103934665603ull;
    for (uint32_t e : h.lent) hsh = (hsh ^ e) * 1099511628211ull;
    for (uint32_t e : h.th) hsh = (hsh ^ e) * 1099511628211ull;
    for (uint32_t e : h.gidx) hsh = (hsh ^ e) * 1099511628211ull;
    stats[3] = hsh;
    stats[4] = h.lent.size();
    stats[5] = h.T;
  }
  {  // node -> arc lists against a host construction
    std::vector<uint64_t> ptr(p + 1, 0);
    for (size_t j = 0; j < m; ++j)
      if (tail[j] != head[j]) {
        ++ptr[tail[j] + 1];
        ++ptr[head[j] + 1];
      }
    for (size_t u = 0; u < p; ++u) ptr[u + 1] += ptr[u];
    std::vector<uint32_t> want(ptr[p]), got;
    std::vector<uint64_t> fill(ptr.begin(), ptr.end() - 1);
    for (size_t j = 0; j < m; ++j)
      if (tail[j] != head[j]) {
        want[fill[tail[j]]++] = (uint32_t)j;
        want[fill[head[j]]++] = (uint32_t)j | tpl::kSignBit;
      }
    if (int rc = fetch(got, op->inc.lr.ent_idx, want.size())) return rc;
    uint64_t bad = 0;
    for (size_t e = 0; e < want.size(); ++e) bad += want[e] != got[e];
    stats[6] = bad;
    stats[7] = want.size();
  }
  return TPL_OK;
}

int tpl_blocks_plan(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len, int ctas,
                    size_t smem_limit, int threads, uint64_t stats[16]) {
  tpl::clear_error();
  if (!tail || !head || !stats || (d_len && !d)) return fail(TPL_ERR_PANIC, "null argument");
  if (d_len > m) return tpl::fail_parameter_mismatch("d", m, d_len);
  for (size_t j = 0; j < m; ++j)
    if (tail[j] >= p || head[j] >= p)
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  tpl::HostBlocks hb;
  tpl::build_blocks(m, p, tail, head, d, d_len, ctas, smem_limit, hb, threads);
  std::fill(stats, stats + 16, 0ull);
  stats[0] = hb.ok;
  if (!hb.ok) return TPL_OK;
  stats[1] = hb.GR;
  stats[2] = hb.GC;
  stats[3] = hb.PT;
  stats[4] = hb.PH;
  stats[5] = hb.Mpad;
  stats[6] = hb.T;
  stats[7] = hb.ntile;
  stats[8] = hb.ring1 | ((uint64_t)hb.ring2 << 8) | ((uint64_t)hb.ring2v << 16);
  uint64_t cmax = 0, cmin = ~0ull, hsh = 1469598103934665603ull;
  for (uint32_t c = 0; c + 1 < hb.cell_off.size(); ++c) {
    cmax = std::max<uint64_t>(cmax, hb.cell_off[c + 1] - hb.cell_off[c]);
    cmin = std::min<uint64_t>(cmin, hb.cell_off[c + 1] - hb.cell_off[c]);
  }
  stats[9] = cmax;
  stats[10] = cmin;
  stats[11] = hb.lent.size();
  stats[12] = 0;
  for (const uint4& d : hb.pdesc) stats[12] += (d.x != tpl::kBNoPiece) + (d.y != tpl::kBNoPiece) + (d.z != tpl::kBNoPiece) + (d.w != tpl::kBNoPiece);
  for (uint32_t e : hb.lent) hsh = (hsh ^ e) * 1099511628211ull;
  for (uint32_t e : hb.piece) hsh = (hsh ^ e) * 1099511628211ull;
  for (uint32_t e : hb.th) hsh = (hsh ^ e) * 1099511628211ull;
  for (uint32_t e : hb.gidx) hsh = (hsh ^ e) * 1099511628211ull;
  stats[13] = hsh;
  // TPL_PLAN_CORRUPT=<kind> (tests of the checker itself): one deliberate defect in the lists of the first non-empty tile before
  // they are checked -- 1: the minus bit of an entry flipped, 2: the slot a node's sum is flushed to replaced by the dummy,
  // 3: a chain depth changed, 4: two entries of one slice exchanged, 5: the slot of a slice's last node changed.
  if (const char* e = std::getenv("TPL_PLAN_CORRUPT")) {
    const int kind = std::atoi(e);
    const uint32_t B = tpl::kFoldThreads, PL = hb.PT + hb.PH;
    for (const uint4& h4 : hb.thdr) {
      const uint32_t L = h4.y & 0xffffffu;
      if (L < 3) continue;
      uint32_t* blk = hb.lent.data() + h4.x;
      if (kind == 1) {
        blk[B * 2] ^= tpl::kBEntMinus;
      } else if (kind == 2) {
        bool hit = false;
        for (size_t w = 2 * (size_t)B; w < (size_t)(L + 1) * B && !hit; ++w)  // rows 2 .. L: not the first entry of a slice
          if (blk[w] & tpl::kBEntNew) {
            blk[w] = (blk[w] & ~(tpl::kBEntNodeMask << tpl::kBEntNodeShift)) | (PL << tpl::kBEntNodeShift);
            hit = true;
          }
        if (!hit) continue;
      } else if (kind == 3) {
        blk[1] = (blk[1] & ~0xffu) | (((blk[1] & 0xffu) + 1u) & 0xffu);
      } else if (kind == 4) {
        std::swap(blk[(size_t)1 * B], blk[(size_t)L * B]);
      } else if (kind == 5) {
        blk[0] = (blk[0] & 0xffu) | (((((blk[0] >> 8) & tpl::kBEntNodeMask) + 1u) % (PL + 1u)) << 8);
      }
      break;
    }
  }
  stats[14] = (uint64_t)tpl::check_blocks(m, p, tail, head, d, d_len, hb);
  stats[15] = tpl::block_smem_bytes(hb.PT + hb.PH, hb.T, (int)hb.ring2, hb.lblk, hb.nl, hb.ntb, true, false, hb.ntile);
  stats[8] |= (uint64_t)hb.nl << 24;
  return TPL_OK;
}

int tpl_cells_plan(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int ctas, size_t smem_limit,
                   uint64_t stats[16]) {
  tpl::clear_error();
  if (!tail || !head || !stats) return fail(TPL_ERR_PANIC, "null argument");
  for (size_t j = 0; j < m; ++j)
    if (tail[j] >= p || head[j] >= p)
      return fail(TPL_ERR_SPARSE_CONSTRUCTION, "Internal error: Failed to construct the sparse matrix from triplets.");
  tpl::HostCells hc;
  tpl::build_cells(m, p, tail, head, ctas, smem_limit, hc);
  std::fill(stats, stats + 16, 0ull);
  stats[0] = hc.ok;
  stats[1] = hc.GR;
  stats[2] = hc.GC;
  stats[3] = hc.Amax;
  stats[4] = hc.L;
  stats[5] = hc.max_rows;
  stats[6] = hc.max_groups;
  stats[7] = hc.max_lines;
  stats[8] = hc.max_slots;
  stats[9] = hc.max_own;
  stats[10] = hc.inbox_atoms;
  if (!hc.hdr.empty()) {
    uint64_t amax = 0, amin = ~0ull;
    for (uint32_t c = 0; c < hc.Gc; ++c) {
      amax = std::max<uint64_t>(amax, hc.hdr[(size_t)c * 8]);
      amin = std::min<uint64_t>(amin, hc.hdr[(size_t)c * 8]);
    }
    stats[11] = amax;
    stats[12] = amin;
  }
  if (hc.ok) {
    stats[13] = tpl::cell_smem_bytes(tpl::cells_probe(hc), true);
    stats[14] = (uint64_t)tpl::check_cells(m, p, tail, head, hc);
    uint32_t ca = 0, cb = 0;
    tpl::cell_conflicts(hc, ca, cb);
    stats[15] = (uint64_t)ca | ((uint64_t)cb << 32);
  }
  return TPL_OK;
}

int tpl_op_set_mode(tpl_op* op, int mode) {
  if (!op || mode < 0 || mode > 5) return fail(TPL_ERR_PANIC, "invalid mode");
  op->mode = mode;
  return TPL_OK;
}

// which kernels a whole-pass solve through this handle runs (see tpl_op_set_mode)
static const char* shape_name(const tpl_op* op);
const char* tpl_op_kernel_shape(const tpl_op* op) { return op ? shape_name(op) : ""; }

static bool replicated(const tpl_op* op) { return op->inner != nullptr && op->mode == 0; }

int tpl_op_last_timing(const tpl_op* op, double* pass_one_ms, double* pass_two_ms, double* gemv_ms) {
  if (!op) return fail(TPL_ERR_PANIC, "null argument");
  if (replicated(op)) return tpl_op_last_timing(op->inner, pass_one_ms, pass_two_ms, gemv_ms);
  DeviceGuard g(op->device);
  double* outs[3] = {pass_one_ms, pass_two_ms, gemv_ms};
  for (int i = 0; i < 3; ++i) {
    if (!outs[i]) continue;
    *outs[i] = 0.0;
    if (!op->timed[i]) continue;
    float ms = 0.f;
    CUDA_TRY(cudaEventSynchronize(op->ev[2 * i + 1]));
    CUDA_TRY(cudaEventElapsedTime(&ms, op->ev[2 * i], op->ev[2 * i + 1]));
    *outs[i] = ms;
  }
  return TPL_OK;
}

}  // extern "C"

// ============================================================================ launches
namespace {

template <class KERNEL, class ARGS>
int launch_resident(tpl_op* op, KERNEL kernel, const ARGS& args, size_t smem) {
  void* params[] = {&op->inc, &op->res, const_cast<ARGS*>(&args)};
  CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(op->G), dim3(tpl::kBlock), params,
                                       smem, op->stream));
  op->launches += 1;
  return TPL_OK;
}

template <class KERNEL, class OP, class ARGS>
int launch_coop(tpl_op* op, KERNEL kernel, const OP& dop, const ARGS& args, size_t smem = SIZE_MAX) {
  void* params[] = {const_cast<OP*>(&dop), const_cast<ARGS*>(&args)};
  CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(op->G), dim3(tpl::kBlock), params,
                                       smem == SIZE_MAX ? op->smem_bytes : smem, op->stream));
  op->launches += 1;
  return TPL_OK;
}

// mode 0: cell kernels (2-D partition, shared-memory resident) whenever a cell fits and the whole pass runs in one
// launch, else the chunk-resident kernels (mode 4 forces these), else (and in mode 2) the streaming kernels with tiled
// node sums; mode 3 (and any operator the tiles do not fit) the gather kernels
bool use_cells(const tpl_op* op) { return op->format == 2 && op->cells_ok && op->mode == 0; }
bool use_resident(const tpl_op* op) { return op->format == 2 && op->resident_ok && (op->mode == 0 || op->mode == 4); }

template <class KERNEL, class ARGS>
int launch_cells(tpl_op* op, KERNEL kernel, const ARGS& args, size_t smem) {
  CUDA_TRY(cudaMemsetAsync(op->cell_xchg, 0, op->cell_xchg_bytes, op->stream));
  void* params[] = {&op->inc, &op->cell, const_cast<ARGS*>(&args)};
  CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(op->cell.Gc), dim3(tpl::kBlock), params,
                                       smem, op->stream));
  op->launches += 1;
  return TPL_OK;
}
// mode 0: the blocked streaming kernels from ~1M arcs on one GPU (cells of >= 6.5k arcs; measured: 0.51 vs 0.49 of the HBM
// roofline at 1M arcs, 0.67 vs 0.57 at 2M, 0.95 vs 0.74 at 20M, but 0.37 vs 0.38 at 700k, where a sweep is two tiles long), the
// tiled ones below; mode 5 forces the blocked kernels, mode 2 the tiled ones.  A sharded handle runs the family that owns its
// exchange block.
bool use_blocked(const tpl_op* op) {
  if (op->format != 2 || !op->blocked_ok) return false;
  if (op->mode == 5 || (op->comm && op->mode == 0)) return true;
  return op->mode == 0 && (op->blk_max_cell >= 6528 || !op->tiled_ok);
}
bool use_tiled(const tpl_op* op) { return op->format == 2 && op->tiled_ok && (op->mode == 0 || op->mode == 2); }

}  // namespace
static const char* shape_name(const tpl_op* op) {
  if (op->comm && op->inner && op->mode == 0) return "replicated";
  if (op->comm) return op->fab_connected && op->mode == 0 ? (op->blocked_ok ? "sharded-blocked" : "sharded-fused") : "sharded";
  if (op->format == 3) return "dense";
  if (op->format != 2) return "csr";
  if (op->mode == 1) return "gather";
  if (use_cells(op)) return "cells";
  if (use_resident(op)) return "chunks";
  if (use_blocked(op)) return "blocked";
  if (use_tiled(op)) return "tiled";
  return "gather";
}
namespace {

template <class KERNEL, class ARGS>
int launch_tiled(tpl_op* op, KERNEL kernel, const ARGS& args, size_t smem) {
  void* params[] = {&op->inc, &op->tile, const_cast<ARGS*>(&args)};
  CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(op->G), dim3(tpl::kBlock), params,
                                       smem, op->stream));
  op->launches += 1;
  return TPL_OK;
}

template <class KERNEL, class ARGS>
int launch_blocked(tpl_op* op, KERNEL kernel, ARGS args, size_t smem) {
  for (int i = 0; i < 3; ++i) args.buf[i] = op->bbuf[i];  // the rotating vectors live in cell order
  void* params[] = {&op->inc, &op->blk, &args};
  CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(op->blk.GR * op->blk.GC), dim3(tpl::kBlock),
                                       params, smem, op->stream));
  op->launches += 1;
  return TPL_OK;
}

int launch_pass1(tpl_op* op, const tpl::Pass1Args& a, bool whole_pass) {
  const bool with_v = a.V != nullptr;
  if (whole_pass && use_cells(op))
    return with_v ? launch_cells(op, tpl::pass1_cell_kernel<true>, a, op->smem_cell1)
                  : launch_cells(op, tpl::pass1_cell_kernel<false>, a, op->smem_cell1);
  if (whole_pass && use_resident(op))
    return with_v ? launch_resident(op, tpl::pass1_resident_kernel<true>, a, op->smem_res1)
                  : launch_resident(op, tpl::pass1_resident_kernel<false>, a, op->smem_res1);
  if (whole_pass && use_blocked(op))
    return with_v ? launch_blocked(op, tpl::pass1_blocked_kernel<true>, a, op->smem_blk1)
                  : launch_blocked(op, tpl::pass1_blocked_kernel<false>, a, op->smem_blk1);
  if (whole_pass && use_tiled(op))
    return with_v ? launch_tiled(op, tpl::pass1_tiled_kernel<true>, a, op->smem_tile1)
                  : launch_tiled(op, tpl::pass1_tiled_kernel<false>, a, op->smem_tile1);
  if (op->format == 3)
    return with_v ? launch_coop(op, tpl::pass1_dense_kernel<true>, op->dense, a) : launch_coop(op, tpl::pass1_dense_kernel<false>, op->dense, a);
  if (op->format == 2)
    return with_v ? launch_coop(op, tpl::pass1_kernel<tpl::IncidenceOp, true>, op->inc, a)
                  : launch_coop(op, tpl::pass1_kernel<tpl::IncidenceOp, false>, op->inc, a);
  return with_v ? launch_coop(op, tpl::pass1_csr_kernel<true>, op->sell, a) : launch_coop(op, tpl::pass1_csr_kernel<false>, op->sell, a);
}
int launch_pass2(tpl_op* op, const tpl::Pass2Args& a) {
  const bool with_v = a.V != nullptr;
  if (use_cells(op))
    return with_v ? launch_cells(op, tpl::pass2_cell_kernel<true>, a, op->smem_cell2)
                  : launch_cells(op, tpl::pass2_cell_kernel<false>, a, op->smem_cell2);
  if (use_resident(op))
    return with_v ? launch_resident(op, tpl::pass2_resident_kernel<true>, a, op->smem_res2)
                  : launch_resident(op, tpl::pass2_resident_kernel<false>, a, op->smem_res2);
  if (use_blocked(op))
    return with_v ? launch_blocked(op, tpl::pass2_blocked_kernel<true>, a, op->smem_blk2v)
                  : launch_blocked(op, tpl::pass2_blocked_kernel<false>, a, op->smem_blk2);
  if (use_tiled(op))
    return with_v ? launch_tiled(op, tpl::pass2_tiled_kernel<true>, a, op->smem_tile2)
                  : launch_tiled(op, tpl::pass2_tiled_kernel<false>, a, op->smem_tile2);
  if (op->format == 3)
    return with_v ? launch_coop(op, tpl::pass2_dense_kernel<true>, op->dense, a) : launch_coop(op, tpl::pass2_dense_kernel<false>, op->dense, a);
  if (op->format == 2)
    return with_v ? launch_coop(op, tpl::pass2_kernel<tpl::IncidenceOp, true>, op->inc, a)
                  : launch_coop(op, tpl::pass2_kernel<tpl::IncidenceOp, false>, op->inc, a);
  return with_v ? launch_coop(op, tpl::pass2_csr_kernel<true>, op->sell, a) : launch_coop(op, tpl::pass2_csr_kernel<false>, op->sell, a);
}

// Brings b to the device (no copy when it already lives there).
int stage_b(tpl_op* op, const double* b, const double** b_dev) {
  if (is_device_ptr(b)) {
    *b_dev = b;
    return TPL_OK;
  }
  CUDA_TRY(cudaMemcpyAsync(op->b_d, b, sizeof(double) * op->n, cudaMemcpyHostToDevice, op->stream));
  *b_dev = op->b_d;
  return TPL_OK;
}

int reset_sync_state(tpl_op* op) {
  tpl::State st{};
  st.s_cur = 1.0;
  st.s_prev = 1.0;
  st.status = tpl::ST_RUNNING;
  st.epoch = op->fab_connected ? op->fab_epoch : 0;
  std::memcpy(op->h_pin, &st, sizeof st);
  CUDA_TRY(cudaMemsetAsync(op->slots, 0, sizeof(uint4) * 2 * op->G * tpl::kSlotAtoms, op->stream));
  CUDA_TRY(cudaMemcpyAsync(op->coef_d, op->h_pin, sizeof st, cudaMemcpyHostToDevice, op->stream));
  return TPL_OK;
}

struct Decomp {  // LanczosDecomposition (src/algorithms/mod.rs:94-108) on the host
  std::vector<double> alphas, betas;
  size_t steps = 0;
  double b_norm = 0.0;
};

// Reads [State | alphas | betas] back in one copy and applies the reference's push rules
// (lanczos_two_pass.rs:86-98: alpha every step, beta only when not broken down and i < k-1).
int fetch_decomp(tpl_op* op, size_t k, Decomp& out, int& status) {
  const size_t doubles = kHeaderDoubles + 2 * op->coef_cap;
  (void)k;
  CUDA_TRY(cudaMemcpyAsync(op->h_pin, op->coef_d, doubles * sizeof(double), cudaMemcpyDeviceToHost, op->stream));
  CUDA_TRY(cudaStreamSynchronize(op->stream));
  tpl::State st;
  std::memcpy(&st, op->h_pin, sizeof st);
  status = st.status;
  if (op->fab_connected) op->fab_epoch = st.epoch;
  out.b_norm = st.b_norm;
  out.steps = (size_t)st.steps;
  const double* al = op->h_pin + kHeaderDoubles;
  const double* be = op->h_pin + kHeaderDoubles + op->coef_cap;
  out.alphas.assign(al, al + out.steps);
  out.betas.assign(be, be + (out.steps ? out.steps - 1 : 0));
  return TPL_OK;
}

// ---------------------------------------------------------------- arc-partitioned (sharded) drivers
template <class KERNEL>
int launch_shard(tpl_op* op, KERNEL kernel, const tpl::ShardArgs& a, size_t smem) {
  void* params[] = {&op->inc, const_cast<tpl::ShardArgs*>(&a)};
  CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(op->G), dim3(tpl::kBlock), params, smem,
                                       op->stream));
  op->launches += 1;
  return TPL_OK;
}
int shard_allreduce(tpl_op* op, double* buf, size_t count) {
  NCCL_TRY(nccl::api().AllReduce(buf, buf, count, nccl::kFloat64, nccl::kSum, op->comm, op->stream));
  return TPL_OK;
}
tpl::ShardArgs shard_args(tpl_op* op, const double* b_dev) {
  tpl::ShardArgs a{};
  for (int i = 0; i < 3; ++i) a.buf[i] = op->buf[i];
  a.b = b_dev;
  a.alphas = op->alphas_d();
  a.betas = op->betas_d();
  a.y = op->y_d();
  a.red = op->red_d;
  a.red_in = op->red_d;
  a.red2 = op->red2_d;
  a.st = op->st_d();
  a.gs = op->gs();
  a.tol = tpl::kBreakdownTol;
  return a;
}

// Pass 1 / basis generation on rank-local slices: 2 launches + 2 all-reduces per step, all asynchronous; breakdown is
// detected on the device (later launches return at once), the host reads the state back once at the end.
int run_pass_one_sharded(tpl_op* op, const double* b_dev, size_t k, double* V_dev, size_t ldv, tpl::ShardArgs a) {
  const size_t p1 = (size_t)op->inc.p + 1;
  a.V = V_dev;
  a.ldv = ldv;
  if (int rc = launch_shard(op, tpl::shard_init_kernel, a, 0)) return rc;
  if (int rc = shard_allreduce(op, op->red2_d, 1)) return rc;
  for (size_t j = 0; j <= k; ++j) {
    a.j = (int)j;
    a.head_only = j == k;
    if (int rc = V_dev ? launch_shard(op, tpl::shard_phase_a_kernel<true>, a, op->smem_bytes)
                       : launch_shard(op, tpl::shard_phase_a_kernel<false>, a, op->smem_bytes))
      return rc;
    if (j == k) break;
    if (int rc = shard_allreduce(op, op->red_d, p1)) return rc;
    if (int rc = launch_shard(op, tpl::shard_phase_b_kernel, a, 0)) return rc;
    if (int rc = shard_allreduce(op, op->red2_d, 1)) return rc;
  }
  return TPL_OK;
}

int run_pass_two_sharded(tpl_op* op, size_t steps, double b_norm, double* x_dev, double* V_dev, size_t ldv,
                         tpl::ShardArgs a) {
  const size_t p1 = (size_t)op->inc.p + 1;
  a.x = x_dev;
  a.V = V_dev;
  a.ldv = ldv;
  a.steps = (int)steps;
  a.b_norm = b_norm;
  a.j = -1;
  auto launch = [&]() {
    return V_dev ? launch_shard(op, tpl::shard_pass2_kernel<true>, a, op->smem_bytes)
                 : launch_shard(op, tpl::shard_pass2_kernel<false>, a, op->smem_bytes);
  };
  if (int rc = launch()) return rc;
  for (size_t j = 0; j + 1 < steps; ++j) {
    a.j = (int)j;
    a.head_only = 0;
    a.red = op->red_d + (j & 1) * p1;
    a.red_in = op->red_d + ((j + 1) & 1) * p1;
    if (int rc = launch()) return rc;
    if (int rc = shard_allreduce(op, a.red, p1)) return rc;
  }
  if (steps > 1) {
    const size_t j = steps - 1;
    a.j = (int)j;
    a.head_only = 1;
    a.red = op->red_d + (j & 1) * p1;
    a.red_in = op->red_d + ((j + 1) & 1) * p1;
    if (int rc = launch()) return rc;
  }
  return TPL_OK;
}

// lanczos_pass_one / basis generation of lanczos_standard on a device-resident b.
int run_pass_one(tpl_op* op, const double* b_dev, size_t k, double* V_dev, size_t ldv, tpl_step_callback cb,
                 void* user, Decomp& out) {
  if (k == 0) return fail(TPL_ERR_PANIC, "capacity overflow (k == 0; the reference panics in Vec::with_capacity(k - 1))");
  if (k > 0x7ffffff0ull) return fail(TPL_ERR_PANIC, "k too large");
  if (int rc = ensure_coef(op, k)) return rc;
  if (int rc = reset_sync_state(op)) return rc;
  tpl::Pass1Args a{};
  for (int i = 0; i < 3; ++i) a.buf[i] = op->buf[i];
  a.b = b_dev;
  a.alphas = op->alphas_d();
  a.betas = op->betas_d();
  a.V = V_dev;
  a.ldv = ldv;
  a.n = op->n;
  a.st = op->st_d();
  a.gs = op->gs();
  a.tol = tpl::kBreakdownTol;
  int status = tpl::ST_RUNNING;
  CUDA_TRY(cudaEventRecord(op->ev[0], op->stream));
  if (op->comm && cb) return fail(TPL_ERR_COMM, "step callbacks are not supported on a sharded operator");
  if (op->comm && !(op->fab_connected && op->mode == 0)) {
    if (int rc = run_pass_one_sharded(op, b_dev, k, V_dev, ldv, shard_args(op, b_dev))) return rc;
    CUDA_TRY(cudaEventRecord(op->ev[1], op->stream));
    if (int rc = fetch_decomp(op, k, out, status)) return rc;
  } else if (!cb && op->mode != 1) {
    a.j_begin = 0;
    a.j_end = (int)k;
    if (int rc = launch_pass1(op, a, true)) return rc;
    CUDA_TRY(cudaEventRecord(op->ev[1], op->stream));
    if (int rc = fetch_decomp(op, k, out, status)) return rc;
  } else {
    // one cooperative launch per step: lets the host run the LanczosCallback (lanczos.rs:93-106)
    for (size_t j = 0; j < k; ++j) {
      a.j_begin = (int)j;
      a.j_end = (int)j + 1;
      if (int rc = launch_pass1(op, a, false)) return rc;
      if (!cb && j + 1 < k) continue;  // mode 1 without a callback: no host round trip needed
      if (int rc = fetch_decomp(op, k, out, status)) return rc;
      // the reference calls the callback after EVERY completed step, also the one whose beta is <= tol, and only then
      // looks at the breakdown (lanczos.rs:93-112); a zero b never completes a step
      if (cb && status != tpl::ST_ZERO_B && out.steps == j + 1) {
        if (!cb(out.steps, V_dev, ldv, out.alphas.data(), out.betas.data(), user)) break;
      }
      if (status != tpl::ST_RUNNING) break;
    }
    CUDA_TRY(cudaEventRecord(op->ev[1], op->stream));
    if (int rc = fetch_decomp(op, k, out, status)) return rc;
  }
  op->timed[0] = true;
  if (status == tpl::ST_ZERO_B) return tpl::fail_input("Input vector `b` must not be a zero vector.");  // mod.rs:268-273
  return TPL_OK;
}

// lanczos_pass_two_impl on device-resident b / x (host alphas, betas, y).
int run_pass_two(tpl_op* op, const double* b_dev, const double* alphas, const double* betas, size_t steps,
                 double b_norm, const double* y, size_t y_len, double* x_dev, double* V_dev, size_t ldv) {
  if (steps != y_len) return tpl::fail_parameter_mismatch("y_k", steps, y_len);  // lanczos_two_pass.rs:220-227
  if (b_norm <= tpl::kBreakdownTol)                                                // :229-235
    return tpl::fail_input("The initial vector `b` must not be a zero vector.");
  if (steps == 0) {                                                                // :237-244
    CUDA_TRY(cudaMemsetAsync(x_dev, 0, sizeof(double) * op->n, op->stream));
    return TPL_OK;
  }
  if (int rc = ensure_coef(op, steps)) return rc;
  if (int rc = reset_sync_state(op)) return rc;
  // coefficients: [alphas | betas | y] -> device in one copy through the pinned mirror
  double* hp = op->h_pin + kHeaderDoubles;
  std::memcpy(hp, alphas, steps * sizeof(double));
  if (steps > 1) std::memcpy(hp + op->coef_cap, betas, (steps - 1) * sizeof(double));
  std::memcpy(hp + 2 * op->coef_cap, y, steps * sizeof(double));
  CUDA_TRY(cudaMemcpyAsync(op->alphas_d(), hp, 3 * op->coef_cap * sizeof(double), cudaMemcpyHostToDevice, op->stream));
  tpl::Pass2Args a{};
  for (int i = 0; i < 3; ++i) a.buf[i] = op->buf[i];
  a.b = b_dev;
  a.alphas = op->alphas_d();
  a.betas = op->betas_d();
  a.y = op->y_d();
  a.x = x_dev;
  a.V = V_dev;
  a.ldv = ldv;
  a.n = op->n;
  a.steps = (int)steps;
  a.b_norm = b_norm;
  a.st = op->st_d();
  a.gs = op->gs();
  CUDA_TRY(cudaEventRecord(op->ev[2], op->stream));
  if (op->comm && !(op->fab_connected && op->mode == 0)) {
    if (int rc = run_pass_two_sharded(op, steps, b_norm, x_dev, V_dev, ldv, shard_args(op, b_dev))) return rc;
  } else if (int rc = launch_pass2(op, a)) {
    return rc;
  }
  CUDA_TRY(cudaEventRecord(op->ev[3], op->stream));
  op->timed[1] = true;
  if (op->fab_connected && op->mode == 0) {
    // the next fused pass continues from the epoch this one ended with (the kernel stored it in the state block); the
    // read-back is bookkeeping between passes, outside the pass-2 event pair
    CUDA_TRY(cudaMemcpyAsync(op->h_pin, op->coef_d, sizeof(tpl::State), cudaMemcpyDeviceToHost, op->stream));
    CUDA_TRY(cudaStreamSynchronize(op->stream));
    tpl::State st;
    std::memcpy(&st, op->h_pin, sizeof st);
    op->fab_epoch = st.epoch;
  }
  return TPL_OK;
}

int gemv_vy(tpl_op* op, const double* V_dev, size_t ldv, size_t steps, const double* y_host, double b_norm,
            double* x_dev) {
  if (int rc = ensure_coef(op, steps)) return rc;
  double* hp = op->h_pin + kHeaderDoubles + 2 * op->coef_cap;
  std::memcpy(hp, y_host, steps * sizeof(double));
  CUDA_TRY(cudaMemcpyAsync(op->y_d(), hp, steps * sizeof(double), cudaMemcpyHostToDevice, op->stream));
  const int block = 256;
  const int grid = (int)std::min<size_t>((op->n + block - 1) / block, (size_t)op->G * 16);
  CUDA_TRY(cudaEventRecord(op->ev[4], op->stream));
  tpl::gemv_vy_kernel<<<grid, block, steps * sizeof(double), op->stream>>>(V_dev, ldv, op->n, (int)steps, op->y_d(),
                                                                          b_norm, x_dev);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(op->ev[5], op->stream));
  op->timed[2] = true;
  op->launches += 1;
  return TPL_OK;
}

int ensure_sweep_buffer(tpl_op* op, size_t elems) {
  if (elems <= op->sweep_elems) return TPL_OK;
  if (op->sweep_d) {
    if (int rc = dev_free(op, op->sweep_d)) return rc;
    op->sweep_d = nullptr;
    op->sweep_elems = 0;
  }
  if (int rc = dev_alloc(op, &op->sweep_d, elems)) return rc;
  op->sweep_elems = elems;
  return TPL_OK;
}
int ensure_sweep_coef(tpl_op* op, size_t elems) {
  if (elems <= op->sweep_y_elems) return TPL_OK;
  if (op->sweep_y_d) {
    if (int rc = dev_free(op, op->sweep_y_d)) return rc;
    op->sweep_y_d = nullptr;
    op->sweep_y_elems = 0;
  }
  if (int rc = dev_alloc(op, &op->sweep_y_d, elems)) return rc;
  op->sweep_y_elems = elems;
  return TPL_OK;
}

int ensure_internal_basis(tpl_op* op, size_t elems) {
  if (elems <= op->V_int_elems) return TPL_OK;
  if (op->V_int) {
    if (int rc = dev_free(op, op->V_int)) return rc;
    op->V_int = nullptr;
    op->V_int_elems = 0;
  }
  if (int rc = dev_alloc(op, &op->V_int, elems)) return rc;
  op->V_int_elems = elems;
  return TPL_OK;
}

// copies a column-major n x cols device matrix (leading dimension n) into a host matrix with ldv
int basis_to_host(tpl_op* op, const double* V_dev, double* V_host, size_t ldv, size_t cols) {
  if (!cols) return TPL_OK;
  CUDA_TRY(cudaMemcpy2DAsync(V_host, ldv * sizeof(double), V_dev, (size_t)op->n * sizeof(double),
                             (size_t)op->n * sizeof(double), cols, cudaMemcpyDeviceToHost, op->stream));
  return TPL_OK;
}

int finish_x(tpl_op* op, const double* x_dev, double* x) {
  if (x_dev != x) CUDA_TRY(cudaMemcpyAsync(x, x_dev, sizeof(double) * op->n, cudaMemcpyDeviceToHost, op->stream));
  CUDA_TRY(cudaStreamSynchronize(op->stream));
  return TPL_OK;
}

int call_ftk(tpl_ftk_solver f, void* user, const Decomp& d, std::vector<double>& y) {
  // closure call + shape validation (solvers.rs:71-87, 155-165)
  y.assign(d.steps + 1, 0.0);
  size_t y_len = d.steps;
  const std::string before = tpl::g_err;
  int rc = f(d.alphas.data(), d.alphas.size(), d.betas.data(), d.betas.size(), y.data(), &y_len, user);
  if (rc != 0) {
    std::string inner = tpl::g_err != before && !tpl::g_err.empty() ? tpl::g_err : ("callback returned " + std::to_string(rc));
    return tpl::fail_solver(inner.c_str());
  }
  if (y_len != d.steps) return tpl::fail_parameter_mismatch("y_k_prime", d.steps, y_len);
  y.resize(d.steps);
  return TPL_OK;
}

// ---------------------------------------------------------------- replicated execution of a sharded handle (see tpl_op::inner)
// rank-local vector [arc slice | p node entries] (host or device) -> full device vector [m arcs | p nodes]: every rank
// contributes its arc slice to a zero-padded copy, one all-reduce (sum) assembles the arc part; the node part is replicated
// by contract and taken from this rank's copy.
int rep_gather(tpl_op* op, const double* v_local, double* full) {
  const size_t m = op->rep_m, p = op->inc.p, lo = op->rep_lo, ml = op->rep_hi - op->rep_lo;
  const cudaMemcpyKind kind = is_device_ptr(v_local) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CUDA_TRY(cudaMemsetAsync(full, 0, m * sizeof(double), op->stream));
  if (ml) CUDA_TRY(cudaMemcpyAsync(full + lo, v_local, ml * sizeof(double), kind, op->stream));
  CUDA_TRY(cudaMemcpyAsync(full + m, v_local + ml, p * sizeof(double), kind, op->stream));
  if (kind == cudaMemcpyHostToDevice) CUDA_TRY(cudaStreamSynchronize(op->stream));  // the caller's buffer is free again
  NCCL_TRY(nccl::api().AllReduce(full, full, m, nccl::kFloat64, nccl::kSum, op->comm, op->stream));
  return TPL_OK;
}
// `cols` columns of a full device matrix (leading dimension m + p) -> this rank's rows of the caller's matrix (ld_local)
int rep_scatter(tpl_op* op, const double* full, double* v_local, size_t ld_local, size_t cols) {
  const size_t m = op->rep_m, p = op->inc.p, lo = op->rep_lo, ml = op->rep_hi - op->rep_lo, ldf = m + p;
  const cudaMemcpyKind kind = is_device_ptr(v_local) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (!cols) return TPL_OK;
  if (ml)
    CUDA_TRY(cudaMemcpy2DAsync(v_local, ld_local * sizeof(double), full + lo, ldf * sizeof(double), ml * sizeof(double), cols, kind,
                               op->stream));
  CUDA_TRY(cudaMemcpy2DAsync(v_local + ml, ld_local * sizeof(double), full + m, ldf * sizeof(double), p * sizeof(double), cols, kind,
                             op->stream));
  CUDA_TRY(cudaStreamSynchronize(op->stream));
  return TPL_OK;
}
int rep_basis(tpl_op* op, size_t elems) {
  if (elems <= op->rep_V_elems) return TPL_OK;
  if (op->rep_V) {
    if (int rc = dev_free(op, op->rep_V)) return rc;
    op->rep_V = nullptr;
    op->rep_V_elems = 0;
  }
  if (int rc = dev_alloc(op, &op->rep_V, elems)) return rc;
  op->rep_V_elems = elems;
  return TPL_OK;
}

}  // namespace

// ============================================================================ public compute entry points
extern "C" {

int tpl_op_apply(tpl_op* op, const double* x, double* y) {
  tpl::clear_error();
  if (!op || !x || !y) return fail(TPL_ERR_PANIC, "null argument");
  DeviceGuard g(op->device);
  if (replicated(op)) {
    if (int rc = rep_gather(op, x, op->rep_in)) return rc;
    if (int rc = tpl_op_apply(op->inner, op->rep_in, op->rep_out)) return rc;
    return rep_scatter(op, op->rep_out, y, op->n, 1);
  }
  const double* x_dev = nullptr;
  if (int rc = stage_b(op, x, &x_dev)) return rc;
  double* y_dev = is_device_ptr(y) ? y : op->x_d;
  if (op->format == 3)
    tpl::apply_dense_kernel<<<op->G, tpl::kBlock, op->smem_bytes, op->stream>>>(op->dense, x_dev, y_dev);
  else if (op->format == 2)
    tpl::apply_kernel<tpl::IncidenceOp><<<op->G, tpl::kBlock, op->smem_bytes, op->stream>>>(op->inc, x_dev, y_dev);
  else
    tpl::apply_csr_kernel<<<op->G, tpl::kBlock, op->smem_bytes, op->stream>>>(op->sell, x_dev, y_dev);
  CUDA_TRY(cudaGetLastError());
  op->launches += 1;
  if (op->comm)  // node rows of the local operator are partial sums over this rank's arcs
    if (int rc = shard_allreduce(op, y_dev + op->inc.m, op->inc.p)) return rc;
  return finish_x(op, y_dev, y);
}

int tpl_pass_one(tpl_op* op, const double* b, size_t k, double* alphas, double* betas, size_t* steps,
                 double* b_norm) {
  tpl::clear_error();
  if (!op || !b || !alphas || !steps || !b_norm || (k > 1 && !betas)) return fail(TPL_ERR_PANIC, "null argument");
  DeviceGuard g(op->device);
  if (replicated(op)) {
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    return tpl_pass_one(op->inner, op->rep_in, k, alphas, betas, steps, b_norm);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  Decomp d;
  int rc = run_pass_one(op, b_dev, k, nullptr, 0, nullptr, nullptr, d);
  *b_norm = d.b_norm;
  if (rc) return rc;
  *steps = d.steps;
  std::copy(d.alphas.begin(), d.alphas.end(), alphas);
  if (!d.betas.empty()) std::copy(d.betas.begin(), d.betas.end(), betas);
  return TPL_OK;
}

int tpl_pass_two(tpl_op* op, const double* b, const double* alphas, const double* betas, size_t steps, double b_norm,
                 const double* y, size_t y_len, double* x, double* V, size_t ldv) {
  tpl::clear_error();
  if (!op || !b || !x || (steps && (!alphas || !y)) || (steps > 1 && !betas)) return fail(TPL_ERR_PANIC, "null argument");
  if (V && ldv < op->n) return tpl::fail_parameter_mismatch("ldv", op->n, ldv);
  DeviceGuard g(op->device);
  if (replicated(op)) {
    const size_t nf = op->rep_m + op->inc.p;
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    if (V && steps == y_len && steps > 0)
      if (int rc = rep_basis(op, nf * steps)) return rc;
    if (int rc = tpl_pass_two(op->inner, op->rep_in, alphas, betas, steps, b_norm, y, y_len, op->rep_out, V ? op->rep_V : nullptr, nf))
      return rc;
    if (V && steps > 0)
      if (int rc = rep_scatter(op, op->rep_V, V, ldv, steps)) return rc;
    return rep_scatter(op, op->rep_out, x, op->n, 1);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  double* x_dev = is_device_ptr(x) ? x : op->x_d;
  double* V_dev = V;
  size_t ld_dev = ldv;
  const bool v_host = V && !is_device_ptr(V);
  if (v_host && steps == y_len && steps > 0) {
    if (int rc = ensure_internal_basis(op, (size_t)op->n * steps)) return rc;
    V_dev = op->V_int;
    ld_dev = op->n;
  }
  if (int rc = run_pass_two(op, b_dev, alphas, betas, steps, b_norm, y, y_len, x_dev, V_dev, ld_dev)) return rc;
  if (v_host && steps > 0)
    if (int rc = basis_to_host(op, V_dev, V, ldv, steps)) return rc;
  return finish_x(op, x_dev, x);
}

int tpl_standard(tpl_op* op, const double* b, size_t k, double* V, size_t ldv, double* alphas, double* betas,
                 size_t* steps, double* b_norm, tpl_step_callback cb, void* user) {
  tpl::clear_error();
  if (!op || !b || !V || !alphas || !steps || !b_norm || (k > 1 && !betas)) return fail(TPL_ERR_PANIC, "null argument");
  if (k == 0) return fail(TPL_ERR_PANIC, "capacity overflow (k == 0; the reference panics in Vec::with_capacity(k - 1))");
  if (ldv < op->n) return tpl::fail_parameter_mismatch("ldv", op->n, ldv);
  DeviceGuard g(op->device);
  if (replicated(op)) {
    if (cb) return fail(TPL_ERR_COMM, "step callbacks are not supported on a sharded operator");
    const size_t nf = op->rep_m + op->inc.p;
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    if (int rc = rep_basis(op, nf * k)) return rc;
    if (int rc = tpl_standard(op->inner, op->rep_in, k, op->rep_V, nf, alphas, betas, steps, b_norm, nullptr, nullptr)) return rc;
    return rep_scatter(op, op->rep_V, V, ldv, k);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  const bool v_host = !is_device_ptr(V);
  double* V_dev = V;
  size_t ld_dev = ldv;
  if (v_host) {
    if (int rc = ensure_internal_basis(op, (size_t)op->n * k)) return rc;
    V_dev = op->V_int;
    ld_dev = op->n;
  }
  // Mat::zeros(n, k) (lanczos.rs:70): columns that are never reached stay zero
  CUDA_TRY(cudaMemset2DAsync(V_dev, ld_dev * sizeof(double), 0, (size_t)op->n * sizeof(double), k, op->stream));
  Decomp d;
  int rc = run_pass_one(op, b_dev, k, V_dev, ld_dev, cb, user, d);
  *b_norm = d.b_norm;
  if (rc) return rc;
  *steps = d.steps;
  std::copy(d.alphas.begin(), d.alphas.end(), alphas);
  if (!d.betas.empty()) std::copy(d.betas.begin(), d.betas.end(), betas);
  if (v_host) {
    if (int rc2 = basis_to_host(op, V_dev, V, ldv, k)) return rc2;
    CUDA_TRY(cudaStreamSynchronize(op->stream));
  }
  return TPL_OK;
}

// solvers::lanczos (src/solvers.rs:46-107): V_k stays in HBM, x = ||b|| * (V_k y') is one streaming GEMV.
int tpl_lanczos(tpl_op* op, const double* b, size_t k, tpl_ftk_solver f_tk, void* user, double* x) {
  tpl::clear_error();
  if (!op || !b || !x || !f_tk) return fail(TPL_ERR_PANIC, "null argument");
  if (k == 0) return fail(TPL_ERR_PANIC, "capacity overflow (k == 0; the reference panics in Vec::with_capacity(k - 1))");
  DeviceGuard g(op->device);
  if (replicated(op)) {
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    if (int rc = tpl_lanczos(op->inner, op->rep_in, k, f_tk, user, op->rep_out)) return rc;
    return rep_scatter(op, op->rep_out, x, op->n, 1);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  if (int rc = ensure_internal_basis(op, (size_t)op->n * k)) return rc;
  Decomp d;
  if (int rc = run_pass_one(op, b_dev, k, op->V_int, op->n, nullptr, nullptr, d)) return rc;
  double* x_dev = is_device_ptr(x) ? x : op->x_d;
  if (d.steps == 0) {  // solvers.rs:65-67
    CUDA_TRY(cudaMemsetAsync(x_dev, 0, sizeof(double) * op->n, op->stream));
    return finish_x(op, x_dev, x);
  }
  std::vector<double> y;
  if (int rc = call_ftk(f_tk, user, d, y)) return rc;
  if (int rc = gemv_vy(op, op->V_int, op->n, d.steps, y.data(), d.b_norm, x_dev)) return rc;
  return finish_x(op, x_dev, x);
}

// solvers::lanczos_two_pass (src/solvers.rs:133-175)
int tpl_lanczos_two_pass(tpl_op* op, const double* b, size_t k, tpl_ftk_solver f_tk, void* user, double* x) {
  tpl::clear_error();
  if (!op || !b || !x || !f_tk) return fail(TPL_ERR_PANIC, "null argument");
  DeviceGuard g(op->device);
  if (replicated(op)) {
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    if (int rc = tpl_lanczos_two_pass(op->inner, op->rep_in, k, f_tk, user, op->rep_out)) return rc;
    return rep_scatter(op, op->rep_out, x, op->n, 1);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  Decomp d;
  if (int rc = run_pass_one(op, b_dev, k, nullptr, 0, nullptr, nullptr, d)) return rc;
  double* x_dev = is_device_ptr(x) ? x : op->x_d;
  if (d.steps == 0) {  // solvers.rs:150-152
    CUDA_TRY(cudaMemsetAsync(x_dev, 0, sizeof(double) * op->n, op->stream));
    return finish_x(op, x_dev, x);
  }
  std::vector<double> y;
  if (int rc = call_ftk(f_tk, user, d, y)) return rc;
  for (double& yi : y) yi = yi * d.b_norm;  // solvers.rs:169
  if (int rc = run_pass_two(op, b_dev, d.alphas.data(), d.betas.data(), d.steps, d.b_norm, y.data(), y.size(), x_dev,
                            nullptr, 0))
    return rc;
  return finish_x(op, x_dev, x);
}

// the k-sweep of the reference's benches (src/bin/tradeoff.rs:262-290 re-solves for every k) from ONE basis
// generation / ONE pass 1 to max(ks).  Step j of the recurrence does not depend on k, so alphas[:k], betas[:k-1] of the long
// run are those of a k-step run, and every x_q below is bit-identical to the corresponding single solve.
static int sweep_prepare(tpl_op* op, const size_t* ks, size_t nk, size_t& kmax) {
  if (!ks || nk == 0) return fail(TPL_ERR_PANIC, "null argument");
  kmax = 0;
  for (size_t q = 0; q < nk; ++q) {
    if (ks[q] == 0) return fail(TPL_ERR_PANIC, "capacity overflow (k == 0; the reference panics in Vec::with_capacity(k - 1))");
    kmax = std::max(kmax, ks[q]);
  }
  (void)op;
  return TPL_OK;
}
// y'_q = f(T_{steps_q}) e1 with steps_q = min(ks[q], steps_taken), through the caller's closure
static int sweep_ftk(tpl_ftk_solver f_tk, void* user, const Decomp& d, size_t k, std::vector<double>& y, size_t& steps_q) {
  Decomp dq;
  steps_q = std::min(k, d.steps);
  dq.steps = steps_q;
  dq.b_norm = d.b_norm;
  dq.alphas.assign(d.alphas.begin(), d.alphas.begin() + steps_q);
  dq.betas.assign(d.betas.begin(), d.betas.begin() + (steps_q ? steps_q - 1 : 0));
  return call_ftk(f_tk, user, dq, y);
}

int tpl_lanczos_sweep(tpl_op* op, const double* b, const size_t* ks, size_t nk, tpl_ftk_solver f_tk, void* user, double* X,
                      size_t ldx) {
  tpl::clear_error();
  if (!op || !b || !X || !f_tk) return fail(TPL_ERR_PANIC, "null argument");
  if (ldx < op->n) return tpl::fail_parameter_mismatch("ldx", op->n, ldx);
  size_t kmax = 0;
  if (int rc = sweep_prepare(op, ks, nk, kmax)) return rc;
  DeviceGuard g(op->device);
  if (replicated(op)) {
    const size_t nf = op->rep_m + op->inc.p;
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    if (int rc = rep_basis(op, nf * nk)) return rc;
    if (int rc = tpl_lanczos_sweep(op->inner, op->rep_in, ks, nk, f_tk, user, op->rep_V, nf)) return rc;
    return rep_scatter(op, op->rep_V, X, ldx, nk);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  if (int rc = ensure_internal_basis(op, (size_t)op->n * kmax)) return rc;
  Decomp d;
  if (int rc = run_pass_one(op, b_dev, kmax, op->V_int, op->n, nullptr, nullptr, d)) return rc;
  const bool x_dev_out = is_device_ptr(X);
  double* Xd = X;
  size_t ldd = ldx;
  if (!x_dev_out) {
    if (int rc = ensure_sweep_buffer(op, (size_t)op->n * std::min<size_t>(nk, tpl::kSweepChunk))) return rc;
    Xd = op->sweep_d;
    ldd = op->n;
  }
  std::vector<double> y, Y;
  CUDA_TRY(cudaEventRecord(op->ev[4], op->stream));
  for (size_t q0 = 0; q0 < nk; q0 += tpl::kSweepChunk) {
    const size_t nq = std::min<size_t>(tpl::kSweepChunk, nk - q0);
    size_t kc = 0;
    Y.assign((std::max<size_t>(d.steps, 1)) * nq, 0.0);
    for (size_t q = 0; q < nq; ++q) {
      size_t steps_q = 0;
      if (d.steps) {
        if (int rc = sweep_ftk(f_tk, user, d, ks[q0 + q], y, steps_q)) return rc;
        for (size_t j = 0; j < steps_q; ++j) Y[j * nq + q] = y[j];
      }
      kc = std::max(kc, steps_q);
    }
    double* out = x_dev_out ? X + q0 * ldx : Xd;
    if (kc == 0) {  // solvers.rs:65-67: no step was taken
      CUDA_TRY(cudaMemset2DAsync(out, ldd * sizeof(double), 0, (size_t)op->n * sizeof(double), nq, op->stream));
    } else {
      if (int rc = ensure_sweep_coef(op, Y.size())) return rc;
      CUDA_TRY(cudaMemcpyAsync(op->sweep_y_d, Y.data(), Y.size() * sizeof(double), cudaMemcpyHostToDevice, op->stream));
      const int block = 256;
      const int grid = (int)std::min<size_t>((op->n + block - 1) / block, (size_t)op->G * 16);
      tpl::gemv_vy_sweep_kernel<<<grid, block, 0, op->stream>>>(op->V_int, op->n, op->n, (int)kc, op->sweep_y_d, (int)nq, d.b_norm,
                                                                out, ldd);
      CUDA_TRY(cudaGetLastError());
      op->launches += 1;
      CUDA_TRY(cudaStreamSynchronize(op->stream));  // Y (pageable host memory) is rewritten by the next chunk
    }
    if (!x_dev_out)
      CUDA_TRY(cudaMemcpy2DAsync(X + q0 * ldx, ldx * sizeof(double), Xd, ldd * sizeof(double), (size_t)op->n * sizeof(double), nq,
                                 cudaMemcpyDeviceToHost, op->stream));
  }
  CUDA_TRY(cudaEventRecord(op->ev[5], op->stream));
  op->timed[2] = true;
  CUDA_TRY(cudaStreamSynchronize(op->stream));
  return TPL_OK;
}

int tpl_lanczos_two_pass_sweep(tpl_op* op, const double* b, const size_t* ks, size_t nk, tpl_ftk_solver f_tk, void* user, double* X,
                               size_t ldx) {
  tpl::clear_error();
  if (!op || !b || !X || !f_tk) return fail(TPL_ERR_PANIC, "null argument");
  if (ldx < op->n) return tpl::fail_parameter_mismatch("ldx", op->n, ldx);
  size_t kmax = 0;
  if (int rc = sweep_prepare(op, ks, nk, kmax)) return rc;
  DeviceGuard g(op->device);
  if (replicated(op)) {
    const size_t nf = op->rep_m + op->inc.p;
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    if (int rc = rep_basis(op, nf * nk)) return rc;
    if (int rc = tpl_lanczos_two_pass_sweep(op->inner, op->rep_in, ks, nk, f_tk, user, op->rep_V, nf)) return rc;
    return rep_scatter(op, op->rep_V, X, ldx, nk);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  Decomp d;
  if (int rc = run_pass_one(op, b_dev, kmax, nullptr, 0, nullptr, nullptr, d)) return rc;
  const bool x_dev_out = is_device_ptr(X);
  std::vector<double> y;
  for (size_t q = 0; q < nk; ++q) {
    double* xq = X + q * ldx;
    double* x_dev = x_dev_out ? xq : op->x_d;
    if (d.steps == 0) {  // solvers.rs:150-152
      CUDA_TRY(cudaMemsetAsync(x_dev, 0, sizeof(double) * op->n, op->stream));
    } else {
      size_t steps_q = 0;
      if (int rc = sweep_ftk(f_tk, user, d, ks[q], y, steps_q)) return rc;
      for (double& yi : y) yi = yi * d.b_norm;  // solvers.rs:169
      if (int rc = run_pass_two(op, b_dev, d.alphas.data(), d.betas.data(), steps_q, d.b_norm, y.data(), y.size(), x_dev, nullptr, 0))
        return rc;
    }
    if (int rc = finish_x(op, x_dev, xq)) return rc;
  }
  return TPL_OK;
}

// k chosen from the residual estimates of one pass 1
int tpl_lanczos_two_pass_inv_adaptive(tpl_op* op, const double* b, size_t k_max, double rtol, double* x, size_t* k_used,
                                      double* res_est) {
  tpl::clear_error();
  if (!op || !b || !x) return fail(TPL_ERR_PANIC, "null argument");
  if (!(rtol >= 0.0)) return tpl::fail_input("rtol must be a non-negative number");
  DeviceGuard g(op->device);
  if (replicated(op)) {
    if (int rc = rep_gather(op, b, op->rep_in)) return rc;
    if (int rc = tpl_lanczos_two_pass_inv_adaptive(op->inner, op->rep_in, k_max, rtol, op->rep_out, k_used, res_est)) return rc;
    return rep_scatter(op, op->rep_out, x, op->n, 1);
  }
  const double* b_dev = nullptr;
  if (int rc = stage_b(op, b, &b_dev)) return rc;
  Decomp d;
  if (k_max == 0) return fail(TPL_ERR_PANIC, "capacity overflow (k == 0; the reference panics in Vec::with_capacity(k - 1))");
  // one step more than asked for: the decomposition holds beta_1 .. beta_{steps-1}, and the estimate of iterate j needs beta_j
  if (int rc = run_pass_one(op, b_dev, k_max + 1, nullptr, 0, nullptr, nullptr, d)) return rc;
  double* x_dev = is_device_ptr(x) ? x : op->x_d;
  if (k_used) *k_used = 0;
  if (res_est) *res_est = d.b_norm;
  if (d.steps == 0) {
    CUDA_TRY(cudaMemsetAsync(x_dev, 0, sizeof(double) * op->n, op->stream));
    return finish_x(op, x_dev, x);
  }
  const size_t cand = std::min(d.steps, k_max);  // iterates 1 .. cand are candidates
  std::vector<double> betas(d.betas), res(cand);
  if (betas.size() < cand) betas.push_back(0.0);  // pass 1 broke down at step `steps`: beta_steps <= 1000 eps, x_steps is exact
  if (int rc = tpl_ftk_inv_residuals(d.alphas.data(), cand, betas.data(), betas.size(), d.b_norm, res.data())) return rc;
  size_t k = 0;
  for (size_t j = 0; j < cand; ++j) {
    if (!(res[j] == res[j])) continue;  // NaN
    if (res[j] <= rtol * d.b_norm) {
      k = j + 1;
      break;
    }
    if (k == 0 || res[j] < res[k - 1]) k = j + 1;
  }
  if (k == 0) return tpl::fail_solver("no Lanczos iterate exists (T_j singular for every j)");
  std::vector<double> y(k);
  size_t y_len = 0;
  if (int rc = tpl_ftk_inv(d.alphas.data(), k, d.betas.data(), k - 1, y.data(), &y_len, nullptr)) return rc;
  for (double& yi : y) yi = yi * d.b_norm;
  if (int rc = run_pass_two(op, b_dev, d.alphas.data(), d.betas.data(), k, d.b_norm, y.data(), y.size(), x_dev, nullptr, 0))
    return rc;
  if (k_used) *k_used = k;
  if (res_est) *res_est = res[k - 1];
  return finish_x(op, x_dev, x);
}

// ---------------------------------------------------------------------------- multi-GPU
int tpl_comm_unique_id(uint8_t id_out[128]) {
  tpl::clear_error();
  if (!id_out) return fail(TPL_ERR_PANIC, "null argument");
  if (!nccl::api().ok) return fail(TPL_ERR_COMM, "NCCL error: %s", nccl::api().why.c_str());
  nccl::UniqueId id;
  NCCL_TRY(nccl::api().GetUniqueId(&id));
  std::memcpy(id_out, id.internal, 128);
  return TPL_OK;
}

int tpl_op_from_kkt_sharded(size_t m, size_t p, size_t arc_begin, size_t arc_end, const uint32_t* tail,
                            const uint32_t* head, const double* d, size_t d_len, int device, int rank, int world,
                            const uint8_t nccl_id[128], tpl_op** out) {
  tpl::clear_error();
  if (!out || !nccl_id || (m && (!tail || !head)) || (d_len && !d)) return fail(TPL_ERR_PANIC, "null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(TPL_ERR_COMM, "invalid rank %d of %d", rank, world);
  if (arc_begin > arc_end || arc_end > m)
    return tpl::fail_parameter_mismatch("arc_end", m, arc_end);
  if (d_len > m) return tpl::fail_parameter_mismatch("d", m, d_len);
  if (!nccl::api().ok) return fail(TPL_ERR_COMM, "NCCL error: %s", nccl::api().why.c_str());
  const size_t m_r = arc_end - arc_begin;
  const size_t d_local = d_len > arc_begin ? std::min(d_len - arc_begin, m_r) : 0;
  tpl_op* op = nullptr;
  if (int rc = tpl_op_from_kkt(m_r, p, tail ? tail + arc_begin : nullptr, head ? head + arc_begin : nullptr,
                               d_local ? d + arc_begin : nullptr, d_local, device, &op))
    return rc;
  auto bail = [&](int rc) {
    std::string keep = tpl::g_err;
    tpl_op_free(op);
    tpl::g_err = keep;
    return rc;
  };
  DeviceGuard g(op->device);
  if (!op->inc.stage_nodes)
    return bail(fail(TPL_ERR_COMM, "sharded mode needs the node segment (%zu doubles) to fit in shared memory", p));
  op->resident_ok = false;
  op->cells_ok = false;
  op->rank = rank;
  op->world = world;
  if (op->blocked_ok) op->tiled_ok = false;  // one family runs fused (and owns the exported exchange block)
  if ((op->tiled_ok || op->blocked_ok) && world <= tpl::kMaxRanks)
    if (int rc = setup_fabric(op, rank, world)) return bail(rc);
  if (int rc = dev_alloc(op, &op->red_d, 2 * (p + 1))) return bail(rc);
  if (int rc = dev_alloc(op, &op->red2_d, 1)) return bail(rc);
  if (cudaMemset(op->red_d, 0, sizeof(double) * 2 * (p + 1)) != cudaSuccess) return bail(fail(TPL_ERR_CUDA, "CUDA error: memset"));
  if (set_smem(tpl::shard_phase_a_kernel<false>, op->smem_bytes) || set_smem(tpl::shard_phase_a_kernel<true>, op->smem_bytes) ||
      set_smem(tpl::shard_pass2_kernel<false>, op->smem_bytes) || set_smem(tpl::shard_pass2_kernel<true>, op->smem_bytes))
    return bail(TPL_ERR_CUDA);
  nccl::UniqueId id;
  std::memcpy(id.internal, nccl_id, 128);
  nccl::Result r = nccl::api().CommInitRank(&op->comm, world, id, rank);
  if (r != 0) return bail(fail(TPL_ERR_COMM, "NCCL error: %s (ncclCommInitRank)", nccl::api().GetErrorString(r)));
  // Replicated execution when the WHOLE operator fits the on-chip cell kernels (every rank takes the same decision: it
  // depends on the global arrays only).  TPL_NO_REPLICATE=1 or any mode other than 0 keeps the arc-partitioned paths.
  if (world > 1 && m <= (size_t)op->G * tpl::kCellArcs && !std::getenv("TPL_NO_REPLICATE")) {
    tpl_op* in = nullptr;
    if (tpl_op_from_kkt(m, p, tail, head, d, d_len, op->device, &in) == TPL_OK) {
      if (in->cells_ok && tpl_op_set_stream(in, op->stream) == TPL_OK) {
        op->inner = in;
        op->rep_lo = arc_begin;
        op->rep_hi = arc_end;
        op->rep_m = m;
        if (dev_alloc(op, &op->rep_in, m + p) || dev_alloc(op, &op->rep_out, m + p)) return bail(TPL_ERR_CUDA);
      } else {
        tpl_op_free(in);
      }
    }
    tpl::clear_error();
  }
  *out = op;
  return TPL_OK;
}

int tpl_op_fabric_export(tpl_op* op, uint8_t handle[64]) {
  tpl::clear_error();
  if (!op || !handle) return fail(TPL_ERR_PANIC, "null argument");
  if (op->inner)
    return fail(TPL_ERR_COMM, "the operator runs replicated (the whole problem fits one GPU's on-chip kernels): no exchange block");
  if (!op->comm || !(op->tiled_ok || op->blocked_ok) || !op->fab_block || op->world > tpl::kMaxRanks)
    return fail(TPL_ERR_COMM, "the operator has no exchange block (not sharded, or neither the blocked nor the tiled kernels fit)");
  DeviceGuard g(op->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, op->fab_block));
  std::memcpy(handle, &h, 64);
  return TPL_OK;
}

int tpl_op_fabric_import(tpl_op* op, const uint8_t* handles, int count) {
  tpl::clear_error();
  if (!op || !handles) return fail(TPL_ERR_PANIC, "null argument");
  if (!op->comm || !op->fab_block) return fail(TPL_ERR_COMM, "the operator has no exchange block");
  if (count != op->world) return tpl::fail_parameter_mismatch("handles", (size_t)op->world, (size_t)count);
  if (op->fab_connected) return TPL_OK;
  DeviceGuard g(op->device);
  tpl::Fabric& f = active_fabric(op);
  size_t part, nodes, slot_off;
  fabric_layout(op, op->blocked_ok, f, part, nodes, slot_off);
  for (int r = 0; r < op->world; ++r) {
    if (r == op->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * 64, 64);
    void* peer = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&peer, h, cudaIpcMemLazyEnablePeerAccess));
    op->fab_peers.push_back(peer);
    f.partials[r] = reinterpret_cast<double*>(peer);
    f.nodebuf[r] = f.partials[r] + part;
    f.slots[r] = reinterpret_cast<uint4*>(static_cast<char*>(peer) + slot_off);
  }
  op->fab_connected = true;
  op->fab_epoch = 0;
  return TPL_OK;
}

int tpl_op_shard_info(const tpl_op* op, int* rank, int* world, size_t* local_arcs, size_t* nodes) {
  if (!op) return fail(TPL_ERR_PANIC, "null argument");
  if (rank) *rank = op->rank;
  if (world) *world = op->world;
  if (local_arcs) *local_arcs = op->format == 2 ? op->inc.m : 0;
  if (nodes) *nodes = op->format == 2 ? op->inc.p : 0;
  return TPL_OK;
}

}  // extern "C"
