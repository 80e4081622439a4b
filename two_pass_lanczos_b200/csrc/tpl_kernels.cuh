// tpl_kernels.cuh -- sm_100a device code of the two-pass Lanczos engine.
//
// Design (DESIGN.md has the long version):
//   * ONE persistent cooperative kernel per pass.  The grid is one CTA per SM; Lanczos steps are a loop
//     inside the kernel and the two dependent reductions of a step (alpha, beta) are grid-wide
//     flag barriers that carry the reduction payload (no atomics, fixed summation order).
//   * The operator is either the KKT network-incidence form (arc rows: thread per arc, node segment
//     staged in shared memory; node rows: warp per fixed-length segment of the node->arc list) or a
//     generic CSR (short rows: thread per row; long rows: the same segment machinery).
//   * Two execution shapes share every per-element expression:
//       - "resident": the CTA's slice of the operator (d, tail, head, node->arc lists) and of the
//         Lanczos vectors lives in shared memory for the whole pass; per step only the node segment
//         and the node-row gathers come from L2 and only the new vector is published to global memory.
//         Used whenever the slice fits in the 227 KB of an SM (up to ~800k arcs on 148 SMs).
//       - "streaming": every sweep streams the vectors and the operator from HBM/L2; any size.
//   * Every per-element expression of the recurrence is written once with explicit round-to-nearest
//     mul/sub so that pass 1, the one-pass variant and pass 2 produce bit-identical basis vectors
//     (reference invariant `basis_drift == 0`, results/orthogonality_*.csv) and follow the reference's
//     two-rounding `sub(w, mul(c, v))` (src/algorithms/mod.rs:183-198).  Reductions group the rows the
//     same way in both shapes, so alpha/beta are bit-identical between them too.
//   * Vectors written inside a kernel are read by other CTAs only through L2 (ld.global.cg / st.global.cg).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpl {

constexpr int kBlock = 512;  // threads per CTA (one CTA per SM, up to 128 registers per thread)
constexpr int kWarps = kBlock / 32;
constexpr uint32_t kSignBit = 0x80000000u;
constexpr uint32_t kSpinLimit = 1u << 24;  // grid-barrier watchdog: trap instead of hanging the GPU
constexpr int kSegChunk = 8;               // entries a lane gathers per batch (segment length 256 = 8 * 32)

enum : int { ST_RUNNING = 0, ST_BREAKDOWN = 1, ST_ZERO_B = 2 };

// Rows whose product is summed by warp-per-segment: KKT node rows, long CSR rows.
struct LongRows {
  uint32_t nlong;           // number of long rows
  uint32_t max_segs;        // max segments / entries / rows owned by one CTA (shared-memory sizing)
  uint32_t max_ents;
  uint32_t max_rows;
  const uint32_t* row;      // [nlong]   global row id
  const uint32_t* seg_ptr;  // [nlong+1] first segment of each long row
  const uint32_t* ent_ptr;  // [nseg+1]  entry range of each segment
  const uint32_t* ent_idx;  // [nent]    column (incidence: bit 31 set = coefficient -1)
  const double* ent_val;    // [nent]    CSR only
  const uint32_t* cta_ptr;  // [G+1]     long rows owned by each CTA
  double* seg_scratch;      // [nseg] or nullptr: where the segment sums go when a CTA's share does not fit in shared memory
                            // (operators with millions of rows in this format: every node row of a sparse graph)
};

struct IncidenceOp {  // A = [[D, E^T], [E, 0]], arc j: +1 at tail, -1 at head
  uint32_t m, p;
  const double* d;       // [m] (zero beyond the loader's d_len)
  const uint32_t* tail;  // [m]
  const uint32_t* head;  // [m]
  LongRows lr;
  int stage_nodes;       // node segment fits in shared memory
};

struct CsrOp {
  uint32_t n;
  uint32_t long_thresh;     // rows with more entries are handled by the segment path
  const uint32_t* row_ptr;  // [n+1]
  const uint32_t* col;      // [nnz]
  const double* val;        // [nnz]
  LongRows lr;
};

struct State {  // persists in HBM between launches of the same handle
  double s_cur, s_prev, beta_prev, b_norm;
  unsigned int epoch;
  int rot, steps, status;
};

struct Trace {  // optional per-CTA phase timestamps (diagnostics; buf == nullptr in production)
  unsigned long long* buf;  // [G][max_steps][kTraceMarks] SM clock (globaltimer in the last slot)
  int max_steps;
};
constexpr int kTraceMarks = 64;  // 0..31 thread 0 of the CTA, 32..47 / 48..62 lane 0 of warp w (two stamps), 63 globaltimer

struct GridSync {
  uint4* slots;  // [2][G][kSlotAtoms] one 128-byte line of {payload lo, epoch, payload hi, epoch} atoms per CTA, double-buffered by epoch parity
  Trace trace;
  int trace_step;  // step the marks taken inside grid_sync belong to
  int trace_base;  // first mark index used by this grid_sync call
};

__device__ __forceinline__ void trace_mark(const Trace& t, int step, int m) {
  if (t.buf != nullptr && threadIdx.x == 0 && step >= 0 && step < t.max_steps) {
    unsigned long long* p = t.buf + ((size_t)blockIdx.x * t.max_steps + step) * kTraceMarks;
    p[m] = clock64();
    if (m == 0) {
      unsigned long long g;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      p[kTraceMarks - 1] = g;
    }
  }
}

// stamp of warp w (lane 0) in slot base + w
__device__ __forceinline__ void trace_mark_warp(const Trace& t, int step, int base) {
  const int w = threadIdx.x >> 5;
  if (t.buf != nullptr && (threadIdx.x & 31) == 0 && step >= 0 && step < t.max_steps && base + w < kTraceMarks - 1)
    t.buf[((size_t)blockIdx.x * t.max_steps + step) * kTraceMarks + base + w] = clock64();
}

__device__ __forceinline__ void trace_value(const Trace& t, int step, int m, unsigned long long v) {
  if (t.buf != nullptr && threadIdx.x == 0 && step >= 0 && step < t.max_steps)
    t.buf[((size_t)blockIdx.x * t.max_steps + step) * kTraceMarks + m] = v;
}

struct Pass1Args {
  double* buf[3];
  const double* b;
  double* alphas;
  double* betas;
  double* V;  // optional n x k basis (one-pass), column-major
  size_t ldv;
  uint32_t n;
  int j_begin, j_end;
  State* st;
  GridSync gs;
  double tol;
};

struct Pass2Args {
  double* buf[3];
  const double* b;
  const double* alphas;
  const double* betas;
  const double* y;
  double* x;
  double* V;  // optional regenerated basis
  size_t ldv;
  uint32_t n;
  int steps;
  double b_norm;
  State* st;
  GridSync gs;
};

// ----------------------------------------------------------------------------- primitives
__device__ __forceinline__ double rec_sub(double t, double c, double u) {
  return __dsub_rn(t, __dmul_rn(c, u));  // t - c*u, two roundings, never contracted
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_v4(uint4* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

constexpr unsigned int kMaxGridCtas = 160;  // CTAs of a persistent grid (one per SM)
constexpr unsigned int kSlotAtoms = 8;      // 16-byte atoms of a CTA's barrier line
struct CtaShared {
  double warp_part[kWarps];
  double result;
  double gather[kMaxGridCtas];  // the payloads of all CTAs (grid_sync)
};

// Grid-wide sum (REDUCE) and/or barrier.  Every CTA owns one 128-byte LINE per epoch parity: eight 16-byte atoms
// {payload lo, epoch, payload hi, epoch} -- each 8-byte half carries its own flag, so a reader never combines halves of
// different epochs (the layout of NCCL's LL protocol) -- all holding the CTA's partial sum.  After a CTA-wide __syncthreads
// eight lanes publish the line in ONE store instruction; thread t of every CTA then polls line t (atom blockIdx.x mod 8, so
// that the G readers of a line spread over its four sectors), the payloads meet in shared memory and every warp adds them in
// the same fixed order: the result is identical in every CTA and from run to run.  Double-buffered by epoch parity (no CTA
// can be two episodes ahead of another).  No atomics.
// Measured (scripts/ubench/sync_bench.cu, 148 CTAs): 16-byte slots packed 8 to a line, polled by one warp -- the first
// form of this barrier -- 5 200 cycles; whole lines polled by a thread each ~2 300: partially written 32-byte sectors and
// 148 pollers on 19 hot lines were the cost, not latency.
//   FENCED = true : full barrier semantics for plain global data written before the call (release fence before
//                   the line is published, acquire fence after all lines were seen) -- streaming kernels.
//   FENCED = false: pure all-reduce of self-validating lines; nothing else is ordered -- dataflow kernels, whose
//                   other exchanged data carries its own tags.
template <bool REDUCE, bool FENCED = true>
__device__ __forceinline__ double grid_sync(double v, const GridSync& gs, unsigned int& epoch, CtaShared& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;  // <= kMaxGridCtas (checked at launch)
  epoch += 1;
  if (REDUCE) {
    v = warp_sum(v);
    if (lane == 0) sh.warp_part[warp] = v;
  }
  __syncthreads();
  trace_mark(gs.trace, gs.trace_step, gs.trace_base + 0);
  uint4* lines = gs.slots + (size_t)(epoch & 1u) * G * kSlotAtoms;
  if (warp == 0) {
    double t = 0.0;
    if (REDUCE) {
      t = lane < kWarps ? sh.warp_part[lane] : 0.0;
      t = warp_sum(t);
    }
    if (lane < (int)kSlotAtoms) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
      if (FENCED) fence_acq_rel_gpu();  // release: orders the CTA's earlier writes (cumulative over the barrier above)
      st_relaxed_gpu_v4(lines + (size_t)blockIdx.x * kSlotAtoms + lane, make_uint4((unsigned)bits, epoch, (unsigned)(bits >> 32), epoch));
    }
  }
  trace_mark(gs.trace, gs.trace_step, gs.trace_base + 1);
  for (unsigned int i = threadIdx.x; i < G; i += kBlock) {
    const uint4* src = lines + (size_t)i * kSlotAtoms + (blockIdx.x & (kSlotAtoms - 1));
    uint4 f;
    unsigned int spins = 0;
    for (;;) {
      f = ld_relaxed_gpu_v4(src);
      if ((f.y == epoch) & (f.w == epoch)) break;
      if (++spins > kSpinLimit) __trap();
    }
    if (FENCED) fence_acq_rel_gpu();  // acquire: the poller has seen the line; the barrier below passes it on to the CTA
    if (REDUCE) sh.gather[i] = __longlong_as_double((long long)(((unsigned long long)f.z << 32) | f.x));
  }
  __syncthreads();
  trace_mark(gs.trace, gs.trace_step, gs.trace_base + 2);
  double s = 0.0;
  if (REDUCE) {  // every warp adds the same numbers in the same order
    for (unsigned int i = lane; i < G; i += 32) s += sh.gather[i];
    s = warp_sum(s);
  }
  trace_mark(gs.trace, gs.trace_step, gs.trace_base + 3);
  return REDUCE ? s : 0.0;
}

// ----------------------------------------------------------------------------- operator products
// x is always addressed as X[i]*s (lazy normalisation: pass 1 keeps the un-normalised w and the
// reciprocal norm; X[i]*s is the single rounding the reference performs when it scales w in place,
// src/algorithms/mod.rs:312-315).  Pass 2 and apply() use s = 1.

// (A x)_j for arc row j in the reference's CSC accumulation order: D_jj x_j first, then the incident
// node columns in ascending node index (columns m+tail, m+head).  A self-loop's merged E entry is an
// explicit 0 (data_loader.rs:118-133) and contributes nothing.
__device__ __forceinline__ double arc_row(double dj, double xj, uint32_t t, uint32_t h, double xt, double xh) {
  double acc = __dmul_rn(dj, xj);
  if (t == h) {
  } else if (t < h) {
    acc = __dadd_rn(acc, xt);
    acc = __dsub_rn(acc, xh);
  } else {
    acc = __dsub_rn(acc, xh);
    acc = __dadd_rn(acc, xt);
  }
  return acc;
}
// one entry of a node row: +x_j for an out-arc, -x_j for an in-arc
__device__ __forceinline__ double node_entry(uint32_t idx, double xval, double s, double acc) {
  const double x = __dmul_rn(xval, s);
  return (idx & kSignBit) ? __dsub_rn(acc, x) : __dadd_rn(acc, x);
}

struct IncidenceDev {
  const IncidenceOp& op;
  const double* sm_node;  // staged, already scaled node segment (or nullptr)
  __device__ __forceinline__ uint32_t num_short() const { return op.m; }
  __device__ __forceinline__ bool is_short(uint32_t) const { return true; }
  __device__ __forceinline__ double node(uint32_t u, const double* X, double s) const {
    return sm_node ? sm_node[u] : __dmul_rn(__ldcg(X + op.m + u), s);
  }
  __device__ __forceinline__ double short_row(uint32_t j, double xj, const double* X, double s) const {
    const uint32_t t = __ldg(op.tail + j), h = __ldg(op.head + j);
    return arc_row(__ldg(op.d + j), xj, t, h, node(t, X, s), node(h, X, s));
  }
  __device__ __forceinline__ double entry(const LongRows& lr, uint32_t e, const double* X, double s, double acc) const {
    const uint32_t idx = __ldg(lr.ent_idx + e);
    return node_entry(idx, __ldcg(X + (idx & ~kSignBit)), s, acc);
  }
};

struct CsrDev {
  const CsrOp& op;
  __device__ __forceinline__ uint32_t num_short() const { return op.n; }
  __device__ __forceinline__ bool is_short(uint32_t i) const {
    return __ldg(op.row_ptr + i + 1) - __ldg(op.row_ptr + i) <= op.long_thresh;
  }
  __device__ __forceinline__ double short_row(uint32_t i, double, const double* X, double s) const {
    const uint32_t p0 = __ldg(op.row_ptr + i), p1 = __ldg(op.row_ptr + i + 1);
    double acc = 0.0;
    for (uint32_t p = p0; p < p1; ++p) {
      const double x = __dmul_rn(__ldcg(X + __ldg(op.col + p)), s);
      acc = __dadd_rn(acc, __dmul_rn(__ldg(op.val + p), x));
    }
    return acc;
  }
  __device__ __forceinline__ double entry(const LongRows& lr, uint32_t e, const double* X, double s, double acc) const {
    const double x = __dmul_rn(__ldcg(X + __ldg(lr.ent_idx + e)), s);
    return __dadd_rn(acc, __dmul_rn(__ldg(lr.ent_val + e), x));
  }
};

__device__ __forceinline__ const LongRows& long_rows(const IncidenceOp& op) { return op.lr; }
__device__ __forceinline__ const LongRows& long_rows(const CsrOp& op) { return op.lr; }

// Stage the (scaled) node segment of X into shared memory.  Caller syncs.
__device__ __forceinline__ const double* stage_nodes(const IncidenceOp& op, const double* X, double s, double* sm) {
  if (!op.stage_nodes) return nullptr;
  for (uint32_t u = threadIdx.x; u < op.p; u += kBlock) sm[u] = __dmul_rn(__ldcg(X + op.m + u), s);
  return sm;
}
__device__ __forceinline__ const double* stage_nodes(const CsrOp&, const double*, double, double*) { return nullptr; }
__device__ __forceinline__ IncidenceDev make_dev(const IncidenceOp& op, const double* sm_node) { return IncidenceDev{op, sm_node}; }
__device__ __forceinline__ CsrDev make_dev(const CsrOp& op, const double*) { return CsrDev{op}; }
__device__ __forceinline__ uint32_t node_smem_doubles(const IncidenceOp& op) { return op.stage_nodes ? op.p : 0; }
__device__ __forceinline__ uint32_t node_smem_doubles(const CsrOp&) { return 0; }

// contiguous block of `total` items owned by this CTA
__device__ __forceinline__ void cta_chunk(uint32_t total, uint32_t& lo, uint32_t& hi) {
  const uint32_t chunk = (total + gridDim.x - 1) / gridDim.x;
  const uint64_t a = (uint64_t)chunk * blockIdx.x;
  lo = a < total ? (uint32_t)a : total;
  hi = (a + chunk) < total ? (uint32_t)(a + chunk) : total;
}

// Segment sums live in shared memory, or -- written and read by the same CTA around a CTA barrier, through L2 -- in
// lr.seg_scratch when the CTA's share is too large for it.
__device__ __forceinline__ void seg_put(const LongRows& lr, double* sm_seg, uint32_t sg, uint32_t s0, double v) {
  if (lr.seg_scratch)
    __stcg(lr.seg_scratch + sg, v);
  else
    sm_seg[sg - s0] = v;
}
__device__ __forceinline__ double seg_get(const LongRows& lr, const double* sm_seg, uint32_t sg, uint32_t s0) {
  return lr.seg_scratch ? __ldcg(lr.seg_scratch + sg) : sm_seg[sg - s0];
}
// Sums the segments of the long rows owned by this CTA into sm_seg (one warp per segment, lanes
// stride the entries, xor-shuffle tree).  The order depends only on the operator, never on the grid.
template <class DEV>
__device__ __forceinline__ void long_row_segments(const DEV& dev, const LongRows& lr, const double* X, double s,
                                                  double* sm_seg, uint32_t& r0, uint32_t& r1, uint32_t& s0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  r0 = __ldg(lr.cta_ptr + blockIdx.x);
  r1 = __ldg(lr.cta_ptr + blockIdx.x + 1);
  s0 = __ldg(lr.seg_ptr + r0);
  const uint32_t s1 = __ldg(lr.seg_ptr + r1);
  for (uint32_t sg = s0 + warp; sg < s1; sg += kWarps) {
    const uint32_t e0 = __ldg(lr.ent_ptr + sg), e1 = __ldg(lr.ent_ptr + sg + 1);
    double acc = 0.0;
#pragma unroll 4
    for (uint32_t e = e0 + lane; e < e1; e += 32) acc = dev.entry(lr, e, X, s, acc);
    acc = warp_sum(acc);
    if (lane == 0) seg_put(lr, sm_seg, sg, s0, acc);
  }
}
__device__ __forceinline__ double long_row_total(const LongRows& lr, uint32_t q, const double* sm_seg, uint32_t s0) {
  const uint32_t a = __ldg(lr.seg_ptr + q), b = __ldg(lr.seg_ptr + q + 1);
  double t = 0.0;
  for (uint32_t sg = a; sg < b; ++sg) t = __dadd_rn(t, seg_get(lr, sm_seg, sg, s0));
  return t;
}

// =============================================================================================
// STREAMING kernels (any size, both operator kinds)
// =============================================================================================
// Replaces lanczos_pass_one (src/algorithms/lanczos_two_pass.rs:65-110) and, with WITH_V, the basis
// generation of lanczos_standard (src/algorithms/lanczos.rs:55-156).  Per step (mod.rs:167-212, 292-340):
//   phase A  w~ = A v_j - beta_{j-1} v_{j-1},  alpha_j = <v_j, w~>      (one sweep + grid reduction)
//   phase B  w  = w~ - alpha_j v_j,            beta_j  = ||w||          (one sweep + grid reduction)
// v_j is held as (W_cur, s_cur) with v = W_cur * s_cur.  Every CTA owns a contiguous chunk of the short
// rows plus the long rows cta_ptr assigns to it, in every phase.
template <class OP, bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_kernel(const OP op, const Pass1Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  double* sm_node = smem;
  double* sm_seg = smem + node_smem_doubles(op);

  const State st0 = *a.st;
  unsigned int epoch = st0.epoch;  // continues across the launches of a step-per-launch pass (slots keep their last tags)
  int rot = st0.rot, steps = st0.steps, status = st0.status;
  double sc = st0.s_cur, sp = st0.s_prev, bp = st0.beta_prev, bnorm = st0.b_norm;
  const LongRows& lr = long_rows(op);
  const auto dev0 = make_dev(op, nullptr);
  uint32_t slo, shi;
  cta_chunk(dev0.num_short(), slo, shi);
  const uint32_t r0 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x) : 0;
  const uint32_t r1 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x + 1) : 0;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };

  if (a.j_begin == 0) {
    // K0: ||b||, W_cur = b, W_prev = 0   (lanczos_two_pass.rs:74, mod.rs:261-289)
    double* Wp = pick(rot);
    double* Wc = pick((rot + 1) % 3);
    double acc = 0.0;
#pragma unroll 4
    for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
      if (!dev0.is_short(i)) continue;
      const double bi = __ldg(a.b + i);
      __stcg(Wc + i, bi);
      __stcg(Wp + i, 0.0);
      acc = fma(bi, bi, acc);
    }
    for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
      const uint32_t i = __ldg(lr.row + q);
      const double bi = __ldg(a.b + i);
      __stcg(Wc + i, bi);
      __stcg(Wp + i, 0.0);
      acc = fma(bi, bi, acc);
    }
    bnorm = sqrt(grid_sync<true>(acc, a.gs, epoch, sh));
    steps = 0;
    if (bnorm <= a.tol) {
      status = ST_ZERO_B;
    } else {
      status = ST_RUNNING;
      sc = 1.0 / bnorm;
      sp = 1.0;
      bp = 0.0;
    }
  }

  if (status == ST_RUNNING) {
    for (int j = a.j_begin; j < a.j_end; ++j) {
      const double* Wp = pick(rot);
      const double* Wc = pick((rot + 1) % 3);
      double* Wn = pick((rot + 2) % 3);
      double* Vcol = WITH_V ? a.V + (size_t)j * a.ldv : nullptr;

      // ---------------- phase A
      const double* nodes = stage_nodes(op, Wc, sc, sm_node);
      __syncthreads();
      const auto dev = make_dev(op, nodes);
      double acc = 0.0;
#pragma unroll 2
      for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
        if (!dev.is_short(i)) continue;
        const double v = __dmul_rn(__ldcg(Wc + i), sc);
        const double vp = __dmul_rn(__ldcg(Wp + i), sp);
        const double wt = rec_sub(dev.short_row(i, v, Wc, sc), bp, vp);
        acc = fma(v, wt, acc);
        __stcg(Wn + i, wt);
        if (WITH_V) __stcs(Vcol + i, v);
      }
      if (lr.nlong) {
        uint32_t q0, q1, s0;
        long_row_segments(dev, lr, Wc, sc, sm_seg, q0, q1, s0);
        __syncthreads();
        for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
          const uint32_t i = __ldg(lr.row + q);
          const double t = long_row_total(lr, q, sm_seg, s0);
          const double v = __dmul_rn(__ldcg(Wc + i), sc);
          const double vp = __dmul_rn(__ldcg(Wp + i), sp);
          const double wt = rec_sub(t, bp, vp);
          acc = fma(v, wt, acc);
          __stcg(Wn + i, wt);
          if (WITH_V) __stcs(Vcol + i, v);
        }
      }
      const double alpha = grid_sync<true, false>(acc, a.gs, epoch, sh);  // all-reduce only: phase B reads this CTA's own rows

      // ---------------- phase B (same row ownership: every thread re-reads the w~ it wrote)
      acc = 0.0;
#pragma unroll 4
      for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
        if (!dev.is_short(i)) continue;
        const double v = __dmul_rn(__ldcg(Wc + i), sc);
        const double w = rec_sub(__ldcg(Wn + i), alpha, v);
        __stcg(Wn + i, w);
        acc = fma(w, w, acc);
      }
      for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
        const uint32_t i = __ldg(lr.row + q);
        const double v = __dmul_rn(__ldcg(Wc + i), sc);
        const double w = rec_sub(__ldcg(Wn + i), alpha, v);
        __stcg(Wn + i, w);
        acc = fma(w, w, acc);
      }
      const double beta = sqrt(grid_sync<true>(acc, a.gs, epoch, sh));

      if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.alphas[j] = alpha;
        a.betas[j] = beta;
      }
      steps = j + 1;
      if (beta <= a.tol) {  // breakdown: stop, buffers are not rotated (mod.rs:331-338)
        status = ST_BREAKDOWN;
        break;
      }
      sp = sc;
      sc = 1.0 / beta;  // recip, then multiply (mod.rs:312)
      bp = beta;
      rot = (rot + 1) % 3;
    }
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    State st;
    st.s_cur = sc;
    st.s_prev = sp;
    st.beta_prev = bp;
    st.b_norm = bnorm;
    st.epoch = epoch;
    st.rot = rot;
    st.steps = steps;
    st.status = status;
    *a.st = st;
  }
}

// Replaces lanczos_pass_two_impl (src/algorithms/lanczos_two_pass.rs:206-312): regenerates v_{j+1}
// with the stored alpha_j, beta_{j-1}, beta_j and accumulates x += y_{j+1} v_{j+1} in the same sweep.
// One grid barrier per step.
template <class OP, bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_kernel(const OP op, const Pass2Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  double* sm_node = smem;
  double* sm_seg = smem + node_smem_doubles(op);

  unsigned int epoch = a.st->epoch;
  const LongRows& lr = long_rows(op);
  const auto dev0 = make_dev(op, nullptr);
  uint32_t slo, shi;
  cta_chunk(dev0.num_short(), slo, shi);
  const uint32_t r0 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x) : 0;
  const uint32_t r1 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x + 1) : 0;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  int rot = 0;
  {
    // v_1 = b * (1/||b||), x = y_0 v_1   (lanczos_two_pass.rs:247-258)
    const double inv = 1.0 / a.b_norm;
    const double y0 = __ldg(a.y);
    double* Vp = buf0;
    double* Vc = buf1;
    auto init = [&](uint32_t i) {
      const double v = __dmul_rn(__ldg(a.b + i), inv);
      __stcg(Vc + i, v);
      __stcg(Vp + i, 0.0);
      __stcg(a.x + i, __dmul_rn(v, y0));
      if (WITH_V) __stcs(a.V + i, v);
    };
#pragma unroll 4
    for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock)
      if (dev0.is_short(i)) init(i);
    for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) init(__ldg(lr.row + q));
    grid_sync<false>(0.0, a.gs, epoch, sh);
  }
  for (int j = 0; j + 1 < a.steps; ++j) {
    const double* Vp = pick(rot);
    const double* Vc = pick((rot + 1) % 3);
    double* Vn = pick((rot + 2) % 3);
    double* Vcol = WITH_V ? a.V + (size_t)(j + 1) * a.ldv : nullptr;
    const double alpha = __ldg(a.alphas + j);
    const double beta = __ldg(a.betas + j);
    const double bp = j == 0 ? 0.0 : __ldg(a.betas + j - 1);
    const double s = 1.0 / beta;
    const double yj = __ldg(a.y + j + 1);

    const double* nodes = stage_nodes(op, Vc, 1.0, sm_node);
    __syncthreads();
    const auto dev = make_dev(op, nodes);
    auto finish = [&](uint32_t i, double t, double v) {
      const double w = rec_sub(rec_sub(t, bp, __ldcg(Vp + i)), alpha, v);
      const double vn = __dmul_rn(w, s);
      __stcg(Vn + i, vn);
      __stcg(a.x + i, __dadd_rn(__ldcg(a.x + i), __dmul_rn(yj, vn)));
      if (WITH_V) __stcs(Vcol + i, vn);
    };
#pragma unroll 2
    for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
      if (!dev.is_short(i)) continue;
      const double v = __ldcg(Vc + i);
      finish(i, dev.short_row(i, v, Vc, 1.0), v);
    }
    if (lr.nlong) {
      uint32_t q0, q1, s0;
      long_row_segments(dev, lr, Vc, 1.0, sm_seg, q0, q1, s0);
      __syncthreads();
      for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
        const uint32_t i = __ldg(lr.row + q);
        finish(i, long_row_total(lr, q, sm_seg, s0), __ldcg(Vc + i));
      }
    }
    grid_sync<false>(0.0, a.gs, epoch, sh);
    rot = (rot + 1) % 3;
  }
}

// =============================================================================================
// RESIDENT kernels (KKT incidence operator whose per-CTA slice fits in shared memory)
// =============================================================================================
// Every CTA owns a contiguous chunk of A = ceil(m/G) arcs and a block of R = ceil(p/G) node rows.  The chunk's
// d / tail / head, the arc part of the two rotating Lanczos vectors (and of x in pass 2) live in shared memory
// for the whole pass.  Node rows are NOT gathered from L2 (scattered 8-byte gathers cost ~2 L1TEX cycles per
// element and dominated the first version of these kernels): instead
//   1. right after a CTA has produced its new arc values w it adds them, from shared memory, into the partial
//      node sums P_c[u] = sum over ITS arcs of (+w_j if tail_j == u, -w_j if head_j == u), walking a per-CTA
//      node -> local-arc list in a fixed order, and publishes the p partials as G chunks of R doubles;
//   2. after the grid barrier the owner of node block g reads the contiguous [G][R] block of partials, adds the
//      G partials of every node in a fixed tree order, T_u, and forms (E v)_u = fl(s * T_u) with the lazy scale
//      s = 1/beta of the vector the partials were taken from.
// Both passes and the one-pass variant use the same lists, order and scale, so the regenerated basis is still
// bit-identical to the stored one.  Per step a CTA moves ~p + G*R doubles in and ~G*R + R doubles out, all
// coalesced.
struct ResidentOp {
  uint32_t R;                   // node rows per owner block
  uint32_t max_long;            // max number of long local lists of one CTA
  const uint32_t* nl_ptr;       // [G][p+1] offsets into nl_ent (per-CTA node -> local-arc lists)
  const uint16_t* nl_ent;       // local arc index | 0x8000 when the arc enters the node (coefficient -1)
  const uint32_t* nl_long_ptr;  // [G+1] range of long lists (warp-per-list) of each CTA
  const uint32_t* nl_long;      // node ids
  double* partials;             // [2][G][G][R] workspace
  double* nodebuf;              // [2][p]      published node part of the newest vector (un-normalised)
};
constexpr uint32_t kLongList = 48;  // local lists longer than this are summed by a warp

struct ResidentSmem {
  double* node;  // [p]   scaled node segment of the current vector
  double* w0;    // [A]   arc part of the two rotating (un-normalised) Lanczos vectors
  double* w1;    // [A]
  double* x;     // [A]   pass 2 only
  double* d;     // [A]
  double* blk;   // [G*R] partial node sums addressed to this CTA's node block
  double* n0;    // [R]   node part of the two rotating vectors (owned block)
  double* n1;    // [R]
  double* nx;    // [R]   pass 2 only
  uint32_t* tail;   // [A]
  uint32_t* head;   // [A]
  uint32_t* lptr;   // [p+1] local list offsets (relative)
  uint32_t* llong;  // [max_long]
  uint16_t* lent;   // [2A]
};

__host__ __device__ inline size_t resident_smem_bytes(uint32_t p, uint32_t A, uint32_t G, uint32_t R, uint32_t max_long,
                                                      bool pass2) {
  size_t dbl = (size_t)p + 3 * (size_t)A + (size_t)G * R + 2 * (size_t)R + (pass2 ? (size_t)A + R : 0);
  size_t u32 = 2 * (size_t)A + (p + 1) + max_long;
  size_t u16 = 2 * (size_t)A;
  return dbl * 8 + u32 * 4 + u16 * 2 + 16;
}

template <bool PASS2>
__device__ __forceinline__ ResidentSmem carve_resident(double* base, uint32_t p, uint32_t A, uint32_t G, uint32_t R,
                                                       uint32_t max_long) {
  ResidentSmem s;
  double* d = base;
  s.node = d; d += p;
  s.w0 = d; d += A;
  s.w1 = d; d += A;
  s.x = d; d += PASS2 ? A : 0;
  s.d = d; d += A;
  s.blk = d; d += (size_t)G * R;
  s.n0 = d; d += R;
  s.n1 = d; d += R;
  s.nx = d; d += PASS2 ? R : 0;
  uint32_t* u = reinterpret_cast<uint32_t*>(d);
  s.tail = u; u += A;
  s.head = u; u += A;
  s.lptr = u; u += p + 1;
  s.llong = u; u += max_long;
  s.lent = reinterpret_cast<uint16_t*>(u);
  return s;
}

struct ResidentCtx {
  uint32_t alo, nA;  // owned arcs [alo, alo+nA)
  uint32_t ulo, nU;  // owned node rows [ulo, ulo+nU) (node index, not row index)
  uint32_t nlong;    // long local lists
};

// Loads the CTA's slice of the operator into shared memory (once per kernel).  Caller syncs.
__device__ __forceinline__ ResidentCtx load_resident(const IncidenceOp& op, const ResidentOp& ro, const ResidentSmem& s) {
  ResidentCtx c;
  uint32_t ahi;
  cta_chunk(op.m, c.alo, ahi);
  c.nA = ahi - c.alo;
  c.ulo = min(op.p, blockIdx.x * ro.R);
  c.nU = min(op.p, c.ulo + ro.R) - c.ulo;
  const uint32_t* lp = ro.nl_ptr + (size_t)blockIdx.x * (op.p + 1);
  const uint32_t e0 = __ldg(lp), e1 = __ldg(lp + op.p);
  const uint32_t l0 = __ldg(ro.nl_long_ptr + blockIdx.x);
  c.nlong = __ldg(ro.nl_long_ptr + blockIdx.x + 1) - l0;
  for (uint32_t i = threadIdx.x; i < c.nA; i += kBlock) {
    s.d[i] = __ldg(op.d + c.alo + i);
    s.tail[i] = __ldg(op.tail + c.alo + i);
    s.head[i] = __ldg(op.head + c.alo + i);
  }
  for (uint32_t i = threadIdx.x; i <= op.p; i += kBlock) s.lptr[i] = __ldg(lp + i) - e0;
  for (uint32_t i = threadIdx.x; i < e1 - e0; i += kBlock) s.lent[i] = __ldg(ro.nl_ent + e0 + i);
  for (uint32_t i = threadIdx.x; i < c.nlong; i += kBlock) s.llong[i] = __ldg(ro.nl_long + l0 + i);
  return c;
}

// Step 1 of the node-row scheme: partial node sums of this CTA's arc values `w` (shared memory), published to
// Pout = partials[parity] as chunk [g][cta][0..R) for every node block g.
__device__ __forceinline__ void publish_partials(const IncidenceOp& op, const ResidentOp& ro, const ResidentSmem& s,
                                                 const ResidentCtx& c, const double* w, double* Pout) {
  const uint32_t G = gridDim.x, R = ro.R;
  double* mine = Pout + (size_t)blockIdx.x * R;
  for (uint32_t u = threadIdx.x; u < op.p; u += kBlock) {
    const uint32_t e0 = s.lptr[u], e1 = s.lptr[u + 1];
    if (e1 - e0 > kLongList) continue;
    double acc = 0.0;
    for (uint32_t e = e0; e < e1; ++e) {
      const uint32_t ent = s.lent[e];
      const double val = w[ent & 0x7fffu];
      acc = (ent & 0x8000u) ? __dsub_rn(acc, val) : __dadd_rn(acc, val);
    }
    const uint32_t g = u / R;
    __stcg(mine + (size_t)g * G * R + (u - g * R), acc);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t q = warp; q < c.nlong; q += kWarps) {
    const uint32_t u = s.llong[q];
    const uint32_t e0 = s.lptr[u], e1 = s.lptr[u + 1];
    double acc = 0.0;
    for (uint32_t e = e0 + lane; e < e1; e += 32) {
      const uint32_t ent = s.lent[e];
      const double val = w[ent & 0x7fffu];
      acc = (ent & 0x8000u) ? __dsub_rn(acc, val) : __dadd_rn(acc, val);
    }
    acc = warp_sum(acc);
    const uint32_t g = u / R;
    if (lane == 0) __stcg(mine + (size_t)g * G * R + (u - g * R), acc);
  }
}

// Step 2, first half: everything a step reads from global memory, in one batch -- the node segment of the
// current vector (scaled on the way in) and the [G][R] block of partials addressed to this CTA.  Caller syncs.
__device__ __forceinline__ void stage_step_inputs(const IncidenceOp& op, const ResidentOp& ro, const ResidentSmem& s,
                                                  const double* Xnode, double scale, const double* Pin) {
  const uint32_t GR = gridDim.x * ro.R;
  const double* blk = Pin + (size_t)blockIdx.x * GR;
  for (uint32_t i = threadIdx.x; i < GR; i += kBlock) s.blk[i] = __ldcg(blk + i);
  for (uint32_t u = threadIdx.x; u < op.p; u += kBlock) s.node[u] = __dmul_rn(__ldcg(Xnode + u), scale);
}
// Step 2, second half (one warp per owned node): T_u = sum over the G partials in a fixed order.
__device__ __forceinline__ double node_total(const ResidentSmem& s, uint32_t R, uint32_t r, int lane) {
  double acc = 0.0;
  for (uint32_t cta = lane; cta < gridDim.x; cta += 32) acc = __dadd_rn(acc, s.blk[cta * R + r]);
  return warp_sum(acc);
}

template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_resident_kernel(const IncidenceOp op, const ResidentOp ro, const Pass1Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  const uint32_t G = gridDim.x, R = ro.R;
  const uint32_t A = (op.m + G - 1) / G;
  const ResidentSmem s = carve_resident<false>(smem, op.p, A, G, R, ro.max_long);
  const ResidentCtx c = load_resident(op, ro, s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t pstride = (size_t)G * G * R;

  unsigned int epoch = a.st->epoch;
  int steps = 0, status = ST_RUNNING;
  double sc = 1.0, sp = 1.0, bp = 0.0, bnorm = 0.0;
  GridSync gs = a.gs;
  {
    // K0: ||b||; the arc / node slices of b become the current vector, the previous one is zero; the partial
    // node sums of b are published for step 0
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < c.nA; i += kBlock) {
      const double bi = __ldg(a.b + c.alo + i);
      s.w0[i] = bi;
      s.w1[i] = 0.0;
      acc = fma(bi, bi, acc);
    }
    for (uint32_t r = threadIdx.x; r < c.nU; r += kBlock) {
      const double bi = __ldg(a.b + op.m + c.ulo + r);
      s.n0[r] = bi;
      s.n1[r] = 0.0;
      acc = fma(bi, bi, acc);
    }
    __syncthreads();
    publish_partials(op, ro, s, c, s.w0, ro.partials);
    bnorm = sqrt(grid_sync<true>(acc, gs, epoch, sh));
    if (bnorm <= a.tol) status = ST_ZERO_B;
    sc = 1.0 / bnorm;
  }
  if (status == ST_RUNNING) {
    for (int j = 0; j < a.j_end; ++j) {
      gs.trace_step = j;
      trace_mark(gs.trace, j, 0);
      // node part of the current vector: b itself for step 0, then what the owners published last step
      const double* Xnode = j == 0 ? a.b + op.m : ro.nodebuf + (size_t)((j - 1) & 1) * op.p;
      double* Nout = ro.nodebuf + (size_t)(j & 1) * op.p;
      const double* Pin = ro.partials + (size_t)(j & 1) * pstride;
      double* Pout = ro.partials + (size_t)((j + 1) & 1) * pstride;
      double* cur = (j & 1) ? s.w1 : s.w0;
      double* prv = (j & 1) ? s.w0 : s.w1;
      double* ncur = (j & 1) ? s.n1 : s.n0;
      double* nprv = (j & 1) ? s.n0 : s.n1;
      double* Vcol = WITH_V ? a.V + (size_t)j * a.ldv : nullptr;

      // ---------------- phase A
      stage_step_inputs(op, ro, s, Xnode, sc, Pin);
      trace_mark(gs.trace, j, 1);
      __syncthreads();
      trace_mark(gs.trace, j, 2);
      double acc = 0.0;
      for (uint32_t r = warp; r < c.nU; r += kWarps) {  // node rows of the owned block
        const double t = __dmul_rn(sc, node_total(s, R, r, lane));
        if (lane == 0) {
          const double v = __dmul_rn(ncur[r], sc);
          const double vp = __dmul_rn(nprv[r], sp);
          const double wt = rec_sub(t, bp, vp);
          acc = fma(v, wt, acc);
          nprv[r] = wt;
          if (WITH_V) __stcs(Vcol + op.m + c.ulo + r, v);
        }
      }
      for (uint32_t i = threadIdx.x; i < c.nA; i += kBlock) {
        const double v = __dmul_rn(cur[i], sc);
        const double vp = __dmul_rn(prv[i], sp);
        const uint32_t t = s.tail[i], h = s.head[i];
        const double wt = rec_sub(arc_row(s.d[i], v, t, h, s.node[t], s.node[h]), bp, vp);
        acc = fma(v, wt, acc);
        prv[i] = wt;  // the previous vector's slot becomes w~ (then w, then the next current vector)
        if (WITH_V) __stcs(Vcol + c.alo + i, v);
      }
      trace_mark(gs.trace, j, 3);
      gs.trace_base = 4;
      const double alpha = grid_sync<true>(acc, gs, epoch, sh);

      // ---------------- phase B: w = w~ - alpha v; node values and partial node sums of w are published
      acc = 0.0;
      for (uint32_t i = threadIdx.x; i < c.nA; i += kBlock) {
        const double w = rec_sub(prv[i], alpha, __dmul_rn(cur[i], sc));
        prv[i] = w;
        acc = fma(w, w, acc);
      }
      for (uint32_t r = threadIdx.x; r < c.nU; r += kBlock) {
        const double w = rec_sub(nprv[r], alpha, __dmul_rn(ncur[r], sc));
        nprv[r] = w;
        __stcg(Nout + c.ulo + r, w);
        acc = fma(w, w, acc);
      }
      __syncthreads();
      trace_mark(gs.trace, j, 8);
      publish_partials(op, ro, s, c, prv, Pout);
      gs.trace_base = 9;
      const double beta = sqrt(grid_sync<true>(acc, gs, epoch, sh));

      if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.alphas[j] = alpha;
        a.betas[j] = beta;
      }
      steps = j + 1;
      if (beta <= a.tol) {
        status = ST_BREAKDOWN;
        break;
      }
      sp = sc;
      sc = 1.0 / beta;
      bp = beta;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    State st;
    st.s_cur = sc;
    st.s_prev = sp;
    st.beta_prev = bp;
    st.b_norm = bnorm;
    st.epoch = epoch;
    st.rot = 0;
    st.steps = steps;
    st.status = status;
    *a.st = st;
  }
}

template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_resident_kernel(const IncidenceOp op, const ResidentOp ro, const Pass2Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  const uint32_t G = gridDim.x, R = ro.R;
  const uint32_t A = (op.m + G - 1) / G;
  const ResidentSmem s = carve_resident<true>(smem, op.p, A, G, R, ro.max_long);
  const ResidentCtx c = load_resident(op, ro, s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t pstride = (size_t)G * G * R;
  unsigned int epoch = a.st->epoch;
  GridSync gs = a.gs;
  // coefficients are loaded one step ahead of their use so that no step starts with a dependent load
  double c_alpha = 0.0, c_beta = 1.0, c_y = 0.0;
  if (a.steps > 1) {
    c_alpha = __ldg(a.alphas);
    c_beta = __ldg(a.betas);
    c_y = __ldg(a.y + 1);
  }
  double sc = 1.0 / a.b_norm, sp = 1.0, bp = 0.0;
  {
    // v_1 = b * (1/||b||) held lazily as (b, 1/||b||); x = y_0 v_1   (lanczos_two_pass.rs:247-258)
    const double y0 = __ldg(a.y);
    for (uint32_t i = threadIdx.x; i < c.nA; i += kBlock) {
      const double bi = __ldg(a.b + c.alo + i);
      const double v = __dmul_rn(bi, sc);
      s.w0[i] = bi;
      s.w1[i] = 0.0;
      s.x[i] = __dmul_rn(v, y0);
      if (WITH_V) __stcs(a.V + c.alo + i, v);
    }
    for (uint32_t r = threadIdx.x; r < c.nU; r += kBlock) {
      const double bi = __ldg(a.b + op.m + c.ulo + r);
      const double v = __dmul_rn(bi, sc);
      s.n0[r] = bi;
      s.n1[r] = 0.0;
      s.nx[r] = __dmul_rn(v, y0);
      if (WITH_V) __stcs(a.V + op.m + c.ulo + r, v);
    }
    __syncthreads();
    publish_partials(op, ro, s, c, s.w0, ro.partials);
    grid_sync<false>(0.0, gs, epoch, sh);
  }
  gs.trace_base = 4;
  for (int j = 0; j + 1 < a.steps; ++j) {
    gs.trace_step = j;
    trace_mark(gs.trace, j, 0);
    const double* Xnode = j == 0 ? a.b + op.m : ro.nodebuf + (size_t)((j - 1) & 1) * op.p;
    double* Nout = ro.nodebuf + (size_t)(j & 1) * op.p;
    const double* Pin = ro.partials + (size_t)(j & 1) * pstride;
    double* Pout = ro.partials + (size_t)((j + 1) & 1) * pstride;
    double* cur = (j & 1) ? s.w1 : s.w0;
    double* prv = (j & 1) ? s.w0 : s.w1;
    double* ncur = (j & 1) ? s.n1 : s.n0;
    double* nprv = (j & 1) ? s.n0 : s.n1;
    double* Vcol = WITH_V ? a.V + (size_t)(j + 1) * a.ldv : nullptr;
    const double alpha = c_alpha, beta = c_beta, yj = c_y;
    const double sinv = 1.0 / beta;
    if (j + 2 < a.steps) {  // prefetch the next step's coefficients
      c_alpha = __ldg(a.alphas + j + 1);
      c_beta = __ldg(a.betas + j + 1);
      c_y = __ldg(a.y + j + 2);
    }

    stage_step_inputs(op, ro, s, Xnode, sc, Pin);
    trace_mark(gs.trace, j, 1);
    __syncthreads();
    trace_mark(gs.trace, j, 2);
    for (uint32_t r = warp; r < c.nU; r += kWarps) {
      const double t = __dmul_rn(sc, node_total(s, R, r, lane));
      if (lane == 0) {
        const double v = __dmul_rn(ncur[r], sc);
        const double w = rec_sub(rec_sub(t, bp, __dmul_rn(nprv[r], sp)), alpha, v);
        const double vn = __dmul_rn(w, sinv);
        nprv[r] = w;
        __stcg(Nout + c.ulo + r, w);
        s.nx[r] = __dadd_rn(s.nx[r], __dmul_rn(yj, vn));
        if (WITH_V) __stcs(Vcol + op.m + c.ulo + r, vn);
      }
    }
    for (uint32_t i = threadIdx.x; i < c.nA; i += kBlock) {
      const double v = __dmul_rn(cur[i], sc);
      const uint32_t t = s.tail[i], h = s.head[i];
      const double w =
          rec_sub(rec_sub(arc_row(s.d[i], v, t, h, s.node[t], s.node[h]), bp, __dmul_rn(prv[i], sp)), alpha, v);
      const double vn = __dmul_rn(w, sinv);
      prv[i] = w;
      s.x[i] = __dadd_rn(s.x[i], __dmul_rn(yj, vn));
      if (WITH_V) __stcs(Vcol + c.alo + i, vn);
    }
    __syncthreads();
    trace_mark(gs.trace, j, 3);
    publish_partials(op, ro, s, c, prv, Pout);
    grid_sync<false>(0.0, gs, epoch, sh);
    sp = sc;
    sc = sinv;
    bp = beta;
  }
  for (uint32_t i = threadIdx.x; i < c.nA; i += kBlock) a.x[c.alo + i] = s.x[i];
  for (uint32_t r = threadIdx.x; r < c.nU; r += kBlock) a.x[op.m + c.ulo + r] = s.nx[r];
  if (blockIdx.x == 0 && threadIdx.x == 0) a.st->epoch = epoch;
}

// ----------------------------------------------------------------------------- LinOp::apply
template <class OP>
__global__ void __launch_bounds__(kBlock, 1) apply_kernel(const OP op, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ double smem[];
  double* sm_node = smem;
  double* sm_seg = smem + node_smem_doubles(op);
  const double* nodes = stage_nodes(op, x, 1.0, sm_node);
  __syncthreads();
  const auto dev = make_dev(op, nodes);
  uint32_t slo, shi;
  cta_chunk(dev.num_short(), slo, shi);
  for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock)
    if (dev.is_short(i)) y[i] = dev.short_row(i, __ldcg(x + i), x, 1.0);
  const LongRows& lr = long_rows(op);
  if (lr.nlong) {
    uint32_t r0, r1, s0;
    long_row_segments(dev, lr, x, 1.0, sm_seg, r0, r1, s0);
    __syncthreads();
    for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) y[__ldg(lr.row + q)] = long_row_total(lr, q, sm_seg, s0);
  }
}

// ----------------------------------------------------------------------------- one-pass reconstruction
// x = b_norm * (V y')   (matmul(x, Replace, V, y', alpha = ||b||), src/solvers.rs:96-104): one streaming
// pass over the n x steps basis, y' staged in shared memory, 8 independent column streams per thread.
__global__ void __launch_bounds__(256) gemv_vy_kernel(const double* __restrict__ V, size_t ldv, uint32_t n, int steps,
                                                       const double* __restrict__ y, double b_norm,
                                                       double* __restrict__ x) {
  extern __shared__ double sy[];
  for (int j = threadIdx.x; j < steps; j += blockDim.x) sy[j] = y[j];
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double* col = V + i;
    double acc = 0.0;
    int j = 0;
    for (; j + 8 <= steps; j += 8) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldcs(col + (size_t)(j + u) * ldv);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, __dmul_rn(t[u], sy[j + u]));
    }
    for (; j < steps; ++j) acc = __dadd_rn(acc, __dmul_rn(__ldcs(col + (size_t)j * ldv), sy[j]));
    x[i] = __dmul_rn(b_norm, acc);
  }
}

// k-sweep reconstruction: x_q = b_norm * (V[:, :steps_q] y'_q) for up to kSweepChunk right-hand coefficient vectors in ONE
// streaming pass over the basis (the reference's sweep re-solves for every k, src/bin/tradeoff.rs:262-290).  Y is kmax x nq
// row-major, zero beyond steps_q: adding 0 * v leaves an accumulator unchanged, and every x_q sees exactly the sequence of
// roundings of gemv_vy_kernel -- bit-identical to the per-k solves.
constexpr int kSweepChunk = 16;
__global__ void __launch_bounds__(256) gemv_vy_sweep_kernel(const double* __restrict__ V, size_t ldv, uint32_t n, int kmax,
                                                             const double* __restrict__ Y, int nq, double b_norm,
                                                             double* __restrict__ X, size_t ldx) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double* col = V + i;
    double acc[kSweepChunk];
#pragma unroll
    for (int q = 0; q < kSweepChunk; ++q) acc[q] = 0.0;
    int j = 0;
    for (; j + 4 <= kmax; j += 4) {
      double t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = __ldcs(col + (size_t)(j + u) * ldv);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double* yrow = Y + (size_t)(j + u) * nq;  // the same words for every thread: one broadcast load each
#pragma unroll
        for (int q = 0; q < kSweepChunk; ++q)
          if (q < nq) acc[q] = __dadd_rn(acc[q], __dmul_rn(t[u], __ldg(yrow + q)));
      }
    }
    for (; j < kmax; ++j) {
      const double t = __ldcs(col + (size_t)j * ldv);
      const double* yrow = Y + (size_t)j * nq;
#pragma unroll
      for (int q = 0; q < kSweepChunk; ++q)
        if (q < nq) acc[q] = __dadd_rn(acc[q], __dmul_rn(t, __ldg(yrow + q)));
    }
#pragma unroll
    for (int q = 0; q < kSweepChunk; ++q)
      if (q < nq) X[(size_t)q * ldx + i] = __dmul_rn(b_norm, acc[q]);
  }
}

}  // namespace tpl
