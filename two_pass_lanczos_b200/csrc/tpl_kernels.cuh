// tpl_kernels.cuh -- sm_100a device code of the two-pass Lanczos engine.
//
// Design (DESIGN.md has the long version):
//   * ONE persistent cooperative kernel per pass.  The grid is one CTA per SM; Lanczos steps are a loop
//     inside the kernel and the two dependent reductions of a step (alpha, beta) are grid-wide
//     flag barriers that carry the reduction payload (no atomics, fixed summation order).
//   * The operator is either the KKT network-incidence form (arc rows: thread per arc, node segment
//     staged in shared memory; node rows: warp per fixed-length segment of the node->arc list) or a
//     generic CSR (short rows: thread per row; long rows: the same segment machinery).
//   * Every per-element expression of the recurrence is written once (rec_sub below) with explicit
//     round-to-nearest mul/sub so that pass 1, the one-pass variant and pass 2 produce bit-identical
//     basis vectors (reference invariant `basis_drift == 0`, results/orthogonality_*.csv) and follow the
//     reference's two-rounding `sub(w, mul(c, v))` (src/algorithms/mod.rs:183-198).
//   * Vectors written inside a kernel are read by other CTAs only through L2 (ld.global.cg / st.global.cg).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpl {

constexpr int kBlock = 1024;  // threads per CTA (one CTA per SM)
constexpr int kWarps = kBlock / 32;
constexpr uint32_t kSignBit = 0x80000000u;
constexpr uint32_t kSpinLimit = 1u << 24;  // grid-barrier watchdog: trap instead of hanging the GPU

enum : int { ST_RUNNING = 0, ST_BREAKDOWN = 1, ST_ZERO_B = 2 };

// Rows whose product is summed by warp-per-segment: KKT node rows, long CSR rows.
struct LongRows {
  uint32_t nlong;           // number of long rows
  uint32_t max_segs;        // max segments owned by one CTA (shared-memory sizing)
  const uint32_t* row;      // [nlong]   global row id
  const uint32_t* seg_ptr;  // [nlong+1] first segment of each long row
  const uint32_t* ent_ptr;  // [nseg+1]  entry range of each segment
  const uint32_t* ent_idx;  // [nent]    column (incidence: bit 31 set = coefficient -1)
  const double* ent_val;    // [nent]    CSR only
  const uint32_t* cta_ptr;  // [G+1]     long rows owned by each CTA
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

struct GridSync {
  unsigned int* flags;  // [G] epoch reached by each CTA
  double* partials;     // [2][G] reduction payload, double-buffered by epoch parity
};

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
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

struct CtaShared {
  double warp_part[kWarps];
  double result;
};

// Grid-wide sum (REDUCE) or plain barrier.  Every CTA publishes its partial and an epoch flag with
// release semantics; warp 0 of every CTA polls all flags, then adds the G partials in a fixed order,
// so the value is identical in all CTAs and from run to run.  Two __syncthreads per call.
template <bool REDUCE>
__device__ __forceinline__ double grid_sync(double v, const GridSync& gs, unsigned int& epoch, CtaShared& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;
  epoch += 1;
  if (REDUCE) {
    v = warp_sum(v);
    if (lane == 0) sh.warp_part[warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
    double* part = gs.partials + (size_t)(epoch & 1u) * G;
    if (REDUCE) {
      double t = lane < kWarps ? sh.warp_part[lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) __stcg(part + blockIdx.x, t);
    }
    if (lane == 0) st_release_gpu(gs.flags + blockIdx.x, epoch);
    for (unsigned int i = lane; i < G; i += 32) {
      unsigned int spins = 0;
      while ((int)(ld_relaxed_gpu(gs.flags + i) - epoch) < 0) {
        if (++spins > kSpinLimit) __trap();
      }
    }
    __syncwarp();
    fence_acq_rel_gpu();
    if (REDUCE) {
      double s = 0.0;
      for (unsigned int i = lane; i < G; i += 32) s += __ldcg(part + i);
      s = warp_sum(s);
      if (lane == 0) sh.result = s;
    }
  }
  __syncthreads();
  return REDUCE ? sh.result : 0.0;
}

// ----------------------------------------------------------------------------- operator products
// x is always addressed as X[i]*s (lazy normalisation: pass 1 keeps the un-normalised w and the
// reciprocal norm; X[i]*s is the single rounding the reference performs when it scales w in place,
// src/algorithms/mod.rs:312-315).  Pass 2 and apply() use s = 1.
struct IncidenceDev {
  const IncidenceOp& op;
  const double* sm_node;  // staged, already scaled node segment (or nullptr)
  __device__ __forceinline__ uint32_t num_short() const { return op.m; }
  __device__ __forceinline__ double node(uint32_t u, const double* X, double s) const {
    return sm_node ? sm_node[u] : __dmul_rn(__ldcg(X + op.m + u), s);
  }
  // (A x)_j for arc row j in the reference's CSC accumulation order: D_jj x_j first, then the
  // incident node columns in ascending node index (columns m+tail, m+head).
  __device__ __forceinline__ bool short_row(uint32_t j, double xj, const double* X, double s, double& out) const {
    const uint32_t t = __ldg(op.tail + j), h = __ldg(op.head + j);
    const double dj = __ldg(op.d + j);
    const double xt = node(t, X, s), xh = node(h, X, s);
    double acc = __dmul_rn(dj, xj);
    if (t == h) {
      // self-loop: the loader's merged E entry is an explicit 0 (data_loader.rs:118-133), contributes nothing
    } else if (t < h) {
      acc = __dadd_rn(acc, xt);
      acc = __dsub_rn(acc, xh);
    } else {
      acc = __dsub_rn(acc, xh);
      acc = __dadd_rn(acc, xt);
    }
    out = acc;
    return true;
  }
  __device__ __forceinline__ double entry(const LongRows& lr, uint32_t e, const double* X, double s, double acc) const {
    const uint32_t idx = __ldg(lr.ent_idx + e);
    const double x = __dmul_rn(__ldcg(X + (idx & ~kSignBit)), s);
    return (idx & kSignBit) ? __dsub_rn(acc, x) : __dadd_rn(acc, x);
  }
};

struct CsrDev {
  const CsrOp& op;
  __device__ __forceinline__ uint32_t num_short() const { return op.n; }
  __device__ __forceinline__ bool short_row(uint32_t i, double, const double* X, double s, double& out) const {
    const uint32_t p0 = __ldg(op.row_ptr + i), p1 = __ldg(op.row_ptr + i + 1);
    if (p1 - p0 > op.long_thresh) return false;  // summed by the segment path
    double acc = 0.0;
    for (uint32_t p = p0; p < p1; ++p) {
      const double x = __dmul_rn(__ldcg(X + __ldg(op.col + p)), s);
      acc = __dadd_rn(acc, __dmul_rn(__ldg(op.val + p), x));
    }
    out = acc;
    return true;
  }
  __device__ __forceinline__ double entry(const LongRows& lr, uint32_t e, const double* X, double s, double acc) const {
    const double x = __dmul_rn(__ldcg(X + __ldg(lr.ent_idx + e)), s);
    return __dadd_rn(acc, __dmul_rn(__ldg(lr.ent_val + e), x));
  }
};

__device__ __forceinline__ const LongRows& long_rows(const IncidenceOp& op) { return op.lr; }
__device__ __forceinline__ const LongRows& long_rows(const CsrOp& op) { return op.lr; }
__device__ __forceinline__ uint32_t op_rows(const IncidenceOp& op) { return op.m + op.p; }
__device__ __forceinline__ uint32_t op_rows(const CsrOp& op) { return op.n; }

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
    if (lane == 0) sm_seg[sg - s0] = acc;
  }
}
__device__ __forceinline__ double long_row_total(const LongRows& lr, uint32_t q, const double* sm_seg, uint32_t s0) {
  const uint32_t a = __ldg(lr.seg_ptr + q), b = __ldg(lr.seg_ptr + q + 1);
  double t = 0.0;
  for (uint32_t sg = a; sg < b; ++sg) t = __dadd_rn(t, sm_seg[sg - s0]);
  return t;
}

// ----------------------------------------------------------------------------- pass 1 / one-pass
// Replaces lanczos_pass_one (src/algorithms/lanczos_two_pass.rs:65-110) and, with WITH_V, the basis
// generation of lanczos_standard (src/algorithms/lanczos.rs:55-156).  Per step (mod.rs:167-212, 292-340):
//   phase A  w~ = A v_j - beta_{j-1} v_{j-1},  alpha_j = <v_j, w~>      (one sweep + grid reduction)
//   phase B  w  = w~ - alpha_j v_j,            beta_j  = ||w||          (one sweep + grid reduction)
// v_j is held as (W_cur, s_cur) with v = W_cur * s_cur.
template <class OP, bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_kernel(const OP op, const Pass1Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  double* sm_node = smem;
  double* sm_seg = smem + node_smem_doubles(op);

  const State st0 = *a.st;
  unsigned int epoch = st0.epoch;
  int rot = st0.rot, steps = st0.steps, status = st0.status;
  double sc = st0.s_cur, sp = st0.s_prev, bp = st0.beta_prev, bnorm = st0.b_norm;
  const uint32_t n = a.n;
  uint32_t lo, hi;
  cta_chunk(n, lo, hi);

  if (a.j_begin == 0) {
    // K0: ||b||, W_cur = b, W_prev = 0   (lanczos_two_pass.rs:74, mod.rs:261-289)
    double* Wp = a.buf[rot];
    double* Wc = a.buf[(rot + 1) % 3];
    double acc = 0.0;
#pragma unroll 4
    for (uint32_t i = lo + threadIdx.x; i < hi; i += kBlock) {
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
      const double* Wp = a.buf[rot];
      const double* Wc = a.buf[(rot + 1) % 3];
      double* Wn = a.buf[(rot + 2) % 3];
      double* Vcol = WITH_V ? a.V + (size_t)j * a.ldv : nullptr;

      // ---------------- phase A
      const double* nodes = stage_nodes(op, Wc, sc, sm_node);
      __syncthreads();
      const auto dev = make_dev(op, nodes);
      double acc = 0.0;
      {
        uint32_t slo, shi;
        cta_chunk(dev.num_short(), slo, shi);
#pragma unroll 2
        for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
          const double v = __dmul_rn(__ldcg(Wc + i), sc);
          double t;
          if (dev.short_row(i, v, Wc, sc, t)) {
            const double vp = __dmul_rn(__ldcg(Wp + i), sp);
            const double wt = rec_sub(t, bp, vp);
            acc = fma(v, wt, acc);
            __stcg(Wn + i, wt);
            if (WITH_V) __stcs(Vcol + i, v);
          }
        }
      }
      const LongRows& lr = long_rows(op);
      if (lr.nlong) {
        uint32_t r0, r1, s0;
        long_row_segments(dev, lr, Wc, sc, sm_seg, r0, r1, s0);
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
      const double alpha = grid_sync<true>(acc, a.gs, epoch, sh);

      // ---------------- phase B
      acc = 0.0;
#pragma unroll 4
      for (uint32_t i = lo + threadIdx.x; i < hi; i += kBlock) {
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

// ----------------------------------------------------------------------------- pass 2
// Replaces lanczos_pass_two_impl (src/algorithms/lanczos_two_pass.rs:206-312): regenerates v_{j+1}
// with the stored alpha_j, beta_{j-1}, beta_j and accumulates x += y_{j+1} v_{j+1} in the same sweep.
// One grid barrier per step.
template <class OP, bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_kernel(const OP op, const Pass2Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  double* sm_node = smem;
  double* sm_seg = smem + node_smem_doubles(op);

  const State st0 = *a.st;
  unsigned int epoch = st0.epoch;
  const uint32_t n = a.n;
  uint32_t lo, hi;
  cta_chunk(n, lo, hi);
  int rot = 0;
  {
    // v_1 = b * (1/||b||), x = y_0 v_1   (lanczos_two_pass.rs:247-258)
    const double inv = 1.0 / a.b_norm;
    const double y0 = __ldg(a.y);
    double* Vp = a.buf[0];
    double* Vc = a.buf[1];
#pragma unroll 4
    for (uint32_t i = lo + threadIdx.x; i < hi; i += kBlock) {
      const double v = __dmul_rn(__ldg(a.b + i), inv);
      __stcg(Vc + i, v);
      __stcg(Vp + i, 0.0);
      __stcg(a.x + i, __dmul_rn(v, y0));
      if (WITH_V) __stcs(a.V + i, v);
    }
    grid_sync<false>(0.0, a.gs, epoch, sh);
  }
  for (int j = 0; j + 1 < a.steps; ++j) {
    const double* Vp = a.buf[rot];
    const double* Vc = a.buf[(rot + 1) % 3];
    double* Vn = a.buf[(rot + 2) % 3];
    double* Vcol = WITH_V ? a.V + (size_t)(j + 1) * a.ldv : nullptr;
    const double alpha = __ldg(a.alphas + j);
    const double beta = __ldg(a.betas + j);
    const double bp = j == 0 ? 0.0 : __ldg(a.betas + j - 1);
    const double s = 1.0 / beta;
    const double yj = __ldg(a.y + j + 1);

    const double* nodes = stage_nodes(op, Vc, 1.0, sm_node);
    __syncthreads();
    const auto dev = make_dev(op, nodes);
    {
      uint32_t slo, shi;
      cta_chunk(dev.num_short(), slo, shi);
#pragma unroll 2
      for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
        const double v = __ldcg(Vc + i);
        double t;
        if (dev.short_row(i, v, Vc, 1.0, t)) {
          const double w = rec_sub(rec_sub(t, bp, __ldcg(Vp + i)), alpha, v);
          const double vn = __dmul_rn(w, s);
          __stcg(Vn + i, vn);
          __stcg(a.x + i, __dadd_rn(__ldcg(a.x + i), __dmul_rn(yj, vn)));
          if (WITH_V) __stcs(Vcol + i, vn);
        }
      }
    }
    const LongRows& lr = long_rows(op);
    if (lr.nlong) {
      uint32_t r0, r1, s0;
      long_row_segments(dev, lr, Vc, 1.0, sm_seg, r0, r1, s0);
      __syncthreads();
      for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
        const uint32_t i = __ldg(lr.row + q);
        const double t = long_row_total(lr, q, sm_seg, s0);
        const double v = __ldcg(Vc + i);
        const double w = rec_sub(rec_sub(t, bp, __ldcg(Vp + i)), alpha, v);
        const double vn = __dmul_rn(w, s);
        __stcg(Vn + i, vn);
        __stcg(a.x + i, __dadd_rn(__ldcg(a.x + i), __dmul_rn(yj, vn)));
        if (WITH_V) __stcs(Vcol + i, vn);
      }
    }
    grid_sync<false>(0.0, a.gs, epoch, sh);
    rot = (rot + 1) % 3;
  }
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
  for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
    double t;
    if (dev.short_row(i, __ldcg(x + i), x, 1.0, t)) y[i] = t;
  }
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

}  // namespace tpl
