// tpl_csr.cuh -- generic sparse symmetric operator (what `&SparseColMat<usize, f64>` is to the reference, src/algorithms/mod.rs:177):
// row-binned CSR kernels fused with the Lanczos recurrence.
//
// Kernel-side format (built once from the host CSC, which is the CSR of a symmetric matrix):
//   * rows with at most `long_thresh` entries live in SELL-32 slices: 32 consecutive rows, entry e of row r at
//     scol / sval[sptr[slice] + 32 e + (r mod 32)] -- a warp reads entry e of its 32 rows as ONE coalesced 128 / 256-byte
//     access, no per-row pointer on the way (the thread-per-row CSR loop it replaces chased row_ptr -> col -> x with two
//     rows per thread in flight and was latency-bound at 0.36 of the HBM roofline on a 5M-arc KKT matrix).  A slice is as
//     wide as its longest short row; rlen[r] predicates the padding away.  Entries keep their ascending column order, i.e.
//     the reference's accumulation order;
//   * longer rows (the node rows of a KKT matrix: ~2m/p entries) are cut into 256-entry segments summed by a warp each
//     (lanes stride the entries, xor-shuffle tree), balanced over the CTAs -- LongRows of tpl_kernels.cuh.
// A warp works on kCsrU slices at a time, level by level: slice pointers, row lengths and the Lanczos vectors of 4 x 32 rows
// first, then every (col, val) pair, then every operand gather -- 4 rows x up to 4 entries per thread in flight.
// Per-element arithmetic is the shared one (rec_sub, two roundings; lazy scaling): pass 1, the one-pass variant and pass 2
// regenerate bit-identical vectors.  Byte model: B_csr = 12 nnz + 4 (n + 1) per product.
#pragma once
#include "tpl_kernels.cuh"

namespace tpl {

constexpr int kCsrU = 4;  // slices a warp has in flight
constexpr int kCsrW = 4;  // entries of a row loaded together (longer short rows continue in a loop)

struct SellOp {
  uint32_t n;
  uint32_t nslice;          // ceil(n / 32)
  const uint32_t* sptr;     // [nslice + 1] first word of every slice in scol / sval (multiples of 32)
  const uint32_t* scol;     // columns, slice-interleaved; padding = the row itself
  const double* sval;       // values, 0 in padding
  const uint8_t* rlen;      // [n] entries of a short row; 0xff = long row (segment path)
  LongRows lr;
};

// (A x)_i, x = X * s, for the lane's row of each of the warp's kCsrU slices; `live[q]` = the row exists and is short.
struct SellBatch {
  uint32_t row[kCsrU];
  bool live[kCsrU];
  double t[kCsrU];
};
// First load level of a batch (slice pointers, row lengths): requested one batch AHEAD by the sweeps, so that a batch costs two
// dependent round trips (entries, operand gathers) instead of three.
struct SellHeads {
  uint32_t base[kCsrU], len[kCsrU];  // len: 0xff = no short row here
};
__device__ __forceinline__ void sell_heads(const SellOp& op, uint32_t slice0, uint32_t stride, uint32_t shi, SellHeads& h) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < kCsrU; ++q) {
    const uint32_t sl = slice0 + q * stride, row = sl * 32u + lane;
    const bool in = sl < shi;
    h.base[q] = in ? __ldg(op.sptr + sl) : 0u;
    h.len[q] = in && row < op.n ? (uint32_t)__ldg(op.rlen + row) : 0xffu;
  }
}
__device__ __forceinline__ void sell_rows(const SellOp& op, uint32_t slice0, uint32_t stride, const SellHeads& h, const double* X, double s,
                                          SellBatch& b) {
  const int lane = threadIdx.x & 31;
  uint32_t base[kCsrU], len[kCsrU];
#pragma unroll
  for (int q = 0; q < kCsrU; ++q) {
    b.row[q] = (slice0 + q * stride) * 32u + lane;
    base[q] = h.base[q];
    b.live[q] = h.len[q] != 0xffu;
    len[q] = b.live[q] ? h.len[q] : 0u;
  }
  uint32_t c[kCsrU][kCsrW];
  double a[kCsrU][kCsrW];
#pragma unroll
  for (int q = 0; q < kCsrU; ++q)
#pragma unroll
    for (int e = 0; e < kCsrW; ++e) {
      const bool on = (uint32_t)e < len[q];
      const uint32_t w = base[q] + 32u * e + lane;
      c[q][e] = on ? __ldcs(op.scol + w) : 0u;  // the matrix is streamed once per sweep: evict-first, the gathered operand stays
      a[q][e] = on ? __ldcs(op.sval + w) : 0.0;
    }
  double x[kCsrU][kCsrW];
#pragma unroll
  for (int q = 0; q < kCsrU; ++q)
#pragma unroll
    for (int e = 0; e < kCsrW; ++e) x[q][e] = (uint32_t)e < len[q] ? __ldcg(X + c[q][e]) : 0.0;
#pragma unroll
  for (int q = 0; q < kCsrU; ++q) {
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < kCsrW; ++e)
      if ((uint32_t)e < len[q]) acc = __dadd_rn(acc, __dmul_rn(a[q][e], __dmul_rn(x[q][e], s)));
    for (uint32_t e = kCsrW; e < len[q]; ++e) {  // rows with 5 .. long_thresh entries
      const uint32_t w = base[q] + 32u * e + lane;
      acc = __dadd_rn(acc, __dmul_rn(__ldg(op.sval + w), __dmul_rn(__ldcg(X + __ldg(op.scol + w)), s)));
    }
    b.t[q] = acc;
  }
}

// Sums the segments of the long rows owned by this CTA into sm_seg: one warp per segment, lane l adds the entries e0 + l,
// e0 + l + 32, ... in that order, then the xor-shuffle tree (the order of long_row_segments in tpl_kernels.cuh).  The eight
// entries a lane owns of a 256-entry chunk are loaded level by level: all (column, value) pairs, then all operand gathers.
__device__ __forceinline__ void sell_long_segments(const LongRows& lr, const double* X, double s, double* sm_seg, uint32_t& r0,
                                                   uint32_t& r1, uint32_t& s0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  r0 = __ldg(lr.cta_ptr + blockIdx.x);
  r1 = __ldg(lr.cta_ptr + blockIdx.x + 1);
  s0 = __ldg(lr.seg_ptr + r0);
  const uint32_t s1 = __ldg(lr.seg_ptr + r1);
  for (uint32_t sg = s0 + warp; sg < s1; sg += kWarps) {
    const uint32_t e0 = __ldg(lr.ent_ptr + sg), e1 = __ldg(lr.ent_ptr + sg + 1);
    double acc = 0.0;
    for (uint32_t eb = e0; eb < e1; eb += 256) {
      uint32_t idx[8];
      double val[8], x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t e = eb + lane + 32 * q;
        idx[q] = e < e1 ? __ldcs(lr.ent_idx + e) : 0u;  // streamed once per sweep: evict-first
        val[q] = e < e1 ? __ldcs(lr.ent_val + e) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = eb + lane + 32 * q < e1 ? __ldcg(X + idx[q]) : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (eb + lane + 32 * q < e1) acc = __dadd_rn(acc, __dmul_rn(val[q], __dmul_rn(x[q], s)));
    }
    acc = warp_sum(acc);
    if (lane == 0) seg_put(lr, sm_seg, sg, s0, acc);
  }
}

// slices [slo, shi) of this CTA
__device__ __forceinline__ void cta_slices(const SellOp& op, uint32_t& slo, uint32_t& shi) { cta_chunk(op.nslice, slo, shi); }

// =============================================================================================
// pass 1 / one-pass basis generation (same step structure and state hand-over as pass1_kernel of tpl_kernels.cuh, so that
// a pass can also run one cooperative launch per step -- the LanczosCallback path, src/algorithms/lanczos.rs:93-106)
// =============================================================================================
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_csr_kernel(const SellOp op, const Pass1Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  double* sm_seg = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const State st0 = *a.st;
  unsigned int epoch = st0.epoch;
  int rot = st0.rot, steps = st0.steps, status = st0.status;
  double sc = st0.s_cur, sp = st0.s_prev, bp = st0.beta_prev, bnorm = st0.b_norm;
  const LongRows& lr = op.lr;
  uint32_t slo, shi;
  cta_slices(op, slo, shi);
  const uint32_t rlo = slo * 32u, rhi = min(op.n, shi * 32u);
  const uint32_t r0 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x) : 0;
  const uint32_t r1 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x + 1) : 0;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };

  if (a.j_begin == 0) {
    double* Wp = pick(rot);
    double* Wc = pick((rot + 1) % 3);
    double acc = 0.0;
#pragma unroll 4
    for (uint32_t i = rlo + threadIdx.x; i < rhi; i += kBlock) {
      if (__ldg(op.rlen + i) == 0xffu) continue;
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

      // ---------------- phase A: w~ = A v - beta_{j-1} v_{j-1}, alpha partial
      double acc = 0.0;
      SellHeads hd, hd_next;
      sell_heads(op, slo + warp, kWarps, shi, hd);
      for (uint32_t s0 = slo + warp; s0 < shi; s0 += kCsrU * kWarps) {
        SellBatch bt;
        double wc[kCsrU], wp[kCsrU];
        sell_heads(op, s0 + kCsrU * kWarps, kWarps, shi, hd_next);  // the next batch's pointers and lengths
#pragma unroll
        for (int q = 0; q < kCsrU; ++q) {  // the vectors of the four rows are requested together with the slice data
          const uint32_t i = min((s0 + q * kWarps) * 32u + lane, op.n - 1);
          wc[q] = __ldcg(Wc + i);
          wp[q] = __ldcg(Wp + i);
        }
        sell_rows(op, s0, kWarps, hd, Wc, sc, bt);
        hd = hd_next;
#pragma unroll
        for (int q = 0; q < kCsrU; ++q)
          if (bt.live[q]) {
            const double v = __dmul_rn(wc[q], sc);
            const double wt = rec_sub(bt.t[q], bp, __dmul_rn(wp[q], sp));
            acc = fma(v, wt, acc);
            __stcg(Wn + bt.row[q], wt);
            if (WITH_V) __stcs(Vcol + bt.row[q], v);
          }
      }
      if (lr.nlong) {
        uint32_t q0, q1, sg0;
        sell_long_segments(lr, Wc, sc, sm_seg, q0, q1, sg0);
        __syncthreads();
        for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
          const uint32_t i = __ldg(lr.row + q);
          const double t = long_row_total(lr, q, sm_seg, sg0);
          const double v = __dmul_rn(__ldcg(Wc + i), sc);
          const double wt = rec_sub(t, bp, __dmul_rn(__ldcg(Wp + i), sp));
          acc = fma(v, wt, acc);
          __stcg(Wn + i, wt);
          if (WITH_V) __stcs(Vcol + i, v);
        }
      }
      // (all-reduce only: phase B reads what this CTA wrote in phase A, ordered by the CTA barriers inside the call)
      const double alpha = grid_sync<true, false>(acc, a.gs, epoch, sh);

      // ---------------- phase B: w = w~ - alpha v, beta partial (same row ownership: the CTA re-reads the w~ it wrote)
      acc = 0.0;
      // eight rows per thread in flight, loads unconditional (a row-length test in front of the loads serialised them: the
      // sweep ran at 40 % of the HBM rate, ncu long-scoreboard stalls, profiles/r2_ncu_csr5M.txt)
      for (uint32_t i0 = rlo + threadIdx.x; i0 < rhi; i0 += 8 * kBlock) {
        double wn[8], wc[8];
        uint32_t rl[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t i = min(i0 + u * kBlock, rhi - 1);
          wn[u] = __ldcg(Wn + i);
          wc[u] = __ldcg(Wc + i);
          rl[u] = __ldg(op.rlen + i);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t i = i0 + u * kBlock;
          if (i < rhi && rl[u] != 0xffu) {
            const double w = rec_sub(wn[u], alpha, __dmul_rn(wc[u], sc));
            __stcg(Wn + i, w);
            acc = fma(w, w, acc);
          }
        }
      }
      for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
        const uint32_t i = __ldg(lr.row + q);
        const double w = rec_sub(__ldcg(Wn + i), alpha, __dmul_rn(__ldcg(Wc + i), sc));
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

// =============================================================================================
// pass 2 (lanczos_pass_two_impl, src/algorithms/lanczos_two_pass.rs:206-312): one sweep and one grid barrier per step
// =============================================================================================
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_csr_kernel(const SellOp op, const Pass2Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  double* sm_seg = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int epoch = a.st->epoch;
  const LongRows& lr = op.lr;
  uint32_t slo, shi;
  cta_slices(op, slo, shi);
  const uint32_t rlo = slo * 32u, rhi = min(op.n, shi * 32u);
  const uint32_t r0 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x) : 0;
  const uint32_t r1 = lr.nlong ? __ldg(lr.cta_ptr + blockIdx.x + 1) : 0;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  int rot = 0;
  {
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
    for (uint32_t i = rlo + threadIdx.x; i < rhi; i += kBlock)
      if (__ldg(op.rlen + i) != 0xffu) init(i);
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
    auto finish = [&](uint32_t i, double t, double v, double vp, double xx) {
      const double w = rec_sub(rec_sub(t, bp, vp), alpha, v);
      const double vn = __dmul_rn(w, s);
      __stcg(Vn + i, vn);
      __stcg(a.x + i, __dadd_rn(xx, __dmul_rn(yj, vn)));
      if (WITH_V) __stcs(Vcol + i, vn);
    };
    SellHeads hd, hd_next;
    sell_heads(op, slo + warp, kWarps, shi, hd);
    for (uint32_t s0 = slo + warp; s0 < shi; s0 += kCsrU * kWarps) {
      SellBatch bt;
      double vc[kCsrU], vp[kCsrU], xx[kCsrU];
      sell_heads(op, s0 + kCsrU * kWarps, kWarps, shi, hd_next);
#pragma unroll
      for (int q = 0; q < kCsrU; ++q) {
        const uint32_t i = min((s0 + q * kWarps) * 32u + lane, op.n - 1);
        vc[q] = __ldcg(Vc + i);
        vp[q] = __ldcg(Vp + i);
        xx[q] = __ldcg(a.x + i);
      }
      sell_rows(op, s0, kWarps, hd, Vc, 1.0, bt);
      hd = hd_next;
#pragma unroll
      for (int q = 0; q < kCsrU; ++q)
        if (bt.live[q]) finish(bt.row[q], bt.t[q], vc[q], vp[q], xx[q]);
    }
    if (lr.nlong) {
      uint32_t q0, q1, sg0;
      sell_long_segments(lr, Vc, 1.0, sm_seg, q0, q1, sg0);
      __syncthreads();
      for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) {
        const uint32_t i = __ldg(lr.row + q);
        finish(i, long_row_total(lr, q, sm_seg, sg0), __ldcg(Vc + i), __ldcg(Vp + i), __ldcg(a.x + i));
      }
    }
    grid_sync<false>(0.0, a.gs, epoch, sh);
    rot = (rot + 1) % 3;
  }
}

// LinOp::apply
__global__ void __launch_bounds__(kBlock, 1) apply_csr_kernel(const SellOp op, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ double smem[];
  double* sm_seg = smem;
  const int warp = threadIdx.x >> 5;
  uint32_t slo, shi;
  cta_slices(op, slo, shi);
  SellHeads hd, hd_next;
  sell_heads(op, slo + warp, kWarps, shi, hd);
  for (uint32_t s0 = slo + warp; s0 < shi; s0 += kCsrU * kWarps) {
    SellBatch bt;
    sell_heads(op, s0 + kCsrU * kWarps, kWarps, shi, hd_next);
    sell_rows(op, s0, kWarps, hd, x, 1.0, bt);
    hd = hd_next;
#pragma unroll
    for (int q = 0; q < kCsrU; ++q)
      if (bt.live[q]) y[bt.row[q]] = bt.t[q];
  }
  const LongRows& lr = op.lr;
  if (lr.nlong) {
    uint32_t r0, r1, sg0;
    sell_long_segments(lr, x, 1.0, sm_seg, r0, r1, sg0);
    __syncthreads();
    for (uint32_t q = r0 + threadIdx.x; q < r1; q += kBlock) y[__ldg(lr.row + q)] = long_row_total(lr, q, sm_seg, sg0);
  }
}

}  // namespace tpl
