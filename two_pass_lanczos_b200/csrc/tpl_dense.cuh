// tpl_dense.cuh -- dense symmetric operator behind the same boundary: the `Mat<f64>` of
// src/bin/dense_tradeoff.rs:154-162 used as `&impl LinOp<f64>`.  A matvec streams the n x n matrix once (8 n^2 bytes,
// HBM bound); it is fused with the three-term recurrence exactly like the sparse kernels: one persistent cooperative
// kernel per pass, phase A (rows of A v, beta_{j-1} v_{j-1} subtracted, alpha partial) and phase B (alpha v subtracted,
// beta partial) separated by the payload-carrying grid barrier, pass 2 regenerating the basis with the stored coefficients
// and accumulating x in the same sweep.
// Rows are summed by a warp: lanes stride the columns (coalesced 256-byte reads of the row, the operand vector staged in
// shared memory when it fits), eight independent accumulators per lane combined in a fixed order, xor-shuffle tree.  The
// matrix is stored column-major as in faer; it is symmetric in every use of the reference (Lanczos requires it), so row i is
// read as column i.
#pragma once
#include "tpl_kernels.cuh"

namespace tpl {

struct DenseOp {
  uint32_t n;       // REAL dimension of the vectors (a Hermitian operator of n_c complex rows: n = 2 n_c)
  uint32_t stage;   // the operand vector (n doubles) is staged in shared memory
  uint32_t cplx;    // 1: complex Hermitian, entries (re, im) interleaved; vectors are n_c complex numbers, interleaved
  size_t lda;       // column stride in ENTRIES (doubles, or complex numbers)
  const double* a;  // column-major, symmetric / Hermitian
};
// Complex Hermitian operators (`T: ComplexField`, src/algorithms/mod.rs:167) ride on the real code: with real alpha and beta
// (T::Real) every vector operation of the recurrence -- <v, w> = Re(v^H w), w - alpha v, ||w||, the scaling by 1 / beta -- IS
// the real operation on the interleaved (re, im) storage; only the product A x is complex.  A "row unit" below is a row of
// the real operator or a complex row, which finishes two real rows (2 i: real part, 2 i + 1: imaginary part).

constexpr int kDenseUnroll = 8;

// (A x)_i with x_j = X[j] * s (s = 1: already normalised), or x staged (and scaled) in shared memory
__device__ __forceinline__ double dense_row(const DenseOp& op, uint32_t i, const double* X, double s, const double* sm_x, int lane) {
  const double* row = op.a + (size_t)i * op.lda;
  double acc[kDenseUnroll];
#pragma unroll
  for (int u = 0; u < kDenseUnroll; ++u) acc[u] = 0.0;
  for (uint32_t j0 = lane; j0 < op.n; j0 += 32 * kDenseUnroll) {
    double av[kDenseUnroll];
#pragma unroll
    for (int u = 0; u < kDenseUnroll; ++u) {
      const uint32_t j = j0 + 32 * u;
      av[u] = j < op.n ? __ldcs(row + j) : 0.0;  // streamed once per matvec
    }
#pragma unroll
    for (int u = 0; u < kDenseUnroll; ++u) {
      const uint32_t j = j0 + 32 * u;
      if (j < op.n) {
        const double xj = sm_x ? sm_x[j] : __dmul_rn(__ldcg(X + j), s);
        acc[u] = __dadd_rn(acc[u], __dmul_rn(av[u], xj));
      }
    }
  }
  double t = 0.0;
#pragma unroll
  for (int u = 0; u < kDenseUnroll; ++u) t = __dadd_rn(t, acc[u]);
  return warp_sum(t);
}

// (A x)_i of a Hermitian operator: row i is the conjugate of column i, so with c_j = a_ji (read coalesced, 16 bytes per entry)
//   Re y_i = sum_j Re c_j Re x_j + Im c_j Im x_j,   Im y_i = sum_j Re c_j Im x_j - Im c_j Re x_j.
__device__ __forceinline__ double2 dense_row_c(const DenseOp& op, uint32_t i, const double* X, double s, const double* sm_x, int lane) {
  constexpr int U = kDenseUnroll / 2;
  const double2* col = reinterpret_cast<const double2*>(op.a) + (size_t)i * op.lda;
  const uint32_t nc = op.n / 2;
  double ar[U], ai[U];
#pragma unroll
  for (int u = 0; u < U; ++u) ar[u] = ai[u] = 0.0;
  for (uint32_t j0 = lane; j0 < nc; j0 += 32 * U) {
    double2 c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t j = j0 + 32 * u;
      c[u] = j < nc ? __ldcs(col + j) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t j = j0 + 32 * u;
      if (j < nc) {
        double xr, xi;
        if (sm_x) {
          xr = sm_x[2 * j];
          xi = sm_x[2 * j + 1];
        } else {
          const double2 xv = __ldcg(reinterpret_cast<const double2*>(X) + j);
          xr = __dmul_rn(xv.x, s);
          xi = __dmul_rn(xv.y, s);
        }
        ar[u] = __dadd_rn(__dadd_rn(ar[u], __dmul_rn(c[u].x, xr)), __dmul_rn(c[u].y, xi));
        ai[u] = __dsub_rn(__dadd_rn(ai[u], __dmul_rn(c[u].x, xi)), __dmul_rn(c[u].y, xr));
      }
    }
  }
  double tr = 0.0, ti = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    tr = __dadd_rn(tr, ar[u]);
    ti = __dadd_rn(ti, ai[u]);
  }
  return make_double2(warp_sum(tr), warp_sum(ti));
}
// the rows a warp finishes for row unit i: `fin(row, t)` once (real) or twice (complex: real and imaginary part)
template <class FIN>
__device__ __forceinline__ void dense_unit(const DenseOp& op, uint32_t i, const double* X, double s, const double* sm_x, int lane, FIN fin) {
  if (op.cplx) {
    const double2 t = dense_row_c(op, i, X, s, sm_x, lane);
    if (lane == 0) {
      fin(2 * i, t.x);
      fin(2 * i + 1, t.y);
    }
  } else {
    const double t = dense_row(op, i, X, s, sm_x, lane);
    if (lane == 0) fin(i, t);
  }
}
__device__ __forceinline__ uint32_t dense_units(const DenseOp& op) { return op.cplx ? op.n / 2 : op.n; }

__device__ __forceinline__ const double* dense_stage(const DenseOp& op, const double* X, double s, double* sm) {
  if (!op.stage) return nullptr;
  for (uint32_t j = threadIdx.x; j < op.n; j += kBlock) sm[j] = __dmul_rn(__ldcg(X + j), s);
  return sm;
}

// Replaces lanczos_pass_one / the basis generation of lanczos_standard for a dense operator; steps [j_begin, j_end) per launch.
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_dense_kernel(const DenseOp op, const Pass1Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  const State st0 = *a.st;
  unsigned int epoch = st0.epoch;
  int rot = st0.rot, steps = st0.steps, status = st0.status;
  double sc = st0.s_cur, sp = st0.s_prev, bp = st0.beta_prev, bnorm = st0.b_norm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t lo, hi;  // row units of this CTA; rlo .. rhi the real rows they finish
  cta_chunk(dense_units(op), lo, hi);
  const uint32_t rlo = op.cplx ? 2 * lo : lo, rhi = op.cplx ? 2 * hi : hi;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };

  if (a.j_begin == 0) {
    double* Wp = pick(rot);
    double* Wc = pick((rot + 1) % 3);
    double acc = 0.0;
    for (uint32_t i = rlo + threadIdx.x; i < rhi; i += kBlock) {
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
      const double* sm_x = dense_stage(op, Wc, sc, smem);
      __syncthreads();
      double acc = 0.0;
      for (uint32_t i = lo + warp; i < hi; i += kWarps)
        dense_unit(op, i, Wc, sc, sm_x, lane, [&](uint32_t r, double t) {
          const double v = __dmul_rn(__ldcg(Wc + r), sc);
          const double wt = rec_sub(t, bp, __dmul_rn(__ldcg(Wp + r), sp));
          acc = fma(v, wt, acc);
          __stcg(Wn + r, wt);
          if (WITH_V) __stcs(Vcol + r, v);
        });
      const double alpha = grid_sync<true, false>(acc, a.gs, epoch, sh);  // all-reduce only: phase B reads this CTA's own rows
      // ---------------- phase B: w = w~ - alpha v, beta partial (same row ownership: lane 0 of the row's warp wrote w~)
      acc = 0.0;
      for (uint32_t i = lo + warp; i < hi; i += kWarps) {
        if (lane == 0)
          for (uint32_t r = op.cplx ? 2 * i : i; r < (op.cplx ? 2 * i + 2 : i + 1); ++r) {
            const double w = rec_sub(__ldcg(Wn + r), alpha, __dmul_rn(__ldcg(Wc + r), sc));
            __stcg(Wn + r, w);
            acc = fma(w, w, acc);
          }
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
      sc = 1.0 / beta;
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

// Replaces lanczos_pass_two_impl for a dense operator.
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_dense_kernel(const DenseOp op, const Pass2Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  unsigned int epoch = a.st->epoch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t lo, hi;
  cta_chunk(dense_units(op), lo, hi);
  const uint32_t rlo = op.cplx ? 2 * lo : lo, rhi = op.cplx ? 2 * hi : hi;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  int rot = 0;
  // The basis is regenerated in the LAZY representation of pass 1 -- un-normalised w in the buffers, the scale applied
  // at every use (the single rounding of the reference's in-place scaling) -- so that it is bit-identical to pass 1's.
  double sc = 1.0 / a.b_norm, sp = 1.0, bp = 0.0;
  {
    const double y0 = __ldg(a.y);
    for (uint32_t i = rlo + threadIdx.x; i < rhi; i += kBlock) {
      const double bi = __ldg(a.b + i);
      const double v = __dmul_rn(bi, sc);
      __stcg(buf1 + i, bi);
      __stcg(buf0 + i, 0.0);
      __stcg(a.x + i, __dmul_rn(v, y0));
      if (WITH_V) __stcs(a.V + i, v);
    }
    grid_sync<false>(0.0, a.gs, epoch, sh);
  }
  for (int j = 0; j + 1 < a.steps; ++j) {
    const double* Wp = pick(rot);
    const double* Wc = pick((rot + 1) % 3);
    double* Wn = pick((rot + 2) % 3);
    double* Vcol = WITH_V ? a.V + (size_t)(j + 1) * a.ldv : nullptr;
    const double alpha = __ldg(a.alphas + j);
    const double beta = __ldg(a.betas + j);
    const double sinv = 1.0 / beta;
    const double yj = __ldg(a.y + j + 1);
    const double* sm_x = dense_stage(op, Wc, sc, smem);
    __syncthreads();
    for (uint32_t i = lo + warp; i < hi; i += kWarps)
      dense_unit(op, i, Wc, sc, sm_x, lane, [&](uint32_t r, double t) {
        const double v = __dmul_rn(__ldcg(Wc + r), sc);
        const double w = rec_sub(rec_sub(t, bp, __dmul_rn(__ldcg(Wp + r), sp)), alpha, v);
        const double vn = __dmul_rn(w, sinv);
        __stcg(Wn + r, w);
        __stcg(a.x + r, __dadd_rn(__ldcg(a.x + r), __dmul_rn(yj, vn)));
        if (WITH_V) __stcs(Vcol + r, vn);
      });
    grid_sync<false>(0.0, a.gs, epoch, sh);
    sp = sc;
    sc = sinv;
    bp = beta;
    rot = (rot + 1) % 3;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.st->epoch = epoch;
}

// LinOp::apply
__global__ void __launch_bounds__(kBlock, 1) apply_dense_kernel(const DenseOp op, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* sm_x = dense_stage(op, x, 1.0, smem);
  __syncthreads();
  uint32_t lo, hi;
  cta_chunk(dense_units(op), lo, hi);
  for (uint32_t i = lo + warp; i < hi; i += kWarps) dense_unit(op, i, x, 1.0, sm_x, lane, [&](uint32_t r, double t) { y[r] = t; });
}

}  // namespace tpl
