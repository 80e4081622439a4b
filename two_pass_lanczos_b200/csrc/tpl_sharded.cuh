// tpl_sharded.cuh -- phase kernels of the arc-partitioned multi-GPU engine.
//
// Rank r owns the contiguous arc block [arc_begin, arc_end) -- rows of A, and the matching slices of every Lanczos
// vector, of d / tail / head -- plus a REPLICA of the p node entries of every vector.  A local vector is laid out
// [arc slice (m_r) | node part (p)], i.e. exactly like a vector of the local operator A_r = [[D_r, E_r^T], [E_r, 0]],
// whose node rows are this rank's PARTIAL node sums E_r x_arc.  A Lanczos step therefore is
//     phase A (local)    arc rows of w~ (complete: the node segment is replicated), partial node sums, partial alpha
//     all-reduce         p partial node sums + 1 partial alpha (one NCCL call, p + 1 doubles)
//     phase B (local)    node rows of w~ and alpha (every rank, every CTA: same values, same order), w = w~ - alpha v,
//                        partial beta^2 over the arcs
//     all-reduce         1 partial beta^2
// and the node part of beta^2, the breakdown test and the rotation happen at the head of the next phase-A launch.
// The kernels are split exactly where a cross-rank sum is needed; inside a kernel the G CTAs reduce with the same
// flag-carrying grid barrier as the single-GPU kernels.  Every per-element expression is shared with them
// (rec_sub / arc_row / node_entry / lazy scaling), pass 2 replays pass 1's node sums through the same all-reduce
// (same communicator, same count, same buffer layout), so the regenerated basis is bit-identical here too.
#pragma once
#include "tpl_kernels.cuh"

namespace tpl {

struct ShardArgs {
  double* buf[3];
  const double* b;
  double* alphas;  // pass 1: out, pass 2: in
  double* betas;
  const double* y;  // pass 2
  double* x;        // pass 2
  double* V;        // optional basis (one-pass / pass 2 with basis), local layout, column-major
  size_t ldv;
  double* red;      // [p + 1] node sums + alpha partial written by this launch (all-reduced after it)
  const double* red_in;  // pass 2: the all-reduced sums of the previous step (other half of the double buffer)
  double* red2;     // [1]     beta^2 / ||b||^2 partial over the arcs (all-reduced)
  State* st;
  GridSync gs;
  double tol;
  double b_norm;    // pass 2
  int j;            // step index
  int steps;        // pass 2: number of Lanczos steps of the decomposition
  int head_only;    // finish the previous step and return
};

// Deterministic CTA-wide sum: thread-strided partials, xor-shuffle tree per warp, then every thread adds the kWarps
// warp results in a fixed order.  Same code, same data => the same bits in every CTA of every rank.
__device__ __forceinline__ double cta_sum(double v, double* sm_warp) {
  v = warp_sum(v);
  __syncthreads();  // protects sm_warp against the previous use
  if ((threadIdx.x & 31) == 0) sm_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += sm_warp[w];
  return t;
}

// ||b||^2 over the local arcs; W_cur = b, W_prev = 0 (arc slice by chunks, node part by every CTA's share)
__global__ void __launch_bounds__(kBlock, 1) shard_init_kernel(const IncidenceOp op, const ShardArgs a) {
  __shared__ CtaShared sh;
  unsigned int epoch = a.st->epoch;
  uint32_t slo, shi;
  cta_chunk(op.m, slo, shi);
  double* Wp = a.buf[0];
  double* Wc = a.buf[1];
  double acc = 0.0;
  for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
    const double bi = __ldg(a.b + i);
    Wc[i] = bi;
    Wp[i] = 0.0;
    acc = fma(bi, bi, acc);
  }
  uint32_t ulo, uhi;
  cta_chunk(op.p, ulo, uhi);
  for (uint32_t u = ulo + threadIdx.x; u < uhi; u += kBlock) {
    Wc[op.m + u] = __ldg(a.b + op.m + u);
    Wp[op.m + u] = 0.0;
  }
  const double tot = grid_sync<true>(acc, a.gs, epoch, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.red2[0] = tot;
    State st = *a.st;
    st.epoch = epoch;
    st.rot = 0;
    st.steps = 0;
    st.status = ST_RUNNING;
    st.s_cur = 1.0;
    st.s_prev = 1.0;
    st.beta_prev = 0.0;
    st.b_norm = 0.0;
    *a.st = st;
  }
}

// Head: finishes step j-1 (or the norm of b for j == 0) from the all-reduced arc partial and the replicated node part.
// Body: phase A of step j on the local operator.
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) shard_phase_a_kernel(const IncidenceOp op, const ShardArgs a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  __shared__ double sm_warp[kWarps];
  double* sm_node = smem;
  double* sm_seg = smem + op.p;

  const State st0 = *a.st;
  if (st0.status != ST_RUNNING) return;
  unsigned int epoch = st0.epoch;
  int rot = st0.rot;
  double sc = st0.s_cur, sp = st0.s_prev, bp = st0.beta_prev, bnorm = st0.b_norm;
  int status = ST_RUNNING, steps = st0.steps;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  const uint32_t m = op.m, p = op.p;

  {
    // node part of the pending squared norm: b itself (j == 0) or the w of step j-1 (in the "next" buffer)
    const double* Wnode = (a.j == 0 ? pick((rot + 1) % 3) : pick((rot + 2) % 3)) + m;
    double acc = 0.0;
    for (uint32_t u = threadIdx.x; u < p; u += kBlock) {
      const double w = __ldcg(Wnode + u);
      acc = fma(w, w, acc);
    }
    const double sq = a.red2[0] + cta_sum(acc, sm_warp);
    const double nrm = sqrt(sq);
    if (a.j == 0) {
      bnorm = nrm;
      if (bnorm <= a.tol) status = ST_ZERO_B;
      sc = 1.0 / bnorm;
      sp = 1.0;
      bp = 0.0;
    } else {
      if (blockIdx.x == 0 && threadIdx.x == 0) a.betas[a.j - 1] = nrm;
      steps = a.j;
      if (nrm <= a.tol) {
        status = ST_BREAKDOWN;  // buffers are not rotated (mod.rs:331-338)
      } else {
        sp = sc;
        sc = 1.0 / nrm;
        bp = nrm;
        rot = (rot + 1) % 3;
      }
    }
  }
  auto save = [&]() {
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
  };
  if (status != ST_RUNNING || a.head_only) {
    save();
    return;
  }

  const double* Wp = pick(rot);
  const double* Wc = pick((rot + 1) % 3);
  double* Wn = pick((rot + 2) % 3);
  double* Vcol = WITH_V ? a.V + (size_t)a.j * a.ldv : nullptr;
  const LongRows& lr = op.lr;
  uint32_t slo, shi;
  cta_chunk(m, slo, shi);
  for (uint32_t u = threadIdx.x; u < p; u += kBlock) sm_node[u] = __dmul_rn(__ldcg(Wc + m + u), sc);
  __syncthreads();
  const IncidenceDev dev{op, sm_node};
  double acc = 0.0;
#pragma unroll 2
  for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
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
    for (uint32_t q = q0 + threadIdx.x; q < q1; q += kBlock) {
      const uint32_t u = __ldg(lr.row + q) - m;
      a.red[u] = long_row_total(lr, q, sm_seg, s0);  // partial node sum of this rank
      if (WITH_V) __stcs(Vcol + m + u, sm_node[u]);
    }
  }
  const double alpha_arc = grid_sync<true>(acc, a.gs, epoch, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.red[p] = alpha_arc;
  save();
}

// Node rows of w~ and alpha from the all-reduced sums (redundantly in every CTA), then phase B over the arcs.
__global__ void __launch_bounds__(kBlock, 1) shard_phase_b_kernel(const IncidenceOp op, const ShardArgs a) {
  __shared__ CtaShared sh;
  __shared__ double sm_warp[kWarps];
  const State st0 = *a.st;
  if (st0.status != ST_RUNNING) return;
  unsigned int epoch = st0.epoch;
  const int rot = st0.rot;
  const double sc = st0.s_cur, sp = st0.s_prev, bp = st0.beta_prev;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  const double* Wp = pick(rot);
  const double* Wc = pick((rot + 1) % 3);
  double* Wn = pick((rot + 2) % 3);
  const uint32_t m = op.m, p = op.p;

  double acc = 0.0;
  for (uint32_t u = threadIdx.x; u < p; u += kBlock) {
    const double v = __dmul_rn(__ldcg(Wc + m + u), sc);
    const double vp = __dmul_rn(__ldcg(Wp + m + u), sp);
    const double wt = rec_sub(__ldcg(a.red + u), bp, vp);
    acc = fma(v, wt, acc);
  }
  const double alpha = __ldcg(a.red + p) + cta_sum(acc, sm_warp);

  uint32_t ulo, uhi;
  cta_chunk(p, ulo, uhi);
  for (uint32_t u = ulo + threadIdx.x; u < uhi; u += kBlock) {
    const double v = __dmul_rn(__ldcg(Wc + m + u), sc);
    const double vp = __dmul_rn(__ldcg(Wp + m + u), sp);
    const double wt = rec_sub(__ldcg(a.red + u), bp, vp);
    __stcg(Wn + m + u, rec_sub(wt, alpha, v));
  }
  uint32_t slo, shi;
  cta_chunk(m, slo, shi);
  acc = 0.0;
#pragma unroll 4
  for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
    const double v = __dmul_rn(__ldcg(Wc + i), sc);
    const double w = rec_sub(__ldcg(Wn + i), alpha, v);
    __stcg(Wn + i, w);
    acc = fma(w, w, acc);
  }
  const double beta2_arc = grid_sync<true>(acc, a.gs, epoch, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.red2[0] = beta2_arc;
    a.alphas[a.j] = alpha;
    a.st->epoch = epoch;
  }
}

// Pass 2.  j < 0: v_1 = b / ||b||, x = y_0 v_1.  j >= 0: head finishes the node part of step j-1 from the all-reduced
// node sums (every CTA builds the whole node segment in shared memory, the owners also store it), body regenerates the
// arc part of v_{j+2} (0-based step j) and the partial node sums.
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) shard_pass2_kernel(const IncidenceOp op, const ShardArgs a) {
  extern __shared__ double smem[];
  double* sm_node = smem;
  double* sm_seg = smem + op.p;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  const uint32_t m = op.m, p = op.p;
  uint32_t slo, shi, ulo, uhi;
  cta_chunk(m, slo, shi);
  cta_chunk(p, ulo, uhi);
  if (a.j < 0) {
    const double inv = 1.0 / a.b_norm;
    const double y0 = __ldg(a.y);
    double* Vp = buf0;
    double* Vc = buf1;
    auto init = [&](uint32_t i) {
      const double v = __dmul_rn(__ldg(a.b + i), inv);
      Vc[i] = v;
      Vp[i] = 0.0;
      a.x[i] = __dmul_rn(v, y0);
      if (WITH_V) __stcs(a.V + i, v);
    };
    for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) init(i);
    for (uint32_t u = ulo + threadIdx.x; u < uhi; u += kBlock) init(m + u);
    return;
  }
  const int j = a.j;
  int rot = j % 3;  // rotation of step j: prev = pick(rot), cur = pick(rot + 1), next = pick(rot + 2)
  if (j == 0) {
    for (uint32_t u = threadIdx.x; u < p; u += kBlock) sm_node[u] = __ldcg(pick(1) + m + u);
  } else {
    // node part of step j-1: v_{j+1} = ((t - beta_{j-2} v_{j-1}) - alpha_{j-1} v_j) * (1 / beta_{j-1})
    const int rp = (j - 1) % 3;
    const double* Vp = pick(rp);
    const double* Vc = pick((rp + 1) % 3);
    double* Vn = pick((rp + 2) % 3);
    const double alpha = __ldg(a.alphas + j - 1);
    const double beta = __ldg(a.betas + j - 1);
    const double bp = j == 1 ? 0.0 : __ldg(a.betas + j - 2);
    const double s = 1.0 / beta;
    const double yj = __ldg(a.y + j);
    double* Vcol = WITH_V ? a.V + (size_t)j * a.ldv : nullptr;
    for (uint32_t u = threadIdx.x; u < p; u += kBlock) {
      const double w = rec_sub(rec_sub(__ldcg(a.red_in + u), bp, __ldcg(Vp + m + u)), alpha, __ldcg(Vc + m + u));
      const double vn = __dmul_rn(w, s);
      sm_node[u] = vn;
      if (u >= ulo && u < uhi) {
        __stcg(Vn + m + u, vn);
        a.x[m + u] = __dadd_rn(a.x[m + u], __dmul_rn(yj, vn));
        if (WITH_V) __stcs(Vcol + m + u, vn);
      }
    }
  }
  if (a.head_only) return;
  __syncthreads();
  const double* Vp = pick(rot);
  const double* Vc = pick((rot + 1) % 3);
  double* Vn = pick((rot + 2) % 3);
  double* Vcol = WITH_V ? a.V + (size_t)(j + 1) * a.ldv : nullptr;
  const double alpha = __ldg(a.alphas + j);
  const double beta = __ldg(a.betas + j);
  const double bp = j == 0 ? 0.0 : __ldg(a.betas + j - 1);
  const double s = 1.0 / beta;
  const double yj = __ldg(a.y + j + 1);
  const IncidenceDev dev{op, sm_node};
  const LongRows& lr = op.lr;
#pragma unroll 2
  for (uint32_t i = slo + threadIdx.x; i < shi; i += kBlock) {
    const double v = __ldcg(Vc + i);
    const double w = rec_sub(rec_sub(dev.short_row(i, v, Vc, 1.0), bp, __ldcg(Vp + i)), alpha, v);
    const double vn = __dmul_rn(w, s);
    __stcg(Vn + i, vn);
    __stcg(a.x + i, __dadd_rn(__ldcg(a.x + i), __dmul_rn(yj, vn)));
    if (WITH_V) __stcs(Vcol + i, vn);
  }
  if (lr.nlong) {
    uint32_t q0, q1, s0;
    long_row_segments(dev, lr, Vc, 1.0, sm_seg, q0, q1, s0);
    __syncthreads();
    for (uint32_t q = q0 + threadIdx.x; q < q1; q += kBlock)
      a.red[__ldg(lr.row + q) - m] = long_row_total(lr, q, sm_seg, s0);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.red[p] = 0.0;  // same all-reduce count and layout as pass 1
}

}  // namespace tpl
