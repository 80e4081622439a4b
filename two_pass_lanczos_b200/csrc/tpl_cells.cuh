// tpl_cells.cuh -- shared-memory / register resident kernels of the KKT incidence operator on a 2-D ("cell") partition
// of the arcs.
//
// The resident kernels of tpl_kernels.cuh cut the arcs into G contiguous chunks: every CTA then needs the whole node
// segment (p values) and contributes to every node sum, so a Lanczos step moves G*p partial sums and G*p node values
// through L2 and is bound by the latency of two (pass 1) or one (pass 2) barrier-fenced exchanges of that size.
// Here the arcs are partitioned like a 2-D SpMV: the nodes are cut into GR contiguous TAIL blocks (balanced by
// out-degree) and GC contiguous HEAD blocks (balanced by in-degree); CTA (a, b) owns the arcs with tail in block a and
// head in block b.  It only needs the node values of those two blocks (~p/GR + p/GC instead of p) and only contributes
// to their sums.  Nodes are grouped in LINES of 8 (one 128-byte line of eight 16-byte LL atoms {lo, tag, hi, tag});
// line l is owned by CTA l % Gc, which adds the contributions in a fixed order, runs the recurrence of its 8 node rows
// and publishes their new values.  All exchanged data is self-validating (tagged atoms, NCCL-LL style), so there is
// no fence and no barrier other than the alpha / beta all-reduces the algorithm itself needs:
//     inbox [2][slots][8]  partial node sums pushed by the contributing CTAs (tail side +, head side already negated)
//     gather[2][L][8]      node values published by the owners, pulled by the CTAs that touch the line
//     ar    [2][Gc][8]     one line per CTA: all-to-all pull all-reduce, fixed summation order
// Every line is written as a whole by 8 adjacent lanes in one store instruction: 16-byte stores that leave a 32-byte
// sector partially written take a slow path in L2 (measured: the same exchange with atom-at-a-time stores was 2-3x slower).
// A thread keeps its 8 arcs (D, tail / head, both vectors, x) in registers for the whole pass.  The arcs of a cell are
// stored in jagged-diagonal order of the tail lists (row e holds the e-th arc of every tail that has one, tails sorted
// by decreasing count), so that the lanes that sum 16 tail lists side by side read consecutive shared-memory words.
// Pass 1 per step: poll {beta, node values, node sums} -> arc rows + node rows -> alpha all-reduce -> w = w~ - alpha v,
// node sums of w pushed, node values and the beta partial published.  Pass 2 per step: poll -> rows -> push/publish.
// Pass 1, the one-pass variant and pass 2 share every expression and every summation order: the regenerated basis is
// bit-identical to the stored one (results/orthogonality_*.csv: drift == 0).
#pragma once
#include "tpl_kernels.cuh"

namespace tpl {

constexpr uint32_t kLine = 8;        // nodes per line = atoms per 128-byte line
constexpr int kArcRegs = 9;          // arcs per worker thread; the arc part of the vectors and of x lives in REGISTERS for a whole pass
constexpr uint32_t kWorkers = kBlock - 32;         // warps 0..14 work on the arcs; warp 15 is the OWNER warp (node rows of the owned lines)
constexpr uint32_t kWorkerWarps = kWorkers / 32;
constexpr uint32_t kCellArcs = kArcRegs * kWorkers;  // arc slots of a cell (4320)

struct CellOp {
  uint32_t GR, GC, Gc;   // cell grid; Gc = GR * GC CTAs
  uint32_t Amax;         // arc slots per cell
  uint32_t L;            // node lines
  uint32_t max_lines;    // most touched lines of any cell
  uint32_t max_slots;    // most pushed lines (inbox slots written) of any cell
  uint32_t max_groups;   // most node-sum groups (16 lists, one warp) of any cell
  uint32_t max_rows;     // most entry rows (32 lanes x 4 entries) of any cell
  uint32_t max_own;      // owned lines per CTA = ceil(L / Gc)
  uint32_t max_tl;       // most tail lists of any cell
  uint32_t inbox_atoms;  // atoms per inbox parity
  uint32_t flags;        // experiment switches (TPL_CELL_FLAGS): 1 serial inbox polls, 2 back-off in spin loops
  const uint32_t* hdr;        // [Gc][8] {arcs, lines, slots, groups, entry rows, 0, 0, 0}
  const uint32_t* gidx;       // [Gc][Amax] arc index in the caller's order
  const uint32_t* lth;        // [Gc][Amax] tail slot | tail-first << 15 | head slot << 16: indices into the cell's node-value array
  const uint16_t* tmap;       // [Gc][8 * max_lines] tail slot a polled node value is mirrored into (0xffff: none)
  const uint32_t* lines;      // [Gc][max_lines] touched lines, ascending (local node = 8 * position + node % 8)
  const uint32_t* push;       // [Gc][max_slots] inbox slot of every pushed line
  // node sums: group g (one warp) sums 16 lists, lanes l and l + 16 taking the even / odd entries of list l:
  const uint4* walk;          // [Gc][max_groups * 32] per lane {index into the cell's sums or ~0, sign mask, first row, rows}
  const uint2* ent4;          // [Gc][max_rows * 32]   per lane and row four 16-bit arc positions (Amax = the zero slot)
  const uint32_t* slot_base;  // [L + 1] first inbox slot of every line
  uint4* inbox;
  uint4* gather;
  uint4* ar;
};

struct CellSmem {
  double* w;                    // [Amax + 8] arc part of the newest vector (what the node sums are formed from); w[Amax] = 0
  double* d;                    // [Amax] D of the cell's arcs
  // node values of the current vector: [8 * max_lines by local node | max_tl by tail rank (the arcs of a half-warp have
  // consecutive ranks: conflict-free) | zero slot (self-loops, empty arc slots)]
  double* nodev;
  uint16_t* tmap;               // [8 * max_lines]
  double* sums;                 // [8 * max_slots] node sums of this cell in push order
  double *n0, *n1, *nx, *T;     // [8 * max_own] owned node rows: current / previous vector, x (pass 2), node sums
  double* arv;                  // [Gc] all-reduce values by slot
  double* wpart;                // [kWarps]
  uint4* walk;                  // [max_groups * 32]
  uint2* ent4;                  // [max_rows * 32]
  uint32_t* lines;              // [max_lines]
  uint32_t* push;               // [max_slots]
  uint32_t* own;                // [2 * max_own] {first slot, slots} of the owned lines
};

__host__ __device__ inline size_t cell_smem_bytes(const CellOp& co, bool pass2) {
  size_t dbl = 2 * (size_t)co.Amax + 8 + kLine * co.max_lines + co.max_tl + 8 + kLine * co.max_slots + (size_t)kLine * co.max_own * (pass2 ? 4 : 3) +
               co.Gc + kWarps;
  dbl = (dbl + 1) & ~(size_t)1;
  const size_t u32 = (size_t)co.max_lines + co.max_slots + 2 * (size_t)co.max_own;
  return dbl * 8 + (size_t)co.max_groups * 32 * 16 + (size_t)co.max_rows * 32 * 8 + u32 * 4 + (size_t)kLine * co.max_lines * 2 + 16;
}

template <bool PASS2>
__device__ __forceinline__ CellSmem carve_cell(double* base, const CellOp& co) {
  CellSmem s;
  double* d = base;
  s.w = d; d += co.Amax + 8;
  s.d = d; d += co.Amax;
  s.nodev = d; d += kLine * co.max_lines + co.max_tl + 8;
  s.sums = d; d += kLine * co.max_slots;
  s.n0 = d; d += kLine * co.max_own;
  s.n1 = d; d += kLine * co.max_own;
  s.nx = d; d += PASS2 ? kLine * co.max_own : 0;
  s.T = d; d += kLine * co.max_own;
  s.arv = d; d += co.Gc;
  s.wpart = d; d += kWarps;
  size_t off = (size_t)(d - base);
  off = (off + 1) & ~(size_t)1;  // 16-byte alignment
  s.walk = reinterpret_cast<uint4*>(base + off);
  s.ent4 = reinterpret_cast<uint2*>(s.walk + (size_t)co.max_groups * 32);
  uint32_t* u = reinterpret_cast<uint32_t*>(s.ent4 + (size_t)co.max_rows * 32);
  s.lines = u; u += co.max_lines;
  s.push = u; u += co.max_slots;
  s.own = u; u += 2 * co.max_own;
  s.tmap = reinterpret_cast<uint16_t*>(u);
  return s;
}

struct CellCtx {
  uint32_t nA, nlines, nslots, ngroups, nown;
  const uint32_t* gidx;
};

// Loads the cell's tables into shared memory (once per kernel).  Caller syncs.
__device__ __forceinline__ CellCtx load_cell(const CellOp& co, const CellSmem& s) {
  const uint32_t c = blockIdx.x;
  const uint32_t* h = co.hdr + (size_t)c * 8;
  CellCtx x;
  x.nA = __ldg(h);
  x.nlines = __ldg(h + 1);
  x.nslots = __ldg(h + 2);
  x.ngroups = __ldg(h + 3);
  const uint32_t nrows = __ldg(h + 4);
  x.nown = co.L > c ? (co.L - c + co.Gc - 1) / co.Gc : 0;
  x.gidx = co.gidx + (size_t)c * co.Amax;
  const uint32_t* ln = co.lines + (size_t)c * co.max_lines;
  for (uint32_t i = threadIdx.x; i < x.nlines; i += kBlock) s.lines[i] = __ldg(ln + i);
  const uint16_t* tm = co.tmap + (size_t)c * kLine * co.max_lines;
  for (uint32_t i = threadIdx.x; i < x.nlines * kLine; i += kBlock) s.tmap[i] = __ldg(tm + i);
  for (uint32_t i = threadIdx.x; i < kLine * co.max_lines + co.max_tl + 8; i += kBlock) s.nodev[i] = 0.0;
  const uint32_t* ps = co.push + (size_t)c * co.max_slots;
  for (uint32_t i = threadIdx.x; i < x.nslots; i += kBlock) s.push[i] = __ldg(ps + i);
  const uint4* wk = co.walk + (size_t)c * co.max_groups * 32;
  for (uint32_t i = threadIdx.x; i < x.ngroups * 32; i += kBlock) s.walk[i] = __ldg(wk + i);
  const uint2* e4 = co.ent4 + (size_t)c * co.max_rows * 32;
  for (uint32_t i = threadIdx.x; i < nrows * 32; i += kBlock) s.ent4[i] = __ldg(e4 + i);
  for (uint32_t i = threadIdx.x; i < x.nslots * kLine; i += kBlock) s.sums[i] = 0.0;  // atoms without a list stay zero
  for (uint32_t i = threadIdx.x; i < 8; i += kBlock) s.w[co.Amax + i] = 0.0;          // the zero slot padding entries point at
  for (uint32_t o = threadIdx.x; o < x.nown; o += kBlock) {
    const uint32_t l = c + o * co.Gc;
    const uint32_t b0 = __ldg(co.slot_base + l);
    s.own[2 * o] = b0;
    s.own[2 * o + 1] = __ldg(co.slot_base + l + 1) - b0;
  }
  return x;
}

// ----------------------------------------------------------------------------- tagged atoms
__device__ __forceinline__ uint4 atom_pack(double v, uint32_t tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return make_uint4((unsigned)b, tag, (unsigned)(b >> 32), tag);
}
__device__ __forceinline__ double atom_value(uint4 f) {
  return __longlong_as_double((long long)(((unsigned long long)f.z << 32) | f.x));
}
__device__ __forceinline__ uint4 ld_cg_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// polling load: flags 8 = ld.global.cg, 16 = ld.volatile, else ld.relaxed.gpu
__device__ __forceinline__ uint4 ld_atom(const uint4* p, uint32_t flags) {
  if (flags & 8u) return ld_cg_v4(p);
  if (flags & 16u) return ld_volatile_v4(p);
  return ld_relaxed_gpu_v4(p);
}
__device__ __forceinline__ double atom_poll(const uint4* p, uint32_t tag, uint32_t flags = 0) {
  uint4 f = ld_atom(p, flags);
  uint32_t spins = 0;
  while (f.y != tag || f.w != tag) {
    if (++spins > kSpinLimit) __trap();
    if (flags & 2u) __nanosleep(100);
    f = ld_atom(p, flags);
  }
  return atom_value(f);
}

// ----------------------------------------------------------------------------- named barriers
// barrier 0 (__syncthreads): all 16 warps; barrier 1: the 15 worker warps; barrier 2: the workers ARRIVE (without waiting)
// once the all-reduce values of the top of a step are in shared memory, the owner warp waits for them.
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory"); }
__device__ __forceinline__ void bar_arrive_top() { asm volatile("bar.arrive 2, %0;" ::"n"(kBlock) : "memory"); }
__device__ __forceinline__ void bar_wait_top() { asm volatile("bar.sync 2, %0;" ::"n"(kBlock) : "memory"); }

// ----------------------------------------------------------------------------- exchange steps
// Node sums of the arc values s.w of this cell -> s.sums.  One warp per group of 16 lists; lane l (< 16) adds the even
// entries of list l in order, lane l + 16 the odd ones, then the two halves are added.  A row of entries is one 8-byte
// word per lane (four 16-bit positions), rows of a group are 32 words apart: every load below is conflict-free except the
// gathers of the head lists.  The sign (tail +, head -) is applied by flipping the sign bit (a - x == a + (-x) exactly).
// Caller syncs before cell_push_lines.
__device__ __forceinline__ void cell_node_sums(const CellSmem& s, const CellCtx& c) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t g = warp; g < c.ngroups; g += kWorkerWarps) {
    const uint4 d = s.walk[g * 32 + lane];
    const uint2* row = s.ent4 + (size_t)d.z * 32 + lane;
    double acc = 0.0;
    uint2 e = d.w ? row[0] : make_uint2(0, 0);
    for (uint32_t k = 0; k < d.w; ++k) {
      const uint2 nxt = k + 1 < d.w ? row[(k + 1) * 32] : e;  // the next row is requested before this one is consumed
      const double v0 = s.w[e.x & 0xffffu], v1 = s.w[e.x >> 16], v2 = s.w[e.y & 0xffffu], v3 = s.w[e.y >> 16];
      acc = __dadd_rn(acc, __hiloint2double(__double2hiint(v0) ^ (int)d.y, __double2loint(v0)));
      acc = __dadd_rn(acc, __hiloint2double(__double2hiint(v1) ^ (int)d.y, __double2loint(v1)));
      acc = __dadd_rn(acc, __hiloint2double(__double2hiint(v2) ^ (int)d.y, __double2loint(v2)));
      acc = __dadd_rn(acc, __hiloint2double(__double2hiint(v3) ^ (int)d.y, __double2loint(v3)));
      e = nxt;
    }
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
    if (lane < 16 && d.x != 0xffffffffu) s.sums[d.x] = acc;
  }
}
__device__ __forceinline__ void st_volatile_v4(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// s.sums -> the owners' inboxes as generation `gen`, one whole line per 8 adjacent lanes.
__device__ __forceinline__ void cell_push_lines(const CellOp& co, const CellSmem& s, const CellCtx& c, uint32_t gen) {
  uint4* box = co.inbox + (size_t)(gen & 1u) * co.inbox_atoms;
  const uint32_t tag = gen + 1;
  for (uint32_t t = threadIdx.x; t < c.nslots * kLine; t += kWorkers) {
    uint4* dst = box + (size_t)s.push[t >> 3] * kLine + (t & 7);
    if (co.flags & 4u) st_volatile_v4(dst, atom_pack(s.sums[t], tag));
    else st_relaxed_gpu_v4(dst, atom_pack(s.sums[t], tag));
  }
}

// Owner side: the sum of the contributions to node r = lane & 7 of owned line o, generation `gen` (same value in the
// four lanes of a node).  Lane = r + 8 q adds slots q, q+4, ... in order; the four partial sums are combined by a fixed xor tree.
__device__ __forceinline__ double cell_poll_inbox(const CellOp& co, const CellSmem& s, uint32_t o, uint32_t gen) {
  const uint4* box = co.inbox + (size_t)(gen & 1u) * co.inbox_atoms;
  const uint32_t tag = gen + 1;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t r = lane & 7, q = lane >> 3;
  const uint32_t b0 = s.own[2 * o], K = s.own[2 * o + 1];
  const uint4* mine = box + (size_t)b0 * kLine + r;
  double t = 0.0;
  for (uint32_t k0 = 0; k0 < K; k0 += 32) {
    // all of a lane's atoms are requested together and re-requested together until every one carries the tag: one L2
    // round trip per attempt, not one per atom
    uint4 f[8];
    uint32_t spins = 0;
    bool pending = true;
    while (pending) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t k = k0 + q + 4 * i;
        if (k < K) f[i] = ld_atom(mine + (size_t)k * kLine, co.flags);
      }
      pending = false;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t k = k0 + q + 4 * i;
        if (k < K) pending |= (f[i].y != tag) | (f[i].w != tag);
      }
      if (++spins > kSpinLimit) __trap();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t k = k0 + q + 4 * i;
      if (k < K) t = __dadd_rn(t, atom_value(f[i]));
    }
  }
  __syncwarp();
  t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 8));
  t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 16));
  return t;
}

// Node values of the touched lines, generation `gen`, times `scale` -> nodev.
__device__ __forceinline__ void cell_poll_gather(const CellOp& co, const CellSmem& s, const CellCtx& c, uint32_t gen,
                                                 uint32_t warp0, uint32_t nwarps, double scale) {
  const uint4* g = co.gather + (size_t)(gen & 1u) * co.L * kLine;
  const uint32_t tag = gen + 1;
  for (uint32_t a = threadIdx.x - warp0 * 32; a < c.nlines * kLine; a += nwarps * 32) {
    const double v = __dmul_rn(atom_poll(g + (size_t)s.lines[a >> 3] * kLine + (a & 7), tag, co.flags), scale);
    const uint32_t mirror = s.tmap[a];
    s.nodev[a] = v;
    if (mirror != 0xffffu) s.nodev[mirror] = v;
  }
}

__device__ __forceinline__ void cell_publish_node(const CellOp& co, uint32_t line, uint32_t r, double v, uint32_t gen) {
  st_relaxed_gpu_v4(co.gather + ((size_t)(gen & 1u) * co.L + line) * kLine + r, atom_pack(v, gen + 1));
}

// All-reduce, publishing half: the CTA's partial (already summed over the warps into wpart, caller synced) goes out as
// one whole line (eight copies of the atom, lanes 0..7).
__device__ __forceinline__ void cell_ar_publish(const CellOp& co, const CellSmem& s, uint32_t epoch) {
  if (threadIdx.x >= kBlock - 32) {  // the last warp: the first ones own the longest node-sum lists
    const uint32_t lane = threadIdx.x & 31;
    double t = lane < kWarps ? s.wpart[lane] : 0.0;
    t = warp_sum(t);
    if (lane < kLine) st_relaxed_gpu_v4(co.ar + ((size_t)(epoch & 1u) * co.Gc + blockIdx.x) * kLine + lane, atom_pack(t, epoch));
  }
}
// All-reduce, polling half: thread t of the first ceil(Gc / 32) warps polls the line of CTA (t + cta) % Gc (rotated so
// that the CTAs do not all start on the same line).  Caller syncs, then every warp calls cell_ar_total.
__device__ __forceinline__ void cell_ar_poll(const CellOp& co, const CellSmem& s, uint32_t epoch) {
  if (threadIdx.x < co.Gc) {
    uint32_t slot = threadIdx.x + blockIdx.x;
    slot = slot >= co.Gc ? slot - co.Gc : slot;
    s.arv[slot] = atom_poll(co.ar + ((size_t)(epoch & 1u) * co.Gc + slot) * kLine + (blockIdx.x & 7u), epoch, co.flags);
  }
}
__device__ __forceinline__ double cell_ar_total(const CellOp& co, const CellSmem& s) {
  double t = 0.0;
  for (uint32_t i = threadIdx.x & 31; i < co.Gc; i += 32) t = __dadd_rn(t, s.arv[i]);
  return warp_sum(t);
}
__device__ __forceinline__ void cell_block_partial(const CellSmem& s, double acc) {
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s.wpart[threadIdx.x >> 5] = acc;
}

// The arcs of a cell are spread over the worker threads, arc r of thread t sitting at position t + r * kWorkers of the
// cell's (jagged-diagonal) order.  A thread keeps its arcs' tail / head, the current vector W (un-normalised, v = W * sc),
// the previous vector v_{j-1} (normalised) and, in pass 2, x in registers for the whole pass; D stays in shared memory.
template <bool PASS2>
struct ArcRegs {
  double W[kArcRegs], P[kArcRegs], X[PASS2 ? kArcRegs : 1];
  uint32_t TH[kArcRegs];
};

// W = b, previous vector = 0; also writes b to s.w for the first node sums.  Returns the partial of ||b||^2.
template <bool PASS2>
__device__ __forceinline__ double cell_load_arcs(const IncidenceOp& op, const CellOp& co, const CellSmem& s, const CellCtx& c,
                                                 const double* b, ArcRegs<PASS2>& R) {
  const uint32_t* lth = co.lth + (size_t)blockIdx.x * co.Amax;
  const uint32_t zero_slot = kLine * co.max_lines + co.max_tl;
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < kArcRegs; ++r) {
    const uint32_t i = threadIdx.x + r * kWorkers;
    R.W[r] = 0.0;
    R.P[r] = 0.0;
    R.TH[r] = zero_slot | (zero_slot << 16);
    if (i < c.nA) {
      const uint32_t g = __ldg(c.gidx + i);
      R.W[r] = __ldg(b + g);
      R.TH[r] = __ldg(lth + i);
      s.d[i] = __ldg(op.d + g);
      s.w[i] = R.W[r];
      acc = fma(R.W[r], R.W[r], acc);
    }
  }
  return acc;
}

// Arc rows of one step: v = W sc replaces the previous vector (the same single rounding the reference performs when it
// scales w in place, mod.rs:312-315) and W becomes
//   pass 1 (PASS2 = false): w~ = A v - beta_{j-1} v_{j-1}; returns the partial of alpha = <v, w~>
//   pass 2 (PASS2 = true):  w  = w~ - alpha v, written to s.w, and x += y_{j+1} (w / beta_j)
template <bool PASS2, bool WITH_V>
__device__ __forceinline__ double cell_arc_rows(const CellSmem& s, const CellCtx& c, ArcRegs<PASS2>& R, double sc, double bp,
                                                double alpha, double sinv, double yj, double* Vcol) {
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < kArcRegs; ++r) {
    if (r * kWorkers < c.nA) {  // uniform: rows of arc slots beyond the cell's last arc are skipped
      const uint32_t i = threadIdx.x + r * kWorkers;
      const uint32_t ii = i < c.nA ? i : 0;
      const uint32_t t = R.TH[r] & 0x7fffu, h = R.TH[r] >> 16;
      const double v = __dmul_rn(R.W[r], sc);
      // pass 2 scales the node values once when it polls them; pass 1 only learns sc together with them
      const double xt = PASS2 ? s.nodev[t] : __dmul_rn(s.nodev[t], sc), xh = PASS2 ? s.nodev[h] : __dmul_rn(s.nodev[h], sc);
      // (A v)_arc in the reference's CSC accumulation order: D v first, then the two node columns by ascending node index
      // (bit 15: the tail comes first); a self-loop reads the zero slot twice
      double row = __dmul_rn(s.d[ii], v);
      if (R.TH[r] & 0x8000u) {
        row = __dadd_rn(row, xt);
        row = __dsub_rn(row, xh);
      } else {
        row = __dsub_rn(row, xh);
        row = __dadd_rn(row, xt);
      }
      const double wt = rec_sub(row, bp, R.P[r]);
      R.P[r] = v;
      if (PASS2) {
        const double w = rec_sub(wt, alpha, v);
        const double vn = __dmul_rn(w, sinv);
        R.W[r] = w;
        R.X[r] = __dadd_rn(R.X[r], __dmul_rn(yj, vn));
        if (i < c.nA) {
          s.w[i] = w;
          if (WITH_V) __stcs(Vcol + __ldg(c.gidx + i), vn);
        }
      } else {
        acc = fma(v, wt, acc);
        R.W[r] = wt;
        if (WITH_V && i < c.nA) __stcs(Vcol + __ldg(c.gidx + i), v);
      }
    }
  }
  return acc;
}

// Replaces lanczos_pass_one (src/algorithms/lanczos_two_pass.rs:65-110) and, with WITH_V, the basis generation of
// lanczos_standard (src/algorithms/lanczos.rs:55-156) for a cell-partitioned incidence operator; whole pass per launch.
// Warps 0..14 hold the arcs; warp 15 owns the node rows of the CTA's lines: it waits for their node sums while the
// workers are busy with the arc rows, so that the exchange started at the end of a step is off the critical path.
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_cell_kernel(const IncidenceOp op, const CellOp co, const Pass1Args a) {
  extern __shared__ double smem[];
  const CellSmem s = carve_cell<false>(smem, co);
  const CellCtx c = load_cell(co, s);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool owner = warp == kWorkerWarps;
  const uint32_t ar_warps = (co.Gc + 31) / 32;  // polling roles of the worker warps at the top of a step
  const Trace& tr = a.gs.trace;
  ArcRegs<false> R;

  uint32_t epoch = 0;
  int steps = 0, status = ST_RUNNING;
  double sc = 1.0, bp = 0.0, bnorm = 0.0;
  {
    // K0: ||b||; the cell's arcs / the owned node rows of b become the current vector, the previous one is zero; node
    // values and node sums of b are published as generation 0
    double acc = 0.0;
    if (!owner) {
      acc = cell_load_arcs<false>(op, co, s, c, a.b, R);
    } else {
      for (uint32_t t = lane; t < c.nown * kLine; t += 32) {
        const uint32_t line = blockIdx.x + (t >> 3) * co.Gc, u = line * kLine + (t & 7);
        const double bi = u < op.p ? __ldg(a.b + op.m + u) : 0.0;
        s.n0[t] = bi;
        s.n1[t] = 0.0;
        acc = fma(bi, bi, acc);
        cell_publish_node(co, line, t & 7, bi, 0);
      }
    }
    cell_block_partial(s, acc);
    __syncthreads();
    cell_ar_publish(co, s, ++epoch);
    if (!owner) {
      cell_node_sums(s, c);
      bar_workers();
      cell_push_lines(co, s, c, 0);
    }
  }
  for (int j = 0; j < a.j_end; ++j) {
    double* Vcol = WITH_V ? a.V + (size_t)j * a.ldv : nullptr;
    trace_mark(tr, j, 0);

    // ---------------- top of the step: ||b||^2 or beta_{j-1}^2 and the node values of the current vector (workers);
    // the owner warp meanwhile waits for the node sums of its lines
    if (!owner) {
      if (warp < ar_warps) cell_ar_poll(co, s, epoch);
      else cell_poll_gather(co, s, c, (uint32_t)j, ar_warps, kWorkerWarps - ar_warps, 1.0);
      trace_mark(tr, j, 1);
      trace_mark_warp(tr, j, 32);
      bar_workers();
      bar_arrive_top();
      trace_mark(tr, j, 2);
    } else {
      for (uint32_t o = 0; o < c.nown; ++o) {
        const double t = cell_poll_inbox(co, s, o, (uint32_t)j);
        if (lane < kLine) s.T[o * kLine + lane] = t;
      }
      trace_mark_warp(tr, j, 32);
      bar_wait_top();
    }
    const double tot = cell_ar_total(co, s);
    if (j == 0) {
      bnorm = sqrt(tot);
      if (bnorm <= a.tol) {
        status = ST_ZERO_B;
        break;
      }
      sc = 1.0 / bnorm;
    } else {
      const double beta = sqrt(tot);
      if (blockIdx.x == 0 && tid == 0) a.betas[j - 1] = beta;
      if (beta <= a.tol) {  // breakdown: stop (mod.rs:331-338)
        status = ST_BREAKDOWN;
        break;
      }
      sc = 1.0 / beta;  // recip, then multiply (mod.rs:312)
      bp = beta;
    }

    // ---------------- phase A: v = W sc, w~ = A v - beta_{j-1} v_{j-1}, alpha partial
    double acc = 0.0;
    if (owner) {
      // owned node rows: n0 holds W (then w~, then the next W), n1 the previous (normalised) vector
      __syncwarp();
      for (uint32_t t = lane; t < c.nown * kLine; t += 32) {
        const uint32_t u = (blockIdx.x + (t >> 3) * co.Gc) * kLine + (t & 7);
        const double v = __dmul_rn(s.n0[t], sc);
        const double wt = rec_sub(__dmul_rn(sc, s.T[t]), bp, s.n1[t]);
        acc = fma(v, wt, acc);
        s.n0[t] = wt;
        s.n1[t] = v;
        if (WITH_V && u < op.p) __stcs(Vcol + op.m + u, v);
      }
    } else {
      acc = cell_arc_rows<false, WITH_V>(s, c, R, sc, bp, 0.0, 0.0, 0.0, Vcol);
    }
    cell_block_partial(s, acc);
    trace_mark(tr, j, 3);
    trace_mark_warp(tr, j, 48);
    __syncthreads();
    cell_ar_publish(co, s, ++epoch);
    trace_mark(tr, j, 4);
    if (warp < ar_warps) cell_ar_poll(co, s, epoch);
    trace_mark(tr, j, 5);
    __syncthreads();
    const double alpha = cell_ar_total(co, s);
    trace_mark(tr, j, 6);

    // ---------------- phase B: w = w~ - alpha v; node values, node sums of w and the beta partial are published
    acc = 0.0;
    if (owner) {
      for (uint32_t t = lane; t < c.nown * kLine; t += 32) {
        const double w = rec_sub(s.n0[t], alpha, s.n1[t]);
        s.n0[t] = w;
        acc = fma(w, w, acc);
        cell_publish_node(co, blockIdx.x + (t >> 3) * co.Gc, t & 7, w, (uint32_t)j + 1);
      }
    } else {
#pragma unroll
      for (int r = 0; r < kArcRegs; ++r) {
        if (r * kWorkers < c.nA) {
          const uint32_t i = tid + r * kWorkers;
          const double w = rec_sub(R.W[r], alpha, R.P[r]);
          R.W[r] = w;
          acc = fma(w, w, acc);
          if (i < c.nA) s.w[i] = w;
        }
      }
    }
    cell_block_partial(s, acc);
    trace_mark(tr, j, 7);
    __syncthreads();
    cell_ar_publish(co, s, ++epoch);
    trace_mark(tr, j, 8);
    if (!owner) {
      cell_node_sums(s, c);
      bar_workers();
      trace_mark(tr, j, 10);
      cell_push_lines(co, s, c, (uint32_t)j + 1);
    }
    trace_mark(tr, j, 9);
    if (blockIdx.x == 0 && tid == 0) a.alphas[j] = alpha;
    steps = j + 1;
  }
  if (blockIdx.x == 0 && tid == 0) {
    State st;
    st.s_cur = sc;
    st.s_prev = 1.0;
    st.beta_prev = bp;
    st.b_norm = bnorm;
    st.epoch = epoch;
    st.rot = 0;
    st.steps = steps;
    st.status = status;
    *a.st = st;
  }
}

// Replaces lanczos_pass_two_impl (src/algorithms/lanczos_two_pass.rs:206-312) for a cell-partitioned operator.  The owner
// warp runs on its own: node sums of generation j in, node values of generation j + 1 out; the workers only ever wait for
// node values, which were published a whole step earlier.
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_cell_kernel(const IncidenceOp op, const CellOp co, const Pass2Args a) {
  extern __shared__ double smem[];
  const CellSmem s = carve_cell<true>(smem, co);
  const CellCtx c = load_cell(co, s);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool owner = warp == kWorkerWarps;
  const Trace& tr = a.gs.trace;
  ArcRegs<true> R;
  // coefficients are loaded one step ahead of their use so that no step starts with a dependent load
  double c_alpha = 0.0, c_beta = 1.0, c_y = 0.0;
  if (a.steps > 1) {
    c_alpha = __ldg(a.alphas);
    c_beta = __ldg(a.betas);
    c_y = __ldg(a.y + 1);
  }
  double sc = 1.0 / a.b_norm, bp = 0.0;
  {
    // v_1 = b * (1/||b||) held lazily as (b, 1/||b||); x = y_0 v_1   (lanczos_two_pass.rs:247-258)
    const double y0 = __ldg(a.y);
    if (!owner) {
      cell_load_arcs<true>(op, co, s, c, a.b, R);
#pragma unroll
      for (int r = 0; r < kArcRegs; ++r) {
        const uint32_t i = tid + r * kWorkers;
        const double v = __dmul_rn(R.W[r], sc);
        R.X[r] = __dmul_rn(v, y0);
        if (WITH_V && i < c.nA) __stcs(a.V + __ldg(c.gidx + i), v);
      }
    } else {
      for (uint32_t t = lane; t < c.nown * kLine; t += 32) {
        const uint32_t line = blockIdx.x + (t >> 3) * co.Gc, u = line * kLine + (t & 7);
        const double bi = u < op.p ? __ldg(a.b + op.m + u) : 0.0;
        const double v = __dmul_rn(bi, sc);
        s.n0[t] = bi;
        s.n1[t] = 0.0;
        s.nx[t] = __dmul_rn(v, y0);
        cell_publish_node(co, line, t & 7, bi, 0);
        if (WITH_V && u < op.p) __stcs(a.V + op.m + u, v);
      }
    }
    __syncthreads();
    if (!owner) {
      cell_node_sums(s, c);
      bar_workers();
      cell_push_lines(co, s, c, 0);
    }
  }
  for (int j = 0; j + 1 < a.steps; ++j) {
    double* Vcol = WITH_V ? a.V + (size_t)(j + 1) * a.ldv : nullptr;
    const double alpha = c_alpha, beta = c_beta, yj = c_y;
    const double sinv = 1.0 / beta;
    if (j + 2 < a.steps) {  // prefetch the next step's coefficients
      c_alpha = __ldg(a.alphas + j + 1);
      c_beta = __ldg(a.betas + j + 1);
      c_y = __ldg(a.y + j + 2);
    }
    if (owner) {
      for (uint32_t o = 0; o < c.nown; ++o) {
        const double T = cell_poll_inbox(co, s, o, (uint32_t)j);
        if (lane < kLine) {
          const uint32_t t = o * kLine + lane, line = blockIdx.x + o * co.Gc, u = line * kLine + lane;
          const double v = __dmul_rn(s.n0[t], sc);
          const double w = rec_sub(rec_sub(__dmul_rn(sc, T), bp, s.n1[t]), alpha, v);
          const double vn = __dmul_rn(w, sinv);
          s.n0[t] = w;
          s.n1[t] = v;
          cell_publish_node(co, line, lane, w, (uint32_t)j + 1);
          s.nx[t] = __dadd_rn(s.nx[t], __dmul_rn(yj, vn));
          if (WITH_V && u < op.p) __stcs(Vcol + op.m + u, vn);
        }
        __syncwarp();
      }
      trace_mark_warp(tr, j, 32);
    } else {
      trace_mark(tr, j, 0);
      cell_poll_gather(co, s, c, (uint32_t)j, 0, kWorkerWarps, sc);
      trace_mark(tr, j, 1);
      trace_mark_warp(tr, j, 32);
      bar_workers();  // also: every warp has finished the node sums of the previous step, s.w may be rewritten
      trace_mark(tr, j, 2);
      cell_arc_rows<true, WITH_V>(s, c, R, sc, bp, alpha, sinv, yj, Vcol);
      trace_mark(tr, j, 3);
      bar_workers();
      trace_mark(tr, j, 4);
      cell_node_sums(s, c);
      trace_mark_warp(tr, j, 48);
      bar_workers();
      trace_mark(tr, j, 6);
      cell_push_lines(co, s, c, (uint32_t)j + 1);
      trace_mark(tr, j, 5);
    }
    sc = sinv;
    bp = beta;
  }
  if (!owner) {
#pragma unroll
    for (int r = 0; r < kArcRegs; ++r) {
      const uint32_t i = tid + r * kWorkers;
      if (i < c.nA) a.x[__ldg(c.gidx + i)] = R.X[r];
    }
  } else {
    for (uint32_t t = lane; t < c.nown * kLine; t += 32) {
      const uint32_t u = (blockIdx.x + (t >> 3) * co.Gc) * kLine + (t & 7);
      if (u < op.p) a.x[op.m + u] = s.nx[t];
    }
  }
}

}  // namespace tpl
