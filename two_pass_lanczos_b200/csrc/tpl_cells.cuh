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
constexpr int kArcBatch = 3;         // arcs whose shared-memory operands are requested together
constexpr uint32_t kArCopies = 12;   // copies of every all-reduce line
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

// ----------------------------------------------------------------------------- shared memory
// Every array is addressed through its 32-bit shared-window byte address with explicit ld/st.shared: with generic pointers
// the compiler re-derives the window base (S2UR SR_CgaCtaId + ULEA) and 64-bit addresses at every access (measured: ~15 of
// the ~40 instructions per arc row in the first version of these kernels).
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct CellSmem {  // shared-window byte addresses
  uint32_t w;      // f64 [Amax + 8] arc part of the newest vector (what the node sums are formed from); w[Amax] = 0
  uint32_t d;      // f64 [Amax] D of the cell's arcs
  // node values of the current vector: [8 * max_lines by local node | max_tl by tail rank (the arcs of a half-warp have
  // consecutive ranks: conflict-free) | zero slot (self-loops, empty arc slots)]
  uint32_t nodev;
  uint32_t sums;   // f64 [8 * max_slots] node sums of this cell in push order
  uint32_t n0, n1, nx, T;  // f64 [8 * max_own] owned node rows: current / previous vector, x (pass 2), node sums
  uint32_t arv;    // f64 [Gc] all-reduce values by slot
  uint32_t wpart;  // f64 [kWarps + 2]
  uint32_t walk;   // uint4 [max_groups * 32]
  uint32_t ent4;   // uint2 [max_rows * 32]
  uint32_t lines;  // u32 [max_lines]
  uint32_t push;   // u32 [max_slots]
  uint32_t own;    // u32 [2 * max_own] {first slot, slots} of the owned lines
  uint32_t tmap;   // u16 [8 * max_lines]
  uint32_t th;     // u32 [Amax] tail / head slots of the arcs (pass 2 only: pass 1 has the registers to keep them)
};

__host__ __device__ inline size_t cell_smem_bytes(const CellOp& co, bool pass2) {
  size_t dbl = 2 * (size_t)co.Amax + 8 + kLine * co.max_lines + co.max_tl + 8 + kLine * co.max_slots +
               (size_t)kLine * co.max_own * (pass2 ? 4 : 3) + co.Gc + kWarps + 2;
  dbl = (dbl + 1) & ~(size_t)1;
  const size_t u32 = (size_t)co.max_lines + co.max_slots + 2 * (size_t)co.max_own;
  return dbl * 8 + (size_t)co.max_groups * 32 * 16 + (size_t)co.max_rows * 32 * 8 + u32 * 4 + (size_t)kLine * co.max_lines * 2 + 16 +
         (pass2 ? (size_t)co.Amax * 4 + 4 : 0);
}

template <bool PASS2>
__device__ __forceinline__ CellSmem carve_cell(double* base, const CellOp& co) {
  CellSmem s;
  uint32_t a = (uint32_t)__cvta_generic_to_shared(base);
  s.w = a; a += (co.Amax + 8) * 8;
  s.d = a; a += co.Amax * 8;
  s.nodev = a; a += (kLine * co.max_lines + co.max_tl + 8) * 8;
  s.sums = a; a += kLine * co.max_slots * 8;
  s.n0 = a; a += kLine * co.max_own * 8;
  s.n1 = a; a += kLine * co.max_own * 8;
  s.nx = a; a += PASS2 ? kLine * co.max_own * 8 : 0;
  s.T = a; a += kLine * co.max_own * 8;
  s.arv = a; a += co.Gc * 8;
  s.wpart = a; a += (kWarps + 2) * 8;  // + the two scalars of the top of a pass-1 step (norm, reciprocal)
  a = (a + 15u) & ~15u;
  s.walk = a; a += co.max_groups * 32 * 16;
  s.ent4 = a; a += co.max_rows * 32 * 8;
  s.lines = a; a += co.max_lines * 4;
  s.push = a; a += co.max_slots * 4;
  s.own = a; a += 2 * co.max_own * 4;
  s.tmap = a; a += kLine * co.max_lines * 2;
  s.th = (a + 3u) & ~3u;
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
  for (uint32_t i = threadIdx.x; i < x.nlines; i += kBlock) sts_u32(s.lines + i * 4, __ldg(ln + i));
  const uint16_t* tm = co.tmap + (size_t)c * kLine * co.max_lines;
  for (uint32_t i = threadIdx.x; i < x.nlines * kLine; i += kBlock) sts_u16(s.tmap + i * 2, __ldg(tm + i));
  for (uint32_t i = threadIdx.x; i < kLine * co.max_lines + co.max_tl + 8; i += kBlock) sts_f64(s.nodev + i * 8, 0.0);
  const uint32_t* ps = co.push + (size_t)c * co.max_slots;
  for (uint32_t i = threadIdx.x; i < x.nslots; i += kBlock) sts_u32(s.push + i * 4, __ldg(ps + i));
  const uint4* wk = co.walk + (size_t)c * co.max_groups * 32;
  for (uint32_t i = threadIdx.x; i < x.ngroups * 32; i += kBlock) sts_v4(s.walk + i * 16, __ldg(wk + i));
  const uint2* e4 = co.ent4 + (size_t)c * co.max_rows * 32;
  for (uint32_t i = threadIdx.x; i < nrows * 32; i += kBlock) sts_v2(s.ent4 + i * 8, __ldg(e4 + i));
  for (uint32_t i = threadIdx.x; i < x.nslots * kLine; i += kBlock) sts_f64(s.sums + i * 8, 0.0);  // atoms without a list stay zero
  for (uint32_t i = threadIdx.x; i < 8; i += kBlock) sts_f64(s.w + (co.Amax + i) * 8, 0.0);        // the zero slot padding entries point at
  for (uint32_t o = threadIdx.x; o < x.nown; o += kBlock) {
    const uint32_t l = c + o * co.Gc;
    const uint32_t b0 = __ldg(co.slot_base + l);
    sts_u32(s.own + (2 * o) * 4, b0);
    sts_u32(s.own + (2 * o + 1) * 4, __ldg(co.slot_base + l + 1) - b0);
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
__device__ __forceinline__ double atom_poll(const uint4* p, uint32_t tag) {
  uint4 f = ld_relaxed_gpu_v4(p);
  uint32_t spins = 0;
  while (f.y != tag || f.w != tag) {
    if (++spins > kSpinLimit) __trap();
    f = ld_relaxed_gpu_v4(p);
  }
  return atom_value(f);
}

// ----------------------------------------------------------------------------- named barriers
// barrier 0 (__syncthreads): all 16 warps; barrier 1: the 15 worker warps; barrier 2: the workers ARRIVE (without waiting)
// once the all-reduce values of the top of a step are in shared memory, the owner warp waits for them.
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory"); }
__device__ __forceinline__ void bar_arrive_top() { asm volatile("bar.arrive 2, %0;" ::"n"(kBlock) : "memory"); }
__device__ __forceinline__ void bar_wait_top() { asm volatile("bar.sync 2, %0;" ::"n"(kBlock) : "memory"); }
// barrier 3: the warps that polled the all-reduce lines of the top of a step, among themselves
__device__ __forceinline__ void bar_ar_warps(uint32_t threads) { asm volatile("bar.sync 3, %0;" ::"r"(threads) : "memory"); }

// ----------------------------------------------------------------------------- exchange steps
__device__ __forceinline__ double flip_sign(double v, uint32_t mask) {
  return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}
// Node sums of the arc values s.w of this cell -> s.sums.  One warp per group of 16 lists; lane l (< 16) adds the even
// entries of list l in order, lane l + 16 the odd ones, then the two halves are added.  A row of entries is one 8-byte
// word per lane (four 16-bit positions), rows of a group are 32 words apart.  The sign (tail +, head -) is applied by
// flipping the sign bit (a - x == a + (-x) exactly).  Caller syncs before cell_push_lines.
__device__ __forceinline__ void cell_node_sums(const CellSmem& s, const CellCtx& c) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t g = warp; g < c.ngroups; g += kWorkerWarps) {
    const uint4 d = lds_v4(s.walk + (g * 32 + lane) * 16);
    uint32_t row = s.ent4 + (d.z * 32 + lane) * 8;
    double acc = 0.0;
    uint2 e = make_uint2(0, 0);
    if (d.w) e = lds_v2(row);
    for (uint32_t k = 0; k < d.w; ++k) {
      row += 32 * 8;
      uint2 nxt = e;
      if (k + 1 < d.w) nxt = lds_v2(row);  // the next row is requested before this one is consumed
      const double v0 = lds_f64(s.w + (e.x & 0xffffu) * 8), v1 = lds_f64(s.w + (e.x >> 16) * 8);
      const double v2 = lds_f64(s.w + (e.y & 0xffffu) * 8), v3 = lds_f64(s.w + (e.y >> 16) * 8);
      acc = __dadd_rn(acc, flip_sign(v0, d.y));
      acc = __dadd_rn(acc, flip_sign(v1, d.y));
      acc = __dadd_rn(acc, flip_sign(v2, d.y));
      acc = __dadd_rn(acc, flip_sign(v3, d.y));
      e = nxt;
    }
    acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
    if (lane < 16 && d.x != 0xffffffffu) sts_f64(s.sums + d.x * 8, acc);
  }
}
// s.sums -> the owners' inboxes as generation `gen`, one whole line per 8 adjacent lanes.
__device__ __forceinline__ void cell_push_lines(const CellOp& co, const CellSmem& s, const CellCtx& c, uint32_t gen) {
  uint4* box = co.inbox + (size_t)(gen & 1u) * co.inbox_atoms;
  const uint32_t tag = gen + 1;
  for (uint32_t t = threadIdx.x; t < c.nslots * kLine; t += kWorkers)
    st_relaxed_gpu_v4(box + (size_t)lds_u32(s.push + (t >> 3) * 4) * kLine + (t & 7), atom_pack(lds_f64(s.sums + t * 8), tag));
}

// Owner side: the sum of the contributions to node r = lane & 7 of owned line o, generation `gen` (same value in the
// four lanes of a node).  Lane = r + 8 q adds slots q, q+4, ... in order; the four partial sums are combined by a fixed xor tree.
__device__ __forceinline__ double cell_poll_inbox(const CellOp& co, const CellSmem& s, uint32_t o, uint32_t gen) {
  const uint4* box = co.inbox + (size_t)(gen & 1u) * co.inbox_atoms;
  const uint32_t tag = gen + 1;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t r = lane & 7, q = lane >> 3;
  const uint32_t b0 = lds_u32(s.own + (2 * o) * 4), K = lds_u32(s.own + (2 * o + 1) * 4);
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
        if (k < K) f[i] = ld_relaxed_gpu_v4(mine + (size_t)k * kLine);
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

// Node values of the touched lines, generation `gen`, times `scale` -> nodev (and the by-rank mirror of the tails).
__device__ __forceinline__ void cell_poll_gather(const CellOp& co, const CellSmem& s, const CellCtx& c, uint32_t gen,
                                                 uint32_t warp0, uint32_t nwarps, double scale) {
  const uint4* g = co.gather + (size_t)(gen & 1u) * co.L * kLine;
  const uint32_t tag = gen + 1;
  for (uint32_t a = threadIdx.x - warp0 * 32; a < c.nlines * kLine; a += nwarps * 32) {
    const double v = __dmul_rn(atom_poll(g + (size_t)lds_u32(s.lines + (a >> 3) * 4) * kLine + (a & 7), tag), scale);
    const uint32_t mirror = lds_u16(s.tmap + a * 2);
    sts_f64(s.nodev + a * 8, v);
    if (mirror != 0xffffu) sts_f64(s.nodev + mirror * 8, v);
  }
}

__device__ __forceinline__ void cell_publish_node(const CellOp& co, uint32_t line, uint32_t r, double v, uint32_t gen) {
  st_relaxed_gpu_v4(co.gather + ((size_t)(gen & 1u) * co.L + line) * kLine + r, atom_pack(v, gen + 1));
}

// All-reduce, publishing half (owner warp): the CTA's partial (already summed over the warps into wpart, caller synced)
// goes out as one whole line (eight copies of the atom, lanes 0..7).
__device__ __forceinline__ void cell_ar_publish(const CellOp& co, const CellSmem& s, uint32_t epoch) {
  const uint32_t lane = threadIdx.x & 31;
  double t = lane < kWarps ? lds_f64(s.wpart + lane * 8) : 0.0;
  t = warp_sum(t);
  // kArCopies copies of the line: a copy is polled by Gc / kArCopies CTAs instead of all of them
  const uint4 atom = atom_pack(t, epoch);
  for (uint32_t q = lane; q < kArCopies * kLine; q += 32)
    st_relaxed_gpu_v4(co.ar + (((size_t)(epoch & 1u) * kArCopies + (q >> 3)) * co.Gc + blockIdx.x) * kLine + (q & 7), atom);
}
// All-reduce, polling half: thread t of the first ceil(Gc / 32) warps polls the line of CTA (t + cta) % Gc (rotated so
// that the CTAs do not all start on the same line).  Caller syncs, then every warp calls cell_ar_total.
__device__ __forceinline__ void cell_ar_poll(const CellOp& co, const CellSmem& s, uint32_t epoch) {
  if (threadIdx.x < co.Gc) {
    uint32_t slot = threadIdx.x + blockIdx.x;
    slot = slot >= co.Gc ? slot - co.Gc : slot;
    sts_f64(s.arv + slot * 8, atom_poll(co.ar + (((size_t)(epoch & 1u) * kArCopies + blockIdx.x % kArCopies) * co.Gc + slot) * kLine + ((blockIdx.x / kArCopies) & 7u), epoch));
  }
}
__device__ __forceinline__ double cell_ar_total(const CellOp& co, const CellSmem& s) {
  const uint32_t lane = threadIdx.x & 31;
  double t = 0.0;
  for (uint32_t i0 = 0; i0 < co.Gc; i0 += 32 * 6) {  // one trip for any grid of up to 192 CTAs: six loads in flight, then the adds
    double v[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const uint32_t i = i0 + lane + 32 * u;
      v[u] = i < co.Gc ? lds_f64(s.arv + i * 8) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 6; ++u)
      if (i0 + lane + 32 * u < co.Gc) t = __dadd_rn(t, v[u]);
  }
  return warp_sum(t);
}
__device__ __forceinline__ void cell_block_partial(const CellSmem& s, double acc) {
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sts_f64(s.wpart + (threadIdx.x >> 5) * 8, acc);
}

// The arcs of a cell are spread over the worker threads, arc r of thread t sitting at position t + r * kWorkers of the
// cell's (jagged-diagonal) order.  A thread keeps its arcs' tail / head, the current vector W (un-normalised, v = W * sc),
// the previous vector v_{j-1} (normalised) and, in pass 2, x in registers for the whole pass; D stays in shared memory.
template <bool PASS2>
struct ArcRegs {
  double W[kArcRegs], P[kArcRegs], X[PASS2 ? kArcRegs : 1];
  uint32_t TH[PASS2 ? 1 : kArcRegs];  // pass 2 reads them from shared memory (s.th)
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
    if (!PASS2) R.TH[r] = zero_slot | (zero_slot << 16);
    if (i < c.nA) {
      const uint32_t g = __ldg(c.gidx + i);
      R.W[r] = __ldg(b + g);
      if (PASS2) sts_u32(s.th + i * 4, __ldg(lth + i));
      else R.TH[r] = __ldg(lth + i);
      sts_f64(s.d + i * 8, __ldg(op.d + g));
      sts_f64(s.w + i * 8, R.W[r]);
      acc = fma(R.W[r], R.W[r], acc);
    }
  }
  return acc;
}

// Arc rows of one step: v = W sc replaces the previous vector (the same single rounding the reference performs when it
// scales w in place, mod.rs:312-315) and W becomes
//   pass 1 (PASS2 = false): w~ = A v - beta_{j-1} v_{j-1}; returns the partial of alpha = <v, w~>
//   pass 2 (PASS2 = true):  w  = w~ - alpha v, written to s.w, and x += y_{j+1} (w / beta_j)
// (A v)_arc follows the reference's CSC accumulation order: D v first, then the two node columns by ascending node index
// (bit 15 of TH: the tail comes first); a self-loop or an empty arc slot reads the zero slot twice.
template <bool PASS2, bool WITH_V>
__device__ __forceinline__ double cell_arc_rows(const CellOp& co, const CellSmem& s, const CellCtx& c, ArcRegs<PASS2>& R, double sc,
                                                double bp, double alpha, double sinv, double yj, double* Vcol) {
  const uint32_t zero_slot = kLine * co.max_lines + co.max_tl;
  double acc = 0.0;
  const uint32_t dbase = s.d + threadIdx.x * 8, wbase = s.w + threadIdx.x * 8;
#pragma unroll
  for (int r0 = 0; r0 < kArcRegs; r0 += kArcBatch) {
    if (r0 * kWorkers < c.nA) {  // uniform: batches of arc slots beyond the cell's last arc are skipped
      double dd[kArcBatch], xt[kArcBatch], xh[kArcBatch];
      uint32_t th[kArcBatch];
#pragma unroll
      for (int u = 0; u < kArcBatch; ++u) {
        const int r = r0 + u;
        const bool live = threadIdx.x + r * kWorkers < c.nA;
        if (PASS2) th[u] = live ? lds_u32(s.th + (threadIdx.x + r * kWorkers) * 4) : (zero_slot | (zero_slot << 16));
        else th[u] = R.TH[r];
        dd[u] = live ? lds_f64(dbase + r * kWorkers * 8) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kArcBatch; ++u) {
        xt[u] = lds_f64(s.nodev + (th[u] & 0x7fffu) * 8);
        xh[u] = lds_f64(s.nodev + (th[u] >> 16) * 8);
      }
#pragma unroll
      for (int u = 0; u < kArcBatch; ++u) {
        const int r = r0 + u;
        const uint32_t i = threadIdx.x + r * kWorkers;
        const double v = __dmul_rn(R.W[r], sc);
        // pass 2 scales the node values once when it polls them; pass 1 only learns sc together with them
        const double a = PASS2 ? xt[u] : __dmul_rn(xt[u], sc), nb = flip_sign(PASS2 ? xh[u] : __dmul_rn(xh[u], sc), 0x80000000u);
        const bool tail_first = th[u] & 0x8000u;
        double row = __dmul_rn(dd[u], v);
        row = __dadd_rn(row, tail_first ? a : nb);
        row = __dadd_rn(row, tail_first ? nb : a);
        const double wt = rec_sub(row, bp, R.P[r]);
        R.P[r] = v;
        if (PASS2) {
          const double w = rec_sub(wt, alpha, v);
          const double vn = __dmul_rn(w, sinv);
          R.W[r] = w;
          R.X[r] = __dadd_rn(R.X[r], __dmul_rn(yj, vn));
          if (i < c.nA) {
            sts_f64(wbase + r * kWorkers * 8, w);
            if (WITH_V) __stcs(Vcol + __ldg(c.gidx + i), vn);
          }
        } else {
          acc = fma(v, wt, acc);
          R.W[r] = wt;
          if (WITH_V && i < c.nA) __stcs(Vcol + __ldg(c.gidx + i), v);
        }
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
        sts_f64(s.n0 + t * 8, bi);
        sts_f64(s.n1 + t * 8, 0.0);
        acc = fma(bi, bi, acc);
        cell_publish_node(co, line, t & 7, bi, 0);
      }
    }
    cell_block_partial(s, acc);
    __syncthreads();
    ++epoch;
    if (owner) {
      cell_ar_publish(co, s, epoch);
    } else {
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
      if (warp < ar_warps) {
        // the all-reduce usually completes before the node values do: the norm and its reciprocal (a sum of Gc numbers, a
        // square root and a division -- ~500 cycles of dependent FP64) are formed by warp 0 in the shadow of the other
        // warps' gather polls and handed over in shared memory
        cell_ar_poll(co, s, epoch);
        bar_ar_warps(ar_warps * 32);
        if (warp == 0) {
          const double root = sqrt(cell_ar_total(co, s));
          if (lane == 0) {
            sts_f64(s.wpart + kWarps * 8, root);
            sts_f64(s.wpart + (kWarps + 1) * 8, 1.0 / root);
          }
        }
      } else {
        cell_poll_gather(co, s, c, (uint32_t)j, ar_warps, kWorkerWarps - ar_warps, 1.0);
      }
      trace_mark(tr, j, 1);
      trace_mark_warp(tr, j, 32);
      bar_workers();
      bar_arrive_top();
      trace_mark(tr, j, 2);
    } else {
      for (uint32_t o = 0; o < c.nown; ++o) {
        const double t = cell_poll_inbox(co, s, o, (uint32_t)j);
        if (lane < kLine) sts_f64(s.T + (o * kLine + lane) * 8, t);
      }
      trace_mark_warp(tr, j, 32);
      bar_wait_top();
    }
    const double root = lds_f64(s.wpart + kWarps * 8), rinv = lds_f64(s.wpart + (kWarps + 1) * 8);
    if (j == 0) {
      bnorm = root;
      if (bnorm <= a.tol) {
        status = ST_ZERO_B;
        break;
      }
      sc = rinv;
    } else {
      const double beta = root;
      if (blockIdx.x == 0 && tid == 0) a.betas[j - 1] = beta;
      if (beta <= a.tol) {  // breakdown: stop (mod.rs:331-338)
        status = ST_BREAKDOWN;
        break;
      }
      sc = rinv;  // recip, then multiply (mod.rs:312)
      bp = beta;
    }

    // ---------------- phase A: v = W sc, w~ = A v - beta_{j-1} v_{j-1}, alpha partial
    trace_mark(tr, j, 11);
    double acc = 0.0;
    if (owner) {
      // owned node rows: n0 holds W (then w~, then the next W), n1 the previous (normalised) vector
      __syncwarp();
      for (uint32_t t = lane; t < c.nown * kLine; t += 32) {
        const uint32_t u = (blockIdx.x + (t >> 3) * co.Gc) * kLine + (t & 7);
        const double v = __dmul_rn(lds_f64(s.n0 + t * 8), sc);
        const double wt = rec_sub(__dmul_rn(sc, lds_f64(s.T + t * 8)), bp, lds_f64(s.n1 + t * 8));
        acc = fma(v, wt, acc);
        sts_f64(s.n0 + t * 8, wt);
        sts_f64(s.n1 + t * 8, v);
        if (WITH_V && u < op.p) __stcs(Vcol + op.m + u, v);
      }
    } else {
      acc = cell_arc_rows<false, WITH_V>(co, s, c, R, sc, bp, 0.0, 0.0, 0.0, Vcol);
    }
    trace_mark(tr, j, 12);
    cell_block_partial(s, acc);
    trace_mark(tr, j, 3);
    __syncthreads();
    ++epoch;
    if (owner) cell_ar_publish(co, s, epoch);
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
        const double w = rec_sub(lds_f64(s.n0 + t * 8), alpha, lds_f64(s.n1 + t * 8));
        sts_f64(s.n0 + t * 8, w);
        acc = fma(w, w, acc);
        cell_publish_node(co, blockIdx.x + (t >> 3) * co.Gc, t & 7, w, (uint32_t)j + 1);
      }
    } else {
#pragma unroll
      for (int r = 0; r < kArcRegs; ++r) {
        if (r * kWorkers < c.nA) {
          const double w = rec_sub(R.W[r], alpha, R.P[r]);
          R.W[r] = w;
          acc = fma(w, w, acc);
          if (tid + r * kWorkers < c.nA) sts_f64(s.w + (tid + r * kWorkers) * 8, w);
        }
      }
    }
    cell_block_partial(s, acc);
    trace_mark(tr, j, 7);
    __syncthreads();
    ++epoch;
    if (owner) cell_ar_publish(co, s, epoch);
    trace_mark(tr, j, 8);
    if (!owner) {
      cell_node_sums(s, c);
      trace_mark_warp(tr, j, 48);
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
        sts_f64(s.n0 + t * 8, bi);
        sts_f64(s.n1 + t * 8, 0.0);
        sts_f64(s.nx + t * 8, __dmul_rn(v, y0));
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
          const double v = __dmul_rn(lds_f64(s.n0 + t * 8), sc);
          const double w = rec_sub(rec_sub(__dmul_rn(sc, T), bp, lds_f64(s.n1 + t * 8)), alpha, v);
          const double vn = __dmul_rn(w, sinv);
          sts_f64(s.n0 + t * 8, w);
          sts_f64(s.n1 + t * 8, v);
          cell_publish_node(co, line, lane, w, (uint32_t)j + 1);
          sts_f64(s.nx + t * 8, __dadd_rn(lds_f64(s.nx + t * 8), __dmul_rn(yj, vn)));
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
      cell_arc_rows<true, WITH_V>(co, s, c, R, sc, bp, alpha, sinv, yj, Vcol);
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
      if (u < op.p) a.x[op.m + u] = lds_f64(s.nx + t * 8);
    }
  }
}

}  // namespace tpl
