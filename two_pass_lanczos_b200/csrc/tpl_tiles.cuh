// tpl_tiles.cuh -- streaming kernels of the KKT incidence operator with TILED node sums (any size that fits the
// shared-memory budget below; the gather kernels of tpl_kernels.cuh stay as the generic fallback and the CSR path).
//
// Why: the node rows of A = [[D, E^T], [E, 0]] are sums over ~2m/p arc values each.  Gathering them from HBM through a
// node -> arc list costs one 32-byte sector per 8-byte value (measured: 24-33 % of the HBM roofline at 5M-50M arcs).
// Here every CTA instead forms the PARTIAL node sums of its own contiguous arc chunk while the freshly computed arc
// values are still in shared memory:
//   * the chunk is cut into tiles of T arcs; a tile's values are written to shared memory as they are produced;
//   * thread i of the CTA owns nodes [i*npt, (i+1)*npt) and walks, per tile, its own slice of a host-built entry list
//     sorted by node: entry = (node offset, +-, index into the tile), adding into a private slice of a p-long
//     shared-memory accumulator (no atomics, fixed order);
//   * runs of >= 32 consecutive arcs with the same tail (netgen emits arcs grouped by tail) are summed by one warp as a
//     "piece" (lane-strided + xor tree) whose result is appended to the tile as a virtual arc, so that the walk treats
//     it like any other entry; instances without such runs simply have every tail in the entry list.
//   * after the last tile the p partial sums of the CTA go to HBM once per step ([G][p] doubles, ~0.3 % of the step's
//     traffic); after the grid barrier that the step needs anyway, the owner of a node block adds the G partials in a
//     fixed order.
// Per step the operator costs 16 B/arc (d, tail, head) + 4..8 B/arc (entries) of coalesced HBM traffic, which is the
// algorithmic figure of SURVEY 8d (B_inc = 24m + 4p).  Summation order depends only on the operator and the grid, so
// pass 1, the one-pass variant and pass 2 still produce bit-identical basis vectors.
#pragma once
#include "tpl_kernels.cuh"

namespace tpl {

constexpr uint32_t kPieceMin = 32;     // shortest same-tail run summed as a piece
constexpr uint32_t kPieceMax = 256;    // longest piece (longer runs are split at fixed offsets)
constexpr uint32_t kMaxPieces = 512;   // per tile (shared-memory slots after the T arc values); T + kMaxPieces <= 16384
constexpr int kUnroll = 4;             // arcs per thread per batch of the streaming loops
constexpr int kUnroll2 = 2;            // pass 2: six loads per arc, two batches of registers in flight
constexpr int kUnrollB = 8;            // pass 1 phase B: two loads per arc
constexpr int kPre = 16;               // list entries per fold thread requested together
constexpr int kStreamWarps = 8;        // warps 0..7 stream a tile from HBM while warps 8..15 fold the previous one into the node sums
                                       // (10 / 6 is better at 5M arcs, worse at 50M; 12 / 4 is worse everywhere)
constexpr int kFoldWarps = kWarps - kStreamWarps;
constexpr int kStreamThreads = kStreamWarps * 32, kFoldThreads = kFoldWarps * 32;

// The exchange buffers of every rank of an arc-partitioned multi-GPU operator as seen from THIS rank (peer memory mapped
// through CUDA IPC over NVLink); world == 1: the local buffers only.  The persistent kernels below write partial node sums,
// published node values and barrier slots straight into the owners' / every rank's buffers and only ever poll and read
// local memory: the collective of a Lanczos step is fused into the kernels, no NCCL call and no extra launch on the data path.
constexpr int kMaxRanks = 8;
struct Fabric {
  int rank, world;
  uint32_t Gtot;  // CTAs of all ranks
  uint32_t Bp;    // node rows owned by one rank = G * R
  double* partials[kMaxRanks];  // [2][Gtot][Bp]  partial node sums addressed to the rank's node rows, by source CTA
  double* nodebuf[kMaxRanks];   // [2][p]         node part of the newest vector (un-normalised), replicated on every rank
  uint4* slots[kMaxRanks];      // [2][Gtot]      barrier / all-reduce slots, replicated on every rank
};

struct TileOp {
  uint32_t T;         // arcs per tile
  uint32_t ntile;     // tiles per CTA chunk
  uint32_t R;         // node rows per owner block = ceil(p / (world * G))
  Fabric fab;
  // per tile: {first entry word, entries per thread L, first piece, end piece}
  const uint4* thdr;      // [G * ntile]
  // entries, per tile L x kFoldThreads words, thread-interleaved (word q of fold thread i at q * kFoldThreads + i => coalesced):
  //   node << 15 | minus << 14 | index into the tile (arcs, then pieces);  0xffffffff = padding.
  // A tile's entries are sorted by node and cut into kBlock slices of (nearly) equal length at node boundaries, so a
  // node is folded by exactly one thread per tile and all threads carry the same load.
  const uint32_t* lent;
  const uint32_t* piece;  // first | (len - 1) << 16
};
constexpr uint32_t kEntPad = 0xffffffffu;

// Shared-memory arrays are addressed through 32-bit shared-window addresses and explicit ld/st.shared: with generic
// `double*` members the compiler re-derives the window base (S2UR SR_CgaCtaId + address arithmetic) inside divergent code.
struct SmArr {
  uint32_t a;  // shared-window byte address of element 0
};
__device__ __forceinline__ double sm_ld(SmArr arr, uint32_t i) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(arr.a + i * 8u) : "memory");
  return v;
}
__device__ __forceinline__ void sm_st(SmArr arr, uint32_t i, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(arr.a + i * 8u), "d"(v) : "memory");
}
struct TileSmem {
  SmArr node;  // [p]  scaled node segment of the current vector (phases that form arc rows)
  SmArr acc;   // [p]  partial node sums of this CTA            (phases that produce a new vector)
  SmArr wt;    // [2][T + kMaxPieces] arc values of a tile + its piece sums, double-buffered (stream / fold)
  uint32_t wt_stride;  // bytes between the two buffers
};
// pass 1 never needs node and acc at the same time (they alias); pass 2 needs both.
__host__ __device__ inline size_t tile_smem_bytes(uint32_t p, uint32_t T, bool pass2) {
  return ((pass2 ? 2 : 1) * (size_t)p + 2 * ((size_t)T + kMaxPieces)) * sizeof(double);
}
__device__ __forceinline__ TileSmem carve_tiles(double* base, uint32_t p, uint32_t T, bool pass2) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(base);
  TileSmem s;
  s.node.a = b;
  s.acc.a = pass2 ? b + p * 8u : b;
  s.wt.a = s.acc.a + p * 8u;
  s.wt_stride = (T + kMaxPieces) * 8u;
  return s;
}

struct TileHdr {
  uint32_t e0, L;   // first entry word of the tile, entries per fold thread
  uint32_t q0, q1;  // the tile's pieces
};
__device__ __forceinline__ TileHdr tile_hdr(const TileOp& to, uint32_t tile_id) {
  TileHdr h;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(h.e0), "=r"(h.L), "=r"(h.q0), "=r"(h.q1)
               : "l"(to.thdr + tile_id));
  return h;
}

// One list entry: acc[node] += (+-) wt[index], a plain read-modify-write (a thread's entries are folded in list order, so
// repeated nodes need no special care).  The sign is applied by flipping the sign bit of the tile value: a - x and
// a + (-x) are the same IEEE operation.  The fold is instruction-issue bound (one entry per arc per step), hence the
// minimal instruction count per entry.
__device__ __forceinline__ void fold_entry(uint32_t ent, SmArr wt, SmArr acc) {
  const uint32_t node = ent >> 15;
  const long long x = __double_as_longlong(sm_ld(wt, ent & 0x3fffu)) ^ ((long long)(ent & 0x4000u) << 49);
  sm_st(acc, node, __dadd_rn(sm_ld(acc, node), __longlong_as_double(x)));
}

// named barriers of the stream / fold hand-off (barrier 0 is __syncthreads): full[b] = tile buffer b holds a streamed tile
// (the stream warps arrive, the fold warps wait), empty[b] = the fold warps are done with buffer b (they arrive, the stream
// warps wait before refilling it), kBarFold = the fold warps among themselves.  bar.arrive / bar.sync order the shared-memory
// accesses of the arriving threads before those of the waiting threads (PTX producer / consumer pattern).
constexpr int kBarFull = 1, kBarEmpty = 3, kBarFold = 5;
__device__ __forceinline__ void bar_sync_n(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// List and piece words of the fold are static data: a fold thread requests them one batch AHEAD of their use -- the next
// kPre entries (of this tile, or the first ones of the next tile) before it folds the current batch, the next tile's piece
// words while it folds this tile -- so that no global-load latency sits between two hand-offs of the stream / fold pipeline
// (before: ~3 dependent piece loads + L / kPre list round trips per tile, about a third of the fold's time at 50M arcs).
struct FoldRegs {
  uint32_t ent[kPre];  // the batch to fold next
  uint32_t pc[2];      // piece words of this warp: lane l holds pieces fwarp + kFoldWarps * (l + 32 * i)
};
static_assert(kMaxPieces <= 2 * 32 * kFoldWarps, "two piece words per lane cover a tile");
__device__ __forceinline__ void fold_request(const TileOp& to, const TileHdr& h, uint32_t q0, uint32_t (&ent)[kPre]) {
  const uint32_t* mine = to.lent + h.e0 + (threadIdx.x - kStreamThreads);
#pragma unroll
  for (int q = 0; q < kPre; ++q) ent[q] = q0 + q < h.L ? __ldg(mine + (size_t)(q0 + q) * kFoldThreads) : kEntPad;
}
__device__ __forceinline__ void piece_request(const TileOp& to, const TileHdr& h, uint32_t (&pc)[2]) {
  const int ftid = threadIdx.x - kStreamThreads, lane = ftid & 31, fwarp = ftid >> 5;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t q = h.q0 + fwarp + kFoldWarps * (lane + 32 * i);
    pc[i] = q < h.q1 ? __ldg(to.piece + q) : 0u;
  }
}

// Fold warps: adds the node sums of the tile held in buffer `wt` into s.acc.  `fr` holds the tile's first list batch and
// its piece words on entry, those of tile `next` (if there is one) on return.
__device__ __forceinline__ void tile_node_sums(const TileOp& to, const TileSmem& s, SmArr wt, const TileHdr& h, const TileHdr& next,
                                               bool has_next, FoldRegs& fr) {
  const int ftid = threadIdx.x - kStreamThreads, lane = ftid & 31, fwarp = ftid >> 5;
  if (h.q1 > h.q0) {
    uint32_t i = 0;
    for (uint32_t q = h.q0 + fwarp; q < h.q1; q += kFoldWarps, ++i) {
      const uint32_t pc = __shfl_sync(0xffffffffu, i < 32 ? fr.pc[0] : fr.pc[1], i & 31);
      const uint32_t first = pc & 0xffffu, len = (pc >> 16) + 1;
      const SmArr w{wt.a + first * 8u};
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four independent lane-strided chains, combined in a fixed order
      for (uint32_t e = lane; e < len; e += 128) {
        const double x0 = sm_ld(w, e);
        const double x1 = e + 32 < len ? sm_ld(w, e + 32) : 0.0;
        const double x2 = e + 64 < len ? sm_ld(w, e + 64) : 0.0;
        const double x3 = e + 96 < len ? sm_ld(w, e + 96) : 0.0;
        a0 = __dadd_rn(a0, x0);
        a1 = __dadd_rn(a1, x1);
        a2 = __dadd_rn(a2, x2);
        a3 = __dadd_rn(a3, x3);
      }
      const double a = warp_sum(__dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3)));
      if (lane == 0) sm_st(wt, to.T + (q - h.q0), a);
    }
    bar_sync_n(kBarFold, kFoldThreads);
  }
  if (has_next) piece_request(to, next, fr.pc);
  uint32_t q0 = 0;
  do {  // (runs once for an empty list: the first batch of the next tile still has to be requested)
    uint32_t nxt[kPre];
    if (q0 + kPre < h.L) {
      fold_request(to, h, q0 + kPre, nxt);
    } else if (has_next) {
      fold_request(to, next, 0, nxt);
    } else {
#pragma unroll
      for (int q = 0; q < kPre; ++q) nxt[q] = kEntPad;
    }
#pragma unroll
    for (int q = 0; q < kPre; ++q)
      if (fr.ent[q] != kEntPad) fold_entry(fr.ent[q], wt, s.acc);
#pragma unroll
    for (int q = 0; q < kPre; ++q) fr.ent[q] = nxt[q];
    q0 += kPre;
  } while (q0 < h.L);
}

// Fold warps, plain form (pass 2): words requested when they are needed.
__device__ __forceinline__ void tile_node_sums_plain(const TileOp& to, const TileSmem& s, SmArr wt, const TileHdr& h) {
  const int ftid = threadIdx.x - kStreamThreads, lane = ftid & 31, fwarp = ftid >> 5;
  if (h.q1 > h.q0) {
    for (uint32_t q = h.q0 + fwarp; q < h.q1; q += kFoldWarps) {
      const uint32_t pc = __ldg(to.piece + q);
      const uint32_t first = pc & 0xffffu, len = (pc >> 16) + 1;
      const SmArr w{wt.a + first * 8u};
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four independent lane-strided chains, combined in a fixed order
      for (uint32_t e = lane; e < len; e += 128) {
        const double x0 = sm_ld(w, e);
        const double x1 = e + 32 < len ? sm_ld(w, e + 32) : 0.0;
        const double x2 = e + 64 < len ? sm_ld(w, e + 64) : 0.0;
        const double x3 = e + 96 < len ? sm_ld(w, e + 96) : 0.0;
        a0 = __dadd_rn(a0, x0);
        a1 = __dadd_rn(a1, x1);
        a2 = __dadd_rn(a2, x2);
        a3 = __dadd_rn(a3, x3);
      }
      const double a = warp_sum(__dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3)));
      if (lane == 0) sm_st(wt, to.T + (q - h.q0), a);
    }
    bar_sync_n(kBarFold, kFoldThreads);
  }
  const uint32_t* mine = to.lent + h.e0 + ftid;
  for (uint32_t q0 = 0; q0 < h.L; q0 += kPre) {  // kPre list words requested together, then folded in order
    uint32_t ent[kPre];
#pragma unroll
    for (int q = 0; q < kPre; ++q) ent[q] = q0 + q < h.L ? __ldg(mine + (size_t)(q0 + q) * kFoldThreads) : kEntPad;
#pragma unroll
    for (int q = 0; q < kPre; ++q)
      if (ent[q] != kEntPad) fold_entry(ent[q], wt, s.acc);
  }
}

struct TileCtx {
  uint32_t alo, ahi;  // owned arcs
  uint32_t ulo, uhi;  // owned node rows (node index)
};
__device__ __forceinline__ TileCtx tile_ctx(const IncidenceOp& op, const TileOp& to) {
  TileCtx c;
  cta_chunk(op.m, c.alo, c.ahi);
  c.ulo = min(op.p, (to.fab.rank * gridDim.x + blockIdx.x) * to.R);
  c.uhi = min(op.p, c.ulo + to.R);
  return c;
}

// writes this CTA's p partial sums (s.acc) to the owners' buffers (parity `par`); caller synchronised before
__device__ __forceinline__ void publish_tile_partials(const IncidenceOp& op, const TileOp& to, const TileSmem& s, uint32_t par) {
  const Fabric& f = to.fab;
  const uint32_t src = f.rank * gridDim.x + blockIdx.x;
  for (int r = 0; r < f.world; ++r) {
    const uint32_t u0 = r * f.Bp, u1 = min(op.p, u0 + f.Bp);
    double* dst = f.partials[r] + ((size_t)par * f.Gtot + src) * f.Bp;
    for (uint32_t u = u0 + threadIdx.x; u < u1; u += kBlock) __stcg(dst + (u - u0), sm_ld(s.acc, u));
  }
}
// T_u = sum over the partials of all CTAs of all ranks in a fixed order (one warp per owned node, lanes stride the
// sources, xor tree)
__device__ __forceinline__ double tile_node_total(const TileOp& to, uint32_t par, uint32_t u, int lane) {
  const Fabric& f = to.fab;
  const double* Pin = f.partials[f.rank] + (size_t)par * f.Gtot * f.Bp + (u - f.rank * f.Bp);
  double a = 0.0;
  for (uint32_t c0 = lane; c0 < f.Gtot; c0 += 256) {  // eight independent loads in flight per lane
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t c = c0 + 32 * q;
      v[q] = c < f.Gtot ? __ldcg(Pin + (size_t)c * f.Bp) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) a = __dadd_rn(a, v[q]);
  }
  return warp_sum(a);
}
// node value of an owned row -> every rank's replica (parity `par`)
__device__ __forceinline__ void publish_node(const TileOp& to, uint32_t p, uint32_t par, uint32_t u, double w) {
  for (int r = 0; r < to.fab.world; ++r) __stcg(to.fab.nodebuf[r] + (size_t)par * p + u, w);
}
__device__ __forceinline__ const double* local_nodebuf(const TileOp& to, uint32_t p, uint32_t par) {
  return to.fab.nodebuf[to.fab.rank] + (size_t)par * p;
}

__device__ __forceinline__ void st_relaxed_sys_v4(uint4* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_acquire_sys_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.acquire.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// Barrier + all-reduce over the CTAs of ALL ranks (world > 1): the same self-validating lines as grid_sync, but every
// CTA pushes its line into every rank's replica over NVLink and polls only its local replica.  System-scope
// fences order the peer-memory data written before the call (partials, node values) before the slot, and the local reads
// after the poll.  The payloads are added in slot order: the result is identical in every CTA of every rank.
// FENCED = false: pure barrier + all-reduce, for the alpha reduction of pass 1 -- no peer memory is written in the phase
// before it and none is read in the phase after it, so the two system-scope fences (an NVLink round trip each) are skipped.
template <bool REDUCE, bool FENCED>
__device__ __forceinline__ double fabric_sync(double v, const TileOp& to, unsigned int& epoch, CtaShared& sh) {
  __shared__ double chunk_sum[kWarps];
  const Fabric& f = to.fab;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  epoch += 1;
  if (REDUCE) {
    v = warp_sum(v);
    if (lane == 0) sh.warp_part[warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
    double t = 0.0;
    if (REDUCE) {
      t = lane < kWarps ? sh.warp_part[lane] : 0.0;
      t = warp_sum(t);
    }
    const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
    const uint32_t src = f.rank * gridDim.x + blockIdx.x;
    if (FENCED) fence_acq_rel_sys();
    // a whole 128-byte line per (source CTA, target rank): eight lanes store its eight atoms in one instruction, four ranks
    // per instruction (as in grid_sync: partially written sectors and packed slots were the cost of the first form)
    for (int r0 = 0; r0 < f.world; r0 += 4) {
      const int r = r0 + (lane >> 3);
      if (r < f.world)
        st_relaxed_sys_v4(f.slots[r] + ((size_t)(epoch & 1u) * f.Gtot + src) * kSlotAtoms + (lane & 7),
                          make_uint4((unsigned)bits, epoch, (unsigned)(bits >> 32), epoch));
    }
  }
  // The slots are polled in chunks of 160 (five per lane); chunk w by warp w, all chunks at the same time (8 ranks: 1184
  // slots, one L2 round trip instead of eight).  Slot order inside a chunk is lane-strided, the lanes are combined by the
  // xor tree and the chunk sums are added in chunk order: a fixed order, identical in every CTA of every rank.
  const uint32_t nchunks = (f.Gtot + 159) / 160;
  const uint4* slots = f.slots[f.rank] + (size_t)(epoch & 1u) * f.Gtot * kSlotAtoms + (blockIdx.x & (kSlotAtoms - 1));
  for (uint32_t ch = warp; ch < nchunks; ch += kWarps) {
    const uint32_t base = ch * 160;
    uint4 q4[5];
    unsigned int spins = 0;
    for (;;) {
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const uint32_t i = base + lane + 32 * q;
        // FENCED: acquire loads pair with the peers' fence + store (a trailing fence would also wait for the acknowledgement
        // of our own slot stores, a whole NVLink round trip)
        const uint4* src_line = slots + (size_t)(i < f.Gtot ? i : f.Gtot - 1) * kSlotAtoms;  // one atom of the source's line
        q4[q] = FENCED ? ld_acquire_sys_v4(src_line) : ld_relaxed_sys_v4(src_line);
      }
      bool ok = true;
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const uint32_t i = base + lane + 32 * q;
        ok = ok & ((i >= f.Gtot) | ((q4[q].y == epoch) & (q4[q].w == epoch)));
      }
      if (ok) break;
      if (++spins > kSpinLimit) __trap();
    }
    __syncwarp();
    if (REDUCE) {
      double c = 0.0;
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (base + lane + 32 * q < f.Gtot)
          c = __dadd_rn(c, __longlong_as_double((long long)(((unsigned long long)q4[q].z << 32) | q4[q].x)));
      c = warp_sum(c);
      if (lane == 0) chunk_sum[ch % kWarps] = c;  // (more than kWarps chunks never happens: Gtot <= 8 * 148)
    }
  }
  __syncthreads();
  double s = 0.0;
  if (REDUCE)
    for (uint32_t ch = 0; ch < nchunks; ++ch) s = __dadd_rn(s, chunk_sum[ch]);
  __syncthreads();  // chunk_sum may be rewritten by the next call
  return s;
}
// the barrier of the tiled kernels: the single-GPU grid barrier, or the fabric-wide one
template <bool REDUCE, bool FENCED = true>
__device__ __forceinline__ double tile_sync(double v, const TileOp& to, const GridSync& gs, unsigned int& epoch, CtaShared& sh) {
  if (to.fab.world > 1) return fabric_sync<REDUCE, FENCED>(v, to, epoch, sh);
  return grid_sync<REDUCE, FENCED>(v, gs, epoch, sh);
}

// Tile loop shared by every phase that produces a new vector, warp-specialised: the STREAM warps (0..7) move the chunk
// through registers in batches of BATCH arcs -- `issue(i0, regs)` starts the global loads of batch [i0, i0 + BATCH),
// `consume(i0, t0, regs, wt)` computes the new arc values, stores them and writes them to the tile buffer wt[i - t0] -- while
// the FOLD warps (8..15) add the node sums of the PREVIOUS tile into s.acc from the other tile buffer.  The HBM stream and
// the (shared-memory bound) list walk therefore overlap; before, they alternated and the walk was 57 % of the phase.
// Inside issue / consume a stream thread addresses arcs i0 + q * kStreamThreads + threadIdx.x.  The caller has zeroed s.acc
// and synchronised; on return every fold is complete (CTA-wide barrier).
template <int BATCH, class REGS, class ISSUE, class CONSUME>
__device__ __forceinline__ void tile_loop(const TileOp& to, const TileSmem& s, const TileCtx& c, ISSUE issue, CONSUME consume,
                                          const Trace* tr = nullptr, int tr_step = -1) {
  const uint32_t tile0 = blockIdx.x * to.ntile;
  uint32_t ntiles = 0;
  if (c.alo < c.ahi) ntiles = min(to.ntile, (c.ahi - c.alo + to.T - 1) / to.T);
  const bool timed = tr != nullptr && tr->buf != nullptr;
  if (threadIdx.x < kStreamThreads) {
    long long c_stream = 0, c_wait = 0, t_a = 0;
    // Flat loop over the batches of the chunk with two register sets used alternately (no copies): batch bi + 1 is requested
    // before batch bi is consumed, across tile boundaries.
    const uint32_t nb = c.alo < c.ahi ? (c.ahi - c.alo + BATCH - 1) / BATCH : 0, bpt = to.T / BATCH;
    uint32_t t = 0, in_tile = 0;  // tile of the current batch, its position inside the tile
    auto step = [&](uint32_t bi, REGS& mine, REGS& other) __attribute__((always_inline)) {
      const uint32_t i0 = c.alo + bi * BATCH;
      if (in_tile == 0 && t >= 2) {  // the fold of tile t - 2 has left this buffer
        if (timed) t_a = clock64();
        bar_sync_n(kBarEmpty + (t & 1u), kBlock);
        if (timed) c_wait += clock64() - t_a;
      }
      if (bi + 1 < nb) issue(i0 + BATCH, other);
      consume(i0, c.alo + t * to.T, mine, SmArr{s.wt.a + (t & 1u) * s.wt_stride});
      if (++in_tile == bpt || bi + 1 == nb) {
        bar_arrive_n(kBarFull + (t & 1u), kBlock);
        ++t;
        in_tile = 0;
      }
    };
    if (timed) c_stream = -clock64();
    REGS ra, rb;
    if (nb) issue(c.alo, ra);
    for (uint32_t bi = 0; bi < nb; bi += 2) {
      step(bi, ra, rb);
      if (bi + 1 < nb) step(bi + 1, rb, ra);
    }
    // drain: every arrival of the fold warps is matched by a wait, so that the barriers are clean for the next call
    for (uint32_t u = ntiles > 2 ? ntiles - 2 : 0; u < ntiles; ++u) bar_sync_n(kBarEmpty + (u & 1u), kBlock);
    if (timed) {
      trace_value(*tr, tr_step, 16, c_stream + clock64());
      trace_value(*tr, tr_step, 17, c_wait);
    }
  } else {
    TileHdr hdr = tile_hdr(to, tile0);
    FoldRegs fr;
    if (ntiles) {  // the first tile's list batch and piece words are on their way while the stream warps fill the buffer
      fold_request(to, hdr, 0, fr.ent);
      piece_request(to, hdr, fr.pc);
    }
    for (uint32_t t = 0; t < ntiles; ++t) {
      TileHdr next = hdr;
      if (t + 1 < ntiles) next = tile_hdr(to, tile0 + t + 1);
      const SmArr wt{s.wt.a + (t & 1u) * s.wt_stride};
      bar_sync_n(kBarFull + (t & 1u), kBlock);
      tile_node_sums(to, s, wt, hdr, next, t + 1 < ntiles, fr);
      bar_arrive_n(kBarEmpty + (t & 1u), kBlock);
      hdr = next;
    }
  }
  __syncthreads();
}

// The first form of the tile loop, kept for PASS 2: one register set copied per batch, conditional loads, fold words requested
// when needed.  Measured on B200 (k = 500): with the loop above pass 2 is 7-10 % slower at 20M / 50M arcs (its six input
// streams already keep the memory system busy; requests issued ahead by the fold warps delay them), pass 1 9-12 % faster.
template <int BATCH, class REGS, class ISSUE, class CONSUME>
__device__ __forceinline__ void tile_loop_plain(const TileOp& to, const TileSmem& s, const TileCtx& c, ISSUE issue, CONSUME consume,
                                          const Trace* tr = nullptr, int tr_step = -1) {
  const uint32_t tile0 = blockIdx.x * to.ntile;
  uint32_t ntiles = 0;
  if (c.alo < c.ahi) ntiles = min(to.ntile, (c.ahi - c.alo + to.T - 1) / to.T);
  const bool timed = tr != nullptr && tr->buf != nullptr;
  if (threadIdx.x < kStreamThreads) {
    long long c_stream = 0, c_wait = 0, t_a = 0, t_b = 0;
    REGS cur;
    if (ntiles) issue(c.alo, cur);
    for (uint32_t t = 0; t < ntiles; ++t) {
      const uint32_t t0 = c.alo + t * to.T, t1 = min(c.ahi, t0 + to.T);
      const SmArr wt{s.wt.a + (t & 1u) * s.wt_stride};
      if (timed) t_a = clock64();
      if (t >= 2) bar_sync_n(kBarEmpty + (t & 1u), kBlock);  // the fold of tile t - 2 has left this buffer
      if (timed) { t_b = clock64(); c_wait += t_b - t_a; }
      for (uint32_t i0 = t0; i0 < t1; i0 += BATCH) {
        REGS nxt;
        if (i0 + BATCH < c.ahi) issue(i0 + BATCH, nxt);  // (two batches ahead was tried: the third register set spills)
        consume(i0, t0, cur, wt);
        cur = nxt;
      }
      bar_arrive_n(kBarFull + (t & 1u), kBlock);
      if (timed) c_stream += clock64() - t_b;
    }
    // drain: every arrival of the fold warps is matched by a wait, so that the barriers are clean for the next call
    for (uint32_t t = ntiles > 2 ? ntiles - 2 : 0; t < ntiles; ++t) bar_sync_n(kBarEmpty + (t & 1u), kBlock);
    if (timed) {
      trace_value(*tr, tr_step, 16, c_stream);
      trace_value(*tr, tr_step, 17, c_wait);
    }
  } else {
    TileHdr hdr = tile_hdr(to, tile0);
    for (uint32_t t = 0; t < ntiles; ++t) {
      TileHdr next = hdr;
      if (t + 1 < ntiles) next = tile_hdr(to, tile0 + t + 1);
      const SmArr wt{s.wt.a + (t & 1u) * s.wt_stride};
      bar_sync_n(kBarFull + (t & 1u), kBlock);
      tile_node_sums_plain(to, s, wt, hdr);
      bar_arrive_n(kBarEmpty + (t & 1u), kBlock);
      hdr = next;
    }
  }
  __syncthreads();
}

// node partial sums of an arbitrary arc vector X (init: b) over the CTA's chunk
__device__ __forceinline__ void tile_sums_of(const IncidenceOp& op, const TileOp& to, const TileSmem& s, const TileCtx& c,
                                             const double* X) {
  for (uint32_t u = threadIdx.x; u < op.p; u += kBlock) sm_st(s.acc, u, 0.0);
  __syncthreads();
  struct R {
    double x[kUnroll];
  };
  tile_loop<kUnroll * kStreamThreads, R>(
      to, s, c,
      [&](uint32_t i0, R& r) __attribute__((always_inline)) {
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          const uint32_t i = i0 + q * kStreamThreads + threadIdx.x;
          r.x[q] = __ldg(X + min(i, c.ahi - 1));
        }
      },
      [&](uint32_t i0, uint32_t t0, const R& r, SmArr wt) __attribute__((always_inline)) {
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          const uint32_t i = i0 + q * kStreamThreads + threadIdx.x;
          if (i < c.ahi) sm_st(wt, i - t0, r.x[q]);
        }
      });
}

// =============================================================================================
// pass 1 / one-pass basis generation, streaming with tiled node sums (single GPU, persistent, cooperative)
// =============================================================================================
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_tiled_kernel(const IncidenceOp op, const TileOp to, const Pass1Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  const TileSmem s = carve_tiles(smem, op.p, to.T, false);
  const TileCtx c = tile_ctx(op, to);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t m = op.m, p = op.p;

  unsigned int epoch = a.st->epoch;
  int steps = 0, status = ST_RUNNING, rot = 0;
  double sc = 1.0, sp = 1.0, bp = 0.0, bnorm = 0.0;
  GridSync gs = a.gs;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  {
    // K0: ||b||, W_cur = b, W_prev = 0, partial node sums of b for step 0
    double* Wp = pick(0);
    double* Wc = pick(1);
    double acc = 0.0;
    for (uint32_t i = c.alo + threadIdx.x; i < c.ahi; i += kBlock) {
      const double bi = __ldg(a.b + i);
      __stcg(Wc + i, bi);
      __stcg(Wp + i, 0.0);
      acc = fma(bi, bi, acc);
    }
    for (uint32_t u = c.ulo + threadIdx.x; u < c.uhi; u += kBlock) {
      const double bi = __ldg(a.b + m + u);
      __stcg(Wc + m + u, bi);
      __stcg(Wp + m + u, 0.0);
      publish_node(to, p, 1, u, bi);  // parity of "step -1"
      acc = fma(bi, bi, acc);
    }
    tile_sums_of(op, to, s, c, a.b);
    publish_tile_partials(op, to, s, 0);
    bnorm = sqrt(tile_sync<true>(acc, to, a.gs, epoch, sh));
    if (bnorm <= a.tol) status = ST_ZERO_B;
    sc = 1.0 / bnorm;
  }
  if (status == ST_RUNNING) {
    for (int j = 0; j < a.j_end; ++j) {
      const double* Wp = pick(rot);
      const double* Wc = pick((rot + 1) % 3);
      double* Wn = pick((rot + 2) % 3);
      const double* Xnode = local_nodebuf(to, p, (j + 1) & 1);  // node part of the current vector
      double* Vcol = WITH_V ? a.V + (size_t)j * a.ldv : nullptr;
      gs.trace_step = j;
      trace_mark(gs.trace, j, 0);

      // ---------------- phase A: w~ = A v - beta_{j-1} v_{j-1}, alpha partial
      for (uint32_t u = threadIdx.x; u < p; u += kBlock) sm_st(s.node, u, __dmul_rn(__ldcg(Xnode + u), sc));
      __syncthreads();
      trace_mark(gs.trace, j, 1);
      if (WITH_V) {  // node part of the basis column: replicated on every rank, each CTA writes its share
        uint32_t vlo, vhi;
        cta_chunk(p, vlo, vhi);
        for (uint32_t u = vlo + threadIdx.x; u < vhi; u += kBlock) __stcs(Vcol + m + u, sm_ld(s.node, u));
      }
      double acc = 0.0;
      for (uint32_t u = c.ulo + warp; u < c.uhi; u += kWarps) {  // node rows of the owned block
        const double t = __dmul_rn(sc, tile_node_total(to, j & 1, u, lane));
        if (lane == 0) {
          const double v = sm_ld(s.node, u);
          const double vp = __dmul_rn(__ldcg(Wp + m + u), sp);
          const double wt = rec_sub(t, bp, vp);
          acc = fma(v, wt, acc);
          __stcg(Wn + m + u, wt);
        }
      }
      trace_mark(gs.trace, j, 2);
      for (uint32_t base = c.alo; base < c.ahi; base += kUnroll * kBlock) {
        double wc[kUnroll], wp[kUnroll], dd[kUnroll];
        uint32_t tl[kUnroll], hd[kUnroll];
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          const uint32_t i = min(base + q * kBlock + threadIdx.x, c.ahi - 1);  // (unconditional: see phase B)
          wc[q] = __ldcg(Wc + i);
          wp[q] = __ldcg(Wp + i);
          dd[q] = __ldg(op.d + i);
          tl[q] = __ldg(op.tail + i);
          hd[q] = __ldg(op.head + i);
        }
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          const uint32_t i = base + q * kBlock + threadIdx.x;
          if (i < c.ahi) {
            const double v = __dmul_rn(wc[q], sc);
            const double vp = __dmul_rn(wp[q], sp);
            const double wt = rec_sub(arc_row(dd[q], v, tl[q], hd[q], sm_ld(s.node, tl[q]), sm_ld(s.node, hd[q])), bp, vp);
            acc = fma(v, wt, acc);
            __stcg(Wn + i, wt);
            if (WITH_V) __stcs(Vcol + i, v);
          }
        }
      }
      trace_mark(gs.trace, j, 3);
      gs.trace_base = 4;
      const double alpha = tile_sync<true, false>(acc, to, gs, epoch, sh);

      // ---------------- phase B: w = w~ - alpha v, beta partial, partial node sums of w
      acc = 0.0;
      for (uint32_t u = threadIdx.x; u < p; u += kBlock) sm_st(s.acc, u, 0.0);  // (aliases s.node: phase A is over)
      for (uint32_t u = c.ulo + threadIdx.x; u < c.uhi; u += kBlock) {
        const double v = __dmul_rn(__ldcg(Wc + m + u), sc);
        const double w = rec_sub(__ldcg(Wn + m + u), alpha, v);
        __stcg(Wn + m + u, w);
        publish_node(to, p, j & 1, u, w);
        acc = fma(w, w, acc);
      }
      __syncthreads();  // accumulators are zero before the first fold
      struct RB {
        double wn[kUnrollB], wc[kUnrollB];
      };
      tile_loop<kUnrollB * kStreamThreads, RB>(
          to, s, c,
          [&](uint32_t i0, RB& r) __attribute__((always_inline)) {
#pragma unroll
            for (int q = 0; q < kUnrollB; ++q) {
              // (unconditional: a conditionally written register stays live around the whole loop; the batch past the end of
              // the chunk re-reads its last arc and is discarded by consume)
              const uint32_t i = min(i0 + q * kStreamThreads + threadIdx.x, c.ahi - 1);
              r.wn[q] = __ldcg(Wn + i);
              r.wc[q] = __ldcg(Wc + i);
            }
          },
          [&](uint32_t i0, uint32_t t0, const RB& r, SmArr wt) __attribute__((always_inline)) {
#pragma unroll
            for (int q = 0; q < kUnrollB; ++q) {
              const uint32_t i = i0 + q * kStreamThreads + threadIdx.x;
              if (i < c.ahi) {
                const double w = rec_sub(r.wn[q], alpha, __dmul_rn(r.wc[q], sc));
                __stcg(Wn + i, w);
                sm_st(wt, i - t0, w);
                acc = fma(w, w, acc);
              }
            }
          },
          &gs.trace, j);
      trace_mark(gs.trace, j, 8);
      publish_tile_partials(op, to, s, (j + 1) & 1);
      gs.trace_base = 9;
      const double beta = sqrt(tile_sync<true>(acc, to, gs, epoch, sh));

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
// pass 2, streaming with tiled node sums: one sweep and one grid barrier per step
// =============================================================================================
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_tiled_kernel(const IncidenceOp op, const TileOp to, const Pass2Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  const TileSmem s = carve_tiles(smem, op.p, to.T, true);
  const TileCtx c = tile_ctx(op, to);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t m = op.m, p = op.p;
  uint32_t vlo, vhi;  // share of the (replicated) node part of x / V this CTA writes
  cta_chunk(p, vlo, vhi);
  unsigned int epoch = a.st->epoch;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  int rot = 0;
  double sc_cur = 1.0 / a.b_norm;  // scale of the vector whose un-normalised partial node sums are in `partials`
  {
    // v_1 = b * (1/||b||), x = y_0 v_1; partial node sums of b
    const double inv = 1.0 / a.b_norm;
    const double y0 = __ldg(a.y);
    double* Vp = buf0;
    double* Vc = buf1;
    for (uint32_t u = c.ulo + threadIdx.x; u < c.uhi; u += kBlock) {
      const double v = __dmul_rn(__ldg(a.b + m + u), inv);
      __stcg(Vc + m + u, v);
      __stcg(Vp + m + u, 0.0);
      publish_node(to, p, 1, u, v);
    }
    for (uint32_t u = vlo + threadIdx.x; u < vhi; u += kBlock) {
      const double v = __dmul_rn(__ldg(a.b + m + u), inv);
      __stcg(a.x + m + u, __dmul_rn(v, y0));
      if (WITH_V) __stcs(a.V + m + u, v);
    }
    for (uint32_t i = c.alo + threadIdx.x; i < c.ahi; i += kBlock) {
      const double v = __dmul_rn(__ldg(a.b + i), inv);
      __stcg(Vc + i, v);
      __stcg(Vp + i, 0.0);
      __stcg(a.x + i, __dmul_rn(v, y0));
      if (WITH_V) __stcs(a.V + i, v);
    }
    // node sums are taken over the UN-normalised vector and scaled afterwards, exactly as pass 1 does with its lazily
    // scaled w (bit-identical node rows)
    tile_sums_of(op, to, s, c, a.b);
    publish_tile_partials(op, to, s, 0);
    tile_sync<false>(0.0, to, a.gs, epoch, sh);
  }
  for (int j = 0; j + 1 < a.steps; ++j) {
    const double* Vp = pick(rot);
    const double* Vc = pick((rot + 1) % 3);
    double* Vn = pick((rot + 2) % 3);
    const double* Xnode = local_nodebuf(to, p, (j + 1) & 1);
    double* Vcol = WITH_V ? a.V + (size_t)(j + 1) * a.ldv : nullptr;
    const double alpha = __ldg(a.alphas + j);
    const double beta = __ldg(a.betas + j);
    const double bp = j == 0 ? 0.0 : __ldg(a.betas + j - 1);
    const double sinv = 1.0 / beta;
    const double yj = __ldg(a.y + j + 1);

    for (uint32_t u = threadIdx.x; u < p; u += kBlock) {
      sm_st(s.node, u, __ldcg(Xnode + u));
      sm_st(s.acc, u, 0.0);
    }
    __syncthreads();
    if (j > 0) {
      // the staged node values are v_{j+1}, regenerated by the previous step: their share of x (and of the basis column)
      // is added here, by every rank for its replica (each CTA its share of the nodes)
      const double yprev = __ldg(a.y + j);
      for (uint32_t u = vlo + threadIdx.x; u < vhi; u += kBlock) {
        const double vn = sm_ld(s.node, u);
        __stcg(a.x + m + u, __dadd_rn(__ldcg(a.x + m + u), __dmul_rn(yprev, vn)));
        if (WITH_V) __stcs(a.V + (size_t)j * a.ldv + m + u, vn);
      }
    }
    for (uint32_t u = c.ulo + warp; u < c.uhi; u += kWarps) {  // node rows of the owned block
      const double t = __dmul_rn(sc_cur, tile_node_total(to, j & 1, u, lane));
      if (lane == 0) {
        const double w = rec_sub(rec_sub(t, bp, __ldcg(Vp + m + u)), alpha, sm_ld(s.node, u));
        const double vn = __dmul_rn(w, sinv);
        __stcg(Vn + m + u, vn);
        publish_node(to, p, j & 1, u, vn);
      }
    }
    __syncthreads();  // node rows are done with s.node reads of other warps; accumulators are zero
    struct R2 {
      double vc[kUnroll2], vp[kUnroll2], dd[kUnroll2], xx[kUnroll2];
      uint32_t tl[kUnroll2], hd[kUnroll2];
    };
    tile_loop_plain<kUnroll2 * kStreamThreads, R2>(
        to, s, c,
        [&](uint32_t i0, R2& r) __attribute__((always_inline)) {
#pragma unroll
          for (int q = 0; q < kUnroll2; ++q) {
            const uint32_t i = i0 + q * kStreamThreads + threadIdx.x;
            if (i < c.ahi) {
              r.vc[q] = __ldcg(Vc + i);
              r.vp[q] = __ldcg(Vp + i);
              r.xx[q] = __ldcg(a.x + i);
              r.dd[q] = __ldg(op.d + i);
              r.tl[q] = __ldg(op.tail + i);
              r.hd[q] = __ldg(op.head + i);
            }
          }
        },
        [&](uint32_t i0, uint32_t t0, const R2& r, SmArr wt) __attribute__((always_inline)) {
#pragma unroll
          for (int q = 0; q < kUnroll2; ++q) {
            const uint32_t i = i0 + q * kStreamThreads + threadIdx.x;
            if (i < c.ahi) {
              const double v = r.vc[q];
              const double w = rec_sub(
                  rec_sub(arc_row(r.dd[q], v, r.tl[q], r.hd[q], sm_ld(s.node, r.tl[q]), sm_ld(s.node, r.hd[q])), bp, r.vp[q]), alpha, v);
              const double vn = __dmul_rn(w, sinv);
              __stcg(Vn + i, vn);
              __stcg(a.x + i, __dadd_rn(r.xx[q], __dmul_rn(yj, vn)));
              if (WITH_V) __stcs(Vcol + i, vn);
              sm_st(wt, i - t0, w);
            }
          }
        });
    publish_tile_partials(op, to, s, (j + 1) & 1);
    tile_sync<false>(0.0, to, a.gs, epoch, sh);
    rot = (rot + 1) % 3;
    sc_cur = sinv;
  }
  if (a.steps > 1) {  // node part of the last regenerated vector v_steps (published by the last step, parity (steps - 2) & 1)
    const double* Xnode = local_nodebuf(to, p, (a.steps - 2) & 1);
    const double ylast = __ldg(a.y + a.steps - 1);
    for (uint32_t u = vlo + threadIdx.x; u < vhi; u += kBlock) {
      const double vn = __ldcg(Xnode + u);
      __stcg(a.x + m + u, __dadd_rn(__ldcg(a.x + m + u), __dmul_rn(ylast, vn)));
      if (WITH_V) __stcs(a.V + (size_t)(a.steps - 1) * a.ldv + m + u, vn);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.st->epoch = epoch;
}

}  // namespace tpl
