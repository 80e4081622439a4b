// tpl_blocks.cuh -- BLOCKED streaming kernels of the KKT incidence operator: the 2-D node-block partition of the cell
// kernels (tpl_cells.cuh) applied to the streaming regime, with an asynchronous bulk-copy input ring.
//
// Why: the tiled kernels of tpl_tiles.cuh keep two p-long f64 arrays (node values, node accumulators) in
// shared memory -- 184 KB at 50M arcs -- which leaves 2048-arc tiles, and they stage six input streams through registers
// with 2 arcs per thread in flight.  Both made pass 2 latency-bound at 0.59 of the HBM roofline.  Here
//   * the nodes are cut into GR contiguous TAIL blocks (balanced by out-degree) and GC contiguous HEAD blocks (balanced by
//     in-degree); CTA (r, c) owns the arcs with tail in block r and head in block c.  It therefore stages only the node
//     values of two blocks (p/GR + p/GC instead of p) and accumulates only their sums: 31 KB instead of 184 KB at 50M arcs;
//   * the operator (d, packed local tail/head word) and the arc part of every Lanczos vector live in HBM in CELL ORDER
//     (cell after cell, inside a cell by tail, every cell padded to a multiple of 128 arcs); b is gathered into that order
//     once at the start of a pass and x scattered back once at its end (gidx);
//   * the sweeps that also fold node sums (phase B of pass 1, the single sweep of pass 2) read their inputs through a ring
//     of shared-memory slots filled by cp.async.bulk (global -> shared, completion on an mbarrier): every compute warp owns
//     kBRing slots of 128 arcs, its lane 0 issues the bulk copies of stage k + ring as soon as stage k is consumed.  No
//     registers are tied up by loads in flight (3 x 8 x 4.6 KB = 110 KB per SM in pass 2);
//   * the node sums of a tile are folded by the fold warps exactly as in tpl_tiles.cuh (same list format, over LOCAL node
//     ids: tails [0, PT), heads [PT, PT + PH)), tiles of up to 4096 arcs;
//   * per step a CTA publishes PT + PH partial sums, destination-indexed: the owner of a node row finds the GC + GR
//     contributions of every rank in consecutive words and adds them in a fixed order.  A node collects 24 partials per
//     rank instead of 148, and a sharded operator moves 24 p doubles per rank per step over NVLink instead of 148 p.
// Arithmetic: the same per-element expressions as every other shape (rec_sub, the reference's CSC order inside an arc row,
// lazy scaling by the rounded reciprocal); pass 1, the one-pass variant and pass 2 share the lists and therefore regenerate
// bit-identical vectors.  alpha / beta differ from the other shapes by rounding only (different fixed summation order).
#pragma once
#include "tpl_tiles.cuh"

namespace tpl {

constexpr int kBStage = 128;                 // arcs of one warp-stage (4 per lane)
// Warp roles of the folding sweeps: warps [0, NCW) compute, the rest fold.  Pass 2 (8 / 8): its compute side is the heavy one
// per arc but HBM-bound; pass 1 (12 / 4): phase B moves only 28 B per arc, its compute warps were latency-bound on their
// dependent chains (ring wait -> loads -> run sums through five shuffle rounds) with two warps per scheduler, while the list
// walk alone costs 8 us of a 250-us step at 20M arcs -- so the fold gets 4 warps, each thread walking two of the 256 slices.
constexpr int kBComputeWarps1 = 12, kBComputeWarps2 = 8;
constexpr int kBSlices = kFoldThreads;        // list slices per tile (host format; independent of how many threads walk them)
constexpr int kBMaxRing = 4;
constexpr int kBMaxList = 8;
constexpr int kBMaxTileBufs = 4;
constexpr int kBBarFull = 6, kBBarEmpty = 10;  // named barriers of the tile buffers (tpl_tiles.cuh uses 1..5)
constexpr uint32_t kBLoop = 0x80000000u;     // th word: self-loop or padding (no incidence entries)
constexpr uint32_t kBTailFirst = 0x40000000u;  // th word: global tail index < global head index (CSC accumulation order)
constexpr uint32_t kBPad = 0xffffffffu;      // gidx of a padding slot
constexpr uint32_t kBNoPiece = 0xffffffffu;  // unused run descriptor of a stage
constexpr uint32_t kBPieceMin = 24;          // shortest same-tail run (inside one stage) summed by the compute warp

struct BlockOp {
  uint32_t GR, GC;   // tail blocks x head blocks; the grid has GR * GC CTAs, CTA c = r * GC + cc
  uint32_t PT, PH;   // largest tail / head block: local node ids are tails [0, PT), heads [PT, PT + PH)
  uint32_t Mpad;     // arcs in cell order, padding included; the node part of a cell-order vector starts here
  uint32_t ring1, ring2, ring2v;  // ring slots per compute warp: pass 1, pass 2, pass 2 with a basis
  uint32_t lblk;     // bytes of the largest tile list block ((L + 1) rows of kFoldThreads words): size of a list buffer
  uint32_t nl;       // list buffers (2 .. kBMaxList): the list block of a tile is bulk-copied nl - 1 tiles ahead
  uint32_t ntb;      // tile buffers (2 .. kBMaxTileBufs): the compute warps run at most ntb tiles ahead of the fold
  uint32_t m;        // arcs in natural order (node part of a natural-order vector starts here)
  uint32_t dbg;      // timing experiments only (results are wrong): 1 = fold warps skip their work, 2 = compute warps skip theirs
  const uint32_t* cell_off;  // [G + 1] first cell-order position of every cell (multiples of kBStage)
  // A block is a run of the ACTIVE nodes of its side (nodes with at least one out-arc / in-arc among this handle's arcs, in
  // ascending order): nodes without arcs on a side -- sinks, sources, and on a sharded handle the tails of other ranks'
  // arcs -- take no shared memory and no partial sums.
  const uint32_t* tbs;       // [GR + 1] first position in tbn of every tail block
  const uint32_t* hbs;       // [GC + 1] first position in hbn of every head block
  const uint32_t* tbn;       // [active tails] node ids, ascending
  const uint32_t* hbn;       // [active heads] node ids, ascending
  const double* d;           // [Mpad] quadratic costs in cell order (0 in padding and beyond the loader's short D)
  const uint32_t* th;        // [Mpad] tail_local | head_local << 15 | kBTailFirst | kBLoop
  const uint32_t* gidx;      // [Mpad] natural arc index of a cell-order position (kBPad in padding)
  const uint4* pdesc;        // [Mpad / 128] per stage up to four same-tail runs: start | (len - 1) << 8 | tile slot << 16 (kBNoPiece: none)
  double* xc;                // [Mpad] arc part of x in cell order (pass 2 workspace)
  TileOp tl;                 // tile lists over local node ids (T, ntile, thdr, lent, piece), R, and the exchange buffers:
                             // tl.fab.partials[rank] is [2][Bp][world * (GC + GR)], destination-indexed
};

struct BlockSmem {
  SmArr node, acc, wt;
  uint32_t wt_stride;
  uint32_t ring;   // shared-window address of the slot area: [compute warp][slot][slot_bytes]
  uint32_t mbar;   // [compute warp][kBMaxRing] mbarriers of the rings, then kBMaxList mbarriers of the list buffers
  uint32_t lst;    // [nl][lblk] list blocks of the tile being folded and of the next ones (bulk-copied nl - 1 tiles ahead)
  uint32_t scr;    // [kBSlices] doubles: a slice's share of a node that earlier slices also hold
  uint32_t hdr;    // [ntile] uint4: the headers of the cell's tiles (copied once per launch: a global load per tile sat on the
                   // fold's critical path with the latency of a saturated memory system)
};
// List format of the blocked kernels (host: build_cell_lists).  A tile's block is (L + 1) rows of kBSlices words,
// slice-interleaved: rows 1..L are the slice's entries
//     minus << 31 | new_node << 30 | slot << 16 | 8 * index into the tile buffer (arcs, then run sums)
// sorted by node and cut into slices of EQUAL length: a node may straddle slices.  The walker adds the values of one node in
// a register and touches shared memory once per node: an entry that opens a new node names the SLOT that takes the finished
// sum of the node before it -- that node's accumulator, or the slice's scratch slot (PL + kBAccPad + slice) when it was the
// slice's first node and earlier slices hold entries of the same node; any other entry names the dummy slot (PL).  The
// slice's FIRST entry has nothing to flush and names its node instead (for the depth phases).  Row 0 = depth | slot of the
// slice's last node << 8.  depth = position in the chain of slices that share the slice's first node: the scratch share is
// added to the node's accumulator `depth` barrier phases after the walk, so the order in which the shares of a node are added
// is fixed and no two threads ever update one accumulator in the same phase.
// Padding (only behind the last entry of a tile) is a harmless entry instead of a branch: it adds the tile's ZERO slot
// (the last of the block_piece_slots(T) words behind the T arc values; written once) to the running sum and names the dummy.
constexpr uint32_t kBEntMinus = 0x80000000u, kBEntNew = 0x40000000u, kBEntNodeShift = 16, kBEntNodeMask = 0x3fffu, kBEntOffMask = 0xffffu;
constexpr uint32_t kBMaxLocalNodes = kBEntNodeMask;  // accumulators + dummy + scratch slots must fit the 14-bit slot field
constexpr uint32_t kBAccPad = 2;  // accumulator slots behind the PL real ones (dummy + alignment)
constexpr int kBPre = 8;          // list entries per fold thread requested together
constexpr int kBFlush = 4;        // the short form of the last batch of a slice
// words of a tile buffer behind its T arc values: at most four run sums per 128-arc stage, then the zero slot (even total)
__host__ __device__ inline uint32_t block_piece_slots(uint32_t T) { return T / 32 + 8; }
__host__ __device__ inline uint32_t block_pad_entry(uint32_t PL, uint32_t T) {
  return (PL << kBEntNodeShift) | ((T + block_piece_slots(T) - 1) * 8u);
}
__host__ __device__ inline size_t block_slot_bytes(int n8, int n4) { return (size_t)n8 * kBStage * 8 + (size_t)n4 * kBStage * 4 + 16; }  // + the stage's run descriptors
// pass 1 never needs node values and accumulators at the same time (they alias), pass 2 needs both
constexpr uint32_t kBMbarBytes = (kWarps * kBMaxRing + kBMaxList) * 8;
__host__ __device__ inline size_t block_smem_bytes(uint32_t PL, uint32_t T, int ring, uint32_t lblk, uint32_t nl, uint32_t ntb, bool pass2, bool with_v, uint32_t ntile) {
  const size_t slot = pass2 ? block_slot_bytes(4, with_v ? 2 : 1) : block_slot_bytes(2, 0);
  return ((pass2 ? 2 : 1) * ((size_t)PL + kBAccPad) + ntb * ((size_t)T + block_piece_slots(T))) * sizeof(double) +
         (size_t)(pass2 ? kBComputeWarps2 : kBComputeWarps1) * ring * slot +
         (size_t)nl * lblk + kBMbarBytes + kFoldThreads * 8 + (size_t)ntile * 16 + 16;  // + 16: the carve-up starts at the next 16-byte boundary
}
__device__ __forceinline__ BlockSmem carve_blocks(double* base, uint32_t PL, uint32_t T, uint32_t lblk, uint32_t nl, uint32_t ntb, uint32_t ntile, bool pass2) {
  // bulk copies need 16-byte aligned shared-memory addresses: PL and T + block_piece_slots(T) are even (host), the base is rounded up
  const uint32_t b = ((uint32_t)__cvta_generic_to_shared(base) + 15u) & ~15u;
  BlockSmem s;
  s.node.a = b;
  s.acc.a = pass2 ? b + (PL + kBAccPad) * 8u : b;
  s.scr = s.acc.a + (PL + kBAccPad) * 8u;  // list slots PL + kBAccPad + slice
  s.wt.a = s.scr + kBSlices * 8u;
  s.wt_stride = (T + block_piece_slots(T)) * 8u;
  s.mbar = s.wt.a + ntb * s.wt_stride;
  s.hdr = s.mbar + kBMbarBytes;
  s.lst = s.hdr + ntile * 16u;
  s.ring = s.lst + nl * lblk;
  return s;
}

// ---------------------------------------------------------------------------- mbarrier / bulk-copy primitives (PTX)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned int spins = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > kSpinLimit) __trap();  // a bulk copy that never lands must not hang the GPU
  }
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); completion is counted on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// orders this thread's earlier generic-proxy accesses (st.global of a vector, ld.shared of a ring slot) before later
// async-proxy accesses (the bulk copies that re-read that vector / refill that slot)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ld / st.shared with a compile-time byte offset (one instruction each, [reg + imm]).  `volatile` keeps them in program order
// relative to the mbarrier wait and the named barriers (volatile asm with a memory clobber); without a clobber of their own
// the compiler is free to schedule the arithmetic of the four arcs of a lane between them.
template <int OFF>
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v));
}
__device__ __forceinline__ double lds64_at(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32_at(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// (A x)_j of an arc row in the reference's CSC accumulation order (arc_row of tpl_kernels.cuh), branch-free: the th word says
// whether the tail column comes first; a - x and a + (-x) are the same IEEE operation, and a self-loop / padding slot adds
// +0.0 twice (its merged incidence entry is an explicit zero).
__device__ __forceinline__ double arc_row_b(double dj, double xj, uint32_t th, double xt, double xh) {
  const double nxh = -xh;
  const bool loop = (th & kBLoop) != 0, tf = (th & kBTailFirst) != 0;
  const double first = loop ? 0.0 : (tf ? xt : nxh);
  const double second = loop ? 0.0 : (tf ? nxh : xt);
  return __dadd_rn(__dadd_rn(__dmul_rn(dj, xj), first), second);
}

struct BlockCtx {
  uint32_t r, cc;     // this CTA's tail / head block
  uint32_t c0, c1;    // its cell-order range (c1 - c0 is a multiple of kBStage)
  uint32_t t0, nt;    // first position (in tbn) and size of its tail block
  uint32_t h0, nh;    // first position (in hbn) and size of its head block
  uint32_t ulo, uhi;  // node rows it owns
  uint32_t nst;       // stages of the cell
  uint32_t ntiles;    // tiles of the cell
};
__device__ __forceinline__ BlockCtx block_ctx(const BlockOp& bo, uint32_t p) {
  BlockCtx c;
  c.r = blockIdx.x / bo.GC;
  c.cc = blockIdx.x % bo.GC;
  c.c0 = __ldg(bo.cell_off + blockIdx.x);
  c.c1 = __ldg(bo.cell_off + blockIdx.x + 1);
  c.t0 = __ldg(bo.tbs + c.r);
  c.nt = __ldg(bo.tbs + c.r + 1) - c.t0;
  c.h0 = __ldg(bo.hbs + c.cc);
  c.nh = __ldg(bo.hbs + c.cc + 1) - c.h0;
  c.ulo = min(p, (bo.tl.fab.rank * gridDim.x + blockIdx.x) * bo.tl.R);
  c.uhi = min(p, c.ulo + bo.tl.R);
  c.nst = (c.c1 - c.c0) / kBStage;
  c.ntiles = (c.c1 - c.c0 + bo.tl.T - 1) / bo.tl.T;
  return c;
}

// node values of the CTA's two blocks -> s.node (local ids), scaled by `sc` (one rounding, the reference's in-place scaling)
__device__ __forceinline__ void stage_block_nodes(const BlockOp& bo, const BlockSmem& s, const BlockCtx& c, const double* Xnode, double sc) {
  for (uint32_t i = threadIdx.x; i < c.nt; i += kBlock) sm_st(s.node, i, __dmul_rn(__ldcg(Xnode + __ldg(bo.tbn + c.t0 + i)), sc));
  for (uint32_t i = threadIdx.x; i < c.nh; i += kBlock) sm_st(s.node, bo.PT + i, __dmul_rn(__ldcg(Xnode + __ldg(bo.hbn + c.h0 + i)), sc));
}
__device__ __forceinline__ void zero_block_acc(const BlockOp& bo, const BlockSmem& s, int tid = threadIdx.x, int nthreads = kBlock) {
  for (uint32_t i = tid; i < bo.PT + bo.PH + kBAccPad; i += nthreads) sm_st(s.acc, i, 0.0);
}

// This CTA's partial sums (s.acc) -> the owners' buffers, parity `par`.  Destination-indexed: node u of rank rk keeps
// world * (GC + GR) consecutive words; rank s, cell (r, cc) writes the tail side of its nodes to word s * (GC + GR) + cc and
// the head side to word s * (GC + GR) + GC + r.  Every word of an active node is written exactly once per step (a cell writes
// all nodes of its two blocks, zeros included); the words of a node that is not active on a side are never written and keep
// the zero they were allocated with.  Nothing has to be cleared.  Caller synchronised before.
__device__ __forceinline__ void publish_block_partials(const BlockOp& bo, const BlockSmem& s, const BlockCtx& c, uint32_t par) {
  const Fabric& f = bo.tl.fab;
  const uint32_t per = bo.GC + bo.GR, SL = f.world * per;
  for (uint32_t i = threadIdx.x; i < c.nt; i += kBlock) {
    const uint32_t u = __ldg(bo.tbn + c.t0 + i), rk = u / f.Bp;
    __stcg(f.partials[rk] + ((size_t)par * f.Bp + (u - rk * f.Bp)) * SL + f.rank * per + c.cc, sm_ld(s.acc, i));
  }
  for (uint32_t i = threadIdx.x; i < c.nh; i += kBlock) {
    const uint32_t u = __ldg(bo.hbn + c.h0 + i), rk = u / f.Bp;
    __stcg(f.partials[rk] + ((size_t)par * f.Bp + (u - rk * f.Bp)) * SL + f.rank * per + bo.GC + c.r, sm_ld(s.acc, bo.PT + i));
  }
}
// T_u = sum of the world * (GC + GR) partials of an owned node: lanes stride the words, xor tree -- a fixed order.
// block_node_lanes is the first half (the lane's share, loads only), warp_sum of it the total: the callers request the
// shares of several nodes (and the node's other operands) before they wait for any of them.
__device__ __forceinline__ double block_node_lanes(const BlockOp& bo, uint32_t par, uint32_t u, int lane) {
  const Fabric& f = bo.tl.fab;
  const uint32_t SL = f.world * (bo.GC + bo.GR);
  const double* Pin = f.partials[f.rank] + ((size_t)par * f.Bp + (u - f.rank * f.Bp)) * SL;
  double a = 0.0;
  for (uint32_t q = lane; q < SL; q += 32) a = __dadd_rn(a, __ldcg(Pin + q));
  return a;
}
constexpr int kBOwnRounds = 2;  // owned node rows a warp has in flight

// Same-tail runs of a stage, summed by the compute warp that has just produced the values (they are still in its registers):
// lane l holds arcs l, l + 32, l + 64, l + 96 of the stage.  A run [start, start + len) adds its arcs per lane in that order,
// then the xor tree; lane 0 stores the sum into the tile buffer behind the T arc values, where the list walk finds it like
// any other value.  Both passes run this code: the tail-side sums are bit-identical.  (Round-2 measurement: summed by the
// fold warps from shared memory, these runs were 45 % of the fold's time, and the fold bounds phase B of pass 1.)
__device__ __forceinline__ void stage_pieces(const double (&w)[4], uint32_t slot_desc, uint32_t wt_tile, uint32_t T, int lane) {
  uint4 pd;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pd.x), "=r"(pd.y), "=r"(pd.z), "=r"(pd.w) : "r"(slot_desc));
  const uint32_t d[4] = {pd.x, pd.y, pd.z, pd.w};
  // (running the xor trees of the runs side by side instead of one after the other was measured 14 % SLOWER: 299 vs 263 us
  // per pass-1 step at 20M arcs)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (d[i] == kBNoPiece) continue;  // the same for every lane
    const uint32_t start = d[i] & 0xffu, len = ((d[i] >> 8) & 0xffu) + 1u;
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) a = __dadd_rn(a, (uint32_t)(lane + 32 * q) - start < len ? w[q] : 0.0);
    a = warp_sum(a);
    if (lane == 0) asm volatile("st.shared.f64 [%0], %1;" ::"r"(wt_tile + (T + (d[i] >> 16)) * 8u), "d"(a));
  }
}

// ---------------------------------------------------------------------------- node sums of a tile (fold warps)
struct BlockTileHdr {
  uint32_t e0, L, D;  // first word of the tile's block, entries per thread, deepest chain
  uint32_t q0, q1;    // the tile's pieces
};
__device__ __forceinline__ BlockTileHdr block_tile_hdr(const BlockSmem& s, uint32_t t) {  // tile t of this CTA's cell
  uint32_t e0, L, q0, q1;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e0), "=r"(L), "=r"(q0), "=r"(q1) : "r"(s.hdr + t * 16u));
  return BlockTileHdr{e0, L & 0xffffffu, L >> 24, q0, q1};
}
// Adds the node sums of the tile held in buffer `wt` (arc values + run sums) into s.acc; `lst` is the shared-memory copy of
// the tile's list block (bulk-copied ahead: a global load per batch of entries made the fold latency-bound -- 2.2 cycles
// per arc with nothing else running, against 1.2 available in phase B of pass 1).  NFT fold threads walk the kBSlices slices
// of the tile, thread f the slices f, f + NFT, ...
// Walk of a slice: straight-line code per entry (value load, sign, add; a node change is a predicated read-modify-write of
// the previous node's accumulator).  The share of the slice's first node goes to the slice's scratch word instead when
// earlier slices hold entries of the same node (depth > 0) and is added chain position by chain position afterwards.
template <int NFT>
__device__ __forceinline__ void block_fold_tile(const BlockSmem& s, uint32_t wt, uint32_t lst, const BlockTileHdr& h, uint32_t dummy,
                                                uint32_t pad, int ftid) {
  constexpr int kPer = kBSlices / NFT;  // slices per thread, walked side by side (independent chains: twice the work in flight)
  uint32_t depth[kPer], fin[kPer], row[kPer];
  double sum[kPer];
  auto slot_addr = [&](uint32_t word, int shift) __attribute__((always_inline)) {  // 14-bit slot field at bit `shift` (>= 3) -> address
    return s.acc.a + ((word >> (shift - 3)) & (kBEntNodeMask << 3));
  };
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const uint32_t slice = (uint32_t)ftid + j * NFT;
    const uint32_t mine = lst + slice * 4u;  // row q of the slice at mine + (q + 1) * 4 * kBSlices
    const uint32_t r0 = lds32<0>(mine);
    depth[j] = r0 & 0xffu;
    fin[j] = slot_addr(r0, 8);
    if (depth[j]) asm volatile("st.shared.f64 [%0], %1;" ::"r"(s.scr + slice * 8u), "d"(0.0));
    sum[j] = 0.0;
    row[j] = mine + 4u * kBSlices;
  }
  // NB entries of every slice of this thread: the entry words, then the tile values they name (independent loads), then the
  // adds.  An entry that opens a new node first adds the finished sum to the slot it names (a slot is named by one entry of
  // one slice per tile: no two threads update one slot in the same phase).  A branch-free form -- every entry one
  // read-modify-write, the dummy slot when no node ends, the writes of four entries issued together -- walks faster alone
  // but loads the shared-memory pipe the compute warps also need: pass 2 at 20M arcs 210 us per step against 203.5.
  // FIRST: the slice's first entry names its node, not a slot (there is nothing to flush yet).  GUARD: only the first `rem`
  // rows exist, the others are taken as padding entries (zero into the running sum).
  auto batch = [&](auto nb_tag, auto guard_tag, auto first_tag, uint32_t rem) __attribute__((always_inline)) {
    constexpr int NB = decltype(nb_tag)::value;
    constexpr bool GUARD = decltype(guard_tag)::value, FIRST = decltype(first_tag)::value;
    uint32_t ent[kPer][NB];
    double x[kPer][NB];
#pragma unroll
    for (int j = 0; j < kPer; ++j)
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        ent[j][q] = pad;
        if (!GUARD || (uint32_t)q < rem) ent[j][q] = lds32_at(row[j] + q * (4u * kBSlices));
      }
#pragma unroll
    for (int j = 0; j < kPer; ++j)
#pragma unroll
      for (int q = 0; q < NB; ++q) x[j][q] = lds64_at(wt + (ent[j][q] & kBEntOffMask));  // the batch's tile values: independent loads
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const uint32_t e = ent[j][q];
        const double xv = x[j][q];
        const double val = __hiloint2double(__double2hiint(xv) ^ (int)(e & kBEntMinus), __double2loint(xv));
        if (e & kBEntNew) {  // the node before is complete: one read-modify-write of the slot the entry names
          if (!(FIRST && q == 0)) {
            const uint32_t a = slot_addr(e, kBEntNodeShift);
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(__dadd_rn(lds64_at(a), sum[j])));
          }
          sum[j] = val;
        } else {
          sum[j] = __dadd_rn(sum[j], val);
        }
      }
#pragma unroll
    for (int j = 0; j < kPer; ++j) row[j] += NB * 4u * kBSlices;
  };
  using std::integral_constant;
  using T8 = integral_constant<int, kBPre>;
  using T4 = integral_constant<int, kBFlush>;
  using Yes = integral_constant<bool, true>;
  using No = integral_constant<bool, false>;
  // L is the same for every slice; the first batch is peeled for the first-entry rule
  auto tail = [&](auto first_tag, uint32_t rem) __attribute__((always_inline)) {
    if (rem > (uint32_t)kBFlush)
      batch(T8{}, Yes{}, first_tag, rem);
    else if (rem > 0u)
      batch(T4{}, Yes{}, first_tag, rem);
  };
  if (h.L >= (uint32_t)kBPre) {
    batch(T8{}, No{}, Yes{}, 0u);
    uint32_t q0 = kBPre;
    for (; q0 + kBPre <= h.L; q0 += kBPre) batch(T8{}, No{}, No{}, 0u);
    tail(No{}, h.L - q0);
  } else {
    tail(Yes{}, h.L);
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j) asm volatile("st.shared.f64 [%0], %1;" ::"r"(fin[j]), "d"(__dadd_rn(lds64_at(fin[j]), sum[j])));
  if (h.D) {
    for (uint32_t d = 1; d <= h.D; ++d) {  // shares of straddling nodes, chain position by chain position
      bar_sync_n(kBarFold, NFT);
#pragma unroll
      for (int j = 0; j < kPer; ++j)
        if (depth[j] == d) {
          const uint32_t slice = (uint32_t)ftid + j * NFT;
          const uint32_t e1 = lds32_at(lst + slice * 4u + 4u * kBSlices);  // the slice's first entry names the node
          const uint32_t dn = s.acc.a + ((e1 >> kBEntNodeShift) & kBEntNodeMask) * 8u;
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(dn), "d"(__dadd_rn(lds64_at(dn), lds64_at(s.scr + slice * 8u))));
        }
    }
  }
}

// ---------------------------------------------------------------------------- the folding sweep
// One sweep over the CTA's cell that produces a new arc vector AND its node sums.  Compute warp w owns the stages
// g = w, w + 8, w + 16, ... of the cell (a stage = 128 consecutive cell-order arcs) and a private ring of RING slots; lane 0
// issues the bulk copies of N8 8-byte arrays, N4 4-byte arrays and the stage's run descriptors.  `consume(pos, slot, wt_stage,
// wt_tile, desc, lane)` handles the arcs pos + lane + 32 q (q < 4): it reads array a of the slot at slot + a * 1024 (+ 8 *
// index) (4-byte arrays behind the 8-byte ones), stores its results to global memory, the new arc value of arc q to
// wt_stage + 8 * (lane + 32 q), and the sums of the stage's same-tail runs behind the tile's arc values (stage_pieces).
// A tile (T arcs) is complete when all compute warps have arrived; the fold warps then add its node sums into s.acc while
// the compute warps fill the other tile buffer (named barriers as in tile_loop).  `rs` (ring slot and mbarrier phase of this
// warp's next stage) lives across the sweeps of a kernel: the mbarriers are initialised once.
struct RingState {
  uint32_t slot, phase;  // compute warps: ring slot and mbarrier phase of the next stage
  uint32_t lphase;       // fold warps: bit b = phase of list buffer b's mbarrier
};
// `pre(ftid, NFT)` runs on the fold threads before their first tile (they would otherwise wait for it): zeroing the
// accumulators and whatever else the sweep's input does not depend on; a barrier among the fold threads follows.
template <int NCW, int N8, int N4, class CONSUME, class PRE>
__device__ __forceinline__ void fold_sweep(const BlockOp& bo, const BlockSmem& s, const BlockCtx& c, uint32_t RING, uint32_t slot_bytes,
                                           const double* const (&src8)[N8], const uint32_t* const (&src4)[N4 ? N4 : 1],
                                           CONSUME consume, PRE pre, RingState& rs, const Trace* tr = nullptr, int tr_step = -1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t SPT = bo.tl.T / kBStage;  // stages per tile (a multiple of 8)
  const bool timed = tr != nullptr && tr->buf != nullptr && tr_step >= 0 && tr_step < tr->max_steps;
  long long c_a = 0, c_b = 0, c_l = 0, c_f = 0, c_tot = 0, t_x = 0;  // diagnostics: cycles waiting for data / for the other role, total
  if (timed) c_tot = -clock64();
  constexpr int NFT = kBlock - 32 * NCW;  // fold threads
  static_assert(kBSlices % NFT == 0, "every fold thread walks the same number of slices");
  if (warp < NCW) {
    const uint32_t nk = c.nst > (uint32_t)warp ? (c.nst - warp + NCW - 1) / NCW : 0;  // my stages: g = warp, warp + NCW, ...
    const uint32_t ring0 = s.ring + (uint32_t)warp * RING * slot_bytes, bar0 = s.mbar + (uint32_t)warp * kBMaxRing * 8u;
    constexpr uint32_t kTx = N8 * kBStage * 8 + N4 * kBStage * 4 + 16;
    constexpr uint32_t kDescOff = N8 * kBStage * 8 + N4 * kBStage * 4;  // the stage's run descriptors sit behind the arrays
    auto issue = [&](uint32_t k, uint32_t sl) __attribute__((always_inline)) {  // lane 0 only: stage k of this warp into slot sl
      const uint32_t bar = bar0 + sl * 8u, dst = ring0 + sl * slot_bytes;
      const size_t pos = (size_t)c.c0 + ((size_t)k * NCW + warp) * kBStage;
      mbar_expect_tx(bar, kTx);
#pragma unroll
      for (int a = 0; a < N8; ++a) bulk_g2s(dst + a * (kBStage * 8), src8[a] + pos, kBStage * 8, bar);
#pragma unroll
      for (int a = 0; a < N4; ++a) bulk_g2s(dst + N8 * (kBStage * 8) + a * (kBStage * 4), src4[a] + pos, kBStage * 4, bar);
      bulk_g2s(dst + kDescOff, bo.pdesc + pos / kBStage, 16, bar);
    };
    // (issuing the copies of a stage from several lanes in one instruction instead of one after the other by lane 0 changes
    // nothing: measured, 2 M - 50 M arcs)
    if (lane == 0) {
      uint32_t sl = rs.slot;
      for (uint32_t k = 0; k < RING && k < nk; ++k) {
        issue(k, sl);
        sl = sl + 1 == RING ? 0 : sl + 1;
      }
    }
    uint32_t k = 0, tb_i = 0;  // tb_i = t % ntb: the tile buffer of tile t
    const uint32_t ntb = bo.ntb;
    for (uint32_t t = 0; t < c.ntiles; ++t) {
      if (timed) t_x = clock64();
      if (t >= ntb) bar_sync_n(kBBarEmpty + tb_i, kBlock);  // the fold of tile t - ntb has left this buffer
      if (timed) c_b += clock64() - t_x;
      const uint32_t wt0 = s.wt.a + tb_i * s.wt_stride;
      // my stages inside tile t: g = warp (mod NCW) -- the round robin runs across tile boundaries, every warp gets the same
      // number of stages (+- 1) whatever the tile size
      const uint32_t tb = t * SPT, te = min(tb + SPT, c.nst);
      uint32_t g = tb + ((uint32_t)warp + NCW - tb % NCW) % NCW;
      for (; g < te; g += NCW) {
        if (timed) t_x = clock64();
        mbar_wait(bar0 + rs.slot * 8u, rs.phase);
        if (timed) c_a += clock64() - t_x;
        if (timed) t_x = clock64();
        if (!(bo.dbg & 2u))
          consume(c.c0 + g * kBStage, ring0 + rs.slot * slot_bytes, wt0 + (g - t * SPT) * (kBStage * 8u), wt0, ring0 + rs.slot * slot_bytes + kDescOff, lane);
        __syncwarp();  // every lane has read its slot words
        if (timed) c_f += clock64() - t_x;
        if (lane == 0 && k + RING < nk) issue(k + RING, rs.slot);
        ++k;
        if (++rs.slot == RING) {
          rs.slot = 0;
          rs.phase ^= 1u;
        }
      }
      bar_arrive_n(kBBarFull + tb_i, kBlock);
      tb_i = tb_i + 1 == ntb ? 0 : tb_i + 1;
    }
    const long long t_loop_end = timed ? clock64() : 0;
    // drain: every arrival of the fold warps is matched by a wait, so that the barriers are clean for the next sweep
    for (uint32_t u = c.ntiles > ntb ? c.ntiles - ntb : 0; u < c.ntiles; ++u) bar_sync_n(kBBarEmpty + u % ntb, kBlock);
    fence_proxy_async();  // the vector just written is bulk-copied by the next sweep (after the grid barrier in between)
    if (timed && threadIdx.x == 0) {
      unsigned long long* q = tr->buf + ((size_t)blockIdx.x * tr->max_steps + tr_step) * kTraceMarks;
      q[16] = (unsigned long long)c_a;                 // compute warp 0: waiting for bulk copies
      q[17] = (unsigned long long)c_b;                 // ... for the fold to release a tile buffer
      q[18] = (unsigned long long)(c_tot + clock64()); // ... whole sweep
      q[23] = (unsigned long long)c_f;                 // ... consuming stages
      q[24] = (unsigned long long)(c_tot + t_loop_end); // ... last stage consumed (since the start of the sweep)
    }
  } else {
    // Fold warps.  The list block of tile t is bulk-copied into list buffer t % nl by fold thread 0: tiles 0 .. nl - 1 at the
    // start of the sweep, tile t - 1 + nl as soon as every fold thread has left tile t - 1 (= has passed the full-barrier of
    // tile t).  Short sweeps (small cells) thus have their whole list in flight from the start.
    const uint32_t dummy = bo.PT + bo.PH, nl = bo.nl;
    const uint32_t lbar = s.mbar + kWarps * kBMaxRing * 8u;
    const int ftid = (int)threadIdx.x - 32 * NCW;
    const bool issuer = ftid == 0;
    auto fetch = [&](const BlockTileHdr& h, uint32_t b) __attribute__((always_inline)) {  // issuer only
      const uint32_t bytes = (h.L + 1u) * (4u * kBSlices);
      mbar_expect_tx(lbar + b * 8u, bytes);
      bulk_g2s(s.lst + b * bo.lblk, bo.tl.lent + h.e0, bytes, lbar + b * 8u);
    };
    if (c.ntiles && issuer)
      for (uint32_t u = 0; u < nl && u < c.ntiles; ++u) fetch(block_tile_hdr(s, u), u);
    pre(ftid, NFT);
    bar_sync_n(kBarFold, NFT);
    const long long t_pre_end = timed ? clock64() : 0;
    long long t_first_full = 0;
    uint32_t b = 0, tb_i = 0;  // list buffer / tile buffer of tile t
    const uint32_t ntb = bo.ntb;
    for (uint32_t t = 0; t < c.ntiles; ++t) {
      const BlockTileHdr h0 = block_tile_hdr(s, t);
      if (timed) t_x = clock64();
      bar_sync_n(kBBarFull + tb_i, kBlock);
      if (timed) c_a += clock64() - t_x;
      if (timed && t == 0) t_first_full = clock64();
      if (issuer && t >= 1 && t - 1 + nl < c.ntiles)  // every fold thread has left tile t - 1: its buffer is free
        fetch(block_tile_hdr(s, t - 1 + nl), b == 0 ? nl - 1 : b - 1);
      if (timed) t_x = clock64();
      mbar_wait(lbar + b * 8u, (rs.lphase >> b) & 1u);
      if (timed) c_l += clock64() - t_x;
      rs.lphase ^= 1u << b;
      if (timed) t_x = clock64();
      if (!(bo.dbg & 1u))
        block_fold_tile<NFT>(s, s.wt.a + tb_i * s.wt_stride, s.lst + b * bo.lblk, h0, dummy, block_pad_entry(dummy, bo.tl.T), ftid);
      if (timed) c_f += clock64() - t_x;
      bar_arrive_n(kBBarEmpty + tb_i, kBlock);
      tb_i = tb_i + 1 == ntb ? 0 : tb_i + 1;
      b = b + 1 == nl ? 0 : b + 1;
    }
    if (timed && ftid == 0) {
      unsigned long long* q = tr->buf + ((size_t)blockIdx.x * tr->max_steps + tr_step) * kTraceMarks;
      q[19] = (unsigned long long)c_a;                 // fold warp 0: waiting for a full tile
      q[20] = (unsigned long long)(c_tot + clock64()); // ... whole sweep
      q[21] = (unsigned long long)c_l;                 // ... waiting for list blocks
      q[22] = (unsigned long long)c_f;                 // ... folding
      q[25] = (unsigned long long)(c_tot + t_pre_end);    // ... preamble done (since the start of the sweep)
      q[26] = (unsigned long long)(c_tot + t_first_full); // ... first tile full
    }
  }
  __syncthreads();
}

// mbarriers of the rings, and the zero slot of both tile buffers that padding list entries read
__device__ __forceinline__ void init_block_smem(const BlockOp& bo, const BlockSmem& s) {
  if (threadIdx.x < kWarps * kBMaxRing + kBMaxList) mbar_init(s.mbar + threadIdx.x * 8u, 1);
  for (uint32_t t = threadIdx.x; t < bo.tl.ntile; t += kBlock) {
    const TileHdr h = tile_hdr(bo.tl, blockIdx.x * bo.tl.ntile + t);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s.hdr + t * 16u), "r"(h.e0), "r"(h.L), "r"(h.q0), "r"(h.q1) : "memory");
  }
  if (threadIdx.x < bo.ntb) sm_st(SmArr{s.wt.a + threadIdx.x * s.wt_stride}, bo.tl.T + block_piece_slots(bo.tl.T) - 1, 0.0);
  fence_mbar_init();
  __syncthreads();
}

// node partial sums of an arbitrary cell-order arc vector X over the CTA's cell (init: the un-normalised b)
template <int NCW>
__device__ __forceinline__ void block_sums_of(const BlockOp& bo, const BlockSmem& s, const BlockCtx& c, uint32_t RING, uint32_t slot_bytes,
                                              const double* X, RingState& rs) {
  const double* const src8[1] = {X};
  const uint32_t* const src4[1] = {nullptr};
  fold_sweep<NCW, 1, 0>(
      bo, s, c, RING, slot_bytes, src8, src4,
      [&](uint32_t, uint32_t slot, uint32_t wt, uint32_t wt_tile, uint32_t desc, int lane) __attribute__((always_inline)) {
        const uint32_t sl = slot + lane * 8u, w = wt + lane * 8u;
        const double x[4] = {lds64<0>(sl), lds64<256>(sl), lds64<512>(sl), lds64<768>(sl)};
        sts64<0>(w, x[0]);
        sts64<256>(w, x[1]);
        sts64<512>(w, x[2]);
        sts64<768>(w, x[3]);
        stage_pieces(x, desc, wt_tile, bo.tl.T, lane);
      },
      [&](int ftid, int nft) __attribute__((always_inline)) { zero_block_acc(bo, s, ftid, nft); }, rs);
}

// =============================================================================================
// pass 1 / one-pass basis generation
// =============================================================================================
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass1_blocked_kernel(const IncidenceOp op, const BlockOp bo, const Pass1Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  const uint32_t PL = bo.PT + bo.PH, p = op.p, m = bo.m, M = bo.Mpad;
  const BlockSmem s = carve_blocks(smem, PL, bo.tl.T, bo.lblk, bo.nl, bo.ntb, bo.tl.ntile, false);
  const BlockCtx c = block_ctx(bo, p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t RING = bo.ring1;
  const uint32_t slot_bytes = (uint32_t)block_slot_bytes(2, 0);
  const TileOp& to = bo.tl;
  init_block_smem(bo, s);

  unsigned int epoch = a.st->epoch;
  RingState rs{0u, 0u, 0u};
  int steps = 0, status = ST_RUNNING, rot = 0;
  double sc = 1.0, sp = 1.0, bp = 0.0, bnorm = 0.0;
  GridSync gs = a.gs;
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  {
    // K0: b gathered into cell order, ||b||, W_cur = b, W_prev = 0, partial node sums of b for step 0
    double* Wp = pick(0);
    double* Wc = pick(1);
    double acc = 0.0;
    for (uint32_t i = c.c0 + threadIdx.x; i < c.c1; i += kBlock) {
      const uint32_t g = __ldg(bo.gidx + i);
      const double bi = g != kBPad ? __ldg(a.b + g) : 0.0;
      __stcg(Wc + i, bi);
      __stcg(Wp + i, 0.0);
      acc = fma(bi, bi, acc);
    }
    for (uint32_t u = c.ulo + threadIdx.x; u < c.uhi; u += kBlock) {
      const double bi = __ldg(a.b + m + u);
      __stcg(Wc + M + u, bi);
      __stcg(Wp + M + u, 0.0);
      publish_node(to, p, 1, u, bi);  // parity of "step -1"
      acc = fma(bi, bi, acc);
    }
    fence_proxy_async();
    __syncthreads();  // the cell's share of W_cur is written (and fenced towards the async proxy) before it is bulk-copied
    block_sums_of<kBComputeWarps1>(bo, s, c, RING, slot_bytes, Wc, rs);
    publish_block_partials(bo, s, c, 0);
    bnorm = sqrt(tile_sync<true>(acc, to, a.gs, epoch, sh));
    if (bnorm <= a.tol) status = ST_ZERO_B;
    sc = 1.0 / bnorm;
  }
  if (status == ST_RUNNING) {
    for (int j = 0; j < a.j_end; ++j) {
      const double* Wp = pick(rot);
      const double* Wc = pick((rot + 1) % 3);
      double* Wn = pick((rot + 2) % 3);
      const double* Xnode = local_nodebuf(to, p, (j + 1) & 1);  // node part of the current vector (un-normalised)
      double* Vcol = WITH_V ? a.V + (size_t)j * a.ldv : nullptr;

      // ---------------- phase A: w~ = A v - beta_{j-1} v_{j-1}, alpha partial
      gs.trace_step = j;
      trace_mark(gs.trace, j, 0);
      // node rows of the owned block: the operands of the warp's first rows are requested BEFORE the node values are staged
      // (two independent chains of L2 round trips at the top of every step; one behind the other they cost 5.7 k cycles)
      double own_t[kBOwnRounds], own_x[kBOwnRounds], own_p[kBOwnRounds];
#pragma unroll
      for (int r = 0; r < kBOwnRounds; ++r) {
        const uint32_t u = c.ulo + warp + r * kWarps;
        own_t[r] = own_x[r] = own_p[r] = 0.0;
        if (u < c.uhi) {
          own_t[r] = block_node_lanes(bo, j & 1, u, lane);
          if (lane == 0) {
            own_x[r] = __ldcg(Xnode + u);
            own_p[r] = __ldcg(Wp + M + u);
          }
        }
      }
      stage_block_nodes(bo, s, c, Xnode, sc);
      __syncthreads();
      trace_mark(gs.trace, j, 1);
      if (WITH_V) {  // node part of the basis column: replicated on every rank, each CTA writes its share
        uint32_t vlo, vhi;
        cta_chunk(p, vlo, vhi);
        for (uint32_t u = vlo + threadIdx.x; u < vhi; u += kBlock) __stcs(Vcol + m + u, __dmul_rn(__ldcg(Xnode + u), sc));
      }
      double acc = 0.0;
      auto own_row = [&](uint32_t u, double lanes, double xn, double wpu) __attribute__((always_inline)) {
        const double t = __dmul_rn(sc, warp_sum(lanes));
        if (lane == 0) {
          const double v = __dmul_rn(xn, sc);
          const double vp = __dmul_rn(wpu, sp);
          const double wt = rec_sub(t, bp, vp);
          acc = fma(v, wt, acc);
          __stcg(Wn + M + u, wt);
        }
      };
#pragma unroll
      for (int r = 0; r < kBOwnRounds; ++r) {
        const uint32_t u = c.ulo + warp + r * kWarps;
        if (u < c.uhi) own_row(u, own_t[r], own_x[r], own_p[r]);
      }
      for (uint32_t u = c.ulo + warp + kBOwnRounds * kWarps; u < c.uhi; u += kWarps) {  // (more than 32 owned rows per CTA)
        const double lanes = block_node_lanes(bo, j & 1, u, lane);
        own_row(u, lanes, lane == 0 ? __ldcg(Xnode + u) : 0.0, lane == 0 ? __ldcg(Wp + M + u) : 0.0);
      }
      trace_mark(gs.trace, j, 2);
      for (uint32_t base = c.c0; base < c.c1; base += kUnroll * kBlock) {
        double wc[kUnroll], wp[kUnroll], dd[kUnroll];
        uint32_t th[kUnroll], gi[kUnroll];
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          const uint32_t i = min(base + q * kBlock + threadIdx.x, c.c1 - 1);  // unconditional loads (see tpl_tiles.cuh)
          wc[q] = __ldcg(Wc + i);
          wp[q] = __ldcg(Wp + i);
          dd[q] = __ldg(bo.d + i);
          th[q] = __ldg(bo.th + i);
          if (WITH_V) gi[q] = __ldg(bo.gidx + i);
        }
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          const uint32_t i = base + q * kBlock + threadIdx.x;
          if (i < c.c1) {
            const double v = __dmul_rn(wc[q], sc);
            const double vp = __dmul_rn(wp[q], sp);
            const double xt = sm_ld(s.node, th[q] & 0x7fffu), xh = sm_ld(s.node, bo.PT + ((th[q] >> 15) & 0x7fffu));
            const double wt = rec_sub(arc_row_b(dd[q], v, th[q], xt, xh), bp, vp);
            acc = fma(v, wt, acc);
            __stcg(Wn + i, wt);
            if (WITH_V)
              if (gi[q] != kBPad) __stcs(Vcol + gi[q], v);
          }
        }
      }
      fence_proxy_async();  // w~ is bulk-copied by phase B
      trace_mark(gs.trace, j, 3);
      gs.trace_base = 4;
      // all-reduce only: what phase B reads of phase A's output (w~ of the own cell and of the owned node rows) was written by
      // this CTA, ordered by the CTA barriers inside the call (and the proxy fence above for the bulk copies) -- no release /
      // acquire fence at GPU scope, which costs 1.6 us each with a sweep's stores in flight
      const double alpha = to.fab.world > 1 ? fabric_sync<true, false>(acc, to, epoch, sh) : grid_sync<true, false>(acc, gs, epoch, sh);

      // ---------------- phase B: w = w~ - alpha v, beta partial, partial node sums of w
      acc = 0.0;
      // the fold threads zero the accumulators (they alias s.node: phase A is over) and finish the owned node rows while the
      // compute warps already stream the first tile
      auto phase_b_pre = [&](int ftid, int nft) __attribute__((always_inline)) {
        zero_block_acc(bo, s, ftid, nft);
        for (uint32_t u = c.ulo + ftid; u < c.uhi; u += nft) {
          const double v = __dmul_rn(__ldcg(Wc + M + u), sc);
          const double w = rec_sub(__ldcg(Wn + M + u), alpha, v);
          __stcg(Wn + M + u, w);
          publish_node(to, p, j & 1, u, w);
          acc = fma(w, w, acc);
        }
      };
      trace_mark(gs.trace, j, 8);
      {
        const double* const src8[2] = {Wn, Wc};
        const uint32_t* const src4[1] = {nullptr};
        fold_sweep<kBComputeWarps1, 2, 0>(
            bo, s, c, RING, slot_bytes, src8, src4,
            [&](uint32_t pos, uint32_t slot, uint32_t wt, uint32_t wt_tile, uint32_t desc, int ln) __attribute__((always_inline)) {
              const uint32_t sl = slot + ln * 8u;
              double wn[4], wc[4];
              wn[0] = lds64<0>(sl), wn[1] = lds64<256>(sl), wn[2] = lds64<512>(sl), wn[3] = lds64<768>(sl);
              wc[0] = lds64<1024>(sl), wc[1] = lds64<1280>(sl), wc[2] = lds64<1536>(sl), wc[3] = lds64<1792>(sl);
              double w[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) w[q] = rec_sub(wn[q], alpha, __dmul_rn(wc[q], sc));
              double* out = Wn + pos + ln;
              __stcg(out, w[0]);
              __stcg(out + 32, w[1]);
              __stcg(out + 64, w[2]);
              __stcg(out + 96, w[3]);
              const uint32_t ws = wt + ln * 8u;
              sts64<0>(ws, w[0]);
              sts64<256>(ws, w[1]);
              sts64<512>(ws, w[2]);
              sts64<768>(ws, w[3]);
              stage_pieces(w, desc, wt_tile, bo.tl.T, ln);
#pragma unroll
              for (int q = 0; q < 4; ++q) acc = fma(w[q], w[q], acc);
            },
            phase_b_pre, rs, &gs.trace, j);
      }
      trace_mark(gs.trace, j, 9);
      publish_block_partials(bo, s, c, (j + 1) & 1);
      trace_mark(gs.trace, j, 10);
      gs.trace_base = 11;
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
// pass 2: one sweep and one grid barrier per step
// =============================================================================================
template <bool WITH_V>
__global__ void __launch_bounds__(kBlock, 1) pass2_blocked_kernel(const IncidenceOp op, const BlockOp bo, const Pass2Args a) {
  extern __shared__ double smem[];
  __shared__ CtaShared sh;
  const uint32_t PL = bo.PT + bo.PH, p = op.p, m = bo.m, M = bo.Mpad;
  const BlockSmem s = carve_blocks(smem, PL, bo.tl.T, bo.lblk, bo.nl, bo.ntb, bo.tl.ntile, true);
  const BlockCtx c = block_ctx(bo, p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t RING = WITH_V ? bo.ring2v : bo.ring2;
  const uint32_t slot_bytes = (uint32_t)block_slot_bytes(4, WITH_V ? 2 : 1);
  const TileOp& to = bo.tl;
  init_block_smem(bo, s);
  uint32_t vlo, vhi;  // share of the (replicated) node part of x / V this CTA writes
  cta_chunk(p, vlo, vhi);
  unsigned int epoch = a.st->epoch;
  RingState rs{0u, 0u, 0u};
  double* const buf0 = a.buf[0];
  double* const buf1 = a.buf[1];
  double* const buf2 = a.buf[2];
  double* const xc = bo.xc;
  auto pick = [&](int r) { return r == 0 ? buf0 : (r == 1 ? buf1 : buf2); };
  int rot = 0;
  double sc_cur = 1.0 / a.b_norm;  // scale of the vector whose un-normalised partial node sums are in `partials`
  {
    // v_1 = b * (1/||b||) gathered into cell order, x = y_0 v_1; partial node sums of the un-normalised b (scaled afterwards,
    // exactly as pass 1 does with its lazily scaled w: bit-identical node rows)
    const double inv = 1.0 / a.b_norm;
    const double y0 = __ldg(a.y);
    double* Vp = buf0;
    double* Vc = buf1;
    double* Braw = buf2;
    for (uint32_t u = c.ulo + threadIdx.x; u < c.uhi; u += kBlock) {
      const double v = __dmul_rn(__ldg(a.b + m + u), inv);
      __stcg(Vc + M + u, v);
      __stcg(Vp + M + u, 0.0);
      publish_node(to, p, 1, u, v);
    }
    for (uint32_t u = vlo + threadIdx.x; u < vhi; u += kBlock) {
      const double v = __dmul_rn(__ldg(a.b + m + u), inv);
      __stcg(a.x + m + u, __dmul_rn(v, y0));
      if (WITH_V) __stcs(a.V + m + u, v);
    }
    for (uint32_t i = c.c0 + threadIdx.x; i < c.c1; i += kBlock) {
      const uint32_t g = __ldg(bo.gidx + i);
      const double bi = g != kBPad ? __ldg(a.b + g) : 0.0;
      const double v = __dmul_rn(bi, inv);
      __stcg(Braw + i, bi);
      __stcg(Vc + i, v);
      __stcg(Vp + i, 0.0);
      __stcg(xc + i, __dmul_rn(v, y0));
      if (WITH_V)
        if (g != kBPad) __stcs(a.V + g, v);
    }
    fence_proxy_async();
    __syncthreads();
    block_sums_of<kBComputeWarps2>(bo, s, c, RING, slot_bytes, Braw, rs);
    publish_block_partials(bo, s, c, 0);
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
    GridSync gs = a.gs;
    gs.trace_step = j;
    trace_mark(gs.trace, j, 0);

    stage_block_nodes(bo, s, c, Xnode, 1.0);
    // What the sweep's input does not depend on runs on the fold threads while the compute warps already stream the first
    // tile: zeroing the accumulators, the node share of x (and of the basis column), the owned node rows.
    auto pass2_pre = [&](int ftid, int nft) __attribute__((always_inline)) {
      const int fw = ftid >> 5, nfw = nft >> 5;
      double own_t[kBOwnRounds], own_x[kBOwnRounds], own_p[kBOwnRounds];  // operands of the warp's first rows, requested together
#pragma unroll
      for (int r = 0; r < kBOwnRounds; ++r) {
        const uint32_t u = c.ulo + fw + r * nfw;
        own_t[r] = own_x[r] = own_p[r] = 0.0;
        if (u < c.uhi) {
          own_t[r] = block_node_lanes(bo, j & 1, u, lane);
          if (lane == 0) {
            own_x[r] = __ldcg(Xnode + u);
            own_p[r] = __ldcg(Vp + M + u);
          }
        }
      }
      zero_block_acc(bo, s, ftid, nft);
      if (j > 0) {
        // the published node values are v_{j+1}, regenerated by the previous step: their share of x (and of the basis column)
        // is added here, by every rank for its replica (each CTA its share of the nodes)
        const double yprev = __ldg(a.y + j);
        for (uint32_t u = vlo + ftid; u < vhi; u += nft) {
          const double vn = __ldcg(Xnode + u);
          __stcg(a.x + m + u, __dadd_rn(__ldcg(a.x + m + u), __dmul_rn(yprev, vn)));
          if (WITH_V) __stcs(a.V + (size_t)j * a.ldv + m + u, vn);
        }
      }
      auto own_row = [&](uint32_t u, double lanes, double xn, double vpu) __attribute__((always_inline)) {
        const double t = __dmul_rn(sc_cur, warp_sum(lanes));
        if (lane == 0) {
          const double w = rec_sub(rec_sub(t, bp, vpu), alpha, xn);
          const double vn = __dmul_rn(w, sinv);
          __stcg(Vn + M + u, vn);
          publish_node(to, p, j & 1, u, vn);
        }
      };
#pragma unroll
      for (int r = 0; r < kBOwnRounds; ++r) {
        const uint32_t u = c.ulo + fw + r * nfw;
        if (u < c.uhi) own_row(u, own_t[r], own_x[r], own_p[r]);
      }
      for (uint32_t u = c.ulo + fw + kBOwnRounds * nfw; u < c.uhi; u += nfw) {
        const double lanes = block_node_lanes(bo, j & 1, u, lane);
        own_row(u, lanes, lane == 0 ? __ldcg(Xnode + u) : 0.0, lane == 0 ? __ldcg(Vp + M + u) : 0.0);
      }
    };
    __syncthreads();  // node values are staged
    trace_mark(gs.trace, j, 1);
    {
      const double* const src8[4] = {Vc, Vp, xc, bo.d};
      const uint32_t* const src4[2] = {bo.th, bo.gidx};
      const uint32_t* const src4n[1] = {bo.th};
      auto body = [&](uint32_t pos, uint32_t slot, uint32_t wt, uint32_t wt_tile, uint32_t desc, int ln) __attribute__((always_inline)) {
        // all loads of the lane's four arcs first, then the node-value gathers, then the arithmetic, then the stores
        const uint32_t sl = slot + ln * 8u, sl4 = slot + 4 * kBStage * 8 + ln * 4u;
        double v[4], vp[4], xx[4], dd[4], xt[4], xh[4];
        uint32_t th[4];
        v[0] = lds64<0>(sl), v[1] = lds64<256>(sl), v[2] = lds64<512>(sl), v[3] = lds64<768>(sl);
        th[0] = lds32<0>(sl4), th[1] = lds32<128>(sl4), th[2] = lds32<256>(sl4), th[3] = lds32<384>(sl4);
        vp[0] = lds64<1024>(sl), vp[1] = lds64<1280>(sl), vp[2] = lds64<1536>(sl), vp[3] = lds64<1792>(sl);
        dd[0] = lds64<3072>(sl), dd[1] = lds64<3328>(sl), dd[2] = lds64<3584>(sl), dd[3] = lds64<3840>(sl);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          xt[q] = lds64_at(s.node.a + (th[q] & 0x7fffu) * 8u);
          xh[q] = lds64_at(s.node.a + (bo.PT + ((th[q] >> 15) & 0x7fffu)) * 8u);
        }
        xx[0] = lds64<2048>(sl), xx[1] = lds64<2304>(sl), xx[2] = lds64<2560>(sl), xx[3] = lds64<2816>(sl);
        double w[4], vn[4], xo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          w[q] = rec_sub(rec_sub(arc_row_b(dd[q], v[q], th[q], xt[q], xh[q]), bp, vp[q]), alpha, v[q]);
          vn[q] = __dmul_rn(w[q], sinv);
          xo[q] = __dadd_rn(xx[q], __dmul_rn(yj, vn[q]));
        }
        double* o1 = Vn + pos + ln;
        double* o2 = xc + pos + ln;
        __stcg(o1, vn[0]);
        __stcg(o1 + 32, vn[1]);
        __stcg(o1 + 64, vn[2]);
        __stcg(o1 + 96, vn[3]);
        __stcg(o2, xo[0]);
        __stcg(o2 + 32, xo[1]);
        __stcg(o2 + 64, xo[2]);
        __stcg(o2 + 96, xo[3]);
        const uint32_t ws = wt + ln * 8u;
        sts64<0>(ws, w[0]);
        sts64<256>(ws, w[1]);
        sts64<512>(ws, w[2]);
        sts64<768>(ws, w[3]);
        stage_pieces(w, desc, wt_tile, bo.tl.T, ln);
        if (WITH_V) {
          const uint32_t sg = sl4 + kBStage * 4;
          const uint32_t g0 = lds32<0>(sg), g1 = lds32<128>(sg), g2 = lds32<256>(sg), g3 = lds32<384>(sg);
          if (g0 != kBPad) __stcs(Vcol + g0, vn[0]);
          if (g1 != kBPad) __stcs(Vcol + g1, vn[1]);
          if (g2 != kBPad) __stcs(Vcol + g2, vn[2]);
          if (g3 != kBPad) __stcs(Vcol + g3, vn[3]);
        }
      };
      if (WITH_V)
        fold_sweep<kBComputeWarps2, 4, 2>(bo, s, c, RING, slot_bytes, src8, src4, body, pass2_pre, rs, &gs.trace, j);
      else
        fold_sweep<kBComputeWarps2, 4, 1>(bo, s, c, RING, slot_bytes, src8, src4n, body, pass2_pre, rs, &gs.trace, j);
    }
    trace_mark(gs.trace, j, 2);
    publish_block_partials(bo, s, c, (j + 1) & 1);
    trace_mark(gs.trace, j, 3);
    gs.trace_base = 4;
    tile_sync<false>(0.0, to, gs, epoch, sh);
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
  // x back to natural order: every CTA scatters the arcs of its own cell
  for (uint32_t i = c.c0 + threadIdx.x; i < c.c1; i += kBlock) {
    const uint32_t g = __ldg(bo.gidx + i);
    if (g != kBPad) __stcg(a.x + g, __ldcg(xc + i));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.st->epoch = epoch;
}

}  // namespace tpl
