// tpl_blocks_host.h -- host construction of the blocked streaming layout (tpl_blocks.cuh) and the checker the CPU tests run
// over it (tpl_blocks_plan).
//
//   1. node blocks: GR contiguous tail blocks with (nearly) equal out-degree, GC contiguous head blocks with (nearly) equal
//      in-degree; G = GR * GC cells;
//   2. cell order: arc j belongs to cell (tail block of tail_j) * GC + (head block of head_j); inside a cell the arcs are
//      ordered by tail node, then by arc index (two stable counting sorts), every cell is padded to a multiple of 128 slots;
//   3. per position: d, the packed th word (local tail | local head << 15 | tail-first | loop-or-padding) and gidx;
//   4. tile lists of every cell over LOCAL node ids (tails [0, PT), heads [PT, PT + PH)): build_cell_lists below.
// Cells are independent: they are built by a pool of host threads and concatenated in cell order (the result does not depend
// on the number of threads).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <numeric>
#include <thread>
#include <vector>

#include "tpl_blocks.cuh"
#include "tpl_tiles_host.h"

namespace tpl {

struct HostBlocks {
  bool ok = false;
  uint32_t GR = 0, GC = 0, PT = 0, PH = 0, Mpad = 0, T = 0, ntile = 0;
  uint32_t ring1 = 0, ring2 = 0, ring2v = 0, lblk = 0, nl = 2, ntb = 2;
  std::vector<uint32_t> cell_off, tbs, hbs, tbn, hbn, th, gidx;
  std::vector<double> d;
  std::vector<uint4> thdr, pdesc;  // pdesc: per 128-arc stage up to four same-tail runs (start | (len - 1) << 8 | tile slot << 16)
  std::vector<uint32_t> lent, piece;
};

// Contiguous blocks of [0, p) with (nearly) equal total weight: bnd[a] = first node of block a
inline std::vector<uint32_t> weight_blocks(const std::vector<uint64_t>& weight, uint32_t parts) {
  const size_t p = weight.size();
  const uint64_t total = std::accumulate(weight.begin(), weight.end(), (uint64_t)0);
  std::vector<uint32_t> bnd(parts + 1, (uint32_t)p);
  bnd[0] = 0;
  uint64_t pre = 0;
  uint32_t a = 1;
  for (size_t u = 0; u < p && a < parts; ++u) {
    while (a < parts && (total > 0 ? (2 * pre + weight[u]) * parts >= 2 * (uint64_t)a * total
                                   : (uint64_t)u * parts >= (uint64_t)a * p))
      bnd[a++] = (uint32_t)u;
    pre += weight[u];
  }
  for (; a < parts; ++a) bnd[a] = (uint32_t)p;
  for (uint32_t i = 1; i <= parts; ++i) bnd[i] = std::max(bnd[i], bnd[i - 1]);
  return bnd;
}

// Ring depths for which the kernels fit in `smem_limit` bytes with tiles of T arcs and list buffers of lblk bytes; `need` = the
// least number of ring slots pass 2 must get.
inline bool blocks_fit(uint32_t PL, size_t smem_limit, uint32_t T, uint32_t lblk, uint32_t ntb, uint32_t ntile, int need, uint32_t& ring1, uint32_t& ring2, uint32_t& ring2v) {
  const size_t budget = smem_limit > 3072 ? smem_limit - 3072 : 0;  // static shared memory of the kernels + margin
  auto fits = [&](int ring, bool pass2, bool v) { return block_smem_bytes(PL, T, ring, lblk, 2, ntb, pass2, v, ntile) <= budget; };
  if (!fits(need, true, false) || !fits(2, true, true) || !fits(2, false, false)) return false;
  ring2 = fits(4, true, false) ? 4 : fits(3, true, false) ? 3 : 2;
  ring2v = fits(3, true, true) ? 3 : 2;
  ring1 = fits(4, false, false) ? 4 : fits(3, false, false) ? 3 : 2;
  return true;
}

// Tile lists of ONE cell in the blocked format (entries over local node ids; see tpl_blocks.cuh "list format"):
// per tile the same-tail runs become pieces, every remaining tail entry, every piece and every head entry is one list entry;
// the entries are sorted by node (stable: a node's tail side in arc order, then its head side) and cut into kFoldThreads
// slices of EQUAL length L -- a node may straddle threads: thread i then adds its share of the node `depth_i` barrier phases
// after the thread that holds the node's first entry (depth = position in the chain of threads that share the node), which
// keeps the summation order fixed.  Row 0 of a tile's block holds the per-thread depth, rows 1..L the entries,
// thread-interleaved.  thdr = {first word, L | max depth << 24, first piece, end piece}.
struct BlockListScratch {
  std::vector<uint32_t> cnt, e_node, e_code, s_node, s_code, order;
};
inline bool build_cell_lists(uint32_t n, uint32_t PL, const uint32_t* tl, const uint32_t* hl, uint32_t T, uint32_t ntile,
                             BlockListScratch& w, std::vector<uint4>& thdr, std::vector<uint32_t>& lent, std::vector<uint32_t>& piece,
                             uint4* pdesc) {
  const uint32_t B = kFoldThreads, pad = block_pad_entry(PL, T);
  w.cnt.resize((size_t)PL + 1);
  thdr.assign(ntile, make_uint4(0, 0, 0, 0));
  lent.clear();
  piece.clear();
  for (uint32_t t = 0; t < ntile; ++t) {
    const uint32_t t0 = std::min(n, t * T), t1 = std::min(n, t0 + T);
    const uint32_t q0 = (uint32_t)piece.size();
    w.e_node.clear();
    w.e_code.clear();
    uint32_t npieces = 0;
    // tail side: maximal runs of equal tail INSIDE a stage of 128 arcs (loops / padding contribute nothing).  Runs of at
    // least kBPieceMin arcs, up to four per stage, are summed by the compute warp that produces the stage (stage_pieces)
    // and enter the list as one entry that reads the sum from slot T + npieces of the tile buffer.
    for (uint32_t s0 = t0; s0 < t1; s0 += kBStage) {
      const uint32_t s1 = std::min(t1, s0 + kBStage);
      uint32_t used = 0;
      uint32_t d4[4] = {kBNoPiece, kBNoPiece, kBNoPiece, kBNoPiece};
      for (uint32_t i = s0; i < s1;) {
        if (tl[i] == hl[i]) {
          ++i;
          continue;
        }
        uint32_t j = i;
        while (j < s1 && tl[j] == tl[i] && tl[j] != hl[j]) ++j;
        const uint32_t len = j - i;
        if (len >= kBPieceMin && used < 4 && npieces < block_piece_slots(T) - 1) {
          d4[used++] = (i - s0) | ((len - 1) << 8) | (npieces << 16);
          w.e_node.push_back(tl[i]);
          w.e_code.push_back((T + npieces) * 8u);
          ++npieces;
        } else {
          for (uint32_t q = i; q < j; ++q) {
            w.e_node.push_back(tl[i]);
            w.e_code.push_back((q - t0) * 8u);
          }
        }
        i = j;
      }
      pdesc[s0 / kBStage] = make_uint4(d4[0], d4[1], d4[2], d4[3]);
    }
    for (uint32_t i = t0; i < t1; ++i)  // head side
      if (tl[i] != hl[i]) {
        w.e_node.push_back(hl[i]);
        w.e_code.push_back(((i - t0) * 8u) | kBEntMinus);
      }
    const uint32_t ne = (uint32_t)w.e_node.size();
    std::fill(w.cnt.begin(), w.cnt.end(), 0u);
    for (uint32_t e = 0; e < ne; ++e) ++w.cnt[w.e_node[e] + 1];
    for (uint32_t u = 0; u < PL; ++u) w.cnt[u + 1] += w.cnt[u];
    w.s_node.resize(ne);
    w.s_code.resize(ne);
    w.order.assign(w.cnt.begin(), w.cnt.end() - 1);
    for (uint32_t e = 0; e < ne; ++e) {
      const uint32_t dst = w.order[w.e_node[e]]++;
      w.s_node[dst] = w.e_node[e];
      w.s_code[dst] = w.e_code[e];
    }
    const uint32_t L = (ne + B - 1) / B;
    const size_t base = lent.size();
    if (base + (size_t)(L + 1) * B >= 0xffffffffull) return false;
    lent.resize(base + (size_t)(L + 1) * B, pad);
    uint32_t maxdepth = 0, prev_depth = 0;
    for (uint32_t i = 0; i < B; ++i) {
      const uint32_t s0 = std::min(ne, i * L), s1 = std::min(ne, s0 + L);
      uint32_t depth = 0;
      if (s0 < s1 && i > 0 && s0 > 0 && w.s_node[s0] == w.s_node[s0 - 1])
        depth = (w.s_node[s0 - L] == w.s_node[s0 - 1]) ? prev_depth + 1 : 1;  // thread i-1 holds only that node: one deeper
      prev_depth = depth;
      maxdepth = std::max(maxdepth, depth);
      // the accumulator slot that takes the sum of node `u` of this slice: the slice's scratch slot for the share of a node
      // that earlier slices also hold, the node's own slot otherwise
      const uint32_t scratch = PL + kBAccPad + i;
      auto slot_of = [&](uint32_t u) { return depth && u == w.s_node[s0] ? scratch : u; };
      lent[base + i] = depth | ((s0 < s1 ? slot_of(w.s_node[s1 - 1]) : PL) << 8);
      for (uint32_t e = s0; e < s1; ++e) {
        const bool first = e == s0 || w.s_node[e] != w.s_node[e - 1];
        const uint32_t field = e == s0 ? w.s_node[e] : first ? slot_of(w.s_node[e - 1]) : PL;
        lent[base + (size_t)(e - s0 + 1) * B + i] = w.s_code[e] | (field << kBEntNodeShift) | (first ? kBEntNew : 0u);
      }
    }
    if (maxdepth > 255) return false;
    thdr[t] = make_uint4((uint32_t)base, L | (maxdepth << 24), q0, (uint32_t)piece.size());
  }
  return true;
}

inline void build_blocks(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len, int G,
                         size_t smem_limit, HostBlocks& h, int threads = 0) {
  h = HostBlocks{};
  if (m == 0 || p == 0 || G < 1) return;
  const bool timing = std::getenv("TPL_BUILD_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "build_blocks %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  h.GR = std::max<uint32_t>(1, (uint32_t)std::floor(std::sqrt((double)G)));
  h.GC = std::max<uint32_t>(1, (uint32_t)G / h.GR);
  const uint32_t Gc = h.GR * h.GC;
  std::vector<uint64_t> outdeg(p, 0), indeg(p, 0);
  for (size_t j = 0; j < m; ++j) {
    ++outdeg[tail[j]];
    ++indeg[head[j]];
  }
  lap("degrees");
  // blocks over the ACTIVE nodes of each side only (ascending node id): tbn / hbn list them, tbs / hbs cut the lists
  std::vector<uint64_t> wt_t, wt_h;
  std::vector<uint32_t> tpos(p, 0), hpos(p, 0);  // position of a node in tbn / hbn
  for (size_t u = 0; u < p; ++u) {
    if (outdeg[u]) {
      tpos[u] = (uint32_t)h.tbn.size();
      h.tbn.push_back((uint32_t)u);
      wt_t.push_back(outdeg[u]);
    }
    if (indeg[u]) {
      hpos[u] = (uint32_t)h.hbn.size();
      h.hbn.push_back((uint32_t)u);
      wt_h.push_back(indeg[u]);
    }
  }
  h.tbs = weight_blocks(wt_t, h.GR);
  h.hbs = weight_blocks(wt_h, h.GC);
  uint32_t PT = 0, PH = 0;
  for (uint32_t a = 0; a < h.GR; ++a) PT = std::max(PT, h.tbs[a + 1] - h.tbs[a]);
  for (uint32_t b = 0; b < h.GC; ++b) PH = std::max(PH, h.hbs[b + 1] - h.hbs[b]);
  PT = std::max(2u, (PT + 1u) & ~1u);  // even: 16-byte aligned shared-memory arrays
  PH = std::max(2u, (PH + 1u) & ~1u);
  if (PT > 0x8000u || PH > 0x8000u || PT + PH + kBAccPad + kFoldThreads > kBMaxLocalNodes + 1) return;  // 15-bit local ids, 14-bit list slots
  h.PT = PT;
  h.PH = PH;
  std::vector<uint32_t> node_tb(p, 0), node_hb(p, 0);
  for (uint32_t a = 0; a < h.GR; ++a)
    for (uint32_t q = h.tbs[a]; q < h.tbs[a + 1]; ++q) node_tb[h.tbn[q]] = a;
  for (uint32_t b = 0; b < h.GC; ++b)
    for (uint32_t q = h.hbs[b]; q < h.hbs[b + 1]; ++q) node_hb[h.hbn[q]] = b;

  lap("blocks");
  // arcs by tail node, then by arc index (identity for a tail-grouped arc list) ...
  std::vector<uint32_t> by_tail(m);
  {
    std::vector<uint64_t> ptr(p + 1, 0);
    for (size_t u = 0; u < p; ++u) ptr[u + 1] = ptr[u] + outdeg[u];
    for (size_t j = 0; j < m; ++j) by_tail[ptr[tail[j]]++] = (uint32_t)j;
  }
  lap("sort by tail");
  // ... then stably by cell
  std::vector<uint64_t> cnt(Gc + 1, 0);
  auto cell_of = [&](uint32_t j) { return node_tb[tail[j]] * h.GC + node_hb[head[j]]; };
  for (size_t j = 0; j < m; ++j) ++cnt[cell_of((uint32_t)j) + 1];
  h.cell_off.assign(Gc + 1, 0);
  uint64_t off = 0, max_cell = 0;
  for (uint32_t c = 0; c < Gc; ++c) {
    h.cell_off[c] = (uint32_t)off;
    max_cell = std::max<uint64_t>(max_cell, cnt[c + 1]);
    off += (cnt[c + 1] + kBStage - 1) / kBStage * kBStage;
    if (off >= 0xfffff000ull) return;  // 32-bit positions
  }
  h.cell_off[Gc] = (uint32_t)off;
  h.Mpad = (uint32_t)off;

  lap("cell counts");
  h.d.assign(h.Mpad, 0.0);
  h.th.assign(h.Mpad, kBLoop);
  h.gidx.assign(h.Mpad, kBPad);
  {
    std::vector<uint32_t> fill(h.cell_off.begin(), h.cell_off.end() - 1);
    for (size_t q = 0; q < m; ++q) {
      const uint32_t j = by_tail[q], c = cell_of(j), pos = fill[c]++;
      const uint32_t t = tail[j], hd = head[j];
      h.gidx[pos] = j;
      h.d[pos] = j < d_len ? d[j] : 0.0;
      h.th[pos] = (tpos[t] - h.tbs[node_tb[t]]) | ((hpos[hd] - h.hbs[node_hb[hd]]) << 15) | (t < hd ? kBTailFirst : 0u) | (t == hd ? kBLoop : 0u);
    }
  }
  std::vector<uint32_t>().swap(by_tail);
  lap("cell order (d, th, gidx)");

  // tile lists per cell, over local node ids; a loop / padding slot gets tail == head (the list builder skips those).
  // Tile size: the largest multiple of 1024 arcs (at most 4096) whose kernels fit with three ring slots in pass 2 (bytes in
  // flight matter more than hand-offs per sweep), else the largest that fits with two.  The list buffers depend on the lists
  // themselves ((L + 1) KB per tile), so a size that does not fit is rebuilt one step smaller.
  if (threads <= 0) threads = m < (1u << 20) ? 1 : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
  threads = std::min<int>(threads, (int)Gc);
  std::vector<std::vector<uint4>> thdr(Gc);
  std::vector<std::vector<uint32_t>> lent(Gc), piece(Gc);
  const uint32_t padded_max = (uint32_t)((max_cell + kBStage - 1) / kBStage * kBStage);
  // the largest tile that fits (up to the cell itself): per-tile hand-offs and depth phases cost more than the overlap of
  // fold and stream gains, at every size (1 M arcs: 0.47 of the HBM peak with one 3072-arc tile rule against 0.43 with
  // eight tiles per cell; 3 M arcs: 0.71 against 0.60)
  uint32_t want = std::min<uint32_t>(4096, std::max<uint32_t>(1024, (padded_max + 1023) / 1024 * 1024));
  if (const char* e = std::getenv("TPL_BLOCK_T")) want = std::min<uint32_t>(4096, std::max<uint32_t>(1024, (uint32_t)std::atoi(e) / 1024 * 1024));  // tuning experiments
  if (const char* e = std::getenv("TPL_BLOCK_NTB")) h.ntb = std::min<uint32_t>(kBMaxTileBufs, std::max<uint32_t>(2, (uint32_t)std::atoi(e)));  // tuning experiments
  bool done = false;
  for (int need = 3; need >= 2 && !done; --need)
    for (uint32_t T = want; T >= 1024 && !done; T -= 1024) {
      // cheap bound first: L >= entries / threads >= (arcs of a full tile) / threads
      if (!blocks_fit(PT + PH, smem_limit, T, (std::min(T, padded_max) / kFoldThreads + 1) * 4u * kFoldThreads, h.ntb, std::max<uint32_t>(1, (padded_max + T - 1) / T), need, h.ring1, h.ring2, h.ring2v))
        continue;
      h.T = T;
      h.ntile = std::max<uint32_t>(1, (padded_max + T - 1) / T);
      std::vector<uint8_t> failed(Gc, 0);
      h.pdesc.assign(h.Mpad / kBStage, make_uint4(kBNoPiece, kBNoPiece, kBNoPiece, kBNoPiece));
      auto run = [&](int first) {
        BlockListScratch w;
        std::vector<uint32_t> tl, hl;
        for (uint32_t c = (uint32_t)first; c < Gc; c += (uint32_t)threads) {
          const uint32_t c0 = h.cell_off[c], n = h.cell_off[c + 1] - c0;
          tl.resize(n);
          hl.resize(n);
          for (uint32_t i = 0; i < n; ++i) {
            const uint32_t w32 = h.th[c0 + i];
            const bool skip = (w32 & kBLoop) != 0;
            tl[i] = skip ? 0u : (w32 & 0x7fffu);
            hl[i] = skip ? 0u : PT + ((w32 >> 15) & 0x7fffu);
          }
          failed[c] = !build_cell_lists(n, PT + PH, tl.data(), hl.data(), h.T, h.ntile, w, thdr[c], lent[c], piece[c], h.pdesc.data() + c0 / kBStage);
        }
      };
      if (threads == 1) {
        run(0);
      } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(run, t);
        for (std::thread& t : pool) t.join();
      }
      uint32_t Lmax = 0;
      bool bad = false;
      for (uint32_t c = 0; c < Gc; ++c) {
        bad = bad || failed[c];
        for (const uint4& x : thdr[c]) Lmax = std::max(Lmax, x.y & 0xffffffu);
      }
      if (bad) return;
      h.lblk = (Lmax + 1) * 4u * kFoldThreads;
      done = blocks_fit(PT + PH, smem_limit, T, h.lblk, h.ntb, h.ntile, need, h.ring1, h.ring2, h.ring2v);
    }
  lap("tile lists");
  if (!done) return;
  {  // as many list buffers as still fit next to the rings (deeper list prefetch; a short sweep then starts with its whole list in flight)
    const size_t budget = smem_limit > 3072 ? smem_limit - 3072 : 0;
    h.nl = 2;
    while (h.nl < (uint32_t)kBMaxList && h.nl < h.ntile &&
           block_smem_bytes(PT + PH, h.T, (int)h.ring2v, h.lblk, h.nl + 1, h.ntb, true, true, h.ntile) <= budget &&
           block_smem_bytes(PT + PH, h.T, (int)h.ring2, h.lblk, h.nl + 1, h.ntb, true, false, h.ntile) <= budget &&
           block_smem_bytes(PT + PH, h.T, (int)h.ring1, h.lblk, h.nl + 1, h.ntb, false, false, h.ntile) <= budget)
      ++h.nl;
  }
  size_t nl = 0, np = 0;
  for (uint32_t c = 0; c < Gc; ++c) {
    nl += lent[c].size();
    np += piece[c].size();
  }
  if (nl >= 0xffffffffull) return;  // entry offsets are 32-bit
  h.thdr.assign((size_t)Gc * h.ntile, make_uint4(0, 0, 0, 0));
  h.lent.reserve(nl);
  h.piece.reserve(np);
  for (uint32_t c = 0; c < Gc; ++c) {
    const uint32_t lb = (uint32_t)h.lent.size(), pb = (uint32_t)h.piece.size();
    for (uint32_t t = 0; t < h.ntile; ++t) {
      const uint4 x = thdr[c][t];
      h.thdr[(size_t)c * h.ntile + t] = make_uint4(x.x + lb, x.y, x.z + pb, x.w + pb);
    }
    h.lent.insert(h.lent.end(), lent[c].begin(), lent[c].end());
    h.piece.insert(h.piece.end(), piece[c].begin(), piece[c].end());
    std::vector<uint32_t>().swap(lent[c]);
    std::vector<uint32_t>().swap(piece[c]);
  }
  lap("concatenate");
  h.ok = true;
}

// 0 when the lists of a cell hold, for every tile, each non-loop arc exactly once on its local head (minus) and once on its
// local tail (directly or inside one piece), slices have equal length with padding only behind the last entry, the new-node
// flags mark exactly the node changes inside a slice, entries are sorted by node across the slices, and every thread's depth
// is its position in the chain of threads that share its first node.
inline int check_cell_lists(uint32_t n, uint32_t PL, const uint32_t* tl, const uint32_t* hl, uint32_t T, uint32_t ntile,
                            const uint4* thdr, const std::vector<uint32_t>& lent, const uint4* pdesc) {
  const uint32_t B = kFoldThreads, pad = block_pad_entry(PL, T);
  std::vector<uint8_t> seen_t(T), seen_h(T);
  for (uint32_t t = 0; t < ntile; ++t) {
    const uint4 hd = thdr[t];
    const uint32_t L = hd.y & 0xffffffu, D = hd.y >> 24;
    const uint32_t t0 = std::min(n, t * T), t1 = std::min(n, t0 + T), na = t1 - t0;
    if ((size_t)hd.x + (size_t)(L + 1) * B > lent.size()) return 2;
    // the tile's runs: slot -> (first arc, length), from the stage descriptors
    std::vector<std::pair<uint32_t, uint32_t>> runs;
    for (uint32_t s0 = t0; s0 < t1; s0 += kBStage) {
      const uint4 d = pdesc[s0 / kBStage];
      for (uint32_t dd : {d.x, d.y, d.z, d.w}) {
        if (dd == kBNoPiece) continue;
        const uint32_t slot = dd >> 16, start = dd & 0xffu, len = ((dd >> 8) & 0xffu) + 1;
        if (slot != runs.size() || start + len > std::min<uint32_t>(kBStage, t1 - s0) || len < kBPieceMin) return 3;
        runs.emplace_back(s0 - t0 + start, len);
      }
    }
    if (runs.size() > block_piece_slots(T) - 1) return 3;
    std::fill(seen_t.begin(), seen_t.end(), 0);
    std::fill(seen_h.begin(), seen_h.end(), 0);
    uint32_t prev_node = 0, prev_depth = 0, maxdepth = 0;
    bool prev_single = false, any = false, ended = false;
    for (uint32_t i = 0; i < B; ++i) {
      const uint32_t depth = lent[hd.x + i] & 0xffu, final_slot = lent[hd.x + i] >> 8;
      uint32_t first_node = 0, last_node = 0, cnt = 0;
      bool single = true;
      for (uint32_t q = 0; q < L; ++q) {
        const uint32_t e = lent[hd.x + (size_t)(q + 1) * B + i];
        if (e == pad) {
          ended = true;  // padding only behind the very last entry of the tile
          continue;
        }
        if (ended) return 4;
        const uint32_t field = (e >> kBEntNodeShift) & kBEntNodeMask, idx = (e & kBEntOffMask) / 8u;
        if ((e & kBEntOffMask) % 8u) return 17;
        // the node is implied by what the entry reads: a head (minus), a run sum, or a tail
        uint32_t node;
        if (e & kBEntMinus) {
          if (idx >= na) return 9;
          node = hl[t0 + idx];
        } else if (idx >= T) {
          if (idx - T >= runs.size()) return 10;
          node = tl[t0 + runs[idx - T].first];
        } else {
          if (idx >= na) return 13;
          node = tl[t0 + idx];
        }
        if (node >= PL) return 5;
        const bool isnew = (e & kBEntNew) != 0;
        const uint32_t scratch = PL + kBAccPad + i;
        if (cnt == 0) {
          if (!isnew || field != node) return 6;  // the first entry names the slice's first node (used by the depth phases)
          first_node = node;
        } else {
          if (isnew != (node != last_node) || node < last_node) return 7;
          // a node change flushes the finished node: into the scratch slot when it is the slice's shared first node
          const uint32_t target = depth && last_node == first_node ? scratch : last_node;
          if (field != (isnew ? target : PL)) return 18;
          if (node != last_node) single = false;
        }
        if (any && cnt == 0 && node < prev_node) return 8;  // sorted across slices
        last_node = node;
        ++cnt;
        if (e & kBEntMinus) {  // head side
          if (idx >= na || hl[t0 + idx] != node || seen_h[idx]) return 9;
          seen_h[idx] = 1;
        } else if (idx >= T) {  // a same-tail run summed by the compute warp
          if (idx - T >= runs.size()) return 10;
          const uint32_t start = runs[idx - T].first, len = runs[idx - T].second;
          if (start + len > na) return 11;
          for (uint32_t a = start; a < start + len; ++a) {
            if (tl[t0 + a] != node || seen_t[a]) return 12;
            seen_t[a] = 1;
          }
        } else {
          if (idx >= na || tl[t0 + idx] != node || seen_t[idx]) return 13;
          seen_t[idx] = 1;
        }
      }
      uint32_t want = 0;
      if (cnt && any && first_node == prev_node) want = prev_single ? prev_depth + 1 : 1;
      if (depth != want) return 14;
      if (final_slot != (cnt == 0 ? PL : depth && last_node == first_node ? PL + kBAccPad + i : last_node)) return 19;
      maxdepth = std::max(maxdepth, depth);
      if (cnt) {
        prev_node = last_node;
        prev_single = single;
        prev_depth = depth;
        any = true;
      }
    }
    if (maxdepth != D) return 15;
    for (uint32_t a = 0; a < na; ++a) {
      const bool loop = tl[t0 + a] == hl[t0 + a];
      if (seen_t[a] != (loop ? 0 : 1) || seen_h[a] != (loop ? 0 : 1)) return 16;
    }
    // The walk itself, as block_fold_tile performs it (tpl_blocks.cuh), on small integers (every sum exact): tile buffer =
    // arc values, run sums behind them, zero slot last; every slice adds its entries into a running sum, an entry that opens
    // a node first adds the finished sum to the slot it names (the first entry of a slice names its node instead: nothing to
    // flush), the last sum goes to the slot of row 0; then the scratch shares, depth by depth.  The accumulators must hold
    // +x on the tail and -x on the head of every non-loop arc.
    {
      const uint32_t slots = block_piece_slots(T), nacc = PL + kBAccPad + B;
      std::vector<double> wt((size_t)T + slots, 0.0), acc(nacc, 0.0), want(PL, 0.0);
      for (uint32_t a = 0; a < na; ++a) {
        wt[a] = (double)((int)((t0 + a) * 7u % 13u) - 6);
        if (tl[t0 + a] != hl[t0 + a]) {
          want[tl[t0 + a]] += wt[a];
          want[hl[t0 + a]] -= wt[a];
        }
      }
      for (size_t r = 0; r < runs.size(); ++r)
        for (uint32_t a = runs[r].first; a < runs[r].first + runs[r].second; ++a) wt[T + r] += wt[a];
      const uint32_t dummy = PL;
      for (uint32_t i = 0; i < B; ++i) {
        const uint32_t row0 = lent[hd.x + i], depth = row0 & 0xffu;
        if (depth) acc[PL + kBAccPad + i] = 0.0;
        double sum = 0.0;
        for (uint32_t q = 0; q < L; ++q) {
          const uint32_t e = lent[hd.x + (size_t)(q + 1) * B + i];
          const double x = wt[(e & kBEntOffMask) / 8u], val = (e & kBEntMinus) ? -x : x;
          if (e & kBEntNew) {
            if (q != 0) acc[(e >> kBEntNodeShift) & kBEntNodeMask] += sum;
            sum = val;
          } else {
            sum += val;
          }
        }
        acc[(row0 >> 8) & kBEntNodeMask] += sum;
      }
      for (uint32_t d = 1; d <= D; ++d)
        for (uint32_t i = 0; i < B; ++i)
          if ((lent[hd.x + i] & 0xffu) == d) {
            const uint32_t e1 = lent[hd.x + B + i];
            acc[(e1 >> kBEntNodeShift) & kBEntNodeMask] += acc[PL + kBAccPad + i];
          }
      (void)dummy;
      for (uint32_t u = 0; u < PL; ++u)
        if (acc[u] != want[u]) return 20;
    }
  }
  return 0;
}

// 0 when the layout is consistent: gidx is a bijection between the non-padding positions and the arcs; every position's
// th word decodes to its arc's tail / head inside the cell's blocks, with the right order / loop flags; d is carried over;
// cells are padded to stage multiples; inside a cell the arcs are sorted by (tail, arc index); the tile lists of every cell
// hold each non-loop arc once on its local tail and once on its local head (check_tiles on the cell as a one-CTA instance).
inline int check_blocks(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len,
                        const HostBlocks& h) {
  const uint32_t Gc = h.GR * h.GC;
  if (!h.ok || h.cell_off.size() != Gc + 1 || h.tbs.size() != h.GR + 1 || h.hbs.size() != h.GC + 1) return 1;
  if (h.tbs[0] != 0 || h.tbs[h.GR] != h.tbn.size() || h.hbs[0] != 0 || h.hbs[h.GC] != h.hbn.size()) return 2;
  for (size_t q = 0; q < h.tbn.size(); ++q)
    if (h.tbn[q] >= p || (q && h.tbn[q] <= h.tbn[q - 1])) return 2;
  for (size_t q = 0; q < h.hbn.size(); ++q)
    if (h.hbn[q] >= p || (q && h.hbn[q] <= h.hbn[q - 1])) return 2;
  if (h.th.size() != h.Mpad || h.gidx.size() != h.Mpad || h.d.size() != h.Mpad || h.cell_off[Gc] != h.Mpad) return 3;
  if (h.T % 1024 != 0 || h.T == 0 || (h.PT & 1u) || (h.PH & 1u)) return 4;
  std::vector<uint8_t> seen(m, 0);
  std::vector<uint32_t> tl, hl;
  for (uint32_t c = 0; c < Gc; ++c) {
    const uint32_t r = c / h.GC, cc = c % h.GC, c0 = h.cell_off[c], c1 = h.cell_off[c + 1];
    if (c0 > c1 || c0 % kBStage || c1 % kBStage) return 5;
    if ((c1 - c0 + h.T - 1) / h.T > h.ntile) return 6;
    uint64_t prev_key = 0;
    bool in_pad = false;
    tl.assign(c1 - c0, 0);
    hl.assign(c1 - c0, 0);
    for (uint32_t pos = c0; pos < c1; ++pos) {
      const uint32_t g = h.gidx[pos], w = h.th[pos];
      if (g == kBPad) {
        in_pad = true;
        if (!(w & kBLoop) || h.d[pos] != 0.0) return 7;
        continue;
      }
      if (in_pad) return 8;  // padding only at the end of a cell
      if (g >= m || seen[g]) return 9;
      seen[g] = 1;
      const uint32_t t = tail[g], hd = head[g];
      const uint32_t lt = w & 0x7fffu, lh = (w >> 15) & 0x7fffu;
      if (lt >= h.tbs[r + 1] - h.tbs[r] || lh >= h.hbs[cc + 1] - h.hbs[cc]) return 10;
      if (h.tbn[h.tbs[r] + lt] != t || h.hbn[h.hbs[cc] + lh] != hd) return 11;
      if (((w & kBLoop) != 0) != (t == hd) || ((w & kBTailFirst) != 0) != (t < hd)) return 12;
      if (h.d[pos] != (g < d_len ? d[g] : 0.0)) return 13;
      const uint64_t key = ((uint64_t)t << 32) | g;
      if (pos > c0 && key <= prev_key) return 14;
      prev_key = key;
      if (t != hd) {
        tl[pos - c0] = w & 0x7fffu;
        hl[pos - c0] = h.PT + ((w >> 15) & 0x7fffu);
      }
    }
    const int rc = check_cell_lists(c1 - c0, h.PT + h.PH, tl.data(), hl.data(), h.T, h.ntile, h.thdr.data() + (size_t)c * h.ntile,
                                    h.lent, h.pdesc.data() + c0 / kBStage);
    if (rc) return 100 + rc;
  }
  for (size_t j = 0; j < m; ++j)
    if (!seen[j]) return 15;
  return 0;
}

}  // namespace tpl
