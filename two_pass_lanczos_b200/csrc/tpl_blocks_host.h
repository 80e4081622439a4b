// tpl_blocks_host.h -- host construction of the blocked streaming layout (tpl_blocks.cuh) and the checker the CPU tests run
// over it (tpl_blocks_plan).
//
//   1. node blocks: GR contiguous tail blocks with (nearly) equal out-degree, GC contiguous head blocks with (nearly) equal
//      in-degree; G = GR * GC cells;
//   2. cell order: arc j belongs to cell (tail block of tail_j) * GC + (head block of head_j); inside a cell the arcs are
//      ordered by tail node, then by arc index (two stable counting sorts), every cell is padded to a multiple of 128 slots;
//   3. per position: d, the packed th word (local tail | local head << 15 | tail-first | loop-or-padding) and gidx;
//   4. tile lists of every cell over LOCAL node ids (tails [0, PT), heads [PT, PT + PH)) by the builder of the tiled
//      kernels (tpl_tiles_host.h): a cell is handed to it as a one-CTA instance; a fold thread's entries stay sorted by node
//      and list padding becomes a harmless entry (the tile's zero slot into a dummy accumulator).
// Cells are independent: they are built by a pool of host threads and concatenated in cell order (the result does not depend
// on the number of threads).
#pragma once
#include <algorithm>
#include <cmath>
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
  uint32_t ring1 = 0, ring2 = 0, ring2v = 0;
  std::vector<uint32_t> cell_off, tbs, hbs, th, gidx;
  std::vector<double> d;
  std::vector<uint4> thdr;
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

// Tile size (multiple of 1024 arcs, at most 4096) and ring depths for which the kernels fit in `smem_limit` bytes: the largest
// tile that still leaves three ring slots to pass 2 (bytes in flight matter more than hand-offs per sweep), else the largest
// tile that fits with two.
inline bool blocks_fit(uint32_t PL, size_t smem_limit, size_t max_cell, uint32_t& T, uint32_t& ring1, uint32_t& ring2, uint32_t& ring2v) {
  const size_t budget = smem_limit > 3072 ? smem_limit - 3072 : 0;  // static shared memory of the kernels + margin
  const uint32_t want = (uint32_t)std::min<size_t>(4096, std::max<size_t>(1024, (max_cell + 1023) / 1024 * 1024));
  for (int need = 3; need >= 2; --need)
    for (uint32_t t = want; t >= 1024; t -= 1024) {
      auto fits = [&](int ring, bool pass2, bool v) { return block_smem_bytes(PL, t, ring, pass2, v) <= budget; };
      if (!fits(need, true, false) || !fits(2, true, true) || !fits(2, false, false)) continue;
      T = t;
      ring2 = fits(4, true, false) ? 4 : fits(3, true, false) ? 3 : 2;
      ring2v = fits(3, true, true) ? 3 : 2;
      ring1 = fits(4, false, false) ? 4 : fits(3, false, false) ? 3 : 2;
      return true;
    }
  return false;
}

inline void build_blocks(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len, int G,
                         size_t smem_limit, HostBlocks& h, int threads = 0) {
  h = HostBlocks{};
  if (m == 0 || p == 0 || G < 1) return;
  h.GR = std::max<uint32_t>(1, (uint32_t)std::floor(std::sqrt((double)G)));
  h.GC = std::max<uint32_t>(1, (uint32_t)G / h.GR);
  const uint32_t Gc = h.GR * h.GC;
  std::vector<uint64_t> outdeg(p, 0), indeg(p, 0);
  for (size_t j = 0; j < m; ++j) {
    ++outdeg[tail[j]];
    ++indeg[head[j]];
  }
  h.tbs = weight_blocks(outdeg, h.GR);
  h.hbs = weight_blocks(indeg, h.GC);
  uint32_t PT = 0, PH = 0;
  for (uint32_t a = 0; a < h.GR; ++a) PT = std::max(PT, h.tbs[a + 1] - h.tbs[a]);
  for (uint32_t b = 0; b < h.GC; ++b) PH = std::max(PH, h.hbs[b + 1] - h.hbs[b]);
  PT = (PT + 1u) & ~1u;  // even: 16-byte aligned shared-memory arrays
  PH = (PH + 1u) & ~1u;
  if (PT > 0x8000u || PH > 0x8000u || PT + PH >= (1u << 17)) return;  // 15-bit local ids, 17-bit list nodes
  h.PT = PT;
  h.PH = PH;
  std::vector<uint32_t> node_tb(p), node_hb(p);
  for (uint32_t a = 0; a < h.GR; ++a)
    for (uint32_t u = h.tbs[a]; u < h.tbs[a + 1]; ++u) node_tb[u] = a;
  for (uint32_t b = 0; b < h.GC; ++b)
    for (uint32_t u = h.hbs[b]; u < h.hbs[b + 1]; ++u) node_hb[u] = b;

  // arcs by tail node, then by arc index (identity for a tail-grouped arc list) ...
  std::vector<uint32_t> by_tail(m);
  {
    std::vector<uint64_t> ptr(p + 1, 0);
    for (size_t u = 0; u < p; ++u) ptr[u + 1] = ptr[u] + outdeg[u];
    for (size_t j = 0; j < m; ++j) by_tail[ptr[tail[j]]++] = (uint32_t)j;
  }
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
  if (!blocks_fit(PT + PH, smem_limit, (size_t)max_cell, h.T, h.ring1, h.ring2, h.ring2v)) return;
  h.ntile = (uint32_t)std::max<uint64_t>(1, ((max_cell + kBStage - 1) / kBStage * kBStage + h.T - 1) / h.T);

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
      h.th[pos] = (t - h.tbs[node_tb[t]]) | ((hd - h.hbs[node_hb[hd]]) << 15) | (t < hd ? kBTailFirst : 0u) | (t == hd ? kBLoop : 0u);
    }
  }
  std::vector<uint32_t>().swap(by_tail);

  // tile lists per cell, over local node ids; a loop / padding slot gets tail == head (the list builder skips those)
  if (threads <= 0) threads = m < (1u << 20) ? 1 : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
  threads = std::min<int>(threads, (int)Gc);
  std::vector<std::vector<uint4>> thdr(Gc);
  std::vector<std::vector<uint32_t>> lent(Gc), piece(Gc);
  auto run = [&](int first) {
    TileScratch w;
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
      build_cta_tiles(n, PT + PH, tl.data(), hl.data(), 1, h.T, h.ntile, 0, w, thdr[c], lent[c], piece[c], false, kMaxPieces - 1);
      const uint32_t pad = block_pad_entry(PT + PH, h.T);  // padding = zero slot of the tile into the dummy accumulator
      for (uint32_t& e : lent[c])
        if (e == kEntPad) e = pad;
    }
  };
  if (threads == 1) {
    run(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(run, t);
    for (std::thread& t : pool) t.join();
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
  h.ok = true;
}

// 0 when the layout is consistent: gidx is a bijection between the non-padding positions and the arcs; every position's
// th word decodes to its arc's tail / head inside the cell's blocks, with the right order / loop flags; d is carried over;
// cells are padded to stage multiples; inside a cell the arcs are sorted by (tail, arc index); the tile lists of every cell
// hold each non-loop arc once on its local tail and once on its local head (check_tiles on the cell as a one-CTA instance).
inline int check_blocks(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len,
                        const HostBlocks& h) {
  const uint32_t Gc = h.GR * h.GC;
  if (!h.ok || h.cell_off.size() != Gc + 1 || h.tbs.size() != h.GR + 1 || h.hbs.size() != h.GC + 1) return 1;
  if (h.tbs[0] != 0 || h.tbs[h.GR] != p || h.hbs[0] != 0 || h.hbs[h.GC] != p) return 2;
  if (h.th.size() != h.Mpad || h.gidx.size() != h.Mpad || h.d.size() != h.Mpad || h.cell_off[Gc] != h.Mpad) return 3;
  if (h.T % 1024 != 0 || h.T == 0 || (h.PT & 1u) || (h.PH & 1u)) return 4;
  std::vector<uint8_t> seen(m, 0);
  std::vector<uint32_t> tl, hl;
  HostTiles one;  // a cell's lists as a one-CTA instance: thdr offsets are global, so lent / piece stay whole
  one.T = h.T;
  one.ntile = h.ntile;
  one.lent = h.lent;
  one.piece = h.piece;
  for (uint32_t& e : one.lent)
    if (e == block_pad_entry(h.PT + h.PH, h.T)) e = kEntPad;
  // a thread's entries are sorted by node (the fold adds a node's values in a register and flushes on a node change)
  for (size_t t = 0; t < h.thdr.size(); ++t) {
    const uint4 hd = h.thdr[t];
    for (uint32_t i = 0; i < (uint32_t)kFoldThreads; ++i) {
      uint32_t prev = 0;
      bool padded = false;
      for (uint32_t q = 0; q < hd.y; ++q) {
        const uint32_t e = one.lent[hd.x + (size_t)q * kFoldThreads + i];
        if (e == kEntPad) {
          padded = true;
          continue;
        }
        if (padded || (e >> 15) < prev) return 16;  // padding only at the end, nodes non-decreasing
        prev = e >> 15;
      }
    }
  }
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
      if (t < h.tbs[r] || t >= h.tbs[r + 1] || hd < h.hbs[cc] || hd >= h.hbs[cc + 1]) return 10;
      if ((w & 0x7fffu) != t - h.tbs[r] || ((w >> 15) & 0x7fffu) != hd - h.hbs[cc]) return 11;
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
    one.thdr.assign(h.thdr.begin() + (size_t)c * h.ntile, h.thdr.begin() + (size_t)(c + 1) * h.ntile);
    const int rc = check_tiles(c1 - c0, h.PT + h.PH, tl.data(), hl.data(), 1, one);
    if (rc) return 100 + rc;
  }
  for (size_t j = 0; j < m; ++j)
    if (!seen[j]) return 15;
  return 0;
}

}  // namespace tpl
