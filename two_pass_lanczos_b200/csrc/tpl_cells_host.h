// tpl_cells_host.h -- host-side construction of the 2-D cell partition used by the kernels of tpl_cells.cuh.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "tpl_cells.cuh"

namespace tpl {

struct HostCells {
  bool ok = false;
  uint32_t GR = 0, GC = 0, Gc = 0, Amax = 0, L = 0;
  uint32_t max_lines = 0, max_slots = 0, max_groups = 0, max_rows = 0, max_own = 0, max_tl = 0, inbox_atoms = 0;
  std::vector<uint32_t> hdr, gidx, lth, lines, push, slot_base;
  std::vector<uint16_t> tmap;
  std::vector<uint4> walk;
  std::vector<uint2> ent4;
};

// the sizing fields of the device-side descriptor (what cell_smem_bytes needs)
inline CellOp cells_probe(const HostCells& h) {
  CellOp co{};
  co.GR = h.GR; co.GC = h.GC; co.Gc = h.Gc; co.Amax = h.Amax; co.L = h.L;
  co.max_lines = h.max_lines; co.max_slots = h.max_slots; co.max_groups = h.max_groups; co.max_rows = h.max_rows;
  co.max_own = h.max_own; co.max_tl = h.max_tl; co.inbox_atoms = h.inbox_atoms;
  return co;
}

// Contiguous blocks of [0, p) with (nearly) equal total weight: bnd[a] = first node of block a.
inline std::vector<uint32_t> balanced_blocks(const std::vector<uint64_t>& weight, uint32_t parts) {
  const size_t p = weight.size();
  const uint64_t total = std::accumulate(weight.begin(), weight.end(), (uint64_t)0);
  std::vector<uint32_t> bnd(parts + 1, (uint32_t)p);
  bnd[0] = 0;
  uint64_t pre = 0;
  uint32_t a = 1;
  for (size_t u = 0; u < p && a < parts; ++u) {
    // node u starts block a when the weight before it (plus half its own) reaches a/parts of the total; without any
    // weight the nodes themselves are split evenly
    while (a < parts && (total > 0 ? (2 * pre + weight[u]) * parts >= 2 * (uint64_t)a * total
                                   : (uint64_t)u * parts >= (uint64_t)a * p))
      bnd[a++] = (uint32_t)u;
    pre += weight[u];
  }
  for (; a < parts; ++a) bnd[a] = (uint32_t)p;
  // Blocks end on line boundaries whenever there are enough lines: a line cut by a block boundary would collect the
  // contributions of two rows (or columns) of cells, and its owner -- which every all-reduce waits for -- would be the last
  // to have its node sums complete (measured: 36 instead of 24 contributions cost its owner ~2x the wait).
  if (p >= (size_t)parts * kLine * 8)  // at least 8 lines per block: rounding then costs < ~10 % of balance
    for (uint32_t i = 1; i < parts; ++i) bnd[i] = std::min<uint32_t>((uint32_t)p, (bnd[i] + kLine / 2) / kLine * kLine);
  for (uint32_t i = 1; i <= parts; ++i) bnd[i] = std::max(bnd[i], bnd[i - 1]);
  return bnd;
}

// Builds the partition for a grid of at most G CTAs.  `smem_limit` is the opt-in shared memory per CTA; the result is
// marked !ok when a cell does not fit (the caller then keeps the other execution shapes).
inline void build_cells(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int G, size_t smem_limit,
                        HostCells& h) {
  h.ok = false;
  if (m == 0 || p == 0 || G < 1 || p > (1u << 22) || m > (size_t)G * kCellArcs) return;
  h.GR = std::max<uint32_t>(1, (uint32_t)std::floor(std::sqrt((double)G)));
  h.GC = std::max<uint32_t>(1, (uint32_t)G / h.GR);
  h.Gc = h.GR * h.GC;
  h.L = (uint32_t)((p + kLine - 1) / kLine);
  const uint32_t Gc = h.Gc;

  std::vector<uint64_t> outdeg(p, 0), indeg(p, 0);
  for (size_t j = 0; j < m; ++j) {
    ++outdeg[tail[j]];
    ++indeg[head[j]];
  }
  const std::vector<uint32_t> tb = balanced_blocks(outdeg, h.GR), hb = balanced_blocks(indeg, h.GC);
  std::vector<uint32_t> node_tb(p), node_hb(p);
  for (uint32_t a = 0; a < h.GR; ++a)
    for (uint32_t u = tb[a]; u < tb[a + 1]; ++u) node_tb[u] = a;
  for (uint32_t b = 0; b < h.GC; ++b)
    for (uint32_t u = hb[b]; u < hb[b + 1]; ++u) node_hb[u] = b;

  // arcs of every cell, ascending arc index
  std::vector<uint32_t> cell_ptr(Gc + 1, 0), cell_arc(m);
  auto cell_of = [&](size_t j) { return node_tb[tail[j]] * h.GC + node_hb[head[j]]; };
  for (size_t j = 0; j < m; ++j) ++cell_ptr[cell_of(j) + 1];
  for (uint32_t c = 0; c < Gc; ++c) cell_ptr[c + 1] += cell_ptr[c];
  {
    std::vector<uint32_t> fill(cell_ptr.begin(), cell_ptr.end() - 1);
    for (size_t j = 0; j < m; ++j) cell_arc[fill[cell_of(j)]++] = (uint32_t)j;
  }
  uint32_t Amax = 1;
  for (uint32_t c = 0; c < Gc; ++c) Amax = std::max(Amax, cell_ptr[c + 1] - cell_ptr[c]);
  if (Amax > kCellArcs) return;  // a thread keeps at most kArcRegs arcs in registers
  h.Amax = (Amax + 31u) & ~31u;

  struct List {
    uint32_t node;
    bool head;
    std::vector<uint16_t> pos;  // arc positions inside the cell, in summation order
  };
  struct Cell {
    std::vector<List> lists;                  // tail lists by rank (longest first), then head lists (longest first)
    std::vector<uint32_t> tlines, hlines, lines;
    std::vector<uint32_t> tnode;              // tail node of every rank
    std::vector<uint32_t> arc_code;           // per position: tail rank | tail-first flag << 15 | local head << 16; ~0 = self-loop / empty
    uint32_t nA = 0, njds = 0, nrows = 0;
  };
  std::vector<Cell> cells(Gc);
  h.hdr.assign((size_t)Gc * 8, 0);
  h.gidx.assign((size_t)Gc * h.Amax, 0xffffffffu);
  h.lth.assign((size_t)Gc * h.Amax, 0);
  std::vector<uint32_t> tcount(p, 0), trank(p, 0), hcount(p, 0), hfill(p, 0), local_of_line(h.L, 0xffffffffu);

  for (uint32_t c = 0; c < Gc; ++c) {
    Cell& ce = cells[c];
    const uint32_t* arcs = cell_arc.data() + cell_ptr[c];
    const uint32_t nA = cell_ptr[c + 1] - cell_ptr[c];
    ce.nA = nA;
    std::vector<uint32_t> tnodes, hnodes;
    for (uint32_t i = 0; i < nA; ++i) {
      const uint32_t j = arcs[i];
      if (tail[j] == head[j]) continue;
      if (tcount[tail[j]]++ == 0) tnodes.push_back(tail[j]);
      if (hcount[head[j]]++ == 0) hnodes.push_back(head[j]);
    }
    // tail lists: decreasing length, then ascending node
    std::sort(tnodes.begin(), tnodes.end(), [&](uint32_t x, uint32_t y) {
      return tcount[x] != tcount[y] ? tcount[x] > tcount[y] : x < y;
    });
    std::sort(hnodes.begin(), hnodes.end());
    std::vector<uint32_t> tlen(tnodes.size());
    for (size_t q = 0; q < tnodes.size(); ++q) {
      tlen[q] = tcount[tnodes[q]];
      trank[tnodes[q]] = (uint32_t)q;
    }
    ce.nrows = tnodes.empty() ? 0 : tlen[0];
    // jagged-diagonal row starts
    std::vector<uint32_t> rs(ce.nrows + 1, 0);
    {
      uint32_t pos = 0;
      size_t alive = tnodes.size();
      for (uint32_t e = 0; e < ce.nrows; ++e) {
        while (alive > 0 && tlen[alive - 1] <= e) --alive;
        rs[e] = pos;
        pos += (uint32_t)alive;
      }
      rs[ce.nrows] = pos;
      ce.njds = pos;
    }
    // touched lines
    for (uint32_t u : tnodes) ce.tlines.push_back(u / kLine);
    for (uint32_t u : hnodes) ce.hlines.push_back(u / kLine);
    std::sort(ce.tlines.begin(), ce.tlines.end());
    ce.tlines.erase(std::unique(ce.tlines.begin(), ce.tlines.end()), ce.tlines.end());
    ce.hlines.erase(std::unique(ce.hlines.begin(), ce.hlines.end()), ce.hlines.end());
    ce.lines.resize(ce.tlines.size() + ce.hlines.size());
    ce.lines.erase(std::set_union(ce.tlines.begin(), ce.tlines.end(), ce.hlines.begin(), ce.hlines.end(), ce.lines.begin()),
                   ce.lines.end());
    for (size_t i = 0; i < ce.lines.size(); ++i) local_of_line[ce.lines[i]] = (uint32_t)i;
    auto local = [&](uint32_t u) { return local_of_line[u / kLine] * kLine + u % kLine; };
    // Arc positions: row e of the jagged-diagonal order holds one arc of every tail list longer than e, at rowstart[e] + q.
    // WHICH arc of list q goes to row e is free, and is chosen so that the 16 arcs of an aligned block of positions (the 16
    // lanes of a half-warp in the arc rows) read their head node values from 16 different shared-memory banks whenever the
    // list still has such an arc.  Self-loops follow the lists.
    uint32_t* gi = h.gidx.data() + (size_t)c * h.Amax;
    std::vector<uint32_t> pos_of(nA);
    ce.tnode = tnodes;
    ce.arc_code.assign(h.Amax, 0xffffffffu);
    {
      std::vector<std::vector<uint32_t>> remaining(tnodes.size());
      uint32_t loops = ce.njds;
      for (uint32_t i = 0; i < nA; ++i) {
        const uint32_t j = arcs[i];
        if (tail[j] == head[j]) {
          gi[loops] = j;
          pos_of[i] = loops++;
        } else {
          remaining[trank[tail[j]]].push_back(i);
        }
      }
      uint32_t used = 0, block = 0xffffffffu;
      for (uint32_t e = 0; e < ce.nrows; ++e)
        for (uint32_t q = 0; q < rs[e + 1] - rs[e]; ++q) {
          const uint32_t pos = rs[e] + q;
          if (pos / 16 != block) {
            block = pos / 16;
            used = 0;
          }
          std::vector<uint32_t>& rem = remaining[q];
          size_t pick = 0;
          for (size_t x = 0; x < rem.size(); ++x)
            if (!(used >> (local(head[arcs[rem[x]]]) & 15u) & 1u)) {
              pick = x;
              break;
            }
          const uint32_t i = rem[pick];
          rem.erase(rem.begin() + (long)pick);
          const uint32_t j = arcs[i], lh = local(head[j]);
          used |= 1u << (lh & 15u);
          gi[pos] = j;
          pos_of[i] = pos;
          ce.arc_code[pos] = q | (tail[j] < head[j] ? 0x8000u : 0u) | (lh << 16);
        }
    }
    // the lists: tail list q = positions rowstart[e] + q, head lists = positions of the in-arcs, ascending arc index
    ce.lists.resize(tnodes.size() + hnodes.size());
    for (size_t q = 0; q < tnodes.size(); ++q) {
      List& li = ce.lists[q];
      li.node = tnodes[q];
      li.head = false;
      for (uint32_t e = 0; e < tlen[q]; ++e) li.pos.push_back((uint16_t)(rs[e] + q));
    }
    std::sort(hnodes.begin(), hnodes.end(), [&](uint32_t x, uint32_t y) {
      return hcount[x] != hcount[y] ? hcount[x] > hcount[y] : x < y;
    });
    for (size_t q = 0; q < hnodes.size(); ++q) {
      List& li = ce.lists[tnodes.size() + q];
      li.node = hnodes[q];
      li.head = true;
      hfill[hnodes[q]] = (uint32_t)(tnodes.size() + q);
    }
    for (uint32_t i = 0; i < nA; ++i) {
      const uint32_t j = arcs[i];
      if (tail[j] != head[j]) ce.lists[hfill[head[j]]].pos.push_back((uint16_t)pos_of[i]);
    }
    if (ce.lines.size() * kLine > 0x7fffu || tnodes.size() > 0x3fffu) return;
    h.max_tl = std::max<uint32_t>(h.max_tl, (uint32_t)tnodes.size());
    // reset the scratch arrays for the next cell
    for (uint32_t l : ce.lines) local_of_line[l] = 0xffffffffu;
    for (uint32_t u : tnodes) tcount[u] = 0;
    for (uint32_t u : hnodes) hcount[u] = 0;
  }
  h.max_own = std::max<uint32_t>(1, (h.L + Gc - 1) / Gc);

  // inbox slots: per line, contributions in the order (cell, tail side before head side)
  std::vector<uint32_t> kcount(h.L, 0);
  std::vector<std::vector<uint32_t>> tslot(Gc), hslot(Gc);
  for (uint32_t c = 0; c < Gc; ++c) {
    for (uint32_t l : cells[c].tlines) tslot[c].push_back(kcount[l]++);
    for (uint32_t l : cells[c].hlines) hslot[c].push_back(kcount[l]++);
  }
  h.slot_base.assign(h.L + 1, 0);
  for (uint32_t l = 0; l < h.L; ++l) h.slot_base[l + 1] = h.slot_base[l] + kcount[l];
  const uint64_t atoms = (uint64_t)h.slot_base[h.L] * kLine;
  if (atoms > 0x7fffffffull) return;
  h.inbox_atoms = (uint32_t)std::max<uint64_t>(atoms, 1);

  // per-cell tables: node-sum groups of 16 lists (one warp: lane l < 16 takes the even entries of list l, lane l + 16 the
  // odd ones), entry rows of four 16-bit positions per lane, padded with the zero slot at position Amax
  struct Tables {
    std::vector<uint4> walk;
    std::vector<uint2> ent4;
    std::vector<uint32_t> push;
  };
  std::vector<Tables> tabs(Gc);
  for (uint32_t c = 0; c < Gc; ++c) {
    const Cell& ce = cells[c];
    Tables& tb = tabs[c];
    const size_t ntl = ce.tlines.size();
    for (size_t i = 0; i < ntl; ++i) tb.push.push_back(h.slot_base[ce.tlines[i]] + tslot[c][i]);
    for (size_t i = 0; i < ce.hlines.size(); ++i) tb.push.push_back(h.slot_base[ce.hlines[i]] + hslot[c][i]);
    const size_t nlists = ce.lists.size(), ngroups = (nlists + 15) / 16;
    uint32_t row0 = 0;
    for (size_t g = 0; g < ngroups; ++g) {
      size_t longest = 0;
      for (size_t q = g * 16; q < std::min(nlists, g * 16 + 16); ++q) longest = std::max(longest, ce.lists[q].pos.size());
      const uint32_t rows = (uint32_t)(((longest + 1) / 2 + 3) / 4);
      tb.ent4.resize(tb.ent4.size() + (size_t)rows * 32, make_uint2(h.Amax | (h.Amax << 16), h.Amax | (h.Amax << 16)));
      for (uint32_t lane = 0; lane < 32; ++lane) {
        const size_t q = g * 16 + (lane & 15);
        uint4 d = make_uint4(0xffffffffu, 0u, row0, rows);
        if (q < nlists) {
          const List& li = ce.lists[q];
          const uint32_t line = li.node / kLine;
          const auto& lv = li.head ? ce.hlines : ce.tlines;
          const size_t slot = (size_t)(std::lower_bound(lv.begin(), lv.end(), line) - lv.begin()) + (li.head ? ntl : 0);
          d.x = (uint32_t)(slot * kLine + li.node % kLine);
          d.y = li.head ? 0x80000000u : 0u;
        }
        tb.walk.push_back(d);
      }
      // entries: lane l < 16 takes entries 0, 2, 4, ... of list l, lane l + 16 entries 1, 3, 5, ...  A tail list keeps its row
      // order (the 16 lists of a group then read consecutive words); the entries of a HEAD list may be taken in any fixed
      // order, and are ordered so that the 16 lanes of a half-warp hit different banks whenever possible.
      {
        uint16_t* e16 = reinterpret_cast<uint16_t*>(tb.ent4.data() + (size_t)row0 * 32);
        std::vector<std::vector<uint16_t>> rem(16);
        for (uint32_t l = 0; l < 16; ++l)
          if (g * 16 + l < nlists) rem[l] = ce.lists[g * 16 + l].pos;
        for (uint32_t n = 0; n < rows * 4; ++n)
          for (uint32_t sub = 0; sub < 2; ++sub) {
            uint32_t used = 0;
            for (uint32_t l = 0; l < 16; ++l) {
              if (rem[l].empty()) continue;
              size_t pick = 0;
              if (ce.lists[g * 16 + l].head)
                for (size_t x = 0; x < rem[l].size(); ++x)
                  if (!(used >> (rem[l][x] & 15u) & 1u)) {
                    pick = x;
                    break;
                  }
              const uint16_t pos = rem[l][pick];
              rem[l].erase(rem[l].begin() + (long)pick);
              used |= 1u << (pos & 15u);
              e16[((size_t)(n / 4) * 32 + sub * 16 + l) * 4 + n % 4] = pos;
            }
          }
      }
      row0 += rows;
    }
    h.max_lines = std::max<uint32_t>(h.max_lines, (uint32_t)ce.lines.size());
    h.max_slots = std::max<uint32_t>(h.max_slots, (uint32_t)tb.push.size());
    h.max_groups = std::max<uint32_t>(h.max_groups, (uint32_t)ngroups);
    h.max_rows = std::max<uint32_t>(h.max_rows, row0);
  }
  h.max_lines = std::max<uint32_t>(h.max_lines, 1);
  h.max_slots = std::max<uint32_t>(h.max_slots, 1);
  h.max_groups = std::max<uint32_t>(h.max_groups, 1);
  h.max_rows = std::max<uint32_t>(h.max_rows, 1);
  h.max_tl = std::max<uint32_t>(h.max_tl, 1);
  // node values live in one shared-memory array: [local node (8 per touched line) | tail value by rank | zero slot]
  const uint32_t NB = kLine * h.max_lines, ZERO = NB + h.max_tl;
  if (ZERO > 0x7fffu) return;
  h.tmap.assign((size_t)Gc * NB, 0xffffu);
  for (uint32_t c = 0; c < Gc; ++c) {
    const Cell& ce = cells[c];
    uint32_t* lt = h.lth.data() + (size_t)c * h.Amax;
    for (uint32_t i = 0; i < h.Amax; ++i) {
      const uint32_t code = ce.arc_code.empty() ? 0xffffffffu : ce.arc_code[i];
      lt[i] = code == 0xffffffffu ? (ZERO | (ZERO << 16)) : ((NB + (code & 0x3fffu)) | (code & 0x8000u) | (code & 0xffff0000u));
    }
    for (size_t q = 0; q < ce.tnode.size(); ++q) {
      const uint32_t u = ce.tnode[q];
      const size_t li = (size_t)(std::lower_bound(ce.lines.begin(), ce.lines.end(), u / kLine) - ce.lines.begin());
      h.tmap[(size_t)c * NB + li * kLine + u % kLine] = (uint16_t)(NB + q);
    }
  }
  h.lines.assign((size_t)Gc * h.max_lines, 0);
  h.push.assign((size_t)Gc * h.max_slots, 0);
  h.walk.assign((size_t)Gc * h.max_groups * 32, make_uint4(0xffffffffu, 0, 0, 0));
  h.ent4.assign((size_t)Gc * h.max_rows * 32, make_uint2(0, 0));
  for (uint32_t c = 0; c < Gc; ++c) {
    const Cell& ce = cells[c];
    const Tables& tb = tabs[c];
    std::copy(ce.lines.begin(), ce.lines.end(), h.lines.begin() + (size_t)c * h.max_lines);
    std::copy(tb.push.begin(), tb.push.end(), h.push.begin() + (size_t)c * h.max_slots);
    std::copy(tb.walk.begin(), tb.walk.end(), h.walk.begin() + (size_t)c * h.max_groups * 32);
    std::copy(tb.ent4.begin(), tb.ent4.end(), h.ent4.begin() + (size_t)c * h.max_rows * 32);
    uint32_t* hd = h.hdr.data() + (size_t)c * 8;
    hd[0] = ce.nA;
    hd[1] = (uint32_t)ce.lines.size();
    hd[2] = (uint32_t)tb.push.size();
    hd[3] = (uint32_t)(tb.walk.size() / 32);
    hd[4] = (uint32_t)(tb.ent4.size() / 32);
  }
  // shared-memory fit (pass 2 is the larger layout)
  if (cell_smem_bytes(cells_probe(h), true) + 1024 > smem_limit) return;
  h.ok = true;
}

// Bank-conflict statistics of a partition (diagnostic): average number of shared-memory wavefronts per half-warp access
// (1.0 = conflict-free), x1000, of (a) the head node values gathered by the arc rows and (b) the arc values gathered by
// the node sums.
inline void cell_conflicts(const HostCells& h, uint32_t& arc_rows_x1000, uint32_t& node_sums_x1000) {
  uint64_t wa = 0, na = 0, wb = 0, nb = 0;
  auto wavefronts = [](const uint32_t* addr, int n) {
    int cnt[16] = {0}, mx = 0;
    for (int i = 0; i < n; ++i) {
      bool dup = false;
      for (int k = 0; k < i; ++k) dup |= addr[k] == addr[i];  // same word: broadcast
      if (!dup) mx = std::max(mx, ++cnt[addr[i] & 15u]);
    }
    return mx;
  };
  for (uint32_t c = 0; c < h.Gc; ++c) {
    const uint32_t* hd = h.hdr.data() + (size_t)c * 8;
    const uint32_t* lt = h.lth.data() + (size_t)c * h.Amax;
    for (uint32_t b0 = 0; b0 < hd[0]; b0 += 16) {
      uint32_t a[16];
      int n = 0;
      for (uint32_t i = b0; i < std::min(hd[0], b0 + 16); ++i) a[n++] = lt[i] >> 16;
      wa += wavefronts(a, n);
      ++na;
    }
    const uint4* wk = h.walk.data() + (size_t)c * h.max_groups * 32;
    const uint2* e4 = h.ent4.data() + (size_t)c * h.max_rows * 32;
    for (uint32_t g = 0; g < hd[3]; ++g)
      for (uint32_t k = 0; k < wk[g * 32].w; ++k)
        for (uint32_t f = 0; f < 4; ++f)
          for (uint32_t half = 0; half < 2; ++half) {
            uint32_t a[16];
            for (uint32_t l = 0; l < 16; ++l) {
              const uint2 e = e4[(size_t)(wk[g * 32].z + k) * 32 + half * 16 + l];
              const uint32_t w = f < 2 ? e.x : e.y;
              a[l] = (f & 1) ? w >> 16 : w & 0xffffu;
            }
            wb += wavefronts(a, 16);
            ++nb;
          }
  }
  arc_rows_x1000 = na ? (uint32_t)(wa * 1000 / na) : 0;
  node_sums_x1000 = nb ? (uint32_t)(wb * 1000 / nb) : 0;
}

// Host-side consistency check of a partition (diagnostic, also run by the CPU test-suite): every arc sits in exactly
// one slot and decodes to its own tail / head, and the node sums formed through the group / entry / push tables from an
// integer-valued arc vector equal E w exactly.  Returns 0 when consistent, else a code naming the broken invariant.
inline int check_cells(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const HostCells& h) {
  std::vector<uint8_t> seen(m, 0);
  std::vector<double> w(m), inbox(h.inbox_atoms, 0.0), direct(p, 0.0);
  std::vector<uint8_t> written(h.inbox_atoms, 0);
  for (size_t j = 0; j < m; ++j) {
    w[j] = (double)((j * 2654435761u) % 2001) - 1000.0;
    if (tail[j] != head[j]) {
      direct[tail[j]] += w[j];
      direct[head[j]] -= w[j];
    }
  }
  for (uint32_t c = 0; c < h.Gc; ++c) {
    const uint32_t* hd = h.hdr.data() + (size_t)c * 8;
    const uint32_t nA = hd[0], nlines = hd[1], nslots = hd[2], ngroups = hd[3], nrows = hd[4];
    if (nA > h.Amax || nlines > h.max_lines || nslots > h.max_slots || ngroups > h.max_groups || nrows > h.max_rows) return 1;
    const uint32_t* gi = h.gidx.data() + (size_t)c * h.Amax;
    const uint32_t* lt = h.lth.data() + (size_t)c * h.Amax;
    const uint32_t* ln = h.lines.data() + (size_t)c * h.max_lines;
    for (uint32_t i = 1; i < nlines; ++i)
      if (ln[i] <= ln[i - 1]) return 2;
    std::vector<double> wl(h.Amax + 8, 0.0);
    for (uint32_t i = 0; i < nA; ++i) {
      const uint32_t j = gi[i];
      if (j >= m || seen[j]) return 3;
      seen[j] = 1;
      wl[i] = w[j];
      const uint32_t NB = kLine * h.max_lines, ZERO = NB + h.max_tl;
      const uint32_t t = lt[i] & 0x7fffu, hh = lt[i] >> 16;
      if (tail[j] == head[j]) {
        if (t != ZERO || hh != ZERO) return 4;
        continue;
      }
      if (t < NB || t >= ZERO || hh / kLine >= nlines) return 5;
      const uint16_t* tm = h.tmap.data() + (size_t)c * NB;
      uint32_t tl = 0xffffffffu;  // the local node whose value is mirrored into tail slot t
      for (uint32_t a = 0; a < nlines * kLine; ++a)
        if (tm[a] == t) tl = a;
      if (tl == 0xffffffffu) return 15;
      if (ln[tl / kLine] * kLine + tl % kLine != tail[j] || ln[hh / kLine] * kLine + hh % kLine != head[j]) return 6;
      if (((lt[i] >> 15) & 1u) != (tail[j] < head[j] ? 1u : 0u)) return 16;
    }
    std::vector<double> sums((size_t)nslots * kLine, 0.0);
    std::vector<uint8_t> sset(sums.size(), 0);
    const uint4* wk = h.walk.data() + (size_t)c * h.max_groups * 32;
    const uint2* e4 = h.ent4.data() + (size_t)c * h.max_rows * 32;
    for (uint32_t g = 0; g < ngroups; ++g) {
      double part[32];
      for (uint32_t lane = 0; lane < 32; ++lane) {
        const uint4 d = wk[g * 32 + lane];
        if (d.z + d.w > nrows) return 7;
        double acc = 0.0;
        for (uint32_t k = 0; k < d.w; ++k) {
          const uint2 e = e4[(size_t)(d.z + k) * 32 + lane];
          const uint32_t pos[4] = {e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16};
          for (uint32_t q = 0; q < 4; ++q) {
            if (pos[q] > h.Amax || (pos[q] >= nA && pos[q] != h.Amax)) return 8;
            acc += d.y ? -wl[pos[q]] : wl[pos[q]];
          }
        }
        part[lane] = acc;
      }
      for (uint32_t lane = 0; lane < 16; ++lane) {
        const uint4 d = wk[g * 32 + lane];
        if (d.x == 0xffffffffu) continue;
        if (d.x >= sums.size() || sset[d.x]) return 9;
        sset[d.x] = 1;
        sums[d.x] = part[lane] + part[lane + 16];
      }
    }
    const uint32_t* ps = h.push.data() + (size_t)c * h.max_slots;
    for (uint32_t sl = 0; sl < nslots; ++sl)
      for (uint32_t r = 0; r < kLine; ++r) {
        const size_t atom = (size_t)ps[sl] * kLine + r;
        if (atom >= h.inbox_atoms || written[atom]) return 11;
        written[atom] = 1;
        inbox[atom] = sums[(size_t)sl * kLine + r];
      }
  }
  for (size_t j = 0; j < m; ++j)
    if (!seen[j]) return 12;
  for (uint32_t l = 0; l < h.L; ++l)
    for (uint32_t r = 0; r < kLine; ++r) {
      double t = 0.0;
      for (uint32_t k = h.slot_base[l]; k < h.slot_base[l + 1]; ++k) {
        if (!written[(size_t)k * kLine + r]) return 13;
        t += inbox[(size_t)k * kLine + r];
      }
      const size_t u = (size_t)l * kLine + r;
      if (u < p ? t != direct[u] : t != 0.0) return 14;
    }
  return 0;
}

}  // namespace tpl
