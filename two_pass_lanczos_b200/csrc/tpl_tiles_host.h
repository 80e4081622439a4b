// tpl_tiles_host.h -- host construction of the per-CTA, per-tile entry lists the tiled streaming kernels (tpl_tiles.cuh)
// fold their node sums from, and the checker the CPU tests run over them (tpl_tiles_plan).
//
// CTA c streams the arcs [c*A, (c+1)*A) in tiles of T.  For a tile, every non-loop arc contributes one entry to its head
// node (sign bit set) and one to its tail node; maximal runs of >= kPieceMin consecutive arcs with the same tail are summed
// first as "pieces" (<= kPieceMax long) and enter the list as one entry each.  The entries are sorted by node (stable: per
// node the tail side in arc order, then the head side), cut into kFoldThreads slices of nearly equal length at node
// boundaries -- a node's entries of one tile belong to ONE thread, so the fold needs no atomics and its order is fixed --
// and stored thread-interleaved: lent[base + q*kFoldThreads + thread].
// The CTAs are independent, so the lists are built by a pool of host threads and concatenated in CTA order: the result does
// not depend on the number of threads.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "tpl_tiles.cuh"

namespace tpl {

struct HostTiles {
  uint32_t T = 0, ntile = 0;
  std::vector<uint4> thdr;  // per tile: {first entry, entries per thread, first piece, end piece}
  std::vector<uint32_t> lent, piece;
};

struct TileScratch {
  std::vector<uint32_t> cnt, e_node, e_code, sorted_node, sorted_code, order, cut;
  std::vector<uint32_t> rem[16];
};

// lists of CTA c, offsets relative to the CTA's own lent / piece arrays
// bank_order = false keeps a thread's entries sorted by node (the blocked kernels add a node's values in a register);
// max_pieces < kMaxPieces reserves the last piece slot of a tile buffer (their padding entries read a zero from it)
inline void build_cta_tiles(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int G, uint32_t T, uint32_t ntile,
                            int c, TileScratch& w, std::vector<uint4>& thdr, std::vector<uint32_t>& lent,
                            std::vector<uint32_t>& piece, bool bank_order = true, uint32_t max_pieces = kMaxPieces) {
  const size_t A = (m + G - 1) / G;
  const int B = kFoldThreads;  // the fold warps walk the lists
  const size_t lo = std::min(m, A * (size_t)c), hi = std::min(m, lo + A);
  w.cnt.resize(p + 1);
  w.cut.resize(B + 1);
  thdr.assign(ntile, make_uint4(0, 0, 0, 0));
  lent.clear();
  piece.clear();
  lent.reserve(2 * (hi - lo) + (hi - lo) / 4);
  for (uint32_t t = 0; t < ntile; ++t) {
    const size_t t0 = std::min(hi, lo + (size_t)t * T), t1 = std::min(hi, t0 + T);
    const uint32_t q0 = (uint32_t)piece.size();
    w.e_node.clear();
    w.e_code.clear();
    uint32_t npieces = 0;
    for (size_t i = t0; i < t1;) {  // tail side: maximal runs of equal tail (self-loops contribute nothing)
      if (tail[i] == head[i]) {
        ++i;
        continue;
      }
      size_t j = i;
      while (j < t1 && tail[j] == tail[i] && tail[j] != head[j]) ++j;
      const size_t len = j - i;
      const size_t need = (len + kPieceMax - 1) / kPieceMax;
      if (len >= kPieceMin && npieces + need <= max_pieces) {
        for (size_t q = i; q < j; q += kPieceMax) {
          const uint32_t l = (uint32_t)std::min<size_t>(kPieceMax, j - q);
          piece.push_back((uint32_t)(q - t0) | ((l - 1) << 16));
          w.e_node.push_back(tail[i]);
          w.e_code.push_back(T + npieces);
          ++npieces;
        }
      } else {
        for (size_t q = i; q < j; ++q) {
          w.e_node.push_back(tail[i]);
          w.e_code.push_back((uint32_t)(q - t0));
        }
      }
      i = j;
    }
    for (size_t i = t0; i < t1; ++i)  // head side
      if (tail[i] != head[i]) {
        w.e_node.push_back(head[i]);
        w.e_code.push_back((uint32_t)(i - t0) | 0x4000u);
      }
    // stable counting sort by node: per node the tail entries come first (ascending index), then the head entries
    const size_t ne = w.e_node.size();
    std::fill(w.cnt.begin(), w.cnt.end(), 0u);
    for (size_t e = 0; e < ne; ++e) ++w.cnt[w.e_node[e] + 1];
    for (size_t u = 0; u < p; ++u) w.cnt[u + 1] += w.cnt[u];
    w.sorted_node.resize(ne);
    w.sorted_code.resize(ne);
    w.order.assign(w.cnt.begin(), w.cnt.end() - 1);
    for (size_t e = 0; e < ne; ++e) {
      const uint32_t dst = w.order[w.e_node[e]]++;
      w.sorted_node[dst] = w.e_node[e];
      w.sorted_code[dst] = w.e_code[e];
    }
    // cut into B slices of nearly equal length at node boundaries (a node never straddles two threads)
    w.cut[0] = 0;
    for (int i = 1; i <= B; ++i) {
      size_t want = std::max<size_t>(w.cut[i - 1], (ne * (size_t)i + B - 1) / B);
      while (want < ne && want > 0 && w.sorted_node[want] == w.sorted_node[want - 1]) ++want;
      w.cut[i] = (uint32_t)std::min(want, ne);
    }
    w.cut[B] = (uint32_t)ne;
    uint32_t L = 0;
    for (int i = 0; i < B; ++i) L = std::max(L, w.cut[i + 1] - w.cut[i]);
    const size_t base = lent.size();
    lent.resize(base + (size_t)L * B, kEntPad);
    // The order in which a thread folds its entries is free (any fixed order is deterministic).  It is chosen so that the
    // 16 threads of a half-warp, which execute fold step q together, read their tile values and their accumulators from
    // different shared-memory banks whenever they can: a random order costs ~3 wavefronts per 8-byte access.
    if (!bank_order) {
      for (int i = 0; i < B; ++i)
        for (uint32_t e = w.cut[i], q = 0; e < w.cut[i + 1]; ++e, ++q)
          lent[base + (size_t)q * B + i] = (w.sorted_node[e] << 15) | w.sorted_code[e];
    }
    for (int i0 = 0; bank_order && i0 < B; i0 += 16) {
      for (int l = 0; l < 16; ++l) {
        w.rem[l].clear();
        if (i0 + l < B)
          for (uint32_t e = w.cut[i0 + l]; e < w.cut[i0 + l + 1]; ++e) w.rem[l].push_back(e);
      }
      for (uint32_t q = 0; q < L; ++q) {
        uint32_t used_w[16] = {0}, used_a[16] = {0};
        for (int l = 0; l < 16 && i0 + l < B; ++l) {
          std::vector<uint32_t>& rem = w.rem[l];
          if (rem.empty()) continue;
          size_t pick = 0;
          uint32_t best = 0xffffffffu;
          for (size_t x = 0; x < rem.size(); ++x) {
            const uint32_t e = rem[x];
            const uint32_t cost = used_w[(w.sorted_code[e] & 0x3fffu) & 15u] + used_a[w.sorted_node[e] & 15u];
            if (cost < best) {
              best = cost;
              pick = x;
              if (!cost) break;
            }
          }
          const uint32_t e = rem[pick];
          rem.erase(rem.begin() + (long)pick);
          ++used_w[(w.sorted_code[e] & 0x3fffu) & 15u];
          ++used_a[w.sorted_node[e] & 15u];
          lent[base + (size_t)q * B + i0 + l] = (w.sorted_node[e] << 15) | w.sorted_code[e];
        }
      }
    }
    thdr[t] = make_uint4((uint32_t)base, L, q0, (uint32_t)piece.size());
  }
}

inline void build_tiles(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int G, uint32_t T, HostTiles& h,
                        int threads = 0) {
  const size_t A = (m + G - 1) / G;
  h.T = T;
  h.ntile = (uint32_t)std::max<size_t>(1, (A + T - 1) / T);
  if (threads <= 0) threads = m < (1u << 20) ? 1 : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
  threads = std::min(threads, G);
  std::vector<std::vector<uint4>> thdr(G);
  std::vector<std::vector<uint32_t>> lent(G), piece(G);
  auto run = [&](int first) {
    TileScratch w;
    for (int c = first; c < G; c += threads) build_cta_tiles(m, p, tail, head, G, T, h.ntile, c, w, thdr[c], lent[c], piece[c]);
  };
  if (threads == 1) {
    run(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(run, t);
    for (std::thread& t : pool) t.join();
  }
  size_t nl = 0, np = 0;
  for (int c = 0; c < G; ++c) {
    nl += lent[c].size();
    np += piece[c].size();
  }
  h.thdr.assign((size_t)G * h.ntile, make_uint4(0, 0, 0, 0));
  h.lent.clear();
  h.piece.clear();
  if (nl >= 0xffffffffull) {  // entry offsets are 32-bit: the caller falls back to the gather kernels
    h.T = 0;
    h.thdr.clear();
    return;
  }
  h.lent.reserve(nl);
  h.piece.reserve(np);
  for (int c = 0; c < G; ++c) {
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
}

// 0 when the lists hold, for every tile, each non-loop arc exactly once on its head node (sign set) and exactly once on its
// tail node (directly or inside one piece), every node of a tile in one thread's slice only, and padding nowhere else.
inline int check_tiles(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int G, const HostTiles& h) {
  const size_t A = (m + G - 1) / G;
  const uint32_t T = h.T, B = kFoldThreads;
  if (h.thdr.size() != (size_t)G * h.ntile) return 1;
  std::vector<uint8_t> seen_t(T), seen_h(T);
  std::vector<int32_t> owner(p);
  for (int c = 0; c < G; ++c) {
    const size_t lo = std::min(m, A * (size_t)c), hi = std::min(m, lo + A);
    for (uint32_t t = 0; t < h.ntile; ++t) {
      const uint4 hd = h.thdr[(size_t)c * h.ntile + t];
      const size_t t0 = std::min(hi, lo + (size_t)t * T), t1 = std::min(hi, t0 + T);
      const uint32_t na = (uint32_t)(t1 - t0);
      if ((size_t)hd.x + (size_t)hd.y * B > h.lent.size() || hd.z > hd.w || hd.w > h.piece.size()) return 2;
      if (hd.w - hd.z > kMaxPieces) return 3;
      std::fill(seen_t.begin(), seen_t.end(), 0);
      std::fill(seen_h.begin(), seen_h.end(), 0);
      std::fill(owner.begin(), owner.end(), -1);
      for (uint32_t q = 0; q < hd.y; ++q)
        for (uint32_t i = 0; i < B; ++i) {
          const uint32_t e = h.lent[hd.x + (size_t)q * B + i];
          if (e == kEntPad) continue;
          const uint32_t node = e >> 15, code = e & 0x7fffu;
          if (node >= p) return 4;
          if (owner[node] >= 0 && owner[node] != (int32_t)i) return 5;  // a node's entries belong to one thread
          owner[node] = (int32_t)i;
          if (code & 0x4000u) {  // head side
            const uint32_t a = code & 0x3fffu;
            if (a >= na || head[t0 + a] != node || seen_h[a]) return 6;
            seen_h[a] = 1;
          } else if (code >= T) {  // a piece
            const uint32_t q1 = hd.z + (code - T);
            if (q1 >= hd.w) return 7;
            const uint32_t start = h.piece[q1] & 0xffffu, len = (h.piece[q1] >> 16) + 1;
            if (len > kPieceMax || start + len > na) return 8;
            for (uint32_t a = start; a < start + len; ++a) {
              if (tail[t0 + a] != node || seen_t[a]) return 9;
              seen_t[a] = 1;
            }
          } else {
            if (code >= na || tail[t0 + code] != node || seen_t[code]) return 10;
            seen_t[code] = 1;
          }
        }
      for (uint32_t a = 0; a < na; ++a) {
        const bool loop = tail[t0 + a] == head[t0 + a];
        if (seen_t[a] != (loop ? 0 : 1) || seen_h[a] != (loop ? 0 : 1)) return 11;
      }
    }
  }
  return 0;
}

}  // namespace tpl
