// tpl_build.cuh -- DEVICE construction of the operator tables of large KKT instances: the blocked streaming layout of
// tpl_blocks.cuh and the node -> arc lists of the generic kernels, from tail / head / d already in HBM.
//
// The host builders (tpl_blocks_host.h, tpl_engine.cu) stay: they build the small instances, and they are the CHECKERS of
// what the device builds -- tpl_op_layout_check downloads the device tables and runs check_blocks over them, and the GPU
// tests compare their hash with the host-built layout (the two are identical word for word).
// Reference: src/utils/data_loader.rs:211-259 assembles the KKT matrix from triplets on the host; the operator tables here
// play that role for the B200 kernels.
//
// Steps (build_blocks_device), mirroring build_blocks:
//   1. out- / in-degrees by warp-aggregated atomics; the node blocks (O(p)) are cut on the host from the degree arrays;
//   2. arcs sorted by (cell, tail, arc index): one stable radix sort (cub) of the arc indices on the key cell << 32 | tail;
//   3. cell order: d, th, gidx scattered to the padded cell positions;
//   4. tile lists: one CTA per tile -- same-tail runs per 128-arc stage, entries keyed node << 13 | order, sorted by a
//      bitonic network in shared memory, cut into 256 slices, depth / slot fields in closed form (see tile_lists_kernel);
//      a first, counting launch gives every tile's list length, from which the host picks the tile size and the offsets.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include <cstddef>
#include <cstdint>
#include <vector>

#include "tpl_blocks_host.h"

namespace tpl {

struct DeviceBlocks {
  HostBlocks meta;  // ok, grid, PT / PH, T, ntile, rings, lblk, nl, ntb and the SMALL tables (cell_off, tbs, hbs, tbn, hbn)
  // the large tables, cudaMalloc'ed here; the caller owns them afterwards (nullptr when !meta.ok)
  double* d = nullptr;
  uint32_t* th = nullptr;
  uint32_t* gidx = nullptr;
  uint4* pdesc = nullptr;
  uint4* thdr = nullptr;
  uint32_t* lent = nullptr;
  size_t lent_words = 0;
  uint32_t max_cell = 0;
};

namespace devbuild {

constexpr int kThreads = 256;
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kSeqBits = 13;  // order of an entry inside its tile: tail side i, head side T + i (T <= 4096)

#define TPL_BUILD_TRY(expr)              \
  do {                                   \
    const cudaError_t e__ = (expr);      \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

// histogram of key[j] (j < m) with one atomic per distinct key per warp
__device__ __forceinline__ void warp_count(uint32_t* hist, uint32_t key, bool valid) {
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const unsigned same = __match_any_sync(act, key);
  if ((int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(hist + key, (uint32_t)__popc(same));
}
__global__ void degrees_kernel(size_t m, const uint32_t* tail, const uint32_t* head, uint32_t* outdeg, uint32_t* indeg, uint32_t* loops) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t rounds = (m + stride - 1) / stride;
  for (size_t r = 0; r < rounds; ++r) {
    const size_t j = r * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = j < m;
    const uint32_t t = ok ? tail[j] : 0u, h = ok ? head[j] : 0u;
    warp_count(outdeg, t, ok);
    warp_count(indeg, h, ok);
    warp_count(loops, t, ok && t == h);
  }
}
// key = cell << 32 | tail, value = arc index
__global__ void cell_keys_kernel(size_t m, const uint32_t* tail, const uint32_t* head, const uint32_t* node_tb, const uint32_t* node_hb,
                                 uint32_t GC, uint64_t* key, uint32_t* val) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (size_t)gridDim.x * blockDim.x) {
    const uint32_t t = tail[j];
    key[j] = ((uint64_t)(node_tb[t] * GC + node_hb[head[j]]) << 32) | t;
    val[j] = (uint32_t)j;
  }
}
// first sorted position of every cell that has arcs (cell_start is pre-set to kNone)
__global__ void cell_starts_kernel(size_t m, const uint64_t* key, uint32_t* cell_start) {
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(key[q] >> 32);
    if (q == 0 || (uint32_t)(key[q - 1] >> 32) != c) cell_start[c] = (uint32_t)q;
  }
}
__global__ void fill_padding_kernel(size_t Mpad, double* d, uint32_t* th, uint32_t* gidx, uint4* pdesc) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < Mpad; i += (size_t)gridDim.x * blockDim.x) {
    d[i] = 0.0;
    th[i] = kBLoop;
    gidx[i] = kBPad;
    if (i % kBStage == 0) pdesc[i / kBStage] = make_uint4(kBNoPiece, kBNoPiece, kBNoPiece, kBNoPiece);
  }
}
__global__ void cell_order_kernel(size_t m, const uint64_t* key, const uint32_t* order, const uint32_t* tail, const uint32_t* head,
                                  const double* d_in, const uint32_t* tloc, const uint32_t* hloc, const uint32_t* cell_start,
                                  const uint32_t* cell_off, double* d, uint32_t* th, uint32_t* gidx) {
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(key[q] >> 32), j = order[q];
    const uint32_t pos = cell_off[c] + ((uint32_t)q - cell_start[c]);
    const uint32_t t = tail[j], hd = head[j];
    gidx[pos] = j;
    d[pos] = d_in[j];
    th[pos] = tloc[t] | (hloc[hd] << 15) | (t < hd ? kBTailFirst : 0u) | (t == hd ? kBLoop : 0u);
  }
}

// ---------------------------------------------------------------------------- tile lists
// One CTA (256 threads = one per slice) per tile of a cell; the device form of build_cell_lists, identical output.
//   a. the tile's local tails / heads into shared memory (loop or padding: tail == head == 0);
//   b. one thread per 128-arc stage finds the stage's same-tail runs of >= kBPieceMin arcs (the first four are candidates);
//      the candidates are numbered across the stages in order and the first block_piece_slots(T) - 1 become run sums (all of them: four per stage fit);
//   c. entries: tail side in arc order (a run = one entry at its first arc), then head side; key = node << 13 | order;
//   d. WRITE: a bitonic sort of the (key, code) pairs puts them in the host's order (by node, a node's tail side in arc order,
//      then its head side); slice i takes entries [i L, (i + 1) L); depth of a slice whose first node X continues from the
//      slice before = i - (slice that holds X's first entry) (the slices in between hold nothing but X); entry / row-0 fields
//      as in build_cell_lists.
// COUNT (WRITE = false) stops after c and reports the number of entries.
struct TileListArgs {
  const uint32_t* th;        // cell-order packed words
  const uint32_t* cell_off;  // [Gc + 1]
  uint32_t T, ntile, PT, PL;
  uint32_t NS;               // sort length: power of two >= 2 T
  uint32_t* ne;              // [Gc * ntile] entries per tile (COUNT: written, WRITE: read)
  const uint32_t* base;      // [Gc * ntile] first word of the tile's block in lent (WRITE)
  uint4* pdesc;
  uint4* thdr;
  uint32_t* lent;
  uint32_t* fail;            // set when a depth exceeds 255
};
template <bool WRITE>
__global__ void __launch_bounds__(kThreads) tile_lists_kernel(const TileListArgs a) {
  extern __shared__ uint32_t sm[];
  const uint32_t T = a.T, B = kBSlices;
  uint16_t* tl = reinterpret_cast<uint16_t*>(sm);  // [T]
  uint16_t* hl = tl + T;                           // [T]
  uint32_t* tcode = sm + T;                        // [T] code of the arc's tail-side entry, kNone: no entry
  uint32_t* cand = tcode + T;                      // [32][4] candidate runs of a stage: start | len << 8
  uint32_t* ncand = cand + 128;                    // [32], then [32] first candidate number of the stage
  uint32_t* red = ncand + 64;                      // [kThreads / 32 + 1] block reductions
  uint32_t* key = red + 16;                        // [NS] (WRITE)
  uint32_t* val = key + a.NS;                      // [NS]
  const uint32_t c = blockIdx.x / a.ntile, t = blockIdx.x % a.ntile;
  const uint32_t c0 = a.cell_off[c], n = a.cell_off[c + 1] - c0;
  const uint32_t t0 = min(n, t * T), t1 = min(n, t0 + T), na = t1 - t0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t i = tid; i < T; i += kThreads) {
    uint32_t w = kBLoop;
    if (i < na) w = a.th[c0 + t0 + i];
    const bool loop = (w & kBLoop) != 0u;
    tl[i] = loop ? 0 : (uint16_t)(w & 0x7fffu);
    hl[i] = loop ? 0 : (uint16_t)(a.PT + ((w >> 15) & 0x7fffu));
    tcode[i] = loop ? kNone : i * 8u;
  }
  __syncthreads();
  const uint32_t nstage = (na + kBStage - 1) / kBStage;
  if ((uint32_t)tid < 32u) {
    uint32_t cnt = 0;
    if ((uint32_t)tid < nstage) {
      const uint32_t s0 = tid * kBStage, s1 = min(na, s0 + kBStage);
      for (uint32_t i = s0; i < s1;) {
        if (tl[i] == hl[i]) {
          ++i;
          continue;
        }
        uint32_t j = i;
        while (j < s1 && tl[j] == tl[i] && tl[j] != hl[j]) ++j;
        if (j - i >= kBPieceMin && cnt < 4) cand[tid * 4 + cnt++] = (i - s0) | ((j - i) << 8);
        i = j;
      }
    }
    ncand[tid] = cnt;
    // exclusive scan over the (at most 32) stages
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    ncand[32 + tid] = inc - cnt;
  }
  __syncthreads();
  if ((uint32_t)tid < nstage) {
    const uint32_t s0 = tid * kBStage, first = ncand[32 + tid];
    uint32_t d4[4] = {kBNoPiece, kBNoPiece, kBNoPiece, kBNoPiece};
    for (uint32_t q = 0; q < ncand[tid]; ++q) {
      const uint32_t id = first + q;
      if (id >= block_piece_slots(T) - 1) break;
      const uint32_t start = cand[tid * 4 + q] & 0xffu, len = cand[tid * 4 + q] >> 8;
      d4[q] = start | ((len - 1) << 8) | (id << 16);
      tcode[s0 + start] = (T + id) * 8u;
      for (uint32_t i = 1; i < len; ++i) tcode[s0 + start + i] = kNone;
    }
    if (WRITE) a.pdesc[(c0 + t0) / kBStage + tid] = make_uint4(d4[0], d4[1], d4[2], d4[3]);
  }
  __syncthreads();
  const uint32_t tile_id = c * a.ntile + t;
  if (!WRITE) {
    uint32_t cnt = 0;
    for (uint32_t i = tid; i < na; i += kThreads) cnt += (tcode[i] != kNone) + (tl[i] != hl[i]);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) red[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
      uint32_t s = 0;
      for (int w = 0; w < kThreads / 32; ++w) s += red[w];
      a.ne[tile_id] = s;
    }
    return;
  }
  const uint32_t ne = a.ne[tile_id];
  for (uint32_t i = tid; i < a.NS; i += kThreads) {
    uint32_t k = kNone, v = 0;
    if (i < T) {
      if (tcode[i] != kNone) {
        k = ((uint32_t)tl[i] << kSeqBits) | i;
        v = tcode[i];
      }
    } else if (i < 2 * T) {
      const uint32_t u = i - T;
      if (tl[u] != hl[u]) {
        k = ((uint32_t)hl[u] << kSeqBits) | i;
        v = (u * 8u) | kBEntMinus;
      }
    }
    key[i] = k;
    val[i] = v;
  }
  __syncthreads();
  for (uint32_t size = 2; size <= a.NS; size <<= 1)
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t x = tid; x < a.NS / 2; x += kThreads) {
        const uint32_t lo = 2 * x - (x & (stride - 1)), hi = lo + stride;  // lo has bit `stride` clear
        const bool up = (lo & size) == 0;
        const uint32_t ka = key[lo], kb = key[hi];
        if ((ka > kb) == up) {
          key[lo] = kb;
          key[hi] = ka;
          const uint32_t va = val[lo];
          val[lo] = val[hi];
          val[hi] = va;
        }
      }
      __syncthreads();
    }
  // slices
  const uint32_t L = (ne + B - 1) / B, pad = block_pad_entry(a.PL, T), base = a.base[tile_id];
  const uint32_t i = tid;  // kThreads == kBSlices
  const uint32_t s0 = min(ne, i * L), s1 = min(ne, s0 + L);
  uint32_t depth = 0;
  if (s0 < s1 && i > 0 && s0 > 0 && (key[s0] >> kSeqBits) == (key[s0 - 1] >> kSeqBits)) {
    const uint32_t want = (key[s0] >> kSeqBits) << kSeqBits;  // first entry of the node: lower bound of node << 13
    uint32_t lo = 0, hi = s0;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (key[mid] < want)
        lo = mid + 1;
      else
        hi = mid;
    }
    depth = i - lo / L;
  }
  const uint32_t scratch = a.PL + kBAccPad + i;
  const uint32_t first_node = s0 < s1 ? key[s0] >> kSeqBits : 0u;
  auto slot_of = [&](uint32_t u) { return depth && u == first_node ? scratch : u; };
  a.lent[base + i] = depth | ((s0 < s1 ? slot_of(key[s1 - 1] >> kSeqBits) : a.PL) << 8);
  for (uint32_t q = 0; q < L; ++q) {
    const uint32_t e = s0 + q;
    uint32_t word = pad;
    if (e < s1) {
      const uint32_t node = key[e] >> kSeqBits;
      const bool first = e == s0 || node != (key[e - 1] >> kSeqBits);
      const uint32_t field = e == s0 ? node : first ? slot_of(key[e - 1] >> kSeqBits) : a.PL;
      word = val[e] | (field << kBEntNodeShift) | (first ? kBEntNew : 0u);
    }
    a.lent[base + (size_t)(q + 1) * B + i] = word;
  }
  uint32_t dmax = __reduce_max_sync(0xffffffffu, depth);
  if (lane == 0) red[warp] = dmax;
  __syncthreads();
  if (tid == 0) {
    uint32_t D = 0;
    for (int w = 0; w < kThreads / 32; ++w) D = max(D, red[w]);
    if (D > 255u) *a.fail = 1u;
    a.thdr[tile_id] = make_uint4(base, L | (D << 24), 0u, 0u);
  }
}
inline size_t tile_lists_smem(uint32_t T, uint32_t NS, bool write) {
  return ((size_t)T + T + 128 + 64 + 16 + (write ? 2 * (size_t)NS : 0)) * sizeof(uint32_t);
}

// node -> arc lists: pairs (tail_j, j), (head_j, j | sign) in arc order; loops get the sentinel key p and sort to the end
__global__ void node_pairs_kernel(size_t m, uint32_t p, const uint32_t* tail, const uint32_t* head, uint32_t* key, uint32_t* val) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (size_t)gridDim.x * blockDim.x) {
    const uint32_t t = tail[j], h = head[j];
    const bool loop = t == h;
    key[2 * j] = loop ? p : t;
    key[2 * j + 1] = loop ? p : h;
    val[2 * j] = (uint32_t)j;
    val[2 * j + 1] = (uint32_t)j | kSignBit;
  }
}

inline int bits_for(uint64_t max_value) {  // bits needed to hold values 0 .. max_value
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}
struct Scratch {  // device temporaries, freed on every exit
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* q : ptrs) cudaFree(q);
  }
  template <class T>
  cudaError_t get(T** out, size_t count) {
    void* q = nullptr;
    const cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(q);
    *out = static_cast<T*>(q);
    return e;
  }
  void drop(void* q) {
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == q) {
        cudaFree(q);
        ptrs.erase(ptrs.begin() + i);
        return;
      }
  }
  void keep(void* q) {  // ownership leaves the scratch
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == q) {
        ptrs.erase(ptrs.begin() + i);
        return;
      }
  }
};

}  // namespace devbuild

inline void free_device_blocks(DeviceBlocks& b) {
  for (void* q : {(void*)b.d, (void*)b.th, (void*)b.gidx, (void*)b.pdesc, (void*)b.thdr, (void*)b.lent})
    if (q) cudaFree(q);
  b = DeviceBlocks{};
}

// tail / head: [m] node ids < p (validated by the caller), d: [m] (zero-padded by the caller), all in device memory.
// Returns a cudaError_t (0 = no CUDA error); out.meta.ok says whether the layout exists (it does not when the node blocks
// or the lists do not fit the shared memory / field widths -- exactly the cases in which build_blocks gives up).
// `work`: four device buffers of at least 8 m bytes each that the builder may overwrite (the handle's vector workspace: freeing
// gigabytes of temporaries costs more than the whole construction).
inline int build_blocks_device(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, int G,
                               size_t smem_limit, cudaStream_t stream, void* const work[4], DeviceBlocks& out) {
  using namespace devbuild;
  out = DeviceBlocks{};
  HostBlocks& h = out.meta;
  if (m == 0 || p == 0 || G < 1 || m >= 0xfffff000ull) return 0;
  const int grid = 148 * 8;
  Scratch sc;
  h.GR = std::max<uint32_t>(1, (uint32_t)std::floor(std::sqrt((double)G)));
  h.GC = std::max<uint32_t>(1, (uint32_t)G / h.GR);
  const uint32_t Gc = h.GR * h.GC;
  // 1. degrees -> node blocks (host, O(p))
  uint32_t* deg = nullptr;  // outdeg | indeg | loops
  TPL_BUILD_TRY(sc.get(&deg, 3 * p));
  TPL_BUILD_TRY(cudaMemsetAsync(deg, 0, 3 * p * sizeof(uint32_t), stream));
  degrees_kernel<<<grid, 256, 0, stream>>>(m, tail, head, deg, deg + p, deg + 2 * p);
  std::vector<uint32_t> hdeg(2 * p);
  TPL_BUILD_TRY(cudaMemcpyAsync(hdeg.data(), deg, 2 * p * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  TPL_BUILD_TRY(cudaStreamSynchronize(stream));
  std::vector<uint64_t> wt_t, wt_h;
  std::vector<uint32_t> tpos(p, 0), hpos(p, 0);
  for (size_t u = 0; u < p; ++u) {
    if (hdeg[u]) {
      tpos[u] = (uint32_t)h.tbn.size();
      h.tbn.push_back((uint32_t)u);
      wt_t.push_back(hdeg[u]);
    }
    if (hdeg[p + u]) {
      hpos[u] = (uint32_t)h.hbn.size();
      h.hbn.push_back((uint32_t)u);
      wt_h.push_back(hdeg[p + u]);
    }
  }
  h.tbs = weight_blocks(wt_t, h.GR);
  h.hbs = weight_blocks(wt_h, h.GC);
  uint32_t PT = 0, PH = 0;
  for (uint32_t a = 0; a < h.GR; ++a) PT = std::max(PT, h.tbs[a + 1] - h.tbs[a]);
  for (uint32_t b = 0; b < h.GC; ++b) PH = std::max(PH, h.hbs[b + 1] - h.hbs[b]);
  PT = std::max(2u, (PT + 1u) & ~1u);
  PH = std::max(2u, (PH + 1u) & ~1u);
  if (PT > 0x8000u || PH > 0x8000u || PT + PH + kBAccPad + kFoldThreads > kBMaxLocalNodes + 1) return 0;
  h.PT = PT;
  h.PH = PH;
  // per node: block and local id on either side (0 for a node that is not active on the side: never read)
  std::vector<uint32_t> nodeinfo(4 * p, 0);  // node_tb | node_hb | tloc | hloc
  for (uint32_t a = 0; a < h.GR; ++a)
    for (uint32_t q = h.tbs[a]; q < h.tbs[a + 1]; ++q) {
      nodeinfo[h.tbn[q]] = a;
      nodeinfo[2 * p + h.tbn[q]] = q - h.tbs[a];
    }
  for (uint32_t b = 0; b < h.GC; ++b)
    for (uint32_t q = h.hbs[b]; q < h.hbs[b + 1]; ++q) {
      nodeinfo[p + h.hbn[q]] = b;
      nodeinfo[3 * p + h.hbn[q]] = q - h.hbs[b];
    }
  uint32_t* ninfo = nullptr;
  TPL_BUILD_TRY(sc.get(&ninfo, 4 * p));
  TPL_BUILD_TRY(cudaMemcpyAsync(ninfo, nodeinfo.data(), 4 * p * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  // 2. arcs by (cell, tail, arc index)
  uint64_t *key_a = static_cast<uint64_t*>(work[0]), *key_b = static_cast<uint64_t*>(work[1]);
  uint32_t *val_a = static_cast<uint32_t*>(work[2]), *val_b = static_cast<uint32_t*>(work[3]);
  cell_keys_kernel<<<grid, 256, 0, stream>>>(m, tail, head, ninfo, ninfo + p, h.GC, key_a, val_a);
  {
    cub::DoubleBuffer<uint64_t> kb(key_a, key_b);
    cub::DoubleBuffer<uint32_t> vb(val_a, val_b);
    const int end_bit = 32 + bits_for(Gc - 1);
    size_t tmp_bytes = 0;
    TPL_BUILD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int64_t)m, 0, end_bit, stream));
    uint8_t* tmp = nullptr;
    TPL_BUILD_TRY(sc.get(&tmp, tmp_bytes));
    // (the tail bits above log2(p) are zero: cub skips nothing, but a pass over 8 zero bits is a plain copy)
    const int tail_bits = bits_for(p - 1);
    // two sorts would be needed for begin/end bit ranges that are not contiguous; sort the tail bits, then the cell bits
    TPL_BUILD_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int64_t)m, 0, tail_bits, stream));
    TPL_BUILD_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int64_t)m, 32, end_bit, stream));
    key_a = kb.Current();
    val_a = vb.Current();
    sc.drop(tmp);
  }
  // 3. cell offsets (host, Gc words) and the cell-order arrays
  uint32_t* cstart = nullptr;
  TPL_BUILD_TRY(sc.get(&cstart, Gc + 1));
  TPL_BUILD_TRY(cudaMemsetAsync(cstart, 0xff, (Gc + 1) * sizeof(uint32_t), stream));
  cell_starts_kernel<<<grid, 256, 0, stream>>>(m, key_a, cstart);
  std::vector<uint32_t> hstart(Gc + 1);
  TPL_BUILD_TRY(cudaMemcpyAsync(hstart.data(), cstart, (Gc + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  TPL_BUILD_TRY(cudaStreamSynchronize(stream));
  hstart[Gc] = (uint32_t)m;
  for (uint32_t c = Gc; c-- > 0;)
    if (hstart[c] == kNone) hstart[c] = hstart[c + 1];  // empty cell
  h.cell_off.assign(Gc + 1, 0);
  uint64_t off = 0, max_cell = 0;
  for (uint32_t c = 0; c < Gc; ++c) {
    const uint64_t cnt = hstart[c + 1] - hstart[c];
    h.cell_off[c] = (uint32_t)off;
    max_cell = std::max(max_cell, cnt);
    off += (cnt + kBStage - 1) / kBStage * kBStage;
    if (off >= 0xfffff000ull) return 0;
  }
  h.cell_off[Gc] = (uint32_t)off;
  h.Mpad = (uint32_t)off;
  uint32_t* coff = nullptr;
  TPL_BUILD_TRY(sc.get(&coff, Gc + 1));
  TPL_BUILD_TRY(cudaMemcpyAsync(coff, h.cell_off.data(), (Gc + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  TPL_BUILD_TRY(cudaMemcpyAsync(cstart, hstart.data(), (Gc + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  TPL_BUILD_TRY(sc.get(&out.d, h.Mpad));
  TPL_BUILD_TRY(sc.get(&out.th, h.Mpad));
  TPL_BUILD_TRY(sc.get(&out.gidx, h.Mpad));
  TPL_BUILD_TRY(sc.get(&out.pdesc, h.Mpad / kBStage));
  fill_padding_kernel<<<grid, 256, 0, stream>>>(h.Mpad, out.d, out.th, out.gidx, out.pdesc);
  cell_order_kernel<<<grid, 256, 0, stream>>>(m, key_a, val_a, tail, head, d, ninfo + 2 * p, ninfo + 3 * p, cstart, coff, out.d, out.th, out.gidx);
  TPL_BUILD_TRY(cudaStreamSynchronize(stream));
  // 4. tile size (the rule of build_blocks) and the lists
  const uint32_t padded_max = (uint32_t)((max_cell + kBStage - 1) / kBStage * kBStage);
  uint32_t want = std::min<uint32_t>(4096, std::max<uint32_t>(1024, (padded_max + 1023) / 1024 * 1024));
  if (const char* e = std::getenv("TPL_BLOCK_T")) want = std::min<uint32_t>(4096, std::max<uint32_t>(1024, (uint32_t)std::atoi(e) / 1024 * 1024));
  if (const char* e = std::getenv("TPL_BLOCK_NTB")) h.ntb = std::min<uint32_t>(kBMaxTileBufs, std::max<uint32_t>(2, (uint32_t)std::atoi(e)));
  uint32_t *ne_dev = nullptr, *fail_dev = nullptr;
  std::vector<uint32_t> ne_host;
  TileListArgs ta{};
  ta.th = out.th;
  ta.cell_off = coff;
  ta.PT = PT;
  ta.PL = PT + PH;
  ta.pdesc = out.pdesc;
  TPL_BUILD_TRY(sc.get(&fail_dev, 1));
  TPL_BUILD_TRY(cudaMemsetAsync(fail_dev, 0, sizeof(uint32_t), stream));
  ta.fail = fail_dev;
  bool done = false;
  for (int need = 3; need >= 2 && !done; --need)
    for (uint32_t T = want; T >= 1024 && !done; T -= 1024) {
      const uint32_t ntile = std::max<uint32_t>(1, (padded_max + T - 1) / T);
      if (!blocks_fit(PT + PH, smem_limit, T, (std::min(T, padded_max) / kFoldThreads + 1) * 4u * kFoldThreads, h.ntb, ntile, need, h.ring1, h.ring2, h.ring2v))
        continue;
      h.T = T;
      h.ntile = ntile;
      const size_t ntiles = (size_t)Gc * ntile;
      if (ne_dev) sc.drop(ne_dev);
      TPL_BUILD_TRY(sc.get(&ne_dev, ntiles));
      ta.T = T;
      ta.ntile = ntile;
      ta.NS = 2048;
      while (ta.NS < 2 * T) ta.NS <<= 1;
      ta.ne = ne_dev;
      tile_lists_kernel<false><<<(unsigned)ntiles, kThreads, tile_lists_smem(T, ta.NS, false), stream>>>(ta);
      ne_host.resize(ntiles);
      TPL_BUILD_TRY(cudaMemcpyAsync(ne_host.data(), ne_dev, ntiles * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      TPL_BUILD_TRY(cudaStreamSynchronize(stream));
      uint32_t Lmax = 0;
      for (uint32_t e : ne_host) Lmax = std::max(Lmax, (e + kFoldThreads - 1) / kFoldThreads);
      h.lblk = (Lmax + 1) * 4u * kFoldThreads;
      done = blocks_fit(PT + PH, smem_limit, T, h.lblk, h.ntb, h.ntile, need, h.ring1, h.ring2, h.ring2v);
    }
  if (!done) {
    for (void* q : {(void*)out.d, (void*)out.th, (void*)out.gidx, (void*)out.pdesc}) sc.drop(q);
    out.d = nullptr;
    out.th = out.gidx = nullptr;
    out.pdesc = nullptr;
    return 0;
  }
  {
    const size_t budget = smem_limit > 3072 ? smem_limit - 3072 : 0;
    h.nl = 2;
    while (h.nl < (uint32_t)kBMaxList && h.nl < h.ntile &&
           block_smem_bytes(PT + PH, h.T, (int)h.ring2v, h.lblk, h.nl + 1, h.ntb, true, true, h.ntile) <= budget &&
           block_smem_bytes(PT + PH, h.T, (int)h.ring2, h.lblk, h.nl + 1, h.ntb, true, false, h.ntile) <= budget &&
           block_smem_bytes(PT + PH, h.T, (int)h.ring1, h.lblk, h.nl + 1, h.ntb, false, false, h.ntile) <= budget)
      ++h.nl;
  }
  const size_t ntiles = (size_t)Gc * h.ntile;
  std::vector<uint32_t> base(ntiles);
  uint64_t words = 0;
  for (size_t q = 0; q < ntiles; ++q) {
    base[q] = (uint32_t)words;
    words += ((uint64_t)(ne_host[q] + kFoldThreads - 1) / kFoldThreads + 1) * kFoldThreads;
  }
  bool fits32 = words < 0xffffffffull;
  uint32_t* base_dev = nullptr;
  if (fits32) {
    TPL_BUILD_TRY(sc.get(&base_dev, ntiles));
    TPL_BUILD_TRY(cudaMemcpyAsync(base_dev, base.data(), ntiles * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    TPL_BUILD_TRY(sc.get(&out.lent, words));
    TPL_BUILD_TRY(sc.get(&out.thdr, ntiles));
    ta.base = base_dev;
    ta.lent = out.lent;
    ta.thdr = out.thdr;
    const size_t smem = tile_lists_smem(h.T, ta.NS, true);
    TPL_BUILD_TRY(cudaFuncSetAttribute(tile_lists_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tile_lists_kernel<true><<<(unsigned)ntiles, kThreads, smem, stream>>>(ta);
    uint32_t failed = 0;
    TPL_BUILD_TRY(cudaMemcpyAsync(&failed, fail_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    TPL_BUILD_TRY(cudaStreamSynchronize(stream));
    TPL_BUILD_TRY(cudaGetLastError());
    fits32 = failed == 0;
  }
  if (!fits32) {
    for (void* q : {(void*)out.d, (void*)out.th, (void*)out.gidx, (void*)out.pdesc, (void*)out.lent, (void*)out.thdr})
      if (q) sc.drop(q);
    out.d = nullptr;
    out.th = out.gidx = out.lent = nullptr;
    out.pdesc = out.thdr = nullptr;
    return 0;
  }
  out.lent_words = (size_t)words;
  out.max_cell = (uint32_t)max_cell;
  for (void* q : {(void*)out.d, (void*)out.th, (void*)out.gidx, (void*)out.pdesc, (void*)out.lent, (void*)out.thdr}) sc.keep(q);
  h.ok = true;
  return 0;
}

// Node -> arc lists: ent_idx (device, cudaMalloc'ed, caller owns) holds for node u the arcs j with tail_j == u (entry j) or
// head_j == u (entry j | kSignBit), ascending j, self-loops left out; row_ent[u] .. row_ent[u + 1] is node u's range.
inline int build_node_lists_device(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, cudaStream_t stream,
                                   void* const work[4], uint32_t** ent_idx, std::vector<uint64_t>& row_ent) {
  using namespace devbuild;
  *ent_idx = nullptr;
  const int grid = 148 * 8;
  Scratch sc;
  const bool timing = std::getenv("TPL_BUILD_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(stream);
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "node lists   %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  uint32_t* deg = nullptr;
  TPL_BUILD_TRY(sc.get(&deg, 3 * p));
  TPL_BUILD_TRY(cudaMemsetAsync(deg, 0, 3 * p * sizeof(uint32_t), stream));
  degrees_kernel<<<grid, 256, 0, stream>>>(m, tail, head, deg, deg + p, deg + 2 * p);
  std::vector<uint32_t> hdeg(3 * p);
  TPL_BUILD_TRY(cudaMemcpyAsync(hdeg.data(), deg, 3 * p * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  lap("degrees");
  uint32_t *key_a = static_cast<uint32_t*>(work[0]), *key_b = static_cast<uint32_t*>(work[1]);
  uint32_t *val_a = static_cast<uint32_t*>(work[2]), *val_b = static_cast<uint32_t*>(work[3]);
  lap("allocate");
  node_pairs_kernel<<<grid, 256, 0, stream>>>(m, (uint32_t)p, tail, head, key_a, val_a);
  lap("pairs");
  cub::DoubleBuffer<uint32_t> kb(key_a, key_b), vb(val_a, val_b);
  size_t tmp_bytes = 0;
  const int end_bit = bits_for(p);
  TPL_BUILD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int64_t)(2 * m), 0, end_bit, stream));
  uint8_t* tmp = nullptr;
  TPL_BUILD_TRY(sc.get(&tmp, tmp_bytes));
  TPL_BUILD_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int64_t)(2 * m), 0, end_bit, stream));
  TPL_BUILD_TRY(cudaStreamSynchronize(stream));
  lap("sort");
  row_ent.assign(p + 1, 0);
  for (size_t u = 0; u < p; ++u) row_ent[u + 1] = row_ent[u] + hdeg[u] + hdeg[p + u] - 2ull * hdeg[2 * p + u];
  // the sorted values are the lists, back to back (loops behind them)
  uint32_t* lists = nullptr;
  TPL_BUILD_TRY(sc.get(&lists, (size_t)row_ent[p]));
  TPL_BUILD_TRY(cudaMemcpyAsync(lists, vb.Current(), (size_t)row_ent[p] * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  TPL_BUILD_TRY(cudaStreamSynchronize(stream));
  lap("copy");
  sc.keep(lists);
  *ent_idx = lists;
  return 0;
}

}  // namespace tpl
#undef TPL_BUILD_TRY
