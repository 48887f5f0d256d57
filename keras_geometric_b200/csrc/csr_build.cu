// K1/K2: device-side COO -> CSR/CSC builder (stable LSD radix sort by segment key), degrees,
// row pointers, hub table, GCN symmetric normalisation.  Integer work only - every output is
// bit-exact with numpy's argsort(kind="stable") / bincount / cumsum.
//
// Sort structure (per 8-bit digit pass, ceil(bits(n_seg-1)/8) passes):
//   upsweep   : SORT_BLOCKS CTAs, each owns a contiguous range of 4096-key tiles and histograms it;
//   scan      : two-level exclusive scan of the digit-major [256][SORT_BLOCKS] table (one CTA per digit row +
//               the 256 digit totals, prefix-summed by every downsweep CTA);
//   downsweep : each CTA re-walks its tiles in order; inside a tile keys are ranked per warp with
//               __match_any_sync (warp-striped layout keeps the original order => stable),
//               warp counts are prefix-summed per digit, the tile is REORDERED BY DIGIT IN SHARED MEMORY and
//               written out in that order, so consecutive threads store consecutive slots of a digit's run
//               (a direct scatter from registers touched up to 32 sectors per store: 1.06 -> 0.74 ms per pass).
// Degrees come from the run boundaries of the sorted keys (two integer atomics per non-empty row) instead of one
// atomicAdd per edge (61.9 M random L2 atomics = 0.9 ms on C4); rowptr is their exclusive scan.
#include "common.cuh"

namespace kgb {

constexpr int SORT_THREADS = 256;
#ifndef KGB_SORT_ITEMS
#define KGB_SORT_ITEMS 16
#endif
constexpr int SORT_ITEMS = KGB_SORT_ITEMS;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX = 256;

// keys for pass 0 + validation
__global__ void csr_prepare_kernel(const int32_t* __restrict__ ei, int64_t E, int by_source, int64_t n_seg,
                                   int64_t n_val, int64_t n_loops, uint32_t* __restrict__ keys,
                                   int32_t* __restrict__ status) {
  const int64_t M = E + n_loops;
  bool bad = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t key, val;
    if (e < E) {
      const int32_t s = __ldg(ei + e), d = __ldg(ei + E + e);
      key = by_source ? s : d;
      val = by_source ? d : s;
    } else {
      key = val = e - E;
    }
    if (key < 0 || key >= n_seg || val < 0 || val >= n_val) {
      bad = true;
      key = 0;
    }
    keys[e] = (uint32_t)key;
  }
  if (bad) atomicOr(status, KGB_STATUS_OOB_INDEX);
}

// deg[k] = length of key k's run in the SORTED key array (deg zeroed on entry): the thread at the first slot of a run
// subtracts its index, the thread at the last slot adds index + 1 - integer atomics, any order gives the same bits
__global__ void csr_degree_kernel(const uint32_t* __restrict__ keys, int64_t M, int32_t* __restrict__ deg) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = __ldg(keys + i);
    const uint32_t prev = i > 0 ? __ldg(keys + i - 1) : 0xffffffffu;      // keys are < 2^31: never equal
    const uint32_t next = i + 1 < M ? __ldg(keys + i + 1) : 0xffffffffu;
    if (prev != k) atomicSub(deg + k, (int32_t)i);
    if (next != k) atomicAdd(deg + k, (int32_t)(i + 1));
  }
}

struct BlockRange {
  int64_t tile_begin, tile_end;
};
__device__ __forceinline__ BlockRange block_range(int64_t n_tiles) {
  const int64_t per = n_tiles / gridDim.x, extra = n_tiles % gridDim.x;
  const int64_t b = blockIdx.x;
  BlockRange r;
  r.tile_begin = b * per + (b < extra ? b : extra);
  r.tile_end = r.tile_begin + per + (b < extra ? 1 : 0);
  return r;
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_upsweep_kernel(const uint32_t* __restrict__ keys, int64_t M, int shift, int32_t* __restrict__ table) {
  __shared__ int32_t hist[RADIX];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t n_tiles = (M + SORT_TILE - 1) / SORT_TILE;
  const BlockRange r = block_range(n_tiles);
  const int64_t begin = r.tile_begin * SORT_TILE;
  int64_t end = r.tile_end * SORT_TILE;
  if (end > M) end = M;
  for (int64_t i = begin + threadIdx.x; i < end; i += SORT_THREADS)
    atomicAdd(&hist[(__ldg(keys + i) >> shift) & 0xff], 1);
  __syncthreads();
  table[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];
}

// Two-level exclusive scan of the digit-major table[256][nb]: CTA d scans digit d's row (offsets[d][b] = keys with
// digit d in blocks before b) and stores the digit's total; every downsweep CTA then adds the exclusive prefix of the
// 256 totals itself.  (A single CTA walking all 303 k entries took 0.30 ms per pass - pure latency.)
constexpr int TSCAN_THREADS = 256;
__global__ void __launch_bounds__(TSCAN_THREADS)
sort_scan_kernel(const int32_t* __restrict__ table, int nb, int64_t* __restrict__ offsets,
                 int64_t* __restrict__ digit_total) {
  __shared__ int64_t warp_sums[TSCAN_THREADS / 32];
  __shared__ int64_t carry_s;
  const int32_t* row = table + (int64_t)blockIdx.x * nb;
  int64_t* out = offsets + (int64_t)blockIdx.x * nb;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int ITEMS = 8;
  for (int base = 0; base < nb; base += TSCAN_THREADS * ITEMS) {
    const int i0 = base + threadIdx.x * ITEMS;
    int32_t v[ITEMS];
    int64_t tot = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      v[k] = (i0 + k < nb) ? row[i0 + k] : 0;
      tot += v[k];
    }
    int64_t x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    int64_t wpre = 0;
#pragma unroll
    for (int w = 0; w < TSCAN_THREADS / 32; ++w) wpre += (w < wid) ? warp_sums[w] : 0;
    const int64_t carry = carry_s;
    const int64_t incl = x + wpre + carry;
    int64_t runv = incl - tot;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if (i0 + k < nb) out[i0 + k] = runv;
      runv += v[k];
    }
    __syncthreads();
    if (threadIdx.x == TSCAN_THREADS - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry_s;
}

// other endpoint of slot e of the (self-loop extended) edge list; out-of-range ids were flagged by csr_prepare_kernel
__device__ __forceinline__ int32_t other_endpoint(const int32_t* __restrict__ ei, int64_t E, int by_source,
                                                  int64_t n_val, int64_t e) {
  int64_t v;
  if (e < E) v = by_source ? __ldg(ei + E + e) : __ldg(ei + e);
  else v = e - E;
  if (v < 0 || v >= n_val) v = 0;
  return (int32_t)v;
}

struct SortIo {
  const uint32_t* keys_in; const int32_t* vals_in;   // pass > 0: the edge id travels with the key
  uint32_t* keys_out; int32_t* vals_out;
  // last pass: col[slot] = other endpoint of the edge that lands there, gathered while the pass's other CTAs rank
  int32_t* col; const int32_t* ei; int64_t E; int by_source; int64_t n_val;
};
constexpr int SORT_DYN_SMEM = 2 * SORT_TILE * 4;   // the tile reordered by digit: key, edge id

// (Carrying the edge's other endpoint as a second payload, to drop the final edge_index[perm] gather, was measured
//  slower: the three passes grew from 2.22 to 3.48 ms on C4, the gather they replace costs 1.05 ms.)
#ifndef KGB_SORT_MINB
#define KGB_SORT_MINB 4   // resident CTAs per SM: 2 -> 2.29 ms for the three C4 passes, 3 -> 1.90, 4 -> 1.43, 5 (spills) -> 1.60;
#endif                    // 2048-key tiles at 6 / 8 CTAs: 1.73 / 1.87
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(SORT_THREADS, KGB_SORT_MINB)
sort_downsweep_kernel(const SortIo io, int64_t M, int shift, const int64_t* __restrict__ offsets,
                      const int64_t* __restrict__ digit_total) {
  __shared__ int32_t warp_cnt[SORT_WARPS][RADIX];
  __shared__ int64_t total_ws[SORT_WARPS];
  __shared__ int64_t digit_base[RADIX];
  __shared__ int32_t tile_start[RADIX];   // first tile-local slot of every digit
  __shared__ int32_t scan_ws[SORT_WARPS];
  // the tile, reordered by digit: a direct scatter from registers touches up to 32 different sectors per store
  // instruction; from here consecutive threads write consecutive slots of a digit's output run
  extern __shared__ __align__(16) uint8_t sort_dyn[];
  uint32_t* skey = reinterpret_cast<uint32_t*>(sort_dyn);
  int32_t* sval = reinterpret_cast<int32_t*>(sort_dyn) + SORT_TILE;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  {
    // global base of (digit, this CTA) = keys with a smaller digit + keys with this digit in earlier CTAs
    const int64_t tot = digit_total[threadIdx.x];
    int64_t x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) total_ws[wid] = x;
    __syncthreads();
    int64_t wpre = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) wpre += (w < wid) ? total_ws[w] : 0;
    digit_base[threadIdx.x] = offsets[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] + wpre + x - tot;
  }
  const int64_t n_tiles = (M + SORT_TILE - 1) / SORT_TILE;
  const BlockRange r = block_range(n_tiles);
  for (int64_t tile = r.tile_begin; tile < r.tile_end; ++tile) {
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t tbase = tile * SORT_TILE;
    const int64_t wbase = tbase + (int64_t)wid * (32 * SORT_ITEMS);
    const int n_valid = (M - tbase) < SORT_TILE ? (int)(M - tbase) : SORT_TILE;
    uint32_t key[SORT_ITEMS];
    int32_t rank[SORT_ITEMS];
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
      const int64_t i = wbase + it * 32 + lane;
      key[it] = (i < M) ? __ldg(io.keys_in + i) : 0u;
    }
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
      const int64_t i = wbase + it * 32 + lane;
      const bool ok = i < M;
      // invalid lanes get a digit id no valid lane can have, so they never join a valid peer set
      const uint32_t d = ok ? ((key[it] >> shift) & 0xffu) : (0x100u + lane);
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      int32_t old = 0;
      if (ok && lane == leader) {
        old = warp_cnt[wid][d];
        warp_cnt[wid][d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[it] = old + __popc(peers & lt_mask);
      __syncwarp();
    }
    __syncthreads();
    // thread d: exclusive prefix over warps for digit d (warp w's keys precede warp w+1's: stable) ...
    const int d = threadIdx.x;
    int32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      const int32_t t = warp_cnt[w][d];
      warp_cnt[w][d] = run;
      run += t;
    }
    // ... and the exclusive prefix of the tile's digit totals = where digit d starts inside the reordered tile
    int32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) scan_ws[wid] = incl;
    __syncthreads();
    int32_t wpre = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) wpre += (w < wid) ? scan_ws[w] : 0;
    tile_start[d] = wpre + incl - run;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
      const int64_t i = wbase + it * 32 + lane;
      if (i < M) {
        const uint32_t dd = (key[it] >> shift) & 0xffu;
        const int32_t slot = tile_start[dd] + warp_cnt[wid][dd] + rank[it];
        skey[slot] = key[it];
        sval[slot] = FIRST ? (int32_t)i : __ldg(io.vals_in + i);
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int j = threadIdx.x; j < n_valid; j += SORT_THREADS) {
      const uint32_t k = skey[j];
      const uint32_t dd = (k >> shift) & 0xffu;
      const int64_t pos = digit_base[dd] + (int64_t)(j - tile_start[dd]);
      io.keys_out[pos] = k;
      const int32_t eid = sval[j];
      io.vals_out[pos] = eid;
      if constexpr (LAST) io.col[pos] = other_endpoint(io.ei, io.E, io.by_source, io.n_val, eid);
    }
    __syncthreads();
    digit_base[d] += run;   // digit_base[d] stayed the tile's base during the scatter above
    __syncthreads();
  }
}

// ---- exclusive scan int32 -> int64 (rowptr) ---------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
  __shared__ int64_t ws[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t s = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += ws[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) scan_i64_inplace_kernel(int64_t* __restrict__ a, int64_t n) {
  // exclusive scan in place, single CTA
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = (i < n) ? a[i] : 0;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int64_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int64_t incl = x + (wid > 0 ? warp_sums[wid - 1] : 0) + carry_s;
    if (i < n) a[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
}

// rowptr[i] = tile_offset + exclusive prefix inside the tile; rowptr[n] = total
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_offsets,
                  int64_t* __restrict__ out) {
  __shared__ int64_t ws[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int64_t x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) ws[wid] = x;
  __syncthreads();
  int64_t wpre = 0;
  for (int w = 0; w < wid; ++w) wpre += ws[w];
  int64_t run = tile_offsets[blockIdx.x] + wpre + (x - s);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
    if (base + k == n - 1) out[n] = run;
  }
}

// n_seg == 1 (no sort pass): perm = identity, col = the other endpoint in edge order
__global__ void csr_identity_kernel(const int32_t* __restrict__ ei, int64_t E, int by_source, int64_t n_val, int64_t M,
                                    int32_t* __restrict__ perm, int32_t* __restrict__ col) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x) {
    perm[k] = (int32_t)k;
    col[k] = other_endpoint(ei, E, by_source, n_val, k);
  }
}

__global__ void csr_hubs_kernel(const int64_t* __restrict__ rowptr, int64_t n_seg, int threshold, int chunk,
                                int32_t* __restrict__ hub_row, int32_t* __restrict__ hub_chunk_base,
                                int32_t* __restrict__ hub_nchunks, int32_t* __restrict__ chunk_hub,
                                int64_t max_hubs, int64_t max_chunks, int32_t* __restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_seg; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = rowptr[i + 1] - rowptr[i];
    if (d > threshold) {
      const int nch = (int)((d + chunk - 1) / chunk);
      const int h = atomicAdd(counts + 0, 1);
      const int base = atomicAdd(counts + 1, nch);
      if (h < max_hubs && (int64_t)base + nch <= max_chunks) {
        hub_row[h] = (int32_t)i;
        hub_chunk_base[h] = base;
        hub_nchunks[h] = nch;
        for (int c = 0; c < nch; ++c) chunk_hub[base + c] = h;
      }
    }
  }
}

__global__ void gcn_dis_kernel(const int32_t* __restrict__ deg, int64_t n, float* __restrict__ dis) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // (deg + 1e-12)^-0.5 in float32: IEEE sqrt then IEEE divide == torch.pow(x, -0.5) on the host
    const float d = __fadd_rn((float)deg[i], 1e-12f);
    float r = __fdiv_rn(1.0f, __fsqrt_rn(d));
    if (isinf(r)) r = 0.f;  // utils/main.py:26-28 (dead for finite deg, kept for parity)
    dis[i] = r;
  }
}

__global__ void gcn_w_kernel(const float* __restrict__ dis, const int32_t* __restrict__ ei, int64_t E,
                             int64_t n_loops, float* __restrict__ w) {
  const int64_t M = E + n_loops;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t s, d;
    if (e < E) { s = __ldg(ei + e); d = __ldg(ei + E + e); }
    else s = d = e - E;
    w[e] = __fmul_rn(__ldg(dis + d), __ldg(dis + s));
  }
}

static int ew_grid(int device, int64_t n, int threads) {
  int64_t g = ceil_div(n, threads);
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int sort_blocks(int device, int64_t M) {
  const int64_t tiles = ceil_div(M, SORT_TILE);
  int64_t b = (int64_t)sm_count(device) * 8;
  if (b > tiles) b = tiles;
  if (b < 1) b = 1;
  return (int)b;
}

struct CsrWs {
  uint32_t* keys[2];
  int32_t* vals[2];

  int32_t* table;
  int64_t* offsets;
  int64_t* tile_sums;
  size_t total;
};

static CsrWs carve_ws(void* ws, int64_t M, int64_t n_seg, int nb) {
  CsrWs w;
  size_t off = 0;
  char* base = reinterpret_cast<char*>(ws);
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t m = (size_t)(M > 0 ? M : 1);
  w.keys[0] = reinterpret_cast<uint32_t*>(take(m * 4));
  w.keys[1] = reinterpret_cast<uint32_t*>(take(m * 4));
  w.vals[0] = reinterpret_cast<int32_t*>(take(m * 4));
  w.vals[1] = reinterpret_cast<int32_t*>(take(m * 4));
  w.table = reinterpret_cast<int32_t*>(take((size_t)RADIX * nb * 4));
  w.offsets = reinterpret_cast<int64_t*>(take(((size_t)RADIX * nb + RADIX) * 8));   // + the 256 digit totals
  w.tile_sums = reinterpret_cast<int64_t*>(take((size_t)(ceil_div(n_seg > 0 ? n_seg : 1, SCAN_TILE)) * 8));
  w.total = off;
  return w;
}

}  // namespace kgb

using namespace kgb;

extern "C" {

size_t kgb_csr_build_workspace_bytes(int64_t n_edges_total, int64_t n_seg) {
  // sized for the largest block count any device could use (sm_count <= 1024 assumed)
  const int64_t tiles = ceil_div(n_edges_total > 0 ? n_edges_total : 1, SORT_TILE);
  int64_t nb = 1024 * 8;
  if (nb > tiles) nb = tiles;
  return carve_ws(nullptr, n_edges_total, n_seg, (int)nb).total;
}

int kgb_csr_build(int device, const int32_t* edge_index, int64_t E, int by_source, int64_t n_seg,
                  int64_t n_val, int64_t n_loops, int64_t* rowptr, int32_t* col, int32_t* perm,
                  int32_t* deg, int32_t* status, void* ws, size_t ws_bytes, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(E >= 0 && n_seg >= 0 && n_val >= 0 && n_loops >= 0, "negative size");
  KGB_REQUIRE(n_seg < (1ll << 31) && n_val < (1ll << 31), "node count exceeds int32");
  const int64_t M = E + n_loops;
  KGB_REQUIRE(M < (1ll << 31), "edge count (incl. self-loops) exceeds int32 slots");
  KGB_REQUIRE(rowptr && deg && status, "rowptr/deg/status must be non-NULL");
  KGB_REQUIRE(M == 0 || (col && perm), "col/perm must be non-NULL");
  KGB_REQUIRE(E == 0 || edge_index, "edge_index is NULL");
  KGB_REQUIRE(n_loops <= n_seg && n_loops <= n_val, "more self-loops than nodes");
  cudaStream_t st = (cudaStream_t)stream;

  KGB_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  if (n_seg > 0) KGB_CHECK_CUDA(cudaMemsetAsync(deg, 0, (size_t)n_seg * sizeof(int32_t), st));
  if (M == 0 || n_seg == 0) {
    KGB_CHECK_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n_seg + 1) * sizeof(int64_t), st));
    if (M > 0 && n_seg == 0) {
      // edges but no segments: every edge is out of range
      int32_t one = KGB_STATUS_OOB_INDEX;
      KGB_CHECK_CUDA(cudaMemcpyAsync(status, &one, sizeof(one), cudaMemcpyHostToDevice, st));
    }
    return KGB_OK;
  }
  const int nb = sort_blocks(device, M);
  CsrWs w = carve_ws(ws, M, n_seg, nb);
  if (ws == nullptr || ws_bytes < w.total) {
    set_error("workspace too small: need %zu bytes, got %zu", w.total, ws_bytes);
    return KGB_ERR_WORKSPACE;
  }

  csr_prepare_kernel<<<ew_grid(device, M, 256), 256, 0, st>>>(edge_index, E, by_source, n_seg, n_val, n_loops,
                                                           w.keys[0], status);
  KGB_CHECK_LAUNCH();

  // stable LSD radix sort of (key, edge id)
  int bits = 0;
  while (((int64_t)1 << bits) < n_seg) ++bits;
  int passes = (bits + 7) / 8;
  if (passes == 0) {
    csr_identity_kernel<<<ew_grid(device, M, 256), 256, 0, st>>>(edge_index, E, by_source, n_val, M, perm, col);
    KGB_CHECK_LAUNCH();
  } else {
    static bool attr_set[64] = {};
    if (device >= 64 || !attr_set[device]) {
      KGB_CHECK_CUDA(cudaFuncSetAttribute(sort_downsweep_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_DYN_SMEM));
      KGB_CHECK_CUDA(cudaFuncSetAttribute(sort_downsweep_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_DYN_SMEM));
      KGB_CHECK_CUDA(cudaFuncSetAttribute(sort_downsweep_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_DYN_SMEM));
      KGB_CHECK_CUDA(cudaFuncSetAttribute(sort_downsweep_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_DYN_SMEM));
      if (device < 64) attr_set[device] = true;
    }
  }
  int cur = 0;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * 8;
    const bool last = (pass == passes - 1);
    sort_upsweep_kernel<<<nb, SORT_THREADS, 0, st>>>(w.keys[cur], M, shift, w.table);
    KGB_CHECK_LAUNCH();
    int64_t* digit_total = w.offsets + (int64_t)RADIX * nb;
    sort_scan_kernel<<<RADIX, TSCAN_THREADS, 0, st>>>(w.table, nb, w.offsets, digit_total);
    KGB_CHECK_LAUNCH();
    SortIo io;
    io.keys_in = w.keys[cur]; io.vals_in = w.vals[cur];
    io.keys_out = w.keys[cur ^ 1];                       // the last pass writes its keys too: sorted keys -> degrees
    io.vals_out = last ? perm : w.vals[cur ^ 1];
    io.col = col; io.ei = edge_index; io.E = E; io.by_source = by_source; io.n_val = n_val;
    const bool fuse = last;   // (a separate gather kernel after the sort: 1.06 + 0.43 ms against 1.36 ms fused)
    if (pass == 0 && fuse) sort_downsweep_kernel<true, true><<<nb, SORT_THREADS, SORT_DYN_SMEM, st>>>(io, M, shift, w.offsets, digit_total);
    else if (pass == 0) sort_downsweep_kernel<true, false><<<nb, SORT_THREADS, SORT_DYN_SMEM, st>>>(io, M, shift, w.offsets, digit_total);
    else if (fuse) sort_downsweep_kernel<false, true><<<nb, SORT_THREADS, SORT_DYN_SMEM, st>>>(io, M, shift, w.offsets, digit_total);
    else sort_downsweep_kernel<false, false><<<nb, SORT_THREADS, SORT_DYN_SMEM, st>>>(io, M, shift, w.offsets, digit_total);
    KGB_CHECK_LAUNCH();
    cur ^= 1;
  }

  // degrees from the run boundaries of the sorted keys, rowptr = their exclusive scan
  csr_degree_kernel<<<ew_grid(device, M, 256), 256, 0, st>>>(w.keys[cur], M, deg);
  KGB_CHECK_LAUNCH();
  {
    const int64_t tiles = ceil_div(n_seg, SCAN_TILE);
    scan_tile_sums_kernel<<<(int)tiles, SCAN_THREADS, 0, st>>>(deg, n_seg, w.tile_sums);
    KGB_CHECK_LAUNCH();
    scan_i64_inplace_kernel<<<1, 1024, 0, st>>>(w.tile_sums, tiles);
    KGB_CHECK_LAUNCH();
    scan_apply_kernel<<<(int)tiles, SCAN_THREADS, 0, st>>>(deg, n_seg, w.tile_sums, rowptr);
    KGB_CHECK_LAUNCH();
  }
  return KGB_OK;
}

int kgb_csr_hubs(int device, const int64_t* rowptr, int64_t n_seg, int32_t threshold, int32_t chunk,
                 int32_t* hub_row, int32_t* hub_chunk_base, int32_t* hub_nchunks, int32_t* chunk_hub,
                 int64_t max_hubs, int64_t max_chunks, int32_t* counts, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(counts, "counts is NULL");
  KGB_REQUIRE(chunk > 0 && threshold >= chunk, "need threshold >= chunk > 0");
  cudaStream_t st = (cudaStream_t)stream;
  KGB_CHECK_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
  if (n_seg == 0) return KGB_OK;
  KGB_REQUIRE(rowptr && hub_row && hub_chunk_base && hub_nchunks && chunk_hub, "NULL pointer");
  csr_hubs_kernel<<<ew_grid(device, n_seg, 256), 256, 0, st>>>(rowptr, n_seg, threshold, chunk, hub_row,
                                                            hub_chunk_base, hub_nchunks, chunk_hub, max_hubs,
                                                            max_chunks, counts);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

int kgb_gcn_norm(int device, const int32_t* deg, int64_t n_nodes, const int32_t* edge_index, int64_t E,
                 int64_t n_loops, float* dis, float* w, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(n_nodes >= 0 && E >= 0 && n_loops >= 0, "negative size");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_nodes > 0) {
    KGB_REQUIRE(deg && dis, "deg/dis is NULL");
    gcn_dis_kernel<<<ew_grid(device, n_nodes, 256), 256, 0, st>>>(deg, n_nodes, dis);
    KGB_CHECK_LAUNCH();
  }
  if (w && E + n_loops > 0) {
    KGB_REQUIRE(E == 0 || edge_index, "edge_index is NULL");
    gcn_w_kernel<<<ew_grid(device, E + n_loops, 256), 256, 0, st>>>(dis, edge_index, E, n_loops, w);
    KGB_CHECK_LAUNCH();
  }
  return KGB_OK;
}

}  // extern "C"
