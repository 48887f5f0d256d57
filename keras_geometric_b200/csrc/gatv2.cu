// K6: fused GATv2 edge kernels for sm_100a (HBM-bound).
//
// Lane layout: a lane group owns (target row, block of HPG heads); inside the group
// LPH = pow2ceil(C / VEC) consecutive lanes own one head, each lane VEC (x CC chunks when
// C / VEC > 32) consecutive channels, so the per-head logit  sum_c a[h,c] * lrelu(h_i + h_j)  is
// a log2(LPH)-step xor-shuffle inside the head's lanes and every head of the block is reduced
// at the same time.  blockIdx.y enumerates head blocks when H > HPG.
//
// Forward: ONE pass over the CSR row with an online softmax (running max m, running sum l,
// rescaled accumulator); the h_j row fetched for the logit is reused for the weighted sum, so
// each edge costs one row gather.  m and l are saved per (row, head); the backward recomputes
// alpha from them instead of storing [nnz, H] tensors.
// Backward: pass 1 walks the forward CSR (per target: g_hdst, r, d att), pass 2 walks the
// transposed structure (per source: g_hsrc).  No float atomics; d att goes through per-CTA
// partials that are summed in fixed order.
//
// Scheduling (all three kernels): rows above the hub threshold are cut into chunks whose partial
// softmax states (m, l, acc) / partial gradient sums are merged in chunk order by a finish kernel;
// chunks first, then 64-row units, handed out by a self-resetting atomic queue (power-law graphs:
// a static stride correlates with the id structure, and one 10^5-edge row would serialise a warp).
#include <math.h>

#include "common.cuh"

namespace kgb {

struct GatP {
  const float* hsrc; const float* hdst;
  int64_t n_src, n_dst;
  int H, C;
  const float* att; float slope;
  const int64_t* rowptr; const int32_t* col;
  const float* bias;
  float* out; float* rowmax; float* rowden;
  // backward
  const float* g; const float* agg; const float* r_in; const float* addend;
  float* g_hdst; float* r_out; float* g_att_part; float* g_hsrc;
  // hub table of the structure being walked + workspaces
  const int32_t* hub_row; const int32_t* hub_chunk_base; const int32_t* hub_nchunks; const int32_t* chunk_hub;
  int n_hubs; int n_chunks; int hub_threshold; int hub_chunk;
  float* partial;  // fwd: acc [n_chunks,HC] | m [n_chunks,H] | l [n_chunks,H];  bwd: sums [n_chunks,HC]
  int32_t* work;   // [2 * gridDim.y] zero on entry (queue head, finished CTAs) or NULL
  const int32_t* unit_order;  // optional permutation of the row units (heaviest first)
  int64_t n_rows;  // rows of the walked structure
  // attention dropout (layers/gatv2_conv.py:252-253), fused: alpha of (edge, head) survives when Philox(edge id,
  // head) >= thr and is scaled by 1 / (1 - p); edge_id[k] = original edge id of slot k of the walked structure
  const int32_t* edge_id; uint32_t drop_thr; float drop_scale; uint32_t seed_lo, seed_hi;
  // per-edge records of the backward (one row of rec_ld floats per CSR slot, written by the per-target pass, read by
  // the per-source pass through slot_map[transposed slot] = CSR slot):
  //   [0, H)   alpha_e * dropout_e            (the message weight)
  //   [H, 2H)  ds_e = alpha_e (d_e <g_i, h_j> - r_i)   (gradient of the logit)
  //   [2H, ..) sign bits of z = h_i + h_j, bit (e * G + lane) for element e of the group's lane `lane`
  // With them the per-source pass needs neither h_i nor the logits: one row gather (g_i) instead of two plus three
  // scalar lookups, no dot products, no exp.  Shapes with one lane group per row (CC == 1, all heads in the group).
  float* rec; int rec_ld; const int32_t* slot_map;
};

// dropout factor of (edge, head): 1 / (1 - p) or 0; 1 when dropout is off
__device__ __forceinline__ float gat_drop(const GatP& p, uint32_t eid, int head) {
  if (p.drop_thr == 0u) return 1.f;
  const uint32_t bits = dropout_keep4(eid, (uint32_t)head >> 2, 1u, p.seed_lo, p.seed_hi, p.drop_thr);
  return ((bits >> (head & 3)) & 1u) ? p.drop_scale : 0.f;
}

// exp(x) for the softmax terms (x <= 0 up to rounding): one multiply + MUFU.EX2 (relative error 2^-22) instead of
// expf's eight-instruction range reduction; the three kernels are issue-bound, and forward and backward use the same
// function, so the recomputed alpha is the forward's alpha bit for bit.
__device__ __forceinline__ float exp_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}

template <int LPH>
__device__ __forceinline__ float head_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

#ifndef KGB_GAT_UNIT_ROWS
#define KGB_GAT_UNIT_ROWS 16
#endif
constexpr int GAT_UNIT_ROWS = KGB_GAT_UNIT_ROWS;  // rows per queue fetch (one atomic per warp)

// Hands (chunk | row) tasks to the lane groups of the CTA; see the header comment.
template <int G, class FC, class FR>
__device__ __forceinline__ void gat_schedule(const GatP& p, FC&& on_chunk, FR&& on_row) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gw = lane / G;
  const int64_t chunk_units = ((int64_t)p.n_chunks + GPW - 1) / GPW;
  const int64_t row_units = (p.n_rows + GAT_UNIT_ROWS - 1) / GAT_UNIT_ROWS;
  const int64_t n_units = chunk_units + row_units;
  const int64_t wpb = blockDim.x >> 5;
  int32_t* work = p.work ? p.work + 2 * blockIdx.y : nullptr;
  int64_t u = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
  while (true) {
    if (work) {
      __syncwarp();
      int32_t t = 0;
      if (lane == 0) t = atomicAdd(work, 1);
      u = __shfl_sync(0xffffffffu, t, 0);
    }
    if (u >= n_units) break;
    if (u < chunk_units) {
      const int64_t t = u * GPW + gw;
      if (t < p.n_chunks) {
        const int h = __ldg(p.chunk_hub + t);
        const int64_t row = __ldg(p.hub_row + h);
        const int64_t ci = t - __ldg(p.hub_chunk_base + h);
        const int64_t rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
        const int64_t k0 = rs + ci * p.hub_chunk;
        const int64_t k1 = (k0 + p.hub_chunk < re) ? k0 + p.hub_chunk : re;
        on_chunk(t, row, k0, k1, ci == 0);
      }
    } else {
      int64_t ui = u - chunk_units;
      if (p.unit_order) ui = __ldg(p.unit_order + ui);
      const int64_t base = ui * GAT_UNIT_ROWS;
      const int64_t lim = (base + GAT_UNIT_ROWS < p.n_rows) ? base + GAT_UNIT_ROWS : p.n_rows;
      for (int64_t row = base + gw; row < lim; row += GPW) {
        const int64_t rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
        if (p.n_hubs > 0 && (re - rs) > p.hub_threshold) continue;  // merged by the finish kernel
        on_row(row, rs, re);
      }
    }
    if (!work) u += (int64_t)gridDim.x * wpb;
  }
}

__device__ __forceinline__ void gat_queue_reset(const GatP& p) {
  if (p.work) {  // the last CTA of this head block re-arms the queue for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      int32_t* work = p.work + 2 * blockIdx.y;
      __threadfence();
      const int done = atomicAdd(work + 1, 1);
      if (done == (int)gridDim.x - 1) {
        work[0] = 0;
        work[1] = 0;
        __threadfence();
      }
    }
  }
}

// Per-lane constants shared by all kernels.
template <int VEC, int LPH, int CC, int HPG>
struct LaneCtx {
  static constexpr int G = LPH * HPG;
  int gl, head;
  unsigned gmask;
  bool on[CC];
  int off[CC];   // element offset inside an [H*C] row (0 for inactive lanes: loads stay unconditional)
  float a[CC][VEC];
  float as[CC][VEC];   // a * slope: a * lrelu(z) = (z > 0 ? a : a * slope) * z, one select instead of multiply + select
  __device__ __forceinline__ void init(const GatP& p) {
    const int lane = threadIdx.x & 31;
    gl = lane % G;
    gmask = group_mask(lane, G);
    head = blockIdx.y * HPG + gl / LPH;
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      const int c0 = ((gl % LPH) + cc * LPH) * VEC;
      on[cc] = (head < p.H) && (c0 < p.C);
      off[cc] = on[cc] ? head * p.C + c0 : 0;
      ld_vec<VEC>(p.att + off[cc], a[cc]);
      if (!on[cc]) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[cc][e] = 0.f;  // inactive lanes contribute nothing to the logits
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) as[cc][e] = a[cc][e] * p.slope;
    }
  }
};

// ------------------------------------------------------------------------------------ forward
// DROPM = 0 compiles the attention dropout out (the common case: inference, or training without it - the Philox
// path otherwise costs registers and issue slots in kernels that are issue-bound: fwd 3.10 -> 2.86 ms, per-target
// backward 4.47 -> 3.97 ms on C4, H = 8, C = 8); DROPM = 1 decides at run time.
template <int VEC, int LPH, int CC, int HPG, int DROPM>
__device__ __forceinline__ void gat_fwd_range(const GatP& p, const LaneCtx<VEC, LPH, CC, HPG>& L, int64_t row,
                                              int64_t k0, int64_t k1, float& m, float& l, float (&acc)[CC][VEC]) {
  constexpr int G = LPH * HPG;
#ifndef KGB_GAT_U
#define KGB_GAT_U 4   // measured on C4 (H8C8 / H1C64 / H8C32): U=4 at 4 CTAs/SM 3.49 / 3.78 / 11.3 ms, U=8 at 3: 3.87 / 4.18 / 11.6
#endif
  constexpr int UMAX = (KGB_GAT_U / CC) < 1 ? 1 : (KGB_GAT_U / CC);
  constexpr int U = (G < UMAX) ? G : UMAX;
  const int HC = p.H * p.C;
  float hi[CC][VEC];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    ld_vec<VEC>(p.hdst + row * HC + L.off[cc], hi[cc]);
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[cc][e] = 0.f;
  }
  m = -INFINITY;
  l = 0.f;
  int64_t k = k0;
  const bool drop = DROPM ? (p.drop_thr != 0u) : false;
  int32_t myc = (k + L.gl < k1) ? __ldg(p.col + k + L.gl) : 0;
  int32_t mye = (drop && k + L.gl < k1) ? __ldg(p.edge_id + k + L.gl) : 0;
  while (k < k1) {
    const int64_t rem = k1 - k;
    const int cnt = rem < G ? (int)rem : G;
    const int64_t kn = k + G;
    const int32_t nc = (kn + L.gl < k1) ? __ldg(p.col + kn + L.gl) : 0;
    const int32_t ne = (drop && kn + L.gl < k1) ? __ldg(p.edge_id + kn + L.gl) : 0;
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float v[U][CC][VEC];
      float s[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t c = __shfl_sync(L.gmask, myc, j + u, G);  // slots past cnt carry row 0: loaded, never used
        const float* rp = p.hsrc + (int64_t)c * HC;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) ld_vec<VEC>(rp + L.off[cc], v[u][cc]);
      }
      float mb = m;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float part = 0.f;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float z = hi[cc][e] + v[u][cc][e];
            part = fmaf(z > 0.f ? L.a[cc][e] : L.as[cc][e], z, part);
          }
        s[u] = head_sum<LPH>(part, L.gmask);
        if ((j + u) < cnt) mb = fmaxf(mb, s[u]);
      }
      const float scale = (m == -INFINITY) ? 0.f : exp_fast(m - mb);
      l *= scale;
#pragma unroll
      for (int cc = 0; cc < CC; ++cc)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[cc][e] *= scale;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if ((j + u) < cnt) {
          const float pe = exp_fast(s[u] - mb);
          l += pe;   // the softmax normalises over ALL edges; dropout only thins the weighted sum
          const float pd = drop ? pe * gat_drop(p, (uint32_t)__shfl_sync(L.gmask, mye, j + u, G), L.head) : pe;
#pragma unroll
          for (int cc = 0; cc < CC; ++cc)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[cc][e] = fmaf(pd, v[u][cc][e], acc[cc][e]);
        }
      }
      m = mb;
    }
    myc = nc;
    mye = ne;
    k = kn;
  }
}

template <int VEC, int LPH, int CC, int HPG>
__device__ __forceinline__ void gat_fwd_store(const GatP& p, const LaneCtx<VEC, LPH, CC, HPG>& L, int64_t row, float m,
                                              float l, float (&acc)[CC][VEC]) {
  const int HC = p.H * p.C;
  const float den = l + 1e-10f;
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    if (!L.on[cc]) continue;
    float o[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) o[e] = __fdiv_rn(acc[cc][e], den);
    if (p.bias) {
      float b[VEC];
      ld_vec<VEC>(p.bias + L.off[cc], b);
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] += b[e];
    }
    st_vec<VEC>(p.out + row * HC + L.off[cc], o);
  }
  if (L.head < p.H && (L.gl % LPH) == 0) {
    p.rowmax[row * p.H + L.head] = m;
    p.rowden[row * p.H + L.head] = l;
  }
}

#ifndef KGB_GAT_MINB_FWD
#define KGB_GAT_MINB_FWD 4
#endif
template <int VEC, int LPH, int CC, int HPG, int DROPM>
__global__ void __launch_bounds__(256, KGB_GAT_MINB_FWD) gatv2_fwd_kernel(const GatP p) {
  using Ctx = LaneCtx<VEC, LPH, CC, HPG>;
  constexpr int G = Ctx::G;
  Ctx L;
  L.init(p);
  const int HC = p.H * p.C;
  float* pm = p.partial + (int64_t)p.n_chunks * HC;
  float* pl = pm + (int64_t)p.n_chunks * p.H;
  gat_schedule<G>(
      p,
      [&](int64_t t, int64_t row, int64_t k0, int64_t k1, bool) {
        float m, l, acc[CC][VEC];
        gat_fwd_range<VEC, LPH, CC, HPG, DROPM>(p, L, row, k0, k1, m, l, acc);
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
          if (L.on[cc]) st_vec<VEC>(p.partial + t * HC + L.off[cc], acc[cc]);
        if (L.head < p.H && (L.gl % LPH) == 0) {
          pm[t * p.H + L.head] = m;
          pl[t * p.H + L.head] = l;
        }
      },
      [&](int64_t row, int64_t rs, int64_t re) {
        float m, l, acc[CC][VEC];
        gat_fwd_range<VEC, LPH, CC, HPG, DROPM>(p, L, row, rs, re, m, l, acc);
        gat_fwd_store<VEC, LPH, CC, HPG>(p, L, row, m, l, acc);
      });
  gat_queue_reset(p);
}

// merge the chunk states of every hub row in chunk order (log-sum-exp merge), then store like a normal row
template <int VEC, int LPH, int CC, int HPG>
__global__ void __launch_bounds__(256) gatv2_fwd_finish_kernel(const GatP p) {
  using Ctx = LaneCtx<VEC, LPH, CC, HPG>;
  constexpr int G = Ctx::G, GPW = 32 / G;
  Ctx L;
  L.init(p);
  const int HC = p.H * p.C;
  const float* pm = p.partial + (int64_t)p.n_chunks * HC;
  const float* pl = pm + (int64_t)p.n_chunks * p.H;
  const int lane = threadIdx.x & 31;
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  for (int64_t h = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + lane / G; h < p.n_hubs;
       h += (int64_t)gridDim.x * gpb) {
    const int64_t row = __ldg(p.hub_row + h);
    const int64_t base = __ldg(p.hub_chunk_base + h);
    const int nch = __ldg(p.hub_nchunks + h);
    float m = -INFINITY, l = 0.f, acc[CC][VEC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc)
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[cc][e] = 0.f;
    const int hd = L.head < p.H ? L.head : 0;
    for (int c = 0; c < nch; ++c) {
      const float cm = __ldg(pm + (base + c) * p.H + hd), cl = __ldg(pl + (base + c) * p.H + hd);
      const float mn = fmaxf(m, cm);
      const float so = (m == -INFINITY) ? 0.f : exp_fast(m - mn);
      const float sn = (cm == -INFINITY) ? 0.f : exp_fast(cm - mn);
      l = l * so + cl * sn;
#pragma unroll
      for (int cc = 0; cc < CC; ++cc) {
        float v[VEC];
        ld_vec<VEC>(p.partial + (base + c) * HC + L.off[cc], v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[cc][e] = acc[cc][e] * so + v[e] * sn;
      }
      m = mn;
    }
    gat_fwd_store<VEC, LPH, CC, HPG>(p, L, row, m, l, acc);
  }
}

// --------------------------------------------------------------------- backward, per target
// REC compiles the per-edge record stores in (opt-in path); without them the kernel is 10 % faster (4.94 -> 4.47 ms).
template <int VEC, int LPH, int CC, int HPG, int DROPM, bool REC>
__device__ __forceinline__ float4 gat_bwd_dst_range(const GatP& p, const LaneCtx<VEC, LPH, CC, HPG>& L, int64_t row,
                                                   int64_t k0, int64_t k1, float (&ghi)[CC][VEC],
                                                   float (&ga)[CC][VEC]) {
  constexpr int G = LPH * HPG;
#ifndef KGB_GAT_U_DST
#define KGB_GAT_U_DST 4   // with 4 CTAs/SM: 4.13 ms (H8C8 on C4); U=8 / 2 CTAs: 5.26, U=4 / 2 CTAs: 5.72
#endif
  constexpr int UMAX = (KGB_GAT_U_DST / CC) < 1 ? 1 : (KGB_GAT_U_DST / CC);
  constexpr int U = (G < UMAX) ? G : UMAX;
  const int HC = p.H * p.C;
  float hi[CC][VEC], gi[CC][VEC];
  float rpart = 0.f;
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    float ag[VEC];
    ld_vec<VEC>(p.hdst + row * HC + L.off[cc], hi[cc]);
    ld_vec<VEC>(p.g + row * HC + L.off[cc], gi[cc]);
    ld_vec<VEC>(p.agg + row * HC + L.off[cc], ag);
    if (p.bias) {  // `agg` is the forward output: take the fused bias off again
      float b[VEC];
      ld_vec<VEC>(p.bias + L.off[cc], b);
#pragma unroll
      for (int e = 0; e < VEC; ++e) ag[e] -= b[e];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (!L.on[cc]) gi[cc][e] = 0.f;
      ghi[cc][e] = 0.f;
      rpart = fmaf(gi[cc][e], ag[e], rpart);
    }
  }
  const float r = head_sum<LPH>(rpart, L.gmask);  // sum_k alpha_k * dalpha_k
  const int hd = L.head < p.H ? L.head : 0;
  const float m = __ldg(p.rowmax + row * p.H + hd);
  const float dinv = 1.f / (__ldg(p.rowden + row * p.H + hd) + 1e-10f);
  int64_t k = k0;
  const bool drop = DROPM ? (p.drop_thr != 0u) : false;
  int32_t myc = (k + L.gl < k1) ? __ldg(p.col + k + L.gl) : 0;
  int32_t mye = (drop && k + L.gl < k1) ? __ldg(p.edge_id + k + L.gl) : 0;
  while (k < k1) {
    const int64_t rem = k1 - k;
    const int cnt = rem < G ? (int)rem : G;
    const int64_t kn = k + G;
    const int32_t nc = (kn + L.gl < k1) ? __ldg(p.col + kn + L.gl) : 0;
    const int32_t ne = (drop && kn + L.gl < k1) ? __ldg(p.edge_id + kn + L.gl) : 0;
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float v[U][CC][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t c = __shfl_sync(L.gmask, myc, j + u, G);
        const float* rp = p.hsrc + (int64_t)c * HC;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) ld_vec<VEC>(rp + L.off[cc], v[u][cc]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float sp = 0.f, dp = 0.f;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float z = hi[cc][e] + v[u][cc][e];
            sp = fmaf(z > 0.f ? L.a[cc][e] : L.as[cc][e], z, sp);
            dp = fmaf(gi[cc][e], v[u][cc][e], dp);
          }
        const float s = head_sum<LPH>(sp, L.gmask);
        float da = head_sum<LPH>(dp, L.gmask);
        const float dfac = drop ? gat_drop(p, (uint32_t)__shfl_sync(L.gmask, mye, j + u, G), L.head) : 1.f;
        da *= dfac;   // d out / d alpha
        if ((j + u) < cnt) {
          const float alpha = exp_fast(s - m) * dinv;
          const float ds = alpha * (da - r);
#pragma unroll
          for (int cc = 0; cc < CC; ++cc)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const float z = hi[cc][e] + v[u][cc][e];
              const float lz = z > 0.f ? z : z * p.slope;
              ghi[cc][e] = fmaf(ds, z > 0.f ? L.a[cc][e] : L.as[cc][e], ghi[cc][e]);
              ga[cc][e] = fmaf(ds, lz, ga[cc][e]);
            }
          if constexpr (CC == 1 && REC) {
            if (p.rec) {   // per-edge record for the per-source pass (group-uniform branch)
              constexpr int NW = (VEC * G + 31) / 32;
              float* rrow = p.rec + (k + j + u) * (int64_t)p.rec_ld;
              if (L.head < p.H && (L.gl % LPH) == 0) {
                rrow[L.head] = alpha * dfac;
                rrow[p.H + L.head] = ds;
              }
              const int gshift = ((threadIdx.x & 31) / G) * G;
              unsigned wbits[NW];
#pragma unroll
              for (int w = 0; w < NW; ++w) wbits[w] = 0u;
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                unsigned b = __ballot_sync(L.gmask, L.on[0] && (hi[0][e] + v[u][0][e]) > 0.f) >> gshift;
                if (G < 32) b &= (1u << G) - 1u;
                wbits[(e * G) / 32] |= b << ((e * G) % 32);   // G divides 32: a ballot never straddles two words
              }
#pragma unroll
              for (int w = 0; w < NW; ++w)
                if (L.gl == w) reinterpret_cast<unsigned*>(rrow + 2 * p.H)[w] = wbits[w];
            }
          }
        }
      }
    }
    myc = nc;
    mye = ne;
    k = kn;
  }
  return make_float4(m, dinv, r, 0.f);   // the per-(target, head) record of the per-source pass
}

#ifndef KGB_GAT_MINB_DST
#define KGB_GAT_MINB_DST 4
#endif
#ifndef KGB_GAT_MINB_SRC
#define KGB_GAT_MINB_SRC 4
#endif
template <int VEC, int LPH, int CC, int HPG, int DROPM, bool REC>
__global__ void __launch_bounds__(256, KGB_GAT_MINB_DST) gatv2_bwd_dst_kernel(const GatP p) {
  using Ctx = LaneCtx<VEC, LPH, CC, HPG>;
  constexpr int G = Ctx::G;
  constexpr int SLOTS = G * CC * VEC;  // floats of the att gradient one group covers
  __shared__ float red[8][SLOTS];
  Ctx L;
  L.init(p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HC = p.H * p.C;
  float ga[CC][VEC];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc)
#pragma unroll
    for (int e = 0; e < VEC; ++e) ga[cc][e] = 0.f;
  gat_schedule<G>(
      p,
      [&](int64_t t, int64_t row, int64_t k0, int64_t k1, bool first) {
        float ghi[CC][VEC];
        const float4 stat = gat_bwd_dst_range<VEC, LPH, CC, HPG, DROPM, REC>(p, L, row, k0, k1, ghi, ga);
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
          if (L.on[cc]) st_vec<VEC>(p.partial + t * HC + L.off[cc], ghi[cc]);
        if (first && L.head < p.H && (L.gl % LPH) == 0) reinterpret_cast<float4*>(p.r_out)[row * p.H + L.head] = stat;
      },
      [&](int64_t row, int64_t rs, int64_t re) {
        float ghi[CC][VEC];
        const float4 stat = gat_bwd_dst_range<VEC, LPH, CC, HPG, DROPM, REC>(p, L, row, rs, re, ghi, ga);
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
          if (L.on[cc]) st_vec<VEC>(p.g_hdst + row * HC + L.off[cc], ghi[cc]);
        if (L.head < p.H && (L.gl % LPH) == 0) reinterpret_cast<float4*>(p.r_out)[row * p.H + L.head] = stat;
      });
  // d att: fold the groups of a warp (fixed order), then the warps of the CTA (fixed order)
  __syncwarp();
#pragma unroll
  for (int cc = 0; cc < CC; ++cc)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float t = L.on[cc] ? ga[cc][e] : 0.f;
#pragma unroll
      for (int o = G; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      ga[cc][e] = t;
    }
  if (lane < G) {
#pragma unroll
    for (int cc = 0; cc < CC; ++cc)
#pragma unroll
      for (int e = 0; e < VEC; ++e) red[warp][(cc * G + L.gl) * VEC + e] = ga[cc][e];
  }
  __syncthreads();
  for (int sidx = threadIdx.x; sidx < SLOTS; sidx += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][sidx];
    const int e = sidx % VEC, lg = (sidx / VEC) % G, cc = sidx / (VEC * G);
    const int hd = blockIdx.y * HPG + lg / LPH;
    const int c0 = ((lg % LPH) + cc * LPH) * VEC + e;
    if (hd < p.H && c0 < p.C) p.g_att_part[(int64_t)blockIdx.x * HC + hd * p.C + c0] = t;
  }
  gat_queue_reset(p);
}

// --------------------------------------------------------------------- backward, per source
// Walks the transposed structure: row j of (colptr, rowidx) lists the targets i of j's out-edges.
template <int VEC, int LPH, int CC, int HPG>
__device__ __forceinline__ void gat_bwd_src_range(const GatP& p, const LaneCtx<VEC, LPH, CC, HPG>& L, int64_t row,
                                                  int64_t k0, int64_t k1, float (&ghj)[CC][VEC]) {
  constexpr int G = LPH * HPG;
#ifndef KGB_GAT_U_SRC
#define KGB_GAT_U_SRC 3   // with 4 CTAs/SM, H8C8 on C4: U=2 6.73 ms, U=3 6.50, U=4 7.30 (U=4 / 3 CTAs: 6.70)
#endif
  constexpr int UMAX = (KGB_GAT_U_SRC / CC) < 1 ? 1 : (KGB_GAT_U_SRC / CC);
  constexpr int U = (G < UMAX) ? G : UMAX;
  const int HC = p.H * p.C;
  float hj[CC][VEC];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    ld_vec<VEC>(p.hsrc + row * HC + L.off[cc], hj[cc]);
#pragma unroll
    for (int e = 0; e < VEC; ++e) ghj[cc][e] = 0.f;
  }
  const int hd = L.head < p.H ? L.head : 0;
  int64_t k = k0;
  const bool drop = p.drop_thr != 0u;
  int32_t myc = (k + L.gl < k1) ? __ldg(p.col + k + L.gl) : 0;
  int32_t mye = (drop && k + L.gl < k1) ? __ldg(p.edge_id + k + L.gl) : 0;
  while (k < k1) {
    const int64_t rem = k1 - k;
    const int cnt = rem < G ? (int)rem : G;
    const int64_t kn = k + G;
    const int32_t nc = (kn + L.gl < k1) ? __ldg(p.col + kn + L.gl) : 0;
    const int32_t ne = (drop && kn + L.gl < k1) ? __ldg(p.edge_id + kn + L.gl) : 0;
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float hi[U][CC][VEC], gi[U][CC][VEC];
      float m[U], dinv[U], r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = __shfl_sync(L.gmask, myc, j + u, G);
        // (row max, 1 / (l + 1e-10), r) of (target i, head): ONE 16-byte record written by the per-target pass
        const float4 st = __ldg(reinterpret_cast<const float4*>(p.r_in) + i * p.H + hd);
        m[u] = st.x;
        dinv[u] = st.y;
        r[u] = st.z;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          ld_vec<VEC>(p.hdst + i * HC + L.off[cc], hi[u][cc]);
          ld_vec<VEC>(p.g + i * HC + L.off[cc], gi[u][cc]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float sp = 0.f, dp = 0.f;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float z = hi[u][cc][e] + hj[cc][e];
            sp = fmaf(z > 0.f ? L.a[cc][e] : L.as[cc][e], z, sp);
            dp = fmaf(L.on[cc] ? gi[u][cc][e] : 0.f, hj[cc][e], dp);
          }
        const float s = head_sum<LPH>(sp, L.gmask);
        float da = head_sum<LPH>(dp, L.gmask);
        const float d = drop ? gat_drop(p, (uint32_t)__shfl_sync(L.gmask, mye, j + u, G), L.head) : 1.f;
        da *= d;
        if ((j + u) < cnt) {
          const float alpha = exp_fast(s - m[u]) * dinv[u];
          const float ds = alpha * (da - r[u]);
          const float ad = alpha * d;   // the message itself: alpha_e * dropout_e * h_j
#pragma unroll
          for (int cc = 0; cc < CC; ++cc)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const float z = hi[u][cc][e] + hj[cc][e];
              ghj[cc][e] = fmaf(ds, z > 0.f ? L.a[cc][e] : L.as[cc][e], fmaf(ad, gi[u][cc][e], ghj[cc][e]));
            }
        }
      }
    }
    myc = nc;
    mye = ne;
    k = kn;
  }
}

template <int VEC, int LPH, int CC, int HPG>
__global__ void __launch_bounds__(256, KGB_GAT_MINB_SRC) gatv2_bwd_src_kernel(const GatP p) {
  using Ctx = LaneCtx<VEC, LPH, CC, HPG>;
  constexpr int G = Ctx::G;
  Ctx L;
  L.init(p);
  const int HC = p.H * p.C;
  gat_schedule<G>(
      p,
      [&](int64_t t, int64_t row, int64_t k0, int64_t k1, bool) {
        float ghj[CC][VEC];
        gat_bwd_src_range<VEC, LPH, CC, HPG>(p, L, row, k0, k1, ghj);
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
          if (L.on[cc]) st_vec<VEC>(p.partial + t * HC + L.off[cc], ghj[cc]);
      },
      [&](int64_t row, int64_t rs, int64_t re) {
        float ghj[CC][VEC];
        gat_bwd_src_range<VEC, LPH, CC, HPG>(p, L, row, rs, re, ghj);
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          if (!L.on[cc]) continue;
          if (p.addend) {  // square graph: + the per-target part of the same node's gradient (one pass less)
            float ad[VEC];
            ld_vec<VEC>(p.addend + row * HC + L.off[cc], ad);
#pragma unroll
            for (int e = 0; e < VEC; ++e) ghj[cc][e] += ad[e];
          }
          st_vec<VEC>(p.g_hsrc + row * HC + L.off[cc], ghj[cc]);
        }
      });
  gat_queue_reset(p);
}

// per-source pass from the per-edge records: g_hsrc[j] = sum over out-edges (alpha d) g_i + ds a lrelu'(z)
template <int VEC, int LPH, int CC, int HPG>
__device__ __forceinline__ void gat_bwd_src_rec_range(const GatP& p, const LaneCtx<VEC, LPH, CC, HPG>& L, int64_t k0,
                                                      int64_t k1, float (&ghj)[CC][VEC]) {
  constexpr int G = LPH * HPG;
#ifndef KGB_GAT_U_SRC_REC
#define KGB_GAT_U_SRC_REC 8
#endif
  constexpr int U = (G < KGB_GAT_U_SRC_REC) ? G : KGB_GAT_U_SRC_REC;
  const int HC = p.H * p.C;
#pragma unroll
  for (int e = 0; e < VEC; ++e) ghj[0][e] = 0.f;
  const int hd = L.head < p.H ? L.head : 0;
  int64_t k = k0;
  int32_t myc = 0, mys = 0;
  if (k + L.gl < k1) {
    myc = __ldg(p.col + k + L.gl);
    mys = __ldg(p.slot_map + k + L.gl);
  }
  while (k < k1) {
    const int64_t rem = k1 - k;
    const int cnt = rem < G ? (int)rem : G;
    const int64_t kn = k + G;
    int32_t nc = 0, ns = 0;
    if (kn + L.gl < k1) {
      nc = __ldg(p.col + kn + L.gl);
      ns = __ldg(p.slot_map + kn + L.gl);
    }
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float gi[U][VEC], ad[U], ds[U];
      unsigned sb[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = __shfl_sync(L.gmask, myc, j + u, G);       // slots past cnt carry row / slot 0: loaded, unused
        const float* rrow = p.rec + (int64_t)__shfl_sync(L.gmask, mys, j + u, G) * p.rec_ld;
        ld_vec<VEC>(p.g + i * HC + L.off[0], gi[u]);
        ad[u] = __ldg(rrow + hd);
        ds[u] = __ldg(rrow + p.H + hd);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          sb[u][e] = (__ldg(reinterpret_cast<const unsigned*>(rrow + 2 * p.H) + (e * G) / 32) >> ((e * G) % 32 + L.gl)) & 1u;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if ((j + u) < cnt) {
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            ghj[0][e] = fmaf(ds[u], sb[u][e] ? L.a[0][e] : L.as[0][e], fmaf(ad[u], gi[u][e], ghj[0][e]));
        }
      }
    }
    myc = nc;
    mys = ns;
    k = kn;
  }
}

template <int VEC, int LPH, int CC, int HPG>
__global__ void __launch_bounds__(256, 4) gatv2_bwd_src_rec_kernel(const GatP p) {
  using Ctx = LaneCtx<VEC, LPH, CC, HPG>;
  constexpr int G = Ctx::G;
  if constexpr (CC == 1) {
    Ctx L;
    L.init(p);
    const int HC = p.H * p.C;
    gat_schedule<G>(
        p,
        [&](int64_t t, int64_t row, int64_t k0, int64_t k1, bool) {
          float ghj[CC][VEC];
          gat_bwd_src_rec_range<VEC, LPH, CC, HPG>(p, L, k0, k1, ghj);
          if (L.on[0]) st_vec<VEC>(p.partial + t * HC + L.off[0], ghj[0]);
        },
        [&](int64_t row, int64_t rs, int64_t re) {
          float ghj[CC][VEC];
          gat_bwd_src_rec_range<VEC, LPH, CC, HPG>(p, L, rs, re, ghj);
          if (!L.on[0]) return;
          if (p.addend) {
            float adv[VEC];
            ld_vec<VEC>(p.addend + row * HC + L.off[0], adv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) ghj[0][e] += adv[e];
          }
          st_vec<VEC>(p.g_hsrc + row * HC + L.off[0], ghj[0]);
        });
    gat_queue_reset(p);
  }
}

// hub rows of either backward pass: out[row,:] = sum over the row's chunks (chunk order) of partial[c,:]
__global__ void __launch_bounds__(256) gat_sum_finish_kernel(const GatP p, float* __restrict__ out) {
  const int HC = p.H * p.C;
  const int64_t total = (int64_t)p.n_hubs * HC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / HC), f = (int)(i % HC);
    const int64_t base = __ldg(p.hub_chunk_base + h);
    const int nch = __ldg(p.hub_nchunks + h);
    float s = 0.f;
    for (int c = 0; c < nch; ++c) s += p.partial[(base + c) * HC + f];
    const int64_t o = (int64_t)__ldg(p.hub_row + h) * HC + f;
    if (p.addend) s += p.addend[o];
    out[o] = s;
  }
}

// ---- dispatch ----------------------------------------------------------------------------------
struct GatShape {
  int vec, lph, cc, hpg, nhb;
};

static bool gat_shape(int H, int C, bool can4, GatShape* s) {
  s->vec = (can4 && C % 4 == 0) ? 4 : 1;
  const int nv = C / s->vec;
  if (nv <= 32) {
    s->lph = pow2_ceil(nv);
    s->cc = 1;
  } else {
    s->lph = 32;
    const int n = (int)ceil_div(nv, 32);
    if (n > 4) return false;
    s->cc = n <= 2 ? 2 : 4;
  }
  int hpg = 32 / s->lph;
  const int hp2 = pow2_ceil(H);
  if (hpg > hp2) hpg = hp2;
  s->hpg = hpg;
  s->nhb = (int)ceil_div(H, hpg);
  return true;
}

enum { GAT_FWD = 0, GAT_BWD_DST = 1, GAT_BWD_SRC = 2, GAT_FWD_FINISH = 3, GAT_BWD_SRC_REC = 4 };

template <int VEC, int LPH, int CC, int HPG>
static void gat_launch(int which, dim3 grid, cudaStream_t st, const GatP& p) {
  // the dropout-free instantiations exist for the 128-bit layouts only (the scalar-width fallback decides at run time)
  const bool nodrop = (VEC == 4) && p.drop_thr == 0u;
  if (which == GAT_FWD) {
    if constexpr (VEC == 4) {
      if (nodrop) { gatv2_fwd_kernel<VEC, LPH, CC, HPG, 0><<<grid, 256, 0, st>>>(p); return; }
    }
    gatv2_fwd_kernel<VEC, LPH, CC, HPG, 1><<<grid, 256, 0, st>>>(p);
  } else if (which == GAT_FWD_FINISH) gatv2_fwd_finish_kernel<VEC, LPH, CC, HPG><<<grid, 256, 0, st>>>(p);
  else if (which == GAT_BWD_DST) {
    if (p.rec) { gatv2_bwd_dst_kernel<VEC, LPH, CC, HPG, 1, true><<<grid, 256, 0, st>>>(p); return; }
    if constexpr (VEC == 4) {
      if (nodrop) { gatv2_bwd_dst_kernel<VEC, LPH, CC, HPG, 0, false><<<grid, 256, 0, st>>>(p); return; }
    }
    gatv2_bwd_dst_kernel<VEC, LPH, CC, HPG, 1, false><<<grid, 256, 0, st>>>(p);
  } else if (which == GAT_BWD_SRC_REC) gatv2_bwd_src_rec_kernel<VEC, LPH, CC, HPG><<<grid, 256, 0, st>>>(p);
  else gatv2_bwd_src_kernel<VEC, LPH, CC, HPG><<<grid, 256, 0, st>>>(p);
}

template <int VEC>
static int gat_dispatch(const GatShape& s, int which, dim3 grid, cudaStream_t st, const GatP& p) {
#define KGB_GAT_CASE(LPH_, CC_, HPG_)                                        \
  if (s.lph == LPH_ && s.cc == CC_ && s.hpg == HPG_) {                       \
    gat_launch<VEC, LPH_, CC_, HPG_>(which, grid, st, p);                    \
    return KGB_OK;                                                           \
  }
  KGB_GAT_CASE(1, 1, 1) KGB_GAT_CASE(1, 1, 2) KGB_GAT_CASE(1, 1, 4) KGB_GAT_CASE(1, 1, 8)
  KGB_GAT_CASE(1, 1, 16) KGB_GAT_CASE(1, 1, 32)
  KGB_GAT_CASE(2, 1, 1) KGB_GAT_CASE(2, 1, 2) KGB_GAT_CASE(2, 1, 4) KGB_GAT_CASE(2, 1, 8) KGB_GAT_CASE(2, 1, 16)
  KGB_GAT_CASE(4, 1, 1) KGB_GAT_CASE(4, 1, 2) KGB_GAT_CASE(4, 1, 4) KGB_GAT_CASE(4, 1, 8)
  KGB_GAT_CASE(8, 1, 1) KGB_GAT_CASE(8, 1, 2) KGB_GAT_CASE(8, 1, 4)
  KGB_GAT_CASE(16, 1, 1) KGB_GAT_CASE(16, 1, 2)
  KGB_GAT_CASE(32, 1, 1) KGB_GAT_CASE(32, 2, 1) KGB_GAT_CASE(32, 4, 1)
#undef KGB_GAT_CASE
  set_error("gatv2: no kernel for lph=%d cc=%d hpg=%d", s.lph, s.cc, s.hpg);
  return KGB_ERR_UNSUPPORTED;
}

static int gat_grid_x(int device, int64_t rows, int64_t chunks, const GatShape& s, int per_sm) {
  const int gpw = 32 / (s.lph * s.hpg);
  const int64_t units = ceil_div(chunks, gpw) + ceil_div(rows, (int64_t)GAT_UNIT_ROWS);
  int64_t need = ceil_div(units, 8);
  const int64_t cap = (int64_t)sm_count(device) * per_sm;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

static bool gat_can4(const GatP& p) {
  return aligned16(p.hsrc) && aligned16(p.hdst) && aligned16(p.att) && (p.C % 4 == 0) &&
         (!p.bias || aligned16(p.bias)) && (!p.out || aligned16(p.out)) && (!p.g || aligned16(p.g)) &&
         (!p.agg || aligned16(p.agg)) && (!p.g_hdst || aligned16(p.g_hdst)) && (!p.g_hsrc || aligned16(p.g_hsrc)) &&
         (!p.partial || aligned16(p.partial));
}

static int gat_set_dropout(GatP& p, const kgb_gat_dropout* d) {
  p.edge_id = nullptr; p.drop_thr = 0u; p.drop_scale = 1.f; p.seed_lo = p.seed_hi = 0u;
  if (!d || d->p <= 0.f) return KGB_OK;
  if (!(d->p < 1.f) || !d->edge_id) {
    set_error("attention dropout needs 0 < p < 1 and the edge ids of the walked structure");
    return KGB_ERR_INVALID;
  }
  const double t = (double)d->p * 4294967296.0;
  p.drop_thr = t >= 4294967295.0 ? 0xffffffffu : (t < 1.0 ? 1u : (uint32_t)t);
  p.drop_scale = 1.f / (1.f - d->p);
  p.edge_id = d->edge_id;
  p.seed_lo = (uint32_t)(d->seed & 0xffffffffull);
  p.seed_hi = (uint32_t)(d->seed >> 32);
  return KGB_OK;
}

static void gat_set_hubs(GatP& p, const kgb_hub_table* hubs) {
  if (hubs && hubs->n_hubs > 0 && hubs->n_chunks > 0 && hubs->partial) {
    p.hub_row = hubs->hub_row; p.hub_chunk_base = hubs->hub_chunk_base; p.hub_nchunks = hubs->hub_nchunks;
    p.chunk_hub = hubs->chunk_hub; p.n_hubs = hubs->n_hubs; p.n_chunks = hubs->n_chunks;
    p.hub_threshold = hubs->threshold; p.hub_chunk = hubs->chunk; p.partial = hubs->partial;
  }
  p.work = hubs ? hubs->work : nullptr;
  p.unit_order = (hubs && hubs->work) ? hubs->unit_order : nullptr;
}

static int gat_run(int device, int which, GatP& p, cudaStream_t st, int per_sm, int grid_x_cap, float* finish_out) {
  GatShape s;
  if (!gat_shape(p.H, p.C, gat_can4(p), &s)) {
    set_error("gatv2: C=%d too wide for the compiled kernels", p.C);
    return KGB_ERR_UNSUPPORTED;
  }
  if (p.work && s.nhb > 32) { p.work = nullptr; p.unit_order = nullptr; }  // queue scratch holds 32 head blocks
  int gx = gat_grid_x(device, p.n_rows, p.n_chunks, s, per_sm);
  if (grid_x_cap > 0 && gx > grid_x_cap) gx = grid_x_cap;
  dim3 grid(gx, s.nhb, 1);
  int rc = (s.vec == 4) ? gat_dispatch<4>(s, which, grid, st, p) : gat_dispatch<1>(s, which, grid, st, p);
  if (rc != KGB_OK) return rc;
  KGB_CHECK_LAUNCH();
  if (p.n_hubs > 0) {
    if (which == GAT_FWD) {
      const int gpw = 32 / (s.lph * s.hpg);
      int64_t fx = ceil_div((int64_t)p.n_hubs, 8 * gpw);
      if (fx > (int64_t)sm_count(device) * 8) fx = (int64_t)sm_count(device) * 8;
      dim3 fgrid((unsigned)fx, s.nhb, 1);
      rc = (s.vec == 4) ? gat_dispatch<4>(s, GAT_FWD_FINISH, fgrid, st, p) : gat_dispatch<1>(s, GAT_FWD_FINISH, fgrid, st, p);
      if (rc != KGB_OK) return rc;
      KGB_CHECK_LAUNCH();
    } else {
      int64_t fx = ceil_div((int64_t)p.n_hubs * p.H * p.C, 256);
      if (fx > (int64_t)sm_count(device) * 8) fx = (int64_t)sm_count(device) * 8;
      gat_sum_finish_kernel<<<(unsigned)fx, 256, 0, st>>>(p, finish_out);
      KGB_CHECK_LAUNCH();
    }
  }
  return KGB_OK;
}

}  // namespace kgb

using namespace kgb;

extern "C" {

int32_t kgb_gatv2_unit_rows(void) { return kgb::GAT_UNIT_ROWS; }

size_t kgb_gatv2_partial_bytes(int32_t n_chunks, int32_t H, int32_t C) {
  if (n_chunks <= 0) return 0;
  return align_up((size_t)n_chunks * ((size_t)H * C + 2 * (size_t)H) * sizeof(float), 256);
}

int kgb_gatv2_fwd(int device, const float* hsrc, const float* hdst, int64_t n_src, int64_t n_dst, int32_t H,
                  int32_t C, const float* att, float slope, const int64_t* rowptr, const int32_t* col,
                  const float* bias, float* out, float* rowmax, float* rowden, const kgb_gat_dropout* drop,
                  const kgb_hub_table* hubs, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(H > 0 && C > 0 && n_dst >= 0 && n_src >= 0, "bad sizes");
  if (n_dst == 0) return KGB_OK;
  KGB_REQUIRE(hsrc && hdst && att && rowptr && out && rowmax && rowden, "NULL pointer");
  GatP p = {};
  p.hsrc = hsrc; p.hdst = hdst; p.n_src = n_src; p.n_dst = n_dst; p.H = H; p.C = C; p.att = att; p.slope = slope;
  p.rowptr = rowptr; p.col = col; p.bias = bias; p.out = out; p.rowmax = rowmax; p.rowden = rowden;
  p.n_rows = n_dst;
  if (gat_set_dropout(p, drop) != KGB_OK) return KGB_ERR_INVALID;
  gat_set_hubs(p, hubs);
  return gat_run(device, GAT_FWD, p, (cudaStream_t)stream, 6, 0, nullptr);
}

int kgb_gatv2_bwd_parts(int device, int64_t n_dst, int32_t H, int32_t C) {
  if (kgb::use_device(device) != KGB_OK) return -1;
  if (H <= 0 || C <= 0 || n_dst < 0) return -1;
  return sm_count(device) * 4;  // upper bound of the CTA rows kgb_gatv2_bwd_dst launches
}

int kgb_gatv2_bwd_dst(int device, const float* g, const float* agg, const float* hsrc, const float* hdst,
                      int64_t n_src, int64_t n_dst, int32_t H, int32_t C, const float* att, float slope,
                      const int64_t* rowptr, const int32_t* col, const float* rowmax, const float* rowden,
                      const float* bias, float* g_hdst, float* stat, float* g_att_part, int32_t n_parts,
                      float* rec, const kgb_gat_dropout* drop, const kgb_hub_table* hubs, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(H > 0 && C > 0 && n_dst >= 0 && n_parts > 0, "bad sizes");
  KGB_REQUIRE(g_att_part, "g_att_part is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  KGB_CHECK_CUDA(cudaMemsetAsync(g_att_part, 0, (size_t)n_parts * H * C * sizeof(float), st));
  if (n_dst == 0) return KGB_OK;
  KGB_REQUIRE(g && agg && hsrc && hdst && att && rowptr && rowmax && rowden && g_hdst && stat, "NULL pointer");
  KGB_REQUIRE(aligned16(stat), "stat must be 16-byte aligned");
  GatP p = {};
  p.hsrc = hsrc; p.hdst = hdst; p.n_src = n_src; p.n_dst = n_dst; p.H = H; p.C = C; p.att = att; p.slope = slope;
  p.rowptr = rowptr; p.col = col; p.rowmax = const_cast<float*>(rowmax); p.rowden = const_cast<float*>(rowden);
  p.g = g; p.agg = agg; p.bias = bias; p.g_hdst = g_hdst; p.r_out = stat; p.g_att_part = g_att_part;
  p.n_rows = n_dst;
  if (rec) {
    p.rec_ld = kgb_gatv2_rec_floats(H, C);
    KGB_REQUIRE(p.rec_ld > 0 && aligned16(rec), "per-edge records are not available for H=%d C=%d", H, C);
    KGB_REQUIRE(gat_can4(p) == (C % 4 == 0), "per-edge records need 16-byte aligned operands");
    p.rec = rec;
  }
  if (gat_set_dropout(p, drop) != KGB_OK) return KGB_ERR_INVALID;
  gat_set_hubs(p, hubs);
  return gat_run(device, GAT_BWD_DST, p, st, 4, n_parts, g_hdst);
}

int32_t kgb_gatv2_rec_floats(int32_t H, int32_t C) {
  if (H <= 0 || C <= 0) return 0;
  GatShape s;
  if (!gat_shape(H, C, C % 4 == 0, &s) || s.cc != 1 || s.nhb != 1) return 0;
  const int G = s.lph * s.hpg;
  const int nw = (s.vec * G + 31) / 32;
  return (2 * H + nw + 3) / 4 * 4;
}

int kgb_gatv2_bwd_src_rec(int device, const float* g, int64_t n_src, int64_t n_dst, int32_t H, int32_t C,
                          const float* att, float slope, const int64_t* colptr, const int32_t* row,
                          const int32_t* slot_map, const float* rec, const float* addend, float* g_hsrc,
                          const kgb_hub_table* hubs, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(H > 0 && C > 0 && n_src >= 0, "bad sizes");
  if (n_src == 0) return KGB_OK;
  KGB_REQUIRE(g && att && colptr && slot_map && rec && g_hsrc, "NULL pointer");
  GatP p = {};
  p.n_src = n_src; p.n_dst = n_dst; p.H = H; p.C = C; p.att = att; p.slope = slope;
  p.rowptr = colptr; p.col = row; p.g = g; p.addend = addend; p.g_hsrc = g_hsrc;
  p.rec = const_cast<float*>(rec); p.slot_map = slot_map; p.rec_ld = kgb_gatv2_rec_floats(H, C);
  p.n_rows = n_src;
  KGB_REQUIRE(p.rec_ld > 0 && aligned16(rec), "per-edge records are not available for H=%d C=%d", H, C);
  gat_set_hubs(p, hubs);
  KGB_REQUIRE(gat_can4(p) == (C % 4 == 0), "per-edge records need 16-byte aligned operands");
  return gat_run(device, GAT_BWD_SRC_REC, p, (cudaStream_t)stream, 4, 0, g_hsrc);
}

int kgb_gatv2_bwd_src(int device, const float* g, const float* hsrc, const float* hdst, int64_t n_src,
                      int64_t n_dst, int32_t H, int32_t C, const float* att, float slope, const int64_t* colptr,
                      const int32_t* row, const float* stat, const float* addend, float* g_hsrc,
                      const kgb_gat_dropout* drop, const kgb_hub_table* hubs, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(H > 0 && C > 0 && n_src >= 0, "bad sizes");
  if (n_src == 0) return KGB_OK;
  KGB_REQUIRE(g && hsrc && hdst && att && colptr && stat && g_hsrc, "NULL pointer");
  KGB_REQUIRE(aligned16(stat), "stat must be 16-byte aligned");
  GatP p = {};
  p.hsrc = hsrc; p.hdst = hdst; p.n_src = n_src; p.n_dst = n_dst; p.H = H; p.C = C; p.att = att; p.slope = slope;
  p.rowptr = colptr; p.col = row;
  p.g = g; p.r_in = stat; p.addend = addend; p.g_hsrc = g_hsrc;
  p.n_rows = n_src;
  if (gat_set_dropout(p, drop) != KGB_OK) return KGB_ERR_INVALID;
  gat_set_hubs(p, hubs);
  return gat_run(device, GAT_BWD_SRC, p, (cudaStream_t)stream, 4, 0, g_hsrc);
}

}  // extern "C"
