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
// transposed structure (per source: g_hsrc).  No atomics; d att goes through per-CTA partials
// that are summed in fixed order.
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
  const float* g; const float* agg; const float* r_in;
  float* g_hdst; float* r_out; float* g_att_part; float* g_hsrc;
};

template <int LPH>
__device__ __forceinline__ float head_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

template <int VEC, int LPH, int CC, int HPG>
struct Lay {
  static constexpr int G = LPH * HPG;
  static constexpr int GPW = 32 / G;
};

// ------------------------------------------------------------------------------------ forward
template <int VEC, int LPH, int CC, int HPG>
__global__ void __launch_bounds__(256) gatv2_fwd_kernel(const GatP p) {
  using L = Lay<VEC, LPH, CC, HPG>;
  constexpr int G = L::G, GPW = L::GPW;
  constexpr int U = (G < 4) ? G : 4;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, gw = lane / G;
  const unsigned gmask = group_mask(lane, G);
  const int head = blockIdx.y * HPG + gl / LPH;
  const int HC = p.H * p.C;
  bool on[CC];
  int off[CC];
  float a[CC][VEC];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    const int c0 = ((gl % LPH) + cc * LPH) * VEC;
    on[cc] = (head < p.H) && (c0 < p.C);
    off[cc] = head * p.C + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) a[cc][e] = 0.f;
    if (on[cc]) ld_vec<VEC>(p.att + off[cc], a[cc]);
  }
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  for (int64_t row = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + gw; row < p.n_dst;
       row += (int64_t)gridDim.x * gpb) {
    float hi[CC][VEC], acc[CC][VEC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) { hi[cc][e] = 0.f; acc[cc][e] = 0.f; }
      if (on[cc]) ld_vec<VEC>(p.hdst + row * HC + off[cc], hi[cc]);
    }
    float m = -INFINITY, l = 0.f;
    const int64_t rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
    int64_t k = rs;
    int32_t myc = (k + gl < re) ? __ldg(p.col + k + gl) : 0;
    while (k < re) {
      const int64_t rem = re - k;
      const int cnt = rem < G ? (int)rem : G;
      const int64_t kn = k + G;
      const int32_t nc = (kn + gl < re) ? __ldg(p.col + kn + gl) : 0;
      for (int j = 0; j < cnt; j += U) {
        float v[U][CC][VEC];
        float s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int32_t c = __shfl_sync(gmask, myc, j + u, G);
          const bool ok = (j + u) < cnt;
          const float* rp = p.hsrc + (int64_t)c * HC;
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            if (ok && on[cc]) ld_vec<VEC>(rp + off[cc], v[u][cc]);
            else {
#pragma unroll
              for (int e = 0; e < VEC; ++e) v[u][cc][e] = 0.f;
            }
          }
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
              part = fmaf(a[cc][e], z > 0.f ? z : z * p.slope, part);
            }
          s[u] = head_sum<LPH>(part, gmask);
          if ((j + u) < cnt) mb = fmaxf(mb, s[u]);
        }
        const float scale = (m == -INFINITY) ? 0.f : expf(m - mb);
        l *= scale;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[cc][e] *= scale;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if ((j + u) < cnt) {
            const float pe = expf(s[u] - mb);
            l += pe;
#pragma unroll
            for (int cc = 0; cc < CC; ++cc)
#pragma unroll
              for (int e = 0; e < VEC; ++e) acc[cc][e] = fmaf(pe, v[u][cc][e], acc[cc][e]);
          }
        }
        m = mb;
      }
      myc = nc;
      k = kn;
    }
    const float den = l + 1e-10f;
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      if (!on[cc]) continue;
      float o[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = __fdiv_rn(acc[cc][e], den);
      if (p.bias) {
        float b[VEC];
        ld_vec<VEC>(p.bias + off[cc], b);
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] += b[e];
      }
      st_vec<VEC>(p.out + row * HC + off[cc], o);
    }
    if (head < p.H && (gl % LPH) == 0) {
      p.rowmax[row * p.H + head] = m;
      p.rowden[row * p.H + head] = l;
    }
  }
}

// --------------------------------------------------------------------- backward, per target
template <int VEC, int LPH, int CC, int HPG>
__global__ void __launch_bounds__(256) gatv2_bwd_dst_kernel(const GatP p) {
  using L = Lay<VEC, LPH, CC, HPG>;
  constexpr int G = L::G, GPW = L::GPW;
  constexpr int U = (G < 4) ? G : 4;
  constexpr int SLOTS = G * CC * VEC;  // floats of the att gradient one group covers
  __shared__ float red[8][SLOTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % G, gw = lane / G;
  const unsigned gmask = group_mask(lane, G);
  const int head = blockIdx.y * HPG + gl / LPH;
  const int HC = p.H * p.C;
  bool on[CC];
  int off[CC];
  float a[CC][VEC], ga[CC][VEC];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    const int c0 = ((gl % LPH) + cc * LPH) * VEC;
    on[cc] = (head < p.H) && (c0 < p.C);
    off[cc] = head * p.C + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) { a[cc][e] = 0.f; ga[cc][e] = 0.f; }
    if (on[cc]) ld_vec<VEC>(p.att + off[cc], a[cc]);
  }
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  for (int64_t row = (int64_t)blockIdx.x * gpb + (int64_t)warp * GPW + gw; row < p.n_dst;
       row += (int64_t)gridDim.x * gpb) {
    float hi[CC][VEC], gi[CC][VEC], ghi[CC][VEC];
    float rpart = 0.f;
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) { hi[cc][e] = 0.f; gi[cc][e] = 0.f; ghi[cc][e] = 0.f; }
      if (on[cc]) {
        float ag[VEC];
        ld_vec<VEC>(p.hdst + row * HC + off[cc], hi[cc]);
        ld_vec<VEC>(p.g + row * HC + off[cc], gi[cc]);
        ld_vec<VEC>(p.agg + row * HC + off[cc], ag);
#pragma unroll
        for (int e = 0; e < VEC; ++e) rpart = fmaf(gi[cc][e], ag[e], rpart);
      }
    }
    const float r = head_sum<LPH>(rpart, gmask);  // sum_k alpha_k * dalpha_k
    float m = 0.f, dinv = 0.f;
    if (head < p.H) {
      m = __ldg(p.rowmax + row * p.H + head);
      dinv = 1.f / (__ldg(p.rowden + row * p.H + head) + 1e-10f);
    }
    const int64_t rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
    int64_t k = rs;
    int32_t myc = (k + gl < re) ? __ldg(p.col + k + gl) : 0;
    while (k < re) {
      const int64_t rem = re - k;
      const int cnt = rem < G ? (int)rem : G;
      const int64_t kn = k + G;
      const int32_t nc = (kn + gl < re) ? __ldg(p.col + kn + gl) : 0;
      for (int j = 0; j < cnt; j += U) {
        float v[U][CC][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int32_t c = __shfl_sync(gmask, myc, j + u, G);
          const bool ok = (j + u) < cnt;
          const float* rp = p.hsrc + (int64_t)c * HC;
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            if (ok && on[cc]) ld_vec<VEC>(rp + off[cc], v[u][cc]);
            else {
#pragma unroll
              for (int e = 0; e < VEC; ++e) v[u][cc][e] = 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float sp = 0.f, dp = 0.f;
#pragma unroll
          for (int cc = 0; cc < CC; ++cc)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const float z = hi[cc][e] + v[u][cc][e];
              sp = fmaf(a[cc][e], z > 0.f ? z : z * p.slope, sp);
              dp = fmaf(gi[cc][e], v[u][cc][e], dp);
            }
          const float s = head_sum<LPH>(sp, gmask);
          const float da = head_sum<LPH>(dp, gmask);
          if ((j + u) < cnt && head < p.H) {
            const float alpha = expf(s - m) * dinv;
            const float ds = alpha * (da - r);
#pragma unroll
            for (int cc = 0; cc < CC; ++cc)
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                const float z = hi[cc][e] + v[u][cc][e];
                const float lz = z > 0.f ? z : z * p.slope;
                const float dz = ds * a[cc][e] * (z > 0.f ? 1.f : p.slope);
                ghi[cc][e] += dz;
                ga[cc][e] = fmaf(ds, lz, ga[cc][e]);
              }
          }
        }
      }
      myc = nc;
      k = kn;
    }
#pragma unroll
    for (int cc = 0; cc < CC; ++cc)
      if (on[cc]) st_vec<VEC>(p.g_hdst + row * HC + off[cc], ghi[cc]);
    if (head < p.H && (gl % LPH) == 0) p.r_out[row * p.H + head] = r;
  }
  // d att: fold the groups of a warp (fixed order), then the warps of the CTA (fixed order)
#pragma unroll
  for (int cc = 0; cc < CC; ++cc)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float t = ga[cc][e];
#pragma unroll
      for (int o = G; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      ga[cc][e] = t;
    }
  if (lane < G) {
#pragma unroll
    for (int cc = 0; cc < CC; ++cc)
#pragma unroll
      for (int e = 0; e < VEC; ++e) red[warp][(cc * G + gl) * VEC + e] = ga[cc][e];
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
}

// --------------------------------------------------------------------- backward, per source
// Walks the transposed structure: row j of (colptr, rowidx) lists the targets i of j's out-edges.
template <int VEC, int LPH, int CC, int HPG>
__global__ void __launch_bounds__(256) gatv2_bwd_src_kernel(const GatP p) {
  using L = Lay<VEC, LPH, CC, HPG>;
  constexpr int G = L::G, GPW = L::GPW;
  constexpr int U = (G < 2) ? G : 2;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, gw = lane / G;
  const unsigned gmask = group_mask(lane, G);
  const int head = blockIdx.y * HPG + gl / LPH;
  const int HC = p.H * p.C;
  bool on[CC];
  int off[CC];
  float a[CC][VEC];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    const int c0 = ((gl % LPH) + cc * LPH) * VEC;
    on[cc] = (head < p.H) && (c0 < p.C);
    off[cc] = head * p.C + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) a[cc][e] = 0.f;
    if (on[cc]) ld_vec<VEC>(p.att + off[cc], a[cc]);
  }
  const int64_t gpb = (int64_t)(blockDim.x >> 5) * GPW;
  for (int64_t row = (int64_t)blockIdx.x * gpb + (int64_t)(threadIdx.x >> 5) * GPW + gw; row < p.n_src;
       row += (int64_t)gridDim.x * gpb) {
    float hj[CC][VEC], ghj[CC][VEC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) { hj[cc][e] = 0.f; ghj[cc][e] = 0.f; }
      if (on[cc]) ld_vec<VEC>(p.hsrc + row * HC + off[cc], hj[cc]);
    }
    const int64_t rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
    int64_t k = rs;
    int32_t myc = (k + gl < re) ? __ldg(p.col + k + gl) : 0;
    while (k < re) {
      const int64_t rem = re - k;
      const int cnt = rem < G ? (int)rem : G;
      const int64_t kn = k + G;
      const int32_t nc = (kn + gl < re) ? __ldg(p.col + kn + gl) : 0;
      for (int j = 0; j < cnt; j += U) {
        float hi[U][CC][VEC], gi[U][CC][VEC];
        float m[U], dinv[U], r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t i = __shfl_sync(gmask, myc, j + u, G);
          const bool ok = (j + u) < cnt;
          m[u] = 0.f; dinv[u] = 0.f; r[u] = 0.f;
          if (ok && head < p.H) {
            m[u] = __ldg(p.rowmax + i * p.H + head);
            dinv[u] = 1.f / (__ldg(p.rowden + i * p.H + head) + 1e-10f);
            r[u] = __ldg(p.r_in + i * p.H + head);
          }
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            if (ok && on[cc]) {
              ld_vec<VEC>(p.hdst + i * HC + off[cc], hi[u][cc]);
              ld_vec<VEC>(p.g + i * HC + off[cc], gi[u][cc]);
            } else {
#pragma unroll
              for (int e = 0; e < VEC; ++e) { hi[u][cc][e] = 0.f; gi[u][cc][e] = 0.f; }
            }
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
              sp = fmaf(a[cc][e], z > 0.f ? z : z * p.slope, sp);
              dp = fmaf(gi[u][cc][e], hj[cc][e], dp);
            }
          const float s = head_sum<LPH>(sp, gmask);
          const float da = head_sum<LPH>(dp, gmask);
          if ((j + u) < cnt && head < p.H) {
            const float alpha = expf(s - m[u]) * dinv[u];
            const float ds = alpha * (da - r[u]);
#pragma unroll
            for (int cc = 0; cc < CC; ++cc)
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                const float z = hi[u][cc][e] + hj[cc][e];
                ghj[cc][e] += ds * a[cc][e] * (z > 0.f ? 1.f : p.slope) + alpha * gi[u][cc][e];
              }
          }
        }
      }
      myc = nc;
      k = kn;
    }
#pragma unroll
    for (int cc = 0; cc < CC; ++cc)
      if (on[cc]) st_vec<VEC>(p.g_hsrc + row * HC + off[cc], ghj[cc]);
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

enum { GAT_FWD = 0, GAT_BWD_DST = 1, GAT_BWD_SRC = 2 };

template <int VEC, int LPH, int CC, int HPG>
static void gat_launch(int which, dim3 grid, cudaStream_t st, const GatP& p) {
  if (which == GAT_FWD) gatv2_fwd_kernel<VEC, LPH, CC, HPG><<<grid, 256, 0, st>>>(p);
  else if (which == GAT_BWD_DST) gatv2_bwd_dst_kernel<VEC, LPH, CC, HPG><<<grid, 256, 0, st>>>(p);
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

static int gat_grid_x(int device, int64_t rows, const GatShape& s, int per_sm) {
  const int64_t gpb = 8 * (32 / (s.lph * s.hpg));
  int64_t need = ceil_div(rows, gpb);
  const int64_t cap = (int64_t)sm_count(device) * per_sm;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

static int gat_run(int device, int which, int64_t rows, GatP& p, cudaStream_t st, int grid_x_override = 0) {
  const bool can4 = aligned16(p.hsrc) && aligned16(p.hdst) && aligned16(p.att) && (p.C % 4 == 0) &&
                    (!p.bias || aligned16(p.bias)) && (!p.out || aligned16(p.out)) && (!p.g || aligned16(p.g)) &&
                    (!p.agg || aligned16(p.agg)) && (!p.g_hdst || aligned16(p.g_hdst)) &&
                    (!p.g_hsrc || aligned16(p.g_hsrc));
  GatShape s;
  if (!gat_shape(p.H, p.C, can4, &s)) {
    set_error("gatv2: C=%d too wide for the compiled kernels", p.C);
    return KGB_ERR_UNSUPPORTED;
  }
  const int gx = grid_x_override ? grid_x_override : gat_grid_x(device, rows, s, 8);
  dim3 grid(gx, s.nhb, 1);
  int rc = (s.vec == 4) ? gat_dispatch<4>(s, which, grid, st, p) : gat_dispatch<1>(s, which, grid, st, p);
  if (rc != KGB_OK) return rc;
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

}  // namespace kgb

using namespace kgb;

extern "C" {

int kgb_gatv2_fwd(int device, const float* hsrc, const float* hdst, int64_t n_src, int64_t n_dst, int32_t H,
                  int32_t C, const float* att, float slope, const int64_t* rowptr, const int32_t* col,
                  const float* bias, float* out, float* rowmax, float* rowden, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(H > 0 && C > 0 && n_dst >= 0 && n_src >= 0, "bad sizes");
  if (n_dst == 0) return KGB_OK;
  KGB_REQUIRE(hsrc && hdst && att && rowptr && out && rowmax && rowden, "NULL pointer");
  GatP p = {};
  p.hsrc = hsrc; p.hdst = hdst; p.n_src = n_src; p.n_dst = n_dst; p.H = H; p.C = C; p.att = att; p.slope = slope;
  p.rowptr = rowptr; p.col = col; p.bias = bias; p.out = out; p.rowmax = rowmax; p.rowden = rowden;
  return gat_run(device, GAT_FWD, n_dst, p, (cudaStream_t)stream);
}

int kgb_gatv2_bwd_parts(int device, int64_t n_dst, int32_t H, int32_t C) {
  if (kgb::use_device(device) != KGB_OK) return -1;
  GatShape s;
  if (H <= 0 || C <= 0 || !gat_shape(H, C, C % 4 == 0, &s)) return -1;
  // upper bound over both vector widths: the scalar layout never needs more CTAs than this
  GatShape s1;
  const int a = gat_grid_x(device, n_dst, s, 4);
  if (!gat_shape(H, C, false, &s1)) return a;  // scalar layout not compiled for this width
  const int b = gat_grid_x(device, n_dst, s1, 4);
  return a > b ? a : b;
}

int kgb_gatv2_bwd_dst(int device, const float* g, const float* agg, const float* hsrc, const float* hdst,
                      int64_t n_src, int64_t n_dst, int32_t H, int32_t C, const float* att, float slope,
                      const int64_t* rowptr, const int32_t* col, const float* rowmax, const float* rowden,
                      float* g_hdst, float* r, float* g_att_part, int32_t n_parts, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(H > 0 && C > 0 && n_dst >= 0 && n_parts > 0, "bad sizes");
  KGB_REQUIRE(g_att_part, "g_att_part is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  KGB_CHECK_CUDA(cudaMemsetAsync(g_att_part, 0, (size_t)n_parts * H * C * sizeof(float), st));
  if (n_dst == 0) return KGB_OK;
  KGB_REQUIRE(g && agg && hsrc && hdst && att && rowptr && rowmax && rowden && g_hdst && r, "NULL pointer");
  GatP p = {};
  p.hsrc = hsrc; p.hdst = hdst; p.n_src = n_src; p.n_dst = n_dst; p.H = H; p.C = C; p.att = att; p.slope = slope;
  p.rowptr = rowptr; p.col = col; p.rowmax = const_cast<float*>(rowmax); p.rowden = const_cast<float*>(rowden);
  p.g = g; p.agg = agg; p.g_hdst = g_hdst; p.r_out = r; p.g_att_part = g_att_part;
  // the CTA count must not exceed the partial rows the caller allocated
  const bool can4 = aligned16(hsrc) && aligned16(hdst) && aligned16(att) && (C % 4 == 0) && aligned16(g) &&
                    aligned16(agg) && aligned16(g_hdst);
  GatShape s;
  if (!gat_shape(H, C, can4, &s)) { set_error("gatv2: C too wide"); return KGB_ERR_UNSUPPORTED; }
  int gx = gat_grid_x(device, n_dst, s, 4);
  if (gx > n_parts) gx = n_parts;
  return gat_run(device, GAT_BWD_DST, n_dst, p, st, gx);
}

int kgb_gatv2_bwd_src(int device, const float* g, const float* hsrc, const float* hdst, int64_t n_src,
                      int64_t n_dst, int32_t H, int32_t C, const float* att, float slope, const int64_t* colptr,
                      const int32_t* row, const float* rowmax, const float* rowden, const float* r,
                      float* g_hsrc, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(H > 0 && C > 0 && n_src >= 0, "bad sizes");
  if (n_src == 0) return KGB_OK;
  KGB_REQUIRE(g && hsrc && hdst && att && colptr && rowmax && rowden && r && g_hsrc, "NULL pointer");
  GatP p = {};
  p.hsrc = hsrc; p.hdst = hdst; p.n_src = n_src; p.n_dst = n_dst; p.H = H; p.C = C; p.att = att; p.slope = slope;
  p.rowptr = colptr; p.col = row; p.rowmax = const_cast<float*>(rowmax); p.rowden = const_cast<float*>(rowden);
  p.g = g; p.r_in = r; p.g_hsrc = g_hsrc;
  return gat_run(device, GAT_BWD_SRC, n_src, p, (cudaStream_t)stream);
}

}  // extern "C"
