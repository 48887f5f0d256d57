// Row-streaming helpers of the training step (HBM-bound, one pass each):
//   * kgb_relu_bwd_colsum: g_pre = g * (y > 0) and the bias gradient sum_r g_pre[r,:] in the same pass
//     (per-CTA column partials in a fixed order -> kgb_reduce_parts; deterministic, no float atomics),
//   * kgb_softmax_xent_fwd / _bwd: mean softmax cross-entropy over integer labels, one warp per row,
//   * kgb_l2_normalize / _bwd: y = x / max(||x||_2, eps) per row (SAGEConv normalize=True, sage_conv.py:432-433),
//     one warp per row, the row is read once and kept in registers.
#include "common.cuh"

namespace kgb {

constexpr int CS_THREADS = 256;

// Thread t owns column group cg = t % CG (4 floats) and row lane rl = t / CG; a CTA streams a contiguous block of rows.
template <bool HAS_Y, bool WRITE>
__global__ void __launch_bounds__(CS_THREADS)
relu_bwd_colsum_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                       int64_t rows, int F, float* __restrict__ out, int64_t ldo, float* __restrict__ partial,
                       int64_t rows_per_cta) {
  extern __shared__ float4 cs_smem[];  // [RL][CG]
  const int CG = F >> 2;
  const int RL = CS_THREADS / CG;      // >= 1 (F <= 1024)
  const int cg = threadIdx.x % CG;
  const int rl = threadIdx.x / CG;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < rows) ? r0 + rows_per_cta : rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < RL) {
#pragma unroll 4
    for (int64_t r = r0 + rl; r < r1; r += RL) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(g + r * ldg) + cg);
      if (HAS_Y) {
        const float4 yy = __ldcs(reinterpret_cast<const float4*>(y + r * ldy) + cg);
        v.x = yy.x > 0.f ? v.x : 0.f;
        v.y = yy.y > 0.f ? v.y : 0.f;
        v.z = yy.z > 0.f ? v.z : 0.f;
        v.w = yy.w > 0.f ? v.w : 0.f;
      }
      if (WRITE) *(reinterpret_cast<float4*>(out + r * ldo) + cg) = v;
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    cs_smem[rl * CG + cg] = acc;
  }
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < RL; ++j) {
      const float4 o = cs_smem[j * CG + cg];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    *(reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * F) + cg) = acc;
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// PER = ceil(C / 32) logits per lane kept in registers (C <= 32 * PER); a warp works on R rows at a time so that R
// rows of loads are in flight per warp (a row of 47 classes is only 188 bytes).
template <int PER, int R, bool BWD>
__global__ void __launch_bounds__(256)
softmax_xent_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int64_t rows, int C,
                    float* __restrict__ row_loss, const float* __restrict__ gup, float scale,
                    float* __restrict__ dlogits, int64_t ldd) {
  const int lane = threadIdx.x & 31;
  const int64_t wpb = blockDim.x >> 5;
  float gs = 0.f;
  if (BWD) gs = __ldg(gup) * scale;
  for (int64_t r0 = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * R; r0 < rows; r0 += (int64_t)gridDim.x * wpb * R) {
    float v[R][PER];
    int64_t lab[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t r = r0 + j;
      lab[j] = (r < rows) ? __ldg(labels + r) : 0;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + i * 32;
        v[j][i] = (c < C && r < rows) ? __ldg(logits + r * ld + c) : -INFINITY;
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t r = r0 + j;
      if (r >= rows) break;   // warp-uniform
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < PER; ++i) m = fmaxf(m, v[j][i]);
      m = warp_max(m);
      float picked = 0.f;   // logit of the label (forward only)
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + i * 32;
        if (!BWD && c < C && (int64_t)c == lab[j]) picked = v[j][i];
        v[j][i] = (c < C) ? expf(v[j][i] - m) : 0.f;
        s += v[j][i];
      }
      s = warp_sum(s);
      if (!BWD) {
        picked = warp_sum(picked);   // loss = logsumexp - logit[label]; labels outside [0, C) contribute 0
        if (lane == 0) row_loss[r] = (logf(s) + m) - picked;
      } else {
        const float inv = 1.f / s;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int c = lane + i * 32;
          if (c < C) dlogits[r * ldd + c] = (v[j][i] * inv - ((int64_t)c == lab[j] ? 1.f : 0.f)) * gs;
        }
      }
    }
  }
}

// ---- row-wise L2 normalisation ----------------------------------------------------------------------------------
// keras.ops.normalize(axis=-1, order=2) on the torch backend: x / clamp(||x||, min=eps).  One warp per row, lane l owns
// elements l, l + 32, ... (PER of them, F <= 32 * PER); the squared norm is a fixed-order warp reduction.
// Backward (y = x / d, d = max(n, eps)): n >= eps -> gx = (g - y <g, y>) / d (clamp passes the gradient, inclusive
// like torch.clamp's mask), n < eps -> gx = g / d.
template <int PER, bool BWD>
__global__ void __launch_bounds__(256)
l2_normalize_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ y, int64_t ldy,
                    float* __restrict__ norm, int64_t rows, int F, float eps, float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    float v[PER];
    float s = 0.f;
    if (!BWD) {
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + i * 32;
        v[i] = (c < F) ? __ldg(a + r * lda + c) : 0.f;
        s = fmaf(v[i], v[i], s);
      }
      const float n = __fsqrt_rn(warp_sum(s));
      const float d = fmaxf(n, eps);
      if (lane == 0) norm[r] = n;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + i * 32;
        if (c < F) out[r * ldo + c] = __fdiv_rn(v[i], d);
      }
    } else {
      float yy[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + i * 32;
        v[i] = (c < F) ? __ldg(a + r * lda + c) : 0.f;        // upstream gradient
        yy[i] = (c < F) ? __ldg(y + r * ldy + c) : 0.f;
        s = fmaf(v[i], yy[i], s);
      }
      const float n = __ldg(norm + r);
      const float d = fmaxf(n, eps);
      const float dot = (n >= eps) ? warp_sum(s) : 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + i * 32;
        if (c < F) out[r * ldo + c] = __fdiv_rn(v[i] - yy[i] * dot, d);
      }
    }
  }
}

// rows wider than 1024 floats: the same arithmetic, the row is read twice instead of being kept in registers
template <bool BWD>
__global__ void __launch_bounds__(256)
l2_normalize_wide_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ y, int64_t ldy,
                         float* __restrict__ norm, int64_t rows, int F, float eps, float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    float s = 0.f;
    for (int c = lane; c < F; c += 32) {
      const float v = __ldg(a + r * lda + c);
      s = fmaf(v, BWD ? __ldg(y + r * ldy + c) : v, s);
    }
    s = warp_sum(s);
    const float n = BWD ? __ldg(norm + r) : __fsqrt_rn(s);
    const float d = fmaxf(n, eps);
    if (!BWD && lane == 0) norm[r] = n;
    const float dot = (BWD && n >= eps) ? s : 0.f;
    for (int c = lane; c < F; c += 32) {
      const float v = __ldg(a + r * lda + c);
      out[r * ldo + c] = __fdiv_rn(BWD ? v - __ldg(y + r * ldy + c) * dot : v, d);
    }
  }
}

template <bool BWD>
static int launch_l2(int device, const float* a, int64_t lda, const float* y, int64_t ldy, float* norm, int64_t rows,
                     int F, float eps, float* out, int64_t ldo, cudaStream_t st) {
  int64_t grid = ceil_div(rows, 8);
  const int64_t cap = (int64_t)sm_count(device) * 8;
  if (grid > cap) grid = cap;
#define KGB_L2(P) l2_normalize_kernel<P, BWD><<<(int)grid, 256, 0, st>>>(a, lda, y, ldy, norm, rows, F, eps, out, ldo)
  if (F <= 32) KGB_L2(1);
  else if (F <= 64) KGB_L2(2);
  else if (F <= 128) KGB_L2(4);
  else if (F <= 256) KGB_L2(8);
  else if (F <= 512) KGB_L2(16);
  else if (F <= 1024) KGB_L2(32);
  else l2_normalize_wide_kernel<BWD><<<(int)grid, 256, 0, st>>>(a, lda, y, ldy, norm, rows, F, eps, out, ldo);
#undef KGB_L2
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

template <bool BWD>
static int launch_xent(int device, const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int C,
                       float* row_loss, const float* gup, float scale, float* dlogits, int64_t ldd, cudaStream_t st) {
  const int64_t cap = (int64_t)sm_count(device) * 8;
#define KGB_XENT(P, R_)                                                                                          \
  do {                                                                                                           \
    int64_t grid = ceil_div(rows, 8 * (R_));                                                                     \
    if (grid > cap) grid = cap;                                                                                  \
    softmax_xent_kernel<P, R_, BWD><<<(int)grid, 256, 0, st>>>(logits, ld, labels, rows, C, row_loss, gup, scale, \
                                                               dlogits, ldd);                                    \
  } while (0)
  if (C <= 32) KGB_XENT(1, 4);
  else if (C <= 64) KGB_XENT(2, 4);
  else if (C <= 128) KGB_XENT(4, 2);
  else if (C <= 256) KGB_XENT(8, 1);
  else if (C <= 512) KGB_XENT(16, 1);
  else KGB_XENT(32, 1);
#undef KGB_XENT
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

}  // namespace kgb

using namespace kgb;

extern "C" {

int32_t kgb_colsum_parts(int device, int64_t rows) {
  if (kgb::use_device(device) != KGB_OK) return -1;
  int64_t parts = (int64_t)sm_count(device) * 4;
  const int64_t by_rows = ceil_div(rows, 64);  // at least 64 rows per CTA
  if (parts > by_rows) parts = by_rows;
  return (int32_t)(parts < 1 ? 1 : parts);
}

int kgb_relu_bwd_colsum(int device, const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t F,
                        float* out, int64_t ldo, float* partial, int32_t n_parts, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(rows > 0 && F > 0 && F % 4 == 0 && F <= 1024, "relu_bwd_colsum: rows > 0, F a multiple of 4 and <= 1024");
  KGB_REQUIRE(g && partial, "NULL pointer");
  KGB_REQUIRE(n_parts == kgb_colsum_parts(device, rows), "n_parts must come from kgb_colsum_parts");
  KGB_REQUIRE(ldg % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && (reinterpret_cast<uintptr_t>(partial) & 15) == 0,
              "g / partial must be 16-byte aligned with ld % 4 == 0");
  KGB_REQUIRE(!y || (ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0), "y must be 16-byte aligned");
  KGB_REQUIRE(!out || (ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0), "out must be 16-byte aligned");
  KGB_REQUIRE(!y || out, "the masked gradient needs an output buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const int CG = F / 4;
  const int RL = CS_THREADS / CG;
  const size_t smem = (size_t)RL * CG * sizeof(float4);
  const int64_t rpc = ceil_div(rows, (int64_t)n_parts);
  if (y)
    relu_bwd_colsum_kernel<true, true><<<n_parts, CS_THREADS, smem, st>>>(g, ldg, y, ldy, rows, F, out, ldo, partial, rpc);
  else if (out)
    relu_bwd_colsum_kernel<false, true><<<n_parts, CS_THREADS, smem, st>>>(g, ldg, y, ldy, rows, F, out, ldo, partial, rpc);
  else
    relu_bwd_colsum_kernel<false, false><<<n_parts, CS_THREADS, smem, st>>>(g, ldg, y, ldy, rows, F, out, ldo, partial, rpc);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

int kgb_softmax_xent_fwd(int device, const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C,
                         float* row_loss, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(rows >= 0 && C > 0 && C <= 1024, "softmax_xent: 0 < C <= 1024");
  if (rows == 0) return KGB_OK;
  KGB_REQUIRE(logits && labels && row_loss, "NULL pointer");
  return launch_xent<false>(device, logits, ld, labels, rows, C, row_loss, nullptr, 0.f, nullptr, 0, (cudaStream_t)stream);
}

int kgb_softmax_xent_bwd(int device, const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C,
                         const float* grad_loss, float scale, float* dlogits, int64_t ldd, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(rows >= 0 && C > 0 && C <= 1024, "softmax_xent: 0 < C <= 1024");
  if (rows == 0) return KGB_OK;
  KGB_REQUIRE(logits && labels && grad_loss && dlogits, "NULL pointer");
  return launch_xent<true>(device, logits, ld, labels, rows, C, nullptr, grad_loss, scale, dlogits, ldd, (cudaStream_t)stream);
}

int kgb_l2_normalize(int device, const float* x, int64_t ldx, int64_t rows, int32_t F, float eps, float* y, int64_t ldy,
                     float* norm, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(rows >= 0 && F > 0, "l2_normalize: F > 0");
  if (rows == 0) return KGB_OK;
  KGB_REQUIRE(x && y && norm, "NULL pointer");
  return launch_l2<false>(device, x, ldx, nullptr, 0, norm, rows, F, eps, y, ldy, (cudaStream_t)stream);
}

int kgb_l2_normalize_bwd(int device, const float* g, int64_t ldg, const float* y, int64_t ldy, const float* norm,
                         int64_t rows, int32_t F, float eps, float* gx, int64_t ldgx, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(rows >= 0 && F > 0, "l2_normalize: F > 0");
  if (rows == 0) return KGB_OK;
  KGB_REQUIRE(g && y && norm && gx, "NULL pointer");
  return launch_l2<true>(device, g, ldg, y, ldy, const_cast<float*>(norm), rows, F, eps, gx, ldgx, (cudaStream_t)stream);
}

}  // extern "C"
