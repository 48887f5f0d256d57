// Shared helpers for libkgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "kgb200.h"

namespace kgb {

void set_error(const char* fmt, ...);
int use_device(int device);  // cudaSetDevice when the calling thread's device differs
int sm_count(int device);
void count_launch();

#define KGB_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      kgb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return KGB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

// every kernel launch is followed by this macro: counts the launch (kgb_launch_count) and checks it
#define KGB_CHECK_LAUNCH()                 \
  do {                                     \
    kgb::count_launch();                   \
    KGB_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

#define KGB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      kgb::set_error(__VA_ARGS__);      \
      return KGB_ERR_INVALID;           \
    }                                   \
  } while (0)

#define KGB_USE_DEVICE(dev)                    \
  do {                                         \
    int _r = kgb::use_device(dev);             \
    if (_r != KGB_OK) return _r;               \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- device helpers ----------------------------------------------------------------------
template <int VEC>
struct Vec;
template <>
struct Vec<4> { using T = float4; };
template <>
struct Vec<2> { using T = float2; };
template <>
struct Vec<1> { using T = float; };

// read-only 128-bit / 64-bit / 32-bit loads through the non-coherent path
template <int VEC>
__device__ __forceinline__ void ld_vec(const float* __restrict__ p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}

// ---- counter-based random numbers for in-kernel dropout ------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): 4 x 32 random bits from a 128-bit counter and a 64-bit key; stateless, so the
// forward pass and both backward passes regenerate the SAME mask for an (edge id, feature block) without storing it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// 4 keep-bits for block `blk` (4 consecutive features, or 4 consecutive heads) of edge `eid`: bit i set <=> element i
// of the block survives dropout (probability 1 - p, thr = p * 2^32).  `stream` separates independent uses.
__device__ __forceinline__ uint32_t dropout_keep4(uint32_t eid, uint32_t blk, uint32_t stream, uint32_t seed_lo,
                                                  uint32_t seed_hi, uint32_t thr) {
  const uint4 r = philox4x32_10(make_uint4(eid, blk, stream, 0x6b676232u), seed_lo, seed_hi);
  return (r.x >= thr ? 1u : 0u) | (r.y >= thr ? 2u : 0u) | (r.z >= thr ? 4u : 0u) | (r.w >= thr ? 8u : 0u);
}

// L2 eviction-priority policies for per-load cache hints (ld.global.L2::cache_hint)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
#ifdef KGB_HOT_COLD_NORMAL
  asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
#else
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#endif
  return p;
}

// read-only load that tells the L2 how long to keep the line (hot rows: evict_last, read-once rows: evict_first)
template <int VEC>
__device__ __forceinline__ void ld_vec_hint(const float* __restrict__ p, float (&v)[VEC], uint64_t policy) {
  if constexpr (VEC == 4) {
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(policy));
  } else if constexpr (VEC == 2) {
    asm("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v[0]), "=f"(v[1]) : "l"(p), "l"(policy));
  } else {
    asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[0]) : "l"(p), "l"(policy));
  }
}

template <int VEC>
__device__ __forceinline__ void st_vec(float* __restrict__ p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
    p[0] = v[0];
  }
}

template <int VEC>
__device__ __forceinline__ void st_vec_i(int32_t* __restrict__ p, const int32_t (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<int2*>(p) = make_int2(v[0], v[1]);
  } else {
    p[0] = v[0];
  }
}

template <int VEC>
__device__ __forceinline__ void ld_vec_i(const int32_t* __restrict__ p, int32_t (&v)[VEC]) {
  if constexpr (VEC == 4) {
    int4 t = __ldg(reinterpret_cast<const int4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    int2 t = __ldg(reinterpret_cast<const int2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}

__device__ __forceinline__ unsigned group_mask(int lane, int G) {
  return G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// smallest power of two >= v (v >= 1)
static inline int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace kgb
