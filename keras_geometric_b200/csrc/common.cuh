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
