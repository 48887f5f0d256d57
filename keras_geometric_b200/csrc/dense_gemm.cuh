// K8: dense node-feature transform X·W on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators,
// TMA operand staging) with fp32-accurate results.
//
// fp32 parity (1e-5 relative) rules out single-pass bf16/tf32 operands, so each fp32 operand tile is
// split IN THE KERNEL (no extra HBM pass) into three bf16 terms and the five most significant product
// bands are accumulated in fp32 in TMEM ("9xBF16" emulation).  The warp-specialised pipeline
// (TMA load -> transform warps -> single-thread tcgen05.mma issue -> TMEM -> epilogue -> TMA store) is
// instantiated from the CuTe/CUTLASS collective templates vendored in the image
// (flashinfer/data/cutlass/include, CUTLASS 4.5); everything around it (layouts, split-K batching,
// epilogue use, C ABI) is ours.  SASS shows UTCHMMA / UTMALDG / UTMASTG / LDTM / STTM.
#pragma once
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"

#include "common.cuh"

namespace kgb {

template <class LayoutA, class LayoutB, class Tile>
struct DenseGemm {
  using Acc = float;
  using LayoutC = cutlass::layout::RowMajor;
  using ClusterShape = cute::Shape<cute::_1, cute::_1, cute::_1>;
  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, Tile, ClusterShape,
      cutlass::epilogue::collective::EpilogueTileAuto, Acc, Acc, float, LayoutC, 4, float, LayoutC, 4,
      cutlass::epilogue::TmaWarpSpecialized1Sm>::CollectiveOp;
  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, float, LayoutA, 4, float, LayoutB, 4, Acc, Tile,
      ClusterShape,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(
          sizeof(typename CollectiveEpilogue::SharedStorage))>,
      cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>::CollectiveOp;
  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<cute::Shape<int, int, int, int>, CollectiveMainloop,
                                                          CollectiveEpilogue, void>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
};

// Generic launcher.  (sa_r, sa_c) etc. are the element strides of the LOGICAL operands
// A(M x K), B(K x N), C/D(M x N); batch strides in elements.
template <class G>
int dense_gemm_launch(const float* A, int64_t lda, int64_t batch_a, const float* B, int64_t ldb, int64_t batch_b,
                      const float* C, float* D, int64_t ldd, int64_t batch_d, int M, int N, int K, int L, float alpha,
                      float beta, void* ws, size_t ws_bytes, size_t* ws_needed, cudaStream_t st) {
  using Gemm = typename G::Gemm;
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sA{};
  StrideB sB{};
  StrideC sC{};
  StrideD sD{};
  // the non-unit stride of each operand is its leading dimension; the batch mode is the last one
  if constexpr (cute::is_same_v<cute::remove_cvref_t<decltype(cute::get<0>(sA))>, cute::Int<1>>) cute::get<1>(sA) = lda;
  else cute::get<0>(sA) = lda;
  if constexpr (cute::is_same_v<cute::remove_cvref_t<decltype(cute::get<0>(sB))>, cute::Int<1>>) cute::get<1>(sB) = ldb;
  else cute::get<0>(sB) = ldb;
  cute::get<0>(sC) = ldd;
  cute::get<0>(sD) = ldd;
  cute::get<2>(sA) = batch_a;
  cute::get<2>(sB) = batch_b;
  cute::get<2>(sC) = batch_d;
  cute::get<2>(sD) = batch_d;
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, L}, {A, sA, B, sB},
                                {{alpha, beta}, C, sC, D, sD}};
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) {
    set_error("dense gemm: shape M=%d N=%d K=%d L=%d not implementable (alignment?)", M, N, K, L);
    return KGB_ERR_UNSUPPORTED;
  }
  const size_t need = Gemm::get_workspace_size(args);
  if (ws_needed) *ws_needed = need;
  if (need > ws_bytes) {
    set_error("dense gemm: workspace too small (need %zu, got %zu)", need, ws_bytes);
    return KGB_ERR_WORKSPACE;
  }
  if (gemm.initialize(args, ws, st) != cutlass::Status::kSuccess) {
    set_error("dense gemm: initialize failed");
    return KGB_ERR_CUDA;
  }
  if (gemm.run(st) != cutlass::Status::kSuccess) {
    set_error("dense gemm: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return KGB_ERR_CUDA;
  }
  count_launch();
  return KGB_OK;
}

using TileWide = cute::Shape<cute::_128, cute::_128, cute::_16>;
using TileNarrow = cute::Shape<cute::_128, cute::_64, cute::_16>;

#define KGB_GEMM_ARGS                                                                                              \
  const float *A, int64_t lda, int64_t batch_a, const float *B, int64_t ldb, int64_t batch_b, const float *C,        \
      float *D, int64_t ldd, int64_t batch_d, int M, int N, int K, int L, float alpha, float beta, void *ws,         \
      size_t ws_bytes, size_t *ws_needed, cudaStream_t st
#define KGB_GEMM_PASS A, lda, batch_a, B, ldb, batch_b, C, D, ldd, batch_d, M, N, K, L, alpha, beta, ws, ws_bytes, ws_needed, st

int dense_gemm_nn(KGB_GEMM_ARGS);  // A row-major [M,K], B row-major [K,N]
int dense_gemm_nt(KGB_GEMM_ARGS);  // A row-major [M,K], B stored [N,K] row-major
int dense_gemm_tn(KGB_GEMM_ARGS);  // A stored [K,M] row-major (i.e. A^T), B row-major [K,N]

}  // namespace kgb
