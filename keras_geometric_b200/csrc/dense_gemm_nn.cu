// K8 instantiation: see dense_gemm.cuh.  LayoutA = RowMajor, LayoutB = RowMajor (CUTLASS tags of the logical
// operands A(MxK), B(KxN)).
#include "dense_gemm.cuh"

namespace kgb {

int dense_gemm_nn(KGB_GEMM_ARGS) {
  using Wide = DenseGemm<cutlass::layout::RowMajor, cutlass::layout::RowMajor, TileWide>;
  using Narrow = DenseGemm<cutlass::layout::RowMajor, cutlass::layout::RowMajor, TileNarrow>;
  if (N > 64) return dense_gemm_launch<Wide>(KGB_GEMM_PASS);
  return dense_gemm_launch<Narrow>(KGB_GEMM_PASS);
}

}  // namespace kgb
