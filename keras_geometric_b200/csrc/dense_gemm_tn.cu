// K8 instantiation: see dense_gemm.cuh.  LayoutA = ColumnMajor, LayoutB = RowMajor (CUTLASS tags of the logical
// operands A(MxK), B(KxN)).
#include "dense_gemm.cuh"

namespace kgb {

int dense_gemm_tn(KGB_GEMM_ARGS) {
  using Wide = DenseGemm<cutlass::layout::ColumnMajor, cutlass::layout::RowMajor, TileWide>;
  using Narrow = DenseGemm<cutlass::layout::ColumnMajor, cutlass::layout::RowMajor, TileNarrow>;
  if (N > 64) return dense_gemm_launch<Wide>(KGB_GEMM_PASS);
  return dense_gemm_launch<Narrow>(KGB_GEMM_PASS);
}

}  // namespace kgb
