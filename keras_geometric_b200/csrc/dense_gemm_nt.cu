// K8 instantiation: see dense_gemm.cuh.  LayoutA = RowMajor, LayoutB = ColumnMajor (CUTLASS tags of the logical
// operands A(MxK), B(KxN)).
#include "dense_gemm.cuh"

namespace kgb {

int dense_gemm_nt(KGB_GEMM_ARGS) {
  using Wide = DenseGemm<cutlass::layout::RowMajor, cutlass::layout::ColumnMajor, TileWide>;
  using Narrow = DenseGemm<cutlass::layout::RowMajor, cutlass::layout::ColumnMajor, TileNarrow>;
  if (N > 64) return dense_gemm_launch<Wide>(KGB_GEMM_PASS);
  return dense_gemm_launch<Narrow>(KGB_GEMM_PASS);
}

}  // namespace kgb
