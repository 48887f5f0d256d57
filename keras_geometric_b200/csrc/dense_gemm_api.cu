// K8 C-ABI entry point (the CUTLASS-templated instantiations live in dense_gemm_{nn,nt,tn}.cu).
#include "common.cuh"

namespace kgb {
#define KGB_GEMM_ARGS                                                                                              \
  const float *A, int64_t lda, int64_t batch_a, const float *B, int64_t ldb, int64_t batch_b, const float *C,        \
      float *D, int64_t ldd, int64_t batch_d, int M, int N, int K, int L, float alpha, float beta, void *ws,         \
      size_t ws_bytes, size_t *ws_needed, cudaStream_t st
int dense_gemm_nn(KGB_GEMM_ARGS);
int dense_gemm_nt(KGB_GEMM_ARGS);
int dense_gemm_tn(KGB_GEMM_ARGS);
}  // namespace kgb

using namespace kgb;

extern "C" {

size_t kgb_dense_gemm_workspace_bytes(int mode, int M, int N, int K, int L) {
  (void)mode; (void)M; (void)N; (void)K; (void)L;
  return (size_t)4 << 20;  // upper bound of what the tile scheduler of these kernels asks for
}

int kgb_dense_gemm(int device, int mode, const float* A, int64_t lda, int64_t batch_a, const float* B, int64_t ldb,
                   int64_t batch_b, const float* C, float* D, int64_t ldd, int64_t batch_d, int M, int N, int K,
                   int L, float alpha, float beta, void* ws, size_t ws_bytes, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(mode >= KGB_GEMM_NN && mode <= KGB_GEMM_TN, "bad gemm mode %d", mode);
  KGB_REQUIRE(M >= 0 && N >= 0 && K >= 0 && L >= 1, "negative size");
  if (M == 0 || N == 0) return KGB_OK;
  KGB_REQUIRE(K > 0, "K must be positive");
  KGB_REQUIRE(A && B && D, "NULL operand");
  KGB_REQUIRE(beta == 0.f || C, "beta != 0 needs C");
  KGB_REQUIRE(aligned16(A) && aligned16(B) && aligned16(D) && (!C || aligned16(C)), "operands must be 16-byte aligned");
  KGB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldd % 4 == 0 && batch_a % 4 == 0 && batch_b % 4 == 0 && batch_d % 4 == 0,
              "leading dimensions / batch strides must be multiples of 4 floats (TMA alignment)");
  const float* Cin = C ? C : D;
  size_t need = 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case KGB_GEMM_NN:
      return dense_gemm_nn(A, lda, batch_a, B, ldb, batch_b, Cin, D, ldd, batch_d, M, N, K, L, alpha, beta, ws, ws_bytes, &need, st);
    case KGB_GEMM_NT:
      return dense_gemm_nt(A, lda, batch_a, B, ldb, batch_b, Cin, D, ldd, batch_d, M, N, K, L, alpha, beta, ws, ws_bytes, &need, st);
    default:
      return dense_gemm_tn(A, lda, batch_a, B, ldb, batch_b, Cin, D, ldd, batch_d, M, N, K, L, alpha, beta, ws, ws_bytes, &need, st);
  }
}

}  // extern "C"
