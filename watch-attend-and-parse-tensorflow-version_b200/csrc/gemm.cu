// Plain-GEMM front ends over the generic core in gemm.cuh.
#include "gemm.cuh"

namespace wap {

int gemm_nn(cudaStream_t st, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            int M, int N, int K, const float* bias, bool accumulate) {
  WAP_REQUIRE((K & 3) == 0 && (lda & 3) == 0, "gemm_nn: K=%d lda=%d must be multiples of 4", K, lda);
  ARowMajor a{A, lda, M, K};
  BRowMajor b{B, ldb, K, N};
  EpiStore e{C, ldc, bias, accumulate, nullptr};
  if (M <= 64 || N <= 64) return launch_gemm<Cfg64x64>(st, a, b, e, M, N, K);
  return launch_gemm<Cfg128x128>(st, a, b, e, M, N, K);
}

int gemm_nt(cudaStream_t st, const float* A, int lda, const float* W, int ldw, float* C, int ldc,
            int M, int N, int K, const float* bias, bool accumulate) {
  WAP_REQUIRE((K & 3) == 0 && (lda & 3) == 0 && (ldw & 3) == 0, "gemm_nt: K=%d lda=%d ldw=%d must be multiples of 4", K, lda, ldw);
  ARowMajor a{A, lda, M, K};
  BColMajor b{W, ldw, K, N};
  EpiStore e{C, ldc, bias, accumulate, nullptr};
  if (M <= 64 || N <= 64) return launch_gemm<Cfg64x64>(st, a, b, e, M, N, K);
  return launch_gemm<Cfg128x128>(st, a, b, e, M, N, K);
}

int gemm_tn_atomic(cudaStream_t st, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                   int M, int N, int K) {
  WAP_REQUIRE((M & 3) == 0 && (lda & 3) == 0, "gemm_tn: M=%d lda=%d must be multiples of 4", M, lda);
  AColMajor a{A, lda, M, K};
  BRowMajor b{B, ldb, K, N};
  EpiAtomic e{C, ldc};
  // enough K-splits to fill the machine: ~4 waves of 148 SMs
  if (M <= 64 || N <= 64) {
    const int tiles = cdiv(M, 64) * cdiv(N, 64);
    int splits = max(1, min(cdiv(K, 4 * BK), cdiv(592, tiles)));
    return launch_gemm<Cfg64x64>(st, a, b, e, M, N, K, splits);
  }
  const int tiles = cdiv(M, 128) * cdiv(N, 128);
  int splits = max(1, min(cdiv(K, 4 * BK), cdiv(592, tiles)));
  return launch_gemm<Cfg128x128>(st, a, b, e, M, N, K, splits);
}

}  // namespace wap
