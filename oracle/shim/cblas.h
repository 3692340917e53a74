/* Minimal cblas.h shim so that /root/reference/src/mat_mul.cu compiles without OpenBLAS.
 * TEST INFRASTRUCTURE ONLY (oracle/): the reference links -lopenblas (Makefile:5, version
 * unpinned, not vendored).  The arithmetic contract of the three calls the reference makes
 * (mat_mul.cu:35,54,67,79) is row-major sgemv / sgemm; cblas_naive.c implements exactly that
 * with a sequential k-loop (same order as the reference's own mat_mul_simple, mat_mul.cu:17-26).
 */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int M, int N, float alpha,
                 const float* A, int lda, const float* X, int incX, float beta, float* Y, int incY);
void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE transA, CBLAS_TRANSPOSE transB, int M, int N,
                 int K, float alpha, const float* A, int lda, const float* B, int ldb, float beta,
                 float* C, int ldc);
void openblas_set_num_threads(int n);
#ifdef __cplusplus
}
#endif
#endif
