/* Row-major sgemv/sgemm behind the cblas.h shim (see cblas.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Every C[i][j] is a sequential fp32 sum over k starting from 0 (the order of the reference's
 * own mat_mul_simple, /root/reference/src/mat_mul.cu:17-26), then C = alpha*acc + beta*C.
 * The loops are arranged i,k,j (B materialised K x N) so gcc vectorises across j WITHOUT
 * reassociating the k-sum: results are bit-identical to the textbook triple loop and
 * reproducible across hosts (-ffp-contract=off in the Makefile), but a few times faster, which
 * matters because the same shim backs the CPU baseline. */
#include "cblas.h"
#include <stdlib.h>
#include <string.h>

void openblas_set_num_threads(int n) { (void)n; }

void cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int M, int N, float alpha,
                 const float* A, int lda, const float* X, int incX, float beta, float* Y, int incY) {
    (void)order;
    if (trans == CblasNoTrans) {
        for (int i = 0; i < M; i++) {
            float acc = 0.0f;
            for (int j = 0; j < N; j++) acc += A[i * lda + j] * X[j * incX];
            Y[i * incY] = alpha * acc + beta * Y[i * incY];
        }
    } else {
        for (int j = 0; j < N; j++) {
            float acc = 0.0f;
            for (int i = 0; i < M; i++) acc += A[i * lda + j] * X[i * incX];
            Y[j * incY] = alpha * acc + beta * Y[j * incY];
        }
    }
}

void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE transA, CBLAS_TRANSPOSE transB, int M, int N,
                 int K, float alpha, const float* A, int lda, const float* B, int ldb, float beta,
                 float* C, int ldc) {
    (void)order;
    /* Bt: K x N row-major view of op(B). */
    const float* Bk = B;
    int ldbk = ldb;
    float* tmpB = NULL;
    if (transB != CblasNoTrans) {
        tmpB = (float*)malloc((size_t)K * N * sizeof(float));
        for (int j = 0; j < N; j++)
            for (int k = 0; k < K; k++) tmpB[(size_t)k * N + j] = B[(size_t)j * ldb + k];
        Bk = tmpB;
        ldbk = N;
    }
    float* acc = (float*)malloc((size_t)N * sizeof(float));
    for (int i = 0; i < M; i++) {
        memset(acc, 0, (size_t)N * sizeof(float));
        for (int k = 0; k < K; k++) {
            const float a = (transA == CblasNoTrans) ? A[(size_t)i * lda + k] : A[(size_t)k * lda + i];
            const float* brow = Bk + (size_t)k * ldbk;
            for (int j = 0; j < N; j++) acc[j] += a * brow[j];
        }
        float* crow = C + (size_t)i * ldc;
        for (int j = 0; j < N; j++) crow[j] = alpha * acc[j] + beta * crow[j];
    }
    free(acc);
    free(tmpB);
}
