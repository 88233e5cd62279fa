// gp.cuh -- host-side launchers shared between gp.cu and criteria.cu
#pragma once
#include "tes_common.cuh"

namespace tes {
constexpr int64_t PS_CHUNK = 148 * 128;  // candidates per pass: one 128-row GEMM tile per SM; K and G chunks stay L2-resident

int launch_se_ard(const double* x, int64_t N, const double* xb, int64_t m, int d, const double* l, double sigma,
                  int64_t noise_from, double sigma0, double* out, int64_t ldo, cudaStream_t st);
struct GemmArgs {
    const double* A; int64_t lda; int transa; int64_t batchA;   // op(A): (M x K)
    const double* B; int64_t sbk, sbn, batchB;                  // B'[k][col] = B[k*sbk + col*sbn]
    double* C; int64_t ldc, batchC;
    int64_t M, N, K;
    double alpha, beta;
    int tri;     // skip CTA tiles strictly above the diagonal
    int kskip;   // triangular operands (sbn == 1 only): 1 = both op(A)(row,k) and B'(k,col) vanish for k < row resp. k < col
                 // (X^T X with X lower-triangular): start at k = max(row0, col0); 2 = B'(k,col) = 0 for k < col: start at col0
    int batch;   // grid.z
};
int launch_gemm_ex(const GemmArgs& g, cudaStream_t st);
// C[M x N] = A[M x K] * B, B addressed as B[k*sbk + col*sbn]
int launch_gemm(const double* A, int64_t lda, const double* B, int64_t sbk, int64_t sbn, double* C, int64_t ldc,
                int64_t M, int64_t N, int64_t K, cudaStream_t st);
// out[i] = max(base - sum_c A[i][c]*B[i][c], clip_min)
int launch_rowdot(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t N, int64_t m, double base,
                  double clip_min, double* out, cudaStream_t st);
int launch_concat_rows(const double* a, int64_t na, const double* b, int64_t nb, int d, double* out,
                       cudaStream_t st);
// Bext = [invpNK | invpNK[:, T:] Y]  (m x (m+1))
int launch_build_bext(const double* invpNK, const double* Y, int64_t m, int64_t T, double* Bext, cudaStream_t st);
}  // namespace tes
