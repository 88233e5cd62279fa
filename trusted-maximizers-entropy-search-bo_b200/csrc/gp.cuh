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

// ---- slice table of the symmetric quadratic forms (criteria.cu; shared with the tensor-core screen in rffprod.cu) ----
struct QSlice {
    int pa, pb, i, diag;
};
__host__ __device__ inline int64_t quad_num_slices(int64_t nb) { return nb * 9 + nb * (nb - 1) / 2 * 16; }

// (a, b) of entry kk of a slice; returns false for the zero padding of slice 8 of a diagonal pair
__device__ __forceinline__ bool quad_slice_ab(const QSlice& q, int kk, int64_t& a, int64_t& b) {
    int al, bl;
    if (!q.diag) {
        al = q.i;
        bl = kk;
    } else if (q.i == 0) {
        al = 0;
        bl = kk;
    } else if (q.i < 8) {
        const bool first = kk < 16 - q.i;
        al = first ? q.i : 16 - q.i;
        bl = first ? q.i + kk : kk;
    } else {
        al = 8;
        bl = 8 + kk;
        if (kk >= 8) return false;
    }
    a = (int64_t)q.pa * 16 + al;
    b = (int64_t)q.pb * 16 + bl;
    return true;
}


// tcgen05 screen of the quadratic forms (rffprod.cu): Q~[x][j] ~ sum_k (M[x,a_k] M[x,b_k]) Bpack[k][j] as a 3xFP16 GEMM.
int64_t quad_screen_kp(int64_t Kdim);                                   // K padded to whole segments
int64_t quad_screen_a_bytes(int64_t nc, int64_t Kdim);                  // operand A of one candidate chunk
int64_t quad_screen_b_bytes(int64_t Kdim);                              // operand B (all maximizers, <= 128)
int launch_quad_screen_build_b(const double* Bpack, int64_t Kdim, int64_t Sp, void* Btc, double* colmax, cudaStream_t st);
int64_t quad_screen_kdim(int64_t T);                                    // K of the screen (run order, see rffprod.cu)
// triangular variant (T <= 128): A = the tile of M, B = row blocks of the packed covariances, dot with M in the epilogue
int64_t quad_tri_b_bytes(int64_t T);
int64_t quad_tri_terms(int64_t T);
int launch_quad_tri_build_b(const double* cov, int64_t Sp, int64_t T, void* Btri, double* colmax, cudaStream_t st);
int launch_quad_screen_tri(const double* G, int64_t ldg, int64_t nc, int64_t T, const void* Btri, const double* colmax,
                           int64_t Sp, double* rowmax, double* Qout, cudaStream_t st);
int launch_quad_screen_ab(const double* cov, int64_t Sp, int64_t T, double* colmax, double* Bpack, cudaStream_t st);
int launch_quad_screen(const double* G, int64_t ldg, int64_t nc, int64_t T, int64_t Kdim, const void* Btc,
                       const double* colmax, int64_t Sp, void* Atc, double* rowmax, double* Qout, cudaStream_t st);
}  // namespace tes
