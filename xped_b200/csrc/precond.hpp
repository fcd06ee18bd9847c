// QR preconditioner of the batched Jacobi SVD (precond.cu): (W, V) <- (A Q_B, Q_B) with A^T = Q_B R_B.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace xb {

struct QrPreBlock
{
    long long w_off, v_off;   // work matrix (m x n, column-major) and accumulator (n x n) of the Jacobi stage
    long long b_off;          // augmented matrix [A^T | 1], n x (m + n), column-major
    long long y_off, t_off;   // Householder panel (<= n x 16) and compact-WY factor (16 x 16)
    long long p_off;          // row order (m entries)
    long long q_off;          // partial row norms (ceil(n / 32) x m entries)
    int m, n;                 // m >= n
};

bool svd_precondition_supported(int max_rows);
void launch_svd_precondition(double* W, double* V, double* Baug, double* Ybuf, double* Tbuf, double* nrm, int* perm, const QrPreBlock* d_blocks,
                             const std::vector<QrPreBlock>& h_blocks, cudaStream_t s, int64_t* launches);

// second step for square blocks (see precond.cu): factorise [W0 | 1], R2 -> Rtmp (offsets q_off); after the caller's GEMM W' = V0 R2^T,
// launch_svd_precondition2_q overwrites V by Q2
void launch_svd_precondition2_factor(const double* W, double* Baug, double* Ybuf, double* Tbuf, double* Rtmp, const QrPreBlock* d_blocks,
                                     const std::vector<QrPreBlock>& h_blocks, cudaStream_t s, int64_t* launches);
void launch_svd_precondition2_q(const double* Baug, double* V, const QrPreBlock* d_blocks, const std::vector<QrPreBlock>& h_blocks, cudaStream_t s, int64_t* launches);

}   // namespace xb
