// Internal interface between the batched Jacobi SVD (linalg.cu) and its callers (tSVD drivers in linalg.cu / rsvd.cu).
#pragma once
#include <string>
#include <vector>

#include "engine.hpp"

namespace xb {

struct SvdBlock
{
    long long w_off;   // work matrix, rows x cols, column-major, ld = rows
    long long v_off;   // accumulator, cols x cols
    long long s_off;   // singular values (cols)
    int rows, cols;
    double tol;
};

struct JacobiResult
{
    std::vector<SvdBlock> blocks;
    std::vector<int> transposed;
    std::vector<double> sigma;   // host copy, per block at s_off (only when requested)
    std::vector<double> shift;   // host copy of the per-block diagonal shift (teigh only)
    double* d_W = nullptr;
    double* d_V = nullptr;
    double* d_sigma = nullptr;   // device copy of the singular values (column norms of W), per block at s_off
    long long n_sigma = 0;
    int sweeps = 0;
};

// the matrices to decompose: column-major blocks (ld = rows) inside one device buffer
struct JacobiInput
{
    const double* d = nullptr;
    std::vector<long long> off;
    std::vector<int> rows, cols;
    std::string sig;   // key of the cached sweep graphs
    int sweeps_hint = 0;            // expected number of sweeps (0: unknown): size of the first batch enqueued without a convergence read-back
    bool no_precondition = false;   // skip the QR preconditioner (projected blocks of the truncated path: columns already close to the singular directions)
};

JacobiInput jacobi_input(const Tensor& t);
// One-sided block Jacobi on every block of `in`: leaves W (= sigma_j u_j columns, tall side first), V and the singular values on the device
JacobiResult jacobi_all(Context& ctx, const JacobiInput& in, bool shifted, bool host_sigma);
void free_jacobi(Context& ctx, JacobiResult& R);
// device-wide sort / truncation (rsvd.cu): capacity check and the full path's finishing step (select + gather of U, Sigma, V^dagger)
bool svd_device_select_supported(long long n_total, int nsec);
SvdResult svd_full_device_select(Context& ctx, const Tensor& t, JacobiResult& J, int64_t max_keep, double eps_svd, bool preserve_multiplets);

}   // namespace xb
