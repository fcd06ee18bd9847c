// QR preconditioner of the batched one-sided Jacobi SVD (linalg.cu).
//
// One-sided Jacobi applied directly to the CTM projector matrices M = Q4 * Q1 (graded spectra over 10-20 orders of
// magnitude) spends ~20 sweeps in its linear phase.  Following Drmac & Veselic ("New fast and accurate Jacobi SVD
// algorithm", SIAM J. Matrix Anal. Appl. 29, 2008) the work matrix is first brought to triangular form by an
// orthogonal transformation from the right:  with B = A^T = Q_B R_B (Householder QR of the n x m matrix B, m >= n),
//     W0 = A Q_B = R_B^T  (m x n, lower trapezoidal),      V0 = Q_B  (n x n),
// and the Jacobi iteration continues on (W0, V0) instead of (A, 1): its invariant W = A V holds from the start, the
// final gather (U = W / sigma, V) is unchanged, and the sweep count drops from ~30 to ~12 on the CTM matrices.
//
// Device form: Q_B^T is accumulated by factorising the augmented matrix [B | 1] (n x (m + n), column-major) in
// place, so that after the last panel it reads [R_B | Q_B^T] and (W0, V0) is simply its transpose.  Per panel of 16
// columns, for ALL sector blocks at once:
//   qr_panel_kernel<CL>  Householder vectors of the panel, register-resident (a thread owns up to 3 rows x 16
//                        columns), one block-wide reduction per column (all inner products v_j^T a_c of a column step
//                        and the Gram entries needed for the compact-WY factor T are reduced together); a cluster of
//                        CL CTAs splits the rows of tall panels and exchanges the partial sums through DSMEM.
//   qr_update_kernel     trailing update C <- (1 - Y T^T Y^T) C of a 32-column slab per CTA on the FP64 tensor cores
//                        (DMMA m8n8k4): Z = Y^T C (K split over the warps), Z <- T^T Z, C -= Y Z.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "engine.hpp"
#include "precond.hpp"

namespace xb {

namespace {

constexpr int PW = 16;          // panel width
constexpr int QR_THREADS = 256;
constexpr int RPT = 3;          // rows per thread in the panel kernel: a CTA holds up to 768 rows of the panel
constexpr int SLAB = 32;        // trailing-update slab width

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- pack: Baug = [W^T | 1]   /   unpack: W = lower(R_B^T), V = (Q_B^T)^T.  32x32 tiles through shared memory.
// grid = (ceil((m_max + n_max) / 32), ceil(n_max / 32), nblocks), block = (32, 8)
// Row order of the factorisation (de Rijk / pivoting by initial norms): column j of B = A^T is row perm[j] of A, rows sorted by
// decreasing norm.  A row permutation of W = A V changes neither the Gram matrices nor V, so the Jacobi stage is unaffected,
// while the triangular factor becomes graded from the start (two sweeps fewer on the CTM matrices).
__global__ void qr_row_norms_kernel(const double* __restrict__ W, double* __restrict__ nrm, const QrPreBlock* __restrict__ blocks)
{
    const QrPreBlock d = blocks[blockIdx.y];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if(r >= d.m) return;
    double s = 0.0;
    for(int c = 0; c < d.n; ++c) {
        const double v = W[d.w_off + r + static_cast<long long>(c) * d.m];
        s = fma(v, v, s);
    }
    nrm[d.p_off + r] = s;
}

__global__ void qr_rank_rows_kernel(const double* __restrict__ nrm, int* __restrict__ perm, const QrPreBlock* __restrict__ blocks)
{
    const QrPreBlock d = blocks[blockIdx.y];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if(r >= d.m) return;
    const double mine = nrm[d.p_off + r];
    int rank = 0;
    for(int j = 0; j < d.m; ++j) {
        const double o = nrm[d.p_off + j];
        rank += (o > mine || (o == mine && j < r)) ? 1 : 0;
    }
    perm[d.p_off + rank] = r;   // ranks are a permutation: ties are broken by the row index
}

__global__ void qr_pack_kernel(const double* __restrict__ W, double* __restrict__ Baug, const int* __restrict__ perm, const QrPreBlock* __restrict__ blocks)
{
    __shared__ double tile[32][33];
    const QrPreBlock d = blocks[blockIdx.z];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;   // r: stacked row (0..m+n), c: column (0..n)
    if(r0 >= d.m + d.n || c0 >= d.n) return;
    for(int j = threadIdx.y; j < 32; j += 8) {
        const int r = r0 + threadIdx.x, c = c0 + j;
        double v = 0.0;
        if(r < d.m + d.n && c < d.n) v = r < d.m ? W[d.w_off + perm[d.p_off + r] + static_cast<long long>(c) * d.m] : (r - d.m == c ? 1.0 : 0.0);
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for(int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + threadIdx.x, r = r0 + j;
        if(r < d.m + d.n && c < d.n) Baug[d.b_off + c + static_cast<long long>(r) * d.n] = tile[threadIdx.x][j];
    }
}

__global__ void qr_unpack_kernel(const double* __restrict__ Baug, double* __restrict__ W, double* __restrict__ V, const int* __restrict__ perm,
                                 const QrPreBlock* __restrict__ blocks)
{
    __shared__ double tile[32][33];
    const QrPreBlock d = blocks[blockIdx.z];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    if(r0 >= d.m + d.n || c0 >= d.n) return;
    for(int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + threadIdx.x, r = r0 + j;
        double v = 0.0;
        if(r < d.m + d.n && c < d.n) v = Baug[d.b_off + c + static_cast<long long>(r) * d.n];
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for(int j = threadIdx.y; j < 32; j += 8) {
        const int r = r0 + threadIdx.x, c = c0 + j;
        if(r < d.m + d.n && c < d.n) {
            const double v = tile[threadIdx.x][j];
            if(r < d.m)   // R_B is upper trapezoidal (below its diagonal Baug holds stale data); rows go back to their original places
                W[d.w_off + perm[d.p_off + r] + static_cast<long long>(c) * d.m] = r >= c ? v : 0.0;
            else
                V[d.v_off + (r - d.m) + static_cast<long long>(c) * d.n] = v;
        }
    }
}

// ---- panel factorisation ----------------------------------------------------------------------------------------
template <int CL>
__global__ void __launch_bounds__(QR_THREADS)
qr_panel_kernel(double* __restrict__ Baug, double* __restrict__ Ybuf, double* __restrict__ Tbuf, const QrPreBlock* __restrict__ blocks, int k0)
{
    __shared__ double red[QR_THREADS / 32][PW];   // per-warp partial sums of one column step
    __shared__ double xch[2][2 * PW];             // [buffer][ sums (16) | diagonal row (16) ] of this CTA, double buffered over the column steps
    __shared__ double tot[2 * PW];                // cluster-wide sums | diagonal row
    __shared__ double Tsm[PW][PW + 1];
    __shared__ double tau_sm[PW];

    const QrPreBlock d = blocks[blockIdx.x / CL];
    const int n = d.n;
    if(k0 >= n) return;   // uniform over the cluster
    const int h = n - k0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned crank = 0;
    if(CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int per = max(PW, (h + CL - 1) / CL);   // the 16 diagonal rows always sit in CTA 0
    const int lo = static_cast<int>(crank) * per;
    double* __restrict__ P = Baug + d.b_off + k0 + static_cast<long long>(k0) * n;   // panel origin, leading dimension n

    // local row l = lo + tid + 256 i  (relative to the panel origin); grow = k0 + l
    double x[RPT][PW];
    int lrow[RPT];
    bool rv[RPT];
#pragma unroll
    for(int i = 0; i < RPT; ++i) {
        lrow[i] = lo + tid + QR_THREADS * i;
        rv[i] = (tid + QR_THREADS * i) < per && lrow[i] < h;
#pragma unroll
        for(int c = 0; c < PW; ++c) x[i][c] = rv[i] ? P[lrow[i] + static_cast<long long>(c) * n] : 0.0;   // m >= 16: columns k0 .. k0+15 always exist in [B | 1]
    }
    for(int e = tid; e < PW * (PW + 1); e += QR_THREADS) (&Tsm[0][0])[e] = 0.0;
    if(tid < PW) tau_sm[tid] = 0.0;
    __syncthreads();

    // One reduction round: slot c >= j carries s_c = sum_{rows below the diagonal of column j} x[.][j] x[.][c];
    // slot a < j - 1 carries z_a = Y[:, a]^T Y[:, j-1] (columns < j are already in their final Y form).
    auto reduce_round = [&](double (&ps)[PW], int buf, bool with_diag_row, int j) {
#pragma unroll
        for(int c = 0; c < PW; ++c) {
            double v = ps[c];
#pragma unroll
            for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            ps[c] = v;
        }
        if(lane == 0) {
#pragma unroll
            for(int c = 0; c < PW; ++c) red[warp][c] = ps[c];
        }
        if(with_diag_row && crank == 0 && tid == j) {   // the owner of the diagonal row publishes it
#pragma unroll
            for(int c = 0; c < PW; ++c) xch[buf][PW + c] = x[0][c];
        }
        __syncthreads();
        if(tid < PW) {
            double s = 0.0;
#pragma unroll
            for(int w = 0; w < QR_THREADS / 32; ++w) s += red[w][tid];
            xch[buf][tid] = s;
        }
        if(CL > 1) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
            if(tid < 2 * PW) {
                double s = 0.0;
                const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(&xch[buf][tid]));
                if(tid < PW) {
                    double part[CL];
#pragma unroll
                    for(int cr = 0; cr < CL; ++cr) {
                        unsigned remote;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cr));
                        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(part[cr]) : "r"(remote));
                    }
#pragma unroll
                    for(int cr = 0; cr < CL; ++cr) s += part[cr];   // fixed order: identical in every CTA
                } else {
                    unsigned remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
                    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(s) : "r"(remote));
                }
                tot[tid] = s;
            }
        } else {
            __syncthreads();
            if(tid < 2 * PW) tot[tid] = xch[buf][tid];
        }
        __syncthreads();
    };

#pragma unroll
    for(int j = 0; j < PW; ++j) {
        const int gd = j;   // diagonal row of column j, relative to the panel origin
        double ps[PW];
#pragma unroll
        for(int c = 0; c < PW; ++c) ps[c] = 0.0;
#pragma unroll
        for(int i = 0; i < RPT; ++i) {
            const bool below = rv[i] && lrow[i] > gd;
#pragma unroll
            for(int c = j; c < PW; ++c) ps[c] = fma(below ? x[i][j] : 0.0, x[i][c], ps[c]);
            if(j >= 2) {
#pragma unroll
                for(int a = 0; a < j - 1; ++a) ps[a] = fma(x[i][a], x[i][j - 1], ps[a]);
            }
        }
        reduce_round(ps, j & 1, true, j);

        // T column j-1 (compact WY, forward columnwise):  T[0:j-1, j-1] = -tau_{j-1} T[0:j-1, 0:j-1] z
        if(j >= 2 && crank == 0 && tid < j - 1) {
            double s = 0.0;
            for(int l = tid; l < j - 1; ++l) s = fma(Tsm[tid][l], tot[l], s);
            Tsm[tid][j - 1] = -tau_sm[j - 1] * s;
        }
        const double sigma2 = tot[j], alpha = tot[PW + j];
        double beta = alpha, tau = 0.0, scale = 0.0;
        if(sigma2 > 0.0) {   // dlarfg; sigma2 == 0  =>  H = 1
            beta = -copysign(sqrt(fma(alpha, alpha, sigma2)), alpha);
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        double f[PW];
#pragma unroll
        for(int c = j + 1; c < PW; ++c) f[c] = tau * fma(scale, tot[c], tot[PW + c]);   // tau * v^T a_c
#pragma unroll
        for(int i = 0; i < RPT; ++i) {
            if(rv[i] && lrow[i] > gd) {
                const double v = x[i][j] * scale;
#pragma unroll
                for(int c = j + 1; c < PW; ++c) x[i][c] = fma(-f[c], v, x[i][c]);
                x[i][j] = v;
            } else if(rv[i] && lrow[i] == gd) {   // the diagonal row: finished row j of R
                P[gd + static_cast<long long>(j) * n] = beta;
#pragma unroll
                for(int c = j + 1; c < PW; ++c) {
                    x[i][c] -= f[c];
                    P[gd + static_cast<long long>(c) * n] = x[i][c];
                }
                x[i][j] = 1.0;
            } else {
                x[i][j] = 0.0;   // rows above the diagonal (their R entries were stored when their row was finished)
            }
        }
        if(crank == 0 && tid == 0) {
            tau_sm[j] = tau;
            Tsm[j][j] = tau;
        }
        __syncthreads();
    }
    {   // Gram entries of the last column, then T column 15
        double ps[PW];
#pragma unroll
        for(int c = 0; c < PW; ++c) ps[c] = 0.0;
#pragma unroll
        for(int i = 0; i < RPT; ++i)
#pragma unroll
            for(int a = 0; a < PW - 1; ++a) ps[a] = fma(x[i][a], x[i][PW - 1], ps[a]);
        reduce_round(ps, 0, false, 0);
        if(crank == 0 && tid < PW - 1) {
            double s = 0.0;
            for(int l = tid; l < PW - 1; ++l) s = fma(Tsm[tid][l], tot[l], s);
            Tsm[tid][PW - 1] = -tau_sm[PW - 1] * s;
        }
        __syncthreads();
    }
    // Y (h x 16, column-major, leading dimension h) and T (16 x 16, row-major, upper triangular)
    double* __restrict__ Y = Ybuf + d.y_off;
#pragma unroll
    for(int i = 0; i < RPT; ++i)
        if(rv[i]) {
#pragma unroll
            for(int c = 0; c < PW; ++c) Y[lrow[i] + static_cast<long long>(c) * h] = x[i][c];
        }
    if(crank == 0) {
        for(int e = tid; e < PW * PW; e += QR_THREADS) Tbuf[d.t_off + e] = Tsm[e / PW][e % PW];
    }
    if(CL > 1) {   // nobody leaves while a peer may still read its xch
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
}

// ---- trailing update: C <- C - Y (T^T (Y^T C)) for one 32-column slab -----------------------------------------------
// grid = (max slabs, nblocks)
__global__ void __launch_bounds__(QR_THREADS)
qr_update_kernel(double* __restrict__ Baug, const double* __restrict__ Ybuf, const double* __restrict__ Tbuf, const QrPreBlock* __restrict__ blocks, int k0)
{
    __shared__ double Zs[QR_THREADS / 32][PW][SLAB + 1];
    __shared__ double Zr[PW][SLAB + 1];
    __shared__ double Zf[PW][SLAB + 4];
    __shared__ double Ts[PW][PW + 1];

    const QrPreBlock d = blocks[blockIdx.y];
    const int n = d.n;
    if(k0 >= n) return;
    const int h = n - k0, ncols = d.m + n;
    const int col0 = k0 + PW + SLAB * blockIdx.x;
    if(col0 >= ncols) return;
    const int nc = min(SLAB, ncols - col0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
    double* __restrict__ C = Baug + d.b_off + k0 + static_cast<long long>(col0) * n;   // h x nc, leading dimension n
    const double* __restrict__ Y = Ybuf + d.y_off;                                      // h x 16, leading dimension h
    for(int e = tid; e < PW * PW; e += QR_THREADS) Ts[e / PW][e % PW] = Tbuf[d.t_off + e];

    // ---- pass A: Z = Y^T C, the rows (K) split over the warps
    {
        const int kw = (((h + 7) / 8) + 3) & ~3;   // rows per warp, multiple of 4
        const int kb = warp * kw, ke = min(h, kb + kw);
        double acc[2][4][2];
#pragma unroll
        for(int it = 0; it < 2; ++it)
#pragma unroll
            for(int jt = 0; jt < 4; ++jt) acc[it][jt][0] = acc[it][jt][1] = 0.0;
#pragma unroll 2
        for(int k = kb; k < ke; k += 4) {
            const int kk = k + fk;
            const bool kv = kk < ke;
            const double a0 = kv ? __ldg(Y + kk + static_cast<long long>(fr) * h) : 0.0;
            const double a1 = kv ? __ldg(Y + kk + static_cast<long long>(8 + fr) * h) : 0.0;
            double b[4];
#pragma unroll
            for(int jt = 0; jt < 4; ++jt) b[jt] = (kv && jt * 8 + fr < nc) ? C[kk + static_cast<long long>(jt * 8 + fr) * n] : 0.0;
#pragma unroll
            for(int jt = 0; jt < 4; ++jt) {
                dmma(acc[0][jt][0], acc[0][jt][1], a0, b[jt]);
                dmma(acc[1][jt][0], acc[1][jt][1], a1, b[jt]);
            }
        }
#pragma unroll
        for(int it = 0; it < 2; ++it)
#pragma unroll
            for(int jt = 0; jt < 4; ++jt) {
                Zs[warp][it * 8 + fr][jt * 8 + 2 * fk] = acc[it][jt][0];
                Zs[warp][it * 8 + fr][jt * 8 + 2 * fk + 1] = acc[it][jt][1];
            }
    }
    __syncthreads();
    for(int e = tid; e < PW * SLAB; e += QR_THREADS) {
        const int i = e / SLAB, c = e % SLAB;
        double s = 0.0;
#pragma unroll
        for(int w = 0; w < QR_THREADS / 32; ++w) s += Zs[w][i][c];   // fixed order
        Zr[i][c] = s;
    }
    __syncthreads();
    for(int e = tid; e < PW * SLAB; e += QR_THREADS) {   // Z <- T^T Z  (T upper triangular)
        const int a = e / SLAB, c = e % SLAB;
        double s = 0.0;
        for(int b = 0; b <= a; ++b) s = fma(Ts[b][a], Zr[b][c], s);
        Zf[a][c] = s;
    }
    __syncthreads();
    // ---- pass B: C -= Y Z, 8-row groups over the warps
    const int ngroups = (h + 7) >> 3;
    for(int g = warp; g < ngroups; g += QR_THREADS / 32) {
        double out[4][2];
#pragma unroll
        for(int jt = 0; jt < 4; ++jt) out[jt][0] = out[jt][1] = 0.0;
        const int row = g * 8 + fr;
#pragma unroll
        for(int kq = 0; kq < PW; kq += 4) {
            const double a = row < h ? __ldg(Y + row + static_cast<long long>(kq + fk) * h) : 0.0;
#pragma unroll
            for(int jt = 0; jt < 4; ++jt) dmma(out[jt][0], out[jt][1], a, Zf[kq + fk][jt * 8 + fr]);
        }
        if(row < h) {
#pragma unroll
            for(int jt = 0; jt < 4; ++jt)
#pragma unroll
                for(int hh = 0; hh < 2; ++hh) {
                    const int c = jt * 8 + 2 * fk + hh;
                    if(c < nc) {
                        double* p = C + row + static_cast<long long>(c) * n;
                        *p -= out[jt][hh];
                    }
                }
        }
    }
}

}   // namespace

bool svd_precondition_supported(int max_rows) { return max_rows <= RPT * QR_THREADS * 4; }

void launch_svd_precondition(double* W, double* V, double* Baug, double* Ybuf, double* Tbuf, double* nrm, int* perm, const QrPreBlock* d_blocks,
                             const std::vector<QrPreBlock>& h_blocks, cudaStream_t s, int64_t* launches)
{
    const int nb = static_cast<int>(h_blocks.size());
    if(nb == 0) return;
    int n_max = 0, mn_max = 0;
    for(const QrPreBlock& b : h_blocks) {
        n_max = std::max(n_max, b.n);
        mn_max = std::max(mn_max, b.m + b.n);
    }
    const dim3 tgrid(static_cast<unsigned>((mn_max + 31) / 32), static_cast<unsigned>((n_max + 31) / 32), static_cast<unsigned>(nb)), tblock(32, 8);
    int m_max = 0;
    for(const QrPreBlock& b : h_blocks) m_max = std::max(m_max, b.m);
    const dim3 rgrid(static_cast<unsigned>((m_max + 127) / 128), static_cast<unsigned>(nb));
    qr_row_norms_kernel<<<rgrid, 128, 0, s>>>(W, nrm, d_blocks);
    qr_rank_rows_kernel<<<rgrid, 128, 0, s>>>(nrm, perm, d_blocks);
    qr_pack_kernel<<<tgrid, tblock, 0, s>>>(W, Baug, perm, d_blocks);
    check_cuda(cudaGetLastError(), "qr_pack_kernel launch");
    *launches += 3;
    const int cluster = n_max <= RPT * QR_THREADS ? 1 : 4;
    for(int k0 = 0; k0 < n_max; k0 += PW) {
        if(cluster == 1) {
            qr_panel_kernel<1><<<nb, QR_THREADS, 0, s>>>(Baug, Ybuf, Tbuf, d_blocks, k0);
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(static_cast<unsigned>(nb) * cluster);
            cfg.blockDim = dim3(QR_THREADS);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cluster;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            check_cuda(cudaLaunchKernelEx(&cfg, qr_panel_kernel<4>, Baug, Ybuf, Tbuf, d_blocks, k0), "qr_panel_kernel<4> launch");
        }
        check_cuda(cudaGetLastError(), "qr_panel_kernel launch");
        const int slabs = (mn_max - k0 - PW + SLAB - 1) / SLAB;
        if(slabs > 0) {
            qr_update_kernel<<<dim3(static_cast<unsigned>(slabs), static_cast<unsigned>(nb)), QR_THREADS, 0, s>>>(Baug, Ybuf, Tbuf, d_blocks, k0);
            check_cuda(cudaGetLastError(), "qr_update_kernel launch");
        }
        *launches += 2;
    }
    qr_unpack_kernel<<<tgrid, tblock, 0, s>>>(Baug, W, V, perm, d_blocks);
    check_cuda(cudaGetLastError(), "qr_unpack_kernel launch");
    ++*launches;
}

}   // namespace xb
