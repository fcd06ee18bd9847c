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
#include <cstdio>
#include <cstdlib>
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
constexpr int NORM_CHUNK = 32;   // columns per partial sum
__global__ void qr_row_norms_kernel(const double* __restrict__ W, double* __restrict__ part, const QrPreBlock* __restrict__ blocks)
{
    // part[chunk][row] = sum over the chunk's columns of W[row][c]^2   (grid: rows/128 x chunks x blocks)
    const QrPreBlock d = blocks[blockIdx.z];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.y * NORM_CHUNK;
    if(r >= d.m || c0 >= d.n) return;
    double s = 0.0;
    const int c1 = min(d.n, c0 + NORM_CHUNK);
    for(int c = c0; c < c1; ++c) {
        const double v = W[d.w_off + r + static_cast<long long>(c) * d.m];
        s = fma(v, v, s);
    }
    part[d.q_off + static_cast<long long>(blockIdx.y) * d.m + r] = s;
}

__global__ void qr_rank_rows_kernel(const double* __restrict__ part, int* __restrict__ perm, const QrPreBlock* __restrict__ blocks)
{
    extern __shared__ double nrm[];   // all row norms of the block (every CTA reduces the partial sums in the same fixed order)
    const QrPreBlock d = blocks[blockIdx.y];
    const int nchunk = (d.n + NORM_CHUNK - 1) / NORM_CHUNK;
    for(int r = threadIdx.x; r < d.m; r += blockDim.x) {
        double s = 0.0;
        for(int k = 0; k < nchunk; ++k) s += part[d.q_off + static_cast<long long>(k) * d.m + r];
        nrm[r] = s;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if(r >= d.m) return;
    const double mine = nrm[r];
    int rank = 0;
    for(int j = 0; j < d.m; ++j) {
        const double o = nrm[j];
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
// The column loop is NOT unrolled (an unrolled 16-step body overflows the instruction cache: 64 us per panel measured):
// after every step the register columns are shifted by one so that the active column is always register column 0.
template <int CL>
__global__ void __launch_bounds__(QR_THREADS)
qr_panel_kernel(double* __restrict__ Baug, double* __restrict__ Ybuf, double* __restrict__ Tbuf, const QrPreBlock* __restrict__ blocks, int k0,
                unsigned long long* __restrict__ tprof)
{
    constexpr int NW = QR_THREADS / 32;
    const bool prof = tprof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long tp = prof ? clock64() : 0;
    auto lap = [&](int slot) {
        if(prof) {
            const long long now = clock64();
            atomicAdd(tprof + slot, static_cast<unsigned long long>(now - tp));
            tp = now;
        }
    };
    __shared__ double red[NW][PW];            // per-warp partial sums of one column step
    __shared__ double xch[2][2 * PW];         // [buffer][ sums (16) | diagonal row (16) ] of this CTA, double buffered over the column steps
    __shared__ double tot[2 * PW];            // cluster-wide sums | diagonal row
    __shared__ double gpart[NW][PW][PW + 1];  // per-warp partial Gram matrices of Y
    __shared__ double gx[PW * PW];            // this CTA's Gram matrix of Y (exchanged over the cluster)
    __shared__ double GY[PW][PW + 1];
    __shared__ double Tsm[PW][PW + 1];
    __shared__ double tau_sm[PW];

    const QrPreBlock d = blocks[blockIdx.x / CL];
    const int n = d.n;
    if(k0 >= n) return;   // uniform over the cluster
    const int h = n - k0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
    unsigned crank = 0;
    if(CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int per = max(PW, (h + CL - 1) / CL);   // the 16 diagonal rows always sit in CTA 0
    const int lo = static_cast<int>(crank) * per;
    double* __restrict__ P = Baug + d.b_off + k0 + static_cast<long long>(k0) * n;   // panel origin, leading dimension n
    double* __restrict__ Y = Ybuf + d.y_off;                                          // h x 16, leading dimension h

    // local row l = lo + tid + 256 i  (relative to the panel origin)
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
    __syncthreads();
    lap(0);

#pragma unroll 1
    for(int j = 0; j < PW; ++j) {
        // register column c holds panel column j + c; slot c: s_c = sum_{rows below the diagonal} x[.][0] x[.][c]  (slot 0: sigma^2)
        double ps[PW];
#pragma unroll
        for(int c = 0; c < PW; ++c) ps[c] = 0.0;
#pragma unroll
        for(int i = 0; i < RPT; ++i) {
            const double xa = (rv[i] && lrow[i] > j) ? x[i][0] : 0.0;
#pragma unroll
            for(int c = 0; c < PW; ++c) ps[c] = fma(xa, x[i][c], ps[c]);
        }
        // transposing butterfly: 16 values over 32 lanes in 8 + 4 + 2 + 1 + 1 shuffles (instead of 16 x 5); after the four
        // halving steps lane l holds the sum over its 16-lane class of value ((l >> 4) & 1) * 8 + ((l >> 3) & 1) * 4 + ...,
        // the last step adds the two classes
#pragma unroll
        for(int k = 0; k < 8; ++k) {
            const bool up = (lane & 16) != 0;
            const double send = up ? ps[k] : ps[k + 8], keep = up ? ps[k + 8] : ps[k];
            ps[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            const bool up = (lane & 8) != 0;
            const double send = up ? ps[k] : ps[k + 4], keep = up ? ps[k + 4] : ps[k];
            ps[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for(int k = 0; k < 2; ++k) {
            const bool up = (lane & 4) != 0;
            const double send = up ? ps[k] : ps[k + 2], keep = up ? ps[k + 2] : ps[k];
            ps[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        {
            const bool up = (lane & 2) != 0;
            const double send = up ? ps[0] : ps[1], keep = up ? ps[1] : ps[0];
            ps[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        ps[0] += __shfl_xor_sync(0xffffffffu, ps[0], 1);
        const int buf = j & 1;
        if((lane & 1) == 0) {   // lane bits (16, 8, 4, 2) select the value index bits (8, 4, 2, 1)
            const int c = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            red[warp][c] = ps[0];
        }
        if(crank == 0 && tid == j) {   // the owner of the diagonal row publishes it
#pragma unroll
            for(int c = 0; c < PW; ++c) xch[buf][PW + c] = x[0][c];
        }
        __syncthreads();
        if(tid < PW) {
            double sum = 0.0;
#pragma unroll
            for(int w = 0; w < NW; ++w) sum += red[w][tid];
            xch[buf][tid] = sum;
        }
        if(CL > 1) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
            if(tid < 2 * PW) {
                double sum = 0.0;
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
                    for(int cr = 0; cr < CL; ++cr) sum += part[cr];   // fixed order: identical in every CTA
                } else {
                    unsigned remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
                    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(sum) : "r"(remote));
                }
                tot[tid] = sum;
            }
        } else {
            __syncthreads();
            if(tid < 2 * PW) tot[tid] = xch[buf][tid];
        }
        __syncthreads();

        lap(1);
        const double sigma2 = tot[0], alpha = tot[PW];
        double beta = alpha, tau = 0.0, scale = 0.0;
        if(sigma2 > 0.0) {   // dlarfg; sigma2 == 0  =>  H = 1.  One reciprocal square root and one division on the chain:
            const double nrm2 = fma(alpha, alpha, sigma2), rn = rsqrt(nrm2), nrm = nrm2 * rn;   // nrm = ||(alpha, x)||
            beta = -copysign(nrm, alpha);
            tau = fma(fabs(alpha), rn, 1.0);                  // (beta - alpha) / beta = 1 + |alpha| / nrm
            scale = copysign(1.0 / (fabs(alpha) + nrm), alpha);   // 1 / (alpha - beta)
        }
        double f[PW];
#pragma unroll
        for(int c = 1; c < PW; ++c) f[c] = tau * fma(scale, tot[c], tot[PW + c]);   // tau * v^T a_c
#pragma unroll
        for(int i = 0; i < RPT; ++i) {
            double yv = 0.0;   // rows above the diagonal (their R entries were stored when their row was finished)
            if(rv[i] && lrow[i] > j) {
                yv = x[i][0] * scale;
#pragma unroll
                for(int c = 1; c < PW; ++c) x[i][c] = fma(-f[c], yv, x[i][c]);
            } else if(rv[i] && lrow[i] == j) {   // the diagonal row: finished row j of R
                P[j + static_cast<long long>(j) * n] = beta;
#pragma unroll
                for(int c = 1; c < PW; ++c) {
                    x[i][c] -= f[c];
                    if(j + c < PW) P[j + static_cast<long long>(j + c) * n] = x[i][c];
                }
                yv = 1.0;
            }
            if(rv[i]) Y[lrow[i] + static_cast<long long>(j) * h] = yv;
#pragma unroll
            for(int c = 1; c < PW; ++c) x[i][c - 1] = x[i][c];   // shift: the next active column becomes register column 0
            x[i][PW - 1] = 0.0;
        }
        if(crank == 0 && tid == 0) {
            tau_sm[j] = tau;
            Tsm[j][j] = tau;
        }
        __syncthreads();
        lap(2);
    }

    // ---- Gram matrix of Y (this CTA's rows) on DMMA, K split over the warps; then the compact-WY factor T
    {
        const int hi = min(h, lo + per);
        const int nrows = max(0, hi - lo);
        const int kw = (((nrows + NW - 1) / NW) + 3) & ~3;
        const int kb = lo + warp * kw, ke = min(hi, kb + kw);
        double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
        for(int k = kb; k < ke; k += 4) {
            const int kk = k + fk;
            const bool kv = kk < ke;
            const double a0 = kv ? Y[kk + static_cast<long long>(fr) * h] : 0.0;
            const double a1 = kv ? Y[kk + static_cast<long long>(8 + fr) * h] : 0.0;
            dmma(acc[0][0][0], acc[0][0][1], a0, a0);
            dmma(acc[0][1][0], acc[0][1][1], a0, a1);
            dmma(acc[1][0][0], acc[1][0][1], a1, a0);
            dmma(acc[1][1][0], acc[1][1][1], a1, a1);
        }
#pragma unroll
        for(int it = 0; it < 2; ++it)
#pragma unroll
            for(int jt = 0; jt < 2; ++jt) {
                gpart[warp][it * 8 + fr][jt * 8 + 2 * fk] = acc[it][jt][0];
                gpart[warp][it * 8 + fr][jt * 8 + 2 * fk + 1] = acc[it][jt][1];
            }
    }
    __syncthreads();
    {
        double sum = 0.0;
#pragma unroll
        for(int w = 0; w < NW; ++w) sum += gpart[w][tid / PW][tid % PW];   // 256 threads: one entry each, fixed order
        gx[tid] = sum;
    }
    if(CL > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        double part[CL];
        const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(&gx[tid]));
#pragma unroll
        for(int cr = 0; cr < CL; ++cr) {
            unsigned remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cr));
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(part[cr]) : "r"(remote));
        }
        double sum = 0.0;
#pragma unroll
        for(int cr = 0; cr < CL; ++cr) sum += part[cr];
        GY[tid / PW][tid % PW] = sum;
    } else {
        GY[tid / PW][tid % PW] = gx[tid];
    }
    __syncthreads();
    lap(3);
    // forward columnwise recurrence (dlarft):  T[0:j, j] = -tau_j T[0:j, 0:j] (Y[:, 0:j]^T y_j);  row a of T only needs row a
    if(crank == 0 && tid < PW) {
        const int a = tid;
        for(int j = a + 1; j < PW; ++j) {
            double sum = 0.0;
            for(int l = a; l < j; ++l) sum = fma(Tsm[a][l], GY[l][j], sum);
            Tsm[a][j] = -tau_sm[j] * sum;
        }
    }
    __syncthreads();
    if(crank == 0) {
        for(int e = tid; e < PW * PW; e += QR_THREADS) Tbuf[d.t_off + e] = Tsm[e / PW][e % PW];
    }
    lap(4);
    if(prof) atomicAdd(tprof + 7, 1ull);
    if(CL > 1) {   // nobody leaves while a peer may still read its shared memory
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

namespace {
// Householder panels + trailing updates over [B | 1] of all blocks (shared by both preconditioning stages)
void run_panels(double* Baug, double* Ybuf, double* Tbuf, const QrPreBlock* d_blocks, int nb, int n_max, int mn_max, cudaStream_t s, int64_t* launches)
{
    const int cluster = n_max <= RPT * QR_THREADS ? 1 : 4;
    unsigned long long* d_tprof = nullptr;
    static const bool qr_prof = std::getenv("XB_SVD_PROF") != nullptr;
    if(qr_prof) {
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_tprof), 8 * sizeof(unsigned long long), s), "qr prof");
        check_cuda(cudaMemsetAsync(d_tprof, 0, 8 * sizeof(unsigned long long), s), "qr prof reset");
    }
    for(int k0 = 0; k0 < n_max; k0 += PW) {
        if(cluster == 1) {
            qr_panel_kernel<1><<<nb, QR_THREADS, 0, s>>>(Baug, Ybuf, Tbuf, d_blocks, k0, d_tprof);
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
            check_cuda(cudaLaunchKernelEx(&cfg, qr_panel_kernel<4>, Baug, Ybuf, Tbuf, d_blocks, k0, d_tprof), "qr_panel_kernel<4> launch");
        }
        check_cuda(cudaGetLastError(), "qr_panel_kernel launch");
        const int slabs = (mn_max - k0 - PW + SLAB - 1) / SLAB;
        if(slabs > 0) {
            qr_update_kernel<<<dim3(static_cast<unsigned>(slabs), static_cast<unsigned>(nb)), QR_THREADS, 0, s>>>(Baug, Ybuf, Tbuf, d_blocks, k0);
            check_cuda(cudaGetLastError(), "qr_update_kernel launch");
        }
        *launches += 2;
    }
    if(qr_prof) {
        unsigned long long hp[8];
        check_cuda(cudaMemcpyAsync(hp, d_tprof, sizeof(hp), cudaMemcpyDeviceToHost, s), "qr prof readback");
        check_cuda(cudaStreamSynchronize(s), "qr prof sync");
        check_cuda(cudaFreeAsync(d_tprof, s), "qr prof free");
        const double np = hp[7] ? static_cast<double>(hp[7]) : 1.0;
        std::fprintf(stderr, "[qr prof] panels %llu | CTA0 cycles/panel: load %.0f reduce(16 steps) %.0f update(16 steps) %.0f gram %.0f T %.0f\n", hp[7], hp[0] / np,
                     hp[1] / np, hp[2] / np, hp[3] / np, hp[4] / np);
    }
}

// ---- stage 2 (square blocks): [W0 | 1] -> [R2 | Q2^T] with W0 = Q2 R2 -------------------------------------------------------
__global__ void qr2_pack_kernel(const double* __restrict__ W, double* __restrict__ Baug, const QrPreBlock* __restrict__ blocks)
{
    const QrPreBlock d = blocks[blockIdx.y];
    const long long cnt = 2ll * d.n * d.n, stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < cnt; e += stride) {
        const int r = static_cast<int>(e % d.n), c = static_cast<int>(e / d.n);
        Baug[d.b_off + e] = c < d.n ? W[d.w_off + e] : (r == c - d.n ? 1.0 : 0.0);
    }
}

__global__ void qr2_unpack_r_kernel(const double* __restrict__ Baug, double* __restrict__ Rtmp, const QrPreBlock* __restrict__ blocks)
{
    const QrPreBlock d = blocks[blockIdx.y];
    const long long cnt = static_cast<long long>(d.n) * d.n, stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < cnt; e += stride) {
        const int r = static_cast<int>(e % d.n), c = static_cast<int>(e / d.n);
        Rtmp[d.q_off + e] = r <= c ? Baug[d.b_off + e] : 0.0;   // below the diagonal Baug holds stale data
    }
}

// V = (Q2^T)^T: 32 x 32 tiles; grid = (ceil(n_max / 32), ceil(n_max / 32), nblocks), block = (32, 8)
__global__ void qr2_unpack_q_kernel(const double* __restrict__ Baug, double* __restrict__ V, const QrPreBlock* __restrict__ blocks)
{
    __shared__ double tile[32][33];
    const QrPreBlock d = blocks[blockIdx.z];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    if(r0 >= d.n || c0 >= d.n) return;
    for(int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + threadIdx.x, r = r0 + j;   // P[c][r] = Baug[c + (n + r) n]
        tile[j][threadIdx.x] = (r < d.n && c < d.n) ? Baug[d.b_off + c + static_cast<long long>(d.n + r) * d.n] : 0.0;
    }
    __syncthreads();
    for(int j = threadIdx.y; j < 32; j += 8) {
        const int r = r0 + threadIdx.x, c = c0 + j;
        if(r < d.n && c < d.n) V[d.v_off + r + static_cast<long long>(c) * d.n] = tile[threadIdx.x][j];
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
    const dim3 ngrid(static_cast<unsigned>((m_max + 127) / 128), static_cast<unsigned>((n_max + NORM_CHUNK - 1) / NORM_CHUNK), static_cast<unsigned>(nb));
    const dim3 rgrid(static_cast<unsigned>((m_max + 127) / 128), static_cast<unsigned>(nb));
    qr_row_norms_kernel<<<ngrid, 128, 0, s>>>(W, nrm, d_blocks);
    qr_rank_rows_kernel<<<rgrid, 128, sizeof(double) * m_max, s>>>(nrm, perm, d_blocks);
    qr_pack_kernel<<<tgrid, tblock, 0, s>>>(W, Baug, perm, d_blocks);
    check_cuda(cudaGetLastError(), "qr_pack_kernel launch");
    *launches += 3;
    run_panels(Baug, Ybuf, Tbuf, d_blocks, nb, n_max, mn_max, s, launches);
    qr_unpack_kernel<<<tgrid, tblock, 0, s>>>(Baug, W, V, perm, d_blocks);
    check_cuda(cudaGetLastError(), "qr_unpack_kernel launch");
    ++*launches;
}

}   // namespace xb

namespace xb {

// Second preconditioning step for SQUARE blocks (Drmac-Veselic: a QR factorisation of the first triangular factor): with the state
// (W0, V0) of the first step and W0 = Q2 R2, the TRANSPOSED problem A^T = V S U^T starts from  W' = A^T Q2 = V0 R2^T,  U' = Q2
// (Gram(W') = R2 R2^T is closer to diagonal than Gram(W0) = R2^T R2: three to four sweeps fewer).  This routine factorises
// [W0 | 1] and leaves R2 (upper triangular) in Rtmp (offsets q_off); the caller forms W' = V0 R2^T with the grouped GEMM and then
// calls launch_svd_precondition2_q, which overwrites V0 by Q2.  The roles of W and V (left / right vectors) are swapped afterwards.
void launch_svd_precondition2_factor(const double* W, double* Baug, double* Ybuf, double* Tbuf, double* Rtmp, const QrPreBlock* d_blocks,
                                     const std::vector<QrPreBlock>& h_blocks, cudaStream_t s, int64_t* launches)
{
    const int nb = static_cast<int>(h_blocks.size());
    if(nb == 0) return;
    int n_max = 0;
    for(const QrPreBlock& b : h_blocks) n_max = std::max(n_max, b.n);
    qr2_pack_kernel<<<dim3(296, static_cast<unsigned>(nb)), 256, 0, s>>>(W, Baug, d_blocks);
    check_cuda(cudaGetLastError(), "qr2_pack_kernel launch");
    ++*launches;
    run_panels(Baug, Ybuf, Tbuf, d_blocks, nb, n_max, 2 * n_max, s, launches);
    qr2_unpack_r_kernel<<<dim3(296, static_cast<unsigned>(nb)), 256, 0, s>>>(Baug, Rtmp, d_blocks);
    check_cuda(cudaGetLastError(), "qr2_unpack_r_kernel launch");
    ++*launches;
}

void launch_svd_precondition2_q(const double* Baug, double* V, const QrPreBlock* d_blocks, const std::vector<QrPreBlock>& h_blocks, cudaStream_t s, int64_t* launches)
{
    const int nb = static_cast<int>(h_blocks.size());
    if(nb == 0) return;
    int n_max = 0;
    for(const QrPreBlock& b : h_blocks) n_max = std::max(n_max, b.n);
    const dim3 grid(static_cast<unsigned>((n_max + 31) / 32), static_cast<unsigned>((n_max + 31) / 32), static_cast<unsigned>(nb));
    qr2_unpack_q_kernel<<<grid, dim3(32, 8), 0, s>>>(Baug, V, d_blocks);
    check_cuda(cudaGetLastError(), "qr2_unpack_q_kernel launch");
    ++*launches;
}

}   // namespace xb
