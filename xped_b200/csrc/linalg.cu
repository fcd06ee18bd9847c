// Batched per-sector decompositions: tSVD (+ global truncation), teigh, tQR.
//
// Reference: Tensor::tSVD / full_svd_impl / cutoff_matrices (src/Core/Tensor.cpp:478-633) loops over sectors and
// calls LAPACK/Eigen per block (src/Interfaces/PlainInterface_Eigen_impl.cpp:114-127); here ALL sector blocks are
// decomposed together by a QR-preconditioned block one-sided Jacobi (Hestenes) method:
//   work matrix W_q (tall side first) and accumulator V_q, started from (A Q_B, Q_B) with A^T = Q_B R_B (precond.cu);
//   one kernel launch per round-robin round handles every (column-block pair, sector) in parallel: Gram matrix of the
//   2x16 columns -> Jacobi rotations on it in shared memory (cyclic by blocks: only the 16x16 cross pairs after the first
//   round of a sweep) -> panel update of W_q and V_q.  Converged columns of W_q are sigma_j * u_j.  Pairs whose blocks did
//   not change since they were found orthogonal are skipped; a sweep is replayed as a CUDA graph.
//   svd_round_resident_kernel (row slices resident in shared memory, clusters of 1 / 2 / 8 CTAs) up to 2432 rows,
//   svd_round_kernel (rows streamed in 64-row chunks, clusters of 4) beyond.
// The singular values of all sectors are then sorted globally on the host (<= ~1e4 numbers) for the truncation
// decision (Tensor.cpp:543-557) and one gather kernel writes the truncated U, Sigma, V^dagger.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>

#include "engine.hpp"
#include "ad.hpp"
#include "precond.hpp"
#include "svd_internal.hpp"

namespace xb {

namespace {

constexpr int SB = 16;    // column-block width
constexpr int SP = 32;    // panel = two column blocks
constexpr int SCH = 64;   // rows per staged chunk
constexpr int SVD_THREADS = 256;

// mode 0: all pairs of the 32-column panel (31 steps), 1: cross pairs only (16 steps); gbi / gbj: indices of the two column blocks in
// the modification-stamp table (gbj = -1: single block)
struct SvdTask { int block, bi, bj, mode, gbi, gbj, pad0, pad1; };

// per-block diagonal shift 1.5 * ||A_q||_F for teigh: A_q + c_q 1 is positive definite with condition number <= 5
__global__ void block_shift_kernel(const double* __restrict__ A, const long long* __restrict__ a_off, const SvdBlock* __restrict__ blocks, int nblocks,
                                   double* __restrict__ shift)
{
    __shared__ double sm[256];
    const int b = blockIdx.x;
    const SvdBlock d = blocks[b];
    const long long cnt = static_cast<long long>(d.rows) * d.cols;
    double s = 0.0;
    for(long long e = threadIdx.x; e < cnt; e += blockDim.x) s = fma(A[a_off[b] + e], A[a_off[b] + e], s);
    sm[threadIdx.x] = s;
    __syncthreads();
    for(int o = 128; o > 0; o >>= 1) {
        if(threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if(threadIdx.x == 0) shift[b] = 1.5 * sqrt(sm[0]);
}

__global__ void svd_init_kernel(const double* __restrict__ A, double* __restrict__ W, double* __restrict__ V, const SvdBlock* __restrict__ blocks,
                                const long long* __restrict__ a_off, const int* __restrict__ a_rows, const int* __restrict__ transposed, int nblocks,
                                const double* __restrict__ shift)
{
    for(int b = blockIdx.y; b < nblocks; b += gridDim.y) {
        const SvdBlock d = blocks[b];
        const long long cnt = static_cast<long long>(d.rows) * d.cols;
        const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
        const int lda = a_rows[b];
        const double sh = shift ? shift[b] : 0.0;
        for(long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < cnt; e += stride) {
            const int r = static_cast<int>(e % d.rows), c = static_cast<int>(e / d.rows);
            const double a = transposed[b] ? A[a_off[b] + c + static_cast<long long>(r) * lda] : A[a_off[b] + r + static_cast<long long>(c) * lda];
            W[d.w_off + e] = r == c ? a + sh : a;
        }
        const long long vc = static_cast<long long>(d.cols) * d.cols;
        for(long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < vc; e += stride)
            V[d.v_off + e] = (e % d.cols == e / d.cols) ? 1.0 : 0.0;
    }
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Jacobi rotation that annihilates g_pq of [[a, c], [c, b]]; false if the pair is already orthogonal to `tol`.
// With d = b - a, h = 2c, D = |d| + sqrt(d^2 + h^2):  t = sign(d) h / D,  cs = 1 / sqrt(1 + t^2) = D / sqrt(D^2 + h^2),
// sn = cs t = sign(d) h / sqrt(D^2 + h^2): two reciprocal square roots and no division on the dependent chain of a Jacobi step.
// cs^2 + sn^2 = (D^2 + h^2) rs^2 = 1 up to the rounding of rs, so the rotation is orthogonal to working precision.
__device__ __forceinline__ bool jacobi_rotation(double a, double b, double c, double tol2, double& cs, double& sn)
{
    cs = 1.0;
    sn = 0.0;
    if(!(a > 0.0 && b > 0.0 && c * c > tol2 * a * b)) return false;
    const double d = b - a, h = 2.0 * c;
    const double r2 = fma(d, d, h * h);
#ifndef XB_JACOBI_EXACT_ANGLE
    // |d| + sqrt(r2) only fixes the rotation ANGLE: a 2^-22 accurate square root leaves cs^2 + sn^2 = 1 intact (see below)
    double rs1;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs1) : "d"(r2));
    if(r2 < 1.0e-280) rs1 = rsqrt(r2);   // the approximation flushes denormals
    const double D = fabs(d) + r2 * rs1;
#else
    const double D = fabs(d) + r2 * rsqrt(r2);
#endif
    const double rs = rsqrt(fma(D, D, h * h));
    cs = D * rs;
    sn = copysign(h, d * h) * rs;
    return true;
}

constexpr int XS_LD = SCH + 4;   // == 4 mod 16: the DMMA fragment loads of both operand roles hit 16 distinct banks per half-warp
constexpr int VL_LD = SP + 4;
constexpr int G_LD = SP + 1;

// Round-robin tournament by physically rotating the players ("caterpillar"): position 0 is fixed, the others move one
// seat per step along 2 -> 4 -> ... -> 30 -> 31 -> 29 -> ... -> 1 -> 2.  The pairs of every step are then always the
// seats (2g, 2g+1), all shared-memory addresses of the Jacobi step are loop invariants, and after 31 steps every pair
// has met once and everybody is back in its original seat.
__device__ __forceinline__ int seat_next(int x)
{
    if(x == 0) return 0;
    if((x & 1) == 0) return x == SP - 2 ? SP - 1 : x + 2;   // top row moves right, the last one drops to the bottom row
    return x == 1 ? 2 : x - 2;                              // bottom row moves left, the first one climbs to the top row
}

// One round of the block one-sided Jacobi method: every (sector block, column-block pair) of the round is handled by a
// cluster of CL CTAs that split the rows.  Phase 1: Gram matrix G = X^T X of the 32-column panel on the FP64 tensor
// cores (partial sums exchanged through distributed shared memory, summed in a fixed order so that all CTAs of the
// cluster hold bit-identical G).  Phase 2: one cyclic two-sided Jacobi sweep on G in shared memory -- thread (k, l) owns
// the 2x2 intersection of rotation pair k's rows and pair l's columns, G and Vl are double buffered and written to the
// rotated seats, so a step costs ONE barrier and no index arithmetic; every CTA of the cluster runs it redundantly
// (deterministic => identical rotations).  Phase 3: X <- X Vl and V <- V Vl on the tensor cores for this CTA's row
// slices, the next 64-row chunk being prefetched into registers while the current one is multiplied.
template <int CL>
__global__ void __launch_bounds__(SVD_THREADS)
svd_round_kernel(double* __restrict__ W, double* __restrict__ V, const SvdBlock* __restrict__ blocks, const SvdTask* __restrict__ tasks, int* __restrict__ rotated,
                 unsigned long long* __restrict__ tprof, int* __restrict__ mod, int* __restrict__ clean, const int* __restrict__ sweep, int round_idx, int n_rounds)
{
    // optional phase timing (XB_SVD_PROF=1): thread 0 of CTA 0 accumulates clock64() deltas per phase
    const bool prof = tprof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long tp = prof ? clock64() : 0;
    auto lap = [&](int slot) {
        if(prof) {
            const long long now = clock64();
            atomicAdd(tprof + slot, static_cast<unsigned long long>(now - tp));
            tp = now;
        }
    };
    __shared__ __align__(16) double Xs[SP * XS_LD];          // chunk staging [col][row]; during phase 2: second Vl buffer
    __shared__ __align__(16) double Gs[2][SP * G_LD];        // Gs[1] doubles as this CTA's partial Gram matrix in phase 1
    __shared__ __align__(16) double Vls[SP * VL_LD];
    __shared__ int colidx[SP];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned crank = 0;
    if(CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const SvdTask task = tasks[blockIdx.x / CL];
    // skip pairs found orthogonal whose column blocks were not rotated since (see the resident kernel)
    const int t_now = *sweep * n_rounds + round_idx + 1;
    {
        const int last_mod = max(mod[task.gbi], task.gbj >= 0 ? mod[task.gbj] : 0);
        if(clean[blockIdx.x / CL] > last_mod) return;
    }
    const SvdBlock blk = blocks[task.block];
    if(tid < SP) {   // interleaved seats: even = block bi, odd = block bj (cross mode rotates only pairs between the two blocks)
        const int i = tid >> 1;
        int c = (tid & 1) ? (task.bj >= 0 ? task.bj * SB + i : -1) : task.bi * SB + i;
        if(c >= blk.cols) c = -1;
        colidx[tid] = c;
    }
    __syncthreads();

    // row slices of this CTA (multiples of SCH so that chunks never straddle CTAs)
    auto slice = [&](int nrows, int& lo, int& hi) {
        const int chunks = (nrows + SCH - 1) / SCH;
        const int per = (chunks + CL - 1) / CL;
        lo = min(nrows, static_cast<int>(crank) * per * SCH);
        hi = min(nrows, (static_cast<int>(crank) + 1) * per * SCH);
    };
    // 64 rows x 32 columns per chunk; a warp reads 32 consecutive rows of one column (256 B, coalesced)
    constexpr int PER_THREAD = (SP * SCH) / SVD_THREADS;
    auto fetch = [&](const double* __restrict__ M, int nrows, int r0, double (&r)[PER_THREAD]) {
#pragma unroll
        for(int it = 0; it < PER_THREAD; ++it) {
            const int e = tid + it * SVD_THREADS;
            const int col = e / SCH, row = e % SCH;
            const int gr = r0 + row, gc = colidx[col];
            r[it] = (gc >= 0 && gr < nrows) ? M[gr + static_cast<long long>(gc) * nrows] : 0.0;
        }
    };
    auto stage = [&](const double (&r)[PER_THREAD]) {
#pragma unroll
        for(int it = 0; it < PER_THREAD; ++it) {
            const int e = tid + it * SVD_THREADS;
            Xs[(e / SCH) * XS_LD + (e % SCH)] = r[it];
        }
    };

    // ---- phase 1: Gram matrix. warp w owns output tiles (ti = w/2, tj = 2*(w&1) + {0,1}) of the 4x4 grid of 8x8 tiles
    double* __restrict__ Gpart = Gs[1];
    {
        const int ti = warp >> 1, tj0 = (warp & 1) * 2;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        int lo, hi;
        slice(blk.rows, lo, hi);
        const double* __restrict__ Wm = W + blk.w_off;
        double reg[PER_THREAD];
        if(lo < hi) fetch(Wm, blk.rows, lo, reg);
        for(int r0 = lo; r0 < hi; r0 += SCH) {
            stage(reg);
            __syncthreads();
            if(r0 + SCH < hi) fetch(Wm, blk.rows, r0 + SCH, reg);   // prefetch the next chunk
#pragma unroll 4
            for(int k0 = 0; k0 < SCH; k0 += 4) {
                const double a = Xs[(ti * 8 + (lane >> 2)) * XS_LD + k0 + (lane & 3)];
#pragma unroll
                for(int j = 0; j < 2; ++j) {
                    const double b = Xs[((tj0 + j) * 8 + (lane >> 2)) * XS_LD + k0 + (lane & 3)];
                    dmma_8x8x4(acc[j][0], acc[j][1], a, b);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for(int j = 0; j < 2; ++j) {
            const int r = ti * 8 + (lane >> 2), c = (tj0 + j) * 8 + 2 * (lane & 3);
            Gpart[r * G_LD + c] = acc[j][0];
            Gpart[r * G_LD + c + 1] = acc[j][1];
        }
    }
    lap(0);   // setup + Gram
    if(CL > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
        __syncthreads();
    }
    double* __restrict__ Vl1 = Xs;   // second Vl buffer (aliases the chunk staging area, unused in phase 2)
    for(int e = tid; e < SP * SP; e += SVD_THREADS) {
        const int r = e / SP, c = e % SP;
        double sum = 0.0;
        if(CL > 1) {
            const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(&Gpart[r * G_LD + c]));
#pragma unroll
            for(int cr = 0; cr < CL; ++cr) {   // fixed order: bit-identical in every CTA of the cluster
                unsigned remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cr));
                double v;
                asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
                sum += v;
            }
        } else {
            sum = Gpart[r * G_LD + c];
        }
        Gs[0][r * G_LD + c] = sum;
        Vl1[r * VL_LD + c] = (r == c) ? 1.0 : 0.0;   // the sweep starts in buffer 1 and ends (31 swaps later) in Vls
    }
    if(CL > 1) {   // peers must not overwrite their Gpart (= Gs[1]) or exit before everybody has read it
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
        __syncthreads();
    }

    lap(1);   // cluster reduction of G
    // ---- phase 2: one cyclic two-sided Jacobi sweep on G; thread (k, l): rows {2k, 2k+1} x columns {2l, 2l+1}
    {
        const int k = tid >> 4, l = tid & 15;
        const double tol2 = blk.tol * blk.tol;
        const int r0 = 2 * k, r1 = 2 * k + 1, c0 = 2 * l, c1 = 2 * l + 1;
        const bool cross = task.mode != 0;
        auto nxt = [&](int x) { return cross ? ((x & 1) ? ((x + 2) & (SP - 1)) : x) : seat_next(x); };
        const int nr0 = nxt(r0), nr1 = nxt(r1), nc0 = nxt(c0), nc1 = nxt(c1);
        // loop-invariant offsets
        const int g00 = r0 * G_LD + c0, g01 = r0 * G_LD + c1, g10 = r1 * G_LD + c0, g11 = r1 * G_LD + c1;
        const int h00 = nr0 * G_LD + nc0, h01 = nr0 * G_LD + nc1, h10 = nr1 * G_LD + nc0, h11 = nr1 * G_LD + nc1;
        const int pa = c0 * G_LD + c0, pb = c1 * G_LD + c1, pc = c0 * G_LD + c1;   // pivots of column pair l
        const int v0p = k * VL_LD + c0, v0q = k * VL_LD + c1, v1p = (k + 16) * VL_LD + c0, v1q = (k + 16) * VL_LD + c1;
        const int w0p = k * VL_LD + nc0, w0q = k * VL_LD + nc1, w1p = (k + 16) * VL_LD + nc0, w1q = (k + 16) * VL_LD + nc1;
        int any = 0;
        auto jacobi_step = [&](const double* __restrict__ Gc, double* __restrict__ Gn, const double* __restrict__ Vc, double* __restrict__ Vn) {
            double cl, sl;
            const bool rl = jacobi_rotation(Gc[pa], Gc[pb], Gc[pc], tol2, cl, sl);
            const double ck = __shfl_sync(0xffffffffu, cl, k);   // lane k of the warp (l == k there) holds pair k's rotation
            const double sk = __shfl_sync(0xffffffffu, sl, k);
            const bool rk = __shfl_sync(0xffffffffu, static_cast<int>(rl), k) != 0;
            any |= rl ? 1 : 0;
            const double a00 = Gc[g00], a01 = Gc[g01], a10 = Gc[g10], a11 = Gc[g11];
            // columns: g[:, p] = c g[:, p] - s g[:, q];  g[:, q] = s g[:, p] + c g[:, q]
            const double b00 = cl * a00 - sl * a01, b01 = sl * a00 + cl * a01;
            const double b10 = cl * a10 - sl * a11, b11 = sl * a10 + cl * a11;
            // rows: g[p, :] = c g[p, :] - s g[q, :];  g[q, :] = s g[p, :] + c g[q, :]
            double n00 = ck * b00 - sk * b10, n01 = ck * b01 - sk * b11;
            double n10 = sk * b00 + ck * b10, n11 = sk * b01 + ck * b11;
            if(k == l && rk) {   // the pivot block itself: off-diagonal annihilated exactly
                n01 = 0.0;
                n10 = 0.0;
            }
            Gn[h00] = n00;
            Gn[h01] = n01;
            Gn[h10] = n10;
            Gn[h11] = n11;
            const double x0p = Vc[v0p], x0q = Vc[v0q], x1p = Vc[v1p], x1q = Vc[v1q];
            Vn[w0p] = cl * x0p - sl * x0q;
            Vn[w0q] = sl * x0p + cl * x0q;
            Vn[w1p] = cl * x1p - sl * x1q;
            Vn[w1q] = sl * x1p + cl * x1q;
            __syncthreads();
        };
        // double buffering: even steps Gs[0] -> Gs[1], Vl1 -> Vls; odd steps back
        const int nsteps = cross ? SB : SP - 1;
#pragma unroll 1
        for(int step = 0; step + 1 < nsteps; step += 2) {
            jacobi_step(Gs[0], Gs[1], Vl1, Vls);
            jacobi_step(Gs[1], Gs[0], Vls, Vl1);
        }
        if(nsteps & 1) jacobi_step(Gs[0], Gs[1], Vl1, Vls);
        any = __syncthreads_or(any);
        lap(2);   // inner Jacobi sweep
        if(!any) {   // panel already orthogonal (uniform over the cluster: identical G everywhere)
            if(tid == 0 && crank == 0) clean[blockIdx.x / CL] = t_now;
            return;
        }
        if(!(nsteps & 1)) {   // an even number of steps leaves the rotation matrix in Vl1 (= the chunk staging area): move it to Vls
            for(int e = tid; e < SP * SP; e += SVD_THREADS) Vls[(e / SP) * VL_LD + (e % SP)] = Vl1[(e / SP) * VL_LD + (e % SP)];
            __syncthreads();
        }
    }
    if(tid == 0 && crank == 0) {
        atomicExch(rotated, 1);
        mod[task.gbi] = t_now;
        if(task.gbj >= 0) mod[task.gbj] = t_now;
    }

    // ---- phase 3: X <- X * Vl for this CTA's rows of W and of V. warp w owns 8 rows of the 64-row chunk, all 32 columns.
    {
        int wlo, whi, vlo, vhi;
        slice(blk.rows, wlo, whi);
        slice(blk.cols, vlo, vhi);
        const int nw = (whi - wlo + SCH - 1) / SCH, nv = (vhi - vlo + SCH - 1) / SCH;
        double reg[PER_THREAD];
        auto chunk_of = [&](int c, double*& M, int& nrows, int& r0) {
            if(c < nw) {
                M = W + blk.w_off;
                nrows = blk.rows;
                r0 = wlo + c * SCH;
            } else {
                M = V + blk.v_off;
                nrows = blk.cols;
                r0 = vlo + (c - nw) * SCH;
            }
        };
        double* M;
        int nrows, r0;
        if(nw + nv > 0) {
            chunk_of(0, M, nrows, r0);
            fetch(M, nrows, r0, reg);
        }
        for(int c = 0; c < nw + nv; ++c) {
            chunk_of(c, M, nrows, r0);
            stage(reg);
            __syncthreads();
            if(c + 1 < nw + nv) {   // prefetch the next chunk while this one is multiplied
                double* M2;
                int n2, r2;
                chunk_of(c + 1, M2, n2, r2);
                fetch(M2, n2, r2, reg);
            }
            double out[4][2];
#pragma unroll
            for(int j = 0; j < 4; ++j) out[j][0] = out[j][1] = 0.0;
#pragma unroll
            for(int k0 = 0; k0 < SP; k0 += 4) {
                const double a = Xs[(k0 + (lane & 3)) * XS_LD + warp * 8 + (lane >> 2)];   // A[row][k] = X[row][k]
#pragma unroll
                for(int j = 0; j < 4; ++j) {
                    const double b = Vls[(k0 + (lane & 3)) * VL_LD + j * 8 + (lane >> 2)];   // B[k][col] = Vl[k][col]
                    dmma_8x8x4(out[j][0], out[j][1], a, b);
                }
            }
            const int gr = r0 + warp * 8 + (lane >> 2);
            if(gr < nrows) {
#pragma unroll
                for(int j = 0; j < 4; ++j)
#pragma unroll
                    for(int h = 0; h < 2; ++h) {
                        const int gc = colidx[j * 8 + 2 * (lane & 3) + h];
                        if(gc >= 0) M[gr + static_cast<long long>(gc) * nrows] = out[j][h];
                    }
            }
            __syncthreads();
        }
    }
    lap(3);   // panel update
    if(prof) atomicAdd(tprof + 7, 1ull);
}

// ---------------------------------------------------------------------------------------------------------------
// Round kernel, resident form (the default): the CTA's row slice of the 32-column W panel is loaded ONCE into shared
// memory ([col][row], pitch == 4 mod 16) and stays there for the Gram phase and the update phase; the V slice is
// prefetched into registers before the Jacobi sweep so that its L2 latency hides behind the sweep.  Shared-memory
// columns are interleaved (even seats: block bi, odd seats: block bj): in cross mode (every round but the first of a
// sweep) only the 16 x 16 pairs between the two blocks are rotated -- 16 steps with the odd seats shifting by one --
// because the pairs inside a block are visited once per sweep in round 0 (standard cyclic-by-blocks ordering).
// DB = false: G and Vl planes are updated in place (read -> barrier -> write -> barrier per Jacobi step instead of one barrier with
// double buffering): 16 KB less shared memory, which lets two CTAs of the 304-row variant share an SM (D=6: 2700 CTAs per round).
// SOLO (the 8-CTA variant): only rank 0 of the cluster reduces the Gram matrix and runs the Jacobi sweep; the peers sleep in the
// cluster barrier (their SM is shared with a CTA of another pair) and then copy the 32 x 32 rotation matrix through DSMEM.
template <int CL, int VREG, bool DB>
__global__ void __launch_bounds__(SVD_THREADS, (VREG <= 20 || (!DB && VREG <= 40)) ? 2 : 1)
svd_round_resident_kernel(double* __restrict__ W, double* __restrict__ V, const SvdBlock* __restrict__ blocks, const SvdTask* __restrict__ tasks,
                          int* __restrict__ rotated, unsigned long long* __restrict__ tprof, int rs, int* __restrict__ mod, int* __restrict__ clean,
                          const int* __restrict__ sweep, int round_idx, int n_rounds)
{
    extern __shared__ __align__(16) double dsm[];
    constexpr bool SOLO = CL >= 8;
    const int LD = rs + 4;                     // rs is a multiple of 16
    double* __restrict__ Xw = dsm;                 // SP x LD
    double* __restrict__ Gp0 = Xw + SP * LD;       // G in 2x2-block planes: plane (r&1)*2 + (c&1), index (r>>1)*16 + (c>>1); two buffers
    double* __restrict__ Gp1 = DB ? Gp0 + SP * SP : Gp0;
    double* __restrict__ Vp0 = Gp1 + SP * SP;      // Vl in planes: plane (r>=16)*2 + (c&1), index (r&15)*16 + (c>>1); two buffers
    double* __restrict__ Vp1 = DB ? Vp0 + SP * SP : Vp0;
    double* __restrict__ G1 = Vp1 + SP * SP;       // partial Gram matrix (phase 1, SP x G_LD); afterwards Vl in [k][n] form (SP x VL_LD)
    __shared__ int colidx[SP];

    const bool prof = tprof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long tp = prof ? clock64() : 0;
    auto lap = [&](int slot) {
        if(prof) {
            const long long now = clock64();
            atomicAdd(tprof + slot, static_cast<unsigned long long>(now - tp));
            tp = now;
        }
    };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    unsigned crank = 0;
    if(CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const SvdTask task = tasks[blockIdx.x / CL];
    // skip pairs that were found orthogonal and whose column blocks have not been rotated since (uniform over the cluster; the
    // tables are only written after the cluster barriers of a round, by the CTAs that own the pair)
    const int t_now = *sweep * n_rounds + round_idx + 1;
    {
        const int last_mod = max(mod[task.gbi], task.gbj >= 0 ? mod[task.gbj] : 0);
        if(clean[blockIdx.x / CL] > last_mod) return;
    }
    const SvdBlock blk = blocks[task.block];
    if(tid < SP) {
        const int i = tid >> 1;
        int c = (tid & 1) ? (task.bj >= 0 ? task.bj * SB + i : -1) : task.bi * SB + i;
        if(c >= blk.cols) c = -1;
        colidx[tid] = c;
    }
    __syncthreads();

    auto slice = [&](int nrows, int& lo, int& hi) {
        const int per = (((nrows + CL - 1) / CL) + 7) & ~7;
        lo = min(nrows, static_cast<int>(crank) * per);
        hi = min(nrows, lo + per);
    };
    int wlo, whi, vlo, vhi;
    slice(blk.rows, wlo, whi);
    slice(blk.cols, vlo, vhi);
    // warp w owns columns w, w+8, w+16, w+24 of the slice; lanes run along the rows (coalesced)
    constexpr int RI = VREG / 4;
    auto fetch = [&](const double* __restrict__ M, int nrows, int lo, int hi, double (&r)[VREG]) {
#pragma unroll
        for(int ci = 0; ci < 4; ++ci) {
            const int gc = colidx[warp + 8 * ci];
            const double* __restrict__ colp = M + static_cast<long long>(gc < 0 ? 0 : gc) * nrows;
#pragma unroll
            for(int ri = 0; ri < RI; ++ri) {
                const int gr = lo + lane + 32 * ri;
                r[ci * RI + ri] = (gc >= 0 && gr < hi) ? colp[gr] : 0.0;
            }
        }
    };
    auto stage = [&](const double (&r)[VREG]) {
#pragma unroll
        for(int ci = 0; ci < 4; ++ci)
#pragma unroll
            for(int ri = 0; ri < RI; ++ri) {
                const int row = lane + 32 * ri;
                if(row < rs) Xw[(warp + 8 * ci) * LD + row] = r[ci * RI + ri];
            }
    };
    double* __restrict__ Wm = W + blk.w_off;
    double* __restrict__ Vm = V + blk.v_off;
    double reg[VREG];
    fetch(Wm, blk.rows, wlo, whi, reg);
    stage(reg);
    __syncthreads();
    fetch(Vm, blk.cols, vlo, vhi, reg);   // in flight during the Gram phase and the Jacobi sweep
    lap(0);

    // ---- phase 1: Gram matrix of the slice. warp w owns tiles (ti = w/2, tj = 2*(w&1) + {0,1}) of the 4x4 grid of 8x8 tiles
    {
        const int ti = warp >> 1, tj0 = (warp & 1) * 2;
        // two k-interleaved accumulator sets per output tile: four independent DMMA chains per warp
        double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
        const int kend = (whi - wlo + 7) & ~7;   // rows beyond the slice are zero-filled up to rs (a multiple of 16)
        const double* __restrict__ pa = Xw + (ti * 8 + fr) * LD + fk;
        const double* __restrict__ pb0 = Xw + (tj0 * 8 + fr) * LD + fk;
        const double* __restrict__ pb1 = pb0 + 8 * LD;
#pragma unroll 2
        for(int k0 = 0; k0 < kend; k0 += 8) {
            const double a0 = pa[k0], a1 = pa[k0 + 4];
            dmma_8x8x4(acc[0][0][0], acc[0][0][1], a0, pb0[k0]);
            dmma_8x8x4(acc[0][1][0], acc[0][1][1], a0, pb1[k0]);
            dmma_8x8x4(acc[1][0][0], acc[1][0][1], a1, pb0[k0 + 4]);
            dmma_8x8x4(acc[1][1][0], acc[1][1][1], a1, pb1[k0 + 4]);
        }
#pragma unroll
        for(int j = 0; j < 2; ++j) {
            const int r = ti * 8 + fr, c = (tj0 + j) * 8 + 2 * fk;
            G1[r * G_LD + c] = acc[0][j][0] + acc[1][j][0];
            G1[r * G_LD + c + 1] = acc[0][j][1] + acc[1][j][1];
        }
    }
    lap(1);
    if(CL > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
        __syncthreads();
    }
    double* __restrict__ Vl = SOLO ? G1 + SP * VL_LD : G1;   // SOLO: a region of its own (peers copy it while G1 may still be read)
    const bool solver = !SOLO || crank == 0;
    if(solver) {
        // every thread sums 4 elements; the remote loads are issued in batches of (up to) 8 ranks before their first use (fixed
        // order: bit-identical G in all CTAs)
        constexpr int BATCH = CL < 8 ? CL : 8;
        double sum[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for(int c0r = 0; c0r < CL; c0r += BATCH) {
            double part[4][BATCH];
#pragma unroll
            for(int i = 0; i < 4; ++i) {
                const int e = tid + i * SVD_THREADS, r = e / SP, c = e % SP;
                if(CL > 1) {
                    const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(&G1[r * G_LD + c]));
#pragma unroll
                    for(int cr = 0; cr < BATCH; ++cr) {
                        unsigned remote;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(c0r + cr));
                        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(part[i][cr]) : "r"(remote));
                    }
                } else {
                    part[i][0] = G1[r * G_LD + c];
                }
            }
#pragma unroll
            for(int i = 0; i < 4; ++i)
#pragma unroll
                for(int cr = 0; cr < BATCH; ++cr) sum[i] += part[i][cr];
        }
#pragma unroll
        for(int i = 0; i < 4; ++i) {
            const int e = tid + i * SVD_THREADS, r = e / SP, c = e % SP;
            Gp0[((r & 1) * 2 + (c & 1)) * (SP * SP / 4) + (r >> 1) * 16 + (c >> 1)] = sum[i];
            Vp0[((r >> 4) * 2 + (c & 1)) * (SP * SP / 4) + (r & 15) * 16 + (c >> 1)] = (r == c) ? 1.0 : 0.0;
        }
    }
    // second barrier.  Not SOLO: peers must not overwrite G1 (with Vl, after the sweep) or exit before everybody has read it: arrive
    // now, wait after the sweep.  SOLO: the peers arrive at once, rank 0 after its sweep; completion = "Vl of rank 0 is ready".
    if(CL > 1 && !(SOLO && crank == 0)) asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    __syncthreads();
    lap(2);

    // ---- phase 2: Jacobi steps on G; thread (k, l) owns the 2x2 block rows {2k, 2k+1} x columns {2l, 2l+1}; pairs are always the
    // seats (2g, 2g+1).  G and Vl live in four planes (one per position inside the 2x2 block), so that all loads of a step are
    // unit-stride over the threads (no bank conflicts) and the stores to the rotated seats are permutations of them.
    int any_rot = 0;
    if(solver) {
        constexpr int PL = SP * SP / 4;
        const bool cross = task.mode != 0;
        auto nxt = [&](int x) { return cross ? ((x & 1) ? ((x + 2) & (SP - 1)) : x) : seat_next(x); };
        const int k = tid >> 4, l = tid & 15;
        const double tol2 = blk.tol * blk.tol;
        const int r0 = 2 * k, r1 = 2 * k + 1, c0 = 2 * l, c1 = 2 * l + 1;
        const int nr0 = nxt(r0), nr1 = nxt(r1), nc0 = nxt(c0), nc1 = nxt(c1);
        auto gidx = [](int r, int c) { return ((r & 1) * 2 + (c & 1)) * PL + (r >> 1) * 16 + (c >> 1); };
        auto vidx = [](int r, int c) { return ((r >> 4) * 2 + (c & 1)) * PL + (r & 15) * 16 + (c >> 1); };
        const int h00 = gidx(nr0, nc0), h01 = gidx(nr0, nc1), h10 = gidx(nr1, nc0), h11 = gidx(nr1, nc1);
        const int pa = 17 * l, pb = 3 * PL + 17 * l, pc = PL + 17 * l;   // pivot block (l, l)
        const int w0p = vidx(k, nc0), w0q = vidx(k, nc1), w1p = vidx(k + 16, nc0), w1q = vidx(k + 16, nc1);
        int any = 0;
        auto jacobi_step = [&](const double* __restrict__ Gc, double* __restrict__ Gn, const double* __restrict__ Vc, double* __restrict__ Vn) {
            double cl, sl;
            const double a00 = Gc[tid], a01 = Gc[PL + tid], a10 = Gc[2 * PL + tid], a11 = Gc[3 * PL + tid];
            const double x0p = Vc[tid], x0q = Vc[PL + tid], x1p = Vc[2 * PL + tid], x1q = Vc[3 * PL + tid];
            const bool rl = jacobi_rotation(Gc[pa], Gc[pb], Gc[pc], tol2, cl, sl);
            const double ck = __shfl_sync(0xffffffffu, cl, k);
            const double sk = __shfl_sync(0xffffffffu, sl, k);
            const bool rk = __shfl_sync(0xffffffffu, static_cast<int>(rl), k) != 0;
            any |= rl ? 1 : 0;
            const double b00 = cl * a00 - sl * a01, b01 = sl * a00 + cl * a01;
            const double b10 = cl * a10 - sl * a11, b11 = sl * a10 + cl * a11;
            double n00 = ck * b00 - sk * b10, n01 = ck * b01 - sk * b11;
            double n10 = sk * b00 + ck * b10, n11 = sk * b01 + ck * b11;
            if(k == l && rk) {
                n01 = 0.0;
                n10 = 0.0;
            }
            if(!DB) __syncthreads();   // in place: everybody has read its 2x2 blocks before the rotated seats are overwritten
            Gn[h00] = n00;
            Gn[h01] = n01;
            Gn[h10] = n10;
            Gn[h11] = n11;
            Vn[w0p] = cl * x0p - sl * x0q;
            Vn[w0q] = sl * x0p + cl * x0q;
            Vn[w1p] = cl * x1p - sl * x1q;
            Vn[w1q] = sl * x1p + cl * x1q;
            __syncthreads();
        };
        const int nsteps = cross ? SB : SP - 1;
#pragma unroll 1
        for(int step = 0; step + 1 < nsteps; step += 2) {
            jacobi_step(Gp0, Gp1, Vp0, Vp1);
            jacobi_step(Gp1, Gp0, Vp1, Vp0);
        }
        const double* Vfin = Vp0;
        if(nsteps & 1) {
            jacobi_step(Gp0, Gp1, Vp0, Vp1);
            Vfin = Vp1;
        }
        any = __syncthreads_or(any);
        any_rot = any;
        if(CL > 1 && !SOLO) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        lap(3);
        if(any || SOLO) {
            // everybody is back in the original seat: Vl[row][col] in the [k][n] form the DMMA fragments of phase 3 read
            Vl[k * VL_LD + c0] = Vfin[tid];
            Vl[k * VL_LD + c1] = Vfin[PL + tid];
            Vl[(k + 16) * VL_LD + c0] = Vfin[2 * PL + tid];
            Vl[(k + 16) * VL_LD + c1] = Vfin[3 * PL + tid];
            if(SOLO && tid == 0) Vl[SP * VL_LD] = any ? 1.0 : 0.0;
        }
        __syncthreads();
    }
    if(SOLO) {
        if(crank == 0) asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if(crank != 0) {   // copy the rotation matrix (and the "rotated at all" flag) of rank 0
            const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(Vl));
            unsigned remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
            double v[5];
#pragma unroll
            for(int i = 0; i < 5; ++i) {
                const int e = tid + i * SVD_THREADS;
                v[i] = 0.0;
                if(e <= SP * VL_LD) asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v[i]) : "r"(remote + 8u * static_cast<unsigned>(e)));
            }
#pragma unroll
            for(int i = 0; i < 5; ++i) {
                const int e = tid + i * SVD_THREADS;
                if(e <= SP * VL_LD) Vl[e] = v[i];
            }
        }
        __syncthreads();
        any_rot = Vl[SP * VL_LD] != 0.0;
        // third barrier: rank 0 may not leave before the peers have copied; everybody arrives now and waits at the very end
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    }
    if(!any_rot) {   // panel already orthogonal (uniform over the cluster)
        if(tid == 0 && crank == 0) clean[blockIdx.x / CL] = t_now;
        if(SOLO) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        return;
    }
    if(tid == 0 && crank == 0) {
        atomicExch(rotated, 1);
        mod[task.gbi] = t_now;
        if(task.gbj >= 0) mod[task.gbj] = t_now;
    }

    // ---- phase 3: X <- X * Vl for the resident W slice, then for the V slice (registers -> shared). warp w: 8-row groups w, w+8, ...
    auto multiply_slice = [&](double* __restrict__ M, int nrows, int lo, int hi) {
        const int ngroups = (hi - lo + 7) >> 3;
        for(int g = 2 * warp; g < ngroups; g += 2 * (SVD_THREADS / 32)) {   // two 8-row groups per pass: eight independent DMMA chains
            double out[2][4][2];
#pragma unroll
            for(int u = 0; u < 2; ++u)
#pragma unroll
                for(int j = 0; j < 4; ++j) out[u][j][0] = out[u][j][1] = 0.0;
#pragma unroll
            for(int k0 = 0; k0 < SP; k0 += 4) {
                const double* __restrict__ xp = Xw + (k0 + fk) * LD + g * 8 + fr;   // rows beyond the slice are zero-filled
                const double a0 = xp[0], a1 = xp[8];
#pragma unroll
                for(int j = 0; j < 4; ++j) {
                    const double b = Vl[(k0 + fk) * VL_LD + j * 8 + fr];
                    dmma_8x8x4(out[0][j][0], out[0][j][1], a0, b);
                    dmma_8x8x4(out[1][j][0], out[1][j][1], a1, b);
                }
            }
#pragma unroll
            for(int u = 0; u < 2; ++u) {
                const int gr = lo + (g + u) * 8 + fr;
                if(gr < hi) {
#pragma unroll
                    for(int j = 0; j < 4; ++j)
#pragma unroll
                        for(int h = 0; h < 2; ++h) {
                            const int gc = colidx[j * 8 + 2 * fk + h];
                            if(gc >= 0) M[gr + static_cast<long long>(gc) * nrows] = out[u][j][h];
                        }
                }
            }
        }
    };
    multiply_slice(Wm, blk.rows, wlo, whi);
    __syncthreads();
    stage(reg);
    __syncthreads();
    multiply_slice(Vm, blk.cols, vlo, vhi);
    if(SOLO) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    lap(4);
    if(prof) atomicAdd(tprof + 7, 1ull);
}

// column norms -> singular values; one warp per column
__global__ void svd_norms_kernel(const double* __restrict__ W, const SvdBlock* __restrict__ blocks, int nblocks, double* __restrict__ sigma)
{
    const int warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for(int b = blockIdx.y; b < nblocks; b += gridDim.y) {
        const SvdBlock d = blocks[b];
        for(int c = blockIdx.x * warps_per_block + warp; c < d.cols; c += gridDim.x * warps_per_block) {
            const double* col = W + d.w_off + static_cast<long long>(c) * d.rows;
            // scaled two-pass norm is unnecessary here: entries are O(sigma_max) at most, squares stay in range
            double s = 0.0;
            for(int r = lane; r < d.rows; r += 32) s = fma(col[r], col[r], s);
            for(int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if(lane == 0) sigma[d.s_off + c] = sqrt(s);
        }
    }
}

struct GatherCol
{
    long long w_src;   // start of the work-matrix column (length wl)
    long long v_src;   // start of the accumulator column (length vl)
    long long u_dst, vd_dst, s_dst;
    int wl, vl;
    int vd_ld;         // leading dimension (= kept count) of the V^dagger block
    int w_is_u;        // 1: W column -> U column, V column -> V^dagger row; 0: the transposed case
    double sigma;
};

__global__ void svd_gather_kernel(const double* __restrict__ W, const double* __restrict__ V, double* __restrict__ U, double* __restrict__ S,
                                  double* __restrict__ Vd, const GatherCol* __restrict__ cols, int n)
{
    for(int i = blockIdx.x; i < n; i += gridDim.x) {
        const GatherCol g = cols[i];
        const double inv = g.sigma > 0.0 ? 1.0 / g.sigma : 0.0;
        if(g.w_is_u) {
            for(int r = threadIdx.x; r < g.wl; r += blockDim.x) U[g.u_dst + r] = W[g.w_src + r] * inv;
            for(int r = threadIdx.x; r < g.vl; r += blockDim.x) Vd[g.vd_dst + static_cast<long long>(r) * g.vd_ld] = V[g.v_src + r];
        } else {
            for(int r = threadIdx.x; r < g.vl; r += blockDim.x) U[g.u_dst + r] = V[g.v_src + r];
            for(int r = threadIdx.x; r < g.wl; r += blockDim.x) Vd[g.vd_dst + static_cast<long long>(r) * g.vd_ld] = W[g.w_src + r] * inv;
        }
        if(threadIdx.x == 0) S[g.s_dst] = g.sigma;
    }
}

// start of a sweep: flag[0] = "a rotation was applied" of the previous sweep (reset here), flag[1] = number of sweeps after which the iteration was first found
// converged (-1: not yet).  Sweeps enqueued behind the converged one skip all their pairs (clean > last_mod), so the host can enqueue several without reading back.
__global__ void svd_sweep_bump_kernel(int* sweep, int* flag)
{
    if(flag[0] == 0 && flag[1] < 0) flag[1] = *sweep;
    flag[0] = 0;
    *sweep += 1;
}

// Round-robin schedule over `nb` column blocks: rounds x pairs; a bye (odd nb / nb == 1) becomes a single-block task.
void schedule(int block, int nb, int gbase, std::vector<std::vector<SvdTask>>& rounds)
{
    if(nb <= 0) return;
    if(nb == 1) {
        if(rounds.empty()) rounds.resize(1);
        rounds[0].push_back({block, 0, -1, 0, gbase, -1, 0, 0});
        return;
    }
    const int n = nb + (nb & 1);   // even number of players, player n-1 may be the dummy
    if(static_cast<int>(rounds.size()) < n - 1) rounds.resize(n - 1);
    for(int r = 0; r < n - 1; ++r) {
        for(int g = 0; g < n / 2; ++g) {
            int a, b;
            if(g == 0) {
                a = n - 1;
                b = r;
            } else {
                a = (r + g) % (n - 1);
                b = (r - g + (n - 1)) % (n - 1);
            }
            if(a > b) std::swap(a, b);
            if(b >= nb) {
                // bye: orthogonalise inside the lone block only in the first round (it meets every other block later)
                if(r == 0) rounds[r].push_back({block, a, -1, 0, gbase + a, -1, 0, 0});
            } else {
                rounds[r].push_back({block, a, b, r == 0 ? 0 : 1, gbase + a, gbase + b, 0, 0});
            }
        }
    }
}

}   // namespace

JacobiInput jacobi_input(const Tensor& t)
{
    const Layout& L = *t.layout;
    JacobiInput in;
    in.d = t.d;
    in.sig = L.sig;
    for(size_t i = 0; i < L.sector.size(); ++i) {
        in.off.push_back(L.off[i]);
        in.rows.push_back(static_cast<int>(L.rows[i]));
        in.cols.push_back(static_cast<int>(L.cols[i]));
    }
    return in;
}

// One-sided block Jacobi on every block of `in`. Leaves W (= sigma_j u_j columns), V and the singular values on the device.
JacobiResult jacobi_all(Context& ctx, const JacobiInput& in, bool shifted, bool host_sigma)
{
    const int nb = static_cast<int>(in.off.size());
    JacobiResult R;
    R.blocks.resize(nb);
    R.transposed.resize(nb);
    long long woff = 0, voff = 0, soff = 0;
    std::vector<long long> a_off(nb);
    std::vector<int> a_rows(nb);
    for(int i = 0; i < nb; ++i) {
        const int m = in.rows[i], n = in.cols[i];
        const bool tr = m < n;
        R.transposed[i] = tr;
        SvdBlock& b = R.blocks[i];
        b.rows = tr ? n : m;
        b.cols = tr ? m : n;
        b.w_off = woff;
        b.v_off = voff;
        b.s_off = soff;
        b.tol = 4.0 * 2.220446049250313e-16 * std::sqrt(static_cast<double>(std::max(b.rows, 1)));
        woff += static_cast<long long>(b.rows) * b.cols;
        voff += static_cast<long long>(b.cols) * b.cols;
        soff += b.cols;
        a_off[i] = in.off[i];
        a_rows[i] = m;
    }
    R.n_sigma = soff;
    if(host_sigma) R.sigma.assign(static_cast<size_t>(soff), 0.0);
    if(nb == 0 || woff == 0) return R;
    cudaStream_t s = ctx.stream;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&R.d_sigma), sizeof(double) * std::max<long long>(soff, 1), s), "svd sigma");
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&R.d_W), sizeof(double) * std::max<long long>(woff, 1), s), "svd W");
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&R.d_V), sizeof(double) * std::max<long long>(voff, 1), s), "svd V");

    // schedule
    std::vector<std::vector<SvdTask>> rounds;
    int n_colblocks = 0;
    for(int i = 0; i < nb; ++i) {
        const int cb = (R.blocks[i].cols + SB - 1) / SB;
        schedule(i, cb, n_colblocks, rounds);
        n_colblocks += cb;
    }
    std::vector<SvdTask> flat;
    std::vector<int> round_begin;
    for(const auto& r : rounds) {
        round_begin.push_back(static_cast<int>(flat.size()));
        flat.insert(flat.end(), r.begin(), r.end());
    }
    round_begin.push_back(static_cast<int>(flat.size()));

    // device metadata: blocks | a_off | a_rows | transposed | tasks | sigma | flag
    const size_t o_blocks = 0, o_aoff = o_blocks + sizeof(SvdBlock) * nb, o_arows = o_aoff + sizeof(long long) * nb, o_tr = o_arows + sizeof(int) * nb;
    size_t o_tasks = o_tr + sizeof(int) * nb;
    o_tasks = (o_tasks + 15) & ~size_t(15);
    size_t o_sigma = o_tasks + sizeof(SvdTask) * flat.size();
    o_sigma = (o_sigma + 15) & ~size_t(15);
    const size_t o_shift = o_sigma;
    const size_t o_flag = o_shift + sizeof(double) * nb;
    // convergence bookkeeping: mod[column block] = time of its last rotation, clean[task] = time the pair was last found orthogonal,
    // one sweep counter; time = sweep * rounds + round + 1.  A pair whose blocks did not change since it was found clean is skipped.
    const size_t o_stamps = o_flag + 16;
    const size_t n_stamps = static_cast<size_t>(n_colblocks) + flat.size() + 4;
    const size_t meta_bytes = o_stamps + sizeof(int) * n_stamps;
    char* d_meta = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_meta), meta_bytes, s), "svd meta");
    std::vector<char> h_meta(o_sigma, 0);
    std::memcpy(h_meta.data() + o_blocks, R.blocks.data(), sizeof(SvdBlock) * nb);
    std::memcpy(h_meta.data() + o_aoff, a_off.data(), sizeof(long long) * nb);
    std::memcpy(h_meta.data() + o_arows, a_rows.data(), sizeof(int) * nb);
    std::memcpy(h_meta.data() + o_tr, R.transposed.data(), sizeof(int) * nb);
    if(!flat.empty()) std::memcpy(h_meta.data() + o_tasks, flat.data(), sizeof(SvdTask) * flat.size());
    ctx.upload_async(d_meta, h_meta.data(), h_meta.size());
    auto* d_blocks = reinterpret_cast<SvdBlock*>(d_meta + o_blocks);
    auto* d_tasks = reinterpret_cast<SvdTask*>(d_meta + o_tasks);
    double* d_sigma = R.d_sigma;
    auto* d_flag = reinterpret_cast<int*>(d_meta + o_flag);
    auto* d_shift = reinterpret_cast<double*>(d_meta + o_shift);
    int* d_mod = reinterpret_cast<int*>(d_meta + o_stamps);
    int* d_clean = d_mod + n_colblocks;
    int* d_sweep = d_clean + flat.size();
    check_cuda(cudaMemsetAsync(d_mod, 0, sizeof(int) * n_stamps, s), "svd stamps reset");

    {
        if(shifted) {
            block_shift_kernel<<<nb, 256, 0, s>>>(in.d, reinterpret_cast<long long*>(d_meta + o_aoff), d_blocks, nb, d_shift);
            check_cuda(cudaGetLastError(), "block_shift_kernel launch");
            ++ctx.launches;
        }
        dim3 grid(64, static_cast<unsigned>(std::min(nb, 64)));
        svd_init_kernel<<<grid, 256, 0, s>>>(in.d, R.d_W, R.d_V, d_blocks, reinterpret_cast<long long*>(d_meta + o_aoff),
                                            reinterpret_cast<int*>(d_meta + o_arows), reinterpret_cast<int*>(d_meta + o_tr), nb, shifted ? d_shift : nullptr);
        check_cuda(cudaGetLastError(), "svd_init_kernel launch");
        ++ctx.launches;
    }
    // QR preconditioning (precond.cu): (W, V) <- (A Q_B, Q_B); blocks narrower than two column blocks gain nothing
    {
        static const bool no_precond = std::getenv("XB_SVD_NOPRECOND") != nullptr;
        int pre_max_rows = 0;
        std::vector<QrPreBlock> pre;
        std::vector<int> pre_index;   // sector block of every entry of `pre`
        long long boff = 0, yoff = 0, poff = 0, qoff = 0;
        for(const SvdBlock& b : R.blocks) {
            if(b.cols < 2 * SB) continue;
            QrPreBlock q;
            q.w_off = b.w_off;
            q.v_off = b.v_off;
            q.b_off = boff;
            q.y_off = yoff;
            q.t_off = static_cast<long long>(pre.size()) * 256;
            q.p_off = poff;
            poff += b.rows;
            q.q_off = qoff;
            qoff += static_cast<long long>((b.cols + 31) / 32) * b.rows;
            q.m = b.rows;
            q.n = b.cols;
            boff += static_cast<long long>(b.cols) * (b.rows + b.cols);
            yoff += static_cast<long long>(b.cols) * 16;
            pre_max_rows = std::max(pre_max_rows, b.cols);
            pre.push_back(q);
            pre_index.push_back(static_cast<int>(&b - R.blocks.data()));
        }
        if(!shifted && !no_precond && !in.no_precondition && !pre.empty() && svd_precondition_supported(pre_max_rows)) {
            static const bool pre_prof = std::getenv("XB_SVD_PROF") != nullptr;
            double* d_aug = nullptr;
            QrPreBlock* d_pre = nullptr;
            const long long toff = 256 * static_cast<long long>(pre.size());
            const long long aug_elems = boff + yoff + toff + qoff + poff;   // [B | 1], Y, T, partial row norms, row order (ints in the last segment)
            check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_aug), sizeof(double) * aug_elems, s), "svd precondition workspace");
            check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_pre), sizeof(QrPreBlock) * pre.size(), s), "svd precondition meta");
            ctx.upload_async(d_pre, pre.data(), sizeof(QrPreBlock) * pre.size());
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if(pre_prof) {
                e0 = ctx.get_event();
                e1 = ctx.get_event();
                cudaEventRecord(e0, s);
            }
            auto run_precondition = [&](int64_t* counter) {
                launch_svd_precondition(R.d_W, R.d_V, d_aug, d_aug + boff, d_aug + boff + yoff, d_aug + boff + yoff + toff,
                                        reinterpret_cast<int*>(d_aug + boff + yoff + toff + qoff), d_pre, pre, s, counter);
            };
            // ~80 short dependent launches: replayed as a CUDA graph like the sweeps (same validity rule: identical buffers)
            static const bool no_qr_graph = std::getenv("XB_SVD_NOGRAPH") != nullptr;
            bool replayed = false;
            if(!no_qr_graph && !pre_prof) {
                const std::string qr_key = in.sig + "#qr";
                if(ctx.sweep_graphs.size() >= 256 && ctx.sweep_graphs.find(qr_key) == ctx.sweep_graphs.end()) {   // bound the cache (layouts that keep changing)
                    for(auto& kv : ctx.sweep_graphs)
                        if(kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
                    ctx.sweep_graphs.clear();
                }
                Context::SweepGraph& g = ctx.sweep_graphs[qr_key];
                // the captured launches write W, the augmented workspace AND V: all three (and the descriptor table) must be the captured buffers
                if(g.exec == nullptr || g.w != R.d_W || g.v != d_aug || g.meta != d_pre || g.x != R.d_V) {
                    if(g.exec) cudaGraphExecDestroy(g.exec);
                    g = Context::SweepGraph();
                    cudaGraph_t graph = nullptr;
                    int64_t counted = 0;
                    bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                    if(ok) {
                        try {
                            run_precondition(&counted);
                        } catch(...) {
                            ok = false;
                        }
                        ok = (cudaStreamEndCapture(s, &graph) == cudaSuccess) && ok && graph != nullptr;
                    }
                    if(ok) ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
                    if(graph) cudaGraphDestroy(graph);
                    if(ok) {
                        g.w = R.d_W;
                        g.v = d_aug;
                        g.meta = d_pre;
                        g.x = R.d_V;
                        g.rounds = static_cast<int>(counted);
                    } else {
                        cudaGetLastError();
                        g = Context::SweepGraph();
                    }
                }
                if(g.exec) {
                    check_cuda(cudaGraphLaunch(g.exec, s), "svd preconditioner graph launch");
                    ctx.launches += g.rounds;
                    replayed = true;
                }
            }
            if(!replayed) run_precondition(&ctx.launches);
            if(pre_prof) {
                cudaEventRecord(e1, s);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                std::fprintf(stderr, "[svd prof] QR preconditioner: %zu blocks, n_max %d, %.3f ms\n", pre.size(), pre_max_rows, ms);
                ctx.event_pool.push_back(e0);
                ctx.event_pool.push_back(e1);
            }
            // second step (square blocks): one more QR for one to three sweeps fewer
            {
                static const int qr2_env = std::getenv("XB_SVD_QR2") ? std::atoi(std::getenv("XB_SVD_QR2")) : -1;
                std::vector<QrPreBlock> pre2;
                std::vector<int> pre2_index;
                long long roff = 0;
                for(size_t k = 0; k < pre.size(); ++k) {
                    if(pre[k].m != pre[k].n) continue;
                    QrPreBlock q = pre[k];
                    q.q_off = roff;
                    roff += static_cast<long long>(q.n) * q.n;
                    pre2.push_back(q);
                    pre2_index.push_back(pre_index[k]);
                }
                const bool want2 = qr2_env > 0;   // opt-in (XB_SVD_QR2=1): measured neutral (D=6: 13 instead of 14 sweeps for one more QR; D=4: 51.9 vs 49.3 ms per step)
                if(want2 && !pre2.empty()) {
                    double* d_rtmp = nullptr;
                    QrPreBlock* d_pre2 = nullptr;
                    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_rtmp), sizeof(double) * roff, s), "svd precondition R2");
                    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_pre2), sizeof(QrPreBlock) * pre2.size(), s), "svd precondition meta 2");
                    ctx.upload_async(d_pre2, pre2.data(), sizeof(QrPreBlock) * pre2.size());
                    launch_svd_precondition2_factor(R.d_W, d_aug, d_aug + boff, d_aug + boff + yoff, d_rtmp, d_pre2, pre2, s, &ctx.launches);
                    std::vector<GemmItem> items;   // W' = V0 * R2^T
                    for(const QrPreBlock& q : pre2) {
                        GemmItem g;
                        std::memset(&g, 0, sizeof(g));
                        g.a_off = q.v_off;
                        g.b_off = q.q_off;
                        g.c_off = q.w_off;
                        g.m = g.n = g.k = q.n;
                        g.lda = g.ldb = g.ldc = q.n;
                        g.ta = 0;
                        g.tb = 1;
                        g.alpha = 1.0;
                        items.push_back(g);
                    }
                    ctx.run_gemm(R.d_V, d_rtmp, R.d_W, items);
                    launch_svd_precondition2_q(d_aug, R.d_V, d_pre2, pre2, s, &ctx.launches);
                    for(int i : pre2_index) R.transposed[i] = !R.transposed[i];   // W now carries the other side's vectors
                    check_cuda(cudaFreeAsync(d_rtmp, s), "svd precondition free");
                    check_cuda(cudaFreeAsync(d_pre2, s), "svd precondition free");
                }
            }
            check_cuda(cudaFreeAsync(d_aug, s), "svd precondition free");
            check_cuda(cudaFreeAsync(d_pre, s), "svd precondition free");
        }
    }
    int* h_flag = static_cast<int*>(ctx.pinned(2 * sizeof(int)));
    unsigned long long* d_tprof = nullptr;
    static const bool svd_prof = std::getenv("XB_SVD_PROF") != nullptr;
    if(svd_prof) {
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_tprof), 8 * sizeof(unsigned long long), s), "svd prof");
        check_cuda(cudaMemsetAsync(d_tprof, 0, 8 * sizeof(unsigned long long), s), "svd prof reset");
    }
    int max_rows = 0;
    for(const SvdBlock& b : R.blocks) max_rows = std::max(max_rows, b.rows);
    // resident kernel: cluster of CL CTAs per panel pair, each keeps its row slice (<= rs rows) of the panel in shared memory
    // up to 640 rows two CTAs per pair (one CTA per SM, nothing shares its FP64 pipe) beat four at two per SM: 55.5 vs 57.8 ms per CTM step;
    // 8 CTAs (304-row slices, two CTAs per SM) up to 2432 rows; a 16-CTA cluster (non-portable size, 224-row slices) up to 3584 rows
    // (the D=6, chi=144 CTM blocks have ~3100); beyond that the chunked kernel
    // 8 CTAs per pair beyond 640 rows: 304-row slices (two CTAs per SM) up to 2432 rows; the chunked kernel beyond (the D=6, chi=144 CTM
    // blocks have ~3100 rows).  Measured on those blocks: chunked 0.57 s per tSVD, 416-row resident slices at one CTA per SM
    // (XB_SVD_WIDE=1) 0.91 s, a 16-CTA cluster (XB_SVD_CL16=1) 1.18 s -- both stay opt-in.
    static const bool cl16 = std::getenv("XB_SVD_CL16") != nullptr && std::atoi(std::getenv("XB_SVD_CL16")) != 0;
    static const bool no_wide = !(std::getenv("XB_SVD_WIDE") != nullptr && std::atoi(std::getenv("XB_SVD_WIDE")) != 0);
    const bool wide = !cl16 && !no_wide && max_rows > 2432 && max_rows <= 3328;
    const int cluster = max_rows <= 160 ? 1 : (max_rows <= 640 ? 2 : ((max_rows <= 2432 || !cl16) ? 8 : 16));
    const bool resident = max_rows <= (cl16 ? 3584 : (no_wide ? 2432 : 3328));
    int rs = 16;
    for(const SvdBlock& b : R.blocks) {
        const int per = (((b.rows + cluster - 1) / cluster) + 7) & ~7;   // cols <= rows, so the V slice is never longer
        rs = std::max(rs, (per + 15) & ~15);
    }
    const size_t dyn_smem = sizeof(double) * (static_cast<size_t>(SP) * (rs + 4) + (cluster >= 8 ? 2 : 4) * SP * SP + (cluster >= 8 ? 2 : 1) * SP * VL_LD + 2);   // <8, 40, false>: single-buffered planes, Vl in a region of its own
    const int n_rounds_total = static_cast<int>(round_begin.size()) - 1;
    auto launch_round = [&](const SvdTask* d_round_tasks, int cnt, int round_idx) {
        int* d_round_clean = d_clean + (d_round_tasks - d_tasks);
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(SVD_THREADS);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const SvdBlock* cb = d_blocks;
        if(!resident) {
            static const int chunked_cl = std::getenv("XB_SVD_CHUNKED_CL") ? std::atoi(std::getenv("XB_SVD_CHUNKED_CL")) : 4;   // experiment knob
            if(chunked_cl == 8 || chunked_cl == 2) {
                cfg.gridDim = dim3(static_cast<unsigned>(cnt) * chunked_cl);
                attr[0].val.clusterDim.x = chunked_cl;
                if(chunked_cl == 8)
                    check_cuda(cudaLaunchKernelEx(&cfg, svd_round_kernel<8>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, d_mod, d_round_clean,
                                                  static_cast<const int*>(d_sweep), round_idx, n_rounds_total),
                               "svd_round_kernel<8> launch");
                else
                    check_cuda(cudaLaunchKernelEx(&cfg, svd_round_kernel<2>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, d_mod, d_round_clean,
                                                  static_cast<const int*>(d_sweep), round_idx, n_rounds_total),
                               "svd_round_kernel<2> launch");
                return;
            }
            cfg.gridDim = dim3(static_cast<unsigned>(cnt) * 4);
            attr[0].val.clusterDim.x = 4;
            check_cuda(cudaLaunchKernelEx(&cfg, svd_round_kernel<4>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, d_mod, d_round_clean, static_cast<const int*>(d_sweep),
                                          round_idx, n_rounds_total),
                       "svd_round_kernel<4> launch");
            return;
        }
        cfg.gridDim = dim3(static_cast<unsigned>(cnt) * cluster);
        cfg.dynamicSmemBytes = dyn_smem;
        attr[0].val.clusterDim.x = cluster;
        if(cluster == 1)
            check_cuda(cudaLaunchKernelEx(&cfg, svd_round_resident_kernel<1, 20, true>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, rs, d_mod, d_round_clean,
                                          static_cast<const int*>(d_sweep), round_idx, n_rounds_total), "svd_round_resident_kernel<1> launch");
        else if(cluster == 2)
            check_cuda(cudaLaunchKernelEx(&cfg, svd_round_resident_kernel<2, 40, true>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, rs, d_mod, d_round_clean,
                                          static_cast<const int*>(d_sweep), round_idx, n_rounds_total), "svd_round_resident_kernel<2> launch");
        else if(cluster == 16)
            check_cuda(cudaLaunchKernelEx(&cfg, svd_round_resident_kernel<16, 28, false>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, rs, d_mod, d_round_clean,
                                          static_cast<const int*>(d_sweep), round_idx, n_rounds_total), "svd_round_resident_kernel<16> launch");
        else if(wide)
            check_cuda(cudaLaunchKernelEx(&cfg, svd_round_resident_kernel<8, 52, false>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, rs, d_mod, d_round_clean,
                                          static_cast<const int*>(d_sweep), round_idx, n_rounds_total), "svd_round_resident_kernel<8, 52> launch");
        else
            check_cuda(cudaLaunchKernelEx(&cfg, svd_round_resident_kernel<8, 40, false>, R.d_W, R.d_V, cb, d_round_tasks, d_flag, d_tprof, rs, d_mod, d_round_clean,
                                          static_cast<const int*>(d_sweep), round_idx, n_rounds_total), "svd_round_resident_kernel<8> launch");
    };
    {
        static unsigned long long configured = 0;
        if(first_use_on_current_device(configured)) {
            const int max_dyn = 200 * 1024;
            check_cuda(cudaFuncSetAttribute(svd_round_resident_kernel<1, 20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn), "svd smem attr");
            check_cuda(cudaFuncSetAttribute(svd_round_resident_kernel<2, 40, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn), "svd smem attr");
            check_cuda(cudaFuncSetAttribute(svd_round_resident_kernel<8, 40, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn), "svd smem attr");
            check_cuda(cudaFuncSetAttribute(svd_round_resident_kernel<8, 52, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn), "svd smem attr");
            check_cuda(cudaFuncSetAttribute(svd_round_resident_kernel<16, 28, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn), "svd smem attr");
            check_cuda(cudaFuncSetAttribute(svd_round_resident_kernel<16, 28, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "svd cluster attr");
        }
    }
    const int max_sweeps = 150;   // bail-out only
    // The rounds of a sweep are launch-latency-bound at CTM sizes: the sweep (flag reset + all rounds) is captured once into a CUDA
    // graph and replayed; the instantiated graph is kept per layout and reused as long as the pool returns the same buffers.
    static const bool no_graph = std::getenv("XB_SVD_NOGRAPH") != nullptr;
    Context::SweepGraph* sg = nullptr;
    int rounds_per_sweep = 0;
    for(size_t r = 0; r + 1 < round_begin.size(); ++r) rounds_per_sweep += round_begin[r + 1] > round_begin[r] ? 1 : 0;
    if(!no_graph && !svd_prof && rounds_per_sweep >= 8) {
        const std::string key = in.sig + (shifted ? "#eigh" : "#svd");
        if(ctx.sweep_graphs.size() >= 256 && ctx.sweep_graphs.find(key) == ctx.sweep_graphs.end()) {   // bound the cache (layouts that keep changing)
            for(auto& kv : ctx.sweep_graphs)
                if(kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
            ctx.sweep_graphs.clear();
        }
        Context::SweepGraph& g = ctx.sweep_graphs[key];
        if(g.exec == nullptr || g.w != R.d_W || g.v != R.d_V || g.meta != d_meta) {
            if(g.exec) cudaGraphExecDestroy(g.exec);
            g = Context::SweepGraph();
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if(ok) {
                try {
                    svd_sweep_bump_kernel<<<1, 1, 0, s>>>(d_sweep, d_flag);
                    for(size_t r = 0; r + 1 < round_begin.size(); ++r) {
                        const int cnt = round_begin[r + 1] - round_begin[r];
                        if(cnt > 0) launch_round(d_tasks + round_begin[r], cnt, static_cast<int>(r));
                    }
                } catch(...) {
                    ok = false;
                }
                ok = (cudaStreamEndCapture(s, &graph) == cudaSuccess) && ok && graph != nullptr;
            }
            if(ok) ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
            if(graph) cudaGraphDestroy(graph);
            if(ok) {
                g.w = R.d_W;
                g.v = R.d_V;
                g.meta = d_meta;
                g.rounds = rounds_per_sweep;
            } else {
                cudaGetLastError();   // clear; fall back to plain launches
                g = Context::SweepGraph();
            }
        }
        if(g.exec) sg = &g;
    }
    // Sweeps are enqueued in batches and the convergence record is read back once per batch: a sweep behind the converged one costs its (empty) launches
    // only, a read-back costs a host round trip with the device idle (~35 us against ~80 us of a D=4 sweep).  First batch = the caller's expectation.
    check_cuda(cudaMemsetAsync(d_flag, 0xff, 2 * sizeof(int), s), "svd flag init");   // flag[0] != 0: "rotated", flag[1] = -1: not converged
    const bool batched = rounds_per_sweep <= 16;
    int enqueued = 0, batch = batched ? std::max(1, std::min(in.sweeps_hint > 0 ? in.sweeps_hint : 2, max_sweeps)) : 1;
    bool converged = false;
    while(!converged && enqueued < max_sweeps) {
        batch = std::min(batch, max_sweeps - enqueued);
        for(int b = 0; b < batch; ++b) {
            if(sg) {
                check_cuda(cudaGraphLaunch(sg->exec, s), "svd sweep graph launch");
                ctx.launches += sg->rounds;
            } else {
                svd_sweep_bump_kernel<<<1, 1, 0, s>>>(d_sweep, d_flag);
                for(size_t r = 0; r + 1 < round_begin.size(); ++r) {
                    const int cnt = round_begin[r + 1] - round_begin[r];
                    if(cnt == 0) continue;
                    launch_round(d_tasks + round_begin[r], cnt, static_cast<int>(r));
                    ++ctx.launches;
                }
            }
        }
        enqueued += batch;
        check_cuda(cudaGetLastError(), "svd_round_kernel launch");
        check_cuda(cudaMemcpyAsync(h_flag, d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, s), "svd flag readback");
        check_cuda(cudaStreamSynchronize(s), "svd sweep sync");
        if(h_flag[1] >= 0) {   // found converged inside the batch: after h_flag[1] sweeps
            R.sweeps = h_flag[1] - 1;
            converged = true;
        } else if(h_flag[0] == 0) {   // the last sweep of the batch applied no rotation
            R.sweeps = enqueued - 1;
            converged = true;
        }
        batch = batched ? 2 : 1;
    }
    if(!converged) R.sweeps = max_sweeps;
    if(R.sweeps >= max_sweeps) {
        cudaFreeAsync(d_meta, s);
        cudaFreeAsync(R.d_W, s);
        cudaFreeAsync(R.d_V, s);
        cudaFreeAsync(R.d_sigma, s);
        throw std::runtime_error("tSVD: block Jacobi did not converge in 150 sweeps");
    }
    {
        dim3 grid(32, static_cast<unsigned>(std::min(nb, 64)));
        svd_norms_kernel<<<grid, 256, 0, s>>>(R.d_W, d_blocks, nb, d_sigma);
        check_cuda(cudaGetLastError(), "svd_norms_kernel launch");
        ++ctx.launches;
    }
    if(host_sigma) check_cuda(cudaMemcpyAsync(R.sigma.data(), d_sigma, sizeof(double) * soff, cudaMemcpyDeviceToHost, s), "sigma readback");
    if(shifted) {
        R.shift.resize(nb);
        check_cuda(cudaMemcpyAsync(R.shift.data(), d_shift, sizeof(double) * nb, cudaMemcpyDeviceToHost, s), "shift readback");
    }
    if(host_sigma || shifted) check_cuda(cudaStreamSynchronize(s), "svd sync");
    if(svd_prof) {
        unsigned long long h[8];
        check_cuda(cudaMemcpy(h, d_tprof, sizeof(h), cudaMemcpyDeviceToHost), "svd prof readback");
        check_cuda(cudaFreeAsync(d_tprof, s), "svd prof free");
        const double n = h[7] ? static_cast<double>(h[7]) : 1.0;
        std::fprintf(stderr, "[svd prof] blocks %d max_rows %d cluster %d rs %d sweeps %d rounds/sweep %zu | CTA0 cycles/round: p0 %.0f p1 %.0f p2 %.0f p3 %.0f p4 %.0f (updates %llu)\n",
                     nb, max_rows, cluster, rs, R.sweeps + 1, round_begin.size() - 1, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[7]);
    }
    check_cuda(cudaFreeAsync(d_meta, s), "svd meta free");
    return R;
}

void free_jacobi(Context& ctx, JacobiResult& R)
{
    if(R.d_W) cudaFreeAsync(R.d_W, ctx.stream);
    if(R.d_V) cudaFreeAsync(R.d_V, ctx.stream);
    if(R.d_sigma) cudaFreeAsync(R.d_sigma, ctx.stream);
    R.d_W = R.d_V = R.d_sigma = nullptr;
}

// ------------------------------------------------------------------------------------------------ tSVD
SvdResult Context::tsvd(const Tensor& t, int64_t max_keep, double eps_svd, bool preserve_multiplets)
{
    if(t.layout->sector.empty()) throw std::invalid_argument("tSVD: tensor has no blocks");
    SvdResult res;
    if(taping()) {
        // reverse rule of full_svd + cutoff (ADTensor.hpp:237-374): needs the full factors in the gauge of the truncated ones, which the deterministic full path
        // gives when tsvd_backward recomputes them -- so the forward pass takes the full path too
        {
            TapePause pause(*this);
            res = tsvd_full(t, max_keep, eps_svd, preserve_multiplets);
        }
        Tensor *ta = tape->keep(t), *tu = tape->keep(*res.U), *ts = tape->keep(*res.S), *tv = tape->keep(*res.Vd);
        tape->record([this, ta, tu, ts, tv](Tape& T) {
            Tensor *gu = T.grad(*tu), *gs = T.grad(*ts), *gv = T.grad(*tv);
            if(!gu && !gs && !gv) return;
            // Sigma is stored as dense diagonal blocks: only the diagonal of its adjoint is a derivative (a product downstream leaves values in the
            // structurally-zero off-diagonal entries too; the reference's rule would add them into dA, ADTensor.hpp:293-297)
            std::unique_ptr<Tensor> gs_diag(gs ? diag_map(*gs, 3) : nullptr);
            T.accumulate(*ta, tsvd_backward(*ta, *tu, *ts, *tv, gu, gs_diag.get(), gv));
        });
        return res;
    }
    if(tsvd_truncated(t, max_keep, eps_svd, preserve_multiplets, &res)) return res;
    return tsvd_full(t, max_keep, eps_svd, preserve_multiplets);
}

SvdResult Context::tsvd_full(const Tensor& t, int64_t max_keep, double eps_svd, bool preserve_multiplets)
{
    const Layout& L = *t.layout;
    if(L.sector.empty()) throw std::invalid_argument("tSVD: tensor has no blocks");
    ProfScope prof(*this, CAT_SVD, 0.0);
    int64_t n_sv = 0;
    for(size_t i = 0; i < L.sector.size(); ++i) n_sv += std::min(L.rows[i], L.cols[i]);
    if(n_sv == 0) throw std::invalid_argument("tSVD: tensor has no singular values");
    static const bool host_select = std::getenv("XB_SVD_HOSTSELECT") != nullptr;
    if(!host_select && svd_device_select_supported(n_sv, static_cast<int>(L.sector.size()))) {
        // singular values stay on the device: sort / truncation there, one small read-back (kept counts, truncWeight, entropy)
        JacobiResult Jd = jacobi_all(*this, jacobi_input(t), false, false);
        SvdResult r = svd_full_device_select(*this, t, Jd, max_keep, eps_svd, preserve_multiplets);
        free_jacobi(*this, Jd);
        return r;
    }
    // more values than the shared-memory sort holds: truncation decision on the host
    JacobiResult J = jacobi_all(*this, jacobi_input(t), false, true);
    struct SV { int sector; int col; int k; double s; };
    std::vector<SV> all;
    for(size_t i = 0; i < L.sector.size(); ++i) {
        const SvdBlock& b = J.blocks[i];
        // per-block descending order first (what LAPACK returns), ties by column index
        std::vector<int> idx(b.cols);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return J.sigma[b.s_off + x] > J.sigma[b.s_off + y]; });
        int k = 0;
        for(int c : idx) all.push_back({static_cast<int>(i), c, k++, J.sigma[b.s_off + c]});
    }
    if(all.empty()) throw std::invalid_argument("tSVD: tensor has no singular values");
    const SV first_entry = all[0];
    // Tensor.cpp:543-547; the reference's std::sort is unstable for ties -- here ties keep (sector, k) order.
    std::stable_sort(all.begin(), all.end(), [](const SV& a, const SV& b) { return a.s > b.s; });
    // Numerically tied values (|diff| <= 1e-12 relative; e.g. the q <-> -q partner sectors of a U(1) double layer) are
    // ordered by (sector index, k) so that the truncation decision does not depend on rounding noise.
    for(size_t i = 0; i < all.size();) {
        size_t j = i + 1;
        while(j < all.size() && all[j - 1].s - all[j].s <= 1.e-12 * all[i].s) ++j;
        if(j - i > 1)
            std::stable_sort(all.begin() + i, all.begin() + j, [](const SV& a, const SV& b) { return a.sector != b.sector ? a.sector < b.sector : a.k < b.k; });
        i = j;
    }
    SvdResult res;
    for(size_t i = static_cast<size_t>(std::max<int64_t>(max_keep, 0)); i < all.size(); ++i)   // :548
        res.trunc_weight += sym->degeneracy(L.sector[all[i].sector]) * all[i].s * all[i].s;
    if(static_cast<int64_t>(all.size()) > max_keep) all.resize(static_cast<size_t>(std::max<int64_t>(max_keep, 0)));
    if(eps_svd > 0.0) all.erase(std::remove_if(all.begin(), all.end(), [eps_svd](const SV& x) { return x.s < eps_svd; }), all.end());   // :551-557
    if(preserve_multiplets) {   // :569-584
        int end_of_multiplet = -1;
        for(int i = static_cast<int>(all.size()) - 1; i > 0; --i) {
            const double rel = 2 * (all[i - 1].s - all[i].s) / (all[i - 1].s + all[i].s);
            if(rel > 0.1) {
                end_of_multiplet = i;
                break;
            }
        }
        if(end_of_multiplet != -1) all.resize(end_of_multiplet);
    }
    if(all.empty()) all.push_back(first_entry);   // :586-590
    std::map<Q, std::vector<SV>> ordered;         // std::map => ascending qarray order (:592-598)
    for(const SV& x : all) {
        ordered[L.sector[x.sector]].push_back(x);
        res.entropy += -sym->degeneracy(L.sector[x.sector]) * x.s * x.s * std::log(x.s * x.s);
    }
    auto trunc = std::make_shared<Qbasis>(sym, 1);
    std::vector<Q> qs;
    for(const auto& [q, v] : ordered) {
        trunc->push_back(q, static_cast<int64_t>(v.size()));
        qs.push_back(q);
    }
    BasisPtr tb = trunc;
    LayoutPtr LU = layout_custom(L.dom, {tb}, qs), LS = layout_custom({tb}, {tb}, qs), LV = layout_custom({tb}, L.cod, qs);
    res.U = new Tensor(this, LU, false);
    res.S = new Tensor(this, LS, true);
    res.Vd = new Tensor(this, LV, false);
    std::vector<GatherCol> cols;
    for(size_t k = 0; k < qs.size(); ++k) {
        const auto& v = ordered[qs[k]];
        const int i = v[0].sector;
        const SvdBlock& b = J.blocks[i];
        const int nret = static_cast<int>(v.size());
        const int m = static_cast<int>(L.rows[i]), n = static_cast<int>(L.cols[i]);
        for(int r = 0; r < nret; ++r) {
            GatherCol g;
            g.w_src = b.w_off + static_cast<long long>(v[r].col) * b.rows;
            g.v_src = b.v_off + static_cast<long long>(v[r].col) * b.cols;
            g.wl = b.rows;
            g.vl = b.cols;
            g.u_dst = LU->off[k] + static_cast<long long>(r) * m;
            g.vd_dst = LV->off[k] + r;
            g.vd_ld = nret;
            g.s_dst = LS->off[k] + r + static_cast<long long>(r) * nret;
            g.w_is_u = J.transposed[i] ? 0 : 1;
            g.sigma = v[r].s;
            (void)n;
            cols.push_back(g);
        }
    }
    GatherCol* d_cols = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_cols), sizeof(GatherCol) * cols.size(), stream), "svd gather meta");
    upload_async(d_cols, cols.data(), sizeof(GatherCol) * cols.size());
    svd_gather_kernel<<<static_cast<unsigned>(std::min<size_t>(cols.size(), 2368)), 128, 0, stream>>>(J.d_W, J.d_V, res.U->d, res.S->d, res.Vd->d, d_cols,
                                                                                                      static_cast<int>(cols.size()));
    check_cuda(cudaGetLastError(), "svd_gather_kernel launch");
    ++launches;
    check_cuda(cudaFreeAsync(d_cols, stream), "svd gather free");
    free_jacobi(*this, J);
    return res;
}

// ------------------------------------------------------------------------------------------------ teigh
// Tensor::teigh (src/Core/Tensor.cpp:782-803; per block SelfAdjointEigenSolver, MatrixInterface_Eigen_impl.hpp).
// Device form: every symmetric block is shifted to A_q + c_q 1 (c_q = 1.5 ||A_q||_F, positive definite), for which the
// one-sided Jacobi SVD *is* the eigen-decomposition: W = (A + c) V, ||w_j|| = lambda_j + c.  Eigenvalues ascending
// like LAPACK/Eigen; eigenvector signs are a gauge freedom.
namespace {
struct EighCol { long long v_src, v_dst, d_dst; int n; double lambda; };

__global__ void eigh_gather_kernel(const double* __restrict__ V, double* __restrict__ Vout, double* __restrict__ D, const EighCol* __restrict__ cols, int n)
{
    for(int i = blockIdx.x; i < n; i += gridDim.x) {
        const EighCol g = cols[i];
        for(int r = threadIdx.x; r < g.n; r += blockDim.x) Vout[g.v_dst + r] = V[g.v_src + r];
        if(threadIdx.x == 0) D[g.d_dst] = g.lambda;
    }
}
}   // namespace

void Context::teigh(const Tensor& t, Tensor** Do, Tensor** Vo)
{
    const Layout& L = *t.layout;
    if(L.rank() != L.corank()) throw std::invalid_argument("teigh: Rank != CoRank");
    for(size_t i = 0; i < L.sector.size(); ++i)
        if(L.rows[i] != L.cols[i]) throw std::invalid_argument("teigh: blocks must be square (coupledDomain == coupledCodomain)");
    BasisPtr md = L.cdom->forget_history(), mc = L.ccod->forget_history();
    LayoutPtr LD = layout_custom({md}, {mc}, L.sector), LV = layout_custom(L.dom, {mc}, L.sector);
    std::unique_ptr<Tensor> D(new Tensor(this, LD, true)), V(new Tensor(this, LV, false));
    if(L.sector.empty() || L.size == 0) {
        *Do = D.release();
        *Vo = V.release();
        return;
    }
    ProfScope prof(*this, CAT_SVD, 0.0);
    JacobiResult J = jacobi_all(*this, jacobi_input(t), true, true);
    std::vector<EighCol> cols;
    for(size_t i = 0; i < L.sector.size(); ++i) {
        const SvdBlock& b = J.blocks[i];
        const int n = b.cols;
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return J.sigma[b.s_off + x] < J.sigma[b.s_off + y]; });
        for(int r = 0; r < n; ++r) {
            EighCol g;
            g.v_src = b.v_off + static_cast<long long>(idx[r]) * n;
            g.v_dst = LV->off[i] + static_cast<long long>(r) * n;
            g.d_dst = LD->off[i] + r + static_cast<long long>(r) * n;
            g.n = n;
            g.lambda = J.sigma[b.s_off + idx[r]] - J.shift[i];
            cols.push_back(g);
        }
    }
    if(!cols.empty()) {
        EighCol* d_cols = nullptr;
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_cols), sizeof(EighCol) * cols.size(), stream), "eigh gather meta");
        upload_async(d_cols, cols.data(), sizeof(EighCol) * cols.size());
        eigh_gather_kernel<<<static_cast<unsigned>(std::min<size_t>(cols.size(), 2368)), 128, 0, stream>>>(J.d_V, V->d, D->d, d_cols, static_cast<int>(cols.size()));
        check_cuda(cudaGetLastError(), "eigh_gather_kernel launch");
        ++launches;
        check_cuda(cudaFreeAsync(d_cols, stream), "eigh gather free");
    }
    free_jacobi(*this, J);
    *Do = D.release();
    *Vo = V.release();
}

// ------------------------------------------------------------------------------------------------ tQR
// Tensor::tQR (src/Core/Tensor.cpp:757-780): per block HouseholderQR (MatrixInterface_Eigen_impl.hpp qr), thin factors
// Q = Q_full * 1(m x n), R = 1(n x m) * R_full.  Device form: one CTA per sector block runs the unblocked Householder
// recurrence with LAPACK's dlarfg sign convention (beta = -sign(alpha) ||x||, the one Eigen's makeHouseholder uses too), so
// R's diagonal signs equal the reference's; the trailing update is spread over the CTA's warps (one column per warp,
// lanes along the contiguous rows).  All blocks of a tensor run concurrently in one launch.
namespace {
struct QrBlock { long long a_off, q_off, r_off, w_off, t_off; int m, n, mid, lq; };

__device__ double block_sum(double v, double* sm)
{
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for(int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s += sm[w];   // fixed order: deterministic
    return s;
}

__global__ void __launch_bounds__(256)
qr_householder_kernel(const double* __restrict__ A, double* __restrict__ W, double* __restrict__ Qo, double* __restrict__ Ro, double* __restrict__ taus,
                      const QrBlock* __restrict__ blocks)
{
    __shared__ double sm[8];
    const QrBlock d = blocks[blockIdx.x];
    const int m = d.m, n = d.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    double* __restrict__ w = W + d.w_off;       // m x n work copy, column-major
    const int mid = d.mid;                         // inner dimension of the middle basis in this sector (Tensor.cpp:760)
    double* __restrict__ tau = taus + d.t_off;   // min(m, n) scalars per block
    const long long cnt = static_cast<long long>(m) * n;
    for(long long e = tid; e < cnt; e += blockDim.x) w[e] = A[d.a_off + e];
    __syncthreads();
    const int kmax = m < n ? m : n;
    for(int j = 0; j < kmax; ++j) {
        double* __restrict__ x = w + j + static_cast<long long>(j) * m;
        const int len = m - j;
        double part = 0.0;
        for(int r = 1 + tid; r < len; r += blockDim.x) part = fma(x[r], x[r], part);
        const double tail = block_sum(part, sm);
        const double alpha = x[0];
        double beta = alpha, tj = 0.0, scale = 0.0;
        if(tail > 0.0) {   // dlarfg; tail == 0  =>  H = 1
            beta = -copysign(sqrt(alpha * alpha + tail), alpha);
            tj = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        __syncthreads();
        for(int r = 1 + tid; r < len; r += blockDim.x) x[r] *= scale;
        if(tid == 0) {
            x[0] = beta;
            tau[j] = tj;
        }
        __syncthreads();
        if(tj != 0.0) {
            for(int k = j + 1 + warp; k < n; k += nwarp) {   // A[j:, k] -= tau v (v^T A[j:, k]),  v = (1, x[1:])
                double* __restrict__ c = w + j + static_cast<long long>(k) * m;
                double dot = lane == 0 ? c[0] : 0.0;
                for(int r = 1 + lane; r < len; r += 32) dot = fma(x[r], c[r], dot);
                for(int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                const double f = tj * dot;
                if(lane == 0) c[0] -= f;
                for(int r = 1 + lane; r < len; r += 32) c[r] = fma(-f, x[r], c[r]);
            }
        }
        __syncthreads();
    }
    // R = 1(mid x m) * R_full : mid x n, rows >= min(m, n) are zero (the reference's 1(n x m) R_full whenever the middle basis is the codomain)
    double* __restrict__ R = Ro + d.r_off;
    double* __restrict__ Q = Qo + d.q_off;
    // Q = H_1 ... H_k * 1(m x mid)  (dorg2r recurrence, backwards), built directly in the output block
    const long long qcnt = static_cast<long long>(m) * mid;
    for(long long e = tid; e < qcnt; e += blockDim.x) {
        const int r = static_cast<int>(e % m), c = static_cast<int>(e / m);
        Q[e] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    for(int j = kmax - 1; j >= 0; --j) {
        const double tj = tau[j];
        if(tj == 0.0) continue;   // uniform across the CTA
        const double* __restrict__ x = w + j + static_cast<long long>(j) * m;
        const int len = m - j;
        for(int k = j + warp; k < mid; k += nwarp) {
            double* __restrict__ c = Q + j + static_cast<long long>(k) * m;
            double dot = lane == 0 ? c[0] : 0.0;
            for(int r = 1 + lane; r < len; r += 32) dot = fma(x[r], c[r], dot);
            for(int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const double f = tj * dot;
            if(lane == 0) c[0] -= f;
            for(int r = 1 + lane; r < len; r += 32) c[r] = fma(-f, x[r], c[r]);
        }
        __syncthreads();
    }
    if(!d.lq) {
        for(long long e = tid; e < static_cast<long long>(mid) * n; e += blockDim.x) {
            const int r = static_cast<int>(e % mid), c = static_cast<int>(e / mid);
            R[e] = (r <= c && r < m) ? w[r + static_cast<long long>(c) * m] : 0.0;
        }
    } else {
        // RETURN_LQ (Tensor.cpp:767-769): the adjoints of both factors; only reached for square blocks (m == n == mid)
        for(long long e = tid; e < static_cast<long long>(n) * n; e += blockDim.x) {
            const int r = static_cast<int>(e % n), c = static_cast<int>(e / n);
            R[e] = (c <= r) ? w[c + static_cast<long long>(r) * m] : 0.0;   // R^T: lower triangular
        }
        for(long long e = tid; e < qcnt; e += blockDim.x) {   // in-place transpose of the square Q
            const int r = static_cast<int>(e % m), c = static_cast<int>(e / m);
            if(r < c) {
                const double a = Q[r + static_cast<long long>(c) * m], b = Q[c + static_cast<long long>(r) * m];
                Q[r + static_cast<long long>(c) * m] = b;
                Q[c + static_cast<long long>(r) * m] = a;
            }
        }
    }
}
}   // namespace

// Tensor::tQR(RETURN_LQ) (src/Core/Tensor.cpp:757-780).  The reference multiplies the full Householder factors by 1(m x n) / 1(n x m), which gives blocks
// that fit the bases of Q (domain -> middle) and R (middle -> codomain) only when the middle basis is the coupled codomain; here the
// identities are 1(m x mid) / 1(mid x m) with mid = middle.inner_dim(q): the same matrices whenever the reference's are consistent, and
// Q (m x m, orthogonal) / R (m x n, upper trapezoidal) when the middle basis is the coupled domain.  RETURN_LQ returns the adjoint blocks
// (Q^T, R^T) like the reference; they fit the bases only for square blocks, anything else is XB_ERR_INVALID.
void Context::tqr(const Tensor& t, bool return_lq, Tensor** Qo, Tensor** Ro)
{
    const Layout& L = *t.layout;
    const bool tall = L.cdom->dim() > L.ccod->dim();
    BasisPtr middle = tall ? L.ccod->forget_history() : L.cdom->forget_history();   // Tensor.cpp:760
    if(return_lq)
        for(size_t i = 0; i < L.sector.size(); ++i)
            if(L.rows[i] != L.cols[i])
                throw std::invalid_argument("tQR(RETURN_LQ): the adjoint factors of a non-square sector block do not fit the bases of the result "
                                            "(the reference builds inconsistent tensors in this case, Tensor.cpp:767-769)");
    LayoutPtr LQ = layout_custom(L.dom, {middle}, L.sector), LR = layout_custom({middle}, L.cod, L.sector);
    std::unique_ptr<Tensor> Q(new Tensor(this, LQ, true)), R(new Tensor(this, LR, true));
    const int nb = static_cast<int>(L.sector.size());
    if(nb == 0 || L.size == 0) {
        *Qo = Q.release();
        *Ro = R.release();
        return;
    }
    std::vector<QrBlock> blocks;
    long long toff = 0;
    for(int i = 0; i < nb; ++i) {
        if(L.rows[i] == 0 || L.cols[i] == 0) continue;
        QrBlock b;
        b.a_off = L.off[i];
        b.q_off = LQ->off[i];
        b.r_off = LR->off[i];
        b.w_off = L.off[i];
        b.t_off = toff;
        b.m = static_cast<int>(L.rows[i]);
        b.n = static_cast<int>(L.cols[i]);
        b.mid = static_cast<int>(middle->inner_dim(L.sector[i]));
        b.lq = return_lq ? 1 : 0;
        toff += std::min(b.m, b.n);
        blocks.push_back(b);
    }
    ProfScope prof(*this, CAT_OTHER, 0.0);
    double *d_W = nullptr, *d_tau = nullptr;
    QrBlock* d_blocks = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_W), sizeof(double) * L.size, stream), "tQR work");
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_tau), sizeof(double) * std::max<long long>(toff, 1), stream), "tQR tau");
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_blocks), sizeof(QrBlock) * blocks.size(), stream), "tQR meta");
    upload_async(d_blocks, blocks.data(), sizeof(QrBlock) * blocks.size());
    qr_householder_kernel<<<static_cast<unsigned>(blocks.size()), 256, 0, stream>>>(t.d, d_W, Q->d, R->d, d_tau, d_blocks);
    check_cuda(cudaGetLastError(), "qr_householder_kernel launch");
    ++launches;
    check_cuda(cudaFreeAsync(d_W, stream), "tQR free");
    check_cuda(cudaFreeAsync(d_tau, stream), "tQR free");
    check_cuda(cudaFreeAsync(d_blocks, stream), "tQR free");
    *Qo = Q.release();
    *Ro = R.release();
}

}   // namespace xb
