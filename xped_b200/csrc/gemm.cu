// Grouped FP64 GEMM for the sector products of TensorBase::operator* (src/Core/TensorBase.cpp:262-279):
//   C_q = alpha * op(A_q) * op(B_q) (+ C_q)   for a ragged list of (m_q, n_q, k_q), one launch for all sectors.
//
// Blackwell's tcgen05 tensor pipe has no FP64 kind, so the product runs on the DMMA path (mma.sync m8n8k4, the native
// SASS shape `DMMA.8x8x4` on sm_100a).  CTA tile BM x 64 with BM = 128 (2x2 warps, warp tile 64x32) or 64 (warp tile
// 32x32) for small sectors; K is streamed in slabs of 8 through a 4-stage cp.async (LDGSTS) ring so that global/L2
// latency is hidden behind the tensor pipe; operands are kept in shared memory with their globally contiguous
// dimension contiguous (no transposing stores) and row pitches == 4 (mod 16) doubles, which makes every m8n8k4
// fragment load hit 16 distinct banks per half-warp for both the normal and the adjoint (transposed) operand layout.
#include <cuda_runtime.h>

#include <algorithm>

#include <cstdint>

#include "engine.hpp"

namespace xb {

namespace {

constexpr int BN = 64, BK = 8, STAGES = 4, GEMM_THREADS = 128;
constexpr int PITCH_K = BK + 4;   // rows of length BK (k contiguous):   12 == 12 (mod 16) -> {0..3,12..15,8..11,4..7}
template <int BM> struct Cfg
{
    static constexpr int PITCH_M = BM + 4;   // rows of length BM (m contiguous): == 4 (mod 16)
    static constexpr int PITCH_N = BN + 4;
    static constexpr int A_ELEMS = (BK * PITCH_M > BM * PITCH_K) ? BK * PITCH_M : BM * PITCH_K;
    static constexpr int B_ELEMS = (BK * PITCH_N > BN * PITCH_K) ? BK * PITCH_N : BN * PITCH_K;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t SMEM_BYTES = sizeof(double) * STAGE_ELEMS * STAGES;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 8-byte async copy global -> shared; src_bytes = 0 zero-fills (out-of-range elements of ragged tiles)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid)
{
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, bool valid)
{
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int BM>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_grouped_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, const GemmItem* __restrict__ items)
{
    using cfg = Cfg<BM>;
    constexpr int WM = BM / 2, WN = BN / 2;   // warp tile
    constexpr int TM = WM / 8, TN = WN / 8;   // DMMA tiles per warp
    extern __shared__ __align__(16) double smem[];

    const GemmItem it = items[blockIdx.x];
    const double* __restrict__ Ap = A + it.a_off;
    const double* __restrict__ Bp = B + it.b_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 1) * WM, wn = (warp >> 1) * WN;
    const int m0 = it.tm, n0 = it.tn;
    const int fr = lane >> 2, fk = lane & 3;
    // 16-byte copies along the contiguous dimension when the block origin and leading dimension allow it
    const bool a16 = ((it.a_off | it.lda) & 1) == 0, b16 = ((it.b_off | it.ldb) & 1) == 0;

    double acc[TM][TN][2];
#pragma unroll
    for(int i = 0; i < TM; ++i)
#pragma unroll
        for(int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // One operand slab: `X` is the extent along the non-K dimension (BM or BN), stored either as [k][x] (x contiguous in
    // global memory: normal A, transposed B) or as [x][k] (k contiguous: transposed A, normal B).
    auto load_operand = [&](double* __restrict__ dst, const double* __restrict__ P, int ld, bool x_contig, bool vec, int x0, int xmax, int k0, int X,
                            int pitch_x) {
        if(x_contig) {   // element (x, kk) at P[x + kk * ld]  ->  dst[kk * pitch_x + x]
            if(vec) {
                for(int e = tid; e < (X / 2) * BK; e += GEMM_THREADS) {
                    const int x = (e % (X / 2)) * 2, kk = e / (X / 2);
                    const int gx = x0 + x, gk = k0 + kk;
                    const bool ok = gx + 1 < xmax && gk < it.k;   // (x, x+1) both valid; a lone tail element is handled below
                    if(ok || gx >= xmax || gk >= it.k) {
                        cp_async16(dst + kk * pitch_x + x, P + (ok ? gx + static_cast<long long>(gk) * ld : 0), ok);
                    } else {   // gx == xmax - 1: one valid element, one padding element
                        cp_async8(dst + kk * pitch_x + x, P + gx + static_cast<long long>(gk) * ld, true);
                        cp_async8(dst + kk * pitch_x + x + 1, P, false);
                    }
                }
            } else {
                for(int e = tid; e < X * BK; e += GEMM_THREADS) {
                    const int x = e % X, kk = e / X;
                    const int gx = x0 + x, gk = k0 + kk;
                    const bool ok = gx < xmax && gk < it.k;
                    cp_async8(dst + kk * pitch_x + x, P + (ok ? gx + static_cast<long long>(gk) * ld : 0), ok);
                }
            }
        } else {   // element (x, kk) at P[kk + x * ld]  ->  dst[x * PITCH_K + kk]
            if(vec) {
                for(int e = tid; e < X * (BK / 2); e += GEMM_THREADS) {
                    const int kk = (e % (BK / 2)) * 2, x = e / (BK / 2);
                    const int gx = x0 + x, gk = k0 + kk;
                    const bool ok = gx < xmax && gk + 1 < it.k;
                    if(ok || gx >= xmax || gk >= it.k) {
                        cp_async16(dst + x * PITCH_K + kk, P + (ok ? gk + static_cast<long long>(gx) * ld : 0), ok);
                    } else {
                        cp_async8(dst + x * PITCH_K + kk, P + gk + static_cast<long long>(gx) * ld, true);
                        cp_async8(dst + x * PITCH_K + kk + 1, P, false);
                    }
                }
            } else {
                for(int e = tid; e < X * BK; e += GEMM_THREADS) {
                    const int kk = e % BK, x = e / BK;
                    const int gx = x0 + x, gk = k0 + kk;
                    const bool ok = gx < xmax && gk < it.k;
                    cp_async8(dst + x * PITCH_K + kk, P + (ok ? gk + static_cast<long long>(gx) * ld : 0), ok);
                }
            }
        }
    };
    auto load_stage = [&](int stage, int k0) {
        double* sa = smem + stage * cfg::STAGE_ELEMS;
        double* sb = sa + cfg::A_ELEMS;
        load_operand(sa, Ap, it.lda, !it.ta, a16, m0, it.m, k0, BM, cfg::PITCH_M);
        load_operand(sb, Bp, it.ldb, it.tb != 0, b16, n0, it.n, k0, BN, cfg::PITCH_N);
    };

    const int nk = (it.k + BK - 1) / BK;
#pragma unroll
    for(int s = 0; s < STAGES - 1; ++s) {
        if(s < nk) load_stage(s, s * BK);
        cp_async_commit();
    }
    // fragment addressing: A[row][k]: normal -> sa[k * PITCH_M + row], transposed -> sa[row * PITCH_K + k]; same for B
    const int a_rs = it.ta ? PITCH_K : 1, a_ks = it.ta ? 1 : cfg::PITCH_M;
    const int b_cs = it.tb ? 1 : PITCH_K, b_ks = it.tb ? cfg::PITCH_N : 1;
    for(int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();   // slab kt has landed for every thread; slab kt-1's buffer is free again
        {
            const int nxt = kt + STAGES - 1;
            if(nxt < nk) load_stage(nxt % STAGES, nxt * BK);
            cp_async_commit();
        }
        const double* sa = smem + (kt % STAGES) * cfg::STAGE_ELEMS;
        const double* sb = sa + cfg::A_ELEMS;
#pragma unroll
        for(int ks = 0; ks < BK; ks += 4) {
            double a[TM], b[TN];
#pragma unroll
            for(int i = 0; i < TM; ++i) a[i] = sa[(wm + i * 8 + fr) * a_rs + (ks + fk) * a_ks];
#pragma unroll
            for(int j = 0; j < TN; ++j) b[j] = sb[(wn + j * 8 + fr) * b_cs + (ks + fk) * b_ks];
#pragma unroll
            for(int i = 0; i < TM; ++i)
#pragma unroll
                for(int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
    double* __restrict__ Cp = C + it.c_off;
#pragma unroll
    for(int i = 0; i < TM; ++i)
#pragma unroll
        for(int j = 0; j < TN; ++j) {
            const int gm = m0 + wm + i * 8 + fr;
            const int gn = n0 + wn + j * 8 + 2 * fk;
            if(gm < it.m) {
#pragma unroll
                for(int h = 0; h < 2; ++h)
                    if(gn + h < it.n) {
                        double* p = Cp + gm + static_cast<long long>(gn + h) * it.ldc;
                        const double v = it.alpha * acc[i][j][h];
                        *p = it.acc ? *p + v : v;
                    }
            }
        }
}

template <int BM> void launch_one(const double* A, const double* B, double* C, const GemmItem* d_items, int n_items, cudaStream_t s)
{
    static unsigned long long configured = 0;
    if(first_use_on_current_device(configured)) {
        check_cuda(cudaFuncSetAttribute(gemm_grouped_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Cfg<BM>::SMEM_BYTES)),
                   "cudaFuncSetAttribute(gemm smem)");
    }
    gemm_grouped_kernel<BM><<<n_items, GEMM_THREADS, Cfg<BM>::SMEM_BYTES, s>>>(A, B, C, d_items);
    check_cuda(cudaGetLastError(), "gemm_grouped_kernel launch");
}

}   // namespace

// split-K epilogue: the K slices of a problem were written to scratch by gemm_grouped_kernel; summed here in slice order (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const double* __restrict__ scratch, double* __restrict__ C, const SplitKJob* __restrict__ jobs)
{
    const SplitKJob j = jobs[blockIdx.y];
    const int cnt = j.m * j.n, stride = gridDim.x * blockDim.x;   // split problems have <= 32 tiles or few waves of tiles: cnt < 2^31
    const double* __restrict__ sc = scratch + j.s_off;
    for(int e = blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += stride) {
        // all slices of the element are in flight at once (<= 8 slices; a rolled loop walked them at one L2 latency each: 12 us per launch at D=4, 51 us at
        // D=6 = 9-12 % of a CTM step); summed in slice order as before
        double v[8];
#pragma unroll
        for(int s = 0; s < 8; ++s) v[s] = s < j.split ? sc[static_cast<long long>(s) * cnt + e] : 0.0;
        double sum = 0.0;
#pragma unroll
        for(int s = 0; s < 8; ++s)
            if(s < j.split) sum += v[s];
        for(int s = 8; s < j.split; ++s) sum += sc[static_cast<long long>(s) * cnt + e];
        const int col = e / j.m;
        C[j.c_off + (e - col * j.m) + static_cast<long long>(col) * j.ldc] = sum;
    }
}

void launch_splitk_reduce(const double* scratch, double* C, const SplitKJob* d_jobs, int n_jobs, int max_job_elems, cudaStream_t s)
{
    if(n_jobs == 0) return;
    // one element per thread up to 2048 blocks per job (the largest job of a D=6 subspace iteration has 490 k elements)
    const int gx = std::max(1, std::min(2048, (max_job_elems + 255) / 256));
    splitk_reduce_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_jobs)), 256, 0, s>>>(scratch, C, d_jobs);
    check_cuda(cudaGetLastError(), "splitk_reduce_kernel launch");
}

void launch_gemm_grouped(const double* A, const double* B, double* C, const GemmItem* d_items, int n_items, int tile_m, cudaStream_t s)
{
    if(n_items == 0) return;
    if(tile_m == 128)
        launch_one<128>(A, B, C, d_items, n_items, s);
    else
        launch_one<64>(A, B, C, d_items, n_items, s);
}

}   // namespace xb
