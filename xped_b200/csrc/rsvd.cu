// tSVD drivers: device-wide singular-value sort + global truncation (Tensor::tSVD, src/Core/Tensor.cpp:543-598) and the
// truncated algorithm used when only a few singular values out of many are kept (CTM projectors: chi of ~10^4).
//
// 1. svd_select_kernel -- ONE CTA sorts the singular values of all sectors (bitonic network in shared memory, <= 16384 values),
//    applies the reference's truncation rules (keep the maxKeep largest, drop values < eps_svd, PRESERVE_MULTIPLETS, never return an
//    empty basis), accumulates truncWeight and the entropy, and writes per sector the kept count and the column order.  The host
//    reads back one small record (kept counts, two scalars) to build the output bases; the gather kernel takes the column
//    lists from device memory.
// 2. Context::tsvd_truncated -- randomized block subspace iteration per sector (all sectors batched in every launch):
//       X <- orth(B^T orth(B X))  with B_q the tall orientation of the sector block and X_q of width K = maxKeep + 32 (<= 160),
//    orthonormalisation by (shifted) Cholesky-QR on grouped DMMA GEMMs, then the batched one-sided Jacobi SVD of the projected
//    blocks Y_q = B_q X_q (tall, K columns) gives Ritz triplets.  A-posteriori check on the device: the residuals
//    ||B^T u - sigma v|| of every kept triplet (<= 16 eps sqrt(m) sigma_max), and at least 8 Ritz values below the cut in every
//    projected sector (the cut lies inside the computed part of each spectrum).  Not satisfied => more iterations (count predicted
//    from the Ritz spectrum), and if the budget is exhausted the caller falls back to the full Jacobi path.  Sectors with
//    min(m, n) <= K are decomposed exactly (X = 1).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>

#include "engine.hpp"
#include "svd_internal.hpp"

namespace xb {

namespace {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAX = 16384;   // capacity of the shared-memory sort
constexpr int SEL_COL_BITS = 14;
constexpr int NORM_PARTS = 128;   // (32 parts: 272 us for the 155 MB matrix of a D=6 chi=144 call, one load in flight per thread)

struct SelSector
{
    long long s_off;   // first singular value of the sector in the sigma array
    int count;
    int direct;        // 1: exact values (no residual / exhaustion check)
    double deg;        // Symmetry::degeneracy(q)
};

struct SelParams
{
    long long max_keep;
    double eps_svd;
    double tol_rel;   // residual tolerance relative to sigma_max (ritz mode)
    int preserve, ritz, nsec, n_total, guard, pad;
};

struct SelHeader
{
    int nkeep, verdict, nkeep0, pad;
    double trunc_weight, entropy, cut, max_res, rate, smax;
};

__device__ __forceinline__ bool before_a(double ka, int pa, double kb, int pb)
{   // (sector asc, sigma desc, column asc)
    const int sa = pa >> SEL_COL_BITS, sb = pb >> SEL_COL_BITS;
    if(sa != sb) return sa < sb;
    if(ka != kb) return ka > kb;
    return pa < pb;
}
__device__ __forceinline__ bool before_b(double ka, int pa, double kb, int pb)
{   // (sigma desc, sector asc, rank inside the sector asc)
    if(ka != kb) return ka > kb;
    return pa < pb;
}

template <bool MODE_A> __device__ void bitonic_sort(double* key, int* pay, int N)
{
    for(int k = 2; k <= N; k <<= 1) {
        for(int j = k >> 1; j > 0; j >>= 1) {
            for(int i = threadIdx.x; i < N; i += blockDim.x) {
                const int x = i ^ j;
                if(x > i) {
                    const double ki = key[i], kx = key[x];
                    const int pi = pay[i], px = pay[x];
                    const bool up = (i & k) == 0;
                    const bool x_first = MODE_A ? before_a(kx, px, ki, pi) : before_b(kx, px, ki, pi);
                    if(x_first == up) {
                        key[i] = kx;
                        key[x] = ki;
                        pay[i] = px;
                        pay[x] = pi;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ double block_sum_det(double v, double* red)
{   // deterministic: fixed tree
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if(threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for(int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if(threadIdx.x == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

__global__ void __launch_bounds__(SEL_THREADS)
svd_select_kernel(const double* __restrict__ sigma, const SelSector* __restrict__ secs, const double* __restrict__ res, const double* __restrict__ normpart,
                  SelParams P, int N, int* __restrict__ perm, SelHeader* __restrict__ hdr, int* __restrict__ kept)
{
    extern __shared__ __align__(16) unsigned char raw[];
    double* key = reinterpret_cast<double*>(raw);
    double* red = key + N;                               // 40
    int* pay = reinterpret_cast<int*>(red + 40);         // N
    int* soff = pay + N;                                 // nsec + 1
    int* cnt1 = soff + P.nsec + 1;                       // entries of each sector among the first nkeep1 (count + eps filter)
    int* cnt2 = cnt1 + P.nsec;                           // kept entries of each sector
    __shared__ int sh_i[8];
    const int tid = threadIdx.x, T = blockDim.x, n = P.n_total;

    for(int q = tid; q < P.nsec; q += T) {
        soff[q] = static_cast<int>(secs[q].s_off);
        cnt1[q] = 0;
        cnt2[q] = 0;
    }
    if(tid == 0) soff[P.nsec] = n;
    __syncthreads();
    for(int i = tid; i < N; i += T) {
        if(i < n) {
            int lo = 0, hi = P.nsec;   // last sector with soff <= i
            while(hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if(soff[mid] <= i) lo = mid; else hi = mid;
            }
            // empty sectors share an offset with their successor: move to the last one starting at soff[lo]
            key[i] = sigma[i];
            pay[i] = (lo << SEL_COL_BITS) | (i - soff[lo]);
        } else {
            key[i] = -1.0;
            pay[i] = INT_MAX;
        }
    }
    __syncthreads();
    bitonic_sort<true>(key, pay, N);
    // per-sector descending order: position p <-> (sector, k = p - soff); remember the column of every (sector, k)
    for(int p = tid; p < n; p += T) {
        const int sec = pay[p] >> SEL_COL_BITS, col = pay[p] & ((1 << SEL_COL_BITS) - 1);
        perm[p] = col;
        pay[p] = (sec << SEL_COL_BITS) | (p - soff[sec]);
    }
    __syncthreads();
    bitonic_sort<false>(key, pay, N);

    const int nkeep0 = static_cast<int>(P.max_keep < 0 ? 0 : (P.max_keep < n ? P.max_keep : n));
    if(tid == 0) {
        // numerically tied values (<= 1e-12 relative) around the cut are ordered by (sector, k): the decision does not depend on rounding noise
        if(nkeep0 > 0 && nkeep0 < n && key[nkeep0 - 1] - key[nkeep0] <= 1.e-12 * key[nkeep0 - 1]) {
            int i = nkeep0 - 1, j = nkeep0 + 1;
            while(i > 0 && key[i - 1] - key[i] <= 1.e-12 * key[i - 1]) --i;
            while(j < n && key[j - 1] - key[j] <= 1.e-12 * key[i]) ++j;
            for(int a = i + 1; a < j; ++a) {
                const double ka = key[a];
                const int pa = pay[a];
                int b = a - 1;
                while(b >= i && pay[b] > pa) {
                    key[b + 1] = key[b];
                    pay[b + 1] = pay[b];
                    --b;
                }
                key[b + 1] = ka;
                pay[b + 1] = pa;
            }
        }
        // eps filter (Tensor.cpp:551-557): the dropped values are the tail of the first nkeep0
        int lo = 0, hi = nkeep0;   // first position with key < eps
        if(P.eps_svd > 0.0) {
            while(lo < hi) {
                const int mid = (lo + hi) >> 1;
                if(key[mid] < P.eps_svd) hi = mid; else lo = mid + 1;
            }
        } else
            lo = nkeep0;
        sh_i[0] = lo;    // nkeep1
        sh_i[1] = -1;    // end of multiplet
    }
    __syncthreads();
    const int nkeep1 = sh_i[0];
    if(P.preserve) {   // Tensor.cpp:569-584: cut at the last gap > 0.1 (relative)
        int best = -1;
        for(int i = tid + 1; i < nkeep1; i += T) {
            const double rel = 2.0 * (key[i - 1] - key[i]) / (key[i - 1] + key[i]);
            if(rel > 0.1) best = i;
        }
        if(best >= 0) atomicMax(&sh_i[1], best);
        __syncthreads();
    }
    int nkeep = (P.preserve && sh_i[1] != -1) ? sh_i[1] : nkeep1;
    for(int p = tid; p < nkeep1; p += T) {
        const int sec = pay[p] >> SEL_COL_BITS;
        atomicAdd(&cnt1[sec], 1);
        if(p < nkeep) atomicAdd(&cnt2[sec], 1);
    }
    // truncWeight (Tensor.cpp:548): everything beyond the first maxKeep values
    double tw = 0.0, ent = 0.0;
    if(P.ritz) {
        for(int p = tid; p < nkeep0; p += T) tw -= secs[pay[p] >> SEL_COL_BITS].deg * key[p] * key[p];
        for(int q = tid; q < P.nsec; q += T) {
            double tot = 0.0;
            for(int k = 0; k < NORM_PARTS; ++k) tot += normpart[q * NORM_PARTS + k];
            tw += secs[q].deg * tot;
        }
    } else {
        for(int p = nkeep0 + tid; p < n; p += T) tw += secs[pay[p] >> SEL_COL_BITS].deg * key[p] * key[p];
    }
    tw = block_sum_det(tw, red);
    if(nkeep == 0 && n > 0) {   // Tensor.cpp:586-590: keep the first entry of the unsorted list = largest value of the first non-empty sector
        if(tid == 0) {
            int q0 = 0;
            while(q0 < P.nsec && secs[q0].count == 0) ++q0;
            cnt2[q0] = 1;
            const double s = sigma[soff[q0] + perm[soff[q0]]];
            sh_i[2] = q0;
            red[33] = -secs[q0].deg * s * s * log(s * s);
        }
        __syncthreads();
        ent = red[33];
        nkeep = 1;
    } else {
        for(int p = tid; p < nkeep; p += T) {
            const double s = key[p];
            ent += -secs[pay[p] >> SEL_COL_BITS].deg * s * s * log(s * s);
        }
        ent = block_sum_det(ent, red);
    }
    __syncthreads();
    // a-posteriori check of Ritz triplets
    int verdict = 0;
    double max_res = 0.0, rate = 0.0;
    const double smax = n > 0 ? key[0] : 0.0;
    const double cutref = nkeep1 > 0 ? key[nkeep1 - 1] : smax;
    if(P.ritz) {
        double mr = 0.0;
        const int ncheck = min(n, nkeep1 + 2);
        for(int p = tid; p < ncheck; p += T) {
            const int sec = pay[p] >> SEL_COL_BITS, k = pay[p] & ((1 << SEL_COL_BITS) - 1);
            if(secs[sec].direct) continue;
            const double r = res[soff[sec] + perm[soff[sec] + k]];
            if(p < nkeep1)
                mr = fmax(mr, r);
            else if(r > 1.e-3 * cutref)
                mr = fmax(mr, 1.e300);   // a discarded value next to the cut that is far from converged could still overtake it
        }
        // max via ordered-int trick is unnecessary: reduce with fmax (order independent)
        for(int o = 16; o > 0; o >>= 1) mr = fmax(mr, __shfl_xor_sync(0xffffffffu, mr, o));
        __syncthreads();
        if((tid & 31) == 0) red[tid >> 5] = mr;
        __syncthreads();
        if(tid == 0) {
            double m = 0.0;
            for(int w = 0; w < (T >> 5); ++w) m = fmax(m, red[w]);
            red[34] = m;
            double rt = 0.0;
            int exhausted = 0;
            for(int q = 0; q < P.nsec; ++q) {
                if(secs[q].direct || secs[q].count == 0) continue;
                if(secs[q].count - cnt1[q] < P.guard) exhausted = 1;
                const double smin = sigma[soff[q] + perm[soff[q] + secs[q].count - 1]];
                rt = fmax(rt, cutref > 0.0 ? smin / cutref : 1.0);
            }
            red[35] = rt;
            sh_i[3] = exhausted;
        }
        __syncthreads();
        max_res = red[34];
        rate = red[35];
        if(!(max_res <= P.tol_rel * smax)) verdict |= 1;
        if(sh_i[3]) verdict |= 2;
    }
    for(int q = tid; q < P.nsec; q += T) kept[q] = cnt2[q];
    if(tid == 0) {
        hdr->nkeep = nkeep;
        hdr->verdict = verdict;
        hdr->nkeep0 = nkeep0;
        hdr->pad = 0;
        hdr->trunc_weight = tw < 0.0 ? 0.0 : tw;
        hdr->entropy = ent;
        hdr->cut = cutref;
        hdr->max_res = smax > 0.0 ? max_res / smax : 0.0;
        hdr->rate = rate;
        hdr->smax = smax;
    }
}

struct GatherSec
{
    long long w_off, v_off, u_dst, vd_dst, s_dst, sel_off;
    int wl, vl, ldw, ldv, nret, w_is_u;
};

// writes the kept columns: U (or V^dagger rows) from W / sigma, the other factor from V, Sigma on the diagonal
__global__ void svd_gather_sel_kernel(const double* __restrict__ W, const double* __restrict__ V, const double* __restrict__ sigma, const int* __restrict__ perm,
                                      double* __restrict__ U, double* __restrict__ S, double* __restrict__ Vd, const GatherSec* __restrict__ secs)
{
    const GatherSec g = secs[blockIdx.y];
    for(int r = blockIdx.x; r < g.nret; r += gridDim.x) {
        const int col = perm[g.sel_off + r];
        const double sg = sigma[g.sel_off + col];
        const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
        const double* w = W + g.w_off + static_cast<long long>(col) * g.ldw;
        const double* v = V + g.v_off + static_cast<long long>(col) * g.ldv;
        if(g.w_is_u) {
            for(int i = threadIdx.x; i < g.wl; i += blockDim.x) U[g.u_dst + static_cast<long long>(r) * g.wl + i] = w[i] * inv;
            for(int i = threadIdx.x; i < g.vl; i += blockDim.x) Vd[g.vd_dst + r + static_cast<long long>(i) * g.nret] = v[i];
        } else {
            for(int i = threadIdx.x; i < g.vl; i += blockDim.x) U[g.u_dst + static_cast<long long>(r) * g.vl + i] = v[i];
            for(int i = threadIdx.x; i < g.wl; i += blockDim.x) Vd[g.vd_dst + r + static_cast<long long>(i) * g.nret] = w[i] * inv;
        }
        if(threadIdx.x == 0) S[g.s_dst + r + static_cast<long long>(r) * g.nret] = sg;
    }
}

// ---- pieces of the truncated path --------------------------------------------------------------------------------------------
struct RsvdSec
{
    long long a_off, x_off, y_off, g_off, s_off;
    int m, n, tr, mb, nb, K, direct, pad;
};

__global__ void rsvd_identity_kernel(double* __restrict__ X, const RsvdSec* __restrict__ secs, const int* __restrict__ list, int nlist)
{
    const RsvdSec s = secs[list[blockIdx.y]];
    const long long cnt = static_cast<long long>(s.nb) * s.K;
    for(long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < cnt; e += static_cast<long long>(gridDim.x) * blockDim.x)
        X[s.x_off + e] = (e % s.nb == e / s.nb) ? 1.0 : 0.0;
}

// warm start: X[:, r] = previous Ritz vector with the r-th largest value of its sector (column order from the previous selection); the last `nrand`
// columns of every projected sector are fresh random vectors (a direction the previous subspace lacked shows up through them)
__global__ void rsvd_warm_start_kernel(double* __restrict__ X, const double* __restrict__ Vprev, const int* __restrict__ perm_prev, const RsvdSec* __restrict__ secs,
                                       int nrand, unsigned long long seed)
{
    const RsvdSec s = secs[blockIdx.y];
    for(int r = blockIdx.x; r < s.K; r += gridDim.x) {
        double* dst = X + s.x_off + static_cast<long long>(r) * s.nb;
        if(s.direct) {
            for(int i = threadIdx.x; i < s.nb; i += blockDim.x) dst[i] = i == r ? 1.0 : 0.0;
        } else if(r < s.K - nrand) {
            const double* src = Vprev + s.x_off + static_cast<long long>(perm_prev[s.s_off + r]) * s.nb;
            for(int i = threadIdx.x; i < s.nb; i += blockDim.x) dst[i] = src[i];
        } else {
            for(int i = threadIdx.x; i < s.nb; i += blockDim.x) {
                unsigned long long z = seed ^ (static_cast<unsigned long long>(s.x_off + static_cast<long long>(r) * s.nb + i) * 0xD1342543DE82EF95ull);
                z += 0x9E3779B97F4A7C15ull;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                dst[i] = static_cast<double>(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
            }
        }
    }
}

// partial squared Frobenius norms of the sector blocks (fixed partition => deterministic)
__global__ void rsvd_norm2_kernel(const double* __restrict__ A, const RsvdSec* __restrict__ secs, double* __restrict__ part)
{
    __shared__ double sm[33];
    const RsvdSec s = secs[blockIdx.y];
    const long long cnt = static_cast<long long>(s.m) * s.n;
    const long long chunk = (cnt + gridDim.x - 1) / gridDim.x;
    const long long e0 = chunk * blockIdx.x, e1 = e0 + chunk < cnt ? e0 + chunk : cnt;
    double v = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;   // four chains: four loads in flight per thread (fixed order => deterministic)
    long long e = e0 + threadIdx.x;
    const long long bd = blockDim.x;
    const double* __restrict__ a = A + s.a_off;
    for(; e + 3 * bd < e1; e += 4 * bd) {
        const double x0 = a[e], x1 = a[e + bd], x2 = a[e + 2 * bd], x3 = a[e + 3 * bd];
        v = fma(x0, x0, v);
        v1 = fma(x1, x1, v1);
        v2 = fma(x2, x2, v2);
        v3 = fma(x3, x3, v3);
    }
    for(; e < e1; e += bd) v = fma(a[e], a[e], v);
    v = (v + v1) + (v2 + v3);
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if(threadIdx.x == 0) {
        double t = 0.0;
        for(int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += sm[w];
        part[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// Cholesky factor of the Gram matrix G (K x K, + shift_rel * max diag) and M = L^-T (upper triangular) so that Y M has orthonormal
// columns.  One CTA per projected sector, the whole matrix REGISTER-resident: 32 x 32 threads, thread (ty, tx) owns the elements
// (i, c) with i = ty (mod 32), c = tx (mod 32) -- NB x NB doubles, NB = ceil(K / 32).  Right-looking factorisation: the owners of
// column j publish it (unscaled) in shared memory, everybody applies the rank-1 update to its own elements: ONE barrier per column
// (the column buffer is double-buffered, pivots through rsqrt: no division).  Then L X = 1 is solved in place the same way (the owners
// publish column j of L and the finished row j of X, rank-1 update of the rows below: no reductions, one barrier per row) and M = X^T is written out.
// diag[sector] = min pivot / max diagonal (~ 1 / cond^2 of the basis); a non-positive pivot sets *bad (the pass is then not an
// orthonormalisation).
template <int NBR, int NBC, int TY>   // rows i = ii * TY + ty (ii < NBR), columns c = cc * 32 + tx (cc < NBC); NBR * TY >= K, NBC * 32 >= K
__global__ void __launch_bounds__(32 * TY, 1)
rsvd_chol_kernel(const double* __restrict__ G, double* __restrict__ M, const RsvdSec* __restrict__ secs, const int* __restrict__ list, double shift_rel,
                 double* __restrict__ diag, int* __restrict__ bad)
{
    __shared__ double col[2][NBC * 32], row[2][NBC * 32], dinv[NBC * 32];
    __shared__ double red[32];
    const int sec = list[blockIdx.x];
    const RsvdSec s = secs[sec];
    const int K = s.K, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const double* g = G + s.g_off;
    double a[NBR][NBC];
    double gmax = 0.0;
#pragma unroll
    for(int ii = 0; ii < NBR; ++ii)
#pragma unroll
        for(int cc = 0; cc < NBC; ++cc) {
            const int i = ii * TY + ty, c = cc * 32 + tx;
            a[ii][cc] = (i < K && c < K) ? g[i + c * K] : (i == c ? 1.0 : 0.0);   // identity padding: decoupled
            if(i == c && i < K) gmax = fmax(gmax, a[ii][cc]);
        }
    // largest diagonal entry (every thread ends up with it)
    for(int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if(tx == 0) red[ty] = gmax;
    __syncthreads();
    gmax = tx < TY ? red[tx] : 0.0;
    for(int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if(!(gmax > 0.0)) gmax = 1.0;
    const double shift = shift_rel * gmax, gmax_inv = 1.0 / gmax;
    // ---- right-looking Cholesky: one barrier per column, no divisions (rsqrt of the pivot, inverse pivot = rsqrt^2).  Few warps with many
    // elements each: the per-column scalar work is paid per warp and the kernel is bound by the instruction issue rate.
    double minpiv = 1.e300;
    bool anybad = false;
    for(int j = 0; j < K; ++j) {
        const int jc = j >> 5, jx = j & 31, b = j & 1;
        if(tx == jx) {
#pragma unroll
            for(int cc = 0; cc < NBC; ++cc)
                if(cc == jc) {
#pragma unroll
                    for(int ii = 0; ii < NBR; ++ii) col[b][ii * TY + ty] = a[ii][cc];
                }
        }
        __syncthreads();
        double d = col[b][j] + shift;
        if(!(d > 0.0) || !isfinite(d)) {
            anybad = true;
            d = gmax * 1.e-30;
        }
        minpiv = fmin(minpiv, d * gmax_inv);
        const double rs = rsqrt(d), inv = rs * rs;
        if(tid == 0) dinv[j] = rs;   // 1 / L_jj for the inversion
        // masked multipliers instead of predicated updates: rows / columns <= j contribute a - 0 * x = a (exactly), so the NBR x NBC update is
        // NBR + NBC selects and NBR * NBC plain FMAs
        double ci[NBR], cj[NBC];
#pragma unroll
        for(int k = 0; k < NBR; ++k) ci[k] = (k * TY + ty > j) ? col[b][k * TY + ty] : 0.0;
#pragma unroll
        for(int k = 0; k < NBC; ++k) cj[k] = (k * 32 + tx > j) ? col[b][k * 32 + tx] * inv : 0.0;
#pragma unroll
        for(int ii = 0; ii < NBR; ++ii)
#pragma unroll
            for(int cc = 0; cc < NBC; ++cc) a[ii][cc] = fma(-ci[ii], cj[cc], a[ii][cc]);
        if(tx == jx) {   // column j of L (zeros above the diagonal)
#pragma unroll
            for(int cc = 0; cc < NBC; ++cc)
                if(cc == jc) {
#pragma unroll
                    for(int ii = 0; ii < NBR; ++ii) {
                        const int i = ii * TY + ty;
                        a[ii][cc] = i > j ? ci[ii] * rs : (i == j ? d * rs : 0.0);
                    }
                }
        }
    }
    __syncthreads();
    // ---- X = L^-1 in place, row by row (solve L X = 1 right-looking): at step j the owners publish column j of L and the finished row j of X
    // (= R_j / L_jj); every element (i > j, c <= j) takes its rank-1 update.  Position (i, c) holds L_ic until step c, the running X_ic afterwards.
#pragma unroll
    for(int ii = 0; ii < NBR; ++ii)
#pragma unroll
        for(int cc = 0; cc < NBC; ++cc)
            if(cc * 32 + tx > ii * TY + ty) a[ii][cc] = 0.0;   // strict upper triangle: symmetric left-overs of the updates
    for(int j = 0; j < K; ++j) {
        const int jc = j >> 5, jx = j & 31, b = j & 1, jr = j % TY, jrr = j / TY;
        const double xjj = dinv[j];
        if(tx == jx) {
#pragma unroll
            for(int cc = 0; cc < NBC; ++cc)
                if(cc == jc) {
#pragma unroll
                    for(int ii = 0; ii < NBR; ++ii) col[b][ii * TY + ty] = a[ii][cc];   // L[:, j] (rows > j are still L)
                }
        }
        if(ty == jr) {
#pragma unroll
            for(int ii = 0; ii < NBR; ++ii)
                if(ii == jrr) {
#pragma unroll
                    for(int cc = 0; cc < NBC; ++cc) {
                        const int c = cc * 32 + tx;
                        const double v = c < j ? a[ii][cc] * xjj : (c == j ? xjj : 0.0);   // finished row j of X
                        a[ii][cc] = v;
                        row[b][c] = v;
                    }
                }
        }
        __syncthreads();
        double ci[NBR], rj[NBC];
#pragma unroll
        for(int k = 0; k < NBR; ++k) ci[k] = (k * TY + ty > j) ? col[b][k * TY + ty] : 0.0;   // L[i][j], rows below j only
#pragma unroll
        for(int k = 0; k < NBC; ++k) rj[k] = (k * 32 + tx < j) ? row[b][k * 32 + tx] : 0.0;   // X[j][c], columns left of j only
#pragma unroll
        for(int ii = 0; ii < NBR; ++ii)
#pragma unroll
            for(int cc = 0; cc < NBC; ++cc) a[ii][cc] = fma(-ci[ii], rj[cc], a[ii][cc]);
        if(tx == jx) {   // column j below the diagonal: L_ij -> X_ij = -L_ij / L_jj
#pragma unroll
            for(int cc = 0; cc < NBC; ++cc)
                if(cc == jc) {
#pragma unroll
                    for(int ii = 0; ii < NBR; ++ii)
                        if(ii * TY + ty > j) a[ii][cc] = -ci[ii] * xjj;
                }
        }
    }
    // M(r, c) = X(c, r)
    double* m_out = M + s.g_off;
#pragma unroll
    for(int ii = 0; ii < NBR; ++ii)
#pragma unroll
        for(int cc = 0; cc < NBC; ++cc) {
            const int i = ii * TY + ty, c = cc * 32 + tx;
            if(i < K && c < K) m_out[c + i * K] = a[ii][cc];
        }
    if(tid == 0) {
        diag[sec] = minpiv;
        if(anybad) atomicOr(bad, 1);
    }
}

// Blocked form of the same factorisation (used for K > 32): the matrix lives in SHARED memory (column-major, odd pitch), 16-column blocks.
// Per block step: (1) warp 0 factorises the 16 x 16 diagonal block in registers (lane = row, pivots / multipliers through shuffles) and inverts its factor;
// (2) panel = A[below, block] * Linv^T (all threads, four outputs each); (3) trailing matrix -= panel * panel^T on 4 x 4 register tiles with rows / columns
// interleaved by a quarter of the trailing size (consecutive lanes -> consecutive addresses).  Three barriers per 16 columns instead of one per column, and the
// per-column latency chain (publish -> barrier -> rsqrt -> update) is paid inside one warp.  X = L^-1 is then built by forward substitution in blocks of 16
// rows, every half-warp owning every 32nd column (columns are independent), straight into the upper triangle of the shared matrix, which thereby becomes M = X^T.
constexpr int CHB = 16;
constexpr int CHOL_THREADS = 512;
__global__ void __launch_bounds__(CHOL_THREADS, 1)
rsvd_chol_blocked_kernel(const double* __restrict__ G, double* __restrict__ M, const RsvdSec* __restrict__ secs, const int* __restrict__ list, double shift_rel,
                         double* __restrict__ diag, int* __restrict__ bad, int bad_tag, unsigned long long* __restrict__ tprof)
{
    extern __shared__ double chol_sh[];
    __shared__ double red[32];
    // optional phase timing (XB_CHOL_PROF=1): lane 0 of warp 0 / warp 1 of CTA 0 accumulate clock64() deltas per phase
    const bool prof = tprof != nullptr && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32);
    long long tp = prof ? clock64() : 0;
    auto lap = [&](int slot) {
        if(prof) {
            const long long now = clock64();
            atomicAdd(tprof + slot, static_cast<unsigned long long>(now - tp));
            tp = now;
        }
    };
    const int sec = list[blockIdx.x];
    const RsvdSec s = secs[sec];
    const int K = s.K, Kp = (K + CHB - 1) / CHB * CHB, LD = Kp + 1, nblk = Kp / CHB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* __restrict__ A = chol_sh;              // Kp x LD: A[c * LD + i]
    double* __restrict__ Pb = A + Kp * LD;         // panel, k-major: Pb[k * Kp + i]
    double* __restrict__ Li = Pb + CHB * Kp;       // inverse of the current diagonal factor: Li[r * 17 + c]
    double* __restrict__ dinv = Li + CHB * (CHB + 1);   // 1 / L_jj
    const double* __restrict__ g = G + s.g_off;
    double gmax = 0.0;
    for(int e0 = tid; e0 < Kp * Kp; e0 += 4 * CHOL_THREADS) {   // four loads in flight per thread (one per iteration: 9 k of the kernel's 127 k cycles at K = 96)
        double v[4];
        int cc[4], ii[4];
#pragma unroll
        for(int u = 0; u < 4; ++u) {
            const int e = e0 + u * CHOL_THREADS;
            cc[u] = e / Kp;
            ii[u] = e - cc[u] * Kp;
            v[u] = (e < Kp * Kp && ii[u] < K && cc[u] < K) ? g[ii[u] + cc[u] * K] : (ii[u] == cc[u] ? 1.0 : 0.0);   // identity padding: decoupled
        }
#pragma unroll
        for(int u = 0; u < 4; ++u) {
            if(e0 + u * CHOL_THREADS < Kp * Kp) {
                A[cc[u] * LD + ii[u]] = v[u];
                if(ii[u] == cc[u] && ii[u] < K) gmax = fmax(gmax, v[u]);
            }
        }
    }
    for(int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if(lane == 0) red[warp] = gmax;
    __syncthreads();
    gmax = lane < CHOL_THREADS / 32 ? red[lane] : 0.0;
    for(int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if(!(gmax > 0.0)) gmax = 1.0;
    const double shift = shift_rel * gmax, gmax_inv = 1.0 / gmax;
    for(int i = tid; i < K; i += CHOL_THREADS) A[i * LD + i] += shift;
    __syncthreads();
    lap(threadIdx.x == 0 ? 0 : 8);   // load

    double minpiv = 1.e300;
    bool anybad = false;
    // diagonal block at c0 (warp 0): lane r (both half-warps redundantly) owns row r.  `fused`: the block first takes the rank-16 update of the panel that
    // was just computed (its rows are the first 16 of the panel buffer) -- look-ahead: the next diagonal block is factorised while the other warps apply
    // the trailing update.
    auto factor_diag = [&](int c0, bool fused) {
        const int r = lane & 15;
        double d[CHB];
#pragma unroll
        for(int c = 0; c < CHB; ++c) d[c] = c <= r ? A[(c0 + c) * LD + c0 + r] : 0.0;
        if(fused) {
#pragma unroll 4
            for(int k = 0; k < CHB; ++k) {
                const double pr = Pb[k * Kp + r];
#pragma unroll
                for(int c = 0; c < CHB; ++c) d[c] = fma(-pr, Pb[k * Kp + c], d[c]);   // (entries above the diagonal are never used)
            }
        }
#pragma unroll
        for(int j = 0; j < CHB; ++j) {
            double djj = __shfl_sync(0xffffffffu, d[j], j);
            if(!(djj > 0.0) || !isfinite(djj)) {
                anybad = true;
                djj = gmax * 1.e-30;
            }
            if(c0 + j < K) minpiv = fmin(minpiv, djj * gmax_inv);
            const double rs = rsqrt(djj);
            if(lane == j) dinv[c0 + j] = rs;
            const double l = r > j ? d[j] * rs : (r == j ? djj * rs : 0.0);
            d[j] = l;
#pragma unroll
            for(int c = j + 1; c < CHB; ++c) {
                const double lc = __shfl_sync(0xffffffffu, l, c);
                d[c] = fma(-l, lc, d[c]);
            }
        }
        __syncwarp();
        // ---- X = L^-1 of the block by columns: lane c solves L x = e_c (L_ik from its row owner)
        double x[CHB];
#pragma unroll
        for(int i = 0; i < CHB; ++i) {
            double acc = (i == r) ? 1.0 : 0.0;
#pragma unroll
            for(int k = 0; k < i; ++k) acc = fma(-__shfl_sync(0xffffffffu, d[k], i), x[k], acc);
            x[i] = acc * dinv[c0 + i];
        }
        if(lane < CHB) {
#pragma unroll
            for(int c = 0; c < CHB; ++c)
                if(c <= r) A[(c0 + c) * LD + c0 + r] = d[c];               // L block
#pragma unroll
            for(int i = 0; i < CHB; ++i) {
                Li[i * (CHB + 1) + r] = x[i];                               // Linv[i][c = r] (zero above the diagonal)
                if(i > r) A[(c0 + i) * LD + c0 + r] = x[i];                 // M part: position (row c, col i) = X[i][c]
            }
        }
    };
    if(warp == 0) factor_diag(0, false);
    lap(threadIdx.x == 0 ? 1 : 9);   // first diagonal block (warp 1: waits)
    __syncthreads();
    lap(threadIdx.x == 0 ? 2 : 10);
    for(int kb = 0; kb + 1 < nblk; ++kb) {
        const int c0 = kb * CHB, r0 = c0 + CHB, nrows = Kp - r0;
        // ---- panel: P[i][j] = sum_{c <= j} A[r0 + i][c0 + c] * Linv[j][c]
        for(int idx = tid; idx < nrows * 4; idx += CHOL_THREADS) {
            const int jq = idx / nrows, i = idx - jq * nrows;
            double a[CHB];
#pragma unroll
            for(int c = 0; c < CHB; ++c) a[c] = A[(c0 + c) * LD + r0 + i];
#pragma unroll
            for(int jj = 0; jj < 4; ++jj) {
                const int j = jq * 4 + jj;
                double p = 0.0;
#pragma unroll
                for(int c = 0; c < CHB; ++c) p = fma(a[c], Li[j * (CHB + 1) + c], p);   // Linv is lower triangular: entries with c > j are zero
                Pb[j * Kp + i] = p;
            }
        }
        lap(threadIdx.x == 0 ? 3 : 11);   // panel
        __syncthreads();
        lap(threadIdx.x == 0 ? 2 : 10);   // barrier wait
        if(warp == 0) {
            factor_diag(r0, true);   // overwrites Li: the panel above was its last reader
            lap(4);   // diagonal block
        } else {
            // ---- trailing update (full square: the matrix stays symmetric; the next diagonal block belongs to warp 0) + the panel goes back into the matrix
            const int q = nrows >> 2, t_id = tid - 32, t_n = CHOL_THREADS - 32;
            for(int t = t_id; t < q * q; t += t_n) {
                const int tc = t / q, ti = t - tc * q;
                double acc[4][4];
#pragma unroll
                for(int a = 0; a < 4; ++a)
#pragma unroll
                    for(int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
                for(int k = 0; k < CHB; ++k) {
                    double pr[4], pc[4];
#pragma unroll
                    for(int a = 0; a < 4; ++a) pr[a] = Pb[k * Kp + ti + a * q];
#pragma unroll
                    for(int b = 0; b < 4; ++b) pc[b] = Pb[k * Kp + tc + b * q];
#pragma unroll
                    for(int a = 0; a < 4; ++a)
#pragma unroll
                        for(int b = 0; b < 4; ++b) acc[a][b] = fma(pr[a], pc[b], acc[a][b]);
                }
#pragma unroll
                for(int b = 0; b < 4; ++b)
#pragma unroll
                    for(int a = 0; a < 4; ++a) {
                        const int gr = ti + a * q, gc = tc + b * q;
                        if(gr >= CHB || gc >= CHB) A[(r0 + gc) * LD + r0 + gr] -= acc[a][b];
                    }
            }
            for(int idx = t_id; idx < nrows * CHB; idx += t_n) {
                const int j = idx / nrows, i = idx - j * nrows;
                A[(c0 + j) * LD + r0 + i] = Pb[j * Kp + i];
            }
            lap(12);   // trailing update
        }
        __syncthreads();
        lap(threadIdx.x == 0 ? 2 : 10);
    }
    // ---- X = L^-1 below the diagonal blocks, one block ROW per step, all threads:  X[ib][c] = -Linv_ib * S,  S[r][c] = sum_{k = c}^{16 ib - 1} L[16 ib + r][k] X[k][c].
    // X[k][c] sits at A[k * LD + c] (upper triangle), Linv_ib[r][k] (k < r) at A[(16 ib + r) * LD + 16 ib + k]; the diagonal of A takes X's diagonal (1 / L_jj) first,
    // so that neither sum has a special case.  Thread = (row r of the block, columns c = group + 32 m): plain dot products on four chains, S goes through the panel
    // buffer, two barriers per block row.  (The column-by-column form with one half-warp per column walked 16-step dependent chains per 16 x 16 block and took
    // 50 k of the kernel's 127 k cycles at K = 96.)
    for(int i = tid; i < Kp; i += CHOL_THREADS) A[i * LD + i] = dinv[i];
    __syncthreads();
    {
        constexpr int MC = 5;   // Kp <= 160 columns / 32 groups
        const int r = tid & 15, grp = tid >> 4;
        for(int ib = 1; ib < nblk; ++ib) {
            const int row = ib * CHB + r, ncol = ib * CHB;
            const double* __restrict__ Lr = A + row;
#pragma unroll
            for(int m = 0; m < MC; ++m) {
                const int c = grp + 32 * m;
                if(c < ncol) {
                    const double* __restrict__ Xc = A + c;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    int k = c;
                    for(; k + 3 < ncol; k += 4) {
                        s0 = fma(Lr[k * LD], Xc[k * LD], s0);
                        s1 = fma(Lr[(k + 1) * LD], Xc[(k + 1) * LD], s1);
                        s2 = fma(Lr[(k + 2) * LD], Xc[(k + 2) * LD], s2);
                        s3 = fma(Lr[(k + 3) * LD], Xc[(k + 3) * LD], s3);
                    }
                    for(; k < ncol; ++k) s0 = fma(Lr[k * LD], Xc[k * LD], s0);
                    Pb[r * Kp + c] = (s0 + s1) + (s2 + s3);
                }
            }
            __syncthreads();
            {
                double acc[MC];
#pragma unroll
                for(int m = 0; m < MC; ++m) acc[m] = 0.0;
                const double* __restrict__ Li_r = A + row * LD + ib * CHB;   // Linv_ib[r][k], k <= r
#pragma unroll 4
                for(int k = 0; k < CHB; ++k) {
                    const double li = k <= r ? -Li_r[k] : 0.0;
#pragma unroll
                    for(int m = 0; m < MC; ++m) {
                        const int c = grp + 32 * m;
                        if(c < ncol) acc[m] = fma(li, Pb[k * Kp + c], acc[m]);
                    }
                }
#pragma unroll
                for(int m = 0; m < MC; ++m) {
                    const int c = grp + 32 * m;
                    if(c < ncol) A[row * LD + c] = acc[m];
                }
            }
            __syncthreads();
        }
    }
    lap(threadIdx.x == 0 ? 5 : 13);   // substitution
    __syncthreads();
    lap(threadIdx.x == 0 ? 2 : 10);
    double* __restrict__ m_out = M + s.g_off;
    for(int e = tid; e < K * K; e += CHOL_THREADS) {
        const int c = e / K, rr = e - c * K;   // M(rr, c) = X(c, rr)
        m_out[e] = rr < c ? A[c * LD + rr] : (rr == c ? dinv[c] : 0.0);
    }
    lap(threadIdx.x == 0 ? 6 : 14);   // store
    if(prof && threadIdx.x == 0) atomicAdd(tprof + 7, 1ull);
    if(tid == 0) {
        diag[sec] = minpiv;
if(anybad) atomicOr(bad, bad_tag);
        if(bad_tag == 16 && minpiv < 0.25) atomicOr(bad, 8);   // single-pass polish on a basis too far from orthonormal
    }
}

// residual of the Ritz triplets: res[id] = || Zr_i - sigma_i^2 v_i || / sigma_i  with Zr = B^T W (W_i = sigma_i u_i); one warp per column
__global__ void rsvd_residual_kernel(const double* __restrict__ Zr, const double* __restrict__ Vf, const double* __restrict__ sigma, const RsvdSec* __restrict__ secs,
                                     const int* __restrict__ list, double* __restrict__ res)
{
    const RsvdSec s = secs[list[blockIdx.y]];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for(int c = blockIdx.x * warps + warp; c < s.K; c += gridDim.x * warps) {
        const double sg = sigma[s.s_off + c];
        const double* z = Zr + s.x_off + static_cast<long long>(c) * s.nb;
        const double* v = Vf + s.x_off + static_cast<long long>(c) * s.nb;
        double acc = 0.0;
        for(int r = lane; r < s.nb; r += 32) {
            const double d = fma(-sg * sg, v[r], z[r]);
            acc = fma(d, d, acc);
        }
        for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if(lane == 0) res[s.s_off + c] = sg > 0.0 ? sqrt(acc) / sg : 0.0;
    }
}

GemmItem gemm_item(long long a_off, long long b_off, long long c_off, int m, int n, int k, int lda, int ldb, int ldc, int ta, int tb)
{
    GemmItem g;
    std::memset(&g, 0, sizeof(g));
    g.a_off = a_off;
    g.b_off = b_off;
    g.c_off = c_off;
    g.m = m;
    g.n = n;
    g.k = k;
    g.lda = lda;
    g.ldb = ldb;
    g.ldc = ldc;
    g.ta = ta;
    g.tb = tb;
    g.alpha = 1.0;
    return g;
}

struct Selection
{
    SelHeader hdr;
    std::vector<int> kept;
    int* d_perm = nullptr;   // column order per sector (device), valid until freed by the caller
};

// runs the device-wide sort / truncation; synchronises once to read the small result record
Selection svd_select(Context& ctx, const double* d_sigma, int n_total, const std::vector<SelSector>& secs, const double* d_res, const double* d_normpart, double tol_rel,
                     int64_t max_keep, double eps_svd, bool preserve, bool ritz)
{
    cudaStream_t s = ctx.stream;
    const int nsec = static_cast<int>(secs.size());
    Selection R;
    int N = 2;
    while(N < n_total) N <<= 1;
    SelSector* d_secs = nullptr;
    char* d_out = nullptr;   // header | kept[nsec]
    const size_t out_bytes = sizeof(SelHeader) + sizeof(int) * nsec;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_secs), sizeof(SelSector) * nsec, s), "svd select meta");
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_out), out_bytes, s), "svd select out");
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&R.d_perm), sizeof(int) * std::max(n_total, 1), s), "svd select perm");
    ctx.upload_async(d_secs, secs.data(), sizeof(SelSector) * nsec);
    SelParams P;
    P.max_keep = max_keep;
    P.eps_svd = eps_svd;
    P.tol_rel = tol_rel;
    P.preserve = preserve ? 1 : 0;
    P.ritz = ritz ? 1 : 0;
    P.nsec = nsec;
    P.n_total = n_total;
    P.guard = 8;
    P.pad = 0;
    const size_t smem = sizeof(double) * (N + 40) + sizeof(int) * (N + 3 * static_cast<size_t>(nsec) + 8);
    static unsigned long long configured = 0;
    if(first_use_on_current_device(configured)) {
        check_cuda(cudaFuncSetAttribute(svd_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024), "svd select smem attr");
    }
    svd_select_kernel<<<1, SEL_THREADS, smem, s>>>(d_sigma, d_secs, d_res, d_normpart, P, N, R.d_perm, reinterpret_cast<SelHeader*>(d_out),
                                                  reinterpret_cast<int*>(d_out + sizeof(SelHeader)));
    check_cuda(cudaGetLastError(), "svd_select_kernel launch");
    ++ctx.launches;
    char* h = static_cast<char*>(ctx.pinned(out_bytes));
    check_cuda(cudaMemcpyAsync(h, d_out, out_bytes, cudaMemcpyDeviceToHost, s), "svd select readback");
    check_cuda(cudaStreamSynchronize(s), "svd select sync");
    std::memcpy(&R.hdr, h, sizeof(SelHeader));
    R.kept.resize(nsec);
    std::memcpy(R.kept.data(), h + sizeof(SelHeader), sizeof(int) * nsec);
    check_cuda(cudaFreeAsync(d_secs, s), "svd select free");
    check_cuda(cudaFreeAsync(d_out, s), "svd select free");
    return R;
}

// per input sector: where its W / V columns live
struct FactorSrc
{
    long long w_off, v_off, s_off;
    int wl, vl, ldw, ldv, w_is_u;
};

// builds the truncated basis and the three output tensors from a selection (Tensor.cpp:592-633) and launches the gather
SvdResult svd_finish(Context& ctx, const Layout& L, const Selection& sel, const std::vector<FactorSrc>& src, const double* d_W, const double* d_V,
                     const double* d_sigma)
{
    SvdResult res;
    res.trunc_weight = sel.hdr.trunc_weight;
    res.entropy = sel.hdr.entropy;
    std::map<Q, int> ordered;   // std::map => ascending qarray order
    for(size_t i = 0; i < L.sector.size(); ++i)
        if(sel.kept[i] > 0) ordered[L.sector[i]] = static_cast<int>(i);
    auto trunc = std::make_shared<Qbasis>(ctx.sym, 1);
    std::vector<Q> qs;
    for(const auto& [q, i] : ordered) {
        trunc->push_back(q, static_cast<int64_t>(sel.kept[i]));
        qs.push_back(q);
    }
    BasisPtr tb = trunc;
    LayoutPtr LU = ctx.layout_custom(L.dom, {tb}, qs), LS = ctx.layout_custom({tb}, {tb}, qs), LV = ctx.layout_custom({tb}, L.cod, qs);
    res.U = new Tensor(&ctx, LU, false);
    res.S = new Tensor(&ctx, LS, true);
    res.Vd = new Tensor(&ctx, LV, false);
    std::vector<GatherSec> gs;
    int max_ret = 1;
    for(size_t k = 0; k < qs.size(); ++k) {
        const int i = ordered[qs[k]];
        const FactorSrc& f = src[i];
        GatherSec g;
        g.w_off = f.w_off;
        g.v_off = f.v_off;
        g.u_dst = LU->off[k];
        g.vd_dst = LV->off[k];
        g.s_dst = LS->off[k];
        g.sel_off = f.s_off;
        g.wl = f.wl;
        g.vl = f.vl;
        g.ldw = f.ldw;
        g.ldv = f.ldv;
        g.nret = sel.kept[i];
        g.w_is_u = f.w_is_u;
        max_ret = std::max(max_ret, g.nret);
        gs.push_back(g);
    }
    if(gs.empty()) return res;
    GatherSec* d_gs = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_gs), sizeof(GatherSec) * gs.size(), ctx.stream), "svd gather meta");
    ctx.upload_async(d_gs, gs.data(), sizeof(GatherSec) * gs.size());
    dim3 grid(static_cast<unsigned>(std::min(max_ret, 512)), static_cast<unsigned>(gs.size()));
    svd_gather_sel_kernel<<<grid, 128, 0, ctx.stream>>>(d_W, d_V, d_sigma, sel.d_perm, res.U->d, res.S->d, res.Vd->d, d_gs);
    check_cuda(cudaGetLastError(), "svd_gather_sel_kernel launch");
    ++ctx.launches;
    check_cuda(cudaFreeAsync(d_gs, ctx.stream), "svd gather free");
    return res;
}

}   // namespace

bool svd_device_select_supported(long long n_total, int nsec) { return n_total <= SEL_MAX && nsec <= 4096; }

// full path: Jacobi on every block, device-wide sort / truncation, gather
SvdResult svd_full_device_select(Context& ctx, const Tensor& t, JacobiResult& J, int64_t max_keep, double eps_svd, bool preserve_multiplets)
{
    const Layout& L = *t.layout;
    std::vector<SelSector> secs(L.sector.size());
    std::vector<FactorSrc> src(L.sector.size());
    for(size_t i = 0; i < L.sector.size(); ++i) {
        const SvdBlock& b = J.blocks[i];
        secs[i] = {b.s_off, b.cols, 1, static_cast<double>(ctx.sym->degeneracy(L.sector[i]))};
        src[i] = {b.w_off, b.v_off, b.s_off, b.rows, b.cols, b.rows, b.cols, J.transposed[i] ? 0 : 1};
    }
    Selection sel = svd_select(ctx, J.d_sigma, static_cast<int>(J.n_sigma), secs, nullptr, nullptr, 0.0, max_keep, eps_svd, preserve_multiplets, false);
    SvdResult res = svd_finish(ctx, L, sel, src, J.d_W, J.d_V, J.d_sigma);
    check_cuda(cudaFreeAsync(sel.d_perm, ctx.stream), "svd select free");
    return res;
}

// ------------------------------------------------------------------------------------------------ truncated tSVD
// Everything that depends only on the layout (block shapes) and K lives in a cached plan: one device arena with fixed offsets (so that the
// pointers baked into the captured CUDA graphs stay valid from call to call), the prepared tile lists of the grouped products, and the
// instantiated graphs of the launch sequences (start, n iterations + polish + projection).
struct RsvdPlan
{
    Context* ctx = nullptr;
    int nsec = 0, n_proj = 0, K0 = 0, mb_max = 0, uses = 0, iters_hint = 0;
    long long xo = 0, yo = 0, go = 0, so = 0, ao = 0;
    std::vector<RsvdSec> secs;
    std::vector<long long> w_off, v_off;
    std::vector<int> proj_list, direct_list;
    char* arena = nullptr;
    double *dA = nullptr, *dX = nullptr, *dX2 = nullptr, *dZ = nullptr, *dY = nullptr, *dY2 = nullptr, *dG = nullptr, *dM = nullptr, *dDiag = nullptr, *dRes = nullptr,
           *dNorm = nullptr;
    RsvdSec* d_secs = nullptr;
    int *d_proj = nullptr, *d_direct = nullptr, *d_bad = nullptr;
    PreparedGemm g_bx, g_bx_all, g_gy, g_ym, g_bty, g_bty_raw, g_gx2, g_x2m, g_gx, g_xm, g_btw, g_xv;
    size_t chol_smem = 0;
    double tol_rel = 0.0, shift_it = 0.0;
    double* dVprev = nullptr;   // Ritz vectors of the previous successful call on this plan (warm start)
    int* d_perm_prev = nullptr;
    bool have_prev = false;
    int hold = 0;   // calls during which iters_hint is not lowered (after a projection that was not converged)
    cudaGraphExec_t graph_warm = nullptr;
    cudaGraphExec_t graph_start = nullptr;
    std::map<int, cudaGraphExec_t> graph_iter;   // key: number of iterations
    std::map<int, int> graph_launches;           // kernel launches inside each graph (key -1: start)
    JacobiInput jin;
    std::string jin_sig_base;
    bool jacobi_precondition = false;   // see the Rayleigh-Ritz step
    int jacobi_sweeps = 0;               // sweeps of the previous projected Jacobi solve from this call site (batch size of the next one)
    std::vector<SelSector> sel_secs;
    ~RsvdPlan()
    {
        if(graph_start) cudaGraphExecDestroy(graph_start);
        if(graph_warm) cudaGraphExecDestroy(graph_warm);
        for(auto& kv : graph_iter)
            if(kv.second) cudaGraphExecDestroy(kv.second);
        for(PreparedGemm* g : {&g_bx, &g_bx_all, &g_gy, &g_ym, &g_bty, &g_bty_raw, &g_gx2, &g_x2m, &g_gx, &g_xm, &g_btw, &g_xv}) {
            for(int w = 0; w < 2; ++w)
                if(g->d_items[w]) cudaFree(g->d_items[w]);
            if(g->d_jobs) cudaFree(g->d_jobs);
            if(g->d_scratch) cudaFree(g->d_scratch);
        }
        if(arena) cudaFree(arena);
    }
};

namespace {

void rsvd_chol(RsvdPlan& P, double shift_rel, int bad_tag = 1)
{
    cudaStream_t s = P.ctx->stream;
    static const bool old_form = std::getenv("XB_CHOL_OLD") != nullptr;
    if(P.K0 > 32 && !old_form) {
        const int Kp = (P.K0 + CHB - 1) / CHB * CHB;
        const size_t smem = sizeof(double) * (static_cast<size_t>(Kp) * (Kp + 1) + CHB * Kp + CHB * (CHB + 1) + Kp);
        static unsigned long long configured = 0;
        if(first_use_on_current_device(configured)) {
            const int most = 227 * 1024 - 512;   // the opt-in maximum covers static + dynamic shared memory
            check_cuda(cudaFuncSetAttribute(rsvd_chol_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, most), "chol smem attr");
        }
        static unsigned long long* d_tprof = nullptr;
        static const bool chol_prof = std::getenv("XB_CHOL_PROF") != nullptr;
        if(chol_prof && !d_tprof) {
            cudaMalloc(reinterpret_cast<void**>(&d_tprof), 16 * sizeof(unsigned long long));
            cudaMemset(d_tprof, 0, 16 * sizeof(unsigned long long));
        }
        rsvd_chol_blocked_kernel<<<P.n_proj, CHOL_THREADS, smem, s>>>(P.dG, P.dM, P.d_secs, P.d_proj, shift_rel, P.dDiag, P.d_bad, bad_tag, d_tprof);
        if(chol_prof) {
            static int calls = 0;
            if(++calls % 100 == 0) {
                cudaStreamSynchronize(s);
                unsigned long long h[16];
                cudaMemcpy(h, d_tprof, sizeof(h), cudaMemcpyDeviceToHost);
                const double n = static_cast<double>(h[7] ? h[7] : 1);
                std::fprintf(stderr, "[chol prof] K0 %d, %.0f launches | warp 0 cycles/launch: load %.0f diag0 %.0f barrier %.0f panel %.0f diag %.0f subst %.0f store %.0f | warp 1: load %.0f "
                                     "wait-diag0 %.0f barrier %.0f panel %.0f trailing %.0f subst %.0f store %.0f\n",
                             P.K0, n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n, h[12] / n, h[13] / n, h[14] / n);
            }
        }
        check_cuda(cudaGetLastError(), "rsvd_chol_blocked_kernel launch");
        ++P.ctx->launches;
        return;
    }
#define XB_CHOL(NBR, NBC, TY) rsvd_chol_kernel<NBR, NBC, TY><<<P.n_proj, 32 * TY, 0, s>>>(P.dG, P.dM, P.d_secs, P.d_proj, shift_rel, P.dDiag, P.d_bad)
    switch((P.K0 + 31) / 32) {
    // few warps with many elements each: the kernel is bound by the instruction issue rate (per-step overhead is paid per warp)
    case 1: XB_CHOL(4, 1, 8); break;
    case 2: XB_CHOL(8, 2, 8); break;
    case 3: {
        static const int ty = std::getenv("XB_CHOL_TY") ? std::atoi(std::getenv("XB_CHOL_TY")) : 8;   // experiment knob
        if(ty == 4)
            XB_CHOL(24, 3, 4);
        else if(ty == 16)
            XB_CHOL(6, 3, 16);
        else
            XB_CHOL(12, 3, 8);
        break;
    }
    case 4: XB_CHOL(16, 4, 8); break;
    default: XB_CHOL(10, 5, 16); break;  // K <= 160: 50 doubles per thread at 512 threads
    }
#undef XB_CHOL
    check_cuda(cudaGetLastError(), "rsvd_chol_kernel launch");
    ++P.ctx->launches;
}
void rsvd_polish(RsvdPlan& P, bool single_pass = false)
{   // X <- orthonormal to working precision: two unshifted Cholesky-QR passes, or ONE when X comes out of an iteration (its orthonormality error is then
    // 1e-6 .. 3e-2 on CTM matrices, one pass leaves 1e-15 -- numpy emulation; the factorisation flags a smallest pivot below 1/4 of the largest diagonal
    // entry, i.e. a basis too far from orthonormal for one pass, as a failed call: bad bit 8)
    Context& c = *P.ctx;
    c.launch_gemm(P.g_gx);
    rsvd_chol(P, 0.0, single_pass ? 16 : 2);
    c.launch_gemm(P.g_xm);
    if(single_pass) {
        check_cuda(cudaMemcpyAsync(P.dX, P.dX2, sizeof(double) * P.xo, cudaMemcpyDeviceToDevice, c.stream), "rsvd basis copy");
        return;
    }
    c.launch_gemm(P.g_gx2);
    rsvd_chol(P, 0.0, 4);
    c.launch_gemm(P.g_x2m);
}
constexpr int WARM_RANDOM_COLS = 16;
void rsvd_start(RsvdPlan& P, bool warm)
{   // X = random (cold) or the previous Ritz vectors + a few random columns (warm), identity on the exactly treated sectors; ||A_q||_F^2; orthonormalise
    Context& c = *P.ctx;
    cudaStream_t s = c.stream;
    check_cuda(cudaMemsetAsync(P.d_bad, 0, sizeof(int), s), "rsvd flag reset");
    check_cuda(cudaMemsetAsync(P.dRes, 0, sizeof(double) * std::max<long long>(P.so, 1), s), "rsvd residual reset");
    if(warm) {
        rsvd_warm_start_kernel<<<dim3(static_cast<unsigned>(P.K0), static_cast<unsigned>(P.nsec)), 128, 0, s>>>(P.dX, P.dVprev, P.d_perm_prev, P.d_secs, WARM_RANDOM_COLS,
                                                                                                                0x5eedc0ffeeull + static_cast<uint64_t>(P.xo));
        check_cuda(cudaGetLastError(), "rsvd_warm_start_kernel launch");
        ++c.launches;
    } else {
        launch_fill_random(P.dX, 0x5eedc0ffeeull + static_cast<uint64_t>(P.xo), P.xo, s);
        ++c.launches;
        if(!P.direct_list.empty()) {
            rsvd_identity_kernel<<<dim3(8, static_cast<unsigned>(P.direct_list.size())), 256, 0, s>>>(P.dX, P.d_secs, P.d_direct, static_cast<int>(P.direct_list.size()));
            check_cuda(cudaGetLastError(), "rsvd_identity_kernel launch");
            ++c.launches;
        }
    }
    rsvd_norm2_kernel<<<dim3(NORM_PARTS, P.nsec), 256, 0, s>>>(P.dA, P.d_secs, P.dNorm);
    check_cuda(cudaGetLastError(), "rsvd_norm2_kernel launch");
    ++c.launches;
    // a warm start (orthonormal Ritz vectors + random columns at the end) goes straight into X <- orth(B^T B X), which needs a well-conditioned, ordered basis,
    // not an orthonormal one; the two Cholesky-QR passes (8 launches) are only spent on a cold start
    static const bool two_orth = std::getenv("XB_RSVD_TWO_ORTH") != nullptr, always_polish = std::getenv("XB_RSVD_START_POLISH") != nullptr;
    if(!warm || two_orth || always_polish || P.K0 <= 32) rsvd_polish(P);
}
void rsvd_iterate(RsvdPlan& P, int iters, bool ordered_basis)
{   // `iters` subspace iterations X <- orth(B^T orth(B X)), orthonormalise, project: Y = B X
    Context& c = *P.ctx;
    // One orthonormalisation per iteration: X <- orth(B^T (B X)).  The basis is ordered (warm start = previous Ritz vectors by decreasing value, Cholesky-QR
    // = Gram-Schmidt in column order), so the iterates stay close to the singular vectors and the Gram matrix of B^T B X is GRADED (diag ~ sigma^4) but nearly
    // diagonal after scaling -- the regime in which a Cholesky factorisation is accurate column by column.  The shift must stay at round-off level here
    // (1e-15 of the largest diagonal entry): the pivots of the weakest kept directions are (sigma_keep / sigma_1)^4 ~ 1e-12 of the largest, and a shift above
    // them spoils the elimination against the leading columns (emulated in numpy on the oracle's CTM matrices: breakdown with the 1e-13 * m shift of the
    // two-stage iteration, identical residuals to the two-stage iteration with 1e-15 or 0).  The two unshifted passes of rsvd_polish restore orthonormality
    // before the projection.  Same iteration counts and residuals as with the intermediate orth(B X) (XB_RSVD_TWO_ORTH=1 restores it), 4 instead of 6
    // products and 1 instead of 2 factorisations per iteration.  A breakdown (exactly rank-deficient input) sets the bad flag: the caller falls back.
    // A cold start (random X) is NOT ordered: B^T B X is then dominated by the leading direction in every column and its Gram matrix is numerically
    // rank-deficient -- the first batch of a cold call keeps the intermediate orthonormalisation.
    // iters < 0: only the tail (orthonormalisation + projection) after -iters - 1 ... see the caller: the iterations and the tail are separate CUDA graphs, so
    // that a changing iteration count does not need a new graph (an instantiation costs 40-60 ms)
    static const bool two_orth = std::getenv("XB_RSVD_TWO_ORTH") != nullptr;
    bool single_orth_path = !two_orth && ordered_basis && P.K0 > 32;
    const bool tail_only = iters < 0, no_tail = iters >= 1000;
    if(no_tail) iters -= 1000;
    if(tail_only) iters = 0;
    for(int it = 0; it < iters; ++it) {
        if(!two_orth && ordered_basis && P.K0 > 32) {   // (K <= 32 uses the register-resident factorisation, which does not rescale)
            c.launch_gemm(P.g_bx);
            c.launch_gemm(P.g_bty_raw);
            c.launch_gemm(P.g_gx2);
            rsvd_chol(P, 1.e-15);
            c.launch_gemm(P.g_x2m);
            single_orth_path = true;
            continue;
        }
        c.launch_gemm(P.g_bx);
        c.launch_gemm(P.g_gy);
        rsvd_chol(P, P.shift_it);
        c.launch_gemm(P.g_ym);
        c.launch_gemm(P.g_bty);
        c.launch_gemm(P.g_gx2);
        rsvd_chol(P, P.shift_it);
        c.launch_gemm(P.g_x2m);
    }
    if(no_tail) return;
    static const bool two_pass_polish = std::getenv("XB_RSVD_POLISH2") != nullptr;
    rsvd_polish(P, single_orth_path && !two_pass_polish);
    c.launch_gemm(P.g_bx_all);
}

// runs `body` directly, or captures it once into a CUDA graph and replays it
template <typename F> void run_or_replay(RsvdPlan& P, cudaGraphExec_t* slot, int key, bool use_graph, F body)
{
    Context& c = *P.ctx;
    cudaStream_t s = c.stream;
    if(!use_graph) {
        body();
        return;
    }
    if(*slot == nullptr) {
        const bool prof_saved = c.profiling;
        c.profiling = false;   // no event records inside a capture
        const int64_t l0 = c.launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if(ok) {
            try {
                body();
            } catch(...) {
                ok = false;
            }
            ok = (cudaStreamEndCapture(s, &graph) == cudaSuccess) && ok && graph != nullptr;
        }
        if(ok) ok = cudaGraphInstantiate(slot, graph, 0) == cudaSuccess;
        if(graph) cudaGraphDestroy(graph);
        c.profiling = prof_saved;
        P.graph_launches[key] = static_cast<int>(c.launches - l0);
        c.launches = l0;
        if(!ok) {
            cudaGetLastError();
            *slot = nullptr;
            body();
            return;
        }
    }
    check_cuda(cudaGraphLaunch(*slot, s), "rsvd graph launch");
    c.launches += P.graph_launches[key];
}

std::shared_ptr<RsvdPlan> build_rsvd_plan(Context& ctx, const Layout& L, int K0)
{
    auto Pp = std::make_shared<RsvdPlan>();
    RsvdPlan& P = *Pp;
    P.ctx = &ctx;
    P.K0 = K0;
    const int nsec = static_cast<int>(L.sector.size());
    P.nsec = nsec;
    P.secs.resize(nsec);
    P.w_off.resize(nsec);
    P.v_off.resize(nsec);
    long long wo = 0, vo = 0;
    for(int i = 0; i < nsec; ++i) {
        RsvdSec& s = P.secs[i];
        s.a_off = L.off[i];
        s.m = static_cast<int>(L.rows[i]);
        s.n = static_cast<int>(L.cols[i]);
        s.tr = s.m < s.n;
        s.mb = std::max(s.m, s.n);
        s.nb = std::min(s.m, s.n);
        s.direct = s.nb <= K0;
        s.K = s.direct ? s.nb : K0;
        s.pad = 0;
        s.x_off = P.xo;
        s.y_off = P.yo;
        s.g_off = P.go;
        s.s_off = P.so;
        P.w_off[i] = wo;   // where jacobi_all puts the work matrix / accumulator of the projected block
        P.v_off[i] = vo;
        wo += static_cast<long long>(s.mb) * s.K;
        vo += static_cast<long long>(s.K) * s.K;
        P.xo += static_cast<long long>(s.nb) * s.K;
        P.yo += static_cast<long long>(s.mb) * s.K;
        P.go += static_cast<long long>(s.K) * s.K;
        P.so += s.K;
        P.mb_max = std::max(P.mb_max, s.mb);
        (s.direct ? P.direct_list : P.proj_list).push_back(i);
        // keep every buffer segment 16-byte aligned for the vectorised GEMM loads
        P.xo += P.xo & 1;
        P.yo += P.yo & 1;
        P.go += P.go & 1;
    }
    P.n_proj = static_cast<int>(P.proj_list.size());
    P.ao = L.size + (L.size & 1);
    // ---- one arena
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t bX = al(sizeof(double) * std::max<long long>(P.xo, 1)), bY = al(sizeof(double) * std::max<long long>(P.yo, 1)),
                 bG = al(sizeof(double) * std::max<long long>(P.go, 1)), bA = al(sizeof(double) * std::max<long long>(P.ao, 1));
    const size_t bDiag = al(sizeof(double) * nsec), bRes = al(sizeof(double) * std::max<long long>(P.so, 1)), bNorm = al(sizeof(double) * nsec * NORM_PARTS);
    const size_t bSecs = al(sizeof(RsvdSec) * nsec), bList = al(sizeof(int) * nsec);
    const size_t bPerm = al(sizeof(int) * std::max<long long>(P.so, 1));
    const size_t total = bA + 4 * bX + 2 * bY + 2 * bG + bDiag + bRes + bNorm + bSecs + 2 * bList + bPerm + 256;
    check_cuda(cudaMalloc(reinterpret_cast<void**>(&P.arena), total), "rsvd arena");
    char* p = P.arena;
    auto take = [&](size_t b) { char* r = p; p += b; return r; };
    P.dA = reinterpret_cast<double*>(take(bA));
    P.dX = reinterpret_cast<double*>(take(bX));
    P.dX2 = reinterpret_cast<double*>(take(bX));
    P.dZ = reinterpret_cast<double*>(take(bX));
    P.dVprev = reinterpret_cast<double*>(take(bX));
    P.d_perm_prev = reinterpret_cast<int*>(take(bPerm));
    P.dY = reinterpret_cast<double*>(take(bY));
    P.dY2 = reinterpret_cast<double*>(take(bY));
    P.dG = reinterpret_cast<double*>(take(bG));
    P.dM = reinterpret_cast<double*>(take(bG));
    P.dDiag = reinterpret_cast<double*>(take(bDiag));
    P.dRes = reinterpret_cast<double*>(take(bRes));
    P.dNorm = reinterpret_cast<double*>(take(bNorm));
    P.d_secs = reinterpret_cast<RsvdSec*>(take(bSecs));
    P.d_proj = reinterpret_cast<int*>(take(bList));
    P.d_direct = reinterpret_cast<int*>(take(bList));
    P.d_bad = reinterpret_cast<int*>(take(256));
    ctx.upload_async(P.d_secs, P.secs.data(), sizeof(RsvdSec) * nsec);
    if(!P.proj_list.empty()) ctx.upload_async(P.d_proj, P.proj_list.data(), sizeof(int) * P.proj_list.size());
    if(!P.direct_list.empty()) ctx.upload_async(P.d_direct, P.direct_list.data(), sizeof(int) * P.direct_list.size());

    // ---- prepared products (tile lists uploaded once)
    std::vector<GemmItem> i_bx, i_bx_all, i_gy, i_ym, i_bty, i_gx2, i_x2m, i_xv, i_btw;
    for(int i = 0; i < nsec; ++i) {
        const RsvdSec& c = P.secs[i];
        const GemmItem bx = gemm_item(c.a_off, c.x_off, c.y_off, c.mb, c.K, c.nb, c.m, c.nb, c.mb, c.tr ? 1 : 0, 0);   // Y = B X (B = A or A^T)
        i_bx_all.push_back(bx);
        i_xv.push_back(gemm_item(c.x_off, P.v_off[i], c.x_off, c.nb, c.K, c.K, c.nb, c.K, c.nb, 0, 0));
        if(c.direct) continue;
        i_bx.push_back(bx);
        i_gy.push_back(gemm_item(c.y_off, c.y_off, c.g_off, c.K, c.K, c.mb, c.mb, c.mb, c.K, 1, 0));
        i_ym.push_back(gemm_item(c.y_off, c.g_off, c.y_off, c.mb, c.K, c.K, c.mb, c.K, c.mb, 0, 0));
        i_bty.push_back(gemm_item(c.a_off, c.y_off, c.x_off, c.nb, c.K, c.mb, c.m, c.mb, c.nb, c.tr ? 0 : 1, 0));
        i_btw.push_back(gemm_item(c.a_off, P.w_off[i], c.x_off, c.nb, c.K, c.mb, c.m, c.mb, c.nb, c.tr ? 0 : 1, 0));
        i_gx2.push_back(gemm_item(c.x_off, c.x_off, c.g_off, c.K, c.K, c.nb, c.nb, c.nb, c.K, 1, 0));
        i_x2m.push_back(gemm_item(c.x_off, c.g_off, c.x_off, c.nb, c.K, c.K, c.nb, c.K, c.nb, 0, 0));
    }
    P.g_bx = ctx.prepare_gemm(P.dA, P.dX, P.dY, i_bx);
    P.g_bx_all = ctx.prepare_gemm(P.dA, P.dX, P.dY, i_bx_all);
    P.g_gy = ctx.prepare_gemm(P.dY, P.dY, P.dG, i_gy);
    P.g_ym = ctx.prepare_gemm(P.dY, P.dM, P.dY2, i_ym);
    P.g_bty = ctx.prepare_gemm(P.dA, P.dY2, P.dX2, i_bty);
    P.g_bty_raw = ctx.prepare_gemm(P.dA, P.dY, P.dX2, i_bty);   // Z = B^T (B X) without the intermediate orthonormalisation
    P.g_gx2 = ctx.prepare_gemm(P.dX2, P.dX2, P.dG, i_gx2);
    P.g_x2m = ctx.prepare_gemm(P.dX2, P.dM, P.dX, i_x2m);
    P.g_gx = ctx.prepare_gemm(P.dX, P.dX, P.dG, i_gx2);    // same shapes, other buffers (polishing)
    P.g_xm = ctx.prepare_gemm(P.dX, P.dM, P.dX2, i_x2m);
    P.g_btw = ctx.prepare_gemm(P.dA, nullptr, P.dZ, i_btw);   // Zr = B^T W: the B operand (Jacobi work matrix) is set per check
    P.g_xv = ctx.prepare_gemm(P.dX, nullptr, P.dX2, i_xv);    // V = X Vrot: the B operand (Jacobi accumulator) is set per check
    P.tol_rel = 16.0 * 2.220446049250313e-16 * std::sqrt(static_cast<double>(P.mb_max));
    P.shift_it = 1.e-13 * P.mb_max;   // keeps the factorisation of a numerically rank-deficient basis going (subspace iteration only needs a basis)
    P.sel_secs.resize(nsec);
    for(int i = 0; i < nsec; ++i) P.sel_secs[i] = {P.secs[i].s_off, P.secs[i].K, P.secs[i].direct, static_cast<double>(ctx.sym->degeneracy(L.sector[i]))};
    P.jin.d = P.dY;
    P.jin.sig = L.sig + "#rsvd" + std::to_string(K0);
    P.jin_sig_base = P.jin.sig;
    for(int i = 0; i < nsec; ++i) {
        P.jin.off.push_back(P.secs[i].y_off);
        P.jin.rows.push_back(P.secs[i].mb);
        P.jin.cols.push_back(P.secs[i].K);
    }
    return Pp;
}

}   // namespace

void Context::clear_rsvd_plans() { rsvd_plans.clear(); }

bool Context::tsvd_truncated(const Tensor& t, int64_t max_keep, double eps_svd, bool preserve_multiplets, SvdResult* out)
{
    static const int env_mode = std::getenv("XB_SVD_TRUNC") ? std::atoi(std::getenv("XB_SVD_TRUNC")) : 1;
    const int mode = opt_svd_truncated >= 0 ? opt_svd_truncated : env_mode;
    static const bool verbose = std::getenv("XB_SVD_PROF") != nullptr;
    static const bool no_graph = std::getenv("XB_SVD_NOGRAPH") != nullptr;
    static const int max_iters = std::getenv("XB_SVD_TRUNC_MAXIT") ? std::atoi(std::getenv("XB_SVD_TRUNC_MAXIT")) : 60;
    static const int first_iters = std::getenv("XB_SVD_TRUNC_IT0") ? std::atoi(std::getenv("XB_SVD_TRUNC_IT0")) : 8;
    if(mode == 0 || max_keep <= 0) return false;
    const Layout& L = *t.layout;
    const int nsec = static_cast<int>(L.sector.size());
    constexpr int KMAX = 160;
    if(max_keep + 16 > KMAX || nsec > 4096) return false;
    static const int oversampling = std::getenv("XB_SVD_TRUNC_OVERSAMPLING") ? std::atoi(std::getenv("XB_SVD_TRUNC_OVERSAMPLING")) : 32;
    const int K0 = static_cast<int>(std::min<int64_t>(KMAX, ((max_keep + oversampling + 15) / 16) * 16));
    {   // enough of a reduction?
        long long sum_nb = 0, sum_k = 0;
        int n_proj = 0;
        for(int i = 0; i < nsec; ++i) {
            const long long nb = std::min(L.rows[i], L.cols[i]);
            sum_nb += nb;
            sum_k += std::min<long long>(nb, K0);
            if(nb > K0) ++n_proj;
        }
        if(n_proj == 0 || sum_k * 2 > sum_nb || sum_k > SEL_MAX) return false;
    }
    const std::string key = L.sig + "#" + std::to_string(K0) + "#" + std::to_string(svd_call_tag);   // one plan (and warm-start state) per call site
    const auto wall0 = std::chrono::steady_clock::now();
    bool new_plan = false;
    auto it = rsvd_plans.find(key);
    if(it == rsvd_plans.end()) {
        new_plan = true;
        if(rsvd_plans.size() >= 12) {
            check_cuda(cudaStreamSynchronize(stream), "rsvd plan eviction sync");
            rsvd_plans.clear();
        }
        it = rsvd_plans.emplace(key, build_rsvd_plan(*this, L, K0)).first;
    }
    RsvdPlan& P = *it->second;
    ++P.uses;
    ++svd_truncated_calls;
    ProfScope prof(*this, CAT_SVD, 0.0);
    cudaStream_t s = stream;
    const bool use_graph = !no_graph && P.uses >= 2;   // a layout seen once (chi bases still moving) is not worth an instantiation
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float t_iter = 0.f, t_jac = 0.f, t_chk = 0.f;
    double host_ms[6] = {0, 0, 0, 0, 0, 0};   // host clock at the marks (verbose only): where the call spends wall time outside the device phases
    auto mark = [&](int k) {
        if(!verbose) return;
        if(!ev[k]) ev[k] = get_event();
        cudaEventRecord(ev[k], s);
        host_ms[k] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    };
    auto lap = [&](int a, int b, float& acc) {
        if(!verbose) return;
        cudaEventSynchronize(ev[b]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[a], ev[b]);
        acc += ms;
    };

    mark(0);
    check_cuda(cudaMemcpyAsync(P.dA, t.d, sizeof(double) * L.size, cudaMemcpyDeviceToDevice, s), "rsvd copy-in");
    static const bool no_warm = std::getenv("XB_SVD_NOWARM") != nullptr;
    const bool warm = P.have_prev && !no_warm;
    if(warm)
        run_or_replay(P, &P.graph_warm, -2, use_graph, [&]() { rsvd_start(P, true); });
    else
        run_or_replay(P, &P.graph_start, -1, use_graph, [&]() { rsvd_start(P, false); });
    P.have_prev = false;

    int iters_done = 0, next_iters = P.iters_hint > 0 ? P.iters_hint : first_iters, checks = 0;
    if(!warm && P.iters_hint > 0) next_iters = std::max(next_iters, first_iters);
    bool ok = false;
    JacobiResult J;
    Selection sel;
    while(true) {
        next_iters = std::max(1, std::min(next_iters, max_iters - iters_done));
        {
            const int n_it = next_iters;
            const bool ordered = warm && iters_done == 0;   // (a continuation starts from the Ritz vectors in Jacobi's column order: not ordered)
            // ONE graph of a single iteration, replayed n_it times, and one graph of the tail (orthonormalisation + projection)
            const int key_it = ordered ? 1 : 2, key_tail = ordered ? 3 : 4;
            for(int i = 0; i < n_it; ++i) {
                cudaGraphExec_t& slot = P.graph_iter[key_it];
                run_or_replay(P, &slot, key_it, use_graph, [&]() { rsvd_iterate(P, 1001, ordered); });
            }
            cudaGraphExec_t& slot = P.graph_iter[key_tail];
            run_or_replay(P, &slot, key_tail, use_graph, [&]() { rsvd_iterate(P, -1, ordered); });
        }
        iters_done += next_iters;
        mark(1);
        // Rayleigh-Ritz: Jacobi SVD of Y = B X.  Y's columns are an orthonormal basis pushed through B: graded like the spectrum and already sorted, which is
        // what the QR preconditioner would establish -- plain Jacobi converges in as many sweeps at D=4 (9-12) without the 0.3 ms of the preconditioner and with
        // cheaper late sweeps (D=4 step 21.1 -> 19.0 ms).  Strongly graded projected blocks (D=6: cut / smax 6e-4) can need 30 sweeps: a plan whose plain solve
        // took more than 16 switches the preconditioner back on for its later calls.
        P.jin.no_precondition = !P.jacobi_precondition;
        P.jin.sig = P.jin_sig_base + (P.jin.no_precondition ? "#np" : "");
        P.jin.sweeps_hint = P.jacobi_sweeps;   // sweeps the previous call from this call site needed
        J = jacobi_all(*this, P.jin, false, false);
        P.jacobi_sweeps = J.sweeps + 1;
        if(P.jin.no_precondition && J.sweeps + 1 > 16) P.jacobi_precondition = true;
        mark(2);
        bool layout_ok = true;
        for(int i = 0; i < nsec; ++i)
            if(J.transposed[i] || J.blocks[i].w_off != P.w_off[i] || J.blocks[i].v_off != P.v_off[i] || J.blocks[i].s_off != P.secs[i].s_off) layout_ok = false;
        if(!layout_ok) break;
        // V = X Vrot (-> X2), Zr = B^T W, residuals of the Ritz triplets
        P.g_xv.B = J.d_V;
        launch_gemm(P.g_xv);
        P.g_btw.B = J.d_W;
        launch_gemm(P.g_btw);
        if(P.n_proj > 0) {
            rsvd_residual_kernel<<<dim3(8, P.n_proj), 256, 0, s>>>(P.dZ, P.dX2, J.d_sigma, P.d_secs, P.d_proj, P.dRes);
            check_cuda(cudaGetLastError(), "rsvd_residual_kernel launch");
            ++launches;
        }
        sel = svd_select(*this, J.d_sigma, static_cast<int>(P.so), P.sel_secs, P.dRes, P.dNorm, P.tol_rel, max_keep, eps_svd, preserve_multiplets, true);
        ++checks;
        int h_bad = 0;
        check_cuda(cudaMemcpy(&h_bad, P.d_bad, sizeof(int), cudaMemcpyDeviceToHost), "rsvd flag readback");
        mark(3);
        if(verbose) {
            lap(0, 1, t_iter);
            lap(1, 2, t_jac);
            lap(2, 3, t_chk);
            std::fprintf(stderr,
                         "[rsvd] K %d sectors %d (projected %d) iters %d: verdict %d max_res %.2e (tol %.2e) rate %.3f nkeep %d cut/smax %.2e jacobi sweeps %d bad %d | ms: "
                         "iterate %.3f jacobi %.3f check %.3f graph %d warm %d\n",
                         K0, nsec, P.n_proj, iters_done, sel.hdr.verdict, sel.hdr.max_res, P.tol_rel, sel.hdr.rate, sel.hdr.nkeep,
                         sel.hdr.smax > 0 ? sel.hdr.cut / sel.hdr.smax : 0.0, J.sweeps + 1, h_bad, t_iter, t_jac, t_chk, use_graph ? 1 : 0, warm ? 1 : 0);
            std::fprintf(stderr, "[rsvd] host clock at marks (ms since entry): start %.3f iterate-enqueued %.3f jacobi-done %.3f check-done %.3f\n", host_ms[0], host_ms[1],
                         host_ms[2], host_ms[3]);
            t_iter = t_jac = t_chk = 0.f;
        }
        if(sel.hdr.verdict == 0 && !h_bad) {
            ok = true;
            break;
        }
        if((sel.hdr.verdict & 2) || h_bad || iters_done >= max_iters) break;
        // predicted number of further iterations: the residual of a kept triplet shrinks by (sigma_{K+1} / sigma_i)^2 <= rate^2 per iteration
        const double r2 = std::min(0.98, std::max(1.e-4, sel.hdr.rate * sel.hdr.rate));
        const double need = std::log(std::max(P.tol_rel / std::max(sel.hdr.max_res, 1.e-300), 1.e-300)) / std::log(r2);
        if(sel.hdr.max_res <= 1.0 && need > static_cast<double>(max_iters - iters_done)) break;   // flat spectrum: the budget cannot be met, hand over to the full path now
        next_iters = static_cast<int>(std::min<double>(std::max(2.0, std::ceil(need) + 1.0), 16.0));
        if(sel.hdr.max_res > 1.0) next_iters = 8;   // sentinel: a value next to the cut is far from converged
        // continue from the Ritz vectors
        mark(0);
        check_cuda(cudaMemcpyAsync(P.dX, P.dX2, sizeof(double) * P.xo, cudaMemcpyDeviceToDevice, s), "rsvd restart copy");
        check_cuda(cudaFreeAsync(sel.d_perm, s), "rsvd free");
        sel.d_perm = nullptr;
        free_jacobi(*this, J);
    }
    svd_truncated_iters += iters_done;
    if(ok) {
        // next call from this call site: when one projection sufficed, try fewer iterations (the warm start may need far fewer than a cold one; the
        // residual floor hides how many were spare); when it did not, remember what was needed and hold that for a while
        if(checks == 1) {
            if(P.hold > 0) {
                --P.hold;
            } else {
                // a failed probe costs a whole extra projection (Jacobi + check): only go down while the observed residual, scaled back by the measured
                // convergence factor per iteration (rate^2), predicts that fewer iterations still meet the tolerance with a factor 4 to spare
                const double per_it = std::min(0.98, std::max(1.e-4, sel.hdr.rate * sel.hdr.rate));
                int spare = 0;
                double predicted = std::max(sel.hdr.max_res, 1.e-300);
                while(spare < iters_done - 2 && predicted / per_it < 0.25 * P.tol_rel) {
                    predicted /= per_it;
                    ++spare;
                }
                P.iters_hint = std::max(2, iters_done - std::min(spare, std::max(1, iters_done / 3)));
            }
        } else {
            P.iters_hint = iters_done;
            P.hold = 6;
        }
        std::vector<FactorSrc> src(nsec);
        for(int i = 0; i < nsec; ++i) {
            const RsvdSec& c = P.secs[i];
            src[i] = {J.blocks[i].w_off, c.x_off, c.s_off, c.mb, c.nb, c.mb, c.nb, c.tr ? 0 : 1};
        }
        *out = svd_finish(*this, L, sel, src, J.d_W, P.dX2, J.d_sigma);
        // keep the Ritz vectors and their order for the next call from this call site
        check_cuda(cudaMemcpyAsync(P.dVprev, P.dX2, sizeof(double) * P.xo, cudaMemcpyDeviceToDevice, s), "rsvd warm-start save");
        check_cuda(cudaMemcpyAsync(P.d_perm_prev, sel.d_perm, sizeof(int) * P.so, cudaMemcpyDeviceToDevice, s), "rsvd warm-start save");
        P.have_prev = true;
    } else
        ++svd_truncated_fallbacks;
    if(sel.d_perm) check_cuda(cudaFreeAsync(sel.d_perm, s), "rsvd free");
    free_jacobi(*this, J);
    for(cudaEvent_t e : ev)
        if(e) event_pool.push_back(e);
    if(verbose)
        std::fprintf(stderr, "[rsvd] wall %.2f ms (new plan %d, cached plans %zu, graph %d)\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count(), new_plan ? 1 : 0, rsvd_plans.size(), use_graph ? 1 : 0);
    return ok;
}

}   // namespace xb
