// Hand-written sm_100a kernels for the Xped hot path.
//
//  permute_bulk_kernel / permute_kernel
//                      batched strided scaled-accumulate tensor transpose: the whole of
//                      Tensor::permute / permute_adj (src/Core/Tensor.cpp:83-214,258-333) in ONE launch.
//                      HBM-bound: every source element is read once, every destination element written once.
//                      Bulk form: tiles move through shared memory with the TMA engine's 1-D cp.async.bulk (one
//                      operation per contiguous run; needs 16-byte aligned runs of even length, fan-in 1);
//                      table-driven form (permute_lut_kernel) for SU(2) recoupling sums / odd run lengths: all index arithmetic tabulated at plan time;
//                      LSU form for what is left (sharded pack / gather, plans whose tables would be large next to the tensor).
//                      Tensor maps (cp.async.bulk.tensor) are not used: sub-block origins and leading dimensions are
//                      arbitrary multiples of 8 bytes and change with every buffer.
//  (gemm.cu)           gemm_grouped_kernel: ragged per-sector C_q += op(A_q) * op(B_q) (src/Core/TensorBase.cpp:262-279) on
//                      the FP64 tensor cores (mma.sync m8n8k4 DMMA; tcgen05 has no FP64 kind).
//  small fused elementwise / reduction kernels for AdjointOp, DiagCoeffUnaryOp, maxNorm, trace.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "engine.hpp"

namespace xb {

// ================================================================================================ permute
// Decompose work item `it` of one phase: x along the fastest tile leg `order[0]`, strip start c1 along `order[1]`, the
// other tile legs; returns false if the item lies outside the (clipped) tile.  `stride` holds the global strides of the
// side being addressed (source strides in the read phase, destination strides in the write phase).
__device__ __forceinline__ bool perm_item(const PermTask& T, const int* __restrict__ order, const int* __restrict__ stride, const int* __restrict__ lim,
                                          int nchunk, unsigned magic_chunk, int strip, int it, int& off, int& pos, int& cnt, int& gstep, int& sstep)
{
    const int ntl = T.ntl;
    if(ntl == 0) {
        off = 0;
        pos = 0;
        cnt = 1;
        gstep = sstep = 0;
        return true;
    }
    const int la = order[0];
    int r = __umulhi(static_cast<unsigned>(it), T.magic[la]);
    const int x = it - r * T.w[la];
    bool ok = x < lim[la];
    off = x * stride[la];
    pos = x * T.smstride[la];
    cnt = 1;
    gstep = sstep = 0;
    if(ntl >= 2) {
        const int lb = order[1];
        int q = nchunk > 1 ? __umulhi(static_cast<unsigned>(r), magic_chunk) : r;
        const int c1 = (r - q * nchunk) * strip;
        r = q;
        cnt = min(strip, lim[lb] - c1);
        gstep = stride[lb];
        sstep = T.smstride[lb];
        off += c1 * gstep;
        pos += c1 * sstep;
        for(int j = 2; j < ntl; ++j) {
            const int l = order[j];
            q = __umulhi(static_cast<unsigned>(r), T.magic[l]);
            const int c = r - q * T.w[l];
            r = q;
            ok = ok && (c < lim[l]);
            off += c * stride[l];
            pos += c * T.smstride[l];
        }
    }
    return ok && cnt > 0;
}

// Read phase for one source: lanes along the source-contiguous leg, each thread walks strips of the second-fastest
// leg; the loads of a strip are issued back to back, then scaled / accumulated into the shared tile.
__device__ __forceinline__ void perm_read(const PermTask& T, const int* __restrict__ sstride, const int* __restrict__ lim, const double* __restrict__ sp,
                                          double coeff, bool first, double* __restrict__ buf, int tid)
{
    const int strip = T.strip_s, items = T.items_s, nchunk = T.nchunk_s;
    const unsigned magic = T.magic_chunk_s;
    for(int it = tid; it < items; it += PERM_THREADS) {
        int off, pos, cnt, gstep, sstep;
        if(!perm_item(T, T.tl_s, sstride, lim, nchunk, magic, strip, it, off, pos, cnt, gstep, sstep)) continue;
        double v[PERM_STRIP];
#pragma unroll
        for(int i = 0; i < PERM_STRIP; ++i)
            if(i < cnt) v[i] = __ldg(sp + off + i * gstep);
        if(first) {
#pragma unroll
            for(int i = 0; i < PERM_STRIP; ++i)
                if(i < cnt) buf[pos + i * sstep] = coeff * v[i];
        } else {
#pragma unroll
            for(int i = 0; i < PERM_STRIP; ++i)
                if(i < cnt) buf[pos + i * sstep] = fma(coeff, v[i], buf[pos + i * sstep]);
        }
    }
}

// Each CTA walks `tiles_per_cta` consecutive entries of the tile list (mostly tiles of one task, so the task record is
// staged once; the next tile's record is prefetched while the current tile moves).  Per tile: read phase (source
// order) -> barrier -> write phase (destination order) -> barrier.
__global__ void __launch_bounds__(PERM_THREADS, 4)
permute_kernel(const double* __restrict__ src, const double* __restrict__ src2, long long split, double* __restrict__ dst,
               const PermTask* __restrict__ tasks, const PermSrc* __restrict__ srcs, const PermTile* __restrict__ tiles, int n_tiles, int tiles_per_cta,
               const PeerDelta delta)
{
    __shared__ PermTask T;
    __shared__ PermSrc S;
    __shared__ double buf[PERM_BUF_ELEMS];
    __shared__ int s_tile[2][sizeof(PermTile) / sizeof(int)];
    __shared__ int s_base[PERM_MAXLEG];
    __shared__ int s_lim[PERM_MAXLEG];

    constexpr int TILE_INTS = static_cast<int>(sizeof(PermTile) / sizeof(int));
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * tiles_per_cta, t1 = min(n_tiles, t0 + tiles_per_cta);
    if(tid < TILE_INTS) s_tile[0][tid] = reinterpret_cast<const int*>(tiles + t0)[tid];
    __syncthreads();
    int staged = -1;
    for(int ti = t0; ti < t1; ++ti) {
        const int cur = (ti - t0) & 1;
        int nxt_word = 0;
        if(tid < TILE_INTS && ti + 1 < t1) nxt_word = reinterpret_cast<const int*>(tiles + ti + 1)[tid];   // prefetch, consumed at the end of the iteration
        const PermTile& tile = *reinterpret_cast<const PermTile*>(s_tile[cur]);
        const int task = tile.task;
        if(task != staged) {
            const int* g = reinterpret_cast<const int*>(tasks + task);
            int* s = reinterpret_cast<int*>(&T);
            for(int i = tid; i < static_cast<int>(sizeof(PermTask) / sizeof(int)); i += PERM_THREADS) s[i] = g[i];
            staged = task;
            __syncthreads();
        }
        if(tid < PERM_MAXLEG) {
            const bool on = tid < T.nleg;
            const int b = on ? tile.c[tid] * T.w[tid] : 0;
            s_base[tid] = b;
            s_lim[tid] = on ? min(T.w[tid], T.dims[tid] - b) : 1;
        }
        __syncthreads();

        // ---- read phase: buf = sum_k coeff_k * src_k tile. The item -> thread map does not depend on k, so the same thread
        // owns the same buf element for every source and the fan-in accumulates in shared memory without barriers.
        if(T.src_count > 0) {
            long long sbase = T.src0_off;
            const double* sb = src;
            if(sbase >= split) {   // sharded permute: this source sits in the receive buffer
                sb = src2;
                sbase -= split;
            }
            for(int l = 0; l < T.nleg; ++l) sbase += static_cast<long long>(s_base[l]) * T.src0_stride[l];
            perm_read(T, T.src0_stride, s_lim, sb + sbase, T.src0_coeff, true, buf, tid);
        }
        for(int k = 1; k < T.src_count; ++k) {
            __syncthreads();   // S of the previous source no longer needed
            if(tid < static_cast<int>(sizeof(PermSrc) / sizeof(int))) reinterpret_cast<int*>(&S)[tid] = reinterpret_cast<const int*>(srcs + T.src_begin + k)[tid];
            __syncthreads();
            long long sbase = S.off;
            const double* sb = src;
            if(sbase >= split) {
                sb = src2;
                sbase -= split;
            }
            for(int l = 0; l < T.nleg; ++l) sbase += static_cast<long long>(s_base[l]) * S.sstride[l];
            perm_read(T, S.sstride, s_lim, sb + sbase, S.coeff, false, buf, tid);
        }
        __syncthreads();

        // ---- write phase: destination order (lanes along the destination-contiguous leg); every element written once
        long long dbase = (T.dst_off & PERM_OFF_MASK) + delta.d[T.dst_off >> PERM_PEER_SHIFT];   // destination shard of this or of a peer rank
        for(int l = 0; l < T.nleg; ++l) dbase += static_cast<long long>(s_base[l]) * T.dstride[l];
        double* __restrict__ dp = dst + dbase;
        const bool zero = T.src_count == 0;
        const int strip = T.strip_d;
        for(int it = tid; it < T.items_d; it += PERM_THREADS) {
            int off, pos, cnt, gstep, sstep;
            if(!perm_item(T, T.tl_d, T.dstride, s_lim, T.nchunk_d, T.magic_chunk_d, strip, it, off, pos, cnt, gstep, sstep)) continue;
#pragma unroll
            for(int i = 0; i < PERM_STRIP; ++i)
                if(i < cnt) dp[off + i * gstep] = zero ? 0.0 : buf[pos + i * sstep];
        }
        if(tid < TILE_INTS) s_tile[cur ^ 1][tid] = nxt_word;
        __syncthreads();   // buf / T / s_base free for the next tile; next tile record visible
    }
}

// ================================================================================================ permute, bulk-copy form
// Tiles move through shared memory with cp.async.bulk (the TMA engine, 1-D form): one instruction per contiguous run,
// completion of the loads on an mbarrier (complete_tx), stores as bulk groups.  The SM's load/store units and the
// integer pipe only see the in-shared-memory transpose (a 2 x 16 lane grid: 16 consecutive source elements, two adjacent
// destination elements; run pitches with pitch/2 odd make both sides conflict-free), no per-element global addressing.
// quotient by a small divisor through its magic number ceil(2^32 / d); d == 1 has no 32-bit magic number
__device__ __forceinline__ int magic_div(int v, int d, unsigned magic) { return d == 1 ? v : static_cast<int>(__umulhi(static_cast<unsigned>(v), magic)); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bulk_load(unsigned dst_smem, const double* gsrc, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_smem), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(double* gdst, unsigned src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    // bounded spin: a byte-count mismatch must surface as a launch failure, never as a hung GPU
    const long long t0 = clock64();
    for(;;) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if(done) return;
        if(clock64() - t0 > 4000000000ll) __trap();
    }
}

// A CTA works on a range of consecutive tiles of ONE task (ranges are cut on the host).  Two staging buffers: copy mode
// alternates them (the loads of tile i+1 are in flight while tile i drains through its bulk stores); transpose mode
// receives the source runs in buffer 0 and builds the destination runs in buffer 1 (the loads of tile i+1 overlap the
// stores of tile i).  Five CTAs per SM hide the rest of the latency -- measured to beat a deeper per-CTA ring at
// three CTAs per SM, because the issue rate of small (192-byte) bulk operations is what limits the short-run cases.
// Runs that are contiguous in global memory AND in the staging buffer are merged into one bulk operation.
constexpr int BULK_MAX_TILES_PER_CTA = 32;
__global__ void __launch_bounds__(BULK_THREADS, 5)
permute_bulk_kernel(const double* __restrict__ src, double* __restrict__ dst, const BulkTask* __restrict__ tasks, const PermTile* __restrict__ tiles,
                    const int2* __restrict__ ranges, const PeerDelta delta)
{
    __shared__ __align__(128) double bufA[BULK_BUF_ELEMS];
    __shared__ __align__(128) double bufB[BULK_BUF_ELEMS];
    __shared__ __align__(8) unsigned long long bar_storage[2];
    __shared__ BulkTask T;
    __shared__ PermTile s_tiles[BULK_MAX_TILES_PER_CTA];

    constexpr int TILE_INTS = static_cast<int>(sizeof(PermTile) / sizeof(int));
    const int tid = threadIdx.x;
    const int2 range = ranges[blockIdx.x];
    const int t0 = range.x, nt = range.y;
    double* const buf[2] = {bufA, bufB};
    const unsigned bar[2] = {smem_u32(&bar_storage[0]), smem_u32(&bar_storage[1])};
    if(tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar[0]));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar[1]));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for(int i = tid; i < nt * TILE_INTS; i += BULK_THREADS) reinterpret_cast<int*>(s_tiles)[i] = reinterpret_cast<const int*>(tiles + t0)[i];
    {
        const int* g = reinterpret_cast<const int*>(tasks + tiles[t0].task);
        int* sdst = reinterpret_cast<int*>(&T);
        for(int i = tid; i < static_cast<int>(sizeof(BulkTask) / sizeof(int)); i += BULK_THREADS) sdst[i] = g[i];
    }
    __syncthreads();
    const bool copy = T.ls == 0;

    // Every thread owns the same (up to two) load runs and store runs in every tile of the CTA, and bulk plans only use tile
    // extents that divide the leg extents (no clipped tiles): byte counts and offsets relative to the tile origin are
    // computed once, before the tile loop.
    //   transpose mode: load run = (x, z), r = z * wx + x;  store run = (y, z), r = z * wy + y.   copy mode: r = z.
    // merge group g > 1: the runs r .. r + g - 1 (consecutive along the first enumerated index) are contiguous in global memory and
    // in the staging buffer; only the leader issues, for the whole group.
    struct Run { long long rel; int smem; unsigned bytes; };   // bytes == 0: nothing to issue
    auto prep = [&](bool is_load, int r) {
        Run R;
        R.rel = 0;
        R.smem = 0;
        R.bytes = 0u;
        const int nrun = copy ? T.nz : T.nz * (is_load ? T.wx : T.wy);
        if(r >= nrun) return R;
        const int group = is_load ? T.mrg_load : T.mrg_store;
        const int* __restrict__ gs = is_load ? T.sstride : T.dstride;
        int z = r, count = 1;
        bool leader = true;
        int group_index = r;   // position of the run's (z_rest, x) group in the staging buffer (loads of the transpose mode)
        int zc_in_group = 0;
        if(!copy) {
            if(is_load && T.nzc > 1) {   // z-merge: r = (z_rest * wx + x) * nzc + zc
                const int t = magic_div(r, T.nzc, T.magic_nzc);
                zc_in_group = r - t * T.nzc;
                const int zr = magic_div(t, T.wx, T.magic_wx);
                const int x = t - zr * T.wx;
                group_index = t;
                z = zr * T.nzc + zc_in_group;
                R.rel += static_cast<long long>(x) * gs[0];
                leader = zc_in_group == 0;
                count = T.nzc;
            } else {
                const int wi = is_load ? T.wx : T.wy;
                z = magic_div(r, wi, is_load ? T.magic_wx : T.magic_wy);
                const int inner = r - z * wi;
                R.rel += static_cast<long long>(inner) * gs[is_load ? 0 : T.ls];
                if(group > 1) {
                    leader = inner == 0;
                    count = group;
                }
            }
        }
        for(int i = 0; i < T.nzleg; ++i) {
            const int l = T.zleg[i];
            const int q = magic_div(z, T.w[l], T.zmagic[i]);
            const int c = z - q * T.w[l];
            z = q;
            R.rel += static_cast<long long>(c) * gs[l];
            if(copy && i == 0 && group > 1) {
                leader = c == 0;
                count = group;
            }
        }
        if(leader) R.bytes = static_cast<unsigned>((is_load ? (copy ? T.wx : T.wy) : T.wx) * count) * 8u;
        R.smem = (!copy && is_load) ? group_index * T.gp + zc_in_group * T.wy : r * (is_load ? T.rs_s : T.rs_d);
        return R;
    };
    const Run ld0 = prep(true, tid), ld1 = prep(true, tid + BULK_THREADS), st0 = prep(false, tid), st1 = prep(false, tid + BULK_THREADS);
    unsigned tile_bytes = 8u;
    for(int l = 0; l < T.nleg; ++l) tile_bytes *= static_cast<unsigned>(T.w[l]);

    auto origin = [&](int i, bool is_load) {
        const PermTile& tile = s_tiles[i];
        long long o = is_load ? T.src_off : (T.dst_off & PERM_OFF_MASK) + delta.d[T.dst_off >> PERM_PEER_SHIFT];
        const int* __restrict__ gs = is_load ? T.sstride : T.dstride;
        for(int l = 0; l < T.nleg; ++l) o += static_cast<long long>(tile.c[l] * T.w[l]) * gs[l];
        return o;
    };
    auto issue_loads = [&](int i, int slot) {
        const double* __restrict__ base = src + origin(i, true);
        if(tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar[slot]), "r"(tile_bytes) : "memory");
        if(ld0.bytes) bulk_load(smem_u32(buf[slot] + ld0.smem), base + ld0.rel, ld0.bytes, bar[slot]);
        if(ld1.bytes) bulk_load(smem_u32(buf[slot] + ld1.smem), base + ld1.rel, ld1.bytes, bar[slot]);
    };
    auto issue_stores = [&](int i, int slot) {
        double* __restrict__ base = dst + origin(i, false);
        if(st0.bytes) bulk_store(base + st0.rel, smem_u32(buf[slot] + st0.smem), st0.bytes);
        if(st1.bytes) bulk_store(base + st1.rel, smem_u32(buf[slot] + st1.smem), st1.bytes);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    };

    if(copy) {
        // tile i: buffer / barrier i & 1
        issue_loads(0, 0);
        for(int i = 0; i < nt; ++i) {
            const int slot = i & 1;
            if(i + 1 < nt) {
                // the other buffer was drained by the stores of tile i - 1: they must have finished reading it
                asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                __syncthreads();
                issue_loads(i + 1, slot ^ 1);
            }
            mbar_wait(bar[slot], static_cast<unsigned>(i >> 1) & 1u);
            issue_stores(i, slot);
        }
    } else {
        // lane grid gx x gy (2 x 16, or 4 x 8 when wy is a multiple of 8 but not of 16): gy consecutive source elements per row
        // (conflict-free for any pitch), gx adjacent destination elements (destination pitch / 2 odd => distinct bank pairs).
        // A warp walks units (z, x_hi); inside a unit only y advances, by constant address increments.
        const int rs_d = T.rs_d, wx = T.wx, wy = T.wy, nxh = T.nxh, nzc = T.nzc, gp = T.gp;
        const unsigned m_nxh = T.magic_nxh, m_nzc = T.magic_nzc;
        const double coeff = T.coeff;
        const int gxs = T.gx_shift, gx = 1 << gxs, gy = 32 >> gxs;
        const int lane = tid & 31, warp = tid >> 5;
        const int y_lo = lane & (gy - 1), x_lo = lane >> (5 - gxs);
        const int units = T.nz * nxh;
        const double* __restrict__ S = bufA;
        double* __restrict__ D = bufB;
        issue_loads(0, 0);
        for(int i = 0; i < nt; ++i) {
            mbar_wait(bar[0], static_cast<unsigned>(i) & 1u);
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the stores of tile i - 1 have finished reading D
            __syncthreads();
            for(int u = warp; u < units; u += BULK_THREADS / 32) {
                const int z = magic_div(u, nxh, m_nxh);
                const int x = (u - z * nxh) * gx + x_lo;
                if(x < wx) {
                    const int zr = magic_div(z, nzc, m_nzc), zc = z - zr * nzc;
                    const double* __restrict__ sp = S + (zr * wx + x) * gp + zc * wy + y_lo;
                    double* __restrict__ dp = D + (z * wy + y_lo) * rs_d + x;
                    for(int y = y_lo; y < wy; y += gy) {
                        *dp = coeff * *sp;
                        sp += gy;
                        dp += gy * rs_d;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes to D -> visible to the bulk stores
            __syncthreads();
            if(i + 1 < nt) issue_loads(i + 1, 0);   // overlaps the stores of this tile
            issue_stores(i, 1);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // all stores complete before the shared memory is released
}

// ================================================================================================ permute, table-driven form
// SU(2) recoupling sums and odd extents (no bulk form): the LSU form spends ~240 thread instructions per destination element on mixed-radix index
// arithmetic (sub-blocks of ~40 k elements with extents like 4 x 3 x 8 x 3 x 3 x 3).  Here the host tabulates, per distinct tile shape, the offsets of
// every tile element for every source and for the destination (LutTask, layout.cpp): a thread reads (table entry, source element) pairs, accumulates its
// <= 8 tile elements over ALL sources in registers, stages them once, and the second phase stores them in destination order.  Three block-wide barriers
// per tile, no barrier between sources.
__global__ void __launch_bounds__(LUT_THREADS, 4)
permute_lut_kernel(const double* __restrict__ src, double* __restrict__ dst, const LutTask* __restrict__ tasks, const LutTileSrc* __restrict__ tsrc,
                   const int* __restrict__ tile_task, const long long* __restrict__ tile_dst, int n_tiles, int tiles_per_cta, const int* __restrict__ tab_int,
                   const unsigned short* __restrict__ tab_pos)
{
    constexpr int PER_THREAD = LUT_TILE_ELEMS / LUT_THREADS;   // 8
    constexpr int BUF = LUT_TILE_ELEMS + LUT_TILE_ELEMS / 32 + 8;
    __shared__ double buf[2][BUF];   // two staging buffers: ONE block-wide barrier per tile
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * tiles_per_cta, t1 = min(n_tiles, t0 + tiles_per_cta);
    for(int ti = t0; ti < t1; ++ti) {
        // task record, tile origins and coefficients are uniform across the CTA: broadcast loads, no staging
        const LutTask* __restrict__ T = tasks + __ldg(tile_task + ti);
        const int vol = __ldg(&T->vol), nsrc = __ldg(&T->src_count);
        const LutTileSrc* __restrict__ ts = tsrc + __ldg(&T->tsrc_begin) + (ti - __ldg(&T->first_tile)) * nsrc;
        const int* __restrict__ tab = tab_int + __ldg(&T->soff_tab);
        double acc[PER_THREAD];
        int o_cur[PER_THREAD];
#pragma unroll
        for(int j = 0; j < PER_THREAD; ++j) {
            acc[j] = 0.0;
            const int e = tid + j * LUT_THREADS;
            o_cur[j] = (nsrc > 0 && e < vol) ? __ldg(tab + e) : 0;
        }
        for(int k = 0; k < nsrc; ++k) {
            const double* __restrict__ sp = src + __ldg(&ts[k].off);
            const double c = __ldg(&ts[k].coeff);
            double v[PER_THREAD];
            int o_nxt[PER_THREAD];
#pragma unroll
            for(int j = 0; j < PER_THREAD; ++j) {
                const int e = tid + j * LUT_THREADS;
                v[j] = e < vol ? __ldg(sp + o_cur[j]) : 0.0;
            }
#pragma unroll
            for(int j = 0; j < PER_THREAD; ++j) {   // offsets of the next source are in flight together with this source's data
                const int e = tid + j * LUT_THREADS;
                o_nxt[j] = (k + 1 < nsrc && e < vol) ? __ldg(tab + (k + 1) * vol + e) : 0;
            }
#pragma unroll
            for(int j = 0; j < PER_THREAD; ++j) {
                acc[j] = fma(c, v[j], acc[j]);
                o_cur[j] = o_nxt[j];
            }
        }
        double* __restrict__ sb = buf[ti & 1];
        const int* __restrict__ dtab = tab_int + __ldg(&T->doff_tab);
        const unsigned short* __restrict__ ptab = tab_pos + __ldg(&T->spos_tab);
        int d_off[PER_THREAD];
        int d_pos[PER_THREAD];
#pragma unroll
        for(int j = 0; j < PER_THREAD; ++j) {   // tables of the write phase: issued before the barrier
            const int e = tid + j * LUT_THREADS;
            d_off[j] = e < vol ? __ldg(dtab + e) : 0;
            d_pos[j] = e < vol ? static_cast<int>(__ldg(ptab + e)) : 0;
        }
#pragma unroll
        for(int j = 0; j < PER_THREAD; ++j) {
            const int e = tid + j * LUT_THREADS;
            if(e < vol) sb[e + (e >> 5)] = acc[j];
        }
        __syncthreads();
        double* __restrict__ dp = dst + __ldg(tile_dst + ti);
#pragma unroll
        for(int j = 0; j < PER_THREAD; ++j) {
            const int e = tid + j * LUT_THREADS;
            if(e < vol) dp[d_off[j]] = sb[d_pos[j]];
        }
        // no second barrier: the next tile stages into the other buffer, and a thread reaches the tile after that only through the next barrier
    }
}

static const PeerDelta NO_PEERS = {{0, 0, 0, 0, 0, 0, 0, 0}};

static void launch_permute2_peers(const double* src, const double* src2, long long split, double* dst, const PermutePlan& plan, const PeerDelta& delta, cudaStream_t s)
{
    if(plan.n_tiles == 0) return;
    const int tpc = plan.tiles_per_cta < 1 ? 1 : plan.tiles_per_cta;
    const int grid = (plan.n_tiles + tpc - 1) / tpc;
    permute_kernel<<<grid, PERM_THREADS, 0, s>>>(src, src2, split, dst, plan.d_tasks, plan.d_srcs, plan.d_tiles, plan.n_tiles, tpc, delta);
    check_cuda(cudaGetLastError(), "permute_kernel launch");
}

void launch_permute2(const double* src, const double* src2, long long split, double* dst, const PermutePlan& plan, cudaStream_t s)
{
    launch_permute2_peers(src, src2, split, dst, plan, NO_PEERS, s);
}

void launch_permute(const double* src, double* dst, const PermutePlan& plan, cudaStream_t s) { launch_permute_peers(src, dst, plan, NO_PEERS, s); }

void launch_permute_peers(const double* src, double* dst, const PermutePlan& plan, const PeerDelta& delta, cudaStream_t s)
{
    // bulk-copy form for bandwidth-bound sizes (XB_PERMUTE_BULK=0 / 1 forces never / whenever the plan qualifies)
    const char* bulk_env = std::getenv("XB_PERMUTE_BULK");   // read per call so that tests can toggle it
    const int bulk_mode = bulk_env ? std::atoi(bulk_env) : -1;
    bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    for(long long d : delta.d) aligned = aligned && (d & 1) == 0;   // peer shards: 16-byte aligned relative to dst
    if(plan.d_btasks && aligned && bulk_mode != 0 && (bulk_mode == 1 || plan.bytes >= (8ll << 20))) {
        permute_bulk_kernel<<<plan.n_branges, BULK_THREADS, 0, s>>>(src, dst, plan.d_btasks, plan.d_btiles, plan.d_branges, delta);
        check_cuda(cudaGetLastError(), "permute_bulk_kernel launch");
        return;
    }
    // table-driven form (plans with recoupling sums / odd extents, no peers): XB_PERMUTE_LUT=0 switches it off
    const char* lut_env = std::getenv("XB_PERMUTE_LUT");
    const int lut_mode = lut_env ? std::atoi(lut_env) : -1;
    if(plan.d_ltasks && lut_mode != 0) {   // (no size threshold: also ahead of the LSU form on the launch-bound tensors of a D=4 CTM step, 1.12 vs 1.24 ms)
        const int tpc = plan.ltiles_per_cta < 1 ? 1 : plan.ltiles_per_cta;
        const int grid = (plan.n_ltiles + tpc - 1) / tpc;
        permute_lut_kernel<<<grid, LUT_THREADS, 0, s>>>(src, dst, plan.d_ltasks, plan.d_ltsrc, plan.d_ltile_task, plan.d_ltile_dst, plan.n_ltiles, tpc, plan.d_lut_int,
                                                        plan.d_lut_pos);
        check_cuda(cudaGetLastError(), "permute_lut_kernel launch");
        return;
    }
    launch_permute2_peers(src, src, INT64_MAX, dst, plan, delta, s);
}

// ================================================================================================ elementwise
__global__ void scale_copy_kernel(const double* __restrict__ x, double* __restrict__ y, double alpha, long long n)
{
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = alpha * x[i];
}

__global__ void axpby_kernel(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ z, double alpha, double beta, long long n)
{
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) z[i] = alpha * x[i] + beta * y[i];
}

static int grid_for(long long n, int threads, int sms = 148)
{
    long long g = (n + threads - 1) / threads;
    const long long cap = static_cast<long long>(sms) * 8;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

void launch_scale_copy(const double* x, double* y, double alpha, int64_t n, cudaStream_t s)
{
    if(n == 0) return;
    scale_copy_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, y, alpha, n);
    check_cuda(cudaGetLastError(), "scale_copy_kernel launch");
}

void launch_axpby(const double* x, const double* y, double* z, double alpha, double beta, int64_t n, cudaStream_t s)
{
    if(n == 0) return;
    axpby_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, y, z, alpha, beta, n);
    check_cuda(cudaGetLastError(), "axpby_kernel launch");
}

__global__ void fill_constant_kernel(double* __restrict__ x, double v, long long n)
{
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}

// Counter-based uniform[-1,1): splitmix64 of (seed, element index) -> 53 random mantissa bits. The oracle's
// helper (oracle/xped_oracle/rng.py) computes the same integers with numpy, so fills are bit-identical.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_random_kernel(double* __restrict__ x, unsigned long long seed, long long n, long long index_base)
{
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const unsigned long long key = splitmix64(seed);
    for(long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long r = splitmix64(key ^ (static_cast<unsigned long long>(i + index_base) * 0xD1342543DE82EF95ull));
        x[i] = static_cast<double>(r >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}

static int grid_for(long long n, int threads, int sms);

void launch_fill_constant(double* x, double v, int64_t n, cudaStream_t s)
{
    if(n == 0) return;
    fill_constant_kernel<<<grid_for(n, 256, 148), 256, 0, s>>>(x, v, n);
    check_cuda(cudaGetLastError(), "fill_constant_kernel launch");
}

void launch_fill_random(double* x, uint64_t seed, int64_t n, cudaStream_t s, int64_t index_base)
{
    if(n == 0) return;
    fill_random_kernel<<<grid_for(n, 256, 148), 256, 0, s>>>(x, seed, n, index_base);
    check_cuda(cudaGetLastError(), "fill_random_kernel launch");
}

// max |x| : order-independent, so a plain atomic max on the bit pattern of a non-negative double is exact.
__global__ void absmax_kernel(const double* __restrict__ x, long long n, unsigned long long* out)
{
    double m = 0.0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) m = fmax(m, fabs(x[i]));
    for(int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double sm[8];
    if((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if(threadIdx.x == 0) {
        for(int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) m = fmax(m, sm[w]);
        atomicMax(out, static_cast<unsigned long long>(__double_as_longlong(m)));
    }
}

void launch_absmax(const double* x, int64_t n, double* d_out, cudaStream_t s)
{
    check_cuda(cudaMemsetAsync(d_out, 0, sizeof(double), s), "absmax memset");
    if(n == 0) return;
    absmax_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, n, reinterpret_cast<unsigned long long*>(d_out));
    check_cuda(cudaGetLastError(), "absmax_kernel launch");
}

// per-block transpose through a 32x33 shared tile: AdjointOp::eval() (include/Xped/Core/AdjointOp.hpp:44-65)
__global__ void transpose_blocks_kernel(const double* __restrict__ src, double* __restrict__ dst, const BlockDesc* __restrict__ blocks,
                                        const long long* __restrict__ dst_off, int nblocks)
{
    __shared__ double tile[32][33];
    for(int b = blockIdx.y; b < nblocks; b += gridDim.y) {
        const BlockDesc d = blocks[b];
        const int tr = (d.rows + 31) / 32, tc = (d.cols + 31) / 32;
        for(int t = blockIdx.x; t < tr * tc; t += gridDim.x) {
            const int r0 = (t % tr) * 32, c0 = (t / tr) * 32;
            for(int j = threadIdx.y; j < 32; j += blockDim.y) {
                const int r = r0 + threadIdx.x, c = c0 + j;
                if(r < d.rows && c < d.cols) tile[j][threadIdx.x] = src[d.off + r + static_cast<long long>(c) * d.rows];
            }
            __syncthreads();
            for(int j = threadIdx.y; j < 32; j += blockDim.y) {
                const int c = c0 + threadIdx.x, r = r0 + j;   // output is cols x rows, column-major
                if(r < d.rows && c < d.cols) dst[dst_off[b] + c + static_cast<long long>(r) * d.cols] = tile[threadIdx.x][j];
            }
            __syncthreads();
        }
    }
}

void launch_transpose_blocks(const double* src, double* dst, const BlockDesc* d_src_blocks, const long long* d_dst_off, int nblocks, cudaStream_t s)
{
    if(nblocks == 0) return;
    dim3 grid(148, static_cast<unsigned>(nblocks < 64 ? nblocks : 64)), block(32, 8);
    transpose_blocks_kernel<<<grid, block, 0, s>>>(src, dst, d_src_blocks, d_dst_off, nblocks);
    check_cuda(cudaGetLastError(), "transpose_blocks_kernel launch");
}

// DiagCoeffUnaryOp (include/Xped/Core/DiagCoeffUnaryOp.hpp:48-51): out = diag(f(diag(in))), off-diagonals zero.
__global__ void diag_map_kernel(const double* __restrict__ src, double* __restrict__ dst, const BlockDesc* __restrict__ blocks, int nblocks, int op)
{
    for(int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const BlockDesc d = blocks[b];
        const long long n = static_cast<long long>(d.rows) * d.cols;
        for(long long e = threadIdx.x; e < n; e += blockDim.x) {
            const int r = static_cast<int>(e % d.rows), c = static_cast<int>(e / d.rows);
            double v = 0.0;
            if(r == c) {
                const double x = src[d.off + e];
                v = op == 0 ? sqrt(x) : (op == 1 ? 1.0 / x : (op == 2 ? 1.0 / sqrt(x) : x));   // op 3: keep the diagonal
            }
            dst[d.off + e] = v;
        }
    }
}

void launch_diag_map(const double* src, double* dst, const BlockDesc* d_blocks, int nblocks, int op, cudaStream_t s)
{
    if(nblocks == 0) return;
    diag_map_kernel<<<nblocks < 296 ? nblocks : 296, 256, 0, s>>>(src, dst, d_blocks, nblocks, op);
    check_cuda(cudaGetLastError(), "diag_map_kernel launch");
}

// Tensor::twist for fermionic symmetries (src/Core/Tensor.cpp:365-375): scale tree-pair sub-blocks by +-1.
__global__ void scale_subblocks_kernel(double* __restrict__ data, const SubScale* __restrict__ subs, int n)
{
    for(int b = blockIdx.x; b < n; b += gridDim.x) {
        const SubScale d = subs[b];
        const long long cnt = static_cast<long long>(d.rows) * d.cols;
        for(long long e = threadIdx.x; e < cnt; e += blockDim.x) {
            const long long idx = d.off + (e % d.rows) + (e / d.rows) * d.ld;
            data[idx] *= d.scale;
        }
    }
}

void launch_scale_subblocks(double* data, const SubScale* d_subs, int n, cudaStream_t s)
{
    if(n == 0) return;
    scale_subblocks_kernel<<<n < 1184 ? n : 1184, 256, 0, s>>>(data, d_subs, n);
    check_cuda(cudaGetLastError(), "scale_subblocks_kernel launch");
}

// trace = sum_q weight_q * tr(block_q) (src/Core/TensorBase.cpp:27-37). One block, fixed summation order => deterministic.
__global__ void trace_kernel(const double* __restrict__ src, const BlockDesc* __restrict__ blocks, int nblocks, double* out)
{
    __shared__ double sm[256];
    double acc = 0.0;
    for(int b = 0; b < nblocks; ++b) {
        const BlockDesc d = blocks[b];
        const int n = d.rows < d.cols ? d.rows : d.cols;
        double part = 0.0;
        for(int i = threadIdx.x; i < n; i += blockDim.x) part += src[d.off + i + static_cast<long long>(i) * d.rows];
        acc += d.weight * part;
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for(int o = 128; o > 0; o >>= 1) {
        if(threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if(threadIdx.x == 0) *out = sm[0];
}

void launch_trace(const double* src, const BlockDesc* d_blocks, int nblocks, double* d_out, cudaStream_t s)
{
    trace_kernel<<<1, 256, 0, s>>>(src, d_blocks, nblocks, d_out);
    check_cuda(cudaGetLastError(), "trace_kernel launch");
}

}   // namespace xb

namespace xb {
// y = x / (*d_scalar): the max-norm normalisation of renormalize_* (src/PEPS/CTM.cpp:1369,1381,1409) without a host sync.
__global__ void scale_by_inv_kernel(const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ d_scalar, long long n)
{
    const double inv = 1. / *d_scalar;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = x[i] * inv;
}

void launch_scale_by_inv(const double* x, double* y, const double* d_scalar, int64_t n, cudaStream_t s)
{
    if(n == 0) return;
    long long g = (n + 255) / 256;
    if(g > 1184) g = 1184;
    scale_by_inv_kernel<<<static_cast<int>(g), 256, 0, s>>>(x, y, d_scalar, n);
    check_cuda(cudaGetLastError(), "scale_by_inv_kernel launch");
}

__global__ void set_identity_kernel(double* __restrict__ x, const BlockDesc* __restrict__ blocks, int nblocks)
{
    for(int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const BlockDesc d = blocks[b];
        const int n = d.rows < d.cols ? d.rows : d.cols;
        for(int i = threadIdx.x; i < n; i += blockDim.x) x[d.off + i + static_cast<long long>(i) * d.rows] = 1.0;
    }
}

void launch_set_identity(double* x, const BlockDesc* d_blocks, int nblocks, cudaStream_t s)
{
    if(nblocks == 0) return;
    set_identity_kernel<<<nblocks < 296 ? nblocks : 296, 256, 0, s>>>(x, d_blocks, nblocks);
    check_cuda(cudaGetLastError(), "set_identity_kernel launch");
}
}   // namespace xb

namespace xb {
// ---- SVD reverse rule (include/Xped/AD/ADTensor.hpp:253-325), elementwise parts ------------------------------------
__global__ void scale_by_diag_kernel(const double* __restrict__ src, double* __restrict__ dst, const double* __restrict__ S,
                                     const DiagScaleDesc* __restrict__ descs, int n, int accumulate)
{
    for(int b = blockIdx.y; b < n; b += gridDim.y) {
        const DiagScaleDesc d = descs[b];
        const long long cnt = static_cast<long long>(d.rows) * d.cols;
        const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
        for(long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < cnt; e += stride) {
            const int r = static_cast<int>(e % d.rows), c = static_cast<int>(e / d.rows);
            const int k = d.by_cols ? c : r;
            const double v = src[d.off + e];
            const double q = v == 0.0 ? 0.0 : v / S[d.s_off + k + static_cast<long long>(k) * d.s_ld];   // padded (zero) adjoints stay zero
            dst[d.off + e] = accumulate ? dst[d.off + e] + q : q;
        }
    }
}

void launch_scale_by_diag(const double* src, double* dst, const double* S, const DiagScaleDesc* d_desc, int n, int accumulate, cudaStream_t s)
{
    if(n == 0) return;
    dim3 grid(148, static_cast<unsigned>(n < 64 ? n : 64));
    scale_by_diag_kernel<<<grid, 256, 0, s>>>(src, dst, S, d_desc, n, accumulate);
    check_cuda(cudaGetLastError(), "scale_by_diag_kernel launch");
}

__global__ void svd_adjoint_core_kernel(const double* __restrict__ G1, const double* __restrict__ G2, const double* __restrict__ dS,
                                        const double* __restrict__ S, double* __restrict__ M, const BlockDesc* __restrict__ blocks, int nblocks)
{
    for(int b = blockIdx.y; b < nblocks; b += gridDim.y) {
        const BlockDesc d = blocks[b];
        const int r = d.rows;
        const long long cnt = static_cast<long long>(r) * r;
        const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
        for(long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < cnt; e += stride) {
            const int i = static_cast<int>(e % r), j = static_cast<int>(e / r);
            const long long ij = d.off + e, ji = d.off + j + static_cast<long long>(i) * r;
            const double si = S[d.off + i + static_cast<long long>(i) * r], sj = S[d.off + j + static_cast<long long>(j) * r];
            // vec_diff: F_inv(i,j) = v(j) - v(i) with v = sigma^2, as (sj - si)(sj + si): exactly 0 on the diagonal (sj * sj - si * si contracts into an FMA whose
            // rounding residue ~1e-16 sigma^2 exceeds the 1e-12 regularisation threshold for sigma^2 > 1e4 and then gives F_ii ~ 1e11)
            const double dij = (sj - si) * (sj + si);
            const double fij = i == j ? 0.0 : (fabs(dij) < 1.e-12 ? dij / (dij * dij + 1.e-12) : 1.0 / dij);
            const double g1ij = G1[ij], g1ji = G1[ji], g2ij = G2[ij], g2ji = G2[ji];
            // F is antisymmetric (also in its regularised form): (F o G + (F o G)^T)_ij = F_ij (G_ij - G_ji).
            double m = dS[ij] + fij * (g1ij - g1ji) * sj + si * fij * (g2ij - g2ji);
            if(g1ij != 0.0) m -= g1ij / sj;
            if(g2ji != 0.0) m -= g2ji / si;
            M[ij] = m;
        }
    }
}

void launch_svd_adjoint_core(const double* G1, const double* G2, const double* dS, const double* S, double* M, const BlockDesc* d_blocks, int nblocks,
                             cudaStream_t s)
{
    if(nblocks == 0) return;
    dim3 grid(148, static_cast<unsigned>(nblocks < 64 ? nblocks : 64));
    svd_adjoint_core_kernel<<<grid, 256, 0, s>>>(G1, G2, dS, S, M, d_blocks, nblocks);
    check_cuda(cudaGetLastError(), "svd_adjoint_core_kernel launch");
}
// ---- reverse rules of the element-wise ops (include/Xped/AD/ADTensor.hpp:511-524, 559-586) -------------------------------------------
// diag_sqrt / diag_inv / their composition: out = diag(g_ii * f'(x_ii)) expressed through the forward result y: sqrt: 0.5 / y, inv: -y^2, inv-sqrt: -0.5 y^3
__global__ void diag_map_backward_kernel(const double* __restrict__ g, const double* __restrict__ y, double* __restrict__ dst, const BlockDesc* __restrict__ blocks,
                                         int nblocks, int op)
{
    for(int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const BlockDesc d = blocks[b];
        const long long n = static_cast<long long>(d.rows) * d.cols;
        for(long long e = threadIdx.x; e < n; e += blockDim.x) {
            const int r = static_cast<int>(e % d.rows), c = static_cast<int>(e / d.rows);
            double v = 0.0;
            if(r == c) {
                const double yy = y[d.off + e];
                v = g[d.off + e] * (op == 0 ? 0.5 / yy : (op == 1 ? -yy * yy : -0.5 * yy * yy * yy));
            }
            dst[d.off + e] = v;
        }
    }
}

void launch_diag_map_backward(const double* g, const double* y, double* dst, const BlockDesc* d_blocks, int nblocks, int op, cudaStream_t s)
{
    if(nblocks == 0) return;
    diag_map_backward_kernel<<<nblocks < 296 ? nblocks : 296, 256, 0, s>>>(g, y, dst, d_blocks, nblocks, op);
    check_cuda(cudaGetLastError(), "diag_map_backward_kernel launch");
}

// <a, b> = sum_i a_i b_i in a fixed order: per-CTA partial sums, then one CTA adds them up (deterministic)
__global__ void dot_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ part)
{
    __shared__ double sm[256];
    const long long chunk = (n + gridDim.x - 1) / gridDim.x;
    const long long e0 = chunk * blockIdx.x, e1 = e0 + chunk < n ? e0 + chunk : n;
    double acc = 0.0;
    for(long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) acc = fma(a[e], b[e], acc);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for(int o = 128; o > 0; o >>= 1) {
        if(threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if(threadIdx.x == 0) part[blockIdx.x] = sm[0];
}
__global__ void dot_final_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out)
{
    if(threadIdx.x == 0) {
        double s = 0.0;
        for(int i = 0; i < nparts; ++i) s += part[i];
        *out = s;
    }
}

void launch_dot(const double* a, const double* b, int64_t n, double* d_part /* 129 doubles */, cudaStream_t s)
{
    dot_partial_kernel<<<128, 256, 0, s>>>(a, b, n, d_part + 1);
    dot_final_kernel<<<1, 32, 0, s>>>(d_part + 1, 128, d_part);
    check_cuda(cudaGetLastError(), "dot kernels launch");
}

// y = x / m, m = max |x|  =>  adj(x) = g / m - (<g, x> / m^2) sign(x_k) e_k at the position(s) k of the maximum (maxNorm rule, ADTensor.hpp:511-524, fused
// with the scaling the CTM applies right after it)
__global__ void maxnorm_scale_backward_kernel(const double* __restrict__ g, const double* __restrict__ x, const double* __restrict__ d_m, const double* __restrict__ d_dot,
                                              double* __restrict__ out, long long n)
{
    const double m = *d_m, inv = 1.0 / m, corr = *d_dot * inv * inv;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double xi = x[i];
        double v = g[i] * inv;
        if(fabs(xi) == m) v -= xi < 0.0 ? -corr : corr;
        out[i] = v;
    }
}

void launch_maxnorm_scale_backward(const double* g, const double* x, const double* d_m, const double* d_dot, double* out, int64_t n, cudaStream_t s)
{
    if(n == 0) return;
    long long gr = (n + 255) / 256;
    if(gr > 1184) gr = 1184;
    maxnorm_scale_backward_kernel<<<static_cast<int>(gr), 256, 0, s>>>(g, x, d_m, d_dot, out, n);
    check_cuda(cudaGetLastError(), "maxnorm_scale_backward_kernel launch");
}

}   // namespace xb
