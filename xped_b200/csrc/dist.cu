// Peer-memory plumbing of the sector-sharded hot path: one process per GPU, every rank owns a device ARENA that all
// other ranks of the box map through CUDA IPC, so the kernels of one rank store into (and load from) the shards of another over NVLink /
// NVSwitch -- the exchange step of a sharded permute is fused into the permute kernel itself (no send buffer, no receive buffer, no collective
// call), and a product can compute a row panel of a block that another rank owns and write the result panel into the owner's shard.
//
//   arena      cudaMalloc'ed once per rank; first 4 KiB = barrier flags; the rest is handed out by a host-side first-fit allocator (256-byte
//              granularity).  Tensors of a context with distributed state live in the arena (DevBuf tries it first), so every shard is
//              addressable by every peer as (mapped arena base of the owner) + (offset published by the owner).
//   mailbox    POSIX shared memory: IPC handles at start-up, then per collective operation one small record per rank (arena offsets of the
//              buffers the peers are about to touch).  Host-to-host, no device involvement; all ranks issue the same operations in the same order.
//   barrier    dist_barrier_kernel: every rank stores a monotone sequence number into its slot of every peer's flag array (st.release.sys over
//              NVLink) and spins on its own array (ld.acquire.sys) -- stream-ordered, a few microseconds, bounded spin (a dead peer traps).
//
// Protocol of a pushed permute on rank r:  allocate the output shard, publish its offset, wait for the peers' offsets (host) -> barrier A on the
// stream (every peer's output shard is allocated and idle) -> one permute launch per destination rank p with dst = peer p's shard (the tasks
// whose SOURCE sub-block r owns; fan-in 1, i.e. Abelian plans) -> barrier B (all peers have finished writing into r's shard).
// The reference's distributed backend (MPI/Cyclops, include/Xped/Util/Mpi.hpp) has no counterpart of this: it distributes dense blocks inside CTF.
#include <cuda_runtime.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>

#include "dist.hpp"
#include "engine.hpp"

namespace xb {

namespace {

constexpr uint64_t SHM_MAGIC = 0x78623264697374ull;   // "xb2dist"
constexpr int RING = 64;
constexpr size_t HEADER_BYTES = 4096;
constexpr size_t GRAIN = 256;

struct DistShm
{
    volatile uint64_t magic;
    volatile int world;
    cudaIpcMemHandle_t handle[DIST_MAX_WORLD];
    volatile uint64_t ready[DIST_MAX_WORLD];   // 1: handle published, 2: peers mapped, 3: shutting down
    volatile uint64_t seq[DIST_MAX_WORLD];     // number of operation records published by the rank
    volatile int64_t slot[DIST_MAX_WORLD][RING][DIST_SLOT_WORDS];
};

struct BarrierArgs
{
    unsigned long long* flags[DIST_MAX_WORLD];   // flag array of every rank (this process' mapping)
    int me, world;
    unsigned long long seq;
};

__global__ void dist_barrier_kernel(BarrierArgs a)
{
    const int p = threadIdx.x;
    if(p >= a.world || p == a.me) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(a.flags[p] + a.me), "l"(a.seq) : "memory");
    const unsigned long long* mine = a.flags[a.me] + p;
    const long long t0 = clock64();
    for(;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(mine) : "memory");
        if(v >= a.seq) break;
        if(clock64() - t0 > 30000000000ll) __trap();   // ~15 s: a peer died; fail the launch instead of hanging the GPU
    }
}

template <typename F> void spin_until(F&& cond, double seconds, const char* what)
{
    const auto t0 = std::chrono::steady_clock::now();
    int n = 0;
    while(!cond()) {
        if(++n > 2000) std::this_thread::yield();
        if(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds)
            throw std::runtime_error(std::string("xped_b200 distributed state: timed out waiting for ") + what);
    }
}

}   // namespace

struct DistState
{
    int rank = 0, world = 1;
    char* arena = nullptr;
    size_t arena_bytes = 0;
    char* peer[DIST_MAX_WORLD] = {};   // mapped arena base of every rank (peer[rank] == arena)
    DistShm* shm = nullptr;
    std::string shm_name;
    uint64_t opseq = 0;
    unsigned long long barrier_seq = 0;
    std::map<int64_t, int64_t> free_list;   // offset -> bytes
    std::map<int64_t, int64_t> used;
    int64_t high_water = 0;
    cudaStream_t side[DIST_MAX_WORLD] = {};
};

void dist_shm_create(const char* name, int world)
{
    if(world < 1 || world > DIST_MAX_WORLD) throw std::invalid_argument("dist_shm_create: world must be 1..8");
    shm_unlink(name);
    const int fd = shm_open(name, O_CREAT | O_RDWR | O_EXCL, 0600);
    if(fd < 0) throw std::runtime_error(std::string("shm_open(create) failed for ") + name);
    if(ftruncate(fd, sizeof(DistShm)) != 0) {
        close(fd);
        throw std::runtime_error("ftruncate on the shared mailbox failed");
    }
    void* p = mmap(nullptr, sizeof(DistShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if(p == MAP_FAILED) throw std::runtime_error("mmap of the shared mailbox failed");
    std::memset(p, 0, sizeof(DistShm));
    auto* s = static_cast<DistShm*>(p);
    s->world = world;
    __sync_synchronize();
    s->magic = SHM_MAGIC;
    munmap(p, sizeof(DistShm));
}

void dist_init(Context& ctx, int rank, int world, size_t arena_bytes, const char* shm_name)
{
    if(ctx.planning_only()) throw std::runtime_error("xped_b200: no CUDA device -- distributed state needs one");
    if(ctx.dist) throw std::invalid_argument("dist_init: context already has distributed state");
    if(world < 1 || world > DIST_MAX_WORLD || rank < 0 || rank >= world) throw std::invalid_argument("dist_init: bad rank / world");
    auto D = std::make_unique<DistState>();
    D->rank = rank;
    D->world = world;
    D->shm_name = shm_name;
    const int fd = shm_open(shm_name, O_RDWR, 0600);
    if(fd < 0) throw std::runtime_error(std::string("shm_open failed for ") + shm_name + " (rank 0 creates it with xb_dist_shm_create before the others open it)");
    void* p = mmap(nullptr, sizeof(DistShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if(p == MAP_FAILED) throw std::runtime_error("mmap of the shared mailbox failed");
    D->shm = static_cast<DistShm*>(p);
    if(D->shm->magic != SHM_MAGIC || D->shm->world != world) {
        munmap(p, sizeof(DistShm));
        throw std::runtime_error("shared mailbox was not created for this job");
    }
    check_cuda(cudaSetDevice(ctx.device), "cudaSetDevice");
    arena_bytes = (arena_bytes + GRAIN - 1) / GRAIN * GRAIN;
    if(arena_bytes < HEADER_BYTES + GRAIN) throw std::invalid_argument("dist_init: arena too small");
    check_cuda(cudaMalloc(reinterpret_cast<void**>(&D->arena), arena_bytes), "cudaMalloc(peer arena)");
    D->arena_bytes = arena_bytes;
    check_cuda(cudaMemset(D->arena, 0, HEADER_BYTES), "cudaMemset(arena header)");
    check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    D->free_list[static_cast<int64_t>(HEADER_BYTES)] = static_cast<int64_t>(arena_bytes - HEADER_BYTES);
    D->peer[rank] = D->arena;
    if(world > 1) {
        cudaIpcMemHandle_t h;
        check_cuda(cudaIpcGetMemHandle(&h, D->arena), "cudaIpcGetMemHandle");
        std::memcpy(const_cast<cudaIpcMemHandle_t*>(&D->shm->handle[rank]), &h, sizeof(h));
        __sync_synchronize();
        D->shm->ready[rank] = 1;
        DistShm* s = D->shm;
        spin_until([&] {
            for(int q = 0; q < world; ++q)
                if(s->ready[q] < 1) return false;
            return true;
        }, 120.0, "the peers' IPC handles");
        __sync_synchronize();
        for(int q = 0; q < world; ++q) {
            if(q == rank) continue;
            cudaIpcMemHandle_t hq;
            std::memcpy(&hq, const_cast<cudaIpcMemHandle_t*>(&s->handle[q]), sizeof(hq));
            void* mapped = nullptr;
            check_cuda(cudaIpcOpenMemHandle(&mapped, hq, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(peer arena)");
            D->peer[q] = static_cast<char*>(mapped);
        }
        s->ready[rank] = 2;
        spin_until([&] {
            for(int q = 0; q < world; ++q)
                if(s->ready[q] < 2) return false;
            return true;
        }, 120.0, "the peers to map the arenas");
    }
    ctx.dist = D.release();
}

void dist_shutdown(Context& ctx)
{
    DistState* D = ctx.dist;
    if(!D) return;
    ctx.dist = nullptr;
    cudaSetDevice(ctx.device);
    cudaDeviceSynchronize();
    if(D->world > 1) {
        // nobody unmaps / frees while a peer may still be storing into this arena: wait (bounded) until every rank is past its last launch
        D->shm->ready[D->rank] = 3;
        try {
            DistShm* s = D->shm;
            const int world = D->world;
            spin_until([&] {
                for(int q = 0; q < world; ++q)
                    if(s->ready[q] < 3) return false;
                return true;
            }, 20.0, "the peers to shut down");
        } catch(...) {
        }
        for(int q = 0; q < D->world; ++q)
            if(q != D->rank && D->peer[q]) cudaIpcCloseMemHandle(D->peer[q]);
    }
    for(cudaStream_t st : D->side)
        if(st) cudaStreamDestroy(st);
    if(D->arena) cudaFree(D->arena);
    if(D->shm) munmap(D->shm, sizeof(DistShm));
    if(D->rank == 0) shm_unlink(D->shm_name.c_str());
    delete D;
}

int dist_rank(const Context& ctx) { return ctx.dist ? ctx.dist->rank : 0; }
int dist_world(const Context& ctx) { return ctx.dist ? ctx.dist->world : 1; }

bool arena_alloc(Context& ctx, size_t bytes, double** p, int64_t* off)
{
    DistState* D = ctx.dist;
    if(!D) return false;
    const int64_t need = static_cast<int64_t>((std::max<size_t>(bytes, 1) + GRAIN - 1) / GRAIN * GRAIN);
    for(auto it = D->free_list.begin(); it != D->free_list.end(); ++it) {
        if(it->second < need) continue;
        const int64_t o = it->first, rest = it->second - need;
        D->free_list.erase(it);
        if(rest > 0) D->free_list[o + need] = rest;
        D->used[o] = need;
        D->high_water = std::max(D->high_water, o + need);
        *p = reinterpret_cast<double*>(D->arena + o);
        *off = o;
        return true;
    }
    return false;
}

void arena_free(Context& ctx, int64_t off)
{
    DistState* D = ctx.dist;
    if(!D) return;   // the arena is already gone (context shutting down)
    auto u = D->used.find(off);
    if(u == D->used.end()) return;
    int64_t o = off, n = u->second;
    D->used.erase(u);
    auto next = D->free_list.lower_bound(o);
    if(next != D->free_list.end() && o + n == next->first) {
        n += next->second;
        next = D->free_list.erase(next);
    }
    if(next != D->free_list.begin()) {
        auto prev = std::prev(next);
        if(prev->first + prev->second == o) {
            prev->second += n;
            return;
        }
    }
    D->free_list[o] = n;
}

void dist_arena_stats(const Context& ctx, int64_t out[3])
{
    out[0] = out[1] = out[2] = 0;
    if(!ctx.dist) return;
    out[0] = static_cast<int64_t>(ctx.dist->arena_bytes);
    for(const auto& kv : ctx.dist->used) out[1] += kv.second;
    out[2] = ctx.dist->high_water;
}

double* dist_peer_ptr(const Context& ctx, int peer, int64_t off)
{
    DistState* D = ctx.dist;
    if(!D || peer < 0 || peer >= D->world) throw std::invalid_argument("dist_peer_ptr: bad peer");
    return reinterpret_cast<double*>(D->peer[peer] + off);
}

// every rank publishes `mine` for the current operation and receives the records of all ranks (host side, shared memory)
void dist_exchange(Context& ctx, const int64_t mine[DIST_SLOT_WORDS], int64_t all[DIST_MAX_WORLD][DIST_SLOT_WORDS])
{
    DistState* D = ctx.dist;
    if(!D) throw std::invalid_argument("dist_exchange: context has no distributed state");
    DistShm* s = D->shm;
    const uint64_t k = D->opseq;
    for(int w = 0; w < DIST_SLOT_WORDS; ++w) s->slot[D->rank][k % RING][w] = mine[w];
    __sync_synchronize();
    s->seq[D->rank] = k + 1;
    for(int q = 0; q < D->world; ++q) {
        if(q != D->rank) spin_until([&] { return s->seq[q] >= k + 1; }, 120.0, "a peer to publish its operation record (ranks must issue the same sharded operations in the same order)");
        __sync_synchronize();
        for(int w = 0; w < DIST_SLOT_WORDS; ++w) all[q][w] = s->slot[q][k % RING][w];
    }
    D->opseq = k + 1;
}

cudaStream_t dist_side_stream(Context& ctx, int i)
{
    DistState* D = ctx.dist;
    if(!D || i < 0 || i >= DIST_MAX_WORLD) throw std::invalid_argument("dist_side_stream: bad index");
    if(!D->side[i]) check_cuda(cudaStreamCreateWithFlags(&D->side[i], cudaStreamNonBlocking), "cudaStreamCreate(side)");
    return D->side[i];
}

void dist_barrier(Context& ctx)
{
    DistState* D = ctx.dist;
    if(!D || D->world == 1) return;
    BarrierArgs a;
    std::memset(&a, 0, sizeof(a));
    for(int q = 0; q < D->world; ++q) a.flags[q] = reinterpret_cast<unsigned long long*>(D->peer[q]);
    a.me = D->rank;
    a.world = D->world;
    a.seq = ++D->barrier_seq;
    dist_barrier_kernel<<<1, 32, 0, ctx.stream>>>(a);
    check_cuda(cudaGetLastError(), "dist_barrier_kernel launch");
    ++ctx.launches;
}

}   // namespace xb
