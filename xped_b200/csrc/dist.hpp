// Peer-memory plumbing of the sector-sharded path (dist.cu): IPC-mapped device arenas, shared-memory mailbox, device-side barrier.
#pragma once
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

namespace xb {

class Context;

constexpr int DIST_MAX_WORLD = 8;    // the GPUs of one box
constexpr int DIST_SLOT_WORDS = 4;   // int64 words of one operation record

void dist_shm_create(const char* name, int world);   // rank 0, before any rank calls dist_init
void dist_init(Context& ctx, int rank, int world, size_t arena_bytes, const char* shm_name);
void dist_shutdown(Context& ctx);
int dist_rank(const Context& ctx);
int dist_world(const Context& ctx);
bool arena_alloc(Context& ctx, size_t bytes, double** p, int64_t* off);
void arena_free(Context& ctx, int64_t off);
void dist_arena_stats(const Context& ctx, int64_t out[3]);   // arena bytes, bytes in use, high-water mark
double* dist_peer_ptr(const Context& ctx, int peer, int64_t off);
void dist_exchange(Context& ctx, const int64_t mine[DIST_SLOT_WORDS], int64_t all[DIST_MAX_WORLD][DIST_SLOT_WORDS]);
extern int64_t dist_pull_mbps;
void dist_barrier(Context& ctx);
cudaStream_t dist_side_stream(Context& ctx, int i);   // one of world - 1 non-blocking side streams (created on first use)   // stream-ordered barrier over all ranks

}   // namespace xb
