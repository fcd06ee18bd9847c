// xped_b200 engine: device-resident block arenas, cached layouts and cached permute plans.
//
// A Tensor is (shared immutable Layout, one contiguous device buffer).  The buffer holds exactly the
// reference's contiguous storage image: per-sector column-major matrices concatenated at
// Layout::off[i] (include/Xped/Core/storage/StorageType_contiguous.hpp:79-97,112-122), so upload /
// download are plain memcpys of `m_data` and structure checks are bit-exact by construction.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "core.hpp"

namespace xb {

struct Layout
{
    std::vector<BasisPtr> dom, cod;   // uncoupled leg bases
    BasisPtr cdom, ccod;              // coupled bases with fusion trees
    std::vector<Q> sector;
    std::unordered_map<Q, int, QHash> dict;
    std::vector<int64_t> off, rows, cols;
    int64_t size = 0;
    std::string sig;   // leg-basis signature (+ explicit sector order for push_back-built tensors)
    int rank() const { return static_cast<int>(dom.size()); }
    int corank() const { return static_cast<int>(cod.size()); }
    int find(const Q& q) const
    {
        auto it = dict.find(q);
        return it == dict.end() ? -1 : it->second;
    }
};
using LayoutPtr = std::shared_ptr<const Layout>;

// ---- device-side plan records (plain structs shared with the kernels) -------------------------------
constexpr int PERM_MAXLEG = 8;
constexpr int PERM_TILE_ELEMS = 2048;   // doubles per tile (16 KiB of payload)
constexpr int PERM_BUF_ELEMS = 3072;    // shared buffer incl. bank-conflict padding
constexpr int PERM_THREADS = 256;
constexpr int PERM_STRIP = 8;   // longest per-thread strip

constexpr int PERM_PEER_SHIFT = 56;   // dst_off = (owner rank of the destination shard << 56) | element offset inside that shard
constexpr long long PERM_OFF_MASK = (1ll << PERM_PEER_SHIFT) - 1;
struct PeerDelta { long long d[8]; };   // element offset of every rank's destination shard relative to the launch's `dst` pointer

struct PermTask   // one destination sub-block (target tree pair) = sum_k coeff_k * transpose(src_k)
{
    long long dst_off;
    int nleg;        // merged legs, dst order (leg 0 fastest in dst)
    int ntl;         // legs with tile extent > 1
    int src_begin;   // first PermSrc
    int src_count;   // fan-in K (0 => zero fill)
    int vol;         // tile volume prod(w)
    int pad_;
    int dims[PERM_MAXLEG];
    int w[PERM_MAXLEG];
    int ntile[PERM_MAXLEG];
    int dstride[PERM_MAXLEG];
    int smstride[PERM_MAXLEG];       // shared-memory stride of each leg (src-order packing, padded)
    int tl_d[PERM_MAXLEG];           // tile legs in dst order
    int tl_s[PERM_MAXLEG];           // tile legs in src order
    unsigned magic[PERM_MAXLEG];     // ceil(2^32 / w[leg]) for fast division by the tile extent
    // Thread work items of the two phases.  An item is (x along the fastest tile leg of that side, a strip of <= L
    // consecutive indices of the second-fastest tile leg, the remaining tile legs): lanes run along the contiguous
    // leg, every thread walks its strip with constant address increments, so the mixed-radix index decomposition is
    // paid once per strip instead of once per element.
    int strip_s, nchunk_s, items_s;   // read phase (source order)
    int strip_d, nchunk_d, items_d;   // write phase (destination order)
    unsigned magic_chunk_s, magic_chunk_d;
    // first source inline: Abelian symmetries have fan-in 1, so most tasks never touch the PermSrc table
    long long src0_off;
    double src0_coeff;
    int src0_stride[PERM_MAXLEG];
};

struct PermSrc
{
    long long off;
    double coeff;
    int sstride[PERM_MAXLEG];   // source stride of each (dst-ordered) leg
};

struct PermTile { int task; unsigned short c[PERM_MAXLEG]; };   // task + tile coordinate along each merged leg (no divisions on the device)

// ---- bulk-copy (TMA) form of a permute plan: used when every contiguous run of every task is 16-byte aligned and an even
// number of doubles long (cp.async.bulk needs 16-byte addresses and sizes) and the fan-in is 1 (Abelian symmetries).
// A tile is wx (destination-contiguous leg 0) x wy (source-contiguous leg `ls`) x nz (all other legs).  Source runs
// (wy doubles each, one per (x, z)) are fetched by cp.async.bulk into shared memory, transposed there (a batch of nz
// wx x wy transposes with the coefficient applied), and the destination runs (wx doubles, one per (y, z)) leave through
// cp.async.bulk stores.  ls == 0 (same fastest leg on both sides): runs are copied through shared memory unchanged.
constexpr int BULK_TILE_ELEMS = 2304;
constexpr int BULK_BUF_ELEMS = 2816;   // one staging buffer incl. run padding
constexpr int BULK_THREADS = 256;
struct BulkTask
{
    long long dst_off, src_off;
    double coeff;
    int nleg, ls;             // ls: source-contiguous leg (0 => copy mode)
    int wx, wy, nz;           // tile extents: leg 0, leg ls, product over the other legs
    int rs_s, rs_d;           // shared-memory pitch of a source / destination run (even)
    int nxh, nyh;             // ceil(wx / gx), ceil(wy / gy): item grid of the transposing pass (lane grid gx x gy = 2 x 16 or 4 x 8)
    int gx_shift;             // log2(gx)
    int nzc, gp;              // transpose mode: extent of the first z leg when its source runs continue each other (else 1) and the
                              // staging pitch of one (z_rest, x) group of nzc source runs
    unsigned magic_nzc;
    int nzleg;                // number of z legs
    int mrg_load, mrg_store;  // > 1: that many consecutive runs are contiguous in global memory and in the staging buffer (one bulk operation)
    int dims[PERM_MAXLEG], w[PERM_MAXLEG];
    int dstride[PERM_MAXLEG], sstride[PERM_MAXLEG];
    int zleg[PERM_MAXLEG];    // the z legs (dst order)
    unsigned zmagic[PERM_MAXLEG];   // ceil(2^32 / w[zleg[i]])
    unsigned magic_wx, magic_wy, magic_nxh, magic_nyh;
};

// ---- table-driven form of a permute plan (SU(2) recoupling sums, odd extents: everything the bulk form cannot take).  Tiles are full boxes (tile extents
// divide the leg extents); for every distinct tile shape (extents + strides of the destination and of all sources) the host tabulates, per tile element,
// the source offset relative to the tile origin for every source, and, in destination order, the destination offset and the staging position.  The kernel
// does no index arithmetic: a thread reads (table, data), accumulates its <= 8 elements over all sources in registers, stages them once, and the second
// phase writes them out in destination order.
constexpr int LUT_TILE_ELEMS = 2048;
constexpr int LUT_THREADS = 256;
struct LutTask
{
    int vol, src_count;       // elements per tile, fan-in
    int soff_tab;             // int table: source k of tile element e at soff_tab + k * vol + e (e in source-0 order = staging order)
    int doff_tab;             // int table: destination offset of element e' (destination order)
    int spos_tab;             // unsigned short table: staging position of element e'
    int first_tile;           // index of the task's first tile
    int tsrc_begin;           // LutTileSrc records of the task's tiles: tsrc_begin + (tile - first_tile) * src_count + k
    int pad_;
};
struct LutTileSrc { long long off; double coeff; };   // origin of one tile inside one source, and that source's recoupling coefficient

// Host copies of a plan's tables.  Kept only in planning-only contexts (no GPU), where the CPU tests walk them element by element the way the kernels do
// (Context::emulate_permute_plan): the plan builder is the most intricate host logic of the path and can be checked against the oracle without a device.
struct PlanHostTables
{
    std::vector<PermTask> tasks;
    std::vector<PermSrc> srcs;
    std::vector<PermTile> tiles;
    std::vector<BulkTask> btasks;   // bulk form (empty if the plan does not qualify)
    std::vector<PermTile> btiles;
    std::vector<LutTask> ltasks;
    std::vector<LutTileSrc> ltsrc;
    std::vector<int> ltile_task;
    std::vector<long long> ltile_dst;
    std::vector<int> lut_int;
    std::vector<unsigned short> lut_pos;
};

struct PermutePlan
{
    LayoutPtr out;
    std::shared_ptr<PlanHostTables> host;   // planning-only contexts
    int n_tasks = 0, n_srcs = 0, n_tiles = 0;
    int tiles_per_cta = 1;
    PermTask* d_tasks = nullptr;
    PermSrc* d_srcs = nullptr;
    PermTile* d_tiles = nullptr;
    // bulk-copy form (nullptr if the plan does not qualify)
    BulkTask* d_btasks = nullptr;
    PermTile* d_btiles = nullptr;
    int2* d_branges = nullptr;   // per CTA: (first tile, tile count), all tiles of one task
    int n_btiles = 0, n_branges = 0;
    // table-driven form (nullptr if the plan does not qualify or has the bulk form)
    LutTask* d_ltasks = nullptr;
    LutTileSrc* d_ltsrc = nullptr;      // per (tile, source)
    int* d_ltile_task = nullptr;        // per tile
    long long* d_ltile_dst = nullptr;   // per tile: origin in the destination
    int* d_lut_int = nullptr;
    unsigned short* d_lut_pos = nullptr;
    int n_ltiles = 0, ltiles_per_cta = 1;
    long long lut_table_bytes = 0;
    int uses = 0;              // cache hits (the table-driven form is built when a plan is used for the second time: ~30-50 ms of host work per plan)
    bool lut_tried = false;
    bool identity_copy = false;   // permute<0, identity>: plain copy (Tensor.cpp:344)
    int64_t bytes = 0;            // algorithmic bytes: 8*(N_src + N_dst)
    int max_fan_in = 0;
    int64_t dst_elems = 0, src_reads = 0;   // census: destination elements, source elements read (fan-in counted)
    ~PermutePlan();
};

// Sector-sharded permute: every rank owns whole sectors; a permute has ONE exchange step.
//   pack    owned source sub-blocks that a peer's destination tasks need -> dense segments of the send buffer (one launch)
//   (NCCL alltoallv of the send buffer, issued by the host framework: torch.distributed.all_to_all_single)
//   local   destination tasks whose sources are all owned by this rank (overlaps the exchange)
//   remote  destination tasks with at least one received source; source offsets >= `split` address the receive buffer
struct DistPermutePlan
{
    int rank = 0, world = 1;
    LayoutPtr in_global, out_global, in_local, out_local;
    std::vector<int> owner_in, owner_out;   // per sector of the global layouts
    std::vector<int64_t> send_counts, recv_counts, send_displs, recv_displs;   // doubles, per peer
    int64_t send_total = 0, recv_total = 0;
    int64_t split = 0;   // = in_local->size
    PermutePlan pack, local, remote;
    // direct-write form (every destination sub-block has at most one source: Abelian plans): push[p] = the tasks whose SOURCE this rank owns and whose
    // destination rank p owns, lowered against p's output shard -- launched with dst = p's shard (peer-mapped); no buffers, no collective (dist.cu)
    std::vector<std::unique_ptr<PermutePlan>> push;
    bool push_zero_fill = false;   // a destination sub-block of this rank has no source at all
    int64_t bytes_local = 0;   // algorithmic bytes moved through this rank's HBM by local + remote: 8 * (N_in_local + N_out_local)
    // host copies of the three work lists for introspection and the CPU (gloo) tests: flat int64 records
    static constexpr int MAXL = 16;
    static constexpr int PACK_REC = 4 + 2 * MAXL;   // peer, send_off, src_off, nleg, dims[16], sstride[16]
    static constexpr int TASK_REC = 4 + 2 * MAXL;   // dst_off, nleg, src_begin, src_count, dims[16], dstride[16]
    static constexpr int SRC_REC = 4 + MAXL;        // remote flag, off (local tensor or receive buffer), coeff bits, 0, sstride[16]
    std::vector<int64_t> pack_records, task_records, src_records;
};

// one row panel C_q[row0 : row0 + rows, :] of a sharded product, computed by `rank` (dist_ops.cpp)
struct DistPanel { int sector; int rank; int64_t row0, rows; };
std::vector<DistPanel> dist_multiply_assignment(const std::vector<int64_t>& m, const std::vector<int64_t>& n, const std::vector<int64_t>& k, const std::vector<int>& owner,
                                                int world, int panel_rows_min, double imbalance_tol);

constexpr int GEMM_TILE_N = 64;   // CTA tile of gemm_grouped_kernel is tile_m x 64 with tile_m = 128 (large sectors) or 64

struct GemmItem   // one (sector, tile) of a grouped product
{
    long long a_off, b_off, c_off;
    int m, n, k;
    int lda, ldb, ldc;
    int tm, tn;   // tile origin
    int ta, tb;   // operand stored transposed (adjoint block)?
    int acc;      // 1: C += alpha * op(A) op(B)   0: C = alpha * op(A) op(B)
    int pad_;
    double alpha;
};

// one split-K problem: C[c_off + i + j ldc] = sum_s scratch[s_off + s m n + i + j m]
struct SplitKJob { long long c_off, s_off; int m, n, ldc, split; };

// host lists of a grouped product (ops.cpp: build_gemm_lists)
struct GemmLists
{
    std::vector<GemmItem> items[2];   // [0]: 128 x 64 tiles, [1]: 64 x 64 tiles
    std::vector<SplitKJob> jobs;
    long long scratch_elems = 0;
    double flops = 0.0;
};
GemmLists build_gemm_lists(const std::vector<GemmItem>& problems, int sm_count);
struct ContractPerms { std::vector<int> p1, p2, pres; int shift1 = 0, shift2 = 0, shiftres = 0; };
ContractPerms contract_permutations(int rank_a, int corank_a, const std::vector<int>& l1, int rank_b, int corank_b, const std::vector<int>& l2, int res_rank);
void emulate_gemm_lists(const GemmLists& G, const double* A, int64_t n_a, const double* B, int64_t n_b, double* C, int64_t n_c);   // test support
std::vector<GemmItem> multiply_problems(const Layout& LA, bool adj_a, const Layout& LB, bool adj_b, const Layout& LO);

// a grouped product whose tile lists are already on the device: launched any number of times (iterative solvers, graph capture)
struct PreparedGemm
{
    const double *A = nullptr, *B = nullptr;
    double* C = nullptr;
    GemmItem* d_items[2] = {nullptr, nullptr};
    int n_items[2] = {0, 0};
    SplitKJob* d_jobs = nullptr;
    int n_jobs = 0, max_job_elems = 0;   // split-K problems and the element count of the largest one (grid of the reduce launch)
    double* d_scratch = nullptr;
    int launches = 0;
    double flops = 0.0;
};

class Context;
class Tape;
struct DistState;   // peer arenas, mailbox, barrier (dist.cu)
struct RsvdPlan;   // cached state of the truncated tSVD for one layout (rsvd.cu)

// device storage of a tensor; shared between a Tensor and its aliases (the AD tape keeps values alive through aliases)
struct DevBuf
{
    Context* ctx;
    double* p = nullptr;
    int64_t arena_off = -1;   // >= 0: lives in the peer-mapped arena of the context (dist.cu) at this byte offset
    DevBuf(Context* c, size_t bytes, bool zero);
    ~DevBuf();
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct Tensor
{
    Context* ctx = nullptr;
    LayoutPtr layout;
    std::shared_ptr<DevBuf> buf;
    double* d = nullptr;
    Tensor(Context* c, LayoutPtr l, bool zero);
    Tensor(const Tensor&) = delete;
    Tensor& operator=(const Tensor&) = delete;
    int64_t size() const { return layout->size; }
    Tensor* alias() const;   // second handle on the same storage (no copy)
    const DevBuf* id() const { return buf.get(); }
private:
    Tensor() = default;
};

struct SvdResult
{
    Tensor *U = nullptr, *S = nullptr, *Vd = nullptr;
    double trunc_weight = 0.0, entropy = 0.0;
};

class Context
{
public:
    Context(int device, const std::vector<Factor>& factors);
    ~Context();

    SymPtr sym;
    int device;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string last_error;

    // ---- host metadata (cached) ----
    BasisPtr coupled(const std::vector<BasisPtr>& uncoupled);
    LayoutPtr layout(const std::vector<BasisPtr>& dom, const std::vector<BasisPtr>& cod);
    LayoutPtr layout_custom(const std::vector<BasisPtr>& dom, const std::vector<BasisPtr>& cod, const std::vector<Q>& sectors);
    std::shared_ptr<PermutePlan> permute_plan(const LayoutPtr& in, int shift, const std::vector<int>& pi, bool adjoint_plan);
    // test support: applies the HOST tables of a plan of a planning-only context to host arrays (form 0: tile list of the LSU form, 1: table-driven form, 2: cp.async.bulk form)
    void emulate_permute_plan(const PermutePlan& plan, int form, const double* src, int64_t n_src, double* dst, int64_t n_dst) const;
    // sector-sharded permute: exchange plan of rank `rank` for the global tensor layout `in` (cached)
    std::shared_ptr<DistPermutePlan> dist_permute_plan(const LayoutPtr& in, int shift, const std::vector<int>& pi, bool adjoint_plan,
                                                       const std::vector<int>& owner_in, const std::vector<int>& owner_out, int rank, int world);
    void dist_pack(const DistPermutePlan& plan, const Tensor& local_in, double* d_send);
    Tensor* dist_local(const DistPermutePlan& plan, const Tensor& local_in);
    void dist_remote(const DistPermutePlan& plan, const Tensor& local_in, const double* d_recv, Tensor& local_out);
    // peer-memory form of the exchange (dist.cu / dist_ops.cpp): needs distributed state (xb_dist_init)
    DistState* dist = nullptr;
    bool arena_force = false;   // the next tensor allocations must come from the arena (buffers that peers address)
    Tensor* dist_permute_push(const DistPermutePlan& plan, const Tensor& local_in);
    // operator* on shards with row-panel splitting of the heavy sectors; owner maps over the GLOBAL sector lists of the operands (common sectors must agree)
    Tensor* dist_multiply(const Tensor& a, const std::vector<int>& owner_a, const Tensor& b, const std::vector<int>& owner_b, int panel_rows_min, std::vector<int>* owner_out);
    Tensor* multiply_sectors(const Tensor& a, bool adj_a, const Tensor& b, bool adj_b, const std::vector<Q>* out_sectors);
    bool planning_only() const { return device < 0; }

    // ---- ops (all asynchronous on `stream` unless they return host values) ----
    Tensor* permute(const Tensor& t, int shift, const std::vector<int>& pi);
    Tensor* permute_adj(const Tensor& t, int shift, const std::vector<int>& pi);
    Tensor* multiply(const Tensor& a, bool adj_a, const Tensor& b, bool adj_b);   // operator* with lazy AdjointOp operands
    Tensor* contract(const Tensor& a, const std::vector<int>& l1, const Tensor& b, const std::vector<int>& l2, int res_rank);
    Tensor* adjoint(const Tensor& t);
    Tensor* twist(const Tensor& t, int leg);
    Tensor* twist_impl(const Tensor& t, int leg);
    Tensor* copy(const Tensor& t);
    Tensor* axpby(double alpha, const Tensor& x, double beta, const Tensor* y);   // alpha*x + beta*y
    Tensor* diag_map(const Tensor& t, int op);                                    // 0 sqrt, 1 inv, 2 inv-sqrt
    double max_norm(const Tensor& t);
    Tensor* scale_by_inv_maxnorm(const Tensor& t);   // t * (1 / t.maxNorm()), scalar stays on the device
    Tensor* scale_by_inv_contraction(const Tensor& x, const Tensor& w);   // x / trace(x * w)  (density-matrix normalisation, CTM.cpp:660-667)
    Tensor* weighted_identity(const LayoutPtr& L);                        // Tensor::WeightedIdentity: degeneracy(q) * 1 per sector (reverse rule of trace)
    double unweighted_trace(const Tensor& t);                             // sum_q tr(block_q)  (ADTensor.hpp:698)
    // ---- reverse-mode AD (include/Xped/AD/ADTensor.hpp): while a tape is installed and recording, the ops above record their reverse rule
    Tape* tape = nullptr;
    bool taping() const;
    double trace(const Tensor& t);
    double squared_norm(const Tensor& t);
    void fill(Tensor& t, int mode, double value, uint64_t seed);                 // 0 constant, 1 identity, 2 counter-based uniform[-1,1]
    SvdResult tsvd(const Tensor& t, int64_t max_keep, double eps_svd, bool preserve_multiplets);
    // the two algorithms behind tsvd(): full batched Jacobi on every sector block, and (few kept values out of many) randomized block subspace
    // iteration + Jacobi on the projected blocks with an a-posteriori residual check (returns false => caller falls back to the full path)
    SvdResult tsvd_full(const Tensor& t, int64_t max_keep, double eps_svd, bool preserve_multiplets);
    bool tsvd_truncated(const Tensor& t, int64_t max_keep, double eps_svd, bool preserve_multiplets, SvdResult* out);
    int64_t svd_truncated_calls = 0, svd_truncated_fallbacks = 0, svd_truncated_iters = 0;
    int svd_call_tag = 0;   // call-site tag of the next tSVD calls ("svd.call_tag"): the truncated path keeps one plan + warm-start subspace per (layout, tag)
    int opt_svd_truncated = -1;   // -1: XB_SVD_TRUNC environment default (on), 0 / 1: forced (xb_ctx_set_int "svd.truncated")
    std::unordered_map<std::string, std::shared_ptr<RsvdPlan>> rsvd_plans;
    void clear_rsvd_plans();
    Tensor* tsvd_backward(const Tensor& A, const Tensor& U, const Tensor& S, const Tensor& Vd, const Tensor* dU, const Tensor* dS, const Tensor* dVd);
    void tqr(const Tensor& t, bool return_lq, Tensor** Qo, Tensor** Ro);
    void teigh(const Tensor& t, Tensor** D, Tensor** V);
    void contract_L(const Tensor& Bold, const Tensor& Bra, const Tensor& Ket, Tensor** Bnew);
    void contract_R(const Tensor& Bold, const Tensor& Bra, const Tensor& Ket, Tensor** Bnew);
    // Mps::leftSweepStep / rightSweepStep with DMRG::BROOM::SVD (src/MPS/Mps.cpp:181-229); `neighbour` may be null (chain end / DISCARD)
    void mps_left_sweep_step(const Tensor& Aloc, const Tensor* Aprev, int64_t max_nsv, double eps_svd, Tensor** Aloc_new, Tensor** Aprev_new, double* trunc_weight,
                             double* entropy);
    void mps_right_sweep_step(const Tensor& Aloc, const Tensor* Anext, int64_t max_nsv, double eps_svd, Tensor** Aloc_new, Tensor** Anext_new, double* trunc_weight,
                              double* entropy);
    void sync();
    // grouped product launcher: one GemmItem per problem (tm = tn = 0); tiles, sorts, uploads and launches
    void run_gemm(const double* A, const double* B, double* C, const std::vector<GemmItem>& problems);
    PreparedGemm prepare_gemm(const double* A, const double* B, double* C, const std::vector<GemmItem>& problems);
    void launch_gemm(const PreparedGemm& P);
    void free_gemm(PreparedGemm& P);
    // operator* keeps its prepared tile lists per (operand layouts, adjoint flags): a CTM move repeats the same ~90 products every step, and building, sorting
    // and uploading the lists per call put ~1.5 ms of copies and allocations on the stream of a 20 ms step.  Offsets are relative to the operand base pointers;
    // only split-K scratch offsets are relative to the RESULT pointer: an entry with scratch is rebuilt when the pool hands out another result address.
    struct GemmCacheEntry { PreparedGemm P; const double* c_ptr = nullptr; };
    std::unordered_map<std::string, GemmCacheEntry> gemm_cache;
    int64_t gemm_cache_hits = 0, gemm_cache_misses = 0;

    // one Jacobi sweep (flag reset + all rounds) as an instantiated CUDA graph, keyed by layout; valid while the stream-ordered pool
    // hands the same work buffers back (it does in a steady CTM run), re-captured otherwise
    struct SweepGraph { cudaGraphExec_t exec = nullptr; const void *w = nullptr, *v = nullptr, *meta = nullptr, *x = nullptr; int rounds = 0; };
    std::unordered_map<std::string, SweepGraph> sweep_graphs;

    // statistics for bench / gpu_launches
    int64_t launches = 0;
    int64_t plan_hits = 0, plan_misses = 0;
    int64_t host_table_bytes = 0;   // planning-only contexts: bytes of the host copies kept with the cached plans (bounded: 1 GB, then plans go without)

    // ---- live per-category device timing (bench.py roofline): CUDA event pairs on `stream` around each
    // launch group while profiling is enabled; prof_read() synchronises, sums and resets.
    enum Cat { CAT_PERMUTE = 0, CAT_GEMM = 1, CAT_SVD = 2, CAT_OTHER = 3, CAT_N = 4 };
    bool profiling = false;
    struct ProfRec { cudaEvent_t e0, e1; int cat; double work; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t get_event();
    void prof_read(double ms[CAT_N], int64_t count[CAT_N], double work[CAT_N]);
    cudaEvent_t timer_e0 = nullptr, timer_e1 = nullptr;
    void timer_start();
    double timer_stop();   // ms between timer_start() and now on `stream` (synchronises)

    // scratch device buffer for reductions / work lists
    void* scratch(size_t bytes);
    void* pinned(size_t bytes);
    void upload_async(void* dst, const void* src, size_t bytes);   // via a growing pinned staging ring

private:
    std::unordered_map<std::string, BasisPtr> coupled_cache_;
    std::unordered_map<std::string, LayoutPtr> layout_cache_;
    std::unordered_map<std::string, std::shared_ptr<PermutePlan>> plan_cache_;
public:
    // recoupling results of treepair::permute keyed by the sector LABELS of the tree pair (+ shift, permutation, dual flags): they do not
    // depend on the multiplicities, so plans for layouts that only differ in sector dimensions (a CTM run whose chi bases still move)
    // reuse them; `dims_follow_permutation` records that the multiplicities of the result trees are those of the argument legs p[k]
    struct RecouplingEntry { std::vector<std::pair<TreePair, double>> items; bool dims_follow_permutation = false; };
    std::unordered_map<std::string, RecouplingEntry> recoupling_cache;
private:
    std::unordered_map<std::string, std::shared_ptr<DistPermutePlan>> dist_plan_cache_;
    void* scratch_ = nullptr;
    size_t scratch_bytes_ = 0;
    void* pinned_ = nullptr;
    size_t pinned_bytes_ = 0;
    // pinned staging ring for small metadata uploads (work lists, block descriptors): chunks are reused once the
    // copies that read them have completed (one fence event per chunk)
    struct StageChunk { char* p = nullptr; size_t cap = 0, used = 0; cudaEvent_t fence = nullptr; bool pending = false; };
    std::vector<StageChunk> staging_;
    size_t stage_cur_ = 0;
};

void check_cuda(cudaError_t e, const char* what);
// cudaFuncSetAttribute is per DEVICE: true the first time the calling site runs on the current device (one bit per device ordinal in `mask`)
inline bool first_use_on_current_device(unsigned long long& mask)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    const bool first = (mask & bit) == 0;
    mask |= bit;
    return first;
}

// ---- state interop (io.cpp)
class Ctm;
void save_tensor(Context& ctx, const Tensor& t, const std::string& path);
Tensor* load_tensor(Context& ctx, const std::string& path);
void save_ctm(Context& ctx, const Ctm& c, int64_t chi, const std::string& path);
Ctm* load_ctm(Context& ctx, const std::string& path);
Tensor* tensor_from_charge_blocks(Context& ctx, int nlegs, int rank, const int* ncharge, const double* const* charges, const double* const* sdims, const int* conj,
                                  int64_t nblocks, const int64_t* voffs_index, const int64_t* voffs_offset, const double* vdat, int64_t nvdat, int qn_scale);

// RAII event pair around a launch group; `work` = algorithmic bytes (permute) or flops (GEMM).
struct ProfScope
{
    Context& c;
    cudaEvent_t e0 = nullptr;
    int cat;
    double work;
    ProfScope(Context& ctx, int category, double w) : c(ctx), cat(category), work(w)
    {
        if(c.profiling) {
            e0 = c.get_event();
            cudaEventRecord(e0, c.stream);
        }
    }
    ~ProfScope()
    {
        if(e0) {
            cudaEvent_t e1 = c.get_event();
            cudaEventRecord(e1, c.stream);
            c.prof_recs.push_back({e0, e1, cat, work});
        }
    }
};

// ---- kernel launchers (kernels.cu) ----
void launch_permute(const double* src, double* dst, const PermutePlan& plan, cudaStream_t s);
// pushed permute over peer memory: destination offsets of the plan carry the owner rank of the destination shard in their top byte (PERM_PEER_SHIFT); the
// element offset of rank p's shard relative to `dst` is delta.d[p] (0 for this rank)
void launch_permute_peers(const double* src, double* dst, const PermutePlan& plan, const PeerDelta& delta, cudaStream_t s);
// sources with offset >= split are read from src2 + (offset - split): sharded permute reading [local shard | receive buffer]
void launch_permute2(const double* src, const double* src2, long long split, double* dst, const PermutePlan& plan, cudaStream_t s);
// deterministic LPT greedy (largest cost first, ties by index, least-loaded rank, ties lowest rank)
std::vector<int> lpt_owner(const std::vector<double>& cost, int world);
void launch_gemm_grouped(const double* A, const double* B, double* C, const GemmItem* d_items, int n_items, int tile_m, cudaStream_t s);
void launch_splitk_reduce(const double* scratch, double* C, const SplitKJob* d_jobs, int n_jobs, int max_job_elems, cudaStream_t s);
void launch_scale_copy(const double* x, double* y, double alpha, int64_t n, cudaStream_t s);
void launch_axpby(const double* x, const double* y, double* z, double alpha, double beta, int64_t n, cudaStream_t s);
void launch_absmax(const double* x, int64_t n, double* d_out, cudaStream_t s);
void launch_scale_by_inv(const double* x, double* y, const double* d_scalar, int64_t n, cudaStream_t s);
void launch_fill_constant(double* x, double v, int64_t n, cudaStream_t s);
void launch_fill_random(double* x, uint64_t seed, int64_t n, cudaStream_t s, int64_t index_base = 0);   // element i gets the value of counter i + index_base
struct BlockDesc { long long off; int rows, cols; double weight; };
void launch_transpose_blocks(const double* src, double* dst, const BlockDesc* d_src_blocks, const long long* d_dst_off, int nblocks, cudaStream_t s);
void launch_set_identity(double* x, const BlockDesc* d_blocks, int nblocks, cudaStream_t s);
void launch_diag_map(const double* src, double* dst, const BlockDesc* d_blocks, int nblocks, int op, cudaStream_t s);
struct SubScale { long long off; int rows, cols, ld; double scale; };
void launch_scale_subblocks(double* data, const SubScale* d_subs, int n, cudaStream_t s);
// SVD reverse rule helpers (include/Xped/AD/ADTensor.hpp:253-325)
struct DiagScaleDesc { long long off, s_off; int rows, cols, s_ld, by_cols; };
// dst = (accumulate ? dst : 0) + src / s  with s taken per column (by_cols) or per row from the diagonal of an S block
void launch_scale_by_diag(const double* src, double* dst, const double* S, const DiagScaleDesc* d_desc, int n, int accumulate, cudaStream_t s);
// M = dS + (J + J^T) S + S (K + K^T) - G1 S^-1 - S^-1 G2^T,  J = F o G1, K = F o G2, F_ij = 1 / (s_j^2 - s_i^2) (regularised)
void launch_svd_adjoint_core(const double* G1, const double* G2, const double* dS, const double* S, double* M, const BlockDesc* d_blocks, int nblocks,
                             cudaStream_t s);
void launch_trace(const double* src, const BlockDesc* d_blocks, int nblocks, double* d_out, cudaStream_t s);
// reverse rules of the element-wise ops (ADTensor.hpp:511-524, 559-586)
void launch_diag_map_backward(const double* g, const double* y, double* dst, const BlockDesc* d_blocks, int nblocks, int op, cudaStream_t s);
void launch_dot(const double* a, const double* b, int64_t n, double* d_part /* 129 doubles: [0] receives the result */, cudaStream_t s);
void launch_maxnorm_scale_backward(const double* g, const double* x, const double* d_m, const double* d_dot, double* out, int64_t n, cudaStream_t s);

}   // namespace xb
