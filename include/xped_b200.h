/* xped_b200 -- C ABI of the B200-native backend for the Xped block-sparse symmetric-tensor hot path.
 *
 * Drop-in boundary: Xped selects its dense backend at compile time through the
 * `PlainInterface` policy struct (include/Xped/Interfaces/PlainInterface.hpp:102-110, CMake
 * XPED_{MATRIX,TENSOR,VECTOR}_LIB + XPED_STORAGE, CMakeLists.txt:35-70).  That per-block interface is too
 * fine grained for a device (one launch per tree pair) and is not sealed (Core reaches into Eigen directly,
 * src/Core/Tensor.cpp:661-676,768-772), so this library sits one level higher: each entry point replaces one
 * public method of `Xped::Tensor<double,Rank,CoRank,Symmetry>` / `TensorBase` as cited below.  A new
 * `XPED_MATRIX_LIB=B200` + `XPED_STORAGE=DeviceArena` choice binds these symbols.
 *
 * Conventions: every function returns 0 on success or an xb_status; xb_last_error() gives the message.
 * Handles own device memory; functions never free their inputs; outputs are fresh handles released with
 * xb_*_free.  Plain pointers and sizes only.  All work is enqueued on the context's CUDA stream; functions
 * that return host values (download, norms, tSVD truncation) synchronise that stream.
 * There is NO CPU fallback: xb_ctx_create fails with XB_ERR_NO_DEVICE when no CUDA device is visible.
 */
#ifndef XPED_B200_H_
#define XPED_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xb_ctx xb_ctx;
typedef struct xb_basis xb_basis;   /* Qbasis<Symmetry,1>  (include/Xped/Core/Qbasis.hpp:37-268) */
typedef struct xb_tensor xb_tensor; /* Tensor<double,Rank,CoRank,Symmetry> (include/Xped/Core/Tensor.hpp:59-444) */
typedef struct xb_ctm xb_ctm;       /* CTM<double,Symmetry,2> (include/Xped/PEPS/CTM.hpp:60-312) */
typedef struct xb_dist_plan xb_dist_plan; /* exchange plan of one sector-sharded permute (no reference counterpart: the
                                             reference's distributed storage is the Cyclops backend, Interfaces/TensorInterface_Cyclops_impl.cpp) */

enum xb_status {
    XB_OK = 0,
    XB_ERR_INVALID = 1,   /* bad argument / basis or shape mismatch (reference: assert / DEBUG_ASSERT) */
    XB_ERR_NO_DEVICE = 2, /* no CUDA device: the product path never falls back to the CPU */
    XB_ERR_CUDA = 3,
    XB_ERR_NOMEM = 4,     /* reference: std::bad_alloc rethrow, StorageType_contiguous.hpp:91-96 */
    XB_ERR_INTERNAL = 5
};

/* Symmetry factors: Sym::U0, Sym::U1<Kind>, Sym::ZN<Kind,N>, Sym::SU2<Kind>; products (Sym::S1xS2 / Combined)
 * are given as a list.  `fermionic` selects the FChargeU1 / FChargeSU2 style kinds. */
enum xb_sym_kind { XB_SYM_U0 = 0, XB_SYM_U1 = 1, XB_SYM_ZN = 2, XB_SYM_SU2 = 3 };
typedef struct xb_sym_factor { int kind; int N; int fermionic; } xb_sym_factor;

const char* xb_last_error(void);
const char* xb_version(void);

/* ---- context: one per GPU / host thread (replaces mpi::XpedWorld threading, include/Xped/Util/Mpi.hpp:17-47) */
int xb_ctx_create(xb_ctx** ctx, int device, const xb_sym_factor* factors, int n_factors);
/* host-metadata-only context (bases, layouts, owner maps, exchange plans) for a process without a GPU; every call that
 * would touch tensor data fails with XB_ERR_NO_DEVICE.  It is NOT a CPU fallback: it cannot hold or move tensor data. */
int xb_ctx_create_planning(xb_ctx** ctx, const xb_sym_factor* factors, int n_factors);
int xb_ctx_free(xb_ctx* ctx);
int xb_ctx_sync(xb_ctx* ctx);
int xb_ctx_nq(const xb_ctx* ctx);
/* integer options / counters.  set: "svd.truncated" = 1 / 0 forces / forbids the truncated tSVD algorithm (randomized subspace iteration + Jacobi on the
 * projected blocks, used when maxKeep << number of singular values; falls back to the full Jacobi path when its residual check fails), -1 = default;
 * "svd.call_tag" = an integer naming the call site of the following tSVD calls: the truncated algorithm starts from the subspace the previous call with the
 * same (tensor layout, tag) ended with (plus fresh random columns), which in an iteration that converges (CTMRG: one tag per move direction) cuts its work
 * several times; any tag is always correct, the a-posteriori check does not depend on the start.
 * get: the same key, and the counters "svd.truncated_calls", "svd.truncated_fallbacks", "svd.truncated_iterations" */
int xb_ctx_set_int(xb_ctx* ctx, const char* key, int64_t value);
int xb_ctx_get_int(const xb_ctx* ctx, const char* key, int64_t* value);
/* counters: [0] kernel launches, [1] permute-plan cache hits, [2] plan cache misses */
int xb_ctx_stats(const xb_ctx* ctx, int64_t out[3]);

/* ---- Qbasis<Symmetry,1> ------------------------------------------------------------------------------
 * q: n * Nq ints (qarray<Nq> per sector, include/Xped/Symmetry/qarray.hpp:29-80), dims: n inner dims.
 * Sectors are push_back()ed in the given order (src/Core/Qbasis.cpp:34-74); equal q merge. */
int xb_basis_create(xb_ctx* ctx, const int* q, const int64_t* dims, int n, xb_basis** out);
int xb_basis_free(xb_basis* b);
int xb_basis_conj(const xb_basis* b, xb_basis** out);                  /* Qbasis::conj, Qbasis.cpp:418-439 */
int xb_basis_sorted(const xb_basis* b, xb_basis** out);                /* copy + Qbasis::sort, Qbasis.cpp:306-331 */
int xb_basis_combine(const xb_basis* a, const xb_basis* b, xb_basis** out); /* combine().forgetHistory(), :441-495,:410-416 */
int xb_basis_size(const xb_basis* b, int* n_sectors, int64_t* dim, int64_t* full_dim, int* is_conj);
int xb_basis_get(const xb_basis* b, int* q /* n*Nq */, int64_t* dims /* n */);

/* ---- Tensor construction / storage ---------------------------------------------------------------------
 * xb_tensor_create: Tensor(basis_domain, basis_codomain) + setZero()  (Tensor.hpp:106-110, Tensor.cpp:46-55) */
int xb_tensor_create(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, xb_tensor** out);
int xb_tensor_free(xb_tensor* t);
int xb_tensor_clone(const xb_tensor* t, xb_tensor** out);
int xb_tensor_rank(const xb_tensor* t, int* rank, int* corank);
int xb_tensor_leg_basis(const xb_tensor* t, int leg, xb_basis** out);  /* uncoupledDomain()/uncoupledCodomain() */
int64_t xb_tensor_plain_size(const xb_tensor* t);                      /* plainSize(), StorageType_contiguous.hpp:149 */
int xb_tensor_num_sectors(const xb_tensor* t);
/* sector(i), block(i).rows()/cols(), m_offsets[i]: the data the bit-exact structure tests compare */
int xb_tensor_structure(const xb_tensor* t, int* q /* nsec*Nq */, int64_t* rows, int64_t* cols, int64_t* offsets);
/* number of (tree, cotree) pairs and, per pair: sector index, row offset (leftOffset), col offset, rows, cols */
int64_t xb_tensor_num_tree_pairs(const xb_tensor* t);
int xb_tensor_tree_pairs(const xb_tensor* t, int64_t* out5 /* npairs*5 */);
/* FusionTree<Rank> f1 / FusionTree<CoRank> f2 of tree pair `pair` (enumeration order of xb_tensor_tree_pairs): q_uncoupled (Rank+CoRank)*Nq ints,
 * q_intermediate (max(Rank-2,0) + max(CoRank-2,0))*Nq ints, q_coupled Nq ints, dims Rank+CoRank (include/Xped/Core/FusionTree.hpp:28-35); any pointer may be NULL */
int xb_tensor_tree_pair_trees(const xb_tensor* t, int64_t pair, int* q_uncoupled, int* q_intermediate, int* q_coupled, int64_t* dims);
/* index of the tree pair (f1, f2) given by its labels (FusionTree::operator== compares labels, FusionTree.cpp:66-85), or -1 */
int64_t xb_tensor_find_tree_pair(const xb_tensor* t, const int* q_uncoupled, const int* q_intermediate, const int* q_coupled);
/* Tensor::view(f1, f2) / l-value subMatrix(f1, f2)  src/Core/Tensor.cpp:830-905, Tensor.hpp:196-216: the f1.dim x f2.dim sub-block IN PLACE --
 * device pointer to its (0, 0) element and the leading dimension (rows of the sector block); column-major, the row / column multi-index runs
 * over the legs with leg 0 fastest.  Valid while the tensor lives; order your work after the context stream (xb_ctx_stream). */
int xb_view(const xb_tensor* t, int64_t pair, double** dev_ptr, int64_t* ld, int64_t* rows, int64_t* cols);
/* Tensor::subBlock(f1, f2) / const subMatrix  src/Core/Tensor.cpp:907-981: a dense copy of the sub-block on the host (n = f1.dim * f2.dim doubles,
 * = the rank-(Rank+CoRank) tensor of dims f1.dims x f2.dims in column-major order); xb_set_sub_block writes one back (setSubTensor on a view) */
int xb_sub_block(const xb_tensor* t, int64_t pair, double* host, int64_t n);
int xb_set_sub_block(xb_tensor* t, int64_t pair, const double* host, int64_t n);
/* raw `m_data` image, host <-> device (Tensor::data()/set_data(), Tensor.hpp:137-152) */
int xb_tensor_upload(xb_tensor* t, const double* host, int64_t n);
int xb_tensor_download(const xb_tensor* t, double* host, int64_t n);
int xb_tensor_set_zero(xb_tensor* t);
int xb_tensor_set_identity(xb_tensor* t);                              /* Tensor.cpp:57-68 */
int xb_tensor_set_constant(xb_tensor* t, double v);                    /* Tensor.cpp:70-81 */
/* counter-based uniform[-1,1] fill keyed by (seed, element index): same values as the oracle's helper */
int xb_tensor_set_random(xb_tensor* t, uint64_t seed);

/* ---- hot path --------------------------------------------------------------------------------------------*/
/* Tensor::permute<shift,p...>()  src/Core/Tensor.cpp:335-355 (-> permute_impl :83-146 / :148-214) */
int xb_permute(xb_ctx* ctx, const xb_tensor* t, int shift, const int* p, int n, xb_tensor** out);
/* Tensor::permute_adj<shift,p...>()  src/Core/Tensor.cpp:258-333 (AD reverse rule, ADTensor.hpp:171-185) */
int xb_permute_adj(xb_ctx* ctx, const xb_tensor* t, int shift, const int* p, int n, xb_tensor** out);
/* TensorBase::operator*  src/Core/TensorBase.cpp:243-288; adj_a/adj_b read an operand as its lazy AdjointOp
 * (include/Xped/Core/AdjointOp.hpp:26-69), which is also what the GEMM reverse rules need (ADTensor.hpp:749-758) */
int xb_multiply(xb_ctx* ctx, const xb_tensor* a, int adj_a, const xb_tensor* b, int adj_b, xb_tensor** out);
/* Tensor::contract<a1,a2,ResRank>()  include/Xped/Core/Tensor.hpp:328-351 (labels as in NCON: >0 contracted, <0 open) */
int xb_contract(xb_ctx* ctx, const xb_tensor* a, const int* labels_a, int na, const xb_tensor* b, const int* labels_b, int nb,
                int res_rank, xb_tensor** out);
/* TensorBase::adjoint().eval()  TensorBase.cpp:82-86, AdjointOp.hpp:44-65 */
int xb_adjoint(xb_ctx* ctx, const xb_tensor* t, xb_tensor** out);
/* Tensor::twist(leg)  src/Core/Tensor.cpp:357-377 */
int xb_twist(xb_ctx* ctx, const xb_tensor* t, int leg, xb_tensor** out);
/* alpha*x (+ beta*y): operator*(scalar), operator+/-  (TensorBase.hpp:197-209, CoeffBinaryOp.hpp:44-64); y may be NULL */
int xb_axpby(xb_ctx* ctx, double alpha, const xb_tensor* x, double beta, const xb_tensor* y, xb_tensor** out);
/* diag_sqrt / diag_inv  TensorBase.cpp:191-202; op: 0 = sqrt, 1 = inv, 2 = inv(sqrt) */
int xb_diag_map(xb_ctx* ctx, const xb_tensor* t, int op, xb_tensor** out);
int xb_max_norm(xb_ctx* ctx, const xb_tensor* t, double* out);        /* TensorBase.cpp:48-56 */
int xb_trace(xb_ctx* ctx, const xb_tensor* t, double* out);           /* TensorBase.cpp:27-37 */
int xb_squared_norm(xb_ctx* ctx, const xb_tensor* t, double* out);    /* TensorBase.cpp:58-64 */
/* Tensor::tSVD(maxKeep, eps_svd, truncWeight, entropy, ..., PRESERVE_MULTIPLETS)  src/Core/Tensor.cpp:478-633.
 * trunc_weight / entropy are in-out accumulators like the reference's reference parameters. */
int xb_tsvd(xb_ctx* ctx, const xb_tensor* t, int64_t max_keep, double eps_svd, int preserve_multiplets, xb_tensor** U, xb_tensor** S,
            xb_tensor** Vdag, double* trunc_weight, double* entropy);
/* reverse rule of the full SVD + truncation  include/Xped/AD/ADTensor.hpp:253-325,356-371 */
int xb_tsvd_backward(xb_ctx* ctx, const xb_tensor* A, const xb_tensor* U, const xb_tensor* S, const xb_tensor* Vdag, const xb_tensor* dU,
                     const xb_tensor* dS, const xb_tensor* dVdag, xb_tensor** dA);
/* Tensor::tQR(bool RETURN_LQ)  src/Core/Tensor.cpp:757-780: Q (domain -> middle), R (middle -> codomain), middle = the smaller coupled basis.
 * Blocks are Q_full 1(m x mid), 1(mid x m) R_full (mid = inner dimension of the middle basis; the reference's matrices wherever those fit
 * its own bases).  return_lq != 0: the adjoint blocks (Q^T, R^T) like the reference, square blocks only (XB_ERR_INVALID otherwise).
 * Tensor::teigh()  :782-803 */
int xb_tqr(xb_ctx* ctx, const xb_tensor* t, int return_lq, xb_tensor** Q, xb_tensor** R);
int xb_teigh(xb_ctx* ctx, const xb_tensor* t, xb_tensor** D, xb_tensor** V);
/* decompose(T1,T2,chi)  src/PEPS/PEPSContractions.cpp:18-43: CTM projectors P1 = S^-1/2 U^+ T1, P2 = T2 V S^-1/2 */
int xb_decompose(xb_ctx* ctx, const xb_tensor* T1, const xb_tensor* T2, int64_t chi, xb_tensor** P1, xb_tensor** P2);
/* contract_L / contract_R  src/MPS/MpsContractions.cpp:11-51 / :54-105 */
int xb_contract_L(xb_ctx* ctx, const xb_tensor* Bold, const xb_tensor* Bra, const xb_tensor* Ket, xb_tensor** Bnew);
int xb_contract_R(xb_ctx* ctx, const xb_tensor* Bold, const xb_tensor* Bra, const xb_tensor* Ket, xb_tensor** Bnew);
/* Mps::leftSweepStep / Mps::rightSweepStep (DMRG::BROOM::SVD)  src/MPS/Mps.cpp:181-202 / :204-229 on the site tensors Ac[loc] (Tensor<2,1>)
 * and its neighbour Ac[loc-1] / Ac[loc+1] (NULL at the chain end or for DISCARD_U / DISCARD_V; *A_nb_new is then NULL).
 * trunc_weight / entropy receive truncWeight[loc] and S[loc]. */
int xb_mps_left_sweep_step(xb_ctx* ctx, const xb_tensor* A_loc, const xb_tensor* A_prev, int64_t max_nsv, double eps_svd, xb_tensor** A_loc_new,
                           xb_tensor** A_nb_new, double* trunc_weight, double* entropy);
int xb_mps_right_sweep_step(xb_ctx* ctx, const xb_tensor* A_loc, const xb_tensor* A_next, int64_t max_nsv, double eps_svd, xb_tensor** A_loc_new,
                            xb_tensor** A_nb_new, double* trunc_weight, double* entropy);

/* ---- sector sharding over the GPUs of one box ----------------------------------------------------
 * One process / context per GPU.  A sharded tensor = on every rank a tensor holding only the sectors that rank owns
 * (xb_tensor_create_sectors), plus an owner map over the canonical sector list that every rank computes identically
 * (xb_dist_owner_lpt: deterministic LPT greedy).  Products, decompositions and element-wise maps act on the shards
 * (sectors are independent, src/Core/TensorBase.cpp:262-279); a permute moves sub-blocks between sectors
 * (src/Core/Tensor.cpp:198-208) and has ONE exchange step: pack -> alltoallv (issued by the host framework over NCCL with the
 * counts below) -> gather.  The reference's distributed path is MPI/Cyclops (XPED_USE_MPI, include/Xped/Util/Mpi.hpp). */
/* canonical structure of Tensor(domain, codomain) without allocating it; q == NULL: only the sector count */
int xb_layout_structure(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int* n_sectors, int* q,
                        int64_t* rows, int64_t* cols, int64_t* offsets);
/* Host-side statistics of the cached plan of permute<shift,p> of the canonical Tensor(domain | codomain) (works in planning-only
 * contexts): out = {destination tasks, LSU tiles, max fan-in, identity copy?, cp.async.bulk tiles (0 = plan not eligible), algorithmic bytes} */
int xb_permute_plan_stats(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n,
                          int64_t out[6]);
/* census of the same plan: destination elements, source elements read (an element feeding k destination sub-blocks counts k times: SU(2) recoupling
 * sums), destination tasks, source records, tiles of the table-driven form (0: the plan does not have it), bytes of its offset tables */
int xb_permute_plan_census(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n, int64_t out[6]);
/* TEST SUPPORT, planning-only contexts only (no GPU; the tensors of such a context own no data, so no product path can reach this): applies the HOST tables
 * of the plan of permute<shift,p> (adjoint != 0: permute_adj) to host arrays the way the kernels address them -- form 0: tile list of the LSU form, form 1:
 * table-driven form, form 2: cp.async.bulk form (every bulk operation and the transposing pass; alignment and mbarrier byte counts checked) -- with every offset bounds-checked and every destination element written exactly once.  `src` / `dst` are the `m_data` images of the
 * input / output tensor.  The CPU tests compare the result with the oracle's Tensor::permute (src/Core/Tensor.cpp:83-214). */
int xb_permute_plan_emulate(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n, int adjoint,
                            int form, const double* src, int64_t n_src, double* dst, int64_t n_dst);
/* TEST SUPPORT, planning-only contexts only: the host lists of op(A) * op(B) (tile items of both kernel variants, split-K jobs, scratch slices; built for a
 * device with `sm_count` SMs, <= 0: the context's default of 148) applied to host arrays the way gemm_grouped_kernel / splitk_reduce_kernel address device
 * memory, bounds-checked and write-once-checked.  counts = {128-row tile items, 64-row tile items, split-K jobs, scratch elements}.  The CPU tests compare the
 * result with per-sector products of the oracle (TensorBase::operator*, src/Core/TensorBase.cpp:243-288). */
int xb_multiply_plan_emulate(xb_ctx* ctx, xb_basis* const* a_domain, int a_rank, xb_basis* const* a_codomain, int a_corank, int adj_a, xb_basis* const* b_domain,
                             int b_rank, xb_basis* const* b_codomain, int b_corank, int adj_b, int sm_count, const double* a, int64_t n_a, const double* b, int64_t n_b,
                             double* c, int64_t n_c, int64_t counts[4]);
/* TEST SUPPORT, planning-only contexts only: contract<labels_a, labels_b, res_rank> (include/Xped/Core/Tensor.hpp:328-351) the way the device path lowers it --
 * get_permutations (Util/Constfct.hpp:55-109), two operand permutes, the grouped product, the permute of the result -- with every plan / list walked on the
 * host (table_form != 0: the table-driven form of the permute plans where a plan has it, else their tile lists). */
int xb_contract_plan_emulate(xb_ctx* ctx, xb_basis* const* a_domain, int a_rank, xb_basis* const* a_codomain, int a_corank, const int* labels_a,
                             xb_basis* const* b_domain, int b_rank, xb_basis* const* b_codomain, int b_corank, const int* labels_b, int res_rank, int table_form,
                             const double* a, int64_t n_a, const double* b, int64_t n_b, double* c, int64_t n_c);
/* a shard: Tensor(domain, codomain) holding only the listed sectors, in that order, zero-initialised */
int xb_tensor_create_sectors(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, const int* q, int n_sectors,
                             xb_tensor** out);
/* counter-based fill of a shard: sector i takes the values the unsharded tensor has at storage offset global_offsets[i] */
int xb_tensor_set_random_at(xb_tensor* t, uint64_t seed, const int64_t* global_offsets);
int xb_dist_owner_lpt(const double* cost, int n, int world, int* owner);
/* plan of permute<shift,p> (adjoint != 0: permute_adj) of the canonical tensor on (domain | codomain) for rank my_rank of world;
 * owner_in: one rank per input sector; owner_out: one rank per output sector or NULL (LPT on rows*cols*(rows+cols)) */
int xb_dist_permute_plan(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n,
                         int adjoint, const int* owner_in, const int* owner_out, int n_out_sectors, int my_rank, int world, xb_dist_plan** out);
int xb_dist_plan_free(xb_dist_plan* plan);
/* alltoallv counts in doubles, `world` entries each; segments are ordered by peer rank */
int xb_dist_plan_counts(const xb_dist_plan* plan, int64_t* send_counts, int64_t* recv_counts);
/* output side: sector count and owner map of the output's canonical sectors, its leg bases (rank + corank fresh handles); any pointer may be NULL */
int xb_dist_plan_out(const xb_dist_plan* plan, int* n_out_sectors, int* owner_out, xb_basis** legs, int* out_rank, int* out_corank);
/* sector labels of this rank's input (which = 0) / output (which = 1) shard; q == NULL: only the count */
int xb_dist_plan_local_sectors(const xb_dist_plan* plan, int which, int* n_sectors, int* q);
/* the plan's work lists as flat int64 records (introspection, CPU tests of the exchange logic); out == NULL: only the count.
 * which = 0 pack:    peer, send_off, src_off, nleg, dims[16], src_stride[16]
 *         1 tasks:   dst_off, nleg, src_begin, src_count, dims[16], dst_stride[16]
 *         2 sources: remote, off (shard offset, or receive-buffer offset if remote), coeff (double bits), 0, src_stride[16] */
int xb_dist_plan_records(const xb_dist_plan* plan, int which, int64_t* n_records, int64_t* out);
/* the three device steps of a sharded permute; d_send / d_recv are device buffers of sum(send_counts) / sum(recv_counts) doubles */
int xb_dist_permute_pack(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, double* d_send);
int xb_dist_permute_local(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, xb_tensor** local_out);
int xb_dist_permute_remote(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, const double* d_recv, xb_tensor* local_out);
/* operator* on shards: the result holds exactly the listed sectors (the ones this rank owns of the product) */
/* ---- peer-memory form of the exchange (csrc/dist.cu, csrc/dist_ops.cpp): every rank maps the device arena of every other rank of the box (CUDA IPC), so the
 * exchange step of a sharded permute is fused into the permute launches (the owner of a source sub-block transposes it straight into the output shard of the
 * destination's owner over NVLink) and a product can be balanced by ROW PANELS of the heavy sectors.  Start-up: rank 0 calls xb_dist_shm_create(name, world),
 * all ranks synchronise (any host barrier), every rank calls xb_dist_init (collective; tensors created afterwards live in the arena).  All ranks must issue the
 * same xb_dist_* operations in the same order.  No reference counterpart (the reference's MPI/Cyclops backend distributes dense blocks inside CTF). */
int xb_dist_shm_create(const char* shm_name, int world);
int xb_dist_init(xb_ctx* ctx, int my_rank, int world, int64_t arena_bytes, const char* shm_name);
int xb_dist_shutdown(xb_ctx* ctx);
int xb_dist_barrier(xb_ctx* ctx);                          /* stream-ordered barrier over all ranks (device side, NVLink flags) */
int xb_dist_arena_stats(const xb_ctx* ctx, int64_t out[3]); /* arena bytes, bytes in use, high-water mark */
int xb_dist_plan_pushable(const xb_dist_plan* plan, int* pushable);   /* 1: fan-in 1 everywhere, xb_dist_permute_push applies */
int xb_dist_permute_push(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, xb_tensor** local_out);
/* operator* on shards (src/Core/TensorBase.cpp:243-288); owner_a / owner_b: owner of every sector of the GLOBAL operands (common sectors must agree); sectors
 * that overload their owner are cut into row panels of >= panel_rows_min rows computed by other ranks; owner_out (may be null): owner of the result's sectors */
int xb_dist_multiply(xb_ctx* ctx, const xb_tensor* a_local, const int* owner_a, const xb_tensor* b_local, const int* owner_b, int panel_rows_min, xb_tensor** c_local,
                     int* owner_out, int n_owner_out);
/* the deterministic panel assignment used by xb_dist_multiply (host only): per panel (sector, rank, row0, rows) as four int64 */
int xb_dist_multiply_assignment(const int64_t* m, const int64_t* n, const int64_t* k, const int* owner, int n_sectors, int world, int panel_rows_min,
                                int64_t* n_panels, int64_t* out);

int xb_multiply_sectors(xb_ctx* ctx, const xb_tensor* a, int adj_a, const xb_tensor* b, int adj_b, const int* q, int n_sectors, xb_tensor** out);

/* ---- CTM driver (1x1 unit cell, TRank = 2, corner projectors)  src/PEPS/CTM.cpp:45-244,385-420,924-1760 ------
 * A: iPEPS site tensor Tensor<double,2,3> (left, top | right, bottom, phys), include/Xped/PEPS/iPEPS.hpp. */
int xb_ctm_create(xb_ctx* ctx, const xb_tensor* A, int64_t chi, xb_ctm** out);   /* CTM(A, chi, FROM_A) + init() */
int xb_ctm_free(xb_ctm* c);
int xb_ctm_step(xb_ctm* c);                                                      /* grow_all(): left, right, top, bottom move */
/* which: 0..3 = C1..C4, 4..7 = T1..T4, 8 = A, 9 = Adag. Returns a clone. */
int xb_ctm_get(const xb_ctm* c, int which, xb_tensor** out);
int xb_ctm_set(xb_ctm* c, int which, const xb_tensor* t);
/* single directional pieces, for parity tests: corner 0..3 = UL, UR, LR, LL (Opts::CORNER) */
int xb_ctm_corner(xb_ctm* c, int corner, xb_tensor** Q);                         /* contractCorner  CTM.cpp:1643-1760 */
int xb_ctm_move(xb_ctm* c, int direction);                                       /* 0 left, 1 right, 2 top, 3 bottom  :924-1183 */
/* computeRDM_h (which = 0) / computeRDM_v (which = 1)  src/PEPS/CTM.cpp:528-874 (TRank == 2): normalised two-site density matrix */
int xb_ctm_rdm(xb_ctm* c, int which, xb_tensor** rho);
/* avg(rho, O, bond) of a two-site observable  src/PEPS/LinearAlgebra.cpp:57-60 */
int xb_avg_two_site(xb_ctx* ctx, const xb_tensor* rho, const xb_tensor* op, double* out);
/* rho1 = normalised partial trace of a two-site density matrix (CTM.cpp:675-678); avg(rho1, O) of a one-site observable (LinearAlgebra.cpp:22-24) */
int xb_rdm_one_site(xb_ctx* ctx, const xb_tensor* rho, xb_tensor** rho1);
int xb_avg_one_site(xb_ctx* ctx, const xb_tensor* rho1, const xb_tensor* op, double* out);

/* ---- state interop ----------------------------------------------------------------------------------------------------
 * Native container: symmetry, leg bases, sector order and the raw m_data image -- the fields the reference serialises through yas
 * (include/Xped/Core/storage/StorageType_contiguous.hpp:191-204, Qbasis.hpp:249-253, Tensor.hpp:409-413); xb_ctm_save / xb_ctm_load write / read chi, C1..C4,
 * T1..T4, A, Adag (checkpoint / resume of an environment, examples/chi_ramp.cpp:201-204).  The yas byte format is not reproduced (no yas, no reference-written
 * file in this image to check against).  Loading checks the symmetry and that the stored block structure follows from the stored bases. */
int xb_tensor_save(xb_ctx* ctx, const xb_tensor* t, const char* path);
int xb_tensor_load(xb_ctx* ctx, const char* path, xb_tensor** out);
int xb_ctm_save(const xb_ctm* c, const char* path);
int xb_ctm_load(xb_ctx* ctx, const char* path, xb_ctm** out);
/* IO::loadMatlabTensor  include/Xped/IO/Matlab.hpp:9-95 on the arrays of the HDF5 group (the container itself needs HighFive / libhdf5):
 * charges[leg][j] / sector_dims[leg][j] = rows 0 / 1 of the leg's `charges` dataset (j < ncharge[leg]); voffs_index / voffs_offset = the columns of `voffs`
 * (combined charge index with leg 0 fastest, offset into vdat); conj[leg] != 0 marks the leg's basis as dual; charges are multiplied by qn_scale and, for Z_N,
 * reduced modulo N.  All legs are placed in the domain of a rank-(N,0) tensor coupled to the vacuum, then permute<N - rank>(identity) (:93).  One Abelian factor. */
int xb_tensor_from_charge_blocks(xb_ctx* ctx, int nlegs, int rank, const int* ncharge, const double* const* charges, const double* const* sector_dims, const int* conj,
                                 int64_t nblocks, const int64_t* voffs_index, const int64_t* voffs_offset, const double* vdat, int64_t nvdat, int qn_scale,
                                 xb_tensor** out);

/* ---- reverse-mode AD (include/Xped/AD/ADTensor.hpp; replaces the stan::math tape): between xb_tape_begin and xb_tape_end the differentiable entry
 * points above record their reverse rule -- xb_permute (:171-185, permute_adj), xb_multiply (:749-758), xb_contract (:193-219, as its three permutes and the
 * product), xb_adjoint (:137-150), xb_twist (:589-602), xb_axpby (:640-676), xb_diag_map (:559-586), xb_tsvd (:237-374, full path + xb_tsvd_backward),
 * xb_tensor_clone, xb_normalize_max (maxNorm rule :502-529 fused with the division), xb_normalize_trace (trace rule :531-546 with operator*(tensor, var)
 * :690-706), xb_decompose.  The loss is a weighted sum of traces: xb_tape_seed_trace(v, w) adds w * d trace(v) / d v (Tensor::WeightedIdentity); any other
 * cotangent can be injected with xb_tape_seed.  xb_tape_backward runs the rules in reverse order; xb_tape_grad returns a copy of the accumulated adjoint
 * (plain d loss / d m_data element, the convention of the reference's adj() tensors) or NULL.  The tape keeps the values it needs alive until xb_tape_end. */
int xb_tape_begin(xb_ctx* ctx);
int xb_tape_end(xb_ctx* ctx);
int xb_tape_seed_trace(xb_ctx* ctx, const xb_tensor* v, double weight);
int xb_tape_seed(xb_ctx* ctx, const xb_tensor* v, const xb_tensor* adj);
int xb_tape_backward(xb_ctx* ctx);
int xb_tape_grad(xb_ctx* ctx, const xb_tensor* v, xb_tensor** g);
int xb_tape_size(xb_ctx* ctx, int64_t* n_nodes);
/* t * (1 / t.maxNorm())  (renormalize_*: src/PEPS/CTM.cpp:1369,1381,1409) and x * (1 / trace(x * w))  (density-matrix normalisation, CTM.cpp:660-667) */
int xb_normalize_max(xb_ctx* ctx, const xb_tensor* t, xb_tensor** out);
int xb_normalize_trace(xb_ctx* ctx, const xb_tensor* x, const xb_tensor* w, xb_tensor** out);
/* E(A) = avg(rho_h, op_h) + avg(rho_v, op_v) of the 1x1 iPEPS A after `steps` grow_all() steps from CTM_INIT::FROM_A, and dE/dA by the reverse sweep
 * (what the reference's energy functor hands to the optimiser, src/PEPS/CTMSolver.cpp:40-75).  checkpoint != 0: one checkpoint per move, replayed during the
 * reverse sweep (src/PEPS/CTM.cpp:399-419,976-987,1327-1335,1407-1416,1751-1758) -- same gradient, less memory.  dA may be NULL (value only). */
int xb_ctm_energy_grad(xb_ctx* ctx, const xb_tensor* A, int64_t chi, int steps, const xb_tensor* op_h, const xb_tensor* op_v, int checkpoint, double* energy,
                       xb_tensor** dA);

/* ---- micro-benchmark hooks (device-resident, timed with CUDA events on the context stream) -------------- */
/* runs `reps` launches of the cached plan of permute<shift,p>(t); returns avg ms per launch and algorithmic bytes */
int xb_bench_permute(xb_ctx* ctx, const xb_tensor* t, int shift, const int* p, int n, int reps, double* ms_per_launch, int64_t* bytes);
int xb_bench_multiply(xb_ctx* ctx, const xb_tensor* a, const xb_tensor* b, int reps, double* ms_per_launch, double* flops);
int xb_flush_l2(xb_ctx* ctx);
/* the context's CUDA stream (cudaStream_t) so that a host framework can order its own work after ours */
void* xb_ctx_stream(xb_ctx* ctx);
/* CUDA-event stopwatch on the context stream: start records an event, stop records a second one, synchronises on
 * it and returns the elapsed device time in ms (covers host gaps between enqueued work, like a wall clock) */
int xb_timer_start(xb_ctx* ctx);
int xb_timer_stop(xb_ctx* ctx, double* ms);
/* live per-category device timing: while enabled every launch group is bracketed by a CUDA event pair.
 * xb_profile_read synchronises, returns {ms, launches, work} summed per category and resets; categories:
 * 0 permute (work = algorithmic bytes), 1 grouped DGEMM (work = flops), 2 tSVD (work = 0), 3 other */
int xb_profile_enable(xb_ctx* ctx, int on);
int xb_profile_read(xb_ctx* ctx, double ms[4], int64_t count[4], double work[4]);

#ifdef __cplusplus
}
#endif
#endif /* XPED_B200_H_ */
