// Sector-sharded operations over peer memory (plumbing in dist.cu).
//
//  dist_permute_push   Tensor::permute (src/Core/Tensor.cpp:83-214) on shards: the exchange is fused into the permute launches -- every rank transposes the
//                      sub-blocks it OWNS straight into the output shard of the destination's owner over NVLink (one launch per destination rank).
//  dist_multiply       TensorBase::operator* (src/Core/TensorBase.cpp:243-288) on shards with the heaviest sectors split into ROW PANELS: sectors are
//                      independent, but a handful of them carry most of the flops (SU(2): 4 of 9 sectors = 80 %; W4: the Q = 0 block = 22 %), so whole-sector
//                      ownership cannot balance 8 GPUs.  A panel C_q[r0:r1, :] = A_q[r0:r1, :] * B_q can run anywhere: the computing rank pulls the A panel and
//                      B_q from the owner's shard (peer-mapped, bulk copies) and the GEMM epilogue stores the C panel into the owner's output shard.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <stdexcept>

#include "dist.hpp"
#include "engine.hpp"

namespace xb {

namespace {
struct ArenaForce   // tensors created in this scope must be addressable by the peers
{
    Context& c;
    bool saved;
    explicit ArenaForce(Context& ctx) : c(ctx), saved(ctx.arena_force) { c.arena_force = true; }
    ~ArenaForce() { c.arena_force = saved; }
};
}   // namespace

Tensor* Context::dist_permute_push(const DistPermutePlan& plan, const Tensor& local_in)
{
    if(!dist) throw std::invalid_argument("dist_permute_push: context has no distributed state (xb_dist_init)");
    if(plan.push.empty()) throw std::invalid_argument("dist_permute_push: plan has recoupling sums (fan-in > 1); use the pack / exchange / gather form");
    if(local_in.layout.get() != plan.in_local.get()) throw std::invalid_argument("dist_permute_push: tensor is not this rank's input shard of the plan");
    if(plan.world != dist_world(*this) || plan.rank != dist_rank(*this)) throw std::invalid_argument("dist_permute_push: plan was built for another rank / world");
    Tensor* out = nullptr;
    {
        ArenaForce af(*this);
        out = new Tensor(this, plan.out_local, plan.push_zero_fill);
    }
    std::unique_ptr<Tensor> guard_out(out);
    int64_t mine[DIST_SLOT_WORDS] = {out->buf->arena_off, 0, 0, 0}, all[DIST_MAX_WORLD][DIST_SLOT_WORDS];
    dist_exchange(*this, mine, all);
    ProfScope prof(*this, CAT_PERMUTE, static_cast<double>(plan.bytes_local));
    dist_barrier(*this);   // A: every output shard is allocated (and zeroed where needed); what lived there before is no longer in use by its owner
    // ONE launch: every tile is transposed in shared memory and leaves for the shard of its destination's owner (this rank's or a peer's, over NVLink)
    PeerDelta delta;
    for(int p = 0; p < 8; ++p) delta.d[p] = 0;
    for(int p = 0; p < plan.world; ++p)
        if(p != plan.rank) delta.d[p] = dist_peer_ptr(*this, p, all[p][0]) - out->d;
    if(plan.push[0]->n_tiles > 0) {
        launch_permute_peers(local_in.d, out->d, *plan.push[0], delta, stream);
        ++launches;
    }
    dist_barrier(*this);   // B: all peers have finished storing into this rank's shard
    return guard_out.release();
}

// ------------------------------------------------------------------------------------------------ row-panel product
int64_t dist_pull_mbps = 600000;   // modelled rate of pulling operand panels from a peer (MB/s; xb_ctx_set_int "dist.pull_mbps"): weighs transfers against flops
std::vector<DistPanel> dist_multiply_assignment(const std::vector<int64_t>& m, const std::vector<int64_t>& n, const std::vector<int64_t>& k, const std::vector<int>& owner,
                                                int world, int panel_rows_min, double imbalance_tol)
{
    // deterministic: every rank computes the same list.  Loads are TIMES (ms): product flops at the grouped-GEMM rate plus, for a panel computed away from
    // its owner, the pull of the A panel and (once per rank and sector) of B_q at the peer-copy rate.  Sectors in descending flops; a sector stays whole on
    // its owner while the owner has room (load <= cap); otherwise it is cut into row panels (multiples of `panel_rows_min` rows) that fill the
    // least-loaded ranks up to the cap; the cap itself is the smallest one the greedy can honour (bisection).
    const double GEMM_FLOP_PER_MS = 29.0e9, PULL_BYTES_PER_MS = 1.0e3 * static_cast<double>(dist_pull_mbps);
    const int ns = static_cast<int>(m.size());
    std::vector<double> fl(ns);
    double total = 0.0;
    for(int i = 0; i < ns; ++i) {
        fl[i] = 2.0 * static_cast<double>(m[i]) * static_cast<double>(n[i]) * static_cast<double>(k[i]) / GEMM_FLOP_PER_MS;
        total += fl[i];
    }
    std::vector<int> order(ns);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fl[a] > fl[b]; });
    auto greedy = [&](double cap, std::vector<DistPanel>& out) {
        out.clear();
        std::vector<double> load(world, 0.0);
        std::vector<char> has_b(static_cast<size_t>(world) * ns, 0);
        for(int s : order) {
            if(m[s] == 0 || n[s] == 0) continue;
            const double comp_row = fl[s] / static_cast<double>(m[s]);
            const double pull_row = 8.0 * static_cast<double>(k[s]) / PULL_BYTES_PER_MS, pull_b = 8.0 * static_cast<double>(k[s]) * static_cast<double>(n[s]) / PULL_BYTES_PER_MS;
            int64_t r0 = 0;
            while(r0 < m[s]) {
                const int64_t left = m[s] - r0;
                int r = owner[s];   // first choice: the owner (no transfer); else the least-loaded rank
                if(load[r] + comp_row * static_cast<double>(std::min<int64_t>(left, panel_rows_min)) > cap) {
                    r = 0;
                    for(int q = 1; q < world; ++q)
                        if(load[q] < load[r]) r = q;
                }
                const bool away = r != owner[s];
                const double fixed = (away && !has_b[static_cast<size_t>(r) * ns + s]) ? pull_b : 0.0, per_row = comp_row + (away ? pull_row : 0.0);
                int64_t rows = left;
                const double room = cap - load[r] - fixed;
                if(per_row * static_cast<double>(left) > room) {
                    rows = room > 0.0 ? static_cast<int64_t>(room / per_row) / panel_rows_min * panel_rows_min : 0;
                    if(rows < panel_rows_min) rows = std::min<int64_t>(left, panel_rows_min);   // everybody is full: overshoot by one minimal panel
                    if(left - rows < panel_rows_min / 2) rows = left;                           // no sliver at the end
                }
                if(away) has_b[static_cast<size_t>(r) * ns + s] = 1;
                out.push_back({s, r, r0, rows});
                load[r] += fixed + per_row * static_cast<double>(rows);
                r0 += rows;
            }
        }
        double worst = 0.0;
        for(double l : load) worst = std::max(worst, l);
        return worst;
    };
    // smallest cap the greedy can honour (a larger one leaves the last ranks short): bisection between the transfer-free mean and twice that
    std::vector<DistPanel> best, trial;
    double lo = total / world, hi = 2.0 * lo + 1e-12, best_worst = greedy(hi, best);
    for(int it = 0; it < 24 && hi - lo > imbalance_tol * 0.25 * lo; ++it) {
        const double mid = 0.5 * (lo + hi), w = greedy(mid, trial);
        if(w <= mid * (1.0 + 1e-9)) {
            hi = mid;
            if(w <= best_worst) {
                best_worst = w;
                best.swap(trial);
            }
        } else {
            lo = mid;
        }
    }
    return best;
}

Tensor* Context::dist_multiply(const Tensor& a, const std::vector<int>& owner_a, const Tensor& b, const std::vector<int>& owner_b, int panel_rows_min,
                               std::vector<int>* owner_out)
{
    if(!dist) throw std::invalid_argument("dist_multiply: context has no distributed state (xb_dist_init)");
    const int rank = dist_rank(*this), world = dist_world(*this);
    LayoutPtr GA = layout(a.layout->dom, a.layout->cod), GB = layout(b.layout->dom, b.layout->cod);
    if(owner_a.size() != GA->sector.size() || owner_b.size() != GB->sector.size()) throw std::invalid_argument("dist_multiply: owner maps do not match the global sector lists");
    if(GA->cod.size() != GB->dom.size()) throw std::invalid_argument("operator*: CoRank of left operand != Rank of right operand");
    for(size_t i = 0; i < GA->cod.size(); ++i)
        if(!GA->cod[i]->same_sectors(*GB->dom[i])) throw std::invalid_argument("operator*: coupledCodomain(left) != coupledDomain(right)");
    LayoutPtr GO = layout(GA->dom, GB->cod);
    // owner of an output sector = owner of the left operand's sector (operands must agree on the common sectors)
    const int no = static_cast<int>(GO->sector.size());
    std::vector<int> own_o(no, 0), ia(no, -1), ib(no, -1);
    std::vector<int64_t> pm(no, 0), pn(no, 0), pk(no, 0);
    for(int o = 0; o < no; ++o) {
        ia[o] = GA->find(GO->sector[o]);
        ib[o] = GB->find(GO->sector[o]);
        own_o[o] = ia[o] >= 0 ? owner_a[ia[o]] : (ib[o] >= 0 ? owner_b[ib[o]] : o % world);
        if(ia[o] >= 0 && ib[o] >= 0) {
            if(owner_a[ia[o]] != owner_b[ib[o]]) throw std::invalid_argument("dist_multiply: operands disagree on the owner of a common sector (redistribute one of them first)");
            pm[o] = GA->rows[ia[o]];
            pk[o] = GA->cols[ia[o]];
            pn[o] = GB->cols[ib[o]];
        }
    }
    if(owner_out) *owner_out = own_o;
    // shard layouts and the offset of every global sector inside its owner's shard
    auto shard_offsets = [&](const LayoutPtr& G, const std::vector<int>& owner, std::vector<int64_t>& off, LayoutPtr* mine) {
        off.assign(G->sector.size(), 0);
        std::vector<int64_t> acc(world, 0);
        std::vector<Q> my_q;
        for(size_t i = 0; i < G->sector.size(); ++i) {
            off[i] = acc[owner[i]];
            acc[owner[i]] += G->rows[i] * G->cols[i];
            if(owner[i] == rank) my_q.push_back(G->sector[i]);
        }
        if(mine) *mine = layout_custom(G->dom, G->cod, my_q);
    };
    std::vector<int64_t> offa, offb, offo;
    LayoutPtr LAm, LBm, LOm;
    shard_offsets(GA, owner_a, offa, &LAm);
    shard_offsets(GB, owner_b, offb, &LBm);
    shard_offsets(GO, own_o, offo, &LOm);
    if(a.layout.get() != LAm.get() || b.layout.get() != LBm.get()) throw std::invalid_argument("dist_multiply: operands are not this rank's shards under the given owner maps");
    for(size_t i = 0; i < LAm->sector.size(); ++i)   // the shard layout is the concatenation assumed above
        if(LAm->off[i] != offa[GA->find(LAm->sector[i])]) throw std::logic_error("dist_multiply: shard offsets are not consecutive");

    const std::vector<DistPanel> panels = dist_multiply_assignment(pm, pn, pk, own_o, world, panel_rows_min, 0.02);
    // operands that peers read must sit in the arena
    std::unique_ptr<Tensor> a_copy, b_copy;
    const Tensor* A = &a;
    const Tensor* B = &b;
    bool anybody_pulls = false;
    for(const DistPanel& p : panels)
        if(p.rank != own_o[p.sector]) anybody_pulls = true;
    if(anybody_pulls) {
        ArenaForce af(*this);
        if(a.buf->arena_off < 0) {
            a_copy.reset(copy(a));
            A = a_copy.get();
        }
        if(b.buf->arena_off < 0) {
            b_copy.reset(&a == &b ? A->alias() : copy(b));
            B = b_copy.get();
        }
    }
    Tensor* out = nullptr;
    {
        ArenaForce af(*this);
        out = new Tensor(this, LOm, true);   // Tout.setZero(): sectors missing in an operand stay zero blocks
    }
    std::unique_ptr<Tensor> guard_out(out);
    int64_t mine[DIST_SLOT_WORDS] = {A->buf->arena_off, B->buf->arena_off, out->buf->arena_off, 0}, all[DIST_MAX_WORLD][DIST_SLOT_WORDS];
    dist_exchange(*this, mine, all);
    dist_barrier(*this);   // A: operands of every rank are complete, outputs allocated and zeroed

    // ---- what this rank computes but does not own is pulled on a side stream while the panels of its own sectors run; then the staged panels
    std::vector<GemmItem> items_own, items_staged;
    std::vector<double*> stage;
    std::vector<int> staged_b(no, -1);   // index into `stage` of the copy of B_q
    cudaStream_t side = nullptr;
    struct Pull { double* dst; const double* src; int64_t rows, ld, cols; };
    std::vector<Pull> pulls;
    for(const DistPanel& p : panels) {
        if(p.rank != rank) continue;
        const int o = p.sector, ow = own_o[o];
        const int64_t m = pm[o], n = pn[o], k = pk[o];
        GemmItem g;
        std::memset(&g, 0, sizeof(g));
        g.m = static_cast<int>(p.rows);
        g.n = static_cast<int>(n);
        g.k = static_cast<int>(k);
        g.alpha = 1.0;
        g.ldc = static_cast<int>(m);
        if(ow == rank) {
            g.a_off = offa[ia[o]] + p.row0;
            g.lda = static_cast<int>(m);
            g.b_off = offb[ib[o]];
            g.ldb = static_cast<int>(k);
            g.c_off = offo[o] + p.row0;
            items_own.push_back(g);
            continue;
        }
        // staging memory comes from the stream-ordered pool of the context stream; the copies are issued on the side stream after a fork event that
        // follows all allocations, and the memory is released on the context stream after the join
        double* sa = nullptr;
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&sa), sizeof(double) * static_cast<size_t>(p.rows) * k, stream), "cudaMallocAsync(panel staging)");
        stage.push_back(sa);
        const double* pa = dist_peer_ptr(*this, ow, all[ow][0]) + offa[ia[o]] + p.row0;
        pulls.push_back({sa, pa, p.rows, m, k});
        if(staged_b[o] < 0) {
            double* sb = nullptr;
            check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&sb), sizeof(double) * static_cast<size_t>(k) * n, stream), "cudaMallocAsync(B staging)");
            staged_b[o] = static_cast<int>(stage.size());
            stage.push_back(sb);
            const double* pb = dist_peer_ptr(*this, ow, all[ow][1]) + offb[ib[o]];
            pulls.push_back({sb, pb, k * n, k * n, 1});
        }
        g.a_off = sa - A->d;
        g.lda = static_cast<int>(p.rows);
        g.b_off = stage[staged_b[o]] - B->d;
        g.ldb = static_cast<int>(k);
        g.c_off = (dist_peer_ptr(*this, ow, all[ow][2]) + offo[o] + p.row0) - out->d;
        items_staged.push_back(g);
    }
    if(!pulls.empty()) {
        side = dist_side_stream(*this, 0);
        cudaEvent_t fork = get_event();
        check_cuda(cudaEventRecord(fork, stream), "cudaEventRecord(fork)");
        check_cuda(cudaStreamWaitEvent(side, fork, 0), "cudaStreamWaitEvent(fork)");
        event_pool.push_back(fork);
        for(const Pull& c : pulls)   // column-major panel: `cols` runs of `rows` doubles, source pitch `ld`
            check_cuda(cudaMemcpy2DAsync(c.dst, sizeof(double) * c.rows, c.src, sizeof(double) * c.ld, sizeof(double) * c.rows, static_cast<size_t>(c.cols), cudaMemcpyDeviceToDevice, side),
                       "cudaMemcpy2DAsync(panel from its owner)");
    }
    if(!items_own.empty()) run_gemm(A->d, B->d, out->d, items_own);
    if(side) {
        cudaEvent_t join = get_event();
        check_cuda(cudaEventRecord(join, side), "cudaEventRecord(join)");
        check_cuda(cudaStreamWaitEvent(stream, join, 0), "cudaStreamWaitEvent(join)");
        event_pool.push_back(join);
        run_gemm(A->d, B->d, out->d, items_staged);
    }
    for(double* s : stage) check_cuda(cudaFreeAsync(s, stream), "cudaFreeAsync(staging)");
    dist_barrier(*this);   // B: every panel of this rank's sectors has arrived; operands may be released
    return guard_out.release();
}

}   // namespace xb
