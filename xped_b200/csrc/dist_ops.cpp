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
    for(int k = 0; k < plan.world; ++k) {
        const int p = (plan.rank + k) % plan.world;   // start with the local part, then round the peers (staggered targets across ranks)
        const PermutePlan& pp = *plan.push[p];
        if(pp.n_tiles == 0) continue;
        double* dst = p == plan.rank ? out->d : dist_peer_ptr(*this, p, all[p][0]);
        launch_permute(local_in.d, dst, pp, stream);
        ++launches;
    }
    dist_barrier(*this);   // B: all peers have finished storing into this rank's shard
    return guard_out.release();
}

// ------------------------------------------------------------------------------------------------ row-panel product
std::vector<DistPanel> dist_multiply_assignment(const std::vector<int64_t>& m, const std::vector<int64_t>& n, const std::vector<int64_t>& k, const std::vector<int>& owner,
                                                int world, int panel_rows_min, double imbalance_tol)
{
    // deterministic: every rank computes the same list.  Sectors in descending flops; a sector stays whole on its owner while the owner has room
    // (load <= mean * (1 + tol)); otherwise it is cut into row panels (multiples of `panel_rows_min` rows) that fill the least-loaded ranks up to the mean.
    const int ns = static_cast<int>(m.size());
    std::vector<double> fl(ns);
    double total = 0.0;
    for(int i = 0; i < ns; ++i) {
        fl[i] = 2.0 * static_cast<double>(m[i]) * static_cast<double>(n[i]) * static_cast<double>(k[i]);
        total += fl[i];
    }
    const double mean = total / world, cap = mean * (1.0 + imbalance_tol);
    std::vector<int> order(ns);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fl[a] > fl[b]; });
    std::vector<double> load(world, 0.0);
    std::vector<DistPanel> out;
    for(int s : order) {
        if(m[s] == 0 || n[s] == 0) continue;
        const double per_row = fl[s] / static_cast<double>(m[s]);
        int64_t r0 = 0;
        while(r0 < m[s]) {
            const int64_t left = m[s] - r0;
            // first choice: the owner (no transfer); else the least-loaded rank
            int r = owner[s];
            if(load[r] + per_row * std::min<int64_t>(left, panel_rows_min) > cap) {
                r = 0;
                for(int q = 1; q < world; ++q)
                    if(load[q] < load[r]) r = q;
            }
            int64_t rows = left;
            const double room = cap - load[r];
            if(per_row * static_cast<double>(left) > room) {
                rows = static_cast<int64_t>(room / per_row) / panel_rows_min * panel_rows_min;
                if(rows < panel_rows_min) rows = std::min<int64_t>(left, panel_rows_min);   // everybody is full: overshoot by one minimal panel
                if(left - rows < panel_rows_min / 2) rows = left;                           // no sliver at the end
            }
            out.push_back({s, r, r0, rows});
            load[r] += per_row * static_cast<double>(rows);
            r0 += rows;
        }
    }
    return out;
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

    // ---- stage what this rank computes but does not own, then ONE grouped product over local and staged panels
    std::vector<GemmItem> items;
    std::vector<double*> stage;
    std::vector<int> staged_b(no, -1);   // index into `stage` of the copy of B_q
    double flops_mine = 0.0;
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
        flops_mine += 2.0 * p.rows * static_cast<double>(n) * k;
        if(ow == rank) {
            g.a_off = offa[ia[o]] + p.row0;
            g.lda = static_cast<int>(m);
            g.b_off = offb[ib[o]];
            g.ldb = static_cast<int>(k);
            g.c_off = offo[o] + p.row0;
        } else {
            double* sa = nullptr;
            check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&sa), sizeof(double) * static_cast<size_t>(p.rows) * k, stream), "cudaMallocAsync(panel staging)");
            stage.push_back(sa);
            const double* pa = dist_peer_ptr(*this, ow, all[ow][0]) + offa[ia[o]] + p.row0;
            check_cuda(cudaMemcpy2DAsync(sa, sizeof(double) * p.rows, pa, sizeof(double) * m, sizeof(double) * p.rows, static_cast<size_t>(k), cudaMemcpyDeviceToDevice, stream),
                       "cudaMemcpy2DAsync(A panel from its owner)");
            if(staged_b[o] < 0) {
                double* sb = nullptr;
                check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&sb), sizeof(double) * static_cast<size_t>(k) * n, stream), "cudaMallocAsync(B staging)");
                staged_b[o] = static_cast<int>(stage.size());
                stage.push_back(sb);
                const double* pb = dist_peer_ptr(*this, ow, all[ow][1]) + offb[ib[o]];
                check_cuda(cudaMemcpyAsync(sb, pb, sizeof(double) * static_cast<size_t>(k) * n, cudaMemcpyDeviceToDevice, stream), "cudaMemcpyAsync(B block from its owner)");
            }
            g.a_off = sa - A->d;
            g.lda = static_cast<int>(p.rows);
            g.b_off = stage[staged_b[o]] - B->d;
            g.ldb = static_cast<int>(k);
            g.c_off = (dist_peer_ptr(*this, ow, all[ow][2]) + offo[o] + p.row0) - out->d;
        }
        items.push_back(g);
    }
    run_gemm(A->d, B->d, out->d, items);
    for(double* s : stage) check_cuda(cudaFreeAsync(s, stream), "cudaFreeAsync(staging)");
    dist_barrier(*this);   // B: every panel of this rank's sectors has arrived; operands may be released
    (void)flops_mine;
    return guard_out.release();
}

}   // namespace xb
