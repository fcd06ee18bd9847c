// extern "C" boundary: include/xped_b200.h.  Exceptions are mapped to status codes + xb_last_error().
#include "../../include/xped_b200.h"

#include <algorithm>
#include <limits>
#include <cstring>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>

#include "ad.hpp"
#include "ctm.hpp"
#include "dist.hpp"
#include "engine.hpp"

using namespace xb;

struct xb_ctx
{
    Context impl;
    std::unique_ptr<Tape> user_tape;   // xb_tape_begin .. xb_tape_end
    xb_ctx(int dev, const std::vector<Factor>& f) : impl(dev, f) {}
};
struct xb_basis { xb_ctx* ctx; BasisPtr b; };
struct xb_tensor { xb_ctx* ctx; std::unique_ptr<Tensor> t; };
struct xb_ctm { xb_ctx* ctx; std::unique_ptr<Ctm> c; };
struct xb_dist_plan { xb_ctx* ctx; std::shared_ptr<DistPermutePlan> p; };

namespace {
thread_local std::string g_err;

template <typename F>
int guard(F&& f)
{
    try {
        f();
        return XB_OK;
    } catch(const std::bad_alloc& e) {
        g_err = std::string("out of memory: ") + e.what();
        return XB_ERR_NOMEM;
    } catch(const std::invalid_argument& e) {
        g_err = e.what();
        return XB_ERR_INVALID;
    } catch(const std::out_of_range& e) {
        g_err = e.what();
        return XB_ERR_INVALID;
    } catch(const std::runtime_error& e) {
        g_err = e.what();
        if(g_err.find("no CUDA device") != std::string::npos) return XB_ERR_NO_DEVICE;
        if(g_err.find("out of memory") != std::string::npos) return XB_ERR_NOMEM;
        return XB_ERR_CUDA;
    } catch(const std::exception& e) {
        g_err = e.what();
        return XB_ERR_INTERNAL;
    }
}

xb_tensor* wrap(xb_ctx* ctx, Tensor* t) { return new xb_tensor{ctx, std::unique_ptr<Tensor>(t)}; }

void need(const void* p, const char* what)
{
    if(!p) throw std::invalid_argument(std::string("null argument: ") + what);
}

// In-place writes (upload, fills): storage can be shared (an identity permute returns an alias of its operand; an AD tape keeps recorded values alive), so a
// tensor whose storage has another owner, or any tensor while a tape exists, moves to fresh storage first (full overwrite: no copy; partial overwrite: the
// old contents are copied) -- copy-on-write; under a tape the handle thereby becomes a new, untracked value, like assigning a new tensor.
void detach_for_write(xb_tensor* t, bool keep_contents)
{
    Context& c = t->ctx->impl;
    Tensor& T = *t->t;
    if(c.tape == nullptr && T.buf.use_count() <= 1) return;   // sole owner and no tape: write in place
    const size_t bytes = static_cast<size_t>(std::max<int64_t>(T.layout->size, 1)) * sizeof(double);
    auto fresh = std::make_shared<DevBuf>(&c, bytes, false);
    if(keep_contents) check_cuda(cudaMemcpyAsync(fresh->p, T.d, bytes, cudaMemcpyDeviceToDevice, c.stream), "detach copy");
    T.buf = fresh;
    T.d = fresh->p;
}

Q make_q(const Symmetry& s, const int* q)
{
    Q out{};
    for(int i = 0; i < s.Nq(); ++i) out[i] = q[i];
    return out;
}
}   // namespace

extern "C" {

const char* xb_last_error(void) { return g_err.c_str(); }
const char* xb_version(void) { return "xped_b200 0.1 (sm_100a)"; }

int xb_ctx_create(xb_ctx** ctx, int device, const xb_sym_factor* factors, int n_factors)
{
    return guard([&] {
        need(ctx, "ctx");
        std::vector<Factor> f;
        for(int i = 0; i < n_factors; ++i) {
            if(factors[i].kind < 0 || factors[i].kind > 3) throw std::invalid_argument("unknown symmetry kind");
            if(factors[i].kind == XB_SYM_ZN && factors[i].N < 2) throw std::invalid_argument("ZN needs N >= 2");
            f.push_back({static_cast<Kind>(factors[i].kind), factors[i].N, factors[i].fermionic != 0});
        }
        if(f.empty()) f.push_back({Kind::U0, 0, false});
        if(device < -1) throw std::invalid_argument("device index < -1");
        *ctx = new xb_ctx(device, f);
    });
}

int xb_ctx_create_planning(xb_ctx** ctx, const xb_sym_factor* factors, int n_factors) { return xb_ctx_create(ctx, -1, factors, n_factors); }

int xb_ctx_free(xb_ctx* ctx)
{
    return guard([&] { delete ctx; });
}

int xb_ctx_sync(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        if(ctx->impl.planning_only()) return;
        ctx->impl.sync();
    });
}

int xb_ctx_set_int(xb_ctx* ctx, const char* key, int64_t value)
{
    return guard([&] {
        need(ctx, "ctx");
        need(key, "key");
        const std::string k(key);
        if(k == "svd.truncated")
            ctx->impl.opt_svd_truncated = value < 0 ? -1 : (value != 0 ? 1 : 0);
        else if(k == "svd.call_tag")
            ctx->impl.svd_call_tag = static_cast<int>(value);
        else if(k == "dist.pull_mbps")
            dist_pull_mbps = std::max<int64_t>(value, 1);
        else
            throw std::invalid_argument("xb_ctx_set_int: unknown key '" + k + "'");
    });
}

int xb_ctx_get_int(const xb_ctx* ctx, const char* key, int64_t* value)
{
    return guard([&] {
        need(ctx, "ctx");
        need(key, "key");
        need(value, "value");
        const std::string k(key);
        if(k == "svd.truncated")
            *value = ctx->impl.opt_svd_truncated;
        else if(k == "svd.call_tag")
            *value = ctx->impl.svd_call_tag;
        else if(k == "dist.pull_mbps")
            *value = dist_pull_mbps;
        else if(k == "svd.truncated_calls")
            *value = ctx->impl.svd_truncated_calls;
        else if(k == "svd.truncated_fallbacks")
            *value = ctx->impl.svd_truncated_fallbacks;
        else if(k == "svd.truncated_iterations")
            *value = ctx->impl.svd_truncated_iters;
        else
            throw std::invalid_argument("xb_ctx_get_int: unknown key '" + k + "'");
    });
}

int xb_ctx_nq(const xb_ctx* ctx) { return ctx ? ctx->impl.sym->Nq() : -1; }

int xb_ctx_stats(const xb_ctx* ctx, int64_t out[3])
{
    return guard([&] {
        need(ctx, "ctx");
        out[0] = ctx->impl.launches;
        out[1] = ctx->impl.plan_hits;
        out[2] = ctx->impl.plan_misses;
    });
}

// ---------------------------------------------------------------------------------------------- state interop
int xb_tensor_save(xb_ctx* ctx, const xb_tensor* t, const char* path)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        need(path, "path");
        save_tensor(ctx->impl, *t->t, path);
    });
}

int xb_tensor_load(xb_ctx* ctx, const char* path, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(path, "path");
        need(out, "out");
        *out = wrap(ctx, load_tensor(ctx->impl, path));
    });
}

int xb_ctm_save(const xb_ctm* c, const char* path)
{
    return guard([&] {
        need(c, "ctm");
        need(path, "path");
        save_ctm(c->ctx->impl, *c->c, c->c->chi(), path);
    });
}

int xb_ctm_load(xb_ctx* ctx, const char* path, xb_ctm** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(path, "path");
        need(out, "out");
        *out = new xb_ctm{ctx, std::unique_ptr<Ctm>(load_ctm(ctx->impl, path))};
    });
}

int xb_tensor_from_charge_blocks(xb_ctx* ctx, int nlegs, int rank, const int* ncharge, const double* const* charges, const double* const* sector_dims, const int* conj,
                                 int64_t nblocks, const int64_t* voffs_index, const int64_t* voffs_offset, const double* vdat, int64_t nvdat, int qn_scale,
                                 xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(ncharge, "ncharge");
        need(charges, "charges");
        need(sector_dims, "sector_dims");
        need(out, "out");
        if(nblocks > 0) {
            need(voffs_index, "voffs_index");
            need(voffs_offset, "voffs_offset");
            need(vdat, "vdat");
        }
        *out = wrap(ctx, tensor_from_charge_blocks(ctx->impl, nlegs, rank, ncharge, charges, sector_dims, conj, nblocks, voffs_index, voffs_offset, vdat, nvdat, qn_scale));
    });
}

// ---------------------------------------------------------------------------------------------- reverse-mode AD
int xb_tape_begin(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        if(ctx->impl.tape) throw std::invalid_argument("xb_tape_begin: a tape is already recording");
        ctx->user_tape.reset(new Tape(ctx->impl));
        ctx->impl.tape = ctx->user_tape.get();
    });
}

int xb_tape_end(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        ctx->impl.tape = nullptr;
        ctx->user_tape.reset();
    });
}

int xb_tape_seed_trace(xb_ctx* ctx, const xb_tensor* v, double weight)
{
    return guard([&] {
        need(ctx, "ctx");
        need(v, "tensor");
        if(!ctx->user_tape) throw std::invalid_argument("no tape: call xb_tape_begin first");
        ctx->user_tape->seed_trace(*v->t, weight);
    });
}

int xb_tape_seed(xb_ctx* ctx, const xb_tensor* v, const xb_tensor* adj)
{
    return guard([&] {
        need(ctx, "ctx");
        need(v, "tensor");
        need(adj, "adjoint");
        if(!ctx->user_tape) throw std::invalid_argument("no tape: call xb_tape_begin first");
        if(adj->t->layout->size != v->t->layout->size || adj->t->layout->sector != v->t->layout->sector)
            throw std::invalid_argument("xb_tape_seed: the adjoint must have the block structure of its value");
        ctx->user_tape->seed(*v->t, *adj->t);
    });
}

int xb_tape_backward(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        if(!ctx->user_tape) throw std::invalid_argument("no tape: call xb_tape_begin first");
        ctx->user_tape->backward();
    });
}

int xb_tape_grad(xb_ctx* ctx, const xb_tensor* v, xb_tensor** g)
{
    return guard([&] {
        need(ctx, "ctx");
        need(v, "tensor");
        need(g, "out");
        if(!ctx->user_tape) throw std::invalid_argument("no tape: call xb_tape_begin first");
        Tensor* a = ctx->user_tape->grad(*v->t);
        if(!a) {
            *g = nullptr;
            return;
        }
        TapePause pause(ctx->impl);
        *g = wrap(ctx, ctx->impl.copy(*a));
    });
}

int xb_tape_size(xb_ctx* ctx, int64_t* n_nodes)
{
    return guard([&] {
        need(ctx, "ctx");
        need(n_nodes, "out");
        *n_nodes = ctx->user_tape ? static_cast<int64_t>(ctx->user_tape->size()) : 0;
    });
}

int xb_normalize_max(xb_ctx* ctx, const xb_tensor* t, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        need(out, "out");
        *out = wrap(ctx, ctx->impl.scale_by_inv_maxnorm(*t->t));
    });
}

int xb_normalize_trace(xb_ctx* ctx, const xb_tensor* x, const xb_tensor* w, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(x, "tensor");
        need(w, "weight");
        need(out, "out");
        *out = wrap(ctx, ctx->impl.scale_by_inv_contraction(*x->t, *w->t));
    });
}

int xb_ctm_energy_grad(xb_ctx* ctx, const xb_tensor* A, int64_t chi, int steps, const xb_tensor* op_h, const xb_tensor* op_v, int checkpoint, double* energy,
                       xb_tensor** dA)
{
    return guard([&] {
        need(ctx, "ctx");
        need(A, "A");
        need(op_h, "op_h");
        need(op_v, "op_v");
        need(energy, "energy");
        Tensor* g = nullptr;
        *energy = ctm_energy_grad(ctx->impl, *A->t, chi, steps, *op_h->t, *op_v->t, checkpoint != 0, dA ? &g : nullptr);
        if(dA) *dA = wrap(ctx, g);
    });
}

// ---------------------------------------------------------------------------------------------- bases
int xb_basis_create(xb_ctx* ctx, const int* q, const int64_t* dims, int n, xb_basis** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        auto b = std::make_shared<Qbasis>(ctx->impl.sym, 1);
        const int nq = ctx->impl.sym->Nq();
        for(int i = 0; i < n; ++i) {
            if(dims[i] < 0) throw std::invalid_argument("negative sector dimension");
            b->push_back(make_q(*ctx->impl.sym, q + static_cast<size_t>(i) * nq), dims[i]);
        }
        *out = new xb_basis{ctx, b};
    });
}

int xb_basis_free(xb_basis* b)
{
    return guard([&] { delete b; });
}

int xb_basis_conj(const xb_basis* b, xb_basis** out)
{
    return guard([&] {
        need(b, "basis");
        *out = new xb_basis{b->ctx, b->b->conj()};
    });
}

int xb_basis_sorted(const xb_basis* b, xb_basis** out)
{
    return guard([&] {
        need(b, "basis");
        auto c = std::make_shared<Qbasis>(*b->b);
        c->sort();
        *out = new xb_basis{b->ctx, c};
    });
}

int xb_basis_combine(const xb_basis* a, const xb_basis* b, xb_basis** out)
{
    return guard([&] {
        need(a, "a");
        need(b, "b");
        *out = new xb_basis{a->ctx, a->b->combine(*b->b)->forget_history()};
    });
}

int xb_basis_size(const xb_basis* b, int* n_sectors, int64_t* dim, int64_t* full_dim, int* is_conj)
{
    return guard([&] {
        need(b, "basis");
        if(n_sectors) *n_sectors = static_cast<int>(b->b->entries().size());
        if(dim) *dim = b->b->dim();
        if(full_dim) *full_dim = b->b->full_dim();
        if(is_conj) *is_conj = b->b->is_conj();
    });
}

int xb_basis_get(const xb_basis* b, int* q, int64_t* dims)
{
    return guard([&] {
        need(b, "basis");
        const int nq = b->b->sym()->Nq();
        size_t i = 0;
        for(const auto& e : b->b->entries()) {
            for(int k = 0; k < nq; ++k) q[i * nq + k] = e.q[k];
            dims[i] = e.dim;
            ++i;
        }
    });
}

// ---------------------------------------------------------------------------------------------- tensors
int xb_tensor_create(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        if(rank < 0 || corank < 0 || rank > MAXR || corank > MAXR) throw std::invalid_argument("rank/corank out of range");
        std::vector<BasisPtr> d, c;
        for(int i = 0; i < rank; ++i) {
            need(domain[i], "domain basis");
            d.push_back(domain[i]->b);
        }
        for(int i = 0; i < corank; ++i) {
            need(codomain[i], "codomain basis");
            c.push_back(codomain[i]->b);
        }
        *out = wrap(ctx, new Tensor(&ctx->impl, ctx->impl.layout(d, c), true));
    });
}

int xb_tensor_free(xb_tensor* t)
{
    return guard([&] { delete t; });
}

int xb_tensor_clone(const xb_tensor* t, xb_tensor** out)
{
    return guard([&] {
        need(t, "tensor");
        *out = wrap(t->ctx, t->ctx->impl.copy(*t->t));
    });
}

int xb_tensor_rank(const xb_tensor* t, int* rank, int* corank)
{
    return guard([&] {
        need(t, "tensor");
        *rank = t->t->layout->rank();
        *corank = t->t->layout->corank();
    });
}

int xb_tensor_leg_basis(const xb_tensor* t, int leg, xb_basis** out)
{
    return guard([&] {
        need(t, "tensor");
        const Layout& L = *t->t->layout;
        if(leg < 0 || leg >= L.rank() + L.corank()) throw std::invalid_argument("leg out of range");
        *out = new xb_basis{t->ctx, leg < L.rank() ? L.dom[leg] : L.cod[leg - L.rank()]};
    });
}

int64_t xb_tensor_plain_size(const xb_tensor* t) { return t ? t->t->layout->size : -1; }
int xb_tensor_num_sectors(const xb_tensor* t) { return t ? static_cast<int>(t->t->layout->sector.size()) : -1; }

int xb_tensor_structure(const xb_tensor* t, int* q, int64_t* rows, int64_t* cols, int64_t* offsets)
{
    return guard([&] {
        need(t, "tensor");
        const Layout& L = *t->t->layout;
        const int nq = t->ctx->impl.sym->Nq();
        for(size_t i = 0; i < L.sector.size(); ++i) {
            for(int k = 0; k < nq; ++k) q[i * nq + k] = L.sector[i][k];
            rows[i] = L.rows[i];
            cols[i] = L.cols[i];
            offsets[i] = L.off[i];
        }
    });
}

int64_t xb_tensor_num_tree_pairs(const xb_tensor* t)
{
    if(!t) return -1;
    const Layout& L = *t->t->layout;
    int64_t n = 0;
    for(const Q& q : L.sector) n += static_cast<int64_t>(L.cdom->trees(q).size()) * static_cast<int64_t>(L.ccod->trees(q).size());
    return n;
}

int xb_tensor_tree_pairs(const xb_tensor* t, int64_t* out5)
{
    return guard([&] {
        need(t, "tensor");
        const Layout& L = *t->t->layout;
        size_t n = 0;
        for(size_t i = 0; i < L.sector.size(); ++i) {
            int64_t ro = 0;
            for(const Tree& dt : L.cdom->trees(L.sector[i])) {
                int64_t co = 0;
                for(const Tree& ct : L.ccod->trees(L.sector[i])) {
                    int64_t* o = out5 + 5 * n++;
                    o[0] = static_cast<int64_t>(i);
                    o[1] = ro;
                    o[2] = co;
                    o[3] = dt.dim;
                    o[4] = ct.dim;
                    co += ct.dim;
                }
                ro += dt.dim;
            }
        }
    });
}

// ---- sub-blocks of one tree pair (Tensor::subMatrix / subBlock / view, include/Xped/Core/Tensor.hpp:191-216, src/Core/Tensor.cpp:830-981)
namespace {
struct PairLoc { int sector; int64_t ro, co; const Tree* dt; const Tree* ct; };
PairLoc locate_pair(const Layout& L, int64_t pair)
{
    if(pair < 0) throw std::invalid_argument("tree pair index < 0");
    int64_t n = 0;
    for(size_t i = 0; i < L.sector.size(); ++i) {
        const auto& dts = L.cdom->trees(L.sector[i]);
        const auto& cts = L.ccod->trees(L.sector[i]);
        const int64_t here = static_cast<int64_t>(dts.size()) * static_cast<int64_t>(cts.size());
        if(pair < n + here) {
            const int64_t a = (pair - n) / static_cast<int64_t>(cts.size()), b = (pair - n) % static_cast<int64_t>(cts.size());
            int64_t ro = 0, co = 0;
            for(int64_t k = 0; k < a; ++k) ro += dts[k].dim;
            for(int64_t k = 0; k < b; ++k) co += cts[k].dim;
            return {static_cast<int>(i), ro, co, &dts[a], &cts[b]};
        }
        n += here;
    }
    throw std::invalid_argument("tree pair index out of range");
}
}   // namespace

int xb_tensor_tree_pair_trees(const xb_tensor* t, int64_t pair, int* q_uncoupled, int* q_intermediate, int* q_coupled, int64_t* dims)
{
    return guard([&] {
        need(t, "tensor");
        const Layout& L = *t->t->layout;
        const int nq = t->ctx->impl.sym->Nq();
        const PairLoc loc = locate_pair(L, pair);
        const int R = L.rank(), C = L.corank();
        const int ni_d = R >= 2 ? R - 2 : 0, ni_c = C >= 2 ? C - 2 : 0;
        for(int i = 0; i < R; ++i) {
            if(q_uncoupled) for(int k = 0; k < nq; ++k) q_uncoupled[i * nq + k] = loc.dt->unc[i][k];
            if(dims) dims[i] = loc.dt->dims[i];
        }
        for(int i = 0; i < C; ++i) {
            if(q_uncoupled) for(int k = 0; k < nq; ++k) q_uncoupled[(R + i) * nq + k] = loc.ct->unc[i][k];
            if(dims) dims[R + i] = loc.ct->dims[i];
        }
        if(q_intermediate) {
            for(int i = 0; i < ni_d; ++i) for(int k = 0; k < nq; ++k) q_intermediate[i * nq + k] = loc.dt->inter[i][k];
            for(int i = 0; i < ni_c; ++i) for(int k = 0; k < nq; ++k) q_intermediate[(ni_d + i) * nq + k] = loc.ct->inter[i][k];
        }
        if(q_coupled) for(int k = 0; k < nq; ++k) q_coupled[k] = loc.dt->coupled[k];
    });
}

int64_t xb_tensor_find_tree_pair(const xb_tensor* t, const int* q_uncoupled, const int* q_intermediate, const int* q_coupled)
{
    if(!t || !q_uncoupled || !q_coupled) return -1;
    const Layout& L = *t->t->layout;
    const Symmetry& sym = *t->ctx->impl.sym;
    const int nq = sym.Nq(), R = L.rank(), C = L.corank();
    const int ni_d = R >= 2 ? R - 2 : 0, ni_c = C >= 2 ? C - 2 : 0;
    auto same = [&](const Q& q, const int* p) {
        for(int k = 0; k < nq; ++k)
            if(q[k] != p[k]) return false;
        return true;
    };
    int64_t n = 0;
    for(size_t i = 0; i < L.sector.size(); ++i) {
        const auto& dts = L.cdom->trees(L.sector[i]);
        const auto& cts = L.ccod->trees(L.sector[i]);
        if(same(L.sector[i], q_coupled)) {
            for(size_t a = 0; a < dts.size(); ++a) {
                bool ok = true;
                for(int l = 0; l < R && ok; ++l) ok = same(dts[a].unc[l], q_uncoupled + l * nq);
                for(int l = 0; l < ni_d && ok; ++l) ok = q_intermediate && same(dts[a].inter[l], q_intermediate + l * nq);
                if(!ok) continue;
                for(size_t b = 0; b < cts.size(); ++b) {
                    bool okc = true;
                    for(int l = 0; l < C && okc; ++l) okc = same(cts[b].unc[l], q_uncoupled + (R + l) * nq);
                    for(int l = 0; l < ni_c && okc; ++l) okc = q_intermediate && same(cts[b].inter[l], q_intermediate + (ni_d + l) * nq);
                    if(okc) return n + static_cast<int64_t>(a) * static_cast<int64_t>(cts.size()) + static_cast<int64_t>(b);
                }
            }
        }
        n += static_cast<int64_t>(dts.size()) * static_cast<int64_t>(cts.size());
    }
    return -1;
}

int xb_view(const xb_tensor* t, int64_t pair, double** dev_ptr, int64_t* ld, int64_t* rows, int64_t* cols)
{
    return guard([&] {
        need(t, "tensor");
        need(dev_ptr, "dev_ptr");
        const Layout& L = *t->t->layout;
        const PairLoc loc = locate_pair(L, pair);
        *dev_ptr = t->t->d + L.off[loc.sector] + loc.ro + loc.co * L.rows[loc.sector];
        if(ld) *ld = L.rows[loc.sector];
        if(rows) *rows = loc.dt->dim;
        if(cols) *cols = loc.ct->dim;
    });
}

int xb_sub_block(const xb_tensor* t, int64_t pair, double* host, int64_t n)
{
    return guard([&] {
        need(t, "tensor");
        need(host, "host buffer");
        const Layout& L = *t->t->layout;
        const PairLoc loc = locate_pair(L, pair);
        const int64_t r = loc.dt->dim, c = loc.ct->dim;
        if(n != r * c) throw std::invalid_argument("sub_block: n != f1.dim * f2.dim");
        if(n == 0) return;
        Context& ctx = t->ctx->impl;
        const double* src = t->t->d + L.off[loc.sector] + loc.ro + loc.co * L.rows[loc.sector];
        check_cuda(cudaMemcpy2DAsync(host, sizeof(double) * r, src, sizeof(double) * L.rows[loc.sector], sizeof(double) * r, static_cast<size_t>(c),
                                     cudaMemcpyDeviceToHost, ctx.stream), "sub_block download");
        check_cuda(cudaStreamSynchronize(ctx.stream), "sub_block sync");
    });
}

int xb_set_sub_block(xb_tensor* t, int64_t pair, const double* host, int64_t n)
{
    return guard([&] {
        need(t, "tensor");
        detach_for_write(t, true);
        need(host, "host buffer");
        const Layout& L = *t->t->layout;
        const PairLoc loc = locate_pair(L, pair);
        const int64_t r = loc.dt->dim, c = loc.ct->dim;
        if(n != r * c) throw std::invalid_argument("set_sub_block: n != f1.dim * f2.dim");
        if(n == 0) return;
        Context& ctx = t->ctx->impl;
        double* dst = t->t->d + L.off[loc.sector] + loc.ro + loc.co * L.rows[loc.sector];
        check_cuda(cudaMemcpy2DAsync(dst, sizeof(double) * L.rows[loc.sector], host, sizeof(double) * r, sizeof(double) * r, static_cast<size_t>(c),
                                     cudaMemcpyHostToDevice, ctx.stream), "sub_block upload");
        check_cuda(cudaStreamSynchronize(ctx.stream), "sub_block sync");   // the caller's buffer is pageable: do not return before it was read
    });
}

int xb_tensor_upload(xb_tensor* t, const double* host, int64_t n)
{
    return guard([&] {
        need(t, "tensor");
        detach_for_write(t, false);
        if(n != t->t->layout->size) throw std::invalid_argument("upload: size != plainSize()");
        Context& c = t->ctx->impl;
        check_cuda(cudaMemcpyAsync(t->t->d, host, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream), "upload");
        c.sync();
    });
}

int xb_tensor_download(const xb_tensor* t, double* host, int64_t n)
{
    return guard([&] {
        need(t, "tensor");
        if(n != t->t->layout->size) throw std::invalid_argument("download: size != plainSize()");
        Context& c = t->ctx->impl;
        check_cuda(cudaMemcpyAsync(host, t->t->d, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream), "download");
        c.sync();
    });
}

int xb_tensor_set_zero(xb_tensor* t)
{
    return guard([&] {
        need(t, "tensor");
        detach_for_write(t, false);
        check_cuda(cudaMemsetAsync(t->t->d, 0, sizeof(double) * t->t->layout->size, t->ctx->impl.stream), "set_zero");
    });
}

int xb_tensor_set_identity(xb_tensor* t)
{
    return guard([&] {
        need(t, "tensor");
        detach_for_write(t, false);
        t->ctx->impl.fill(*t->t, 1, 0.0, 0);
    });
}

int xb_tensor_set_constant(xb_tensor* t, double v)
{
    return guard([&] {
        need(t, "tensor");
        detach_for_write(t, false);
        t->ctx->impl.fill(*t->t, 0, v, 0);
    });
}

int xb_tensor_set_random(xb_tensor* t, uint64_t seed)
{
    return guard([&] {
        need(t, "tensor");
        detach_for_write(t, false);
        t->ctx->impl.fill(*t->t, 2, 0.0, seed);
    });
}


// ---------------------------------------------------------------------------------------------- sector sharding
namespace {
void collect_bases(xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, std::vector<BasisPtr>& d, std::vector<BasisPtr>& c)
{
    if(rank < 0 || corank < 0 || rank > MAXR || corank > MAXR) throw std::invalid_argument("rank/corank out of range");
    for(int i = 0; i < rank; ++i) {
        need(domain[i], "domain basis");
        d.push_back(domain[i]->b);
    }
    for(int i = 0; i < corank; ++i) {
        need(codomain[i], "codomain basis");
        c.push_back(codomain[i]->b);
    }
}
}   // namespace

int xb_layout_structure(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int* n_sectors, int* q, int64_t* rows,
                        int64_t* cols, int64_t* offsets)
{
    return guard([&] {
        need(ctx, "ctx");
        need(n_sectors, "n_sectors");
        std::vector<BasisPtr> d, c;
        collect_bases(domain, rank, codomain, corank, d, c);
        LayoutPtr L = ctx->impl.layout(d, c);
        *n_sectors = static_cast<int>(L->sector.size());
        if(!q) return;
        const int nq = ctx->impl.sym->Nq();
        for(size_t i = 0; i < L->sector.size(); ++i) {
            for(int k = 0; k < nq; ++k) q[i * nq + k] = L->sector[i][k];
            rows[i] = L->rows[i];
            cols[i] = L->cols[i];
            offsets[i] = L->off[i];
        }
    });
}

int xb_permute_plan_stats(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n, int64_t out[6])
{
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        std::vector<BasisPtr> d, c;
        collect_bases(domain, rank, codomain, corank, d, c);
        LayoutPtr L = ctx->impl.layout(d, c);
        auto plan = ctx->impl.permute_plan(L, shift, std::vector<int>(p, p + n), false);
        out[0] = plan->n_tasks;
        out[1] = plan->n_tiles;
        out[2] = plan->max_fan_in;
        out[3] = plan->identity_copy ? 1 : 0;
        out[4] = plan->n_btiles;   // > 0: the plan also has the cp.async.bulk form
        out[5] = plan->bytes;
    });
}

int xb_permute_plan_census(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n, int64_t out[6])
{
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        std::vector<BasisPtr> d, c;
        collect_bases(domain, rank, codomain, corank, d, c);
        LayoutPtr L = ctx->impl.layout(d, c);
        auto plan = ctx->impl.permute_plan(L, shift, std::vector<int>(p, p + n), false);
        out[0] = plan->dst_elems;
        out[1] = plan->src_reads;
        out[2] = plan->n_tasks;
        out[3] = plan->n_srcs;
        out[4] = plan->n_ltiles;        // > 0: the plan has the table-driven form
        out[5] = plan->lut_table_bytes;
    });
}

int xb_permute_plan_emulate(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n, int adjoint,
                            int form, const double* src, int64_t n_src, double* dst, int64_t n_dst)
{
    return guard([&] {
        need(ctx, "ctx");
        need(src, "src");
        need(dst, "dst");
        if(!ctx->impl.planning_only()) throw std::invalid_argument("xb_permute_plan_emulate: test support of planning-only contexts (the device path is xb_permute)");
        std::vector<BasisPtr> d, c;
        collect_bases(domain, rank, codomain, corank, d, c);
        LayoutPtr L = ctx->impl.layout(d, c);
        auto plan = ctx->impl.permute_plan(L, shift, std::vector<int>(p, p + n), adjoint != 0);
        if(n_src != L->size || n_dst != plan->out->size) throw std::invalid_argument("xb_permute_plan_emulate: array sizes do not match the tensors");
        ctx->impl.emulate_permute_plan(*plan, form, src, n_src, dst, n_dst);
    });
}

int xb_multiply_plan_emulate(xb_ctx* ctx, xb_basis* const* a_domain, int a_rank, xb_basis* const* a_codomain, int a_corank, int adj_a, xb_basis* const* b_domain,
                             int b_rank, xb_basis* const* b_codomain, int b_corank, int adj_b, int sm_count, const double* a, int64_t n_a, const double* b, int64_t n_b,
                             double* c, int64_t n_c, int64_t counts[4])
{
    return guard([&] {
        need(ctx, "ctx");
        need(a, "a");
        need(b, "b");
        need(c, "c");
        if(!ctx->impl.planning_only()) throw std::invalid_argument("xb_multiply_plan_emulate: test support of planning-only contexts (the device path is xb_multiply)");
        std::vector<BasisPtr> ad, ac, bd, bc;
        collect_bases(a_domain, a_rank, a_codomain, a_corank, ad, ac);
        collect_bases(b_domain, b_rank, b_codomain, b_corank, bd, bc);
        LayoutPtr LA = ctx->impl.layout(ad, ac), LB = ctx->impl.layout(bd, bc);
        LayoutPtr LO = ctx->impl.layout(adj_a ? ac : ad, adj_b ? bd : bc);
        if(n_a != LA->size || n_b != LB->size || n_c != LO->size) throw std::invalid_argument("xb_multiply_plan_emulate: array sizes do not match the tensors");
        const GemmLists G = build_gemm_lists(multiply_problems(*LA, adj_a != 0, *LB, adj_b != 0, *LO), sm_count > 0 ? sm_count : ctx->impl.sm_count);
        std::fill(c, c + n_c, 0.0);   // Tout.setZero(): sectors missing in an operand stay zero blocks
        emulate_gemm_lists(G, a, n_a, b, n_b, c, n_c);
        if(counts) {
            counts[0] = static_cast<int64_t>(G.items[0].size());
            counts[1] = static_cast<int64_t>(G.items[1].size());
            counts[2] = static_cast<int64_t>(G.jobs.size());
            counts[3] = G.scratch_elems;
        }
    });
}

int xb_contract_plan_emulate(xb_ctx* ctx, xb_basis* const* a_domain, int a_rank, xb_basis* const* a_codomain, int a_corank, const int* labels_a,
                             xb_basis* const* b_domain, int b_rank, xb_basis* const* b_codomain, int b_corank, const int* labels_b, int res_rank, int table_form,
                             const double* a, int64_t n_a, const double* b, int64_t n_b, double* c, int64_t n_c)
{
    return guard([&] {
        need(ctx, "ctx");
        need(a, "a");
        need(b, "b");
        need(c, "c");
        Context& C = ctx->impl;
        if(!C.planning_only()) throw std::invalid_argument("xb_contract_plan_emulate: test support of planning-only contexts (the device path is xb_contract)");
        std::vector<BasisPtr> ad, ac, bd, bc;
        collect_bases(a_domain, a_rank, a_codomain, a_corank, ad, ac);
        collect_bases(b_domain, b_rank, b_codomain, b_corank, bd, bc);
        LayoutPtr LA = C.layout(ad, ac), LB = C.layout(bd, bc);
        if(n_a != LA->size || n_b != LB->size) throw std::invalid_argument("xb_contract_plan_emulate: operand sizes do not match the tensors");
        const ContractPerms cp = contract_permutations(a_rank, a_corank, std::vector<int>(labels_a, labels_a + a_rank + a_corank), b_rank, b_corank,
                                                       std::vector<int>(labels_b, labels_b + b_rank + b_corank), res_rank);
        auto walk = [&](const LayoutPtr& L, int shift, const std::vector<int>& p, const double* src, int64_t n_src, LayoutPtr& out_layout) {
            auto plan = C.permute_plan(L, shift, p, false);
            out_layout = plan->out;
            std::vector<double> dst(static_cast<size_t>(plan->out->size), std::numeric_limits<double>::quiet_NaN());
            const bool has_tables = plan->host && !plan->host->ltile_task.empty();
            C.emulate_permute_plan(*plan, (table_form != 0 && has_tables) ? 1 : 0, src, n_src, dst.data(), static_cast<int64_t>(dst.size()));
            return dst;
        };
        LayoutPtr L1, L2, L3;
        const std::vector<double> ap = walk(LA, cp.shift1, cp.p1, a, n_a, L1);
        const std::vector<double> bp = walk(LB, cp.shift2, cp.p2, b, n_b, L2);
        LayoutPtr LO = C.layout(L1->dom, L2->cod);
        std::vector<double> prod(static_cast<size_t>(LO->size), 0.0);
        emulate_gemm_lists(build_gemm_lists(multiply_problems(*L1, false, *L2, false, *LO), C.sm_count), ap.data(), static_cast<int64_t>(ap.size()), bp.data(),
                           static_cast<int64_t>(bp.size()), prod.data(), static_cast<int64_t>(prod.size()));
        const std::vector<double> res = walk(LO, cp.shiftres, cp.pres, prod.data(), static_cast<int64_t>(prod.size()), L3);
        if(n_c != static_cast<int64_t>(res.size())) throw std::invalid_argument("xb_contract_plan_emulate: result size does not match the contraction");
        std::copy(res.begin(), res.end(), c);
    });
}

int xb_tensor_create_sectors(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, const int* q, int n_sectors,
                             xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        std::vector<BasisPtr> d, c;
        collect_bases(domain, rank, codomain, corank, d, c);
        std::vector<Q> qs;
        for(int i = 0; i < n_sectors; ++i) qs.push_back(make_q(*ctx->impl.sym, q + i * ctx->impl.sym->Nq()));
        *out = wrap(ctx, new Tensor(&ctx->impl, ctx->impl.layout_custom(d, c, qs), true));
    });
}

int xb_tensor_set_random_at(xb_tensor* t, uint64_t seed, const int64_t* global_offsets)
{
    return guard([&] {
        need(t, "tensor");
        detach_for_write(t, false);
        need(global_offsets, "global_offsets");
        Context& c = t->ctx->impl;
        const Layout& L = *t->t->layout;
        for(size_t i = 0; i < L.sector.size(); ++i) {
            launch_fill_random(t->t->d + L.off[i], seed, L.rows[i] * L.cols[i], c.stream, global_offsets[i]);
            ++c.launches;
        }
    });
}

int xb_dist_owner_lpt(const double* cost, int n, int world, int* owner)
{
    return guard([&] {
        need(owner, "owner");
        std::vector<double> v(cost, cost + (n > 0 ? n : 0));
        std::vector<int> o = lpt_owner(v, world);
        std::copy(o.begin(), o.end(), owner);
    });
}

int xb_dist_permute_plan(xb_ctx* ctx, xb_basis* const* domain, int rank, xb_basis* const* codomain, int corank, int shift, const int* p, int n,
                         int adjoint, const int* owner_in, const int* owner_out, int n_out_sectors, int my_rank, int world, xb_dist_plan** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        need(owner_in, "owner_in");
        std::vector<BasisPtr> d, c;
        collect_bases(domain, rank, codomain, corank, d, c);
        LayoutPtr L = ctx->impl.layout(d, c);
        std::vector<int> oi(owner_in, owner_in + L->sector.size()), oo;
        if(owner_out) oo.assign(owner_out, owner_out + n_out_sectors);
        *out = new xb_dist_plan{ctx, ctx->impl.dist_permute_plan(L, shift, std::vector<int>(p, p + n), adjoint != 0, oi, oo, my_rank, world)};
    });
}

int xb_dist_plan_free(xb_dist_plan* plan)
{
    return guard([&] { delete plan; });
}

int xb_dist_plan_counts(const xb_dist_plan* plan, int64_t* send_counts, int64_t* recv_counts)
{
    return guard([&] {
        need(plan, "plan");
        std::copy(plan->p->send_counts.begin(), plan->p->send_counts.end(), send_counts);
        std::copy(plan->p->recv_counts.begin(), plan->p->recv_counts.end(), recv_counts);
    });
}

int xb_dist_plan_out(const xb_dist_plan* plan, int* n_out_sectors, int* owner_out, xb_basis** legs, int* out_rank, int* out_corank)
{
    return guard([&] {
        need(plan, "plan");
        const DistPermutePlan& P = *plan->p;
        if(n_out_sectors) *n_out_sectors = static_cast<int>(P.owner_out.size());
        if(owner_out) std::copy(P.owner_out.begin(), P.owner_out.end(), owner_out);
        const Layout& L = *P.out_global;
        if(out_rank) *out_rank = L.rank();
        if(out_corank) *out_corank = L.corank();
        if(legs)
            for(int i = 0; i < L.rank() + L.corank(); ++i) legs[i] = new xb_basis{plan->ctx, i < L.rank() ? L.dom[i] : L.cod[i - L.rank()]};
    });
}

int xb_dist_plan_local_sectors(const xb_dist_plan* plan, int which, int* n_sectors, int* q)
{
    return guard([&] {
        need(plan, "plan");
        need(n_sectors, "n_sectors");
        const Layout& L = which == 0 ? *plan->p->in_local : *plan->p->out_local;
        *n_sectors = static_cast<int>(L.sector.size());
        if(!q) return;
        const int nq = plan->ctx->impl.sym->Nq();
        for(size_t i = 0; i < L.sector.size(); ++i)
            for(int k = 0; k < nq; ++k) q[i * nq + k] = L.sector[i][k];
    });
}

int xb_dist_plan_records(const xb_dist_plan* plan, int which, int64_t* n_records, int64_t* out)
{
    return guard([&] {
        need(plan, "plan");
        need(n_records, "n_records");
        const DistPermutePlan& P = *plan->p;
        const std::vector<int64_t>& v = which == 0 ? P.pack_records : which == 1 ? P.task_records : P.src_records;
        const int width = which == 0 ? DistPermutePlan::PACK_REC : which == 1 ? DistPermutePlan::TASK_REC : DistPermutePlan::SRC_REC;
        if(which < 0 || which > 2) throw std::invalid_argument("xb_dist_plan_records: which must be 0 (pack), 1 (tasks) or 2 (sources)");
        *n_records = static_cast<int64_t>(v.size()) / width;
        if(out) std::copy(v.begin(), v.end(), out);
    });
}

int xb_dist_permute_pack(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, double* d_send)
{
    return guard([&] {
        need(ctx, "ctx");
        need(plan, "plan");
        need(local_in, "local_in");
        ctx->impl.dist_pack(*plan->p, *local_in->t, d_send);
    });
}

int xb_dist_permute_local(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, xb_tensor** local_out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(plan, "plan");
        need(local_in, "local_in");
        need(local_out, "local_out");
        *local_out = wrap(ctx, ctx->impl.dist_local(*plan->p, *local_in->t));
    });
}

int xb_dist_permute_remote(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, const double* d_recv, xb_tensor* local_out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(plan, "plan");
        need(local_in, "local_in");
        need(local_out, "local_out");
        ctx->impl.dist_remote(*plan->p, *local_in->t, d_recv, *local_out->t);
    });
}

int xb_dist_shm_create(const char* shm_name, int world)
{
    return guard([&] {
        need(shm_name, "shm_name");
        dist_shm_create(shm_name, world);
    });
}

int xb_dist_init(xb_ctx* ctx, int my_rank, int world, int64_t arena_bytes, const char* shm_name)
{
    return guard([&] {
        need(ctx, "ctx");
        need(shm_name, "shm_name");
        dist_init(ctx->impl, my_rank, world, static_cast<size_t>(arena_bytes), shm_name);
    });
}

int xb_dist_shutdown(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        ctx->impl.sync();
        dist_shutdown(ctx->impl);
    });
}

int xb_dist_barrier(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        dist_barrier(ctx->impl);
    });
}

int xb_dist_arena_stats(const xb_ctx* ctx, int64_t out[3])
{
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        dist_arena_stats(ctx->impl, out);
    });
}

int xb_dist_plan_pushable(const xb_dist_plan* plan, int* pushable)
{
    return guard([&] {
        need(plan, "plan");
        need(pushable, "pushable");
        *pushable = plan->p->push.empty() ? 0 : 1;
    });
}

int xb_dist_permute_push(xb_ctx* ctx, const xb_dist_plan* plan, const xb_tensor* local_in, xb_tensor** local_out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(plan, "plan");
        need(local_in, "local_in");
        need(local_out, "local_out");
        *local_out = wrap(ctx, ctx->impl.dist_permute_push(*plan->p, *local_in->t));
    });
}

int xb_dist_multiply(xb_ctx* ctx, const xb_tensor* a_local, const int* owner_a, const xb_tensor* b_local, const int* owner_b, int panel_rows_min, xb_tensor** c_local,
                     int* owner_out, int n_owner_out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(a_local, "a_local");
        need(b_local, "b_local");
        need(owner_a, "owner_a");
        need(owner_b, "owner_b");
        need(c_local, "c_local");
        Context& c = ctx->impl;
        const Layout& LA = *a_local->t->layout;
        const Layout& LB = *b_local->t->layout;
        const size_t na = c.layout(LA.dom, LA.cod)->sector.size(), nb = c.layout(LB.dom, LB.cod)->sector.size();
        std::vector<int> oo;
        *c_local = wrap(ctx, c.dist_multiply(*a_local->t, std::vector<int>(owner_a, owner_a + na), *b_local->t, std::vector<int>(owner_b, owner_b + nb),
                                             panel_rows_min < 2 ? 256 : panel_rows_min, &oo));
        if(owner_out) {
            if(n_owner_out != static_cast<int>(oo.size())) throw std::invalid_argument("xb_dist_multiply: owner_out needs one entry per sector of the result");
            std::copy(oo.begin(), oo.end(), owner_out);
        }
    });
}

int xb_dist_multiply_assignment(const int64_t* m, const int64_t* n, const int64_t* k, const int* owner, int n_sectors, int world, int panel_rows_min,
                                int64_t* n_panels, int64_t* out)
{
    return guard([&] {
        need(n_panels, "n_panels");
        if(world < 1 || world > DIST_MAX_WORLD) throw std::invalid_argument("xb_dist_multiply_assignment: world must be 1..8");
        std::vector<int64_t> vm(m, m + n_sectors), vn(n, n + n_sectors), vk(k, k + n_sectors);
        std::vector<int> vo(owner, owner + n_sectors);
        for(int o : vo)
            if(o < 0 || o >= world) throw std::invalid_argument("xb_dist_multiply_assignment: owner out of range");
        const auto panels = dist_multiply_assignment(vm, vn, vk, vo, world, panel_rows_min < 2 ? 256 : panel_rows_min, 0.02);
        if(out) {
            if(*n_panels < static_cast<int64_t>(panels.size())) throw std::invalid_argument("xb_dist_multiply_assignment: output too small");
            for(size_t i = 0; i < panels.size(); ++i) {
                out[4 * i + 0] = panels[i].sector;
                out[4 * i + 1] = panels[i].rank;
                out[4 * i + 2] = panels[i].row0;
                out[4 * i + 3] = panels[i].rows;
            }
        }
        *n_panels = static_cast<int64_t>(panels.size());
    });
}

int xb_multiply_sectors(xb_ctx* ctx, const xb_tensor* a, int adj_a, const xb_tensor* b, int adj_b, const int* q, int n_sectors, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(a, "a");
        need(b, "b");
        need(out, "out");
        std::vector<Q> qs;
        for(int i = 0; i < n_sectors; ++i) qs.push_back(make_q(*ctx->impl.sym, q + i * ctx->impl.sym->Nq()));
        *out = wrap(ctx, ctx->impl.multiply_sectors(*a->t, adj_a != 0, *b->t, adj_b != 0, &qs));
    });
}

// ---------------------------------------------------------------------------------------------- hot path
int xb_permute(xb_ctx* ctx, const xb_tensor* t, int shift, const int* p, int n, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        *out = wrap(ctx, ctx->impl.permute(*t->t, shift, std::vector<int>(p, p + n)));
    });
}

int xb_permute_adj(xb_ctx* ctx, const xb_tensor* t, int shift, const int* p, int n, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        *out = wrap(ctx, ctx->impl.permute_adj(*t->t, shift, std::vector<int>(p, p + n)));
    });
}

int xb_multiply(xb_ctx* ctx, const xb_tensor* a, int adj_a, const xb_tensor* b, int adj_b, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(a, "a");
        need(b, "b");
        *out = wrap(ctx, ctx->impl.multiply(*a->t, adj_a != 0, *b->t, adj_b != 0));
    });
}

int xb_contract(xb_ctx* ctx, const xb_tensor* a, const int* la, int na, const xb_tensor* b, const int* lb, int nb, int res_rank, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(a, "a");
        need(b, "b");
        *out = wrap(ctx, ctx->impl.contract(*a->t, std::vector<int>(la, la + na), *b->t, std::vector<int>(lb, lb + nb), res_rank));
    });
}

int xb_adjoint(xb_ctx* ctx, const xb_tensor* t, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        *out = wrap(ctx, ctx->impl.adjoint(*t->t));
    });
}

int xb_twist(xb_ctx* ctx, const xb_tensor* t, int leg, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        *out = wrap(ctx, ctx->impl.twist(*t->t, leg));
    });
}

int xb_axpby(xb_ctx* ctx, double alpha, const xb_tensor* x, double beta, const xb_tensor* y, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(x, "x");
        *out = wrap(ctx, ctx->impl.axpby(alpha, *x->t, beta, y ? y->t.get() : nullptr));
    });
}

int xb_diag_map(xb_ctx* ctx, const xb_tensor* t, int op, xb_tensor** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        if(op < 0 || op > 2) throw std::invalid_argument("diag_map: op must be 0, 1 or 2");
        *out = wrap(ctx, ctx->impl.diag_map(*t->t, op));
    });
}

int xb_max_norm(xb_ctx* ctx, const xb_tensor* t, double* out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        *out = ctx->impl.max_norm(*t->t);
    });
}

int xb_trace(xb_ctx* ctx, const xb_tensor* t, double* out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        *out = ctx->impl.trace(*t->t);
    });
}

int xb_squared_norm(xb_ctx* ctx, const xb_tensor* t, double* out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        *out = ctx->impl.squared_norm(*t->t);
    });
}

int xb_tsvd(xb_ctx* ctx, const xb_tensor* t, int64_t max_keep, double eps_svd, int preserve_multiplets, xb_tensor** U, xb_tensor** S, xb_tensor** Vdag,
            double* trunc_weight, double* entropy)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        SvdResult r = ctx->impl.tsvd(*t->t, max_keep, eps_svd, preserve_multiplets != 0);
        *U = wrap(ctx, r.U);
        *S = wrap(ctx, r.S);
        *Vdag = wrap(ctx, r.Vd);
        if(trunc_weight) *trunc_weight += r.trunc_weight;
        if(entropy) *entropy += r.entropy;
    });
}

int xb_tsvd_backward(xb_ctx* ctx, const xb_tensor* A, const xb_tensor* U, const xb_tensor* S, const xb_tensor* Vdag, const xb_tensor* dU,
                     const xb_tensor* dS, const xb_tensor* dVdag, xb_tensor** dA)
{
    return guard([&] {
        need(ctx, "ctx");
        need(A, "A");
        need(U, "U");
        need(S, "S");
        need(Vdag, "Vdag");
        *dA = wrap(ctx, ctx->impl.tsvd_backward(*A->t, *U->t, *S->t, *Vdag->t, dU ? dU->t.get() : nullptr, dS ? dS->t.get() : nullptr,
                                                dVdag ? dVdag->t.get() : nullptr));
    });
}

int xb_tqr(xb_ctx* ctx, const xb_tensor* t, int return_lq, xb_tensor** Qo, xb_tensor** Ro)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        Tensor *q = nullptr, *r = nullptr;
        ctx->impl.tqr(*t->t, return_lq != 0, &q, &r);
        *Qo = wrap(ctx, q);
        *Ro = wrap(ctx, r);
    });
}

int xb_teigh(xb_ctx* ctx, const xb_tensor* t, xb_tensor** D, xb_tensor** V)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        Tensor *d = nullptr, *v = nullptr;
        ctx->impl.teigh(*t->t, &d, &v);
        *D = wrap(ctx, d);
        *V = wrap(ctx, v);
    });
}

int xb_decompose(xb_ctx* ctx, const xb_tensor* T1, const xb_tensor* T2, int64_t chi, xb_tensor** P1, xb_tensor** P2)
{
    return guard([&] {
        need(ctx, "ctx");
        need(T1, "T1");
        need(T2, "T2");
        Tensor *p1 = nullptr, *p2 = nullptr;
        decompose_projectors(ctx->impl, *T1->t, *T2->t, chi, &p1, &p2);
        *P1 = wrap(ctx, p1);
        *P2 = wrap(ctx, p2);
    });
}

int xb_contract_L(xb_ctx* ctx, const xb_tensor* Bold, const xb_tensor* Bra, const xb_tensor* Ket, xb_tensor** Bnew)
{
    return guard([&] {
        need(ctx, "ctx");
        need(Bold, "Bold");
        need(Bra, "Bra");
        need(Ket, "Ket");
        Tensor* o = nullptr;
        ctx->impl.contract_L(*Bold->t, *Bra->t, *Ket->t, &o);
        *Bnew = wrap(ctx, o);
    });
}

int xb_mps_left_sweep_step(xb_ctx* ctx, const xb_tensor* A_loc, const xb_tensor* A_prev, int64_t max_nsv, double eps_svd, xb_tensor** A_loc_new,
                           xb_tensor** A_nb_new, double* trunc_weight, double* entropy)
{
    return guard([&] {
        need(ctx, "ctx");
        need(A_loc, "A_loc");
        Tensor *a = nullptr, *b = nullptr;
        ctx->impl.mps_left_sweep_step(*A_loc->t, A_prev ? A_prev->t.get() : nullptr, max_nsv, eps_svd, &a, &b, trunc_weight, entropy);
        *A_loc_new = wrap(ctx, a);
        *A_nb_new = b ? wrap(ctx, b) : nullptr;
    });
}

int xb_mps_right_sweep_step(xb_ctx* ctx, const xb_tensor* A_loc, const xb_tensor* A_next, int64_t max_nsv, double eps_svd, xb_tensor** A_loc_new,
                            xb_tensor** A_nb_new, double* trunc_weight, double* entropy)
{
    return guard([&] {
        need(ctx, "ctx");
        need(A_loc, "A_loc");
        Tensor *a = nullptr, *b = nullptr;
        ctx->impl.mps_right_sweep_step(*A_loc->t, A_next ? A_next->t.get() : nullptr, max_nsv, eps_svd, &a, &b, trunc_weight, entropy);
        *A_loc_new = wrap(ctx, a);
        *A_nb_new = b ? wrap(ctx, b) : nullptr;
    });
}

int xb_contract_R(xb_ctx* ctx, const xb_tensor* Bold, const xb_tensor* Bra, const xb_tensor* Ket, xb_tensor** Bnew)
{
    return guard([&] {
        need(ctx, "ctx");
        need(Bold, "Bold");
        need(Bra, "Bra");
        need(Ket, "Ket");
        Tensor* o = nullptr;
        ctx->impl.contract_R(*Bold->t, *Bra->t, *Ket->t, &o);
        *Bnew = wrap(ctx, o);
    });
}

// ---------------------------------------------------------------------------------------------- CTM
int xb_ctm_create(xb_ctx* ctx, const xb_tensor* A, int64_t chi, xb_ctm** out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(A, "A");
        *out = new xb_ctm{ctx, std::make_unique<Ctm>(ctx->impl, *A->t, chi)};
    });
}

int xb_ctm_free(xb_ctm* c)
{
    return guard([&] { delete c; });
}

int xb_ctm_step(xb_ctm* c)
{
    return guard([&] {
        need(c, "ctm");
        c->c->grow_all();
    });
}

int xb_ctm_get(const xb_ctm* c, int which, xb_tensor** out)
{
    return guard([&] {
        need(c, "ctm");
        *out = wrap(c->ctx, c->ctx->impl.copy(c->c->get(which)));
    });
}

int xb_ctm_set(xb_ctm* c, int which, const xb_tensor* t)
{
    return guard([&] {
        need(c, "ctm");
        need(t, "tensor");
        c->c->set(which, c->ctx->impl.copy(*t->t));
    });
}

int xb_ctm_corner(xb_ctm* c, int corner, xb_tensor** Qo)
{
    return guard([&] {
        need(c, "ctm");
        *Qo = wrap(c->ctx, c->c->contract_corner(corner));
    });
}

int xb_ctm_rdm(xb_ctm* c, int which, xb_tensor** rho)
{
    return guard([&] {
        need(c, "ctm");
        *rho = wrap(c->ctx, c->c->rdm(which));
    });
}

int xb_rdm_one_site(xb_ctx* ctx, const xb_tensor* rho, xb_tensor** rho1)
{
    return guard([&] {
        need(ctx, "ctx");
        need(rho, "rho");
        *rho1 = wrap(ctx, rdm_one_site(ctx->impl, *rho->t));
    });
}

int xb_avg_one_site(xb_ctx* ctx, const xb_tensor* rho1, const xb_tensor* op, double* out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(rho1, "rho1");
        need(op, "op");
        *out = avg_one_site(ctx->impl, *rho1->t, *op->t);
    });
}

int xb_avg_two_site(xb_ctx* ctx, const xb_tensor* rho, const xb_tensor* op, double* out)
{
    return guard([&] {
        need(ctx, "ctx");
        need(rho, "rho");
        need(op, "op");
        *out = avg_two_site(ctx->impl, *rho->t, *op->t);
    });
}

int xb_ctm_move(xb_ctm* c, int direction)
{
    return guard([&] {
        need(c, "ctm");
        c->c->move(direction);
    });
}

// ---------------------------------------------------------------------------------------------- bench hooks
int xb_bench_permute(xb_ctx* ctx, const xb_tensor* t, int shift, const int* p, int n, int reps, double* ms_per_launch, int64_t* bytes)
{
    return guard([&] {
        need(ctx, "ctx");
        need(t, "tensor");
        Context& c = ctx->impl;
        auto plan = c.permute_plan(t->t->layout, shift, std::vector<int>(p, p + n), false);
        Tensor out(&c, plan->out, false);
        cudaEvent_t e0, e1;
        check_cuda(cudaEventCreate(&e0), "event");
        check_cuda(cudaEventCreate(&e1), "event");
        check_cuda(cudaEventRecord(e0, c.stream), "record");
        for(int r = 0; r < reps; ++r) {
            if(plan->identity_copy)
                check_cuda(cudaMemcpyAsync(out.d, t->t->d, sizeof(double) * t->t->size(), cudaMemcpyDeviceToDevice, c.stream), "copy");
            else
                launch_permute(t->t->d, out.d, *plan, c.stream);
        }
        c.launches += reps;
        check_cuda(cudaEventRecord(e1, c.stream), "record");
        check_cuda(cudaEventSynchronize(e1), "event sync");
        float ms = 0.f;
        check_cuda(cudaEventElapsedTime(&ms, e0, e1), "elapsed");
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms_per_launch = static_cast<double>(ms) / reps;
        *bytes = plan->bytes;
    });
}

int xb_bench_multiply(xb_ctx* ctx, const xb_tensor* a, const xb_tensor* b, int reps, double* ms_per_launch, double* flops)
{
    return guard([&] {
        need(ctx, "ctx");
        need(a, "a");
        need(b, "b");
        Context& c = ctx->impl;
        const Layout& LA = *a->t->layout;
        const Layout& LB = *b->t->layout;
        double fl = 0.0;
        for(size_t i = 0; i < LA.sector.size(); ++i) {
            const int j = LB.find(LA.sector[i]);
            if(j >= 0) fl += 2.0 * static_cast<double>(LA.rows[i]) * static_cast<double>(LA.cols[i]) * static_cast<double>(LB.cols[j]);
        }
        std::unique_ptr<Tensor> warm(c.multiply(*a->t, false, *b->t, false));
        cudaEvent_t e0, e1;
        check_cuda(cudaEventCreate(&e0), "event");
        check_cuda(cudaEventCreate(&e1), "event");
        check_cuda(cudaEventRecord(e0, c.stream), "record");
        for(int r = 0; r < reps; ++r) std::unique_ptr<Tensor> o(c.multiply(*a->t, false, *b->t, false));
        check_cuda(cudaEventRecord(e1, c.stream), "record");
        check_cuda(cudaEventSynchronize(e1), "event sync");
        float ms = 0.f;
        check_cuda(cudaEventElapsedTime(&ms, e0, e1), "elapsed");
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        c.sync();
        *ms_per_launch = static_cast<double>(ms) / reps;
        *flops = fl;
    });
}

int xb_flush_l2(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        Context& c = ctx->impl;
        const size_t bytes = size_t(256) << 20;   // > 126 MB L2
        void* p = c.scratch(bytes);
        check_cuda(cudaMemsetAsync(p, 0, bytes, c.stream), "flush_l2");
    });
}

void* xb_ctx_stream(xb_ctx* ctx) { return ctx ? static_cast<void*>(ctx->impl.stream) : nullptr; }

int xb_timer_start(xb_ctx* ctx)
{
    return guard([&] {
        need(ctx, "ctx");
        ctx->impl.timer_start();
    });
}

int xb_timer_stop(xb_ctx* ctx, double* ms)
{
    return guard([&] {
        need(ctx, "ctx");
        need(ms, "ms");
        *ms = ctx->impl.timer_stop();
    });
}

int xb_profile_enable(xb_ctx* ctx, int on)
{
    return guard([&] {
        need(ctx, "ctx");
        ctx->impl.profiling = on != 0;
    });
}

int xb_profile_read(xb_ctx* ctx, double ms[4], int64_t count[4], double work[4])
{
    return guard([&] {
        need(ctx, "ctx");
        ctx->impl.prof_read(ms, count, work);
    });
}

}   // extern "C"
