// Whole-tensor operations of the B200 backend (the methods of Xped::Tensor that sit on the hot path).
#include <algorithm>
#include <limits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>

#include "ad.hpp"
#include "engine.hpp"

namespace xb {

// ------------------------------------------------------------------------------------------------ permute
Tensor* Context::permute(const Tensor& t, int shift, const std::vector<int>& pi)
{
    auto plan = permute_plan(t.layout, shift, pi, false);
    if(plan->identity_copy && plan->out.get() == t.layout.get()) {
        // permute<0, identity> (short-circuited copy in the reference, Tensor.cpp:344): tensors are immutable values here (writers through the C ABI
        // detach shared storage first), so the result shares the operand's storage -- no launch, no traffic, and the same value on an AD tape
        return t.alias();
    }
    auto* out = new Tensor(this, plan->out, false);
    ProfScope prof(*this, CAT_PERMUTE, static_cast<double>(plan->bytes));
    if(plan->identity_copy) {
        check_cuda(cudaMemcpyAsync(out->d, t.d, sizeof(double) * t.size(), cudaMemcpyDeviceToDevice, stream), "permute identity copy");
    } else {
        launch_permute(t.d, out->d, *plan, stream);
    }
    ++launches;
    if(taping()) {   // ADTensor.hpp:171-185: adj(t) += adj(out).permute_adj<shift, p>()
        Tensor *tin = tape->keep(t), *tout = tape->keep(*out);
        tape->record([this, tin, tout, shift, pi](Tape& T) {
            if(Tensor* g = T.grad(*tout)) T.accumulate(*tin, permute_adj(*g, shift, pi));
        });
    }
    return out;
}

Tensor* Context::permute_adj(const Tensor& t, int shift, const std::vector<int>& pi)
{
    auto plan = permute_plan(t.layout, shift, pi, true);
    auto* out = new Tensor(this, plan->out, false);
    ProfScope prof(*this, CAT_PERMUTE, static_cast<double>(plan->bytes));
    if(plan->identity_copy) {
        check_cuda(cudaMemcpyAsync(out->d, t.d, sizeof(double) * t.size(), cudaMemcpyDeviceToDevice, stream), "permute_adj identity copy");
    } else {
        launch_permute(t.d, out->d, *plan, stream);
    }
    ++launches;
    return out;
}

Tensor* Context::copy(const Tensor& t)
{
    auto* out = new Tensor(this, t.layout, false);
    check_cuda(cudaMemcpyAsync(out->d, t.d, sizeof(double) * t.size(), cudaMemcpyDeviceToDevice, stream), "copy");
    if(taping()) {
        Tensor *tin = tape->keep(t), *tout = tape->keep(*out);
        tape->record([this, tin, tout](Tape& T) {
            if(Tensor* g = T.grad(*tout)) T.accumulate(*tin, copy(*g));
        });
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ grouped GEMM launcher
// Host part of prepare_gemm: tile lists of the two kernel variants, split-K jobs and the size of their scratch buffer.  A pure function of the problem list and
// the SM count -- no device involved, so that the CPU tests can walk the lists element by element (emulate_gemm_lists).  Scratch destinations are marked
// pad_ = 1 with c_off relative to the scratch buffer; prepare_gemm rebases them onto the allocation.
GemmLists build_gemm_lists(const std::vector<GemmItem>& problems, int sm_count)
{
    GemmLists G;
    std::vector<GemmItem>(&items)[2] = G.items;   // [0]: 128 x 64 tiles (sectors with m > 96), [1]: 64 x 64 tiles
    std::vector<SplitKJob>& jobs = G.jobs;        // problems with few tiles and a long K: K is cut into slices, the partial products are summed afterwards
    long long& scratch_elems = G.scratch_elems;
    static const bool no_split = std::getenv("XB_GEMM_NOSPLIT") != nullptr;
    // (measured in round 2 and removed again: summing the slices inside gemm_grouped_kernel by the last slice CTA of a tile (ticket counters) -- 32.2 vs
    // 26.3 ms per D=4 CTM step, and the extra epilogue raised the 128-row kernel from 168 to 207 registers: 2 instead of 3 CTAs per SM, W4 T*T 23.8 vs 29.1
    // TFLOP/s)
    // tiles of the whole launch: a launch that fills less than ~two waves (3 CTAs per SM) with LONG contractions leaves SMs idle for a large fraction of its
    // run time (the 3060 x 160 x 3096 products of the D=6 subspace iteration: 333 CTAs on 148 SMs = 2.25 per SM, 10 TFLOP/s) -- cut K there as well
    // Small launches (the products of a D=4 subspace iteration: 136 CTAs of 128 x 64 tiles, ONE CTA = four warps per SM: tensor pipe 31 % busy, the warps wait on
    // their own DMMA chains, 32 us for 0.15 GFLOP) use 64-row tiles for every problem: twice the CTAs, two CTAs per SM.
    static const bool no_small64 = std::getenv("XB_GEMM_SMALL64") != nullptr && std::atoi(std::getenv("XB_GEMM_SMALL64")) == 0;
    bool force64 = false;
    if(!no_small64) {
        long long ctas = 0;
        for(const GemmItem& g : problems) {
            if(g.m <= 0 || g.n <= 0) continue;
            const int tile_m = g.m > 96 ? 128 : 64;
            const long long ntiles = static_cast<long long>((g.m + tile_m - 1) / tile_m) * ((g.n + GEMM_TILE_N - 1) / GEMM_TILE_N);
            long long split = 1;
            if(!no_split && !g.acc && g.k >= 256 && ntiles <= 32) split = std::max<long long>(1, std::min<long long>(8, std::min<long long>(g.k / 96, sm_count / ntiles)));
            ctas += ntiles * split;
        }
        force64 = ctas < 2ll * sm_count;
    }
    long long launch_tiles = 0;
    for(const GemmItem& g : problems) {
        if(g.m <= 0 || g.n <= 0) continue;
        const int tile_m = (g.m > 96 && !force64) ? 128 : 64;
        launch_tiles += static_cast<long long>((g.m + tile_m - 1) / tile_m) * ((g.n + GEMM_TILE_N - 1) / GEMM_TILE_N);
    }
    static const int wave_split = std::getenv("XB_GEMM_WAVE_SPLIT") ? std::atoi(std::getenv("XB_GEMM_WAVE_SPLIT")) : 6;   // target CTAs per SM (0: off)
    for(GemmItem g : problems) {
        if(g.m <= 0 || g.n <= 0) continue;
        G.flops += 2.0 * g.m * static_cast<double>(g.n) * g.k;
        const int which = (g.m > 96 && !force64) ? 0 : 1, tile_m = which == 0 ? 128 : 64;
        const int ntiles = ((g.m + tile_m - 1) / tile_m) * ((g.n + GEMM_TILE_N - 1) / GEMM_TILE_N);
        int split = 1;
        if(!no_split && !g.acc && g.k >= 256 && ntiles <= 32) split = std::max(1, std::min(8, std::min(g.k / 96, sm_count / ntiles)));
        if(!no_split && wave_split > 0 && !g.acc && split == 1 && g.k >= 768 && launch_tiles < static_cast<long long>(wave_split) * sm_count)
            split = std::max(1, std::min(8, std::min(g.k / 256, static_cast<int>((static_cast<long long>(wave_split) * sm_count + launch_tiles - 1) / launch_tiles))));
        if(split > 1) {
            const int kchunk = ((((g.k + split - 1) / split) + 7) / 8) * 8;
            SplitKJob job;
            job.c_off = g.c_off;
            job.s_off = scratch_elems;
            job.m = g.m;
            job.n = g.n;
            job.ldc = g.ldc;
            job.split = 0;
            for(int k0 = 0; k0 < g.k; k0 += kchunk) {
                GemmItem h = g;
                h.k = std::min(kchunk, g.k - k0);
                h.a_off += g.ta ? k0 : static_cast<long long>(k0) * g.lda;
                h.b_off += g.tb ? static_cast<long long>(k0) * g.ldb : k0;
                h.c_off = scratch_elems + static_cast<long long>(job.split) * g.m * g.n;   // relative to the scratch buffer; rebased below
                h.ldc = g.m;
                h.pad_ = 1;   // marks a scratch destination
                ++job.split;
                for(int tn = 0; tn < g.n; tn += GEMM_TILE_N)
                    for(int tm = 0; tm < g.m; tm += tile_m) {
                        h.tm = tm;
                        h.tn = tn;
                        items[which].push_back(h);
                    }
            }
            scratch_elems += static_cast<long long>(job.split) * g.m * g.n;
            jobs.push_back(job);
            continue;
        }
        g.pad_ = 0;
        for(int tn = 0; tn < g.n; tn += GEMM_TILE_N)
            for(int tm = 0; tm < g.m; tm += tile_m) {
                g.tm = tm;
                g.tn = tn;
                items[which].push_back(g);
            }
    }
    return G;
}

PreparedGemm Context::prepare_gemm(const double* A, const double* B, double* C, const std::vector<GemmItem>& problems)
{
    PreparedGemm P;
    P.A = A;
    P.B = B;
    P.C = C;
    GemmLists G = build_gemm_lists(problems, sm_count);
    std::vector<GemmItem>(&items)[2] = G.items;
    std::vector<SplitKJob>& jobs = G.jobs;
    const long long scratch_elems = G.scratch_elems;
    P.flops = G.flops;
    if(items[0].empty() && items[1].empty()) return P;
    static const bool gemm_log = std::getenv("XB_GEMM_LOG") != nullptr;
    if(gemm_log) {
        int mm = 0, mn = 0, mk = 0;
        for(const GemmItem& g : problems) {
            mm = std::max(mm, g.m);
            mn = std::max(mn, g.n);
            mk = std::max(mk, g.k);
        }
        std::fprintf(stderr, "[gemm] problems %zu flops %.3e max(m,n,k) = (%d,%d,%d) tiles128 %zu tiles64 %zu split-K jobs %zu\n", problems.size(), P.flops, mm, mn, mk,
                     items[0].size(), items[1].size(), jobs.size());
    }
    long long rebase = 0;
    if(scratch_elems > 0) {
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&P.d_scratch), sizeof(double) * scratch_elems, stream), "cudaMallocAsync split-K scratch");
        rebase = (reinterpret_cast<intptr_t>(P.d_scratch) - reinterpret_cast<intptr_t>(C)) / static_cast<intptr_t>(sizeof(double));   // both 256-byte aligned
    }
    for(int which = 0; which < 2; ++which) {
        auto& v = items[which];
        if(v.empty()) continue;
        for(GemmItem& it : v)
            if(it.pad_) {
                it.c_off += rebase;
                it.pad_ = 0;
            }
        // heavy tiles first: k is the per-tile cost
        std::stable_sort(v.begin(), v.end(), [](const GemmItem& x, const GemmItem& y) { return x.k > y.k; });
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&P.d_items[which]), v.size() * sizeof(GemmItem), stream), "cudaMallocAsync gemm items");
        upload_async(P.d_items[which], v.data(), v.size() * sizeof(GemmItem));
        P.n_items[which] = static_cast<int>(v.size());
        ++P.launches;
    }
    if(!jobs.empty()) {
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&P.d_jobs), jobs.size() * sizeof(SplitKJob), stream), "cudaMallocAsync split-K jobs");
        upload_async(P.d_jobs, jobs.data(), jobs.size() * sizeof(SplitKJob));
        P.n_jobs = static_cast<int>(jobs.size());
        for(const SplitKJob& j : jobs) P.max_job_elems = std::max(P.max_job_elems, j.m * j.n);
        ++P.launches;
    }
    return P;
}

void Context::launch_gemm(const PreparedGemm& P)
{
    if(P.launches == 0) return;
    ProfScope prof(*this, CAT_GEMM, P.flops);
    for(int which = 0; which < 2; ++which)
        if(P.n_items[which] > 0) launch_gemm_grouped(P.A, P.B, P.C, P.d_items[which], P.n_items[which], which == 0 ? 128 : 64, stream);
    if(P.n_jobs > 0) launch_splitk_reduce(P.d_scratch, P.C, P.d_jobs, P.n_jobs, P.max_job_elems, stream);
    launches += P.launches;
}

void Context::free_gemm(PreparedGemm& P)
{
    for(int which = 0; which < 2; ++which)
        if(P.d_items[which]) check_cuda(cudaFreeAsync(P.d_items[which], stream), "cudaFreeAsync gemm items");
    if(P.d_jobs) check_cuda(cudaFreeAsync(P.d_jobs, stream), "cudaFreeAsync split-K jobs");
    if(P.d_scratch) check_cuda(cudaFreeAsync(P.d_scratch, stream), "cudaFreeAsync split-K scratch");
    P = PreparedGemm();
}

void Context::run_gemm(const double* A, const double* B, double* C, const std::vector<GemmItem>& problems)
{
    PreparedGemm P = prepare_gemm(A, B, C, problems);
    launch_gemm(P);
    free_gemm(P);
}

// ------------------------------------------------------------------------------------------------ operator*
// TensorBase::operator* (src/Core/TensorBase.cpp:243-288). Operands may be lazy adjoints (AdjointOp.hpp:26-69):
// the block is then read transposed, domain/codomain swap roles, the sector list stays the operand's.
// Test support (planning-only contexts): walks the host lists of a grouped product the way gemm_grouped_kernel / splitk_reduce_kernel address device memory --
// every (problem, tile) item computes its clipped tile of alpha * op(A) op(B) over its K range into C or into its slice of the scratch buffer, the split-K jobs
// sum their slices in slice order -- with every offset bounds-checked and every element of a non-accumulating destination written exactly once.
void emulate_gemm_lists(const GemmLists& G, const double* A, int64_t n_a, const double* B, int64_t n_b, double* C, int64_t n_c)
{
    std::vector<double> scratch(static_cast<size_t>(std::max<long long>(G.scratch_elems, 0)), std::numeric_limits<double>::quiet_NaN());
    std::vector<unsigned char> written_c(static_cast<size_t>(n_c), 0), written_s(scratch.size(), 0);
    auto at = [](const double* p, int64_t n, long long o, const char* what) {
        if(o < 0 || o >= n) throw std::runtime_error(std::string("emulate_gemm_lists: ") + what + " offset out of range");
        return p[o];
    };
    for(int which = 0; which < 2; ++which) {
        const int tile_m = which == 0 ? 128 : 64;
        for(const GemmItem& it : G.items[which]) {
            if(it.tm % tile_m != 0 || it.tn % GEMM_TILE_N != 0 || it.tm >= it.m || it.tn >= it.n) throw std::runtime_error("emulate_gemm_lists: bad tile origin");
            const int i1 = std::min(it.m, it.tm + tile_m), j1 = std::min(it.n, it.tn + GEMM_TILE_N);
            for(int j = it.tn; j < j1; ++j)
                for(int i = it.tm; i < i1; ++i) {
                    double acc = 0.0;
                    for(int kk = 0; kk < it.k; ++kk) {
                        const double a = at(A, n_a, it.a_off + (it.ta ? kk + static_cast<long long>(i) * it.lda : i + static_cast<long long>(kk) * it.lda), "A");
                        const double b = at(B, n_b, it.b_off + (it.tb ? j + static_cast<long long>(kk) * it.ldb : kk + static_cast<long long>(j) * it.ldb), "B");
                        acc = std::fma(a, b, acc);
                    }
                    const long long o = it.c_off + i + static_cast<long long>(j) * it.ldc;
                    if(it.pad_) {   // a K slice: scratch destination
                        if(o < 0 || o >= static_cast<long long>(scratch.size())) throw std::runtime_error("emulate_gemm_lists: scratch offset out of range");
                        if(written_s[static_cast<size_t>(o)]++) throw std::runtime_error("emulate_gemm_lists: scratch element written twice");
                        scratch[static_cast<size_t>(o)] = it.alpha * acc;
                    } else {
                        if(o < 0 || o >= n_c) throw std::runtime_error("emulate_gemm_lists: C offset out of range");
                        if(!it.acc && written_c[static_cast<size_t>(o)]++) throw std::runtime_error("emulate_gemm_lists: C element written twice");
                        C[o] = it.acc ? C[o] + it.alpha * acc : it.alpha * acc;
                    }
                }
        }
    }
    for(const SplitKJob& j : G.jobs) {
        const long long cnt = static_cast<long long>(j.m) * j.n;
        for(long long e = 0; e < cnt; ++e) {
            double sum = 0.0;
            for(int sl = 0; sl < j.split; ++sl) {
                const long long o = j.s_off + sl * cnt + e;
                if(o < 0 || o >= static_cast<long long>(scratch.size()) || !written_s[static_cast<size_t>(o)]) throw std::runtime_error("emulate_gemm_lists: slice element missing");
                sum += scratch[static_cast<size_t>(o)];
            }
            const long long col = e / j.m, o = j.c_off + (e - col * j.m) + col * j.ldc;
            if(o < 0 || o >= n_c) throw std::runtime_error("emulate_gemm_lists: job destination out of range");
            if(written_c[static_cast<size_t>(o)]++) throw std::runtime_error("emulate_gemm_lists: C element written twice");
            C[o] = sum;
        }
    }
}

// the per-sector problems of operator* (src/Core/TensorBase.cpp:262-279): host metadata only
std::vector<GemmItem> multiply_problems(const Layout& LA, bool adj_a, const Layout& LB, bool adj_b, const Layout& LO)
{
    std::vector<GemmItem> items;
    for(size_t i = 0; i < LA.sector.size(); ++i) {
        const int j = LB.find(LA.sector[i]);
        if(j < 0) continue;
        const int o = LO.find(LA.sector[i]);
        if(o < 0) continue;
        GemmItem g;
        std::memset(&g, 0, sizeof(g));
        g.a_off = LA.off[i];
        g.b_off = LB.off[j];
        g.c_off = LO.off[o];
        g.lda = static_cast<int>(LA.rows[i]);
        g.ldb = static_cast<int>(LB.rows[j]);
        g.ldc = static_cast<int>(LO.rows[o]);
        g.ta = adj_a;
        g.tb = adj_b;
        g.m = static_cast<int>(adj_a ? LA.cols[i] : LA.rows[i]);
        g.k = static_cast<int>(adj_a ? LA.rows[i] : LA.cols[i]);
        const int kb = static_cast<int>(adj_b ? LB.cols[j] : LB.rows[j]);
        g.n = static_cast<int>(adj_b ? LB.rows[j] : LB.cols[j]);
        if(kb != g.k || g.m != LO.rows[o] || g.n != LO.cols[o]) throw std::logic_error("operator*: block shape mismatch");
        if(g.m == 0 || g.n == 0) continue;
        g.alpha = 1.0;
        items.push_back(g);
    }
    return items;
}

Tensor* Context::multiply(const Tensor& a, bool adj_a, const Tensor& b, bool adj_b) { return multiply_sectors(a, adj_a, b, adj_b, nullptr); }

// `out_sectors` (sector-sharded operands): the result holds exactly these sectors in this order (the ones this
// rank owns) instead of the canonical list; operands are shards, sectors absent from either stay zero blocks.
Tensor* Context::multiply_sectors(const Tensor& a, bool adj_a, const Tensor& b, bool adj_b, const std::vector<Q>* out_sectors)
{
    const Layout& LA = *a.layout;
    const Layout& LB = *b.layout;
    const auto& a_dom = adj_a ? LA.cod : LA.dom;
    const auto& a_cod = adj_a ? LA.dom : LA.cod;
    const auto& b_dom = adj_b ? LB.cod : LB.dom;
    const auto& b_cod = adj_b ? LB.dom : LB.cod;
    if(a_cod.size() != b_dom.size()) throw std::invalid_argument("operator*: CoRank of left operand != Rank of right operand");
    for(size_t i = 0; i < a_cod.size(); ++i)   // DEBUG_ASSERT(coupledCodomain == coupledDomain), TensorBase.cpp:254
        if(!a_cod[i]->same_sectors(*b_dom[i])) throw std::invalid_argument("operator*: coupledCodomain(left) != coupledDomain(right)");
    LayoutPtr LO = out_sectors ? layout_custom(a_dom, b_cod, *out_sectors) : layout(a_dom, b_cod);
    // Tout.setZero(): sectors missing in an operand stay zero blocks -- but a product that covers every sector of the result overwrites all of it, and the
    // fill (one memset per product, ~90 per CTM step, 155 MB each at D=6) is skipped
    size_t covered = 0;
    for(size_t o = 0; o < LO->sector.size(); ++o) {
        if(LO->rows[o] == 0 || LO->cols[o] == 0) {
            ++covered;
            continue;
        }
        const int i = LA.find(LO->sector[o]);
        if(i >= 0 && LB.find(LO->sector[o]) >= 0) ++covered;
    }
    auto* out = new Tensor(this, LO, covered != LO->sector.size());

    std::vector<GemmItem> items = multiply_problems(LA, adj_a, LB, adj_b, *LO);
    {
        static const bool no_cache = std::getenv("XB_GEMM_NOCACHE") != nullptr;
        const std::string key = LA.sig + "|" + LB.sig + "|" + LO->sig + (adj_a ? "|a" : "|n") + (adj_b ? "a" : "n");
        auto it = no_cache ? gemm_cache.end() : gemm_cache.find(key);
        if(it != gemm_cache.end() && it->second.P.d_scratch != nullptr && it->second.c_ptr != out->d) {   // scratch offsets were taken relative to another result
            free_gemm(it->second.P);
            gemm_cache.erase(it);
            it = gemm_cache.end();
        }
        if(no_cache) {
            run_gemm(a.d, b.d, out->d, items);
        } else {
            if(it == gemm_cache.end()) {
                ++gemm_cache_misses;
                if(gemm_cache.size() >= 512) {
                    for(auto& kv : gemm_cache) free_gemm(kv.second.P);
                    gemm_cache.clear();
                }
                GemmCacheEntry e;
                e.P = prepare_gemm(a.d, b.d, out->d, items);
                e.c_ptr = out->d;
                it = gemm_cache.emplace(key, e).first;
            } else {
                ++gemm_cache_hits;
            }
            PreparedGemm& P = it->second.P;
            P.A = a.d;
            P.B = b.d;
            P.C = out->d;
            launch_gemm(P);
        }
    }
    if(taping()) {   // ADTensor.hpp:749-758: adj(left) += adj(res) * right^+, adj(right) += left^+ * adj(res); lazy-adjoint operands get the adjoint of that
        Tensor *ta = tape->keep(a), *tb = tape->keep(b), *tout = tape->keep(*out);
        tape->record([this, ta, tb, tout, adj_a, adj_b](Tape& T) {
            Tensor* g = T.grad(*tout);
            if(!g) return;
            T.accumulate(*ta, adj_a ? multiply(*tb, adj_b, *g, true) : multiply(*g, false, *tb, !adj_b));
            g = T.grad(*tout);   // (the map may have rehashed)
            T.accumulate(*tb, adj_b ? multiply(*g, true, *ta, adj_a) : multiply(*ta, !adj_a, *g, false));
        });
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ contract
// Tensor::contract (include/Xped/Core/Tensor.hpp:328-351) with get_permutations (Util/Constfct.hpp:55-109).
// get_permutations (include/Xped/Util/Constfct.hpp:55-109) for contract<a1, a2, ResRank>: the leg permutations / shifts that turn the contraction into
// permute(a) * permute(b) followed by a permute of the product.  Host metadata only.
ContractPerms contract_permutations(int rank_a, int corank_a, const std::vector<int>& l1, int rank_b, int corank_b, const std::vector<int>& l2, int res_rank)
{
    ContractPerms cp;
    const int r1 = static_cast<int>(l1.size()), r2 = static_cast<int>(l2.size());
    if(r1 != rank_a + corank_a || r2 != rank_b + corank_b) throw std::invalid_argument("contract: label array length != number of legs");
    const int max1 = std::max(0, r1 ? *std::max_element(l1.begin(), l1.end()) : 0);
    const int max2 = std::max(0, r2 ? *std::max_element(l2.begin(), l2.end()) : 0);
    if(max1 != max2) throw std::invalid_argument("contract: contracted labels do not match");
    auto pos = [](const std::vector<int>& v, int label) {
        auto it = std::find(v.begin(), v.end(), label);
        if(it == v.end()) throw std::invalid_argument("contract: missing contraction label");
        return static_cast<int>(it - v.begin());
    };
    std::vector<int>&p1 = cp.p1, &p2 = cp.p2, &pres = cp.pres;
    std::vector<int> con1;
    for(int k = 1; k <= max2; ++k) p2.push_back(pos(l2, k));
    for(int i = 0; i < r2; ++i)
        if(std::find(p2.begin(), p2.begin() + max2, i) == p2.begin() + max2) p2.push_back(i);
    for(int k = 1; k <= max1; ++k) con1.push_back(pos(l1, k));
    for(int i = 0; i < r1; ++i)
        if(std::find(con1.begin(), con1.end(), i) == con1.end()) p1.push_back(i);
    p1.insert(p1.end(), con1.begin(), con1.end());
    std::vector<int> open;
    for(int x : l1)
        if(x < 0) open.push_back(x);
    for(int x : l2)
        if(x < 0) open.push_back(x);
    pres.resize(open.size());
    std::iota(pres.begin(), pres.end(), 0);
    std::sort(pres.begin(), pres.end(), [&](int i, int j) { return open[i] > open[j]; });
    cp.shift1 = rank_a - r1 + max1;
    cp.shift2 = rank_b - max2;
    cp.shiftres = r1 - max1 - res_rank;
    return cp;
}

Tensor* Context::contract(const Tensor& a, const std::vector<int>& l1, const Tensor& b, const std::vector<int>& l2, int res_rank)
{
    const ContractPerms cp = contract_permutations(a.layout->rank(), a.layout->corank(), l1, b.layout->rank(), b.layout->corank(), l2, res_rank);
    const std::vector<int>&p1 = cp.p1, &p2 = cp.p2, &pres = cp.pres;
    const int shift1 = cp.shift1, shift2 = cp.shift2, shiftres = cp.shiftres;

    std::unique_ptr<Tensor> ap(permute(a, shift1, p1));
    std::unique_ptr<Tensor> bp(permute(b, shift2, p2));
    std::unique_ptr<Tensor> prod(multiply(*ap, false, *bp, false));
    return permute(*prod, shiftres, pres);
}

// ------------------------------------------------------------------------------------------------ adjoint / twist / elementwise
namespace {
std::vector<BlockDesc> block_descs(const Layout& L, const Symmetry* weight_sym)
{
    std::vector<BlockDesc> v(L.sector.size());
    for(size_t i = 0; i < L.sector.size(); ++i)
        v[i] = {L.off[i], static_cast<int>(L.rows[i]), static_cast<int>(L.cols[i]), weight_sym ? static_cast<double>(weight_sym->degeneracy(L.sector[i])) : 1.0};
    return v;
}
}   // namespace

Tensor* Context::adjoint(const Tensor& t)
{
    // AdjointOp::eval(): domain <-> codomain, blocks pushed in the operand's sector order (Tensor.hpp:448-453)
    const Layout& L = *t.layout;
    LayoutPtr LO = layout_custom(L.cod, L.dom, L.sector);
    auto* out = new Tensor(this, LO, false);
    const int nb = static_cast<int>(L.sector.size());
    if(nb == 0) return out;
    auto descs = block_descs(L, nullptr);
    std::vector<long long> doff(LO->off.begin(), LO->off.end());
    const size_t b1 = descs.size() * sizeof(BlockDesc), b2 = doff.size() * sizeof(long long);
    char* d_meta = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_meta), b1 + b2, stream), "cudaMallocAsync adjoint meta");
    upload_async(d_meta, descs.data(), b1);
    upload_async(d_meta + b1, doff.data(), b2);
    launch_transpose_blocks(t.d, out->d, reinterpret_cast<BlockDesc*>(d_meta), reinterpret_cast<long long*>(d_meta + b1), nb, stream);
    check_cuda(cudaFreeAsync(d_meta, stream), "cudaFreeAsync adjoint meta");
    ++launches;
    if(taping()) {   // ADTensor.hpp:137-150
        Tensor *tin = tape->keep(t), *tout = tape->keep(*out);
        tape->record([this, tin, tout](Tape& T) {
            if(Tensor* g = T.grad(*tout)) T.accumulate(*tin, adjoint(*g));
        });
    }
    return out;
}

Tensor* Context::twist(const Tensor& t, int leg)
{
    Tensor* out = nullptr;
    {
        TapePause pause(*this);
        out = twist_impl(t, leg);
    }
    if(taping()) {   // ADTensor.hpp:589-602: adj(t) += adj(res).twist(leg)
        Tensor *tin = tape->keep(t), *tout = tape->keep(*out);
        tape->record([this, tin, tout, leg](Tape& T) {
            if(Tensor* g = T.grad(*tout)) T.accumulate(*tin, twist(*g, leg));
        });
    }
    return out;
}

Tensor* Context::twist_impl(const Tensor& t, int leg)
{
    // Tensor::twist (src/Core/Tensor.cpp:357-377): early-out copy for bosonic symmetries (:362).
    if(!sym->fermionic()) return copy(t);
    const Layout& L = *t.layout;
    // Fermionic: scale every tree-pair sub-block by coeff_twist of the leg's sector label, on the device.
    const int R = L.rank(), N = R + L.corank();
    if(leg < 0 || leg >= N) throw std::invalid_argument("twist: leg out of range");
    auto* out = copy(t);
    std::vector<SubScale> subs;
    for(size_t i = 0; i < L.sector.size(); ++i) {
        const auto& dts = L.cdom->trees(L.sector[i]);
        const auto& cts = L.ccod->trees(L.sector[i]);
        int64_t ro = 0;
        for(const Tree& dt : dts) {
            int64_t co = 0;
            for(const Tree& ct : cts) {
                const double c = leg < R ? sym->coeff_twist(dt.unc[leg]) : sym->coeff_twist(ct.unc[leg - R]);
                if(c != 1.0) subs.push_back({L.off[i] + ro + co * L.rows[i], static_cast<int>(dt.dim), static_cast<int>(ct.dim), static_cast<int>(L.rows[i]), c});
                co += ct.dim;
            }
            ro += dt.dim;
        }
    }
    if(!subs.empty()) {
        SubScale* d_subs = nullptr;
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_subs), subs.size() * sizeof(SubScale), stream), "cudaMallocAsync twist");
        upload_async(d_subs, subs.data(), subs.size() * sizeof(SubScale));
        launch_scale_subblocks(out->d, d_subs, static_cast<int>(subs.size()), stream);
        check_cuda(cudaFreeAsync(d_subs, stream), "cudaFreeAsync twist");
        ++launches;
    }
    return out;
}

Tensor* Context::axpby(double alpha, const Tensor& x, double beta, const Tensor* y)
{
    auto* out = new Tensor(this, x.layout, false);
    if(y == nullptr) {
        launch_scale_copy(x.d, out->d, alpha, x.size(), stream);   // CoeffUnaryOp (CoeffUnaryOp.hpp:48-49)
    } else {
        // CoeffBinaryOp pairs the i-th blocks and requires identical sector lists (CoeffBinaryOp.hpp:44-64)
        if(y->layout->sector != x.layout->sector || y->layout->size != x.layout->size) throw std::invalid_argument("axpby: operands have different block structure");
        launch_axpby(x.d, y->d, out->d, alpha, beta, x.size(), stream);
    }
    ++launches;
    if(taping()) {   // ADTensor.hpp:640-676
        Tensor *tx = tape->keep(x), *ty = y ? tape->keep(*y) : nullptr, *tout = tape->keep(*out);
        tape->record([this, tx, ty, tout, alpha, beta](Tape& T) {
            Tensor* g = T.grad(*tout);
            if(!g) return;
            T.accumulate(*tx, axpby(alpha, *g, 0.0, nullptr));
            if(ty) {
                g = T.grad(*tout);
                T.accumulate(*ty, axpby(beta, *g, 0.0, nullptr));
            }
        });
    }
    return out;
}

Tensor* Context::diag_map(const Tensor& t, int op)
{
    const Layout& L = *t.layout;
    auto* out = new Tensor(this, t.layout, false);
    const int nb = static_cast<int>(L.sector.size());
    if(nb == 0) return out;
    auto descs = block_descs(L, nullptr);
    BlockDesc* d_meta = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_meta), descs.size() * sizeof(BlockDesc), stream), "cudaMallocAsync diag meta");
    upload_async(d_meta, descs.data(), descs.size() * sizeof(BlockDesc));
    launch_diag_map(t.d, out->d, d_meta, nb, op, stream);
    check_cuda(cudaFreeAsync(d_meta, stream), "cudaFreeAsync diag meta");
    ++launches;
    if(taping()) {   // ADTensor.hpp:559-586 (diag_sqrt: adj * 0.5 / res, diag_inv: adj * (-res^2)) and their composition
        Tensor *tin = tape->keep(t), *tout = tape->keep(*out);
        tape->record([this, tin, tout, op](Tape& T) {
            Tensor* g = T.grad(*tout);
            if(!g) return;
            const Layout& Lo = *tout->layout;
            auto* gi = new Tensor(this, tin->layout, false);
            auto dd = block_descs(Lo, nullptr);
            BlockDesc* dm = nullptr;
            check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&dm), dd.size() * sizeof(BlockDesc), stream), "cudaMallocAsync diag meta");
            upload_async(dm, dd.data(), dd.size() * sizeof(BlockDesc));
            launch_diag_map_backward(g->d, tout->d, gi->d, dm, static_cast<int>(dd.size()), op, stream);
            check_cuda(cudaFreeAsync(dm, stream), "cudaFreeAsync diag meta");
            ++launches;
            T.accumulate(*tin, gi);
        });
    }
    return out;
}

double Context::max_norm(const Tensor& t)
{
    double* d_out = static_cast<double*>(scratch(sizeof(double)));
    launch_absmax(t.d, t.size(), d_out, stream);
    ++launches;
    double h = 0.0;
    check_cuda(cudaMemcpyAsync(&h, d_out, sizeof(double), cudaMemcpyDeviceToHost, stream), "max_norm readback");
    sync();
    return h;
}

double Context::trace(const Tensor& t)
{
    const Layout& L = *t.layout;
    auto descs = block_descs(L, sym.get());
    const size_t bytes = descs.size() * sizeof(BlockDesc);
    char* d_meta = static_cast<char*>(scratch(bytes + 64));
    upload_async(d_meta + 64, descs.data(), bytes);
    launch_trace(t.d, reinterpret_cast<BlockDesc*>(d_meta + 64), static_cast<int>(descs.size()), reinterpret_cast<double*>(d_meta), stream);
    ++launches;
    double h = 0.0;
    check_cuda(cudaMemcpyAsync(&h, d_meta, sizeof(double), cudaMemcpyDeviceToHost, stream), "trace readback");
    sync();
    return h;
}

double Context::squared_norm(const Tensor& t)
{
    // squaredNorm = (T * T^dagger).trace() (src/Core/TensorBase.cpp:58-64)
    std::unique_ptr<Tensor> p(multiply(t, false, t, true));
    return trace(*p);
}

}   // namespace xb

namespace xb {

Tensor* Context::scale_by_inv_maxnorm(const Tensor& t)
{
    auto* out = new Tensor(this, t.layout, false);
    double* d_scalar = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_scalar), sizeof(double), stream), "cudaMallocAsync maxnorm scalar");
    launch_absmax(t.d, t.size(), d_scalar, stream);
    launch_scale_by_inv(t.d, out->d, d_scalar, t.size(), stream);
    check_cuda(cudaFreeAsync(d_scalar, stream), "cudaFreeAsync maxnorm scalar");
    launches += 2;
    if(taping()) {   // maxNorm (ADTensor.hpp:502-529) fused with the division that follows it in renormalize_* (CTM.cpp:1369,1381,1409)
        Tensor *tin = tape->keep(t), *tout = tape->keep(*out);
        tape->record([this, tin, tout](Tape& T) {
            Tensor* g = T.grad(*tout);
            if(!g) return;
            auto* gi = new Tensor(this, tin->layout, false);
            double* d_tmp = nullptr;   // [0] max |x|, [1] <g, x>, [2..130] partial sums
            check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_tmp), sizeof(double) * 132, stream), "cudaMallocAsync maxnorm backward");
            launch_absmax(tin->d, tin->size(), d_tmp, stream);
            launch_dot(g->d, tin->d, tin->size(), d_tmp + 1, stream);
            launch_maxnorm_scale_backward(g->d, tin->d, d_tmp, d_tmp + 1, gi->d, tin->size(), stream);
            check_cuda(cudaFreeAsync(d_tmp, stream), "cudaFreeAsync maxnorm backward");
            launches += 4;
            T.accumulate(*tin, gi);
        });
    }
    return out;
}

// x / trace(x * w): normalisation of a density matrix (CTM.cpp:660-667: norm = rho.contract<...>(Id.twist(0).twist(1)).trace(); rho = rho * (1 / norm)).
// Reverse rule (operator*(tensor, var) + trace, ADTensor.hpp:531-546,690-706): adj(x) = g / n - (<g, x> / n^2) * d n / d x, d n / d x_q = degeneracy(q) w_q^T.
Tensor* Context::scale_by_inv_contraction(const Tensor& x, const Tensor& w)
{
    double n = 0.0;
    Tensor* out = nullptr;
    {
        TapePause pause(*this);
        std::unique_ptr<Tensor> prod(multiply(x, false, w, false));
        n = trace(*prod);
        out = axpby(1.0 / n, x, 0.0, nullptr);
    }
    if(taping()) {
        Tensor *tx = tape->keep(x), *tw = tape->keep(w), *tout = tape->keep(*out);
        tape->record([this, tx, tw, tout, n](Tape& T) {
            Tensor* g = T.grad(*tout);
            if(!g) return;
            std::unique_ptr<Tensor> gx(multiply(*g, false, *tx, true));
            const double inner = unweighted_trace(*gx);   // <g, x>
            // degeneracy-weighted adjoint of w, laid out like x
            std::unique_ptr<Tensor> wa(adjoint(*tw));
            const Layout& L = *wa->layout;
            std::vector<SubScale> subs;
            for(size_t i = 0; i < L.sector.size(); ++i)
                subs.push_back({L.off[i], static_cast<int>(L.rows[i]), static_cast<int>(L.cols[i]), static_cast<int>(L.rows[i]), static_cast<double>(sym->degeneracy(L.sector[i]))});
            if(!subs.empty()) {
                SubScale* d_subs = nullptr;
                check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_subs), subs.size() * sizeof(SubScale), stream), "cudaMallocAsync weights");
                upload_async(d_subs, subs.data(), subs.size() * sizeof(SubScale));
                launch_scale_subblocks(wa->d, d_subs, static_cast<int>(subs.size()), stream);
                check_cuda(cudaFreeAsync(d_subs, stream), "cudaFreeAsync weights");
                ++launches;
            }
            if(wa->layout->size != g->layout->size || wa->layout->sector != g->layout->sector) throw std::logic_error("scale_by_inv_contraction: weight layout differs from the operand's");
            T.accumulate(*tx, axpby(1.0 / n, *g, -inner / (n * n), wa.get()));
        });
    }
    return out;
}

Tensor* Context::weighted_identity(const LayoutPtr& L)
{
    auto* out = new Tensor(this, L, true);
    fill(*out, 1, 0.0, 0);
    std::vector<SubScale> subs;
    for(size_t i = 0; i < L->sector.size(); ++i)
        subs.push_back({L->off[i], static_cast<int>(L->rows[i]), static_cast<int>(L->cols[i]), static_cast<int>(L->rows[i]), static_cast<double>(sym->degeneracy(L->sector[i]))});
    if(!subs.empty()) {
        SubScale* d_subs = nullptr;
        check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_subs), subs.size() * sizeof(SubScale), stream), "cudaMallocAsync weights");
        upload_async(d_subs, subs.data(), subs.size() * sizeof(SubScale));
        launch_scale_subblocks(out->d, d_subs, static_cast<int>(subs.size()), stream);
        check_cuda(cudaFreeAsync(d_subs, stream), "cudaFreeAsync weights");
        ++launches;
    }
    return out;
}

double Context::unweighted_trace(const Tensor& t)
{
    const Layout& L = *t.layout;
    auto descs = block_descs(L, nullptr);
    const size_t bytes = descs.size() * sizeof(BlockDesc);
    char* d_meta = static_cast<char*>(scratch(bytes + 64));
    upload_async(d_meta + 64, descs.data(), bytes);
    launch_trace(t.d, reinterpret_cast<BlockDesc*>(d_meta + 64), static_cast<int>(descs.size()), reinterpret_cast<double*>(d_meta), stream);
    ++launches;
    double h = 0.0;
    check_cuda(cudaMemcpyAsync(&h, d_meta, sizeof(double), cudaMemcpyDeviceToHost, stream), "trace readback");
    sync();
    return h;
}

void Context::fill(Tensor& t, int mode, double value, uint64_t seed)
{
    const Layout& L = *t.layout;
    if(mode == 0) {
        launch_fill_constant(t.d, value, L.size, stream);
    } else if(mode == 2) {
        launch_fill_random(t.d, seed, L.size, stream);
    } else {
        check_cuda(cudaMemsetAsync(t.d, 0, sizeof(double) * static_cast<size_t>(L.size), stream), "fill identity memset");
        std::vector<BlockDesc> descs(L.sector.size());
        for(size_t i = 0; i < L.sector.size(); ++i) descs[i] = {L.off[i], static_cast<int>(L.rows[i]), static_cast<int>(L.cols[i]), 1.0};
        if(!descs.empty()) {
            BlockDesc* d_meta = nullptr;
            check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_meta), descs.size() * sizeof(BlockDesc), stream), "cudaMallocAsync identity meta");
            upload_async(d_meta, descs.data(), descs.size() * sizeof(BlockDesc));
            launch_set_identity(t.d, d_meta, static_cast<int>(descs.size()), stream);
            check_cuda(cudaFreeAsync(d_meta, stream), "cudaFreeAsync identity meta");
        }
    }
    ++launches;
}

}   // namespace xb
