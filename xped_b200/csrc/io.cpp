// State interop.
//
// 1. Native container "XB2T": symmetry, leg bases, sector order and the raw `m_data` image of a tensor (what the reference serialises with yas: data, sectors,
//    uncoupled bases; include/Xped/Core/storage/StorageType_contiguous.hpp:191-204, Qbasis.hpp:249-253) -- save / load of tensors and of a whole CTM state
//    (chi, C1..C4, T1..T4, A), i.e. checkpoint / resume of an environment.  The yas byte format itself is NOT reproduced: neither the yas library nor a file
//    written by the reference exists in this offline image, so byte compatibility could not be verified ("parity unpinned").
// 2. The charge-block (MATLAB) tensor format of IO::loadMatlabTensor (include/Xped/IO/Matlab.hpp:9-95): per leg a list of (charge, sector dimension), `vdat` = all
//    block data, `voffs` = (combined charge index, data offset) per block.  All legs go into the domain of a rank-(N,0) tensor coupled to the vacuum, every
//    block is written into the sub-block of its fusion tree, then permute<N - Rank>(identity) moves the trailing legs to the codomain (:93).  The HDF5 container
//    around these arrays needs HighFive / libhdf5 (absent here): the entry point takes the arrays themselves.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "ctm.hpp"
#include "engine.hpp"

namespace xb {

namespace {

constexpr char MAGIC_T[8] = {'X', 'B', '2', 'T', 'E', 'N', 'S', '1'};
constexpr char MAGIC_C[8] = {'X', 'B', '2', 'C', 'T', 'M', 'S', '1'};

template <typename T> void put(std::ostream& o, const T& v) { o.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T> T get(std::istream& i)
{
    T v;
    i.read(reinterpret_cast<char*>(&v), sizeof(T));
    if(!i) throw std::runtime_error("xped_b200 io: truncated file");
    return v;
}

void write_tensor(std::ostream& o, Context& ctx, const Tensor& t)
{
    const Layout& L = *t.layout;
    const Symmetry& sym = *ctx.sym;
    o.write(MAGIC_T, 8);
    put<int32_t>(o, static_cast<int32_t>(sym.factors().size()));
    for(const Factor& f : sym.factors()) {
        put<int32_t>(o, static_cast<int32_t>(f.kind));
        put<int32_t>(o, f.N);
        put<int32_t>(o, f.fermionic ? 1 : 0);
    }
    const int nq = sym.Nq();
    put<int32_t>(o, nq);
    put<int32_t>(o, L.rank());
    put<int32_t>(o, L.corank());
    auto write_basis = [&](const Qbasis& b) {
        put<int32_t>(o, b.is_conj() ? 1 : 0);
        put<int64_t>(o, static_cast<int64_t>(b.entries().size()));
        for(const auto& e : b.entries()) {
            for(int k = 0; k < nq; ++k) put<int32_t>(o, e.q[k]);
            put<int64_t>(o, e.dim);
        }
    };
    for(const auto& b : L.dom) write_basis(*b);
    for(const auto& b : L.cod) write_basis(*b);
    put<int64_t>(o, static_cast<int64_t>(L.sector.size()));
    for(size_t i = 0; i < L.sector.size(); ++i) {
        for(int k = 0; k < nq; ++k) put<int32_t>(o, L.sector[i][k]);
        put<int64_t>(o, L.rows[i]);
        put<int64_t>(o, L.cols[i]);
    }
    put<int64_t>(o, L.size);
    std::vector<double> h(static_cast<size_t>(L.size));
    if(L.size > 0) {
        check_cuda(cudaMemcpyAsync(h.data(), t.d, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx.stream), "save: download");
        check_cuda(cudaStreamSynchronize(ctx.stream), "save: sync");
        o.write(reinterpret_cast<const char*>(h.data()), static_cast<std::streamsize>(sizeof(double) * h.size()));
    }
    if(!o) throw std::runtime_error("xped_b200 io: write failed");
}

Tensor* read_tensor(std::istream& in, Context& ctx)
{
    char magic[8];
    in.read(magic, 8);
    if(!in || std::memcmp(magic, MAGIC_T, 8) != 0) throw std::invalid_argument("xped_b200 io: not a tensor record (bad magic)");
    const Symmetry& sym = *ctx.sym;
    const int nf = get<int32_t>(in);
    if(nf != static_cast<int>(sym.factors().size())) throw std::invalid_argument("xped_b200 io: the file was written for another symmetry");
    for(const Factor& f : sym.factors()) {
        const int kind = get<int32_t>(in), N = get<int32_t>(in), ferm = get<int32_t>(in);
        if(kind != static_cast<int>(f.kind) || N != f.N || (ferm != 0) != f.fermionic) throw std::invalid_argument("xped_b200 io: the file was written for another symmetry");
    }
    const int nq = get<int32_t>(in);
    if(nq != sym.Nq()) throw std::invalid_argument("xped_b200 io: number of quantum numbers differs");
    const int rank = get<int32_t>(in), corank = get<int32_t>(in);
    if(rank < 0 || corank < 0 || rank > MAXR || corank > MAXR) throw std::invalid_argument("xped_b200 io: implausible rank");
    auto read_basis = [&]() -> BasisPtr {
        const bool conj = get<int32_t>(in) != 0;
        const int64_t n = get<int64_t>(in);
        if(n < 0 || n > (1 << 24)) throw std::invalid_argument("xped_b200 io: implausible basis size");
        auto b = std::make_shared<Qbasis>(ctx.sym, 1);
        for(int64_t i = 0; i < n; ++i) {
            Q q{};
            for(int k = 0; k < nq; ++k) q[k] = get<int32_t>(in);
            const int64_t dim = get<int64_t>(in);
            b->push_back(conj ? sym.conj(q) : q, dim);   // a conjugated basis is rebuilt as the conj() of its pre-image (dual trees, Qbasis.cpp:418-439)
        }
        return conj ? b->conj() : BasisPtr(b);
    };
    std::vector<BasisPtr> dom, cod;
    for(int i = 0; i < rank; ++i) dom.push_back(read_basis());
    for(int i = 0; i < corank; ++i) cod.push_back(read_basis());
    const int64_t nsec = get<int64_t>(in);
    if(nsec < 0 || nsec > (1 << 24)) throw std::invalid_argument("xped_b200 io: implausible sector count");
    std::vector<Q> sectors(static_cast<size_t>(nsec));
    std::vector<int64_t> rows(static_cast<size_t>(nsec)), cols(static_cast<size_t>(nsec));
    for(int64_t i = 0; i < nsec; ++i) {
        Q q{};
        for(int k = 0; k < nq; ++k) q[k] = get<int32_t>(in);
        sectors[static_cast<size_t>(i)] = q;
        rows[static_cast<size_t>(i)] = get<int64_t>(in);
        cols[static_cast<size_t>(i)] = get<int64_t>(in);
    }
    const int64_t size = get<int64_t>(in);
    LayoutPtr L = ctx.layout(dom, cod);
    if(L->sector != sectors) L = ctx.layout_custom(dom, cod, sectors);
    if(L->size != size || L->rows != rows || L->cols != cols) throw std::invalid_argument("xped_b200 io: the stored block structure does not match the bases");
    std::unique_ptr<Tensor> t(new Tensor(&ctx, L, false));
    if(size > 0) {
        std::vector<double> h(static_cast<size_t>(size));
        in.read(reinterpret_cast<char*>(h.data()), static_cast<std::streamsize>(sizeof(double) * h.size()));
        if(!in) throw std::runtime_error("xped_b200 io: truncated file");
        check_cuda(cudaMemcpyAsync(t->d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx.stream), "load: upload");
        check_cuda(cudaStreamSynchronize(ctx.stream), "load: sync");
    }
    return t.release();
}

}   // namespace

void save_tensor(Context& ctx, const Tensor& t, const std::string& path)
{
    std::ofstream o(path, std::ios::binary | std::ios::trunc);
    if(!o) throw std::runtime_error("xped_b200 io: cannot open '" + path + "' for writing");
    write_tensor(o, ctx, t);
}

Tensor* load_tensor(Context& ctx, const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if(!in) throw std::runtime_error("xped_b200 io: cannot open '" + path + "'");
    return read_tensor(in, ctx);
}

void save_ctm(Context& ctx, const Ctm& c, int64_t chi, const std::string& path)
{
    std::ofstream o(path, std::ios::binary | std::ios::trunc);
    if(!o) throw std::runtime_error("xped_b200 io: cannot open '" + path + "' for writing");
    o.write(MAGIC_C, 8);
    put<int64_t>(o, chi);
    for(int which = 0; which < 10; ++which) write_tensor(o, ctx, c.get(which));
}

Ctm* load_ctm(Context& ctx, const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if(!in) throw std::runtime_error("xped_b200 io: cannot open '" + path + "'");
    char magic[8];
    in.read(magic, 8);
    if(!in || std::memcmp(magic, MAGIC_C, 8) != 0) throw std::invalid_argument("xped_b200 io: not a CTM state (bad magic)");
    const int64_t chi = get<int64_t>(in);
    std::vector<TP> ts;
    for(int which = 0; which < 10; ++which) ts.emplace_back(read_tensor(in, ctx));
    return new Ctm(ctx, chi, ts);
}

// IO::loadMatlabTensor (include/Xped/IO/Matlab.hpp:9-95) on the arrays an HDF5 reader hands over.
//   charges[leg][j], sdims[leg][j]  : `charges` dataset of the leg, row 0 (charge values) and row 1 (sector dimensions), j < ncharge[leg]
//   voffs_index[b], voffs_offset[b] : `voffs` rows: combined charge index (leg 0 fastest, :17-33) and the offset of the block in `vdat`
// Single Abelian factor (U(1) or Z_N), like the reference (it builds qarray<1> charges and reduces them modulo Symmetry::MOD_N[0]).
Tensor* tensor_from_charge_blocks(Context& ctx, int nlegs, int rank, const int* ncharge, const double* const* charges, const double* const* sdims, const int* conj,
                                  int64_t nblocks, const int64_t* voffs_index, const int64_t* voffs_offset, const double* vdat, int64_t nvdat, int qn_scale)
{
    const Symmetry& sym = *ctx.sym;
    if(sym.Nq() != 1 || sym.non_abelian()) throw std::invalid_argument("charge-block import: one Abelian quantum number expected (U1 or ZN), like IO::loadMatlabTensor");
    if(nlegs < 1 || nlegs > MAXR || rank < 0 || rank > nlegs) throw std::invalid_argument("charge-block import: bad leg count / rank");
    const Factor* fac = nullptr;
    for(const Factor& f : sym.factors())
        if(f.kind != Kind::U0) fac = &f;
    const int mod = (fac && fac->kind == Kind::ZN) ? fac->N : 0;
    auto charge_of = [&](double c) {   // :52: (qn_scale * round(c)) % MOD_N
        int q = qn_scale * static_cast<int>(std::lround(c));
        if(mod > 0) q = ((q % mod) + mod) % mod;
        return q;
    };
    std::vector<BasisPtr> legs;
    for(int r = 0; r < nlegs; ++r) {
        auto b = std::make_shared<Qbasis>(ctx.sym, 1);
        for(int j = 0; j < ncharge[r]; ++j) {
            Q q{};
            q[0] = charge_of(charges[r][j]);
            b->push_back(q, static_cast<int64_t>(std::llround(sdims[r][j])));
        }
        if(conj && conj[r]) b->set_conj_dual();   // :54-60: SET_CONJ() marks the basis (and its trees) as dual WITHOUT conjugating the charges
        b->sort();                                // :61
        legs.push_back(b);
    }
    LayoutPtr L0 = ctx.layout(legs, {});
    std::vector<double> host(static_cast<size_t>(L0->size), 0.0);
    const int vac = L0->find(sym.vacuum());
    for(int64_t bidx = 0; bidx < nblocks; ++bidx) {
        // combined index -> per-leg charge indices, leg 0 fastest (get_indices, :17-33)
        int64_t div = voffs_index[bidx];
        Tree tree;
        tree.rank = nlegs;
        for(int r = 0; r < nlegs; ++r) {
            const int64_t idx = div % ncharge[r];
            div /= ncharge[r];
            tree.unc[r] = Q{};
            tree.unc[r][0] = charge_of(charges[r][idx]);
            tree.dims[r] = static_cast<int64_t>(std::llround(sdims[r][idx]));
            tree.dual[r] = conj && conj[r];
        }
        tree.coupled = sym.vacuum();
        tree.compute_dim();
        // computeIntermediates (:73): Abelian fusion is unique
        Q acc = tree.unc[0];
        for(int r = 1; r + 1 < nlegs; ++r) {
            std::vector<Q> f;
            sym.fuse(acc, tree.unc[r], f);
            acc = f.at(0);
            tree.inter[r - 1] = acc;
        }
        const int64_t begin = voffs_offset[bidx], end = bidx + 1 < nblocks ? voffs_offset[bidx + 1] : nvdat;
        if(begin < 0 || end > nvdat || end < begin) throw std::invalid_argument("charge-block import: block offsets out of range");
        if(vac < 0) throw std::invalid_argument("charge-block import: the legs do not couple to the vacuum");
        int pos = -1;
        const int64_t ro = L0->cdom->left_offset(tree, &pos);
        if(pos < 0) throw std::invalid_argument("charge-block import: a block's charges do not fuse to the vacuum");
        if(end - begin > tree.dim) throw std::invalid_argument("charge-block import: block larger than its sub-block");
        for(int64_t i = begin; i < end; ++i) host[static_cast<size_t>(L0->off[vac] + ro + (i - begin))] = vdat[i];   // setVal(submatrix, i - begin, 0, raw[i])  (:88)
    }
    std::unique_ptr<Tensor> tmp(new Tensor(&ctx, L0, false));
    if(L0->size > 0) {
        check_cuda(cudaMemcpyAsync(tmp->d, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, ctx.stream), "charge-block import: upload");
        check_cuda(cudaStreamSynchronize(ctx.stream), "charge-block import: sync");
    }
    std::vector<int> id(static_cast<size_t>(nlegs));
    for(int i = 0; i < nlegs; ++i) id[static_cast<size_t>(i)] = i;
    return ctx.permute(*tmp, nlegs - rank, id);   // :93
}

}   // namespace xb
