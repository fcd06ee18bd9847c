// xped_b200 host core: symmetry sectors, graded bases, fusion trees and recoupling coefficients.
//
// These are the host-side objects the reference builds with compile-time templates
// (include/Xped/Symmetry/*.hpp, include/Xped/Core/{Qbasis,FusionTree,treepair}.hpp).  Here they are
// runtime-polymorphic, immutable once built and shared by pointer, so that layouts and permute plans
// can be cached by signature and uploaded to the device once (BASELINE.json north_star, "Host and storage").
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace xb {

constexpr int MAXQ = 4;   // quantum numbers per sector label (product symmetries)
constexpr int MAXR = 8;   // legs per fusion tree (CTM needs 7: Tensor<4,3> bent to one side)

using Q = std::array<int, MAXQ>;   // zero padded beyond Nq -> lexicographic compare == qarray<Nq> compare

struct QHash {
    size_t operator()(const Q& q) const noexcept
    {
        size_t h = 1469598103934665603ull;
        for(int v : q) h = (h ^ static_cast<size_t>(static_cast<uint32_t>(v))) * 1099511628211ull;
        return h;
    }
};

enum class Kind : int { U0 = 0, U1 = 1, ZN = 2, SU2 = 3 };

struct Factor
{
    Kind kind;
    int N;            // modulus for ZN
    bool fermionic;   // Sym::FChargeU1 / FChargeSU2 style kinds (include/Xped/Symmetry/kind_dummies.hpp)
};

// Runtime restatement of SymBase (include/Xped/Symmetry/SymBase.hpp:14-102) + U0/U1/ZN/SU2/S1xS2.
class Symmetry
{
public:
    explicit Symmetry(const std::vector<Factor>& factors);

    int Nq() const { return nq_; }
    bool non_abelian() const { return non_abelian_; }
    bool fermionic() const { return fermionic_; }
    const std::vector<Factor>& factors() const { return all_; }

    Q vacuum() const;
    Q conj(const Q& q) const;
    int degeneracy(const Q& q) const;
    void fuse(const Q& l, const Q& r, std::vector<Q>& out) const;   // basis_combine; first factor outermost
    bool triangle(const Q& a, const Q& b, const Q& c) const;

    double coeff_swap(const Q& l, const Q& r, const Q& f) const;
    double coeff_turn(const Q& l, const Q& r, const Q& f) const;
    double coeff_FS(const Q& q) const;
    double coeff_twist(const Q& q) const;
    double coeff_rightOrtho(const Q& a, const Q& b) const;
    double coeff_recouple(const Q& q1, const Q& q2, const Q& q3, const Q& Qt, const Q& Q12, const Q& Q23) const;

    std::string signature() const;

private:
    std::vector<Factor> all_;   // as given (U0 factors carry no label)
    std::vector<Factor> act_;   // label-carrying factors, act_[i] <-> q[i]
    int nq_ = 0;
    bool non_abelian_ = false, fermionic_ = false;
    mutable std::unordered_map<uint64_t, double> sixj_cache_;
    double su2_6j(int a, int b, int c, int d, int e, int f) const;   // arguments 2j
};
using SymPtr = std::shared_ptr<const Symmetry>;

// One fusion tree (include/Xped/Core/FusionTree.hpp:22-115), multiplicity-free symmetries only.
struct Tree
{
    int rank = 0;
    std::array<Q, MAXR> unc{};
    std::array<Q, MAXR> inter{};   // rank-2 used
    Q coupled{};
    std::array<int64_t, MAXR> dims{};
    std::array<bool, MAXR> dual{};
    int64_t dim = 1;

    int n_inter() const { return rank >= 2 ? rank - 2 : 0; }
    void compute_dim()
    {
        dim = 1;
        for(int i = 0; i < rank; ++i) dim *= dims[i];
    }
    bool operator==(const Tree& o) const;   // FusionTree.cpp:66-85: IS_DUAL ignored
    bool operator<(const Tree& o) const;    // FusionTree.cpp:32-64
};

struct TreeHash { size_t operator()(const Tree& t) const noexcept; };
using TreeMap = std::unordered_map<Tree, double, TreeHash>;

struct TreePair
{
    Tree d, c;
    bool operator==(const TreePair& o) const { return d == o.d && c == o.c; }
};
struct TreePairHash
{
    size_t operator()(const TreePair& p) const noexcept { return TreeHash{}(p.d) * 1000003u ^ TreeHash{}(p.c); }
};
using PairMap = std::unordered_map<TreePair, double, TreePairHash>;

// Graded basis with fusion history (include/Xped/Core/Qbasis.hpp:37-268).
class Qbasis
{
public:
    struct Entry { Q q; int64_t num; int64_t dim; };

    Qbasis(SymPtr sym, int depth) : sym_(std::move(sym)), depth_(depth) {}

    const SymPtr& sym() const { return sym_; }
    int depth() const { return depth_; }
    bool is_conj() const { return conj_; }
    const std::vector<Entry>& entries() const { return data_; }
    int find(const Q& q) const;   // index or -1
    int64_t inner_dim(const Q& q) const;
    int64_t dim() const;
    int64_t full_dim() const;
    const std::vector<Tree>& trees(const Q& q) const;
    // Qbasis.cpp:208-225; also returns the matched tree's position through `idx`.
    int64_t left_offset(const Tree& t, int* idx = nullptr) const;

    void push_back(const Q& q, int64_t inner_dim);                              // Qbasis.cpp:34-74
    void push_back(const Q& q, int64_t inner_dim, std::vector<Tree> trees);     // Qbasis.cpp:76-91
    void sort();                                                                // Qbasis.cpp:306-331
    void set_conj_dual();   // SET_CONJ() + IS_DUAL[0] = true on every tree, charges untouched (what IO::loadMatlabTensor does to a conjugated leg, IO/Matlab.hpp:54-60)
    std::shared_ptr<Qbasis> conj() const;                                       // Qbasis.cpp:418-439
    std::shared_ptr<Qbasis> combine(const Qbasis& other) const;                 // Qbasis.cpp:441-495
    std::shared_ptr<Qbasis> forget_history() const;                             // Qbasis.cpp:410-416

    bool same_sectors(const Qbasis& o) const;   // Qbasis::operator== on the (q, dim) sequence (Qbasis.cpp:333-353)
    const std::string& signature() const;   // identifies (q,dim,dual) sequence + conj flag; cached

private:
    SymPtr sym_;
    int depth_;
    std::vector<Entry> data_;
    std::unordered_map<Q, std::vector<Tree>, QHash> trees_;
    int64_t curr_dim_ = 0;
    bool conj_ = false;
    mutable std::string sig_;
};
using BasisPtr = std::shared_ptr<const Qbasis>;

// include/Xped/Util/Permutations.hpp:195-225
std::vector<int> decompose(const std::vector<int>& pi);

// src/Core/FusionTree.cpp:353-412 / :326-351
void tree_swap(const Symmetry& sym, const Tree& t, int pos, std::vector<std::pair<Tree, double>>& out);
TreeMap tree_permute(const Symmetry& sym, const Tree& t, const std::vector<int>& pi);

// src/Core/treepair.cpp:143-217: list of (target pair, coefficient) for one source pair.
PairMap treepair_permute(const Symmetry& sym, int shift, const Tree& t1, const Tree& t2, const std::vector<int>& pi);

// src/Core/TensorHelper.cpp:15-45
std::shared_ptr<Qbasis> build_fusion_tree(const SymPtr& sym, const std::vector<BasisPtr>& uncoupled);

}   // namespace xb
