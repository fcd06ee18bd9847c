// Host core implementation. Reference semantics are cited per function (paths under /root/reference).
#include "core.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace xb {

// ------------------------------------------------------------------------------------------------ Symmetry
namespace {

inline double phase(int q) { return (q % 2) != 0 ? -1.0 : 1.0; }   // Symmetry/functions.hpp:76-81

const long double* fact_table()
{
    static long double tab[512];
    static bool init = false;
    if(!init) {
        tab[0] = 1.0L;
        for(int i = 1; i < 512; ++i) tab[i] = tab[i - 1] * static_cast<long double>(i);
        init = true;
    }
    return tab;
}

inline bool triad_ok(int a, int b, int c) { return ((a + b + c) % 2 == 0) && c >= std::abs(a - b) && c <= a + b; }

inline long double delta2(int a, int b, int c)
{
    const long double* F = fact_table();
    return F[(a + b - c) / 2] * F[(a - b + c) / 2] * F[(-a + b + c) / 2] / F[(a + b + c) / 2 + 1];
}

}   // namespace

Symmetry::Symmetry(const std::vector<Factor>& factors) : all_(factors)
{
    for(const auto& f : all_) {
        if(f.kind == Kind::U0) continue;
        act_.push_back(f);
        if(f.kind == Kind::SU2) non_abelian_ = true;
        if(f.fermionic) fermionic_ = true;
    }
    nq_ = static_cast<int>(act_.size());
    if(nq_ > MAXQ) throw std::invalid_argument("too many symmetry factors");
}

std::string Symmetry::signature() const
{
    std::string s;
    for(const auto& f : all_) { s += std::to_string(static_cast<int>(f.kind)) + ":" + std::to_string(f.N) + (f.fermionic ? "f" : "b") + ","; }
    return s;
}

Q Symmetry::vacuum() const
{
    Q q{};
    for(int i = 0; i < nq_; ++i) q[i] = (act_[i].kind == Kind::SU2) ? 1 : 0;   // SU2.hpp:108, U1.hpp:95
    return q;
}

Q Symmetry::conj(const Q& q) const
{
    Q o{};
    for(int i = 0; i < nq_; ++i) {
        switch(act_[i].kind) {
        case Kind::U1: o[i] = -q[i]; break;                                        // U1.hpp:104
        case Kind::ZN: o[i] = ((-q[i]) % act_[i].N + act_[i].N) % act_[i].N; break;   // ZN.hpp:115
        default: o[i] = q[i];                                                       // SU2.hpp:111
        }
    }
    return o;
}

int Symmetry::degeneracy(const Q& q) const
{
    int d = 1;
    for(int i = 0; i < nq_; ++i)
        if(act_[i].kind == Kind::SU2) d *= q[i];   // SU2.hpp:112
    return d;
}

void Symmetry::fuse(const Q& l, const Q& r, std::vector<Q>& out) const
{
    // S1xS2.hpp:217-229: cartesian product, first factor is the outer loop.
    out.clear();
    out.push_back(Q{});
    std::vector<Q> next;
    for(int i = 0; i < nq_; ++i) {
        next.clear();
        for(const Q& base : out) {
            switch(act_[i].kind) {
            case Kind::U1: {
                Q b = base;
                b[i] = l[i] + r[i];   // U1.hpp:189-194
                next.push_back(b);
                break;
            }
            case Kind::ZN: {
                Q b = base;
                b[i] = ((l[i] + r[i]) % act_[i].N + act_[i].N) % act_[i].N;   // ZN.hpp:200-205
                next.push_back(b);
                break;
            }
            default: {
                for(int v = std::abs(l[i] - r[i]) + 1; v <= l[i] + r[i] - 1; v += 2) {   // SU2.cpp:17-27
                    Q b = base;
                    b[i] = v;
                    next.push_back(b);
                }
            }
            }
        }
        out.swap(next);
    }
}

namespace {
inline bool tri(const Factor& f, int a, int b, int c)
{
    switch(f.kind) {
    case Kind::U1: return a + b == c;                                   // U1.hpp:219-224
    case Kind::ZN: return ((a + b) % f.N + f.N) % f.N == c;             // ZN.hpp:232-237
    default: return (c - 1 >= std::abs(a - b)) && (c - 1 <= a + b - 2);   // SU2.cpp:113-118
    }
}
}   // namespace

bool Symmetry::triangle(const Q& a, const Q& b, const Q& c) const
{
    for(int i = 0; i < nq_; ++i)
        if(!tri(act_[i], a[i], b[i], c[i])) return false;
    return true;
}

double Symmetry::coeff_swap(const Q& l, const Q& r, const Q& f) const
{
    double out = 1.0;
    for(int i = 0; i < nq_; ++i) {
        const Factor& F = act_[i];
        if(!tri(F, l[i], r[i], f[i])) return 0.0;
        if(F.kind == Kind::SU2) {   // SU2.hpp:166-174
            double sign = (F.fermionic && ((l[i] - 1) % 2 != 0) && ((r[i] - 1) % 2 != 0)) ? -1.0 : 1.0;
            out *= sign * phase((l[i] + r[i] - f[i] - 1) / 2);
        } else {   // U1.hpp:174-182, ZN.hpp:185-193
            if(F.fermionic && (l[i] % 2 != 0) && (r[i] % 2 != 0)) out = -out;
        }
    }
    return out;
}

double Symmetry::coeff_turn(const Q& l, const Q& r, const Q& f) const
{
    double out = 1.0;
    for(int i = 0; i < nq_; ++i) {
        const Factor& F = act_[i];
        if(!tri(F, l[i], r[i], f[i])) return 0.0;
        if(F.kind == Kind::SU2) {   // SU2.hpp:139-142: coeff_swap * sqrt(qf/ql)
            double sign = (F.fermionic && ((l[i] - 1) % 2 != 0) && ((r[i] - 1) % 2 != 0)) ? -1.0 : 1.0;
            out *= sign * phase((l[i] + r[i] - f[i] - 1) / 2) * std::sqrt(static_cast<double>(f[i]) / static_cast<double>(l[i]));
        }   // U1.hpp:158: 1
    }
    return out;
}

double Symmetry::coeff_FS(const Q& q) const
{
    double out = 1.0;
    for(int i = 0; i < nq_; ++i)
        if(act_[i].kind == Kind::SU2 && q[i] % 2 == 0) out = -out;   // SU2.cpp:35-39
    return out;
}

double Symmetry::coeff_twist(const Q& q) const
{
    double out = 1.0;
    for(int i = 0; i < nq_; ++i) {
        const Factor& F = act_[i];
        if(!F.fermionic) continue;
        if(F.kind == Kind::SU2) {
            if((q[i] - 1) % 2 != 0) out = -out;   // SU2.hpp:122-126
        } else if(q[i] % 2 != 0) {
            out = -out;   // U1.hpp:130-134
        }
    }
    return out;
}

double Symmetry::coeff_rightOrtho(const Q& a, const Q& b) const
{
    double out = 1.0;
    for(int i = 0; i < nq_; ++i)
        if(act_[i].kind == Kind::SU2) out *= static_cast<double>(a[i]) / static_cast<double>(b[i]);   // SU2.hpp:135
    return out;
}

double Symmetry::su2_6j(int a, int b, int c, int d, int e, int f) const
{
    // {a/2 b/2 c/2; d/2 e/2 f/2}: Racah's single-sum formula; GSL's gsl_sf_coupling_6j is what the
    // reference calls (SU2Wrappers.hpp:73-76).
    if(a < 0 || b < 0 || c < 0 || d < 0 || e < 0 || f < 0) return 0.0;
    if(!(triad_ok(a, b, c) && triad_ok(a, e, f) && triad_ok(d, b, f) && triad_ok(d, e, c))) return 0.0;
    uint64_t key = 0;
    for(int v : {a, b, c, d, e, f}) key = (key << 10) | static_cast<uint64_t>(v & 1023);
    auto it = sixj_cache_.find(key);
    if(it != sixj_cache_.end()) return it->second;
    const long double* F = fact_table();
    const int t1 = (a + b + c) / 2, t2 = (a + e + f) / 2, t3 = (d + b + f) / 2, t4 = (d + e + c) / 2;
    const int u1 = (a + b + d + e) / 2, u2 = (b + c + e + f) / 2, u3 = (a + c + d + f) / 2;
    const int lo = std::max(std::max(t1, t2), std::max(t3, t4)), hi = std::min(u1, std::min(u2, u3));
    long double s = 0.0L;
    for(int t = lo; t <= hi; ++t) {
        long double term = F[t + 1] / (F[t - t1] * F[t - t2] * F[t - t3] * F[t - t4] * F[u1 - t] * F[u2 - t] * F[u3 - t]);
        s += (t % 2) ? -term : term;
    }
    long double pref = std::sqrt(delta2(a, b, c) * delta2(a, e, f) * delta2(d, b, f) * delta2(d, e, c));
    double val = static_cast<double>(pref * s);
    sixj_cache_.emplace(key, val);
    return val;
}

double Symmetry::coeff_recouple(const Q& q1, const Q& q2, const Q& q3, const Q& Qt, const Q& Q12, const Q& Q23) const
{
    double out = 1.0;
    for(int i = 0; i < nq_; ++i) {
        const Factor& F = act_[i];
        if(F.kind == Kind::SU2) {
            // SU2.hpp:161-164: sqrt(Q12*Q23) * (-1)^((q1+q2+q3+Q)/2) * {q1 q2 Q12; q3 Q Q23}
            double six = su2_6j(q1[i] - 1, q2[i] - 1, Q12[i] - 1, q3[i] - 1, Qt[i] - 1, Q23[i] - 1);
            out *= std::sqrt(static_cast<double>(Q12[i] * Q23[i])) * phase(static_cast<int>((q1[i] + q2[i] + q3[i] + Qt[i]) / 2.)) * six;
        } else {
            // U1.hpp:162-165 + :212-216 (6j = product of four triangle deltas)
            bool ok = tri(F, q1[i], q2[i], Q12[i]) && tri(F, q1[i], Q23[i], Qt[i]) && tri(F, q2[i], q3[i], Q23[i]) && tri(F, Q12[i], q3[i], Qt[i]);
            if(!ok) return 0.0;
        }
        if(out == 0.0) return 0.0;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ Tree
bool Tree::operator==(const Tree& o) const
{
    if(rank != o.rank || !(coupled == o.coupled)) return false;
    for(int i = 0; i < rank; ++i)
        if(!(unc[i] == o.unc[i]) || dims[i] != o.dims[i]) return false;
    for(int i = 0; i < n_inter(); ++i)
        if(!(inter[i] == o.inter[i])) return false;
    return true;
}

bool Tree::operator<(const Tree& o) const
{
    for(int i = 0; i < rank; ++i) {
        if(unc[i] < o.unc[i]) return true;
        if(o.unc[i] < unc[i]) return false;
    }
    for(int i = 0; i < n_inter(); ++i) {
        if(inter[i] < o.inter[i]) return true;
        if(o.inter[i] < inter[i]) return false;
    }
    return coupled < o.coupled;
}

size_t TreeHash::operator()(const Tree& t) const noexcept
{
    size_t h = 1469598103934665603ull ^ static_cast<size_t>(t.rank);
    auto mix = [&h](size_t v) { h = (h ^ v) * 1099511628211ull; };
    QHash qh;
    for(int i = 0; i < t.rank; ++i) {
        mix(qh(t.unc[i]));
        mix(static_cast<size_t>(t.dims[i]));
    }
    for(int i = 0; i < t.n_inter(); ++i) mix(qh(t.inter[i]));
    mix(qh(t.coupled));
    return h;
}

// ------------------------------------------------------------------------------------------------ Qbasis
int Qbasis::find(const Q& q) const
{
    for(size_t i = 0; i < data_.size(); ++i)
        if(data_[i].q == q) return static_cast<int>(i);
    return -1;
}

int64_t Qbasis::inner_dim(const Q& q) const
{
    int i = find(q);
    return i < 0 ? 0 : data_[i].dim;
}

int64_t Qbasis::dim() const
{
    int64_t d = 0;
    for(const auto& e : data_) d += e.dim;
    return d;
}

int64_t Qbasis::full_dim() const
{
    int64_t d = 0;
    for(const auto& e : data_) d += e.dim * sym_->degeneracy(e.q);
    return d;
}

const std::vector<Tree>& Qbasis::trees(const Q& q) const
{
    auto it = trees_.find(q);
    if(it == trees_.end()) throw std::out_of_range("Qbasis::trees: sector not present");
    return it->second;
}

int64_t Qbasis::left_offset(const Tree& t, int* idx) const
{
    const auto& v = trees(t.coupled);
    int64_t off = 0;
    for(size_t i = 0; i < v.size(); ++i) {
        if(v[i] == t) {
            if(idx) *idx = static_cast<int>(i);
            return off;
        }
        off += v[i].dim;
    }
    throw std::out_of_range("Qbasis::left_offset: tree not in basis");
}

void Qbasis::push_back(const Q& q, int64_t inner_dim)
{
    if(depth_ > 1) throw std::logic_error("push_back(q,dim) needs depth 0 or 1");
    if(inner_dim == 0) return;
    sig_.clear();
    int i = find(q);
    if(i < 0) {
        Tree t;
        t.rank = depth_;
        t.coupled = q;
        if(depth_ == 1) {
            t.unc[0] = q;
            t.dual[0] = conj_;
            t.dims[0] = inner_dim;
            t.dim = inner_dim;
        } else {
            t.dim = 1;
        }
        trees_.emplace(q, std::vector<Tree>{t});
        data_.push_back({q, curr_dim_, inner_dim});
    } else {
        data_[i].dim += inner_dim;
        for(size_t j = i + 1; j < data_.size(); ++j) data_[j].num += inner_dim;
        Tree& t = trees_[q][0];
        t.dim += inner_dim;
        t.dims[0] += inner_dim;
    }
    curr_dim_ += inner_dim;
}

void Qbasis::push_back(const Q& q, int64_t inner_dim, std::vector<Tree> trees)
{
    if(find(q) >= 0) throw std::logic_error("push_back(q,dim,trees): sector already present");
    sig_.clear();
    trees_.emplace(q, std::move(trees));
    data_.push_back({q, curr_dim_, inner_dim});
    curr_dim_ += inner_dim;
}

void Qbasis::sort()
{
    if(data_.size() < 2) return;
    sig_.clear();
    std::vector<size_t> idx(data_.size());
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
        if(conj_) return sym_->conj(data_[a].q) < sym_->conj(data_[b].q);   // Qbasis.cpp:315-319
        return data_[a].q < data_[b].q;
    });
    std::vector<Entry> nd(data_.size());
    int64_t acc = 0;
    for(size_t i = 0; i < idx.size(); ++i) {
        nd[i] = data_[idx[i]];
        nd[i].num = acc;
        acc += nd[i].dim;
    }
    data_.swap(nd);
}

void Qbasis::set_conj_dual()
{
    if(depth_ != 1) throw std::logic_error("set_conj_dual() needs a depth-1 basis");
    sig_.clear();
    conj_ = true;
    for(auto& kv : trees_)
        for(Tree& t : kv.second) t.dual[0] = true;
}

std::shared_ptr<Qbasis> Qbasis::conj() const
{
    if(depth_ != 1) throw std::logic_error("conj() needs a depth-1 basis");
    auto out = std::make_shared<Qbasis>(sym_, 1);
    out->conj_ = !conj_;
    for(const auto& e : data_) {
        const Tree& old = trees(e.q)[0];
        Tree t;
        t.rank = 1;
        t.unc[0] = sym_->conj(e.q);
        t.coupled = t.unc[0];
        t.dual[0] = !old.dual[0];
        t.dims[0] = old.dims[0];
        t.dim = old.dim;
        out->push_back(t.coupled, e.dim, std::vector<Tree>{t});
    }
    out->sort();
    return out;
}

std::shared_ptr<Qbasis> Qbasis::forget_history() const
{
    auto out = std::make_shared<Qbasis>(sym_, 1);
    for(const auto& e : data_) out->push_back(e.q, e.dim);
    return out;
}

std::shared_ptr<Qbasis> Qbasis::combine(const Qbasis& other) const
{
    if(other.depth_ != 1) throw std::logic_error("combine: right operand must have depth 1");
    if(depth_ + 1 > MAXR) throw std::length_error("combine: too many legs");
    auto out = std::make_shared<Qbasis>(sym_, depth_ + 1);
    std::unordered_map<Q, std::vector<Tree>, QHash> acc;
    std::vector<Q> order;   // first-seen order (only used to make iteration deterministic)
    std::vector<Q> fused;
    for(const auto& e1 : data_) {
        const auto& t1s = trees(e1.q);
        for(const auto& e2 : other.data_) {
            const Tree& t2 = other.trees(e2.q)[0];
            sym_->fuse(e1.q, e2.q, fused);
            for(const Q& q : fused) {
                auto it = acc.find(q);
                if(it == acc.end()) {
                    it = acc.emplace(q, std::vector<Tree>{}).first;
                    order.push_back(q);
                }
                for(const Tree& t1 : t1s) {
                    // FusionTree::enlarge (FusionTree.cpp:309-324); (t1, e2.q, q) is unique, so the
                    // reference's duplicate scan (Qbasis.cpp:461-467) can never hit.
                    Tree t;
                    t.rank = depth_ + 1;
                    for(int i = 0; i < depth_; ++i) {
                        t.unc[i] = t1.unc[i];
                        t.dims[i] = t1.dims[i];
                        t.dual[i] = t1.dual[i];
                    }
                    t.unc[depth_] = t2.unc[0];
                    t.dims[depth_] = t2.dims[0];
                    t.dual[depth_] = t2.dual[0];
                    for(int i = 0; i < t1.n_inter(); ++i) t.inter[i] = t1.inter[i];
                    if(t.n_inter() > 0) t.inter[t.n_inter() - 1] = t1.coupled;
                    t.coupled = q;
                    t.compute_dim();
                    it->second.push_back(t);
                }
            }
        }
    }
    for(const Q& q : order) {
        auto& v = acc[q];
        std::sort(v.begin(), v.end());
        int64_t d = 0;
        for(const auto& t : v) d += t.dim;
        out->push_back(q, d, std::move(v));
    }
    out->sort();
    return out;
}

bool Qbasis::same_sectors(const Qbasis& o) const
{
    // Order-insensitive: the reference compares the *coupled* bases, which combine() has sorted, so two legs
    // that carry the same (q, dim) set in a different push_back order (e.g. X vs X.conj().conj()) are compatible.
    if(data_.size() != o.data_.size()) return false;
    for(const auto& e : data_)
        if(o.inner_dim(e.q) != e.dim) return false;
    return true;
}

const std::string& Qbasis::signature() const
{
    if(sig_.empty()) {
        std::string s = conj_ ? "c" : "n";
        s += std::to_string(depth_);
        for(const auto& e : data_) {
            for(int i = 0; i < sym_->Nq(); ++i) s += std::to_string(e.q[i]) + ",";
            s += ":" + std::to_string(e.dim);
            if(depth_ == 1) s += trees(e.q)[0].dual[0] ? "d" : "u";
            s += ";";
        }
        sig_ = s;
    }
    return sig_;
}

std::shared_ptr<Qbasis> build_fusion_tree(const SymPtr& sym, const std::vector<BasisPtr>& uncoupled)
{
    if(uncoupled.empty()) {
        auto b = std::make_shared<Qbasis>(sym, 0);
        b->push_back(sym->vacuum(), 1);
        return b;
    }
    std::shared_ptr<Qbasis> cur = std::make_shared<Qbasis>(*uncoupled[0]);
    for(size_t i = 1; i < uncoupled.size(); ++i) cur = cur->combine(*uncoupled[i]);
    return cur;
}

// ------------------------------------------------------------------------------------------------ recoupling
std::vector<int> decompose(const std::vector<int>& pi)
{
    const int N = static_cast<int>(pi.size());
    std::vector<int> a(N), out;
    std::iota(a.begin(), a.end(), 0);
    if(a == pi) return out;
    for(int i = 0; i < N; ++i) {
        int index = static_cast<int>(std::find(a.begin(), a.end(), pi[i]) - a.begin());
        if(index == 0) continue;
        for(int j = index - 1; j >= i; --j) {   // for i == 0 the reference's unsigned loop also stops below 0
            std::swap(a[j], a[j + 1]);
            out.push_back(j);
        }
    }
    assert(a == pi);
    return out;
}

constexpr double EPS = 1.e-12;   // ScalarTraits<double>::epsilon()

void tree_swap(const Symmetry& sym, const Tree& t, int pos, std::vector<std::pair<Tree, double>>& out)
{
    out.clear();
    Tree n = t;
    std::swap(n.unc[pos], n.unc[pos + 1]);
    std::swap(n.dual[pos], n.dual[pos + 1]);
    std::swap(n.dims[pos], n.dims[pos + 1]);
    if(pos == 0) {
        const Q& qf = (t.rank == 2) ? t.coupled : t.inter[0];
        double c = sym.coeff_swap(t.unc[0], t.unc[1], qf);
        if(std::abs(c) < EPS) return;
        out.emplace_back(n, c);
        return;
    }
    const Q& q1 = (pos == 1) ? t.unc[0] : t.inter[pos - 2];
    const Q& q2 = t.unc[pos];
    const Q& q3 = t.unc[pos + 1];
    const Q& Qt = (pos == t.rank - 2) ? t.coupled : t.inter[pos];
    const Q& Q12 = t.inter[pos - 1];
    std::vector<Q> fused;
    sym.fuse(q3, q1, fused);
    const double s1 = sym.coeff_swap(Q12, q3, Qt);
    if(s1 == 0.0) return;
    for(const Q& Q31 : fused) {
        double c = s1 * sym.coeff_recouple(q3, q1, q2, Qt, Q31, Q12) * sym.coeff_swap(q3, q1, Q31);
        if(std::abs(c) < EPS) continue;
        n.inter[pos - 1] = Q31;
        out.emplace_back(n, c);
    }
}

TreeMap tree_permute(const Symmetry& sym, const Tree& t, const std::vector<int>& pi)
{
    TreeMap cur, tmp;
    cur.emplace(t, 1.0);
    std::vector<std::pair<Tree, double>> sw;
    for(int pos : decompose(pi)) {
        tmp.clear();
        for(const auto& [tree, coeff] : cur) {
            tree_swap(sym, tree, pos, sw);
            for(const auto& [t2, c2] : sw) {
                auto it = tmp.find(t2);
                if(it == tmp.end()) {
                    tmp.emplace(t2, coeff * c2);
                } else if(std::abs(it->second + coeff * c2) < EPS) {
                    tmp.erase(it);   // FusionTree.cpp:338-340
                } else {
                    it->second += coeff * c2;
                }
            }
        }
        cur.swap(tmp);
    }
    return cur;
}

namespace {

// treepair.cpp:35-88. Returns false if the coefficient vanishes.
bool turn_right(const Symmetry& sym, const Tree& t1, const Tree& t2, Tree& t1p, Tree& t2p, double& coeff)
{
    const int R = t1.rank, C = t2.rank;
    Q a;
    if(C == 1)
        a = sym.vacuum();
    else if(C == 2)
        a = t2.unc[0];
    else
        a = t2.inter[C - 3];
    const Q& b = t2.unc[C - 1];
    coeff = sym.coeff_turn(a, b, t2.coupled);
    if(t2.dual[C - 1]) coeff *= sym.coeff_FS(sym.conj(b));
    if(std::abs(coeff) < EPS) return false;
    t1p = Tree{};
    t1p.rank = R + 1;
    for(int i = 0; i < R; ++i) {
        t1p.unc[i] = t1.unc[i];
        t1p.dims[i] = t1.dims[i];
        t1p.dual[i] = t1.dual[i];
    }
    t1p.unc[R] = sym.conj(b);
    t1p.dims[R] = t2.dims[C - 1];
    t1p.dual[R] = !t2.dual[C - 1];
    for(int i = 0; i < t1.n_inter(); ++i) t1p.inter[i] = t1.inter[i];
    if(R + 1 > 2) t1p.inter[R - 2] = t1.coupled;
    t1p.coupled = a;
    t1p.compute_dim();
    t2p = Tree{};
    t2p.rank = C - 1;
    for(int i = 0; i < C - 1; ++i) {
        t2p.unc[i] = t2.unc[i];
        t2p.dims[i] = t2.dims[i];
        t2p.dual[i] = t2.dual[i];
    }
    for(int i = 0; i < t2p.n_inter(); ++i) t2p.inter[i] = t2.inter[i];
    t2p.coupled = a;
    t2p.compute_dim();
    return true;
}

// treepair.cpp:103-141 with shift < 0 (bend codomain legs into the domain) or > 0 (the reverse).
bool turn(const Symmetry& sym, int shift, Tree& t1, Tree& t2, double& coeff)
{
    coeff = 1.0;
    Tree a, b;
    double c;
    while(shift < 0) {
        if(!turn_right(sym, t1, t2, a, b, c)) return false;
        t1 = a;
        t2 = b;
        coeff *= c;
        ++shift;
    }
    while(shift > 0) {   // turn_left(t1,t2) = turn_right(t2,t1) with the roles swapped back
        if(!turn_right(sym, t2, t1, a, b, c)) return false;
        t2 = a;
        t1 = b;
        coeff *= c;
        --shift;
    }
    return true;
}

Tree abelian_tree(const Symmetry& sym, int n, const Q* qs, const int64_t* dims, const bool* duals)
{
    // FusionTree::computeIntermediates (FusionTree.cpp:133-150)
    Tree t;
    t.rank = n;
    if(n == 0) {
        t.coupled = sym.vacuum();
        t.dim = 1;
        return t;
    }
    std::vector<Q> f;
    Q acc = qs[0];
    for(int i = 0; i < n; ++i) {
        t.unc[i] = qs[i];
        t.dims[i] = dims[i];
        t.dual[i] = duals[i];
        if(i > 0) {
            sym.fuse(acc, qs[i], f);
            acc = f.front();
            if(i < n - 1) t.inter[i - 1] = acc;
        }
    }
    t.coupled = acc;
    t.compute_dim();
    return t;
}

}   // namespace

PairMap treepair_permute(const Symmetry& sym, int shift, const Tree& t1, const Tree& t2, const std::vector<int>& pi)
{
    const int R = t1.rank, C = t2.rank, N = R + C;
    const int newR = R - shift, newC = C + shift;
    PairMap out;
    if(!sym.non_abelian() && !sym.fermionic()) {   // treepair.cpp:147-168
        Q qs[MAXR];
        int64_t ds[MAXR];
        bool du[MAXR];
        auto leg = [&](int k, bool to_domain, int slot) {
            if(pi[k] < R) {
                const int s = pi[k];
                qs[slot] = to_domain ? t1.unc[s] : sym.conj(t1.unc[s]);
                ds[slot] = t1.dims[s];
                du[slot] = to_domain ? t1.dual[s] : !t1.dual[s];
            } else {
                const int s = pi[k] - R;
                qs[slot] = to_domain ? sym.conj(t2.unc[s]) : t2.unc[s];
                ds[slot] = t2.dims[s];
                du[slot] = to_domain ? !t2.dual[s] : t2.dual[s];
            }
        };
        for(int r = 0; r < newR; ++r) leg(r, true, r);
        Tree od = abelian_tree(sym, newR, qs, ds, du);
        for(int c = 0; c < newC; ++c) leg(c + newR, false, c);
        Tree oc = abelian_tree(sym, newC, qs, ds, du);
        out.emplace(TreePair{od, oc}, 1.0);
        return out;
    }
    // general path, treepair.cpp:178-205
    std::vector<int> pi_id(N), pi_tmp(N), pi_corr(N);
    std::iota(pi_id.begin(), pi_id.end(), 0);
    for(int i = 0; i < C; ++i) pi_id[i + R] = (C - 1) - i + R;
    for(int i = 0; i < N; ++i) pi_tmp[i] = pi_id[pi[i]];
    for(int i = 0; i < newR; ++i) pi_corr[i] = pi_tmp[i];
    for(int i = 0; i < newC; ++i) pi_corr[i + newR] = pi_tmp[(newC - 1) - i + newR];

    Tree a = t1, b = t2;
    double c1;
    if(!turn(sym, -C, a, b, c1)) return out;
    for(const auto& [t1pp, c2] : tree_permute(sym, a, pi_corr)) {
        Tree x = t1pp, y = b;
        double c3;
        if(!turn(sym, newC, x, y, c3)) continue;
        out[TreePair{x, y}] += c1 * c2 * c3;
    }
    return out;
}

}   // namespace xb
