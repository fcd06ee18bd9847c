// The reference-side binding of libxped_b200.so: `Xped::Tensor<double, Rank, CoRank, Symmetry>` with the public interface of
// include/Xped/Core/Tensor.hpp:59-444 / TensorBase.hpp:34-148, forwarding every method to the C ABI (include/xped_b200.h).  A maintainer drops this header
// in as include/Xped/Core/TensorB200.hpp and selects it with XPED_STORAGE=DeviceArena.  Compiled and exercised in this repository against the
// mock of Qbasis / the symmetry tags in integration/mock (tests/test_integration_shim.py).
#pragma once
#include <xped_b200.h>

#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <vector>

namespace Xped {

inline void b200_check(int st)
{   // the reference asserts / VERIFYs (TensorBase.cpp:253-254): the device backend throws instead
    if(st != XB_OK) throw std::runtime_error(xb_last_error());
}

// symmetry tag -> xb_sym_factor list (include/Xped/Symmetry/*.hpp)
template <typename Symmetry> struct b200_factors;
template <> struct b200_factors<Sym::U0> { static std::vector<xb_sym_factor> get() { return {{XB_SYM_U0, 0, 0}}; } };
template <typename Kind> struct b200_factors<Sym::U1<Kind>> { static std::vector<xb_sym_factor> get() { return {{XB_SYM_U1, 0, std::is_same<Kind, Sym::FChargeU1>::value ? 1 : 0}}; } };
template <typename Kind> struct b200_factors<Sym::SU2<Kind>> { static std::vector<xb_sym_factor> get() { return {{XB_SYM_SU2, 0, 0}}; } };
template <typename Kind, int N> struct b200_factors<Sym::ZN<Kind, N>> { static std::vector<xb_sym_factor> get() { return {{XB_SYM_ZN, N, std::is_same<Kind, Sym::FChargeU1>::value ? 1 : 0}}; } };

template <typename Symmetry> struct B200Context
{   // one per host thread / GPU; replaces the mpi::XpedWorld plumbing (include/Xped/Util/Mpi.hpp:17-47)
    xb_ctx* h = nullptr;
    static B200Context& get()
    {
        static thread_local B200Context c = [] {
            B200Context r;
            auto f = b200_factors<Symmetry>::get();
            b200_check(xb_ctx_create(&r.h, /*device*/ 0, f.data(), static_cast<int>(f.size())));
            return r;
        }();
        return c;
    }
};

template <typename Symmetry> std::shared_ptr<xb_basis> to_xb(const Qbasis<Symmetry, 1>& b)
{   // Qbasis<Symmetry,1> (Qbasis.hpp:37-268) -> handle; a conjugated basis is the conj() of its pre-image
    std::vector<int> q;
    std::vector<int64_t> d;
    for(const auto& e : b.entries()) {
        for(int x : e.first) q.push_back(x);
        d.push_back(static_cast<int64_t>(e.second));
    }
    if(q.empty()) q.push_back(0);
    xb_basis* out = nullptr;
    b200_check(xb_basis_create(B200Context<Symmetry>::get().h, q.data(), d.data(), static_cast<int>(d.size()), &out));
    if(b.IS_CONJ()) {
        xb_basis* c = nullptr;
        b200_check(xb_basis_conj(out, &c));
        xb_basis_free(out);
        out = c;
    }
    return std::shared_ptr<xb_basis>(out, [](xb_basis* p) { xb_basis_free(p); });
}

template <typename Scalar, std::size_t Rank, std::size_t CoRank, typename Symmetry> class Tensor;

template <std::size_t Rank, std::size_t CoRank, typename Symmetry> class Tensor<double, Rank, CoRank, Symmetry>
{
public:
    using Scalar = double;
    using qType = typename Symmetry::qType;

    Tensor() = default;
    // Tensor(basis_domain, basis_codomain) + setZero()  (Tensor.hpp:106-110, Tensor.cpp:46-55)
    Tensor(const std::array<Qbasis<Symmetry, 1>, Rank>& dom, const std::array<Qbasis<Symmetry, 1>, CoRank>& cod)
    {
        std::vector<std::shared_ptr<xb_basis>> keep;
        std::vector<xb_basis*> d, c;
        for(const auto& b : dom) { keep.push_back(to_xb(b)); d.push_back(keep.back().get()); }
        for(const auto& b : cod) { keep.push_back(to_xb(b)); c.push_back(keep.back().get()); }
        xb_tensor* t = nullptr;
        b200_check(xb_tensor_create(ctx(), d.data(), static_cast<int>(Rank), c.data(), static_cast<int>(CoRank), &t));
        adopt_handle(t);
    }
    static Tensor adopt(xb_tensor* t)
    {
        Tensor out;
        out.adopt_handle(t);
        return out;
    }
    Tensor(const Tensor& o) { *this = o; }   // value semantics: deep copy (Tensor.hpp copies its std::vector)
    Tensor& operator=(const Tensor& o)
    {
        if(this != &o && o.h_) {
            xb_tensor* t = nullptr;
            b200_check(xb_tensor_clone(o.h_.get(), &t));
            adopt_handle(t);
        }
        return *this;
    }
    Tensor(Tensor&&) = default;
    Tensor& operator=(Tensor&&) = default;

    xb_tensor* handle() const { return h_.get(); }
    static constexpr std::size_t rank() { return Rank; }
    static constexpr std::size_t corank() { return CoRank; }

    void setZero() { b200_check(xb_tensor_set_zero(h_.get())); host_.clear(); }
    void setIdentity() { b200_check(xb_tensor_set_identity(h_.get())); host_.clear(); }
    void setConstant(double v) { b200_check(xb_tensor_set_constant(h_.get(), v)); host_.clear(); }
    void setRandom(std::mt19937& engine)
    {   // Tensor.cpp:33-44: uniform [-1, 1] from the caller's engine, filled on the host and uploaded
        std::uniform_real_distribution<double> dist(-1.0, 1.0);
        std::vector<double> v(plainSize());
        for(auto& x : v) x = dist(engine);
        set_data(v.data(), v.size());
    }
    std::size_t plainSize() const { return static_cast<std::size_t>(xb_tensor_plain_size(h_.get())); }   // StorageType_contiguous.hpp:149
    const double* data() const
    {   // Tensor.hpp:137-152: host image of m_data, downloaded on demand
        host_.resize(plainSize());
        b200_check(xb_tensor_download(h_.get(), host_.data(), static_cast<int64_t>(host_.size())));
        return host_.data();
    }
    void set_data(const double* p, std::size_t n)
    {
        b200_check(xb_tensor_upload(h_.get(), p, static_cast<int64_t>(n)));
        host_.clear();
    }
    std::size_t num_sectors() const { return static_cast<std::size_t>(xb_tensor_num_sectors(h_.get())); }

    template <int shift, std::size_t... p> Tensor<double, Rank - shift, CoRank + shift, Symmetry> permute() const
    {   // Tensor.cpp:335-355
        static_assert(sizeof...(p) == Rank + CoRank, "permutation length");
        const int pi[] = {static_cast<int>(p)..., 0};
        xb_tensor* out = nullptr;
        b200_check(xb_permute(ctx(), h_.get(), shift, pi, static_cast<int>(sizeof...(p)), &out));
        return Tensor<double, Rank - shift, CoRank + shift, Symmetry>::adopt(out);
    }
    template <int shift, std::size_t... p> Tensor<double, Rank + shift, CoRank - shift, Symmetry> permute_adj() const
    {   // Tensor.cpp:258-333 (applied to a tensor of the output type of permute<shift, p...>)
        const int pi[] = {static_cast<int>(p)..., 0};
        xb_tensor* out = nullptr;
        b200_check(xb_permute_adj(ctx(), h_.get(), shift, pi, static_cast<int>(sizeof...(p)), &out));
        return Tensor<double, Rank + shift, CoRank - shift, Symmetry>::adopt(out);
    }
    // contract<a1, a2, ResRank>(other)  (Tensor.hpp:328-351): label arrays as template-free arguments here (C++17: no class-type NTTPs)
    template <std::size_t ResRank, std::size_t ResCoRank, std::size_t OR, std::size_t OC>
    Tensor<double, ResRank, ResCoRank, Symmetry> contract(const std::array<int, Rank + CoRank>& a1, const Tensor<double, OR, OC, Symmetry>& other,
                                                          const std::array<int, OR + OC>& a2) const
    {
        xb_tensor* out = nullptr;
        b200_check(xb_contract(ctx(), h_.get(), a1.data(), static_cast<int>(a1.size()), other.handle(), a2.data(), static_cast<int>(a2.size()),
                               static_cast<int>(ResRank), &out));
        return Tensor<double, ResRank, ResCoRank, Symmetry>::adopt(out);
    }
    Tensor<double, CoRank, Rank, Symmetry> adjoint() const
    {   // TensorBase.cpp:82-86 + AdjointOp::eval (AdjointOp.hpp:44-65)
        xb_tensor* out = nullptr;
        b200_check(xb_adjoint(ctx(), h_.get(), &out));
        return Tensor<double, CoRank, Rank, Symmetry>::adopt(out);
    }
    Tensor twist(std::size_t leg) const
    {   // Tensor.cpp:357-377
        xb_tensor* out = nullptr;
        b200_check(xb_twist(ctx(), h_.get(), static_cast<int>(leg), &out));
        return adopt(out);
    }
    Tensor operator*(double s) const
    {   // TensorBase.hpp:197-209
        xb_tensor* out = nullptr;
        b200_check(xb_axpby(ctx(), s, h_.get(), 0.0, nullptr, &out));
        return adopt(out);
    }
    Tensor operator+(const Tensor& o) const
    {   // CoeffBinaryOp.hpp:44-64
        xb_tensor* out = nullptr;
        b200_check(xb_axpby(ctx(), 1.0, h_.get(), 1.0, o.h_.get(), &out));
        return adopt(out);
    }
    Tensor operator-(const Tensor& o) const
    {
        xb_tensor* out = nullptr;
        b200_check(xb_axpby(ctx(), 1.0, h_.get(), -1.0, o.h_.get(), &out));
        return adopt(out);
    }
    double maxNorm() const { double v; b200_check(xb_max_norm(ctx(), h_.get(), &v)); return v; }          // TensorBase.cpp:48-56
    double trace() const { double v; b200_check(xb_trace(ctx(), h_.get(), &v)); return v; }               // TensorBase.cpp:27-37
    double squaredNorm() const { double v; b200_check(xb_squared_norm(ctx(), h_.get(), &v)); return v; }  // TensorBase.cpp:58-64
    double norm() const { return std::sqrt(squaredNorm()); }

    // tSVD(maxKeep, eps_svd, truncWeight, entropy, PRESERVE_MULTIPLETS)  (Tensor.cpp:478-633)
    std::tuple<Tensor<double, Rank, 1, Symmetry>, Tensor<double, 1, 1, Symmetry>, Tensor<double, 1, CoRank, Symmetry>>
    tSVD(std::size_t maxKeep, double eps_svd, double& truncWeight, double& entropy, bool PRESERVE_MULTIPLETS = true) const
    {
        xb_tensor *U = nullptr, *S = nullptr, *Vd = nullptr;
        b200_check(xb_tsvd(ctx(), h_.get(), static_cast<int64_t>(maxKeep), eps_svd, PRESERVE_MULTIPLETS ? 1 : 0, &U, &S, &Vd, &truncWeight, &entropy));
        return std::make_tuple(Tensor<double, Rank, 1, Symmetry>::adopt(U), Tensor<double, 1, 1, Symmetry>::adopt(S), Tensor<double, 1, CoRank, Symmetry>::adopt(Vd));
    }
    std::pair<Tensor<double, Rank, 1, Symmetry>, Tensor<double, 1, CoRank, Symmetry>> tQR(bool RETURN_LQ = false) const
    {   // Tensor.cpp:757-780
        xb_tensor *Q = nullptr, *R = nullptr;
        b200_check(xb_tqr(ctx(), h_.get(), RETURN_LQ ? 1 : 0, &Q, &R));
        return {Tensor<double, Rank, 1, Symmetry>::adopt(Q), Tensor<double, 1, CoRank, Symmetry>::adopt(R)};
    }
    std::pair<Tensor<double, 1, 1, Symmetry>, Tensor<double, Rank, 1, Symmetry>> teigh() const
    {   // Tensor.cpp:782-803
        xb_tensor *D = nullptr, *V = nullptr;
        b200_check(xb_teigh(ctx(), h_.get(), &D, &V));
        return {Tensor<double, 1, 1, Symmetry>::adopt(D), Tensor<double, Rank, 1, Symmetry>::adopt(V)};
    }

    static xb_ctx* ctx() { return B200Context<Symmetry>::get().h; }

private:
    void adopt_handle(xb_tensor* t)
    {
        h_ = std::shared_ptr<xb_tensor>(t, [](xb_tensor* p) { xb_tensor_free(p); });
        host_.clear();
    }
    std::shared_ptr<xb_tensor> h_;
    mutable std::vector<double> host_;
};

// TensorBase::operator*  (TensorBase.cpp:243-288)
template <std::size_t R, std::size_t M, std::size_t C, typename Symmetry>
Tensor<double, R, C, Symmetry> operator*(const Tensor<double, R, M, Symmetry>& a, const Tensor<double, M, C, Symmetry>& b)
{
    xb_tensor* out = nullptr;
    b200_check(xb_multiply(B200Context<Symmetry>::get().h, a.handle(), 0, b.handle(), 0, &out));
    return Tensor<double, R, C, Symmetry>::adopt(out);
}

// decompose(T1, T2, chi)  (src/PEPS/PEPSContractions.cpp:18-43)
template <std::size_t R, std::size_t C, typename Symmetry>
std::pair<Tensor<double, 1, R, Symmetry>, Tensor<double, C, 1, Symmetry>> decompose(const Tensor<double, R, C, Symmetry>& T1, const Tensor<double, C, R, Symmetry>& T2,
                                                                                 std::size_t chi)
{
    xb_tensor *P1 = nullptr, *P2 = nullptr;
    b200_check(xb_decompose(B200Context<Symmetry>::get().h, T1.handle(), T2.handle(), static_cast<int64_t>(chi), &P1, &P2));
    return {Tensor<double, 1, R, Symmetry>::adopt(P1), Tensor<double, C, 1, Symmetry>::adopt(P2)};
}

}   // namespace Xped
