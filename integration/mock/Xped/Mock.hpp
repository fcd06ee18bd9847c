// Minimal stand-ins for the pieces of the reference the shim touches (Qbasis<Symmetry,1>, qarray, symmetry tags), just enough to COMPILE
// integration/TensorB200.hpp without the reference's dependencies (Eigen, Boost, GSL, yas, HighFive, stan-math are absent from this image).
// Interfaces follow include/Xped/Symmetry/qarray.hpp:29-80, include/Xped/Symmetry/{U0,U1,SU2}.hpp and include/Xped/Core/Qbasis.hpp:37-268
// (push_back, Nq, dim, IS_CONJ, conj); they carry no fusion logic -- that lives in libxped_b200.so.
#pragma once
#include <array>
#include <cstddef>
#include <utility>
#include <vector>

namespace Xped {

template <std::size_t Nq> using qarray = std::array<int, Nq>;

namespace Sym {
struct SpinSU2 {};
struct SpinU1 {};
struct FChargeU1 {};
struct U0 { static constexpr std::size_t Nq = 0; using qType = qarray<0>; };
template <typename Kind> struct U1 { static constexpr std::size_t Nq = 1; using qType = qarray<1>; };
template <typename Kind> struct SU2 { static constexpr std::size_t Nq = 1; using qType = qarray<1>; };
template <typename Kind, int N> struct ZN { static constexpr std::size_t Nq = 1; using qType = qarray<1>; static constexpr int MOD = N; };
}   // namespace Sym

template <typename Symmetry, std::size_t depth = 1> class Qbasis
{
public:
    using qType = typename Symmetry::qType;
    void push_back(const qType& q, std::size_t inner_dim) { data_.emplace_back(q, inner_dim); }
    std::size_t Nq() const { return data_.size(); }
    std::size_t dim() const
    {
        std::size_t d = 0;
        for(const auto& e : data_) d += e.second;
        return d;
    }
    bool IS_CONJ() const { return conj_; }
    Qbasis conj() const
    {
        Qbasis out = *this;
        out.conj_ = !conj_;
        return out;
    }
    const std::vector<std::pair<qType, std::size_t>>& entries() const { return data_; }

private:
    std::vector<std::pair<qType, std::size_t>> data_;
    bool conj_ = false;
};

}   // namespace Xped
