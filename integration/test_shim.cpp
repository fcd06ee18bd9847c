// Compiles the reference-side shim (integration/TensorB200.hpp) against the mock of the reference's Qbasis / symmetry tags and, on a machine with a GPU,
// runs the checks of tests/tensor_tests.hpp:2-33 (permute is inverted by the inverse permutation) plus product / tSVD invariants through it.
// Exit code 0 = all good; 77 = no CUDA device (compile + link + symbol binding were still exercised).
#include <cmath>
#include <cstdio>

#include <Xped/Mock.hpp>
#include <TensorB200.hpp>

using namespace Xped;
using Symmetry = Sym::SU2<Sym::SpinSU2>;

static double rel_diff(const Tensor<double, 2, 2, Symmetry>& a, const Tensor<double, 2, 2, Symmetry>& b)
{
    return (a - b).norm() / b.norm();
}

int main()
{
    xb_ctx* probe = nullptr;
    xb_sym_factor f{XB_SYM_SU2, 0, 0};
    const int st = xb_ctx_create(&probe, 0, &f, 1);
    if(st == XB_ERR_NO_DEVICE) {
        std::printf("no CUDA device: shim compiled and linked, nothing to run\n");
        return 77;
    }
    if(st != XB_OK) {
        std::printf("xb_ctx_create: %s\n", xb_last_error());
        return 1;
    }
    xb_ctx_free(probe);
    Qbasis<Symmetry, 1> B, C;
    B.push_back({1}, 2);
    B.push_back({2}, 3);
    B.push_back({3}, 2);
    C.push_back({2}, 2);
    C.push_back({3}, 1);
    std::mt19937 eng(7);
    Tensor<double, 2, 2, Symmetry> t({{B, C}}, {{B, C}});
    t.setRandom(eng);
    // permute and its inverse (tests/tensor_tests.hpp:2-33)
    auto p = t.permute<1, 2, 0, 3, 1>();          // Tensor<1,3>
    auto back = p.permute<-1, 1, 3, 0, 2>();      // Tensor<2,2>
    const double e1 = rel_diff(back, t);
    // operator*, adjoint, trace: squaredNorm = (T * T^+).trace()  (TensorBase.cpp:58-64)
    auto prod = t * t.adjoint();
    const double e2 = std::abs(prod.trace() - t.squaredNorm()) / t.squaredNorm();
    // contract == permute + operator* (Tensor.hpp:328-351)
    auto c1 = t.contract<2, 2>({-1, -2, 1, 2}, t, {1, 2, -3, -4});
    const double e3 = rel_diff(c1, t * t);
    // full tSVD reconstructs the tensor
    double tw = 0.0, ent = 0.0;
    auto [U, S, Vd] = t.tSVD(1000, 0.0, tw, ent, false);
    const double e4 = rel_diff((U * S) * Vd, t);
    auto [Q, R] = t.tQR();
    const double e5 = rel_diff(Q * R, t);
    std::printf("permute round trip %.2e, squaredNorm %.2e, contract %.2e, tSVD %.2e, tQR %.2e, truncWeight %.2e\n", e1, e2, e3, e4, e5, tw);
    return (e1 < 1e-13 && e2 < 1e-13 && e3 < 1e-13 && e4 < 1e-12 && e5 < 1e-12 && tw < 1e-20) ? 0 : 2;
}
