// teigh / tQR / SVD reverse rule / MPS transfer contractions. (Filled in after the first GPU bring-up.)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "engine.hpp"

namespace xb {

// ------------------------------------------------------------------------------------------------ SVD reverse rule
// Reverse pass of tSVD = full_svd_impl + cutoff_matrices (include/Xped/AD/ADTensor.hpp:253-325, 356-371).  The adjoints of
// the truncated factors are padded into the full factors' adjoints (left columns / top-left corner / top rows) and
//   dA = U (dS + (J+J^T) S + S (K+K^T)) V^+  +  (1 - U U^+) dU S^-1 V^+  +  U S^-1 dV^+ (1 - V V^+),   J = F o U^+dU, K = F o V^+dV
// is evaluated as   dA = [U (M - G1 S^-1 - S^-1 G2^T) + dU S^-1] V^+ + U (S^-1 dV^+)   with G1 = U^+ dU, G2 = V^+ dV:
// five grouped products and three elementwise kernels over all sectors at once.  The full factors are recomputed here
// (the forward pass keeps only the truncated ones), like the reference's checkpointed replay does.
Tensor* Context::tsvd_backward(const Tensor& A, const Tensor& U, const Tensor& S, const Tensor& Vd, const Tensor* dU, const Tensor* dS, const Tensor* dVd)
{
    (void)U;
    (void)S;
    (void)Vd;
    const Layout& LA = *A.layout;
    SvdResult F = tsvd_full(A, INT64_MAX / 4, 0.0, false);
    std::unique_ptr<Tensor> FU(F.U), FS(F.S), FV(F.Vd);
    const Layout &LU = *FU->layout, &LS = *FS->layout, &LV = *FV->layout;
    if(std::getenv("XB_DEBUG_SVDBWD")) {
        // compare the recomputed full factors with the caller's (leading columns / rows per sector)
        sync();
        std::vector<double> hf(FU->size()), hu(U.size()), hs(FS->size()), hss(S.size());
        cudaMemcpy(hf.data(), FU->d, sizeof(double) * hf.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(hu.data(), U.d, sizeof(double) * hu.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(hs.data(), FS->d, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(hss.data(), S.d, sizeof(double) * hss.size(), cudaMemcpyDeviceToHost);
        const Layout& Lu = *U.layout;
        for(size_t j = 0; j < Lu.sector.size(); ++j) {
            const int i = LU.find(Lu.sector[j]);
            double worst = 0.0, worst_s = 0.0;
            for(int64_t c = 0; c < Lu.cols[j]; ++c) {
                for(int64_t r = 0; r < Lu.rows[j]; ++r) worst = std::max(worst, std::fabs(hf[LU.off[i] + r + c * LU.rows[i]] - hu[Lu.off[j] + r + c * Lu.rows[j]]));
                worst_s = std::max(worst_s, std::fabs(hs[LS.off[i] + c + c * LS.rows[i]] - hss[S.layout->off[j] + c + c * S.layout->rows[j]]));
            }
            std::fprintf(stderr, "[svd bwd] sector %zu (%lld x %lld kept of %lld): max |FU - U| %.3e, max |FS - S| %.3e\n", j, (long long)Lu.rows[j], (long long)Lu.cols[j],
                         (long long)LU.cols[i], worst, worst_s);
        }
    }
    std::unique_ptr<Tensor> dFU(new Tensor(this, FU->layout, true)), dFS(new Tensor(this, FS->layout, true)), dFV(new Tensor(this, FV->layout, true));
    auto pad = [&](const Tensor* src, Tensor& dst, const char* what) {
        if(!src) return;
        const Layout &Ls = *src->layout, &Ld = *dst.layout;
        for(size_t j = 0; j < Ls.sector.size(); ++j) {
            const int i = Ld.find(Ls.sector[j]);
            if(i < 0) throw std::invalid_argument(std::string("tsvd_backward: sector of ") + what + " not present in the decomposition");
            if(Ls.rows[j] > Ld.rows[i] || Ls.cols[j] > Ld.cols[i]) throw std::invalid_argument(std::string("tsvd_backward: block of ") + what + " too large");
            if(Ls.rows[j] == 0 || Ls.cols[j] == 0) continue;
            check_cuda(cudaMemcpy2DAsync(dst.d + Ld.off[i], sizeof(double) * Ld.rows[i], src->d + Ls.off[j], sizeof(double) * Ls.rows[j],
                                         sizeof(double) * Ls.rows[j], static_cast<size_t>(Ls.cols[j]), cudaMemcpyDeviceToDevice, stream),
                       "tsvd_backward pad");
        }
    };
    pad(dU, *dFU, "dU");
    pad(dS, *dFS, "dS");
    pad(dVd, *dFV, "dVdag");
    std::unique_ptr<Tensor> G1(multiply(*FU, true, *dFU, false)), G2(multiply(*FV, false, *dFV, true));
    if(G1->layout->sector != LS.sector || G2->layout->sector != LS.sector || G1->layout->size != LS.size || G2->layout->size != LS.size)
        throw std::logic_error("tsvd_backward: layout of U^+ dU differs from Sigma's");
    const int nb = static_cast<int>(LS.sector.size());
    std::vector<BlockDesc> bs(nb);
    std::vector<DiagScaleDesc> du(nb), dv(nb);
    for(int i = 0; i < nb; ++i) {
        bs[i] = {LS.off[i], static_cast<int>(LS.rows[i]), static_cast<int>(LS.cols[i]), 1.0};
        du[i] = {LU.off[i], LS.off[i], static_cast<int>(LU.rows[i]), static_cast<int>(LU.cols[i]), static_cast<int>(LS.rows[i]), 1};
        dv[i] = {LV.off[i], LS.off[i], static_cast<int>(LV.rows[i]), static_cast<int>(LV.cols[i]), static_cast<int>(LS.rows[i]), 0};
    }
    const size_t b0 = sizeof(BlockDesc) * nb, b1 = sizeof(DiagScaleDesc) * nb;
    char* d_meta = nullptr;
    check_cuda(cudaMallocAsync(reinterpret_cast<void**>(&d_meta), b0 + 2 * b1, stream), "tsvd_backward meta");
    upload_async(d_meta, bs.data(), b0);
    upload_async(d_meta + b0, du.data(), b1);
    upload_async(d_meta + b0 + b1, dv.data(), b1);
    std::unique_ptr<Tensor> M(new Tensor(this, FS->layout, false));
    launch_svd_adjoint_core(G1->d, G2->d, dFS->d, FS->d, M->d, reinterpret_cast<BlockDesc*>(d_meta), nb, stream);
    if(std::getenv("XB_DEBUG_SVDBWD")) {
        sync();
        auto dump = [&](const char* name, const Tensor& T) {
            std::vector<double> h(T.size());
            cudaMemcpy(h.data(), T.d, sizeof(double) * h.size(), cudaMemcpyDeviceToHost);
            std::string fn = std::string("gpurun_out/dbg_") + name + ".bin";
            if(FILE* f = std::fopen(fn.c_str(), "wb")) {
                std::fwrite(h.data(), sizeof(double), h.size(), f);
                std::fclose(f);
            }
        };
        dump("Mcore", *M);
        dump("G1", *G1);
        dump("G2", *G2);
        dump("FS", *FS);
        dump("dFS", *dFS);
    }
    std::unique_ptr<Tensor> X1(multiply(*FU, false, *M, false));
    if(X1->layout->sector != LU.sector || X1->layout->size != LU.size) throw std::logic_error("tsvd_backward: layout of U*M differs from U's");
    launch_scale_by_diag(dFU->d, X1->d, FS->d, reinterpret_cast<DiagScaleDesc*>(d_meta + b0), nb, 1, stream);   // X1 += dU S^-1
    std::unique_ptr<Tensor> Z(new Tensor(this, FV->layout, false));
    launch_scale_by_diag(dFV->d, Z->d, FS->d, reinterpret_cast<DiagScaleDesc*>(d_meta + b0 + b1), nb, 0, stream);   // Z = S^-1 dV^+
    launches += 3;
    check_cuda(cudaFreeAsync(d_meta, stream), "tsvd_backward meta free");
    std::unique_ptr<Tensor> d1(multiply(*X1, false, *FV, false)), d2(multiply(*FU, false, *Z, false));
    std::unique_ptr<Tensor> sum(axpby(1.0, *d1, 1.0, d2.get()));
    if(sum->layout.get() == A.layout.get()) return sum.release();
    // A was built with an explicit sector order (push_back): move the blocks into A's storage order
    std::unique_ptr<Tensor> out(new Tensor(this, A.layout, true));
    const Layout& Lsum = *sum->layout;
    for(size_t i = 0; i < LA.sector.size(); ++i) {
        const int j = Lsum.find(LA.sector[i]);
        if(j < 0) continue;
        check_cuda(cudaMemcpyAsync(out->d + LA.off[i], sum->d + Lsum.off[j], sizeof(double) * LA.rows[i] * LA.cols[i], cudaMemcpyDeviceToDevice, stream),
                   "tsvd_backward relayout");
    }
    return out.release();
}
// ------------------------------------------------------------------------------------------------ MPS transfer contractions
// contract_L (src/MPS/MpsContractions.cpp:11-51):  Bnew[Q] = sum_{tree (Qin,sigma)->Q} sum_s Bra_sub(s)^T * Bold[Qin] * Ket_sub(s).
// Device form: two grouped launches instead of one small GEMM pair per (Q, tree, s):
//   X = "Bold applied to the left leg of Ket" (same layout as Ket; rows whose Qin is absent from Bold stay zero),
//   Bnew[Q] = Bra[Q]^T * X[Q]   (one full-K product per sector).
namespace {
void check_mps_args(const Tensor& Bold, const Tensor& Bra, const Tensor& Ket, const char* who)
{
    if(Bold.layout->rank() != 1 || Bold.layout->corank() != 1 || Bra.layout->rank() != 2 || Bra.layout->corank() != 1 || Ket.layout->rank() != 2 ||
       Ket.layout->corank() != 1)
        throw std::invalid_argument(std::string(who) + ": expects Bold(1,1), Bra(2,1), Ket(2,1)");
}
GemmItem blank_item()
{
    GemmItem g;
    std::memset(&g, 0, sizeof(g));
    g.alpha = 1.0;
    return g;
}
}   // namespace

void Context::contract_L(const Tensor& Bold, const Tensor& Bra, const Tensor& Ket, Tensor** Bnew)
{
    check_mps_args(Bold, Bra, Ket, "contract_L");
    const Layout &LB = *Bold.layout, &LBra = *Bra.layout, &LK = *Ket.layout;
    LayoutPtr LO = layout({LBra.cod[0]}, {LK.cod[0]});
    std::unique_ptr<Tensor> out(new Tensor(this, LO, true));
    std::unique_ptr<Tensor> X(new Tensor(this, Ket.layout, true));
    std::vector<GemmItem> s1, s2;
    for(size_t i = 0; i < LBra.sector.size(); ++i) {
        const Q& q = LBra.sector[i];
        const int k = LK.find(q);
        if(k < 0) continue;
        const int o = LO->find(q);
        if(o < 0) continue;
        const int dimQ = static_cast<int>(LBra.cols[i]);
        if(LK.cols[k] != dimQ && false) throw std::invalid_argument("contract_L: Bra / Ket blocks differ");
        bool any = false;
        for(const Tree& t : LBra.cdom->trees(q)) {
            const int ib = LB.find(t.unc[0]);
            if(ib < 0) continue;
            const int64_t ro_k = LK.cdom->left_offset(t);
            const int d0 = static_cast<int>(t.dims[0]), d1 = static_cast<int>(t.dims[1]);
            if(LB.rows[ib] != d0 || LB.cols[ib] != d0) throw std::invalid_argument("contract_L: Bold block does not match the left leg of Bra/Ket");
            for(int s = 0; s < d1; ++s) {
                GemmItem g = blank_item();   // X_sub(s) = Bold[Qin] * Ket_sub(s)
                g.a_off = LB.off[ib];
                g.lda = d0;
                g.b_off = LK.off[k] + ro_k + static_cast<int64_t>(s) * d0;
                g.ldb = static_cast<int>(LK.rows[k]);
                g.c_off = g.b_off;
                g.ldc = g.ldb;
                g.m = d0;
                g.k = d0;
                g.n = static_cast<int>(LK.cols[k]);
                s1.push_back(g);
                any = true;
            }
        }
        if(!any) continue;
        GemmItem g = blank_item();   // Bnew[Q] = Bra[Q]^T * X[Q]
        g.a_off = LBra.off[i];
        g.lda = static_cast<int>(LBra.rows[i]);
        g.ta = 1;
        g.b_off = LK.off[k];
        g.ldb = static_cast<int>(LK.rows[k]);
        g.c_off = LO->off[o];
        g.ldc = static_cast<int>(LO->rows[o]);
        g.m = dimQ;
        g.k = static_cast<int>(LBra.rows[i]);
        g.n = static_cast<int>(LK.cols[k]);
        if(LK.rows[k] != LBra.rows[i] || LO->rows[o] != g.m || LO->cols[o] != g.n) throw std::invalid_argument("contract_L: Bra / Ket block shapes differ");
        s2.push_back(g);
    }
    run_gemm(Bold.d, Ket.d, X->d, s1);
    run_gemm(Bra.d, X->d, out->d, s2);
    *Bnew = out.release();
}

// contract_R (MpsContractions.cpp:54-105): Bnew[Qin] += coeff_rightOrtho(Q,Qin) * sum_s Ket_sub(s) * Bold[Q] * Bra_sub(s)^T.
//   Y[Q] = Ket[Q] * Bold[Q] (one product per sector), then the (Q, tree, s) terms are accumulated into Bnew[Qin] in
//   "waves" holding at most one term per destination block, so the summation order is fixed (deterministic).
void Context::contract_R(const Tensor& Bold, const Tensor& Bra, const Tensor& Ket, Tensor** Bnew)
{
    check_mps_args(Bold, Bra, Ket, "contract_R");
    const Layout &LB = *Bold.layout, &LBra = *Bra.layout, &LK = *Ket.layout;
    LayoutPtr LO = layout({LK.dom[0]}, {LBra.dom[0]});
    std::unique_ptr<Tensor> out(new Tensor(this, LO, true));
    std::unique_ptr<Tensor> Y(new Tensor(this, Ket.layout, true));
    std::vector<GemmItem> s1;
    std::vector<std::vector<GemmItem>> waves;
    std::vector<int> used(LO->sector.size(), 0);
    for(size_t i = 0; i < LK.sector.size(); ++i) {
        const Q& q = LK.sector[i];
        const int ib = LB.find(q), ibra = LBra.find(q);
        if(ib < 0 || ibra < 0) continue;
        const int dimQ = static_cast<int>(LK.cols[i]);
        if(LB.rows[ib] != dimQ || LB.cols[ib] != dimQ || LBra.cols[ibra] != dimQ) throw std::invalid_argument("contract_R: Bold block does not match the right leg of Bra/Ket");
        GemmItem g = blank_item();
        g.a_off = LK.off[i];
        g.lda = static_cast<int>(LK.rows[i]);
        g.b_off = LB.off[ib];
        g.ldb = dimQ;
        g.c_off = LK.off[i];
        g.ldc = g.lda;
        g.m = g.lda;
        g.k = dimQ;
        g.n = dimQ;
        s1.push_back(g);
        for(const Tree& t : LK.cdom->trees(q)) {
            const int o = LO->find(t.unc[0]);
            if(o < 0) continue;
            const int64_t ro_k = LK.cdom->left_offset(t), ro_b = LBra.cdom->left_offset(t);
            const int d0 = static_cast<int>(t.dims[0]), d1 = static_cast<int>(t.dims[1]);
            const double coeff = sym->coeff_rightOrtho(q, t.unc[0]);
            for(int s = 0; s < d1; ++s) {
                GemmItem h = blank_item();   // Bnew[Qin] += coeff * Y_sub(s) * Bra_sub(s)^T
                h.a_off = LK.off[i] + ro_k + static_cast<int64_t>(s) * d0;
                h.lda = static_cast<int>(LK.rows[i]);
                h.b_off = LBra.off[ibra] + ro_b + static_cast<int64_t>(s) * d0;
                h.ldb = static_cast<int>(LBra.rows[ibra]);
                h.tb = 1;
                h.c_off = LO->off[o];
                h.ldc = static_cast<int>(LO->rows[o]);
                h.m = d0;
                h.k = dimQ;
                h.n = d0;
                h.acc = 1;
                h.alpha = coeff;
                if(LO->rows[o] != d0 || LO->cols[o] != d0) throw std::invalid_argument("contract_R: result block shape mismatch");
                const int w = used[o]++;
                if(static_cast<int>(waves.size()) <= w) waves.resize(w + 1);
                waves[w].push_back(h);
            }
        }
    }
    run_gemm(Ket.d, Bold.d, Y->d, s1);
    for(const auto& w : waves) run_gemm(Y->d, Bra.d, out->d, w);
    *Bnew = out.release();
}

}   // namespace xb

namespace xb {

// ------------------------------------------------------------------------------------------------ MPS sweep steps
// Mps::leftSweepStep (src/MPS/Mps.cpp:181-202):  [U, S, Vdag] = Ac[loc].permute<+1,0,2,1>().tSVD(max_Nsv, eps_svd, truncWeight, entropy, ., false);
//   Ac[loc] = Vdag.permute<-1,0,2,1>();  Ac[loc-1] = Ac[loc-1] * (U * S).
void Context::mps_left_sweep_step(const Tensor& Aloc, const Tensor* Aprev, int64_t max_nsv, double eps_svd, Tensor** Aloc_new, Tensor** Aprev_new,
                                  double* trunc_weight, double* entropy)
{
    if(Aloc.layout->rank() != 2 || Aloc.layout->corank() != 1) throw std::invalid_argument("leftSweepStep: site tensor must be a Tensor<2,1>");
    std::unique_ptr<Tensor> p(permute(Aloc, +1, {0, 2, 1}));
    SvdResult r = tsvd(*p, max_nsv, eps_svd, false);
    std::unique_ptr<Tensor> U(r.U), S(r.S), Vd(r.Vd);
    if(trunc_weight) *trunc_weight = r.trunc_weight;
    if(entropy) *entropy = r.entropy;
    *Aloc_new = permute(*Vd, -1, {0, 2, 1});
    *Aprev_new = nullptr;
    if(Aprev) {
        std::unique_ptr<Tensor> left(multiply(*U, false, *S, false));
        *Aprev_new = multiply(*Aprev, false, *left, false);
    }
}

// Mps::rightSweepStep (src/MPS/Mps.cpp:204-229):  [U, S, Vdag] = Ac[loc].tSVD(...);  Ac[loc] = U;
//   Ac[loc+1] = ((S * Vdag) * Ac[loc+1].permute<+1,0,1,2>()).permute<-1,0,1,2>().
void Context::mps_right_sweep_step(const Tensor& Aloc, const Tensor* Anext, int64_t max_nsv, double eps_svd, Tensor** Aloc_new, Tensor** Anext_new,
                                   double* trunc_weight, double* entropy)
{
    if(Aloc.layout->rank() != 2 || Aloc.layout->corank() != 1) throw std::invalid_argument("rightSweepStep: site tensor must be a Tensor<2,1>");
    SvdResult r = tsvd(Aloc, max_nsv, eps_svd, false);
    std::unique_ptr<Tensor> U(r.U), S(r.S), Vd(r.Vd);
    if(trunc_weight) *trunc_weight = r.trunc_weight;
    if(entropy) *entropy = r.entropy;
    *Anext_new = nullptr;
    if(Anext) {
        std::unique_ptr<Tensor> right(multiply(*S, false, *Vd, false));
        std::unique_ptr<Tensor> np(permute(*Anext, +1, {0, 1, 2}));
        std::unique_ptr<Tensor> prod(multiply(*right, false, *np, false));
        *Anext_new = permute(*prod, -1, {0, 1, 2});
    }
    *Aloc_new = U.release();
}

}   // namespace xb
