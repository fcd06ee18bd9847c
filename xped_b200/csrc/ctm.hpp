// Directional CTMRG driver re-hosted on the device backend (1x1 unit cell, TRank = 2, corner projectors).
// Mirrors Xped::CTM<double,Symmetry,2> (include/Xped/PEPS/CTM.hpp:60-312, src/PEPS/CTM.cpp).
#pragma once
#include <memory>
#include <vector>

#include "engine.hpp"

namespace xb {

using TP = std::shared_ptr<Tensor>;

// decompose(T1, T2, chi): src/PEPS/PEPSContractions.cpp:18-43
void decompose_projectors(Context& ctx, const Tensor& T1, const Tensor& T2, int64_t chi, Tensor** P1, Tensor** P2);
// avg(rho, O, bond) for a two-site observable (src/PEPS/LinearAlgebra.cpp:57-60): rho.twist(0).twist(1).contract<{1,2,3,4},{3,4,1,2},0>(O).trace()
double avg_two_site(Context& ctx, const Tensor& rho, const Tensor& op);
// rho1 from a two-site density matrix (src/PEPS/CTM.cpp:675-678) and avg(rho1, O) of a one-site observable (LinearAlgebra.cpp:22-24)
Tensor* rdm_one_site(Context& ctx, const Tensor& rho);
double avg_one_site(Context& ctx, const Tensor& rho1, const Tensor& op);
// energy of a 1x1 iPEPS after `steps` CTM steps and its gradient w.r.t. the site tensor by reverse-mode AD (dA may be null: value only)
double ctm_energy_grad(Context& ctx, const Tensor& A, int64_t chi, int steps, const Tensor& op_h, const Tensor& op_v, bool checkpoint, Tensor** dA);

class Ctm
{
public:
    Ctm(Context& ctx, const Tensor& A, int64_t chi);   // CTM(A, chi, FROM_A) + init()  (CTM.cpp:45-244)
    Ctm(Context& ctx, int64_t chi, const std::vector<TP>& tensors);   // view on given C1..C4, T1..T4, A, Adag (checkpoint replay)
    bool checkpoint_moves = false;     // under an AD tape: record one checkpoint per move instead of every op
    std::vector<TP> all_tensors() const;
    int64_t chi() const { return chi_; }

    void grow_all();                    // CTM.cpp:385-420
    void move(int direction);           // 0 left, 1 right, 2 top, 3 bottom  (CTM.cpp:924-1183)
    void move_impl(int direction);
    Tensor* contract_corner(int corner);   // 0 UL, 1 UR, 2 LR, 3 LL  (CTM.cpp:1643-1760)
    Tensor* rdm(int which);             // 0: rho_h, 1: rho_v, normalised  (computeRDM_h / computeRDM_v, CTM.cpp:528-874, TRank == 2)
    const Tensor& get(int which) const;
    void set(int which, Tensor* t);     // takes ownership

private:
    Context& ctx_;
    int64_t chi_;
    TP A_, Adag_;
    TP C1_, C2_, C3_, C4_, T1_, T2_, T3_, T4_;

    TP con(const TP& a, std::initializer_list<int> la, const TP& b, std::initializer_list<int> lb, int res_rank) const;
    TP tw(const TP& t, std::initializer_list<int> legs) const;   // chained twists; no-op (shared) for bosonic symmetries
    TP normalized(const TP& t) const;                            // t * (1 / t.maxNorm()) without a host round trip
    TP corner(int c) const;
    void projectors(int direction, TP& P1, TP& P2) const;        // get_projectors  CTM.cpp:1185-1337
    void init_from_A();
};

}   // namespace xb
