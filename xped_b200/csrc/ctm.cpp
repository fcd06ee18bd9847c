// CTMRG moves as sequences of cached-plan device ops. Tensor network definitions (NCON label arrays) follow
// src/PEPS/CTM.cpp of the reference; what differs is that every permute plan, layout and fusion-tree table is
// built once and reused, no environment deep copies are taken per move (CTM.cpp:929 `auto curr = *this`), and
// the max-norm normalisation never leaves the device.
#include "ctm.hpp"

#include "ad.hpp"

#include <stdexcept>

namespace xb {

void decompose_projectors(Context& ctx, const Tensor& T1, const Tensor& T2, int64_t chi, Tensor** P1, Tensor** P2)
{
    std::unique_ptr<Tensor> prod(ctx.multiply(T1, false, T2, false));
    SvdResult r = ctx.tsvd(*prod, chi, 1.e-10, false);   // PEPSContractions.cpp:27
    std::unique_ptr<Tensor> U(r.U), S(r.S), Vd(r.Vd);
    std::unique_ptr<Tensor> isqrtS(ctx.diag_map(*S, 2));   // S.diag_sqrt().diag_inv()
    std::unique_ptr<Tensor> l(ctx.multiply(*isqrtS, false, *U, true));
    *P1 = ctx.multiply(*l, false, T1, false);   // isqrtS * U^dagger * T1
    std::unique_ptr<Tensor> rgt(ctx.multiply(T2, false, *Vd, true));
    *P2 = ctx.multiply(*rgt, false, *isqrtS, false);   // T2 * Vdag^dagger * isqrtS
}

Ctm::Ctm(Context& ctx, const Tensor& A, int64_t chi) : ctx_(ctx), chi_(chi)
{
    if(A.layout->rank() != 2 || A.layout->corank() != 3) throw std::invalid_argument("CTM: A must be a Tensor<2,3> (left, top | right, bottom, phys)");
    A_.reset(ctx.copy(A));
    // iPEPS::updateAdags (src/PEPS/iPEPS.cpp:200-215): A.adjoint().eval().permute<0,3,4,2,0,1>()
    std::unique_ptr<Tensor> adj(ctx.adjoint(*A_));
    Adag_.reset(ctx.permute(*adj, 0, {3, 4, 2, 0, 1}));
    init_from_A();
}

TP Ctm::con(const TP& a, std::initializer_list<int> la, const TP& b, std::initializer_list<int> lb, int res_rank) const
{
    return TP(ctx_.contract(*a, std::vector<int>(la), *b, std::vector<int>(lb), res_rank));
}

TP Ctm::tw(const TP& t, std::initializer_list<int> legs) const
{
    if(!ctx_.sym->fermionic()) return t;   // Tensor::twist early-out, src/Core/Tensor.cpp:362
    TP cur = t;
    for(int leg : legs) cur.reset(ctx_.twist(*cur, leg));
    return cur;
}

TP Ctm::normalized(const TP& t) const { return TP(ctx_.scale_by_inv_maxnorm(*t)); }

void Ctm::init_from_A()
{
    // CTM.cpp:158-214 (CTM_INIT::FROM_A, TRank == 2)
    const Layout& LA = *A_->layout;
    const Layout& LD = *Adag_->layout;
    BasisPtr ketL = LA.dom[0], ketT = LA.dom[1], ketR = LA.cod[0], ketB = LA.cod[1];
    BasisPtr braL = LD.dom[0], braT = LD.dom[1], braR = LD.cod[0], braB = LD.cod[1];
    auto identity = [&](const std::vector<BasisPtr>& d, const std::vector<BasisPtr>& c) {
        TP t(new Tensor(&ctx_, ctx_.layout(d, c), true));
        ctx_.fill(*t, 1, 0.0, 0);
        return t;
    };
    auto fused = [&](const BasisPtr& k, const BasisPtr& b) -> BasisPtr { return k->combine(*b)->forget_history(); };
    TP fuse_ll = identity({fused(ketL, braL)}, {ketL, braL});
    TP fuse_uu = identity({fused(ketT, braT)}, {ketT, braT});
    TP fuse_rr = identity({ketR, braR}, {fused(ketR, braR)});
    TP fuse_dd = identity({ketB, braB}, {fused(ketB, braB)});
    TP fll = tw(fuse_ll, {1, 2}), fuu = tw(fuse_uu, {1, 2});

    C1_ = con(con(con(A_, {1, 2, -1, -2, 3}, tw(Adag_, {3, 4}), {1, 2, 3, -3, -4}, 2), {1, -1, 2, -2}, fuse_rr, {1, 2, -3}, 2), {1, 2, -2}, fuse_dd,
              {1, 2, -1}, 0);
    C2_ = con(con(con(A_, {-1, 1, 2, -2, 3}, tw(Adag_, {4}), {-3, 1, 3, 2, -4}, 2), {1, -1, 2, -2}, fll, {-3, 1, 2}, 2), {1, 2, -1}, fuse_dd, {1, 2, -2}, 1);
    C3_ = con(con(con(A_, {-1, -2, 1, 2, 3}, Adag_, {-3, -4, 3, 1, 2}, 2), {1, -1, 2, -2}, fll, {-3, 1, 2}, 2), {1, 2, -2}, fuu, {-1, 1, 2}, 2);
    C4_ = con(con(con(A_, {1, -1, -2, 2, 3}, tw(Adag_, {3}), {1, -3, 3, -4, 2}, 2), {1, -1, 2, -2}, fuu, {-3, 1, 2}, 2), {1, 2, -1}, fuse_rr, {1, 2, -2}, 1);

    T1_ = con(con(con(A_, {-1, 1, -2, -3, 2}, tw(Adag_, {3, 4}), {-4, 1, 2, -5, -6}, 3), {1, -1, -2, 2, -3, -4}, fll, {-5, 1, 2}, 4), {1, -3, 2, -4, -1},
              fuse_rr, {1, 2, -2}, 1);
    T2_ = con(con(con(A_, {-1, -2, 1, -3, 2}, tw(Adag_, {4}), {-4, -5, 2, 1, -6}, 3), {-1, 1, -2, -3, 2, -4}, fuu, {-5, 1, 2}, 4), {-1, 1, -2, 2, -3},
              fuse_dd, {1, 2, -4}, 3);
    T3_ = con(con(con(A_, {-1, -2, -3, 1, 2}, tw(Adag_, {3}), {-4, -5, 2, -6, 1}, 3), {1, -1, -2, 2, -3, -4}, fll, {-5, 1, 2}, 4), {-1, 1, -2, 2, -3},
              fuse_rr, {1, 2, -4}, 3);
    T4_ = con(con(con(A_, {1, -1, -2, -3, 2}, tw(Adag_, {3, 4}), {1, -4, 2, -5, -6}, 3), {1, -1, -2, 2, -3, -4}, fuu, {-5, 1, 2}, 4), {-3, 1, -4, 2, -1},
              fuse_dd, {1, 2, -2}, 1);
}

TP Ctm::corner(int c) const
{
    switch(c) {
    case 0: {   // UPPER_LEFT  CTM.cpp:1660-1672
        TP C1T1 = con(C1_, {-1, 1}, T1_, {1, -2, -3, -4}, 1);
        TP T4C1T1 = con(T4_, {1, -1, -2, -3}, tw(C1T1, {0}), {1, -4, -5, -6}, 3);
        TP T4C1T1A = con(T4C1T1, {-1, 1, -2, -3, 2, -4}, A_, {1, 2, -5, -6, -7}, 4);
        return con(T4C1T1A, {-1, 1, -4, 2, -5, -2, 3}, tw(Adag_, {3, 4}), {1, 2, 3, -6, -3}, 3);
    }
    case 3: {   // LOWER_LEFT  CTM.cpp:1688-1694
        TP C4T3 = con(C4_, {-1, 1}, T3_, {-2, -3, 1, -4}, 1);
        TP T4C4T3 = con(T4_, {-1, 1, -2, -3}, C4T3, {1, -4, -5, -6}, 3);
        TP T4C4T3A = con(T4C4T3, {-1, 1, -2, 2, -3, -4}, tw(A_, {3}), {1, -5, -6, 2, -7}, 4);
        return con(T4C4T3A, {-4, 1, 2, -1, -5, -2, 3}, tw(Adag_, {3}), {1, -6, 3, -3, 2}, 3);
    }
    case 1: {   // UPPER_RIGHT  CTM.cpp:1709-1715
        TP T1C2 = con(T1_, {-1, 1, -2, -3}, C2_, {1, -4}, 3);
        TP T1C2T2 = con(T1C2, {-1, -2, -3, 1}, T2_, {-4, -5, 1, -6}, 3);
        TP T1C2T2A = con(T1C2T2, {-1, 1, -2, 2, -3, -4}, tw(A_, {2}), {-5, 1, 2, -6, -7}, 4);
        return con(T1C2T2A, {-1, 1, 2, -4, -2, -5, 3}, tw(Adag_, {4}), {-3, 1, 3, 2, -6}, 3);
    }
    case 2: {   // LOWER_RIGHT  CTM.cpp:1730-1737
        TP C3T3 = con(tw(C3_, {1}), {-1, 1}, T3_, {-2, -3, -4, 1}, 1);
        TP T2C3T3 = con(T2_, {-1, -2, -3, 1}, C3T3, {1, -4, -5, -6}, 3);
        TP T2C3T3A = con(T2C3T3, {1, -1, -2, 2, -3, -4}, tw(A_, {2, 3}), {-5, -6, 1, 2, -7}, 4);
        return con(T2C3T3A, {1, -1, 2, -4, -5, -2, 3}, Adag_, {-6, -3, 3, 1, 2}, 3);
    }
    default: throw std::invalid_argument("CTM: corner must be 0..3 (UL, UR, LR, LL)");
    }
}

Tensor* Ctm::contract_corner(int c) { return ctx_.copy(*corner(c)); }

// Full contraction X[1,2,3,4] Y[3,4,1,2] of two Tensor<2,2> = trace(X * Y): contract<{1,2,3,4},{3,4,1,2},0>(...).trace() of the
// reference goes through a Tensor<0,4> * Tensor<4,0> product; the recoupling factors of the leg bending cancel in a full contraction.
static double full_contraction(Context& ctx, const Tensor& X, const Tensor& Y)
{
    std::unique_ptr<Tensor> prod(ctx.multiply(X, false, Y, false));
    return ctx.trace(*prod);
}

double avg_two_site(Context& ctx, const Tensor& rho, const Tensor& op)
{
    if(rho.layout->rank() != 2 || rho.layout->corank() != 2 || op.layout->rank() != 2 || op.layout->corank() != 2)
        throw std::invalid_argument("avg: rho and the observable must be Tensor<2,2>");
    std::unique_ptr<Tensor> a(ctx.twist(rho, 0)), b(ctx.twist(*a, 1));
    return full_contraction(ctx, *b, op);
}

Tensor* rdm_one_site(Context& ctx, const Tensor& rho)
{
    if(rho.layout->rank() != 2 || rho.layout->corank() != 2) throw std::invalid_argument("rdm_one_site: rho must be a Tensor<2,2>");
    const Layout& L = *rho.layout;
    std::unique_ptr<Tensor> Id(new Tensor(&ctx, ctx.layout({L.cod[1]}, {L.dom[1]}), true));
    ctx.fill(*Id, 1, 0.0, 0);
    std::unique_ptr<Tensor> Idt(ctx.twist(*Id, 0));
    std::unique_ptr<Tensor> rho1(ctx.contract(rho, {-1, 1, -2, 2}, *Idt, {2, 1}, 1));
    std::unique_ptr<Tensor> tw(ctx.twist(*rho1, 0));
    const double norm = ctx.trace(*tw);
    return ctx.axpby(1.0 / norm, *rho1, 0.0, nullptr);
}

double avg_one_site(Context& ctx, const Tensor& rho1, const Tensor& op)
{
    if(rho1.layout->rank() != 1 || rho1.layout->corank() != 1 || op.layout->rank() != 1 || op.layout->corank() != 1)
        throw std::invalid_argument("avg: rho1 and the observable must be Tensor<1,1>");
    std::unique_ptr<Tensor> a(ctx.twist(rho1, 0));
    std::unique_ptr<Tensor> prod(ctx.multiply(*a, false, op, false));   // contract<{1,2},{2,1},0>(O).trace() = trace(rho1 * O)
    return ctx.trace(*prod);
}

Tensor* Ctm::rdm(int which)
{
    TP rho;
    if(which == 0) {   // computeRDM_h, CTM.cpp:585-603
        TP C1T1 = con(C1_, {-1, 1}, T1_, {1, -2, -3, -4}, 1);
        TP T4C1T1 = con(T4_, {1, -1, -2, -3}, tw(C1T1, {0}), {1, -4, -5, -6}, 3);
        TP T4C1T1A = con(T4C1T1, {-1, 1, -2, -3, 2, -4}, A_, {1, 2, -5, -6, -7}, 4);
        TP C4T3 = con(C4_, {-1, 1}, T3_, {-2, -3, 1, -4}, 1);
        TP C4T3Ad = con(C4T3, {-1, -2, 1, -3}, tw(Adag_, {3}), {-4, -5, -6, -7, 1}, 3);
        TP left_half = con(T4C1T1A, {1, 2, -1, 3, -2, 4, -3}, C4T3Ad, {1, 4, -4, 2, 3, -5, -6}, 3);
        TP T1C2 = con(T1_, {-1, 1, -2, -3}, C2_, {1, -4}, 3);
        TP T1C2T2 = con(T1C2, {-1, -2, -3, 1}, T2_, {-4, -5, 1, -6}, 3);
        TP AT1C2T2 = con(tw(A_, {1}), {-1, 1, 2, -2, -3}, T1C2T2, {-4, 1, -5, 2, -6, -7}, 3);
        TP T3C3 = con(T3_, {-1, -2, -3, 1}, C3_, {-4, 1}, 3);
        TP T3C3Ad = con(T3C3, {-1, 1, -2, -3}, Adag_, {-4, -5, -6, -7, 1}, 3);
        TP right_half = con(AT1C2T2, {-1, 1, -2, -3, 2, 3, 4}, T3C3Ad, {1, -4, 4, -5, 2, -6, 3}, 3);
        rho = con(left_half, {1, 2, -3, 3, -1, 4}, right_half, {2, -4, 1, 3, 4, -2}, 2);
    } else if(which == 1) {   // computeRDM_v, CTM.cpp:753-779
        TP C1T1 = con(C1_, {-1, 1}, T1_, {1, -2, -3, -4}, 1);
        TP T4C1T1 = con(T4_, {1, -1, -2, -3}, tw(C1T1, {0}), {1, -4, -5, -6}, 3);
        TP T4C1T1A = con(T4C1T1, {-1, 1, -2, -3, 2, -4}, A_, {1, 2, -5, -6, -7}, 4);
        TP C2T2 = con(C2_, {-1, 1}, T2_, {-2, -3, 1, -4}, 1);
        TP AdC2T2 = con(tw(Adag_, {3, 4}), {-1, -2, -3, 1, -4}, C2T2, {-5, -6, 1, -7}, 4);
        TP upper_half = con(T4C1T1A, {-1, 1, 2, 3, 4, -2, -3}, AdC2T2, {1, 3, -4, -5, 2, 4, -6}, 3);
        TP C4T3 = con(C4_, {-1, 1}, T3_, {-2, -3, 1, -4}, 1);
        TP T4C4T3 = con(T4_, {-1, 1, -2, -3}, C4T3, {1, -4, -5, -6}, 3);
        TP AT4C4T3 = con(tw(A_, {0}), {1, -1, -2, 2, -3}, T4C4T3, {-4, 1, -5, 2, -6, -7}, 3);
        TP T2C3 = con(T2_, {-1, -2, -3, 1}, C3_, {1, -4}, 3);
        TP T2C3Ad = con(T2C3, {-1, 1, -2, -3}, Adag_, {-4, -5, -6, 1, -7}, 3);
        TP lower_half = con(AT4C4T3, {-1, 1, -2, -3, 2, 3, 4}, T2C3Ad, {1, -4, 4, 2, -5, -6, 3}, 3);
        rho = con(upper_half, {1, 2, -3, -1, 3, 4}, lower_half, {2, -4, 1, 4, 3, -2}, 2);
    } else {
        throw std::invalid_argument("CTM: rdm selector must be 0 (horizontal) or 1 (vertical)");
    }
    // norm = rho.contract<{1,2,3,4},{3,4,1,2},0>(Id2.twist(0).twist(1)).trace();  rho *= 1 / norm   (CTM.cpp:660-667)
    const Layout& L = *rho->layout;
    TP Id2(new Tensor(&ctx_, ctx_.layout(L.cod, L.dom), true));
    ctx_.fill(*Id2, 1, 0.0, 0);
    return ctx_.scale_by_inv_contraction(*rho, *tw(Id2, {0, 1}));
}

void Ctm::projectors(int direction, TP& P1, TP& P2) const
{
    // get_projectors, PROJECTION::CORNER (CTM.cpp:1196-1318); decompose argument order per direction.
    TP Qa, Qb;
    switch(direction) {
    case 0: Qa = tw(corner(3), {3, 4, 5}); Qb = corner(0); break;   // LEFT:   decompose(Q4, Q1)
    case 1: Qa = corner(1); Qb = corner(2); break;                  // RIGHT:  decompose(Q2, Q3)
    case 2: Qa = corner(0); Qb = corner(1); break;                  // TOP:    decompose(Q1, Q2)
    case 3: Qa = corner(2); Qb = tw(corner(3), {0, 1, 2}); break;   // BOTTOM: decompose(Q3, Q4)
    default: throw std::invalid_argument("CTM: direction must be 0..3");
    }
    Tensor *p1 = nullptr, *p2 = nullptr;
    const int tag_saved = ctx_.svd_call_tag;
    ctx_.svd_call_tag = 1 + direction;   // one warm-start subspace per move direction
    try {
        decompose_projectors(ctx_, *Qa, *Qb, chi_, &p1, &p2);
    } catch(...) {
        ctx_.svd_call_tag = tag_saved;
        throw;
    }
    ctx_.svd_call_tag = tag_saved;
    P1.reset(p1);
    P2.reset(p2);
}

Ctm::Ctm(Context& ctx, int64_t chi, const std::vector<TP>& t) : ctx_(ctx), chi_(chi)
{
    if(t.size() != 10) throw std::invalid_argument("CTM: 10 tensors expected (C1..C4, T1..T4, A, Adag)");
    C1_ = t[0]; C2_ = t[1]; C3_ = t[2]; C4_ = t[3];
    T1_ = t[4]; T2_ = t[5]; T3_ = t[6]; T4_ = t[7];
    A_ = t[8]; Adag_ = t[9];
}

std::vector<TP> Ctm::all_tensors() const { return {C1_, C2_, C3_, C4_, T1_, T2_, T3_, T4_, A_, Adag_}; }

// AD with checkpointing (src/PEPS/CTM.cpp:976-987 and the like: the reference wraps contractCorner / get_projectors / renormalize_* of a move in
// stan checkpoints that re-run them during the reverse sweep): here ONE checkpoint per move -- the forward pass keeps only the environment at the
// move boundaries, the reverse sweep replays the move under a sub-tape.
void Ctm::move(int direction)
{
    if(ctx_.taping() && checkpoint_moves) {
        Context* ctx = &ctx_;
        const int64_t chi = chi_;
        std::vector<TP> outs = ctx_.tape->checkpoint(all_tensors(), [ctx, chi, direction](const std::vector<TP>& in) {
            Ctm tmp(*ctx, chi, in);
            tmp.move_impl(direction);
            std::vector<TP> all = tmp.all_tensors();
            all.resize(8);   // the environment; A and Adag are inputs only
            return all;
        });
        C1_ = outs[0]; C2_ = outs[1]; C3_ = outs[2]; C4_ = outs[3];
        T1_ = outs[4]; T2_ = outs[5]; T3_ = outs[6]; T4_ = outs[7];
        return;
    }
    move_impl(direction);
}

void Ctm::move_impl(int direction)
{
    TP P1, P2;
    projectors(direction, P1, P2);
    switch(direction) {
    case 0: {   // left_move + renormalize_left (CTM.cpp:924-988, 1339-1419)
        TP C1T1 = con(C1_, {-1, 1}, T1_, {1, -4, -2, -3}, 3);
        TP C1n = con(tw(P1, {1, 2, 3}), {-1, 1, 2, 3}, C1T1, {1, 2, 3, -2}, 0);
        TP C4T3 = con(C4_, {-2, 1}, T3_, {-3, -4, 1, -1}, 1);
        TP C4n = con(C4T3, {-2, 1, 2, 3}, tw(P2, {0, 1, 2}), {1, 2, 3, -1}, 1);
        TP P1T4 = con(tw(P1, {1, 2, 3}), {-1, 1, -2, -3}, T4_, {-4, 1, -5, -6}, 3);
        TP P1T4A = con(P1T4, {-1, 2, -2, -3, 1, -4}, A_, {1, -5, -6, 2, -7}, 4);
        TP P1T4AAdag = con(P1T4A, {-1, 2, -4, 1, -5, -2, 3}, tw(Adag_, {3, 4}), {1, -6, 3, -3, 2}, 3);
        TP T4n = con(P1T4AAdag, {-2, -3, -4, 1, 2, 3}, tw(P2, {0, 1, 2}), {1, 2, 3, -1}, 1);
        C1_ = normalized(C1n);
        T4_ = normalized(T4n);
        C4_ = normalized(C4n);
        break;
    }
    case 1: {   // right_move + renormalize_right (CTM.cpp:990-1053, 1421-1490)
        TP T1C2 = con(T1_, {-1, 1, -3, -4}, C2_, {1, -2}, 1);
        TP C2n = con(T1C2, {-1, 1, 2, 3}, P2, {1, 2, 3, -2}, 1);
        TP C3T3 = con(tw(C3_, {1}), {-1, 1}, T3_, {-2, -3, -4, 1}, 3);
        TP C3n = con(P1, {-1, 1, 2, 3}, C3T3, {1, 2, 3, -2}, 2);
        TP P1T2 = con(P1, {-1, 1, -2, -3}, T2_, {-4, -5, 1, -6}, 3);
        TP P1T2A = con(P1T2, {-1, 1, -2, 2, -3, -4}, tw(A_, {2}), {-5, 1, 2, -6, -7}, 4);
        TP P1T2AAdag = con(P1T2A, {-1, 1, 2, -4, -2, -5, 3}, tw(Adag_, {4}), {-3, 1, 3, 2, -6}, 3);
        TP T2n = con(P1T2AAdag, {-3, -1, -2, 1, 2, 3}, P2, {1, 2, 3, -4}, 3);
        C2_ = normalized(C2n);
        T2_ = normalized(T2n);
        C3_ = normalized(C3n);
        break;
    }
    case 2: {   // top_move + renormalize_top (CTM.cpp:1055-1118, 1492-1558)
        TP T4C1 = con(T4_, {1, -1, -3, -4}, tw(C1_, {0}), {1, -2}, 1);
        TP C1n = con(T4C1, {-1, 1, 2, 3}, P2, {1, 2, 3, -2}, 0);
        TP C2T2 = con(C2_, {-1, 1}, T2_, {-2, -3, 1, -4}, 3);
        TP C2n = con(P1, {-1, 1, 2, 3}, C2T2, {1, 2, 3, -2}, 1);
        TP P1T1 = con(P1, {-1, 1, -2, -3}, T1_, {1, -4, -5, -6}, 3);
        TP P1T1A = con(P1T1, {-1, 1, -2, -3, 2, -4}, A_, {1, 2, -5, -6, -7}, 4);
        TP P1T1AAdag = con(P1T1A, {-1, 1, -4, 2, -5, -2, 3}, tw(Adag_, {3, 4}), {1, 2, 3, -6, -3}, 3);
        TP T1n = con(P1T1AAdag, {-1, -3, -4, 1, 2, 3}, P2, {1, 2, 3, -2}, 1);
        C1_ = normalized(C1n);
        T1_ = normalized(T1n);
        C2_ = normalized(C2n);
        break;
    }
    case 3: {   // bottom_move + renormalize_bottom (CTM.cpp:1120-1183, 1560-1641)
        TP T4C4 = con(T4_, {-4, 1, -2, -3}, C4_, {1, -1}, 3);
        TP C4n = con(tw(P1, {1, 2, 3}), {-2, 1, 2, 3}, T4C4, {1, 2, 3, -1}, 1);
        TP T2C3 = con(T2_, {-3, -4, -1, 1}, C3_, {1, -2}, 1);
        TP C3n = con(T2C3, {-1, 1, 2, 3}, tw(P2, {0, 1, 2}), {1, 2, 3, -2}, 2);
        TP P1T3 = con(tw(P1, {1, 2, 3}), {-1, 1, -2, -3}, tw(T3_, {0}), {-4, -5, -6, 1}, 3);
        TP P1T3A = con(P1T3, {-1, 1, -2, 2, -3, -4}, A_, {-5, -6, 1, 2, -7}, 4);
        TP P1T3AAdag = con(P1T3A, {-1, 1, 2, -4, -5, -2, 3}, tw(Adag_, {3}), {-6, -3, 3, 1, 2}, 3);
        TP T3n = con(P1T3AAdag, {-4, -1, -2, 1, 2, 3}, tw(P2, {0, 1, 2}), {1, 2, 3, -3}, 3);
        C4_ = normalized(C4n);
        T3_ = normalized(T3n);
        C3_ = normalized(C3n);
        break;
    }
    }
}

void Ctm::grow_all()
{
    move(0);
    move(1);
    move(2);
    move(3);
}

// E(A) = avg(rho_h, O_h) + avg(rho_v, O_v) after `steps` grow_all() from CTM_INIT::FROM_A, and dE/dA by the reverse sweep over the tape
// (the energy functional the reference's optimiser differentiates, src/PEPS/CTMSolver.cpp:40-75, for a 1x1 unit cell).
double ctm_energy_grad(Context& ctx, const Tensor& A, int64_t chi, int steps, const Tensor& op_h, const Tensor& op_v, bool checkpoint, Tensor** dA)
{
    if(ctx.tape) throw std::invalid_argument("ctm_energy_grad: a tape is already recording on this context");
    Tape tape(ctx);
    struct Guard
    {
        Context& c;
        ~Guard() { c.tape = nullptr; }
    } guard{ctx};
    ctx.tape = dA ? &tape : nullptr;
    double energy = 0.0;
    {
        Ctm ctm(ctx, A, chi);
        ctm.checkpoint_moves = checkpoint;
        for(int s = 0; s < steps; ++s) ctm.grow_all();
        for(int which = 0; which < 2; ++which) {
            TP rho(ctm.rdm(which));
            TP a(ctx.twist(*rho, 0)), b(ctx.twist(*a, 1));
            TP prod(ctx.multiply(*b, false, which == 0 ? op_h : op_v, false));   // avg(): src/PEPS/LinearAlgebra.cpp:57-60
            energy += ctx.trace(*prod);
            if(dA) tape.seed_trace(*prod, 1.0);
        }
    }
    if(dA) {
        tape.backward();
        Tensor* g = tape.grad(A);
        ctx.tape = nullptr;
        *dA = g ? ctx.copy(*g) : new Tensor(&ctx, A.layout, true);
    }
    return energy;
}

const Tensor& Ctm::get(int which) const
{
    const TP* all[10] = {&C1_, &C2_, &C3_, &C4_, &T1_, &T2_, &T3_, &T4_, &A_, &Adag_};
    if(which < 0 || which > 9) throw std::invalid_argument("CTM: tensor selector must be 0..9");
    return **all[which];
}

void Ctm::set(int which, Tensor* t)
{
    TP* all[10] = {&C1_, &C2_, &C3_, &C4_, &T1_, &T2_, &T3_, &T4_, &A_, &Adag_};
    if(which < 0 || which > 9) {
        delete t;
        throw std::invalid_argument("CTM: tensor selector must be 0..9");
    }
    all[which]->reset(t);
}

}   // namespace xb
