"""Reverse-mode AD on the device (include/Xped/AD/ADTensor.hpp rules on a slim host tape, xb_tape_*): gradients through the C ABI against central finite
differences, the shape of the reference's tests/ad.cpp:44-590 (a scalar loss of a chain of tensor ops, perturb single stored elements), and the CTM energy
gradient with / without checkpoint replay (src/PEPS/CTM.cpp:399-419,976-987)."""
import numpy as np
import pytest

import xped_b200 as xb
import xped_oracle as xo
from helpers import SPECS, rand_entries
from xped_oracle.ctm import heisenberg_bond_SU2, init_basis

pytestmark = pytest.mark.gpu


def fd_check(ctx, inputs, loss_fn, n_probe=6, h=1e-6, tol=2e-6, seed=0):
    """inputs: list of tensors; loss_fn(list) -> (loss value, tensor whose trace is the loss or None).  Compares the tape gradient with central differences
    on n_probe random stored elements of every input."""
    ctx.tape_begin()
    try:
        val, root = loss_fn(inputs)
        ctx.tape_seed_trace(root, 1.0)
        ctx.tape_backward()
        grads = [ctx.tape_grad(t) for t in inputs]
    finally:
        ctx.tape_end()
    rng = np.random.default_rng(seed)
    worst = 0.0
    for t, g in zip(inputs, grads):
        base = t.download()
        gd = g.download() if g is not None else np.zeros_like(base)
        assert gd.shape == base.shape
        for k in rng.choice(base.size, size=min(n_probe, base.size), replace=False):
            vals = []
            for s in (+1, -1):
                d = base.copy()
                d[k] += s * h
                t.upload(d)
                vals.append(loss_fn(inputs)[0])
            t.upload(base)
            fd = (vals[0] - vals[1]) / (2 * h)
            err = abs(fd - gd[k]) / max(1.0, abs(fd))
            worst = max(worst, err)
            assert err < tol, f"d loss / d element {k}: tape {gd[k]:.10e}, finite difference {fd:.10e}"
    return worst


@pytest.mark.parametrize("name", ["SU2", "U1", "fU1xSU2"])
def test_tape_rules_against_finite_differences(name):
    ctx = xb.Context(SPECS[name])
    try:
        rng = np.random.default_rng(5)
        e1, e2 = rand_entries(SPECS[name], 6, rng, 3), rand_entries(SPECS[name], 5, rng, 2)
        B1, B2 = ctx.basis(e1), ctx.basis(e2)
        A = ctx.tensor([B1, B2], [B1, B2]).setRandom(1)
        B = ctx.tensor([B1, B2], [B1, B2]).setRandom(2)

        def loss(ts):   # permute, contract, operator*, adjoint, twist, + / scalar *, maxNorm scaling, trace-normalisation, trace
            a, b = ts
            x = a.permute(1, (2, 0, 3, 1)).permute(-1, (1, 3, 0, 2))         # back to (2,2) in another leg order and back
            y = x.contract([-1, -2, 1, 2], b, [1, 2, -3, -4], 2)
            z = (y @ a.adjoint()).twist(0).twist(2)
            w = z.axpby(0.7, -0.3, b @ b).normalize_max()
            idw = w.clone().setIdentity()
            v = w.normalize_trace(idw) @ a
            return v.trace(), v
        fd_check(ctx, [A, B], loss)

        def loss_svd(ts):   # tSVD (full SVD + truncation rule), diag_sqrt / diag_inv, projector-like products
            a, = ts
            m = a @ a.adjoint() + a                                          # generic, well separated spectrum
            U, S, Vd = m.tSVD(4, 1e-12, False)
            isq = S.diag_sqrt().diag_inv()
            p = (isq @ U.adjoint()) @ m @ (Vd.adjoint() @ isq)
            q = (U @ S.diag_sqrt()) @ Vd
            out = q @ m.adjoint() + q.multiply(q, False, True) + (U @ p) @ U.adjoint()
            return out.trace(), out
        fd_check(ctx, [A], loss_svd, tol=5e-5)
    finally:
        ctx.close()


def heisenberg_setup(ctx, D):
    osym = xo.Symmetry.SU2()
    ophys = xo.Qbasis.from_list(osym, [((2,), 1)])
    aux = init_basis(osym, D, ophys)
    entries = [(tuple(q), d) for q, _, d in aux.data]
    a, p = ctx.basis(entries), ctx.basis([((2,), 1)])
    A = ctx.tensor([a, a], [a, a, p]).setRandom(3)
    bond = heisenberg_bond_SU2(osym, ophys)
    H = ctx.tensor([p, p], [p, p])
    assert H.structure() == bond.structure()
    H.upload(bond.data)
    return A, H


def generic_env_energy(ctx, A, H, chi, env_seed, tape):
    """E(A) of one CTM step started from a GENERIC environment (counter-based random C1..C4, T1..T4 on the bases of the FROM_A environment), RDMs and
    bond energies included.  With tape=True the ops are recorded and (E, dE/dA) is returned."""
    if tape:
        ctx.tape_begin()
    try:
        ctm = xb.CTM(A, chi)
        for k, name in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
            ctm.set(name, ctm.get(name).setRandom(env_seed + k))
        ctm.grow_all()
        e, roots = 0.0, []
        for rho in (ctm.computeRDM_h(), ctm.computeRDM_v()):
            prod = rho.twist(0).twist(1) @ H
            e += prod.trace()
            roots.append(prod)
        if not tape:
            return e, None
        for r in roots:
            ctx.tape_seed_trace(r, 1.0)
        ctx.tape_backward()
        return e, ctx.tape_grad(A)
    finally:
        if tape:
            ctx.tape_end()


@pytest.mark.parametrize("D,chi", [(2, 5), (3, 7)])
def test_ctm_step_gradient_generic_environment(D, chi):
    """d E / d A through one full CTMRG step (4 moves: 4 tSVD with truncation to chi, 64 contractions, max-norm renormalisations), computeRDM_h / computeRDM_v,
    their normalisation and the bond energies: tape vs central finite differences.  The environment is generic (no symmetry-protected degenerate singular values:
    the FROM_A trajectory has them and truncates inside such pairs, where the energy is not differentiable and the reference's rule blows up just the same)."""
    ctx = xb.Context(SPECS["SU2"])
    try:
        A, H = heisenberg_setup(ctx, D)
        e0, g = generic_env_energy(ctx, A, H, chi, 40, tape=True)
        ev, _ = generic_env_energy(ctx, A, H, chi, 40, tape=False)
        assert abs(e0 - ev) < 1e-12 * max(1.0, abs(e0))
        gd = g.download()
        assert np.linalg.norm(gd) > 0
        base = A.download()
        rng = np.random.default_rng(1)
        h = 1e-6
        scale = max(1.0, np.abs(gd).max())
        for k in rng.choice(base.size, size=min(8, base.size), replace=False):
            vals = []
            for s in (+1, -1):
                d = base.copy()
                d[k] += s * h
                A.upload(d)
                vals.append(generic_env_energy(ctx, A, H, chi, 40, tape=False)[0])
            A.upload(base)
            fd = (vals[0] - vals[1]) / (2 * h)
            assert abs(fd - gd[k]) < 2e-5 * scale, f"element {k}: tape {gd[k]:.8e} vs finite difference {fd:.8e}"
    finally:
        ctx.close()


@pytest.mark.parametrize("D,chi,steps", [(2, 6, 2), (2, 4, 3)])
def test_ctm_energy_grad_entry_point(D, chi, steps):
    """xb_ctm_energy_grad (FROM_A start): value with / without tape, and the checkpointed reverse sweep (one replayed checkpoint per move,
    src/PEPS/CTM.cpp:976-987) reproduces the gradient of the fully recorded one.  (D = 2: at D = 3 the FROM_A trajectory truncates inside exactly degenerate
    pairs of singular values, where F_ij = 1 / (s_j^2 - s_i^2) of the SVD rule amplifies rounding differences between the two evaluation orders.)"""
    ctx = xb.Context(SPECS["SU2"])
    try:
        A, H = heisenberg_setup(ctx, D)
        e0, g = xb.ctm_energy_grad(A, chi, steps, H, H)
        e1, gc = xb.ctm_energy_grad(A, chi, steps, H, H, checkpoint=True)
        ev, _ = xb.ctm_energy_grad(A, chi, steps, H, H, want_grad=False)
        assert abs(e0 - e1) < 1e-12 * max(1.0, abs(e0)) and abs(e0 - ev) < 1e-12 * max(1.0, abs(e0))
        gd, gcd = g.download(), gc.download()
        assert np.linalg.norm(gd) > 0 and np.all(np.isfinite(gd))
        assert np.linalg.norm(gd - gcd) < 1e-9 * np.linalg.norm(gd), "checkpoint replay changes the gradient"
    finally:
        ctx.close()
