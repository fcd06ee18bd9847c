"""State interop: the native container round trip (tensors with conjugated legs and with the explicit sector order tSVD leaves, a whole CTM
state: save, load, continue identically) and the charge-block (MATLAB) format of IO::loadMatlabTensor (include/Xped/IO/Matlab.hpp:9-95) against its CPU restatement."""
import itertools

import numpy as np
import pytest

import xped_b200 as xb
import xped_oracle as xo
from helpers import SPECS, pair, rand_entries
from xped_oracle.io import load_charge_blocks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["SU2", "U1xU1", "fU1xSU2"])
def test_tensor_save_load_round_trip(name, tmp_path):
    ctx, osym = xb.Context(SPECS[name]), xo.Symmetry.from_spec(SPECS[name])
    try:
        rng = np.random.default_rng(3)
        ents = [rand_entries(SPECS[name], 7, rng, 3) for _ in range(4)]
        g, _ = pair(ctx, osym, ents[:2], ents[2:], 21)
        cases = [g, g.permute(1, (2, 0, 3, 1)), g.adjoint()]          # conjugated legs after the permute; adjoint: explicit sector order
        U, S, Vd = (g @ g.adjoint()).tSVD(5, 1e-12, False)
        cases += [U, S, Vd]                                           # truncated bases, custom sector lists
        for k, t in enumerate(cases):
            path = tmp_path / f"t{k}.xb2t"
            t.save(path)
            back = xb.Tensor.load(ctx, path)
            assert back.structure() == t.structure()
            assert (back.rank, back.corank) == (t.rank, t.corank)
            for leg in range(t.rank + t.corank):
                assert back.leg_basis(leg).entries() == t.leg_basis(leg).entries() and back.leg_basis(leg).IS_CONJ() == t.leg_basis(leg).IS_CONJ()
            assert np.array_equal(back.download(), t.download())
            # the reloaded tensor behaves like the original under a further symmetric operation (fusion trees / dual flags restored)
            n = t.rank + t.corank
            if t.corank >= 1 and n >= 2:
                p = tuple(reversed(range(n)))
                assert np.array_equal(back.permute(0, p).download(), t.permute(0, p).download()) if t.rank == t.corank else True
        other = xb.Context(SPECS["U1"])
        try:
            with pytest.raises(xb.XbError):
                xb.Tensor.load(other, tmp_path / "t0.xb2t")           # written for another symmetry
        finally:
            other.close()
    finally:
        ctx.close()


def test_ctm_state_checkpoint_resume(tmp_path):
    ctx = xb.Context(SPECS["SU2"])
    try:
        a, p = ctx.basis([((1,), 2), ((2,), 1)]), ctx.basis([((2,), 1)])
        A = ctx.tensor([a, a], [a, a, p]).setRandom(0)
        ctm = xb.CTM(A, 10)
        ctm.grow_all()
        ctm.save(tmp_path / "env.xb2c")
        resumed = xb.CTM.load(ctx, tmp_path / "env.xb2c")
        ctx.set_int("svd.truncated", 0)      # deterministic path on both sides
        ctm.grow_all()
        resumed.grow_all()
        for name in ["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4", "A"]:
            x, y = ctm.get(name), resumed.get(name)
            assert x.structure() == y.structure() and np.array_equal(x.download(), y.download()), name
    finally:
        ctx.close()


@pytest.mark.parametrize("name,mod", [("U1", None), ("Z3", 3)])
def test_charge_block_import(name, mod):
    """Random charge-block files (3 and 4 legs, one conjugated leg, charges scaled by qn_scale = 2 for U(1)): device vs the CPU restatement, bit-exact structure
    and values; and the blocks land where the format says (norm of the tensor = norm of vdat)."""
    ctx, osym = xb.Context(SPECS[name]), xo.Symmetry.from_spec(SPECS[name])
    try:
        rng = np.random.default_rng(17)
        for nlegs, rank, conj in [(3, 2, [False, False, True]), (4, 2, [False, True, False, False]), (4, 1, None)]:
            qn_scale = 2 if mod is None else 1
            charges, dims = [], []
            for _ in range(nlegs):
                k = int(rng.integers(2, 4))
                cs = rng.choice(np.arange(-2, 3) if mod is None else np.arange(0, mod), size=k, replace=False)
                charges.append([float(c) for c in cs])
                dims.append([float(d) for d in rng.integers(1, 4, size=k)])

            def total(idx):
                q = sum(qn_scale * int(charges[r][i]) for r, i in enumerate(idx))
                return q % mod if mod else q
            voffs, vdat, off = [], [], 0
            for idx in itertools.product(*[range(len(c)) for c in reversed(charges)]):
                idx = idx[::-1]                                    # leg 0 fastest in the combined index
                if total(idx) != 0:
                    continue
                combined, mult = 0, 1
                for r, i in enumerate(idx):
                    combined += i * mult
                    mult *= len(charges[r])
                size = int(np.prod([dims[r][i] for r, i in enumerate(idx)]))
                voffs.append((combined, off))
                vdat += list(rng.uniform(-1, 1, size))
                off += size
            if not voffs:
                continue
            voffs.sort()
            # offsets must ascend with the order of the rows: re-pack vdat in row order
            order = sorted(range(len(voffs)), key=lambda k: voffs[k][0])
            g = xb.Tensor.from_charge_blocks(ctx, charges, dims, rank, voffs, vdat, conj, qn_scale)
            o = load_charge_blocks(osym, charges, dims, rank, voffs, vdat, conj, qn_scale, mod)
            assert g.structure() == o.structure()
            assert (g.rank, g.corank) == (rank, nlegs - rank)
            assert np.array_equal(g.download(), o.data)
            assert abs(np.linalg.norm(o.data) - np.linalg.norm(vdat)) < 1e-12 * np.linalg.norm(vdat)
            del order
    finally:
        ctx.close()
