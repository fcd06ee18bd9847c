"""torchrun --nproc-per-node 2 tools/debug_sharded.py : per-operation errors of the sharded path over a real NCCL group."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
    sys.path.insert(0, p)
import xped_b200 as xb, xped_oracle as xo
from xped_b200.sharded import Comm, ShardedTensor
from helpers import SPECS, obasis
from test_sharded_cpu import CASES, PERMS
from xped_oracle.rng import uniform_fill
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
for name in ("U1", "SU2"):
    spec, dom_e, cod_e = CASES[name]
    ctx = xb.Context(spec, device=rank); osym = xo.Symmetry.from_spec(spec); comm = Comm()
    dom, cod = [ctx.basis(e) for e in dom_e], [ctx.basis(e) for e in cod_e]
    ot = xo.Tensor(osym, [obasis(osym, e) for e in dom_e], [obasis(osym, e) for e in cod_e]); ot.data = uniform_fill(8, ot.plainSize())
    T = ShardedTensor.zeros(ctx, comm, dom, cod).setRandom(8)
    print(rank, name, "gather ok", np.array_equal(T.gather(), ot.data), "owner", T.owner, flush=True)
    for shift, p in PERMS:
        r, orr = T.permute(shift, p), ot.permute(shift, p)
        e = float(np.linalg.norm(r.gather() - orr.data) / np.linalg.norm(orr.data))
        if rank == 0: print(name, "permute", shift, p, "err", e, "owner_out", r.owner, flush=True)
    n2 = T.squaredNorm(); ref = float(ot.data @ ot.data); mx = T.maxNorm()
    if rank == 0: print(name, "squaredNorm", n2, ref, "maxNorm", mx, float(np.abs(ot.data).max()), flush=True)
    ctx.close()
dist.destroy_process_group()
