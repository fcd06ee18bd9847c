"""Peer-memory form of the sharded path vs the NCCL form and vs the unsharded tensor (run under torchrun, >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [dim]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import xped_b200 as xb
from xped_b200.sharded import Comm, ShardedTensor, enable_peer_memory

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
DIM = int(sys.argv[1]) if len(sys.argv) > 1 else 24
PERMS = [(0, (1, 0, 2, 3)), (1, (2, 3, 0, 1)), (-1, (3, 2, 1, 0)), (0, (0, 2, 1, 3)), (2, (0, 1, 3, 2))]


def log(*a):
    if rank == 0:
        print(*a, flush=True)


def parity(dim):
    ctx = xb.Context([("U1", 0, 0)], device=local)
    comm = Comm()
    B = ctx.basis([(q, dim) for q in range(-3, 4)])
    full = ctx.tensor([B, B], [B, B]).setRandom(5)
    T = ShardedTensor.zeros(ctx, comm, [B, B], [B, B]).setRandom(5)
    ref = {}
    for shift, p in PERMS:
        ref[(shift, p)] = (T.permute(shift, p).gather(), full.permute(shift, p).download())
        assert np.array_equal(*ref[(shift, p)]), "NCCL form != unsharded"
    prod_full = (full @ full).download()
    enable_peer_memory(ctx, comm, 1 << 30)
    T = ShardedTensor.zeros(ctx, comm, [B, B], [B, B]).setRandom(5)
    for shift, p in PERMS:
        r = T.permute(shift, p)
        assert np.array_equal(r.gather(), ref[(shift, p)][1]), f"push form != unsharded for {shift} {p}"
        back = r.permute(0, (0, 1, 2, 3), owner_out=[(i + 1) % world for i in range(len(r.sectors))])   # redistribution through the push form
        assert np.array_equal(back.gather(), ref[(shift, p)][1])
    log("push permutes: bit-exact vs unsharded")
    worst = 0.0
    n = len(T.sectors)
    for owner in (None, [0] * n, [(world - 1) if i % 2 else 0 for i in range(n)]):
        ctx._panel_rows_min = 16
        Ts = ShardedTensor.zeros(ctx, comm, [B, B], [B, B], owner=owner).setRandom(5)
        g = (Ts @ Ts).gather()
        worst = max(worst, float(np.linalg.norm(g - prod_full) / np.linalg.norm(prod_full)))
    log("row-panel products: worst rel. error", worst)
    assert worst < 1e-13
    c = T.contract([-1, -2, 1, 2], T, [1, 2, -3, -4], 2).gather()
    cf = full.contract([-1, -2, 1, 2], full, [1, 2, -3, -4], 2).download()
    err = float(np.linalg.norm(c - cf) / np.linalg.norm(cf))
    log("sharded contract: rel. error", err)
    assert err < 1e-13
    ctx.close()


def timing(dim):
    ctx = xb.Context([("U1", 0, 0)], device=local)
    comm = Comm()
    B = ctx.basis([(q, dim) for q in range(-4, 5)])

    def timed(fn, reps):
        fn(); ctx.sync(); dist.barrier()
        ctx.timer_start()
        for _ in range(reps):
            fn()
        ms = ctx.timer_stop() / reps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    rows = {}
    for mode in ("nccl", "peer"):
        if mode == "peer":
            enable_peer_memory(ctx, comm, 12 << 30)
        T = ShardedTensor.zeros(ctx, comm, [B, B], [B, B]).setRandom(0)
        nelem = T.plainSize()
        flops = sum(2.0 * r * c * c for r, c in zip(T.rows, T.cols))
        for shift, p in PERMS[:4]:
            ms = timed(lambda: T.permute(shift, p), 3)
            rows.setdefault((shift, p), {})[mode] = (ms, 16.0 * nelem / ms / 1e6)
        ms = timed(lambda: T @ T, 2)
        rows.setdefault("T*T", {})[mode] = (ms, flops / ms / 1e9)
        if mode == "peer":
            for mbps in (150000, 300000, 1200000, 10**9):
                ctx.set_int("dist.pull_mbps", mbps)
                T @ T; ctx.sync(); dist.barrier()
                ctx.timer_start()
                for _ in range(2):
                    T @ T
                mine = ctx.timer_stop() / 2
                t = torch.tensor([mine], dtype=torch.float64, device="cuda")
                box = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(box, t)
                per = [round(float(b.item()), 2) for b in box]
                log(f"T*T pull_mbps {mbps}: per-rank ms {per} -> {flops / max(per) / 1e9:.1f} TFLOP/s")
            ctx.set_int("dist.pull_mbps", 600000)
        del T
    for k, v in rows.items():
        log(k, {m: (round(a, 3), round(b, 1)) for m, (a, b) in v.items()})
    st = (xb.capi.C.c_int64 * 3)() if hasattr(xb.capi, "C") else None
    ctx.close()


parity(6)
if DIM > 0:
    timing(DIM)
dist.barrier()
dist.destroy_process_group()
log("dist_check ok")
