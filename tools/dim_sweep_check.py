"""configs[4] size sweep, T*T only: time per product for several repetition counts (the first timed product of a fresh output buffer re-prepares the tile lists
when the product has split-K problems: host work inside the timed region)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xped_b200 as xb


def u1_basis(ctx, qmax, dim):
    return ctx.basis([(q, dim) for q in range(-qmax, qmax + 1)])


ctx = xb.Context([("U1", 0, 0)])
for dim in (16, 32):
    B = u1_basis(ctx, 4, dim)
    T = ctx.tensor([B, B], [B, B]).setRandom(0)
    for reps in (2, 2, 8):
        ms, fl = T.bench_multiply(T, reps=reps)
        print(f"dim {dim} reps {reps}: {ms:.3f} ms per product, {fl / ms / 1e9:.2f} TFLOP/s; stats {ctx.stats()}", flush=True)
    del T
