#!/usr/bin/env python
"""bench.py -- headline measurement of the xped_b200 hot path (contract: task statement, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras]

Headline metric (BASELINE.json `metric`, configs[1]): CTMRG steps/s of the SU(2) iPEPS environment at D=4, chi=64,
1x1 unit cell, CTM_INIT::FROM_A, one "step" = CTM::grow_all() = left, right, top, bottom move
(src/PEPS/CTM.cpp:385-420), RDM/energy evaluation excluded.  The other two parts of the metric -- permute HBM GB/s and
contract FP64 TFLOP/s -- are measured in the same run on the W4 / W3 tensors of BASELINE.md section 4 and reported under
"kernels"; the `roofline` object is that of the permute kernel (HBM-bound), `roofline_contract` that of the grouped DGEMM.

  value  : steps/s with the environment resident in HBM, CUDA events on the context stream, L2 flushed between steps.
  e2e    : the same step through the C ABI with HOST buffers: every step uploads C1..C4, T1..T4 and A from pinned host
           memory (xb_tensor_upload), calls xb_ctm_step and downloads the new environment (xb_tensor_download).
  --impl reference : the CPU oracle (oracle/xped_oracle, a numpy restatement of the reference's algorithm; the reference
           itself cannot be built here) on the host cores, same config / metric / unit.
N > 1: sectors of one D=4 CTM are too small to shard (9 sectors, < 6 MB): every rank advances its own independent
environment (different seed) -- "replicas", weak scaling, no data-path collective.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_AUX, CHI = 4, 64
ENV = ["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the W3/W4 kernel sections (headline only)")
    ap.add_argument("--cpu-steps", type=int, default=3, help="oracle steps timed for cpu_baseline")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------------------------- oracle (CPU) arm
def oracle_ctm(seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import xped_oracle as xo
    from xped_oracle.ctm import CTM, init_basis
    from xped_oracle.rng import uniform_fill
    sym = xo.Symmetry.SU2()
    phys = xo.Qbasis.from_list(sym, [((2,), 1)])
    aux = init_basis(sym, D_AUX, phys)
    A = xo.Tensor(sym, [aux, aux], [aux, aux, phys])
    A.data = uniform_fill(seed, A.plainSize())
    return CTM(A, CHI)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


CPU_KIND_NOTE = ("oracle/xped_oracle: restatement of the reference algorithm (src/PEPS/CTM.cpp, src/Core/Tensor.cpp) executed the way an optimised "
                 "Eigen + BLAS build runs it -- recoupling coefficients cached per (tensor structure, shift, permutation) like XPED_CACHE_PERMUTE_OUTPUT, "
                 "sub-block transposes in the compiled kernel oracle/csrc/xped_oracle_kernels.c (serial like the reference's loop), per-block GEMM and "
                 "gesdd SVD in OpenBLAS / LAPACK with all host threads")


def install_fast_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from xped_oracle import fast
    fast.install()


def time_oracle(warmup, steps):
    import numpy as np  # noqa: F401  (loads the BLAS the oracle's GEMM / SVD calls use)
    install_fast_oracle()
    ctm = oracle_ctm(0)
    # same workload state as the device arm: grow until the chi bases are saturated (set-up), then the warm-up steps
    prev, same, n = None, 0, 0
    while n < 12 and same < 2:
        ctm.grow_all()
        n += 1
        t = ctm.tensors()
        cur = tuple(t[nm].structure()[1:3] for nm in ("C1", "T1"))
        same = same + 1 if cur == prev else 0
        prev = cur
    for _ in range(warmup):
        ctm.grow_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctm.grow_all()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3


def config_dict(n_gpus, extra=None):
    cfg = {"workload": "configs[1]: SU(2) iPEPS CTMRG, D=4 (aux {q=1:2,q=2:2}), chi=64, 1x1 cell, FROM_A init, step = grow_all() (4 moves)",
           "symmetry": "SU2", "D": D_AUX, "chi": CHI, "unit_cell": "1x1",
           "parallelism": "single GPU" if n_gpus == 1 else f"{n_gpus} replicas (one independent environment per GPU, no data-path collective)",
           "l2": "flushed between timed steps (256 MiB memset outside the event pairs)"}
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use all the host threads it can (set before numpy loads its BLAS)
    if "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(os.cpu_count() or 1)
    sps, ms = time_oracle(args.warmup, args.steps)
    cores = blas_threads()
    line = {"impl": "reference", "metric": "ctmrg_steps_per_s", "value": sps, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (counter-based uniform[-1,1] site tensor, seed 0)", "config": config_dict(1),
            "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} grow_all() steps after {args.warmup} warm-up steps; " + CPU_KIND_NOTE},
            "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- B200 arm
def u1_basis(ctx, qmax, dim):
    return ctx.basis([(q, dim) for q in range(-qmax, qmax + 1)])


def su2_corner_pair(ctx, aux, env, seeds):
    """two random corner-shaped tensors Q(3,3) on (env, aux, aux | env, aux, aux)"""
    a, e = ctx.basis(aux), ctx.basis(env)
    return [ctx.tensor([e, a, a], [e, a, a]).setRandom(s) for s in seeds]


def oracle_tensor_like(xo, osym, g):
    """oracle tensor with the bases and the storage image of a device tensor"""
    r, c = g.rank, g.corank
    legs = []
    for i in range(r + c):
        b = g.leg_basis(i)
        ent = [(tuple(q), d) for q, d in b.entries()]
        if b.IS_CONJ():      # a conjugated leg basis (Qbasis::conj, dual fusion trees): rebuild it as the conj() of its pre-image
            legs.append(xo.Qbasis.from_list(osym, [(tuple(osym.conj(q)), d) for q, d in ent]).conj())
        else:
            legs.append(xo.Qbasis.from_list(osym, ent))
    t = xo.Tensor(osym, legs[:r], legs[r:])
    if t.structure() != g.structure():
        raise RuntimeError("oracle / device structure mismatch")
    t.data = g.download()
    return t


def cpu_d6_left_move(xb, ctx, ctm, A):
    install_fast_oracle()
    import xped_oracle as xo
    from xped_oracle.ctm import CTM as OCTM
    osym = xo.Symmetry.SU2()
    oA = oracle_tensor_like(xo, osym, A)
    octm = OCTM(oA, 144)
    for nm in ENV:
        setattr(octm, nm, oracle_tensor_like(xo, osym, ctm.get(nm)))
    t0 = time.perf_counter()
    octm.move(0)
    dt = time.perf_counter() - t0
    ctx.sync()
    ctx.timer_start()
    ctm.move(0)
    gpu_ms = ctx.timer_stop()
    return {"cpu_ms": dt * 1e3, "gpu_ms_same_move": gpu_ms, "cores": blas_threads(), "kind": "port",
            "sample": "ONE left move (a quarter of a grow_all() step) at D=6 chi=144 from the device's environment after 5 steps; " + CPU_KIND_NOTE}


def kernel_sections(xb, device, torch, cpu_d6=False):
    """permute HBM GB/s and contract FP64 TFLOP/s on the large tensors (W4, W3)"""
    out = {}
    # ---- FP64 tensor peak: cuBLAS DGEMM 8192^3 measured here (MEASURED_PEAKS.json has no FP64 entry)
    n = 8192
    a = torch.rand(n, n, dtype=torch.float64, device="cuda")
    b = torch.rand(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    dgemm_peak = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    out["fp64_peak_tflops_cublas_dgemm_8192"] = dgemm_peak

    # ---- W4: U(1) Tensor<2,2>, legs q in -4..4, 24 states per charge
    ctx = xb.Context([("U1", 0, 0)], device=device)
    B = u1_basis(ctx, 4, 24)
    T = ctx.tensor([B, B], [B, B]).setRandom(0)
    nelem = T.plainSize()
    perms = [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (0, 2, 1, 3), (3, 2, 1, 0)]
    rows = []
    for p in perms:
        for shift in (0, 1, -1, 2, -2):
            T.bench_permute(shift, p, reps=2)            # plan build + warm-up
            ctx.flush_l2()
            ms, nbytes = T.bench_permute(shift, p, reps=5)  # 2.6 GB per launch >> L2
            rows.append({"p": list(p), "shift": shift, "ms": ms, "GBps": nbytes / ms / 1e6})
    gbs = sorted(r["GBps"] for r in rows)
    tot_bytes = sum(2 * 8 * nelem for _ in rows)
    tot_ms = sum(r["ms"] for r in rows)
    out["permute_w4"] = {"tensor": f"U(1) Tensor<2,2>, leg q=-4..4 x 24, {nelem} elements", "cases": len(rows), "GBps_min": gbs[0],
                         "GBps_median": gbs[len(gbs) // 2], "GBps_max": gbs[-1], "GBps_aggregate": tot_bytes / tot_ms / 1e6,
                         "worst": min(rows, key=lambda r: r["GBps"]), "best": max(rows, key=lambda r: r["GBps"])}
    runs = sorted(T.bench_multiply(T, reps=2) for _ in range(3))     # three batches of two launches; the median batch is reported
    ms, fl = runs[1]
    out["contract_w4"] = {"op": "T*T (17 sectors, block side (9-|Q|)*576)", "flops": fl, "ms": ms, "TFLOPs": fl / ms / 1e9,
                          "frac_of_cublas_dgemm": fl / ms / 1e9 / dgemm_peak, "ms_batches": [r[0] for r in runs]}
    del T
    ctx.close()
    # ---- W3: SU(2) D=6 chi=144 corners, M = Q4*Q1
    aux3 = [(1, 3), (2, 3)]
    env3 = [(1, 16), (2, 32), (3, 36), (4, 30), (5, 18), (6, 8), (7, 4)]
    ctx3 = xb.Context([("SU2", 0, 0)], device=device)
    Q4, Q1 = su2_corner_pair(ctx3, aux3, env3, (1, 2))
    Q4.bench_multiply(Q1, reps=2)
    ctx3.sync()
    ms, fl = min(Q4.bench_multiply(Q1, reps=5) for _ in range(3))
    out["contract_w3"] = {"op": "Q4*Q1, SU(2) D=6 chi=144 corners (9 sectors, block side 36..2394)", "flops": fl, "ms": ms, "TFLOPs": fl / ms / 1e9,
                          "frac_of_cublas_dgemm": fl / ms / 1e9 / dgemm_peak}
    prow = []
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5))]:
        Q4.bench_permute(shift, p, reps=2)
        ctx3.flush_l2()
        ms, nbytes = Q4.bench_permute(shift, p, reps=5)
        prow.append({"p": list(p), "shift": shift, "ms": ms, "GBps": nbytes / ms / 1e6})
    out["permute_w3"] = {"tensor": f"SU(2) corner Q(3,3), {Q4.plainSize()} elements (fan-in > 1: recoupling sums)", "cases": prow}
    M = Q4 @ Q1
    ctx3.sync()
    ctx3.timer_start()
    U, S, Vd = M.tSVD(144, 1e-10, False)
    cold_ms = ctx3.timer_stop()   # first call on this layout: plan construction, CUDA-graph capture, no warm start
    del U, S, Vd
    warm = []
    for _ in range(3):
        ctx3.sync()
        ctx3.timer_start()
        U, S, Vd = M.tSVD(144, 1e-10, False)
        warm.append(ctx3.timer_stop())
    out["tsvd_w3"] = {"op": "tSVD(144, 1e-10) of M = Q4*Q1 of two RANDOM corners (blocks 36..2394, 10800 singular values)", "ms": min(warm), "ms_calls": warm,
                      "ms_first_call": cold_ms, "fallbacks_to_full_jacobi": ctx3.get_int("svd.truncated_fallbacks"),
                      "note": "the spectrum of a product of random corners is nearly flat: the subspace iteration cannot separate 144 of 10800 values, its check "
                              "hands over to the full batched Jacobi path (all 10800 triplets).  The projector matrices of a CTM run decay: see "
                              "ctm_d6_chi144.tsvd_ms_per_call for the same call inside the D=6 step (truncated path)"}
    del U, S, Vd, M, Q4, Q1
    # ---- the north-star target workload: one CTMRG step at D=6, chi=144 (SU(2), 1x1 cell, FROM_A; chi saturates after two steps)
    try:
        a6, phys6 = ctx3.basis(aux3), ctx3.basis([(2, 1)])
        A6 = ctx3.tensor([a6, a6], [a6, a6, phys6]).setRandom(0)
        ctm6 = xb.CTM(A6, 144)
        for _ in range(6):   # chi saturates after two to three steps; plans used for the second time build their table-driven form (host work)
            ctm6.grow_all()
        steps6 = []
        for _ in range(2):   # two steady-state steps
            ctx3.sync()
            ctx3.profile(True)
            ctx3.profile_read()
            ctx3.timer_start()
            ctm6.grow_all()
            ms6 = ctx3.timer_stop()
            prof6 = ctx3.profile_read()
            ctx3.profile(False)
            steps6.append({"ms": ms6, "device_ms": {k: v[0] for k, v in prof6.items()}, "launch_groups": {k: v[1] for k, v in prof6.items()},
                           "permute_GBps": prof6["permute"][2] / max(prof6["permute"][0], 1e-9) / 1e6,      # algorithmic bytes of the step's permutes / their device time
                           "gemm_TFLOPs": prof6["gemm"][2] / max(prof6["gemm"][0], 1e-9) / 1e9})
        best = min(steps6, key=lambda r: r["ms"])
        out["ctm_d6_chi144"] = {"op": "grow_all() at D=6 (aux {1:3, 2:3}), chi=144, steps 7 and 8 from FROM_A (chi bases saturated), entirely on the device", "ms_per_step": best["ms"],
                                "tsvd_ms_per_call": best["device_ms"]["tsvd"] / max(best["launch_groups"]["tsvd"], 1), "steps": steps6,
                                "tsvd": {"truncated_calls": ctx3.get_int("svd.truncated_calls"), "fallbacks_to_full_jacobi": ctx3.get_int("svd.truncated_fallbacks"),
                                         "subspace_iterations": ctx3.get_int("svd.truncated_iterations")}}
        if cpu_d6:
            # the reference algorithm on the host cores at the SAME size: one left move (1 of the 4 moves of a step: two corners, one tSVD of blocks up to
            # ~3060^2, renormalisation) from the environment the device has now
            try:
                out["ctm_d6_chi144"]["cpu_left_move"] = cpu_d6_left_move(xb, ctx3, ctm6, A6)
            except Exception as e:
                out["ctm_d6_chi144"]["cpu_left_move"] = {"error": f"{type(e).__name__}: {e}"}
        del ctm6, A6
    except Exception as e:
        out["ctm_d6_chi144"] = {"error": f"{type(e).__name__}: {e}"}
    # ---- configs[4] size sweep on one GPU: the same U(1) Tensor<2,2> with 8 / 16 / 24 / 32 states per charge
    try:
        sweep = []
        ctxs = xb.Context([("U1", 0, 0)], device=device)
        for dim in (8, 16, 32):
            Bs = u1_basis(ctxs, 4, dim)
            Ts = ctxs.tensor([Bs, Bs], [Bs, Bs]).setRandom(0)
            tot_ms, tot_b = 0.0, 0.0
            for p in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (0, 2, 1, 3), (3, 2, 1, 0)]:
                Ts.bench_permute(0, p, reps=1)
                ctxs.flush_l2()
                ms, nbytes = Ts.bench_permute(0, p, reps=3)
                tot_ms += ms
                tot_b += nbytes
            # best of three two-product batches: the first batch after the permutes can carry a pool growth for the fresh result buffers (up to 4 GB at dim 32)
            # and the re-preparation of the tile lists for a new result address inside its timed region (seen once: 10 / 17 instead of 28.6 / 29.6 TFLOP/s)
            ms, fl = min((Ts.bench_multiply(Ts, reps=2) for _ in range(3)), key=lambda r: r[0])
            sweep.append({"dim_per_charge": dim, "elements": Ts.plainSize(), "permute_GBps_aggregate": tot_b / tot_ms / 1e6, "contract_TFLOPs": fl / ms / 1e9})
            del Ts
        out["w4_dim_sweep"] = {"op": "5 permutes (shift 0) + T*T of the U(1) Tensor<2,2>, leg q=-4..4 x dim; dim = 24 is permute_w4 / contract_w4 above", "cases": sweep}
        # ---- configs[0a]: the small U(1) Tensor<2,2> (q = -2..2 x 4 states: ~2e4 elements): latency-bound, microseconds per operation
        Bm = u1_basis(ctxs, 2, 4)
        Tm = ctxs.tensor([Bm, Bm], [Bm, Bm]).setRandom(0)
        import itertools
        us = []
        for shift in (-2, -1, 0, 1, 2):
            for p in itertools.permutations(range(4)):
                Tm.bench_permute(shift, p, reps=1)
                ms, _ = Tm.bench_permute(shift, p, reps=20)
                us.append(1e3 * ms)
        us.sort()
        ms_mm, _ = Tm.bench_multiply(Tm, reps=50)
        out["small_u1"] = {"op": "configs[0a]: U(1) Tensor<2,2>, leg q=-2..2 x 4 (%d elements): all 24 x 5 permute<shift,p> and T*T" % Tm.plainSize(),
                           "permute_us_median": us[len(us) // 2], "permute_us_max": us[-1], "multiply_us": 1e3 * ms_mm}
        del Tm
        ctxs.close()
    except Exception as e:
        out["w4_dim_sweep"] = {"error": f"{type(e).__name__}: {e}"}
    # ---- configs[2]: U(1) x U(1) Hubbard MPS transfer contraction at bond dimension ~2000 (89 sectors, blocks <= 159)
    try:
        from xped_b200.workloads import hubbard_bond_basis
        ctxh = xb.Context([("U1", 0, 1), ("U1", 0, 0)], device=device)
        bond = hubbard_bond_basis(2000)
        loc = [((0, 0), 1), ((1, 1), 1), ((1, -1), 1), ((2, 0), 1)]
        bb, lb = ctxh.basis(bond), ctxh.basis(loc)
        comb = dict((tuple(q), d) for q, d in bb.combine(lb).entries())
        outb = ctxh.basis([(q, min(d, comb[q])) for q, d in bond if q in comb])
        Ah = ctxh.tensor([bb, lb], [outb]).setRandom(5)
        Lh = ctxh.tensor([bb], [bb]).setRandom(6)
        Rh = ctxh.tensor([outb], [outb]).setRandom(7)

        def timed_us(fn, reps=20):
            fn()
            ctxh.sync()
            ctxh.timer_start()
            for _ in range(reps):
                fn()
            return 1e3 * ctxh.timer_stop() / reps
        out["hubbard_m2000"] = {"op": "configs[2]: contract_L / contract_R on A(2,1), bond basis %d states in %d sectors, four-state Hubbard site" %
                                      (sum(d for _, d in bond), len(bond)), "A_elements": Ah.plainSize(),
                                "contract_L_us": timed_us(lambda: xb.contract_L(Lh, Ah, Ah)), "contract_R_us": timed_us(lambda: xb.contract_R(Rh, Ah, Ah)),
                                "permute_multiply_formulation_us": timed_us(lambda: (Ah @ Rh).permute(1, (0, 1, 2)).multiply(Ah.permute(1, (0, 1, 2)), False, True))}
        del Ah, Lh, Rh
        ctxh.close()
    except Exception as e:
        out["hubbard_m2000"] = {"error": f"{type(e).__name__}: {e}"}
    # ---- configs[0]: the reference's own benchmarks/bench.cpp workload (SU(2) spin-1/2 chain, L = 20, Minit = 200, Qinit = 10:
    # <Psi|Psi> by dot() ten times, then one sweep of SVD steps); latency-bound at these sizes: reported as ms
    try:
        psi = xb.random_mps(ctx3, 20, [((2,), 1)], (1,), 200, 10, seed=0)
        xb.dot(psi, psi, "RIGHT")
        ctx3.sync()
        ctx3.timer_start()
        for _ in range(10):
            nrm = xb.dot(psi, psi, "RIGHT")
        ms_dot = ctx3.timer_stop() / 10
        ctx3.timer_start()
        for l in range(psi.length()):
            psi.rightSweepStep(l)
        ms_sweep = ctx3.timer_stop()
        out["mps_bench_cpp"] = {"op": "benchmarks/bench.cpp: SU(2) spin-1/2 chain, L=20, Minit=200, Qinit=10", "dot_ms": ms_dot, "right_sweep_ms": ms_sweep,
                                "norm_before_sweep": nrm, "norm_after_sweep": xb.dot(psi, psi, "RIGHT"),
                                "bond_dims": [sum(d for _, d in psi.Ac[l].leg_basis(2).entries()) for l in range(psi.length())]}
    except Exception as e:
        out["mps_bench_cpp"] = {"error": f"{type(e).__name__}: {e}"}
    ctx3.close()
    return out, dgemm_peak


def sharded_section(xb, torch, dist, local, world):
    """BASELINE.json configs[4]: the W4 contraction + permute sweep sharded by sector over the ranks.
    Peer-memory form (csrc/dist.cu): every rank maps every other rank's device arena (CUDA IPC over NVLink / NVSwitch); a sharded permute is ONE
    launch per rank that transposes the sub-blocks the rank owns straight into the output shard of the destination's owner (no buffers, no
    collective), T*T is balanced by row panels of the heavy sectors (operands pulled, result panels stored over NVLink).  The pack -> NCCL
    alltoallv -> gather form of the same permutes is timed beside it.  Times: CUDA events on the context stream, max over ranks; bytes / flops
    are those of the GLOBAL tensor.  Parity: on a small tensor of the same structure every sharded result is compared with the unsharded one on
    every rank inside this run (bit-exact permutes, products <= 1e-13)."""
    import numpy as np
    from xped_b200.sharded import Comm, ShardedTensor, enable_peer_memory
    ctx = xb.Context([("U1", 0, 0)], device=local)
    comm = Comm()
    perms = [(0, (1, 0, 2, 3)), (1, (2, 3, 0, 1)), (-1, (3, 2, 1, 0)), (0, (0, 2, 1, 3))]

    def timed(fn, reps):
        fn()                       # warm-up: plans, exchange buffers, NCCL channels
        ctx.sync()
        dist.barrier()
        ctx.timer_start()
        for _ in range(reps):
            fn()
        ms = ctx.timer_stop() / reps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the NCCL form first (before the arena exists)
    B = u1_basis(ctx, 4, 24)
    T = ShardedTensor.zeros(ctx, comm, [B, B], [B, B]).setRandom(0)
    nelem = T.plainSize()
    flops = sum(2.0 * r * c * c for r, c in zip(T.rows, T.cols))
    nccl = {"permutes": []}
    for shift, p in perms:
        ms = timed(lambda: T.permute(shift, p), 3)
        nccl["permutes"].append({"p": list(p), "shift": shift, "ms": ms, "GBps_global": 16.0 * nelem / ms / 1e6})
    ms = timed(lambda: T @ T, 2)
    nccl["multiply"] = {"op": "T*T, whole sectors per rank (LPT on flops)", "ms": ms, "TFLOPs_global": flops / ms / 1e9}
    del T
    # ---- peer-memory form
    enable_peer_memory(ctx, comm, 12 << 30)
    parity = {}
    Bs = u1_basis(ctx, 3, 6)
    full = ctx.tensor([Bs, Bs], [Bs, Bs]).setRandom(5)
    Ts = ShardedTensor.zeros(ctx, comm, [Bs, Bs], [Bs, Bs]).setRandom(5)
    ok = True
    for shift, p in perms + [(2, (0, 1, 3, 2))]:
        ok = ok and np.array_equal(Ts.permute(shift, p).gather(), full.permute(shift, p).download())
    parity["permutes_bit_exact"] = bool(ok)
    ref = (full @ full).download()
    worst = 0.0
    for owner in (None, [0] * len(Ts.sectors)):     # the skewed owner map forces row panels onto the other ranks
        ctx._panel_rows_min = 16
        Tp = ShardedTensor.zeros(ctx, comm, [Bs, Bs], [Bs, Bs], owner=owner).setRandom(5)
        worst = max(worst, float(np.linalg.norm((Tp @ Tp).gather() - ref) / np.linalg.norm(ref)))
    ctx._panel_rows_min = 256
    parity["multiply_rel_err"] = worst
    flag = torch.tensor([0.0 if (ok and worst < 1e-13) else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(flag)
    parity["all_ranks_ok"] = bool(flag.item() == 0.0)
    if not parity["all_ranks_ok"]:
        raise AssertionError(f"sharded parity failed: {parity}")
    del full, Ts, Tp
    T = ShardedTensor.zeros(ctx, comm, [B, B], [B, B]).setRandom(0)
    mine = sum(T.rows[i] * T.cols[i] for i in T.my_sectors())
    my_flops = sum(2.0 * T.rows[i] * T.cols[i] * T.cols[i] for i in T.my_sectors())
    out = {"tensor": f"W4: U(1) Tensor<2,2>, leg q=-4..4 x 24, {nelem} elements, 17 sectors over {world} ranks (LPT on flops)",
           "form": "peer memory (CUDA IPC arenas over NVLink): fused permute + exchange launch, row-panel balanced product",
           "local_fraction_of_elements": mine / nelem, "local_fraction_of_flops": my_flops / flops, "permutes": [], "parity": parity}
    for shift, p in perms:
        ms = timed(lambda: T.permute(shift, p), 3)
        out["permutes"].append({"p": list(p), "shift": shift, "ms": ms, "GBps_global": 16.0 * nelem / ms / 1e6})
    ms = timed(lambda: T @ T, 2)
    out["multiply"] = {"op": "T*T, row panels of the heavy sectors spread over the ranks", "flops": flops, "ms": ms, "TFLOPs_global": flops / ms / 1e9}
    ms = timed(lambda: T.contract([-1, -2, 1, 2], T, [1, 2, -3, -4], 2), 2)
    out["contract"] = {"op": "contract<{-1,-2,1,2},{1,2,-3,-4},2>(T, T): 3 sharded permutes + sharded product", "ms": ms}
    out["nccl_form"] = nccl
    del T
    # ---- the same operations unsharded on ONE GPU of this run (rank 0; the other ranks wait): the denominators of the scaling figures
    if comm.rank == 0:
        ctx1 = xb.Context([("U1", 0, 0)], device=local)
        B1 = u1_basis(ctx1, 4, 24)
        T1 = ctx1.tensor([B1, B1], [B1, B1]).setRandom(0)
        one = {"permutes": []}
        for shift, p in perms:
            T1.bench_permute(shift, p, reps=2)
            ms1, nbytes = T1.bench_permute(shift, p, reps=3)
            one["permutes"].append({"p": list(p), "shift": shift, "ms": ms1, "GBps": nbytes / ms1 / 1e6})
        ms1, fl1 = sorted(T1.bench_multiply(T1, reps=2) for _ in range(2))[0]
        one["multiply"] = {"ms": ms1, "TFLOPs": fl1 / ms1 / 1e9}
        del T1
        ctx1.close()
        out["one_gpu_same_run"] = one
        out["speedup_vs_one_gpu"] = {"permutes": [a["ms"] / b["ms"] for a, b in zip(one["permutes"], out["permutes"])],
                                     "multiply": one["multiply"]["ms"] / out["multiply"]["ms"]}
    dist.barrier()
    ctx.close()
    return out


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import xped_b200 as xb

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- xped_b200 has no CPU fallback (use --impl reference for the CPU oracle arm)")
    torch.cuda.set_device(local)
    if world > 1 and not os.environ.get("XB_BENCH_NO_PIN"):
        # one process per GPU, launch-bound steps (~1000 launches in ~21 ms): give every rank its own slice of the host cores so that the launch threads do not
        # migrate / share a core with another rank's spinning synchronisation
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = len(cpus) // world
            if per >= 1:
                os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
        except (AttributeError, OSError):
            pass
    real_stdout = None
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner to the C-level stdout when the first communicator is
        # created, so file descriptor 1 points to stderr for the rest of the run and the line goes to a duplicate of the real stdout
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    ctx = xb.Context([("SU2", 0, 0)], device=local)
    lib = ctx.lib
    aux = [(1, 2), (2, 2)]           # iPEPS::init_basis rule for D=4, spin-1/2 (src/PEPS/iPEPS.cpp:82-116)
    a, phys = ctx.basis(aux), ctx.basis([(2, 1)])
    A = ctx.tensor([a, a], [a, a, phys]).setRandom(rank)
    ctm = xb.CTM(A, CHI)

    def step_timed():
        ctx.timer_start()
        ctm.grow_all()
        return ctx.timer_stop()

    # workload state: CTM_INIT::FROM_A starts with chi = D^2 = 16 and grows; the benchmark configuration is chi = 64, i.e. the environment after its bases
    # have saturated (their sector structure no longer changes from step to step).  Reaching that state is set-up, like building the tensors; the W warm-up
    # steps and the K timed steps all run at chi = 64 with a fixed set of layouts.
    setup_steps, prev_struct, same = 0, None, 0
    while setup_steps < 12 and same < 2:
        ctm.grow_all()
        setup_steps += 1
        cur = tuple(ctm.get(nm).structure()[1:3] for nm in ("C1", "T1"))
        same = same + 1 if cur == prev_struct else 0
        prev_struct = cur
    for _ in range(args.warmup):
        ctm.grow_all()
    ctx.sync()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing
    l0 = ctx.stats()["launches"]
    barrier()
    total_ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        total_ms += step_timed()
    barrier()
    launches = ctx.stats()["launches"] - l0
    # ---- the same steps with per-category event pairs (live kernel durations for the roofline)
    ctx.profile(True)
    ctx.profile_read()
    prof_ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        prof_ms += step_timed()
    prof = ctx.profile_read()
    ctx.profile(False)

    # ---- end to end through the C ABI with host buffers
    host = {}
    def host_buf(name, n):
        if name not in host or host[name].numel() < n:
            host[name] = torch.empty(max(n, 1), dtype=torch.float64).pin_memory()
        return host[name]
    sizes = {}
    for nm in ENV + ["A"]:
        t = ctm.get(nm)
        sizes[nm] = t.plainSize()
        xb.capi.check(lib.xb_tensor_download(t.h, C.cast(host_buf(nm, sizes[nm]).data_ptr(), xb.capi.DP), sizes[nm]))
    barrier()
    e2e_ms, h2d, d2h = 0.0, 0, 0
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.timer_start()
        for nm in ENV + ["A"]:
            t = ctm.get(nm)          # handle with the current layout; its contents are overwritten from the host image
            xb.capi.check(lib.xb_tensor_upload(t.h, C.cast(host[nm].data_ptr(), xb.capi.DP), sizes[nm]))
            ctm.set(nm, t)
            h2d += 8 * sizes[nm]
        ctm.grow_all()
        for nm in ENV:
            t = ctm.get(nm)
            sizes[nm] = t.plainSize()
            xb.capi.check(lib.xb_tensor_download(t.h, C.cast(host_buf(nm, sizes[nm]).data_ptr(), xb.capi.DP), sizes[nm]))
            d2h += 8 * sizes[nm]
        e2e_ms += ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- bond energy of the SU(2) Heisenberg model on the converged-so-far environment: computeRDM_h,
    # computeRDM_v and avg() on the device; S.S is -3/4 on the singlet and +1/4 on the triplet sector of phys x phys
    energy = None
    if rank == 0:
        try:
            H = ctx.tensor([phys, phys], [phys, phys])
            secs, rows, cols, offs = H.structure()
            data = np.zeros(H.plainSize())
            for q, o in zip(secs, offs):
                stot = (q[0] - 1) / 2.0
                data[o] = 0.5 * (stot * (stot + 1.0) - 1.5)
            H.upload(data)
            for _ in range(2):   # plan build; the second use of a plan builds its table-driven permute form (host work)
                ctm.computeRDM_h()
                ctm.computeRDM_v()
            ctx.sync()
            ctx.timer_start()
            rho_h, rho_v = ctm.computeRDM_h(), ctm.computeRDM_v()
            e_h, e_v = xb.avg_two_site(rho_h, H), xb.avg_two_site(rho_v, H)
            energy = {"ms": ctx.timer_stop(), "E_h": e_h, "E_v": e_v, "energy_per_site": e_h + e_v,
                      "what": "computeRDM_h + computeRDM_v + avg(S.S) after the timed CTM steps (random site tensor: not a variational energy)"}
        except Exception as e:
            energy = {"error": f"{type(e).__name__}: {e}"}
    if world > 1:
        tt = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = tt.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())

    sharded = None
    if world > 1 and (not args.no_extras or os.environ.get("XB_BENCH_SHARDED")):
        try:
            sharded = sharded_section(xb, torch, dist, local, world)
        except Exception as e:   # the headline line must survive a failure of the optional section
            sharded = {"error": f"{type(e).__name__}: {e}"}
    extras, dgemm_peak = ({}, None)
    if rank == 0 and not args.no_extras:
        ksampler = ClockSampler(local)
        ksampler.start()
        extras, dgemm_peak = kernel_sections(xb, local, torch, cpu_d6=(world == 1 and args.cpu_steps > 0))
        extras["clocks"] = ksampler.stop()

    if rank == 0:
        ms_step = total_ms / args.steps
        value = world * args.steps / (total_ms * 1e-3)
        p_ms, p_cnt, p_bytes = prof["permute"]
        g_ms, g_cnt, g_fl = prof["gemm"]
        s_ms, s_cnt, _ = prof["tsvd"]
        shares = {k: (v[0] / prof_ms if prof_ms > 0 else None) for k, v in prof.items()}
        # roofline of the HBM-bound kernel on the tensor that is large enough to be bandwidth-bound (W4); the CTM-step
        # permutes (<= 11 MB each at D=4) are launch/latency-bound and are reported beside it.
        roofline = None
        # DRAM bytes per launch (dram__bytes_read + write) of the two roofline kernels from the committed ncu --set full captures
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        if extras:
            w4 = extras["permute_w4"]
            roofline = {"kernel": "permute_bulk_kernel", "bound": "hbm", "achieved": w4["GBps_aggregate"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": w4["GBps_aggregate"] / hbm_peak, "traffic": traffic.get("permute_bulk_kernel"), "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": 16 * 162238464,
                        "on": "W4 (25 permute<shift,p> cases of the 1.30 GB U(1) tensor), algorithmic bytes 8*(N_src+N_dst) per launch",
                        "in_ctm_step": {"launches_per_step": p_cnt / args.steps, "avg_us_per_launch": 1e3 * p_ms / max(p_cnt, 1),
                                        "GBps": p_bytes / max(p_ms, 1e-9) / 1e6, "share_of_step": shares["permute"]}}
            d6 = extras.get("ctm_d6_chi144", {})
            if "steps" in d6:
                b6 = min(d6["steps"], key=lambda r: r["ms"])
                roofline["in_ctm_step_d6_chi144"] = {"launches_per_step": b6["launch_groups"]["permute"], "ms": b6["device_ms"]["permute"], "GBps": b6["permute_GBps"],
                                                     "frac": b6["permute_GBps"] / hbm_peak, "share_of_step": b6["device_ms"]["permute"] / b6["ms"],
                                                     "note": "SU(2) plans: table-driven form (permute_lut_kernel), recoupling sums read every source element 2.1-2.3 times (DESIGN.md 9.2)"}
            roofline_contract = {"kernel": "gemm_grouped_kernel", "bound": "tensor", "achieved": extras["contract_w4"]["TFLOPs"], "peak": dgemm_peak,
                                 "unit": "TFLOP/s", "frac": extras["contract_w4"]["TFLOPs"] / dgemm_peak, "traffic": traffic.get("gemm_grouped_kernel"),
                                 "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                                 "on": "W4 T*T (1.27e12 flop)",
                                 "in_ctm_step": {"launches_per_step": g_cnt / args.steps, "avg_us_per_launch": 1e3 * g_ms / max(g_cnt, 1),
                                                 "TFLOPs": g_fl / max(g_ms, 1e-9) / 1e9, "share_of_step": shares["gemm"]}}
            if "steps" in d6:
                roofline_contract["in_ctm_step_d6_chi144"] = {"launch_groups_per_step": b6["launch_groups"]["gemm"], "ms": b6["device_ms"]["gemm"],
                                                              "TFLOPs": b6["gemm_TFLOPs"], "frac": b6["gemm_TFLOPs"] / dgemm_peak,
                                                              "share_of_step": b6["device_ms"]["gemm"] / b6["ms"]}
        else:
            roofline = {"kernel": "permute_kernel", "bound": "hbm", "achieved": p_bytes / max(p_ms, 1e-9) / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": p_bytes / max(p_ms, 1e-9) / 1e6 / hbm_peak, "traffic": None, "peak_source": peak_src,
                        "on": "permutes inside the CTM step (latency-bound sizes)"}
            roofline_contract = None
        cpu_sps, cpu_ms = time_oracle(3, args.cpu_steps) if (world == 1 and args.cpu_steps > 0) else (None, None)
        line = {"metric": "ctmrg_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (counter-based uniform[-1,1] site tensor, seed = rank)", "config": dict(config_dict(world), setup_steps_until_chi_saturated=setup_steps),
                "clocks": clocks,
                "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": h2d // args.steps,
                        "d2h_bytes_per_step": d2h // args.steps},
                "gpu_launches": int(launches),
                "roofline": roofline, "roofline_contract": roofline_contract,
                "step_breakdown": {"ms_per_step_with_event_pairs": prof_ms / args.steps,
                                   "per_step": {k: {"ms": v[0] / args.steps, "launch_groups": v[1] / args.steps, "share": shares[k]} for k, v in prof.items()}},
                "kernels": extras}
        if sharded is not None:
            line["sharded_w4"] = sharded
        if energy is not None:
            line["rdm_energy"] = energy
        if cpu_sps is not None:
            line["cpu_baseline"] = {"value": cpu_sps, "unit": "steps/s", "cores": blas_threads(), "kind": "port",
                                    "sample": f"{args.cpu_steps} grow_all() steps of the same D=4 chi=64 workload after 3 warm-up steps; " + CPU_KIND_NOTE,
                                    "ms_per_step": cpu_ms}
        print(json.dumps(line), file=real_stdout or sys.stdout, flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
