import sys, os
sys.path.insert(0, "/root/repo")
import torch, bench
import xped_b200 as xb
torch.cuda.set_device(0)
try:
    out, peak = bench.kernel_sections(xb, 0, torch)
    print("OK", out["tsvd_w3"]["ms"], out.get("ctm_d6_chi144", {}).get("ms_per_step"), out.get("mps_bench_cpp"))
except Exception as e:
    import traceback; traceback.print_exc()
