#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded.log 2>&1; tail -5 gpurun_out/pytest_sharded.log
XB_BENCH_SHARDED=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-extras > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_n$N.json") if x.startswith("{")]
d=json.loads(l[-1]); print(d["value"], d["ms_per_step"]); print(json.dumps(d.get("sharded_w4"), indent=1)[:4000])
PY
tail -5 gpurun_out/bench_n$N.err
