#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
nproc; lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket" 
for pin in 0 1; do
  if [ $pin = 0 ]; then export XB_BENCH_NO_PIN=1; else unset XB_BENCH_NO_PIN; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 5 --no-extras > gpurun_out/rep_n${N}_$pin.json 2> gpurun_out/rep_n${N}_$pin.err
  python -c "
import json
l=[x for x in open('gpurun_out/rep_n${N}_$pin.json') if x.startswith('{')]
d=json.loads(l[-1]); print('pin $pin', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
