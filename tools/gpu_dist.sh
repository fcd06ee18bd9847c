#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py ${DIM:-24} > gpurun_out/dist_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_n$N.log
grep -v "NCCL\|^$\|warn" gpurun_out/dist_check_n$N.log | tail -40
