#!/bin/bash
mkdir -p gpurun_out
for tr in 0 1; do
  XB_SVD_TRUNC=$tr timeout 600 python tools/rsvd_check.py 4 64 8 > gpurun_out/rsvd_d4_t$tr.log 2>&1; echo "rc=$?" >> gpurun_out/rsvd_d4_t$tr.log
done
python - <<'PY'
import numpy as np
for D in (4,):
    a=np.load(f"gpurun_out/rsvd_sigma_D{D}_trunc0.npy"); b=np.load(f"gpurun_out/rsvd_sigma_D{D}_trunc1.npy")
    print("sigma shapes", a.shape, b.shape, "rel diff", np.linalg.norm(a-b)/np.linalg.norm(a) if a.shape==b.shape else None)
PY
XB_SVD_TRUNC=1 XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 4 > gpurun_out/rsvd_d4_prof.log 2>&1
grep "^step\|tSVD of" gpurun_out/rsvd_d4_t0.log; grep "^step\|tSVD of" gpurun_out/rsvd_d4_t1.log; grep "\[rsvd\]" gpurun_out/rsvd_d4_prof.log | tail -8
if [ "${D6:-1}" = "1" ]; then
XB_SVD_TRUNC=1 XB_SVD_PROF=1 timeout 900 python tools/rsvd_check.py 6 144 5 > gpurun_out/rsvd_d6_t1.log 2>&1; echo "rc=$?" >> gpurun_out/rsvd_d6_t1.log
grep "^step\|tSVD of\|rc=\|Error\|error" gpurun_out/rsvd_d6_t1.log | tail; grep "\[rsvd\]" gpurun_out/rsvd_d6_t1.log | tail -6
fi
timeout 900 python -m pytest tests -m gpu -x -q -k "tsvd or decompose or ctm or mps" > gpurun_out/pytest_svd.log 2>&1; tail -5 gpurun_out/pytest_svd.log
