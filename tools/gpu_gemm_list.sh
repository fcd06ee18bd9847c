#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/rsvd_check.py 4 64 12 2>&1 | grep "^step 1[01]" | cut -c1-150
timeout 250 python tools/rsvd_check.py 6 144 8 2>&1 | grep "^step [67]" | cut -c1-150
XB_CHECK_NCU_LAST=1 XB_SVD_NOGRAPH=1 timeout 600 ncu --profile-from-start off -k regex:gemm_grouped --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gemm_last_d4.csv python tools/rsvd_check.py 4 64 10 > gpurun_out/ncu_gemm_last_d4.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/gemm_last_d4.csv') if not l.startswith('==')]
rd=csv.DictReader(lines)
agg=collections.OrderedDict()
seq=[]
for r in rd:
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(r['Metric Value'].replace(',','')); unit=r['Metric Unit']
    us=v/1e3 if unit.startswith('ns') else v
    key=(('128' if '<(int)128>' in r['Kernel Name'] or '<128>' in r['Kernel Name'] else '64'), r['Grid Size'])
    agg.setdefault(key,[0,0.0]); agg[key][0]+=1; agg[key][1]+=us
    seq.append((key,us))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(k, n, 'launches', round(t,1), 'us total', round(t/n,1), 'us mean')
PY
