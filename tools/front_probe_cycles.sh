#!/bin/bash
# Role knock-outs of front_tc_kernel measured in SM CYCLES (ncu sm__cycles_elapsed.max), not in time: under the power cap
# the clock moves by 15 % between variants, which made the event-timed version of this probe (front_probe.py) unreadable.
# Needs the diagnostic library (SPN_FRONT_DBG_BUILD=1 python superpoint-nerf-pytorch_b200/build.py --force).
ncu --metrics sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none \
    -k regex:front_tc_kernel --csv --log-file gpurun_out/front_probe_cycles.csv python tools/front_probe.py > gpurun_out/front_probe_cycles.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/front_probe_cycles.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
out={}
for r in rows[1:]:
    out.setdefault((r[h.index('ID')], r[ik][:40]), {})[r[im]]=r[iv]
last=None
for (i,k),m in out.items():
    print(i,k, m.get('sm__cycles_elapsed.max'), m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'))
P
