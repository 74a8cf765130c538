#!/bin/bash
# tools/ab/cycles_multi.sh <lib1.so> ...: SM cycles (ncu sm__cycles_elapsed.max) of front_tc_kernel on 100 forwards for each build
LIB=superpoint-nerf-pytorch_b200/libspn_b200.so
cp $LIB /tmp/cur.so
for L in "$@"; do
  n=$(basename $L .so); cp $L $LIB
  SPN_PROBE_BASELINE_ONLY=1 timeout 120 ncu --metrics sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:front_tc_kernel --csv --log-file gpurun_out/cyc_$n.csv python tools/front_probe.py > /dev/null 2>&1
  python - $n <<'P'
import csv,sys
n=sys.argv[1]
rows=[r for r in csv.reader(open(f'gpurun_out/cyc_{n}.csv')) if len(r)>5]
h=rows[0]; im=h.index('Metric Name'); iv=h.index('Metric Value'); ii=h.index('ID')
out={}
for r in rows[1:]: out.setdefault(r[ii],{})[r[im]]=float(r[iv].replace(',',''))
ids=sorted(out,key=int)[2:]
for k in out[ids[0]]:
    vals=[out[i][k] for i in ids]
    print(n, k, round(sum(vals)/len(vals),2), '/tile', round(sum(vals)/len(vals)/ (60000 if 'sum' in k else 405.4),2))
P
done
cp /tmp/cur.so $LIB
