#!/bin/bash
# A/B of two builds of libspn_b200.so on one box: tools/ab/run_ab.sh <old.so> <tag>
# interleaved bench runs (new, old, new, old) so that clock drift hits both arms alike.
set -u
OLD=$1; TAG=$2
LIB=superpoint-nerf-pytorch_b200/libspn_b200.so
cp $LIB /tmp/new.so
for rep in 1 2; do
  cp /tmp/new.so $LIB; python bench.py --no-cpu-baseline --steps 10 > gpurun_out/ab_${TAG}_new$rep.json 2> gpurun_out/ab_${TAG}_new$rep.err
  cp $OLD $LIB;        python bench.py --no-cpu-baseline --steps 10 > gpurun_out/ab_${TAG}_old$rep.json 2> gpurun_out/ab_${TAG}_old$rep.err
done
cp /tmp/new.so $LIB
python - <<'P'
import json,glob,sys
for f in sorted(glob.glob('gpurun_out/ab_*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k={x['kernel']:x for x in d.get('kernels',[])}
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'], {n:round(v.get('ms_per_launch',0),4) for n,v in k.items()} if k else '')
    except Exception as e: print(f,'ERR',e)
P
