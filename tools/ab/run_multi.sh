#!/bin/bash
# tools/ab/run_multi.sh <tag> <lib1.so> <lib2.so> ...: interleaved bench runs (2 rounds) of several builds of libspn_b200.so
set -u
TAG=$1; shift
LIB=superpoint-nerf-pytorch_b200/libspn_b200.so
cp $LIB /tmp/cur.so
for rep in 1 2; do
  for L in "$@"; do
    n=$(basename $L .so)
    cp $L $LIB
    timeout 150 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/ab_${TAG}_${n}_$rep.json 2> gpurun_out/ab_${TAG}_${n}_$rep.err
  done
done
cp /tmp/cur.so $LIB
python - "$TAG" <<'P'
import json,glob,sys
for f in sorted(glob.glob(f'gpurun_out/ab_{sys.argv[1]}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k={x['kernel']:x for x in d.get('kernels',[])}
        print(f.split('/')[-1], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'], 'front', round(k['backbone.block_2']['ms_per_launch'],4), 'kp', d.get('keypoints_last_step'))
    except Exception as e: print(f,'ERR',e)
P
