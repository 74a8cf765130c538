#!/bin/bash
# bench sweeps over the launch configuration (same library): chunk size (--max-forwards) and streams
for cfg in "100 1" "200 1" "400 1" "800 1" "1600 1" "100 1"; do
  set -- $cfg
  timeout 150 python bench.py --no-cpu-baseline --steps 10 --max-forwards $1 --streams $2 > gpurun_out/sweep_mf$1_s$2.json 2> gpurun_out/sweep_mf$1_s$2.err
  python - $1 $2 <<'P'
import json,sys
mf,s=sys.argv[1:3]
try:
    d=json.loads(open(f'gpurun_out/sweep_mf{mf}_s{s}.json').read().strip().splitlines()[-1])
    print('max_forwards',mf,'streams',s,'img/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'])
except Exception as e: print(mf,s,'ERR',e)
P
done
