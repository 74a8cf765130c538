#!/bin/bash
# End-of-round measurement suite on one B200 box (every step under its own timeout; outputs in gpurun_out/).
#   gpurun --timeout 1500 -- 'bash tools/final_measure.sh'
set -u
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q > $O/gputests_final.log 2>&1; tail -3 $O/gputests_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; tail -3 $O/smoke_final.log
timeout 300 python bench.py > $O/bench_r2_final_1gpu.json 2> $O/bench_r2_final_1gpu.err; tail -c 300 $O/bench_r2_final_1gpu.err
timeout 300 python bench.py --steps 100 --no-cpu-baseline > $O/bench_r2_sustained_100steps.json 2>/dev/null
timeout 400 python bench.py --impl reference --steps 2 > $O/bench_r2_reference_arm.json 2>/dev/null
timeout 300 python tools/bench_configs.py > $O/configs_r2_final.json 2>/dev/null
timeout 200 python bench.py --precision fp32 --images-per-step 4 --steps 3 --no-cpu-baseline > $O/bench_r2_fp32.json 2>/dev/null
timeout 200 python bench.py --precision f16x3 --images-per-step 32 --steps 10 --no-cpu-baseline > $O/bench_r2_f16x3.json 2>/dev/null
# ncu: launch list of one bench command, then --set full of one timed step of the same command (100 forwards per launch
# so that the per-launch figures compare with the earlier captures); the raw page is exported here (the .ncu-rep is > 64 MiB)
CMD="python bench.py --images-per-step 2 --steps 1 --warmup 3 --max-forwards 100 --no-cpu-baseline"
timeout 200 $CMD > $O/ncu_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv $CMD > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -s 81 -c 30 -o /tmp/r2_prof $CMD > $O/ncu2.log 2>&1
ncu -i /tmp/r2_prof.ncu-rep --page raw --csv > $O/r2_prof_raw.csv 2> $O/ncu2b.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"front_tc" -s 2 -c 1 -o $O/r2_prof_front_final $CMD > $O/ncu4.log 2>&1
ls -la $O/*.csv $O/*.ncu-rep | tail -5
for f in bench_r2_final_1gpu bench_r2_sustained_100steps bench_r2_f16x3 bench_r2_fp32; do python - $O/$f.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'clk', d['clocks'], 'roof', {k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline') or {}).items() if k in ('achieved','frac','ms_per_launch','share_of_step')})
except Exception as e: print(sys.argv[1],'ERR',e)
P
done
