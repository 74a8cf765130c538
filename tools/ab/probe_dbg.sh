#!/bin/bash
# knock-out probe (SM cycles per tile, tensor-pipe active %) with the diagnostic library tools/ab/libspn_b200_dbg.so
cp superpoint-nerf-pytorch_b200/libspn_b200.so /tmp/new.so; cp tools/ab/libspn_b200_dbg.so superpoint-nerf-pytorch_b200/libspn_b200.so
timeout 300 bash tools/front_probe_cycles.sh | python -c "
import sys,collections
d=collections.OrderedDict()
for l in sys.stdin:
    p=l.split()
    k=l[l.index('front_tc_kernel'):].split('>')[0]
    d.setdefault(k,[]).append((float(p[-2]),float(p[-1])))
for k,v in d.items(): print(k, round(sum(x[0] for x in v)/len(v)/405.4,1), round(sum(x[1] for x in v)/len(v),1))
"
cp /tmp/new.so superpoint-nerf-pytorch_b200/libspn_b200.so
