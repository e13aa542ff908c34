#!/bin/bash
# A/B of the pyramid call: branches on/off, graph on/off (run on the GPU box).
O=gpurun_out
T=${1:-ab}
: > $O/${T}_pyr.txt
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  SSDR_KNN_BRANCHES=$1 SSDR_KNN_GRAPH=$2 python bench.py --steps 20 --warmup 3 --no-extra --no-multi 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('branches=$1 graph=$2 value %.1f M ms %.4f e2e %.1f M equal=%s launches=%s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d.get('pyramid_call_equals_per_call_results'), d.get('gpu_launches')))" >> $O/${T}_pyr.txt 2>&1
done
cat $O/${T}_pyr.txt
for sp in 1 2 3; do
  SSDR_KNN_SPLIT=$sp python bench.py --steps 20 --warmup 3 --no-extra --no-multi 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('split=$sp value %.1f M ms %.4f e2e %.1f M one_call %.1f M equal=%s launches=%s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['one_call']['value']/1e6, d.get('pyramid_call_equals_per_call_results'), d.get('gpu_launches')))" >> $O/${T}_pyr.txt 2>&1
done
cat $O/${T}_pyr.txt
