#!/bin/bash
# A/B of the k=16 query kernels and the cell occupancy on the headline pyramid (run on the GPU box).
O=gpurun_out
T=${1:-ab}
for q in 0 1; do
  SSDR_KNN_QUERY=$q python bench.py --steps 10 --warmup 3 --no-extra --no-multi 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('query=$q value %.1f M ms %.4f e2e %.1f M main_ms %s evals/q %.1f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, [round(s['main_kernel_ms'],4) for s in d['knn_detail']['stage_ms'] if s['call']=='k16'], d['roofline']['dist_evals_per_query']))"
done > $O/${T}_ab.txt 2>&1
for occ in 0.2 0.4 0.5 0.65 0.8; do
  SSDR_KNN_OCCUPANCY=$occ python bench.py --steps 10 --warmup 3 --no-extra --no-multi 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('occ=$occ value %.1f M ms %.4f main_ms %s evals/q %.1f' % (d['value']/1e6, d['ms_per_step'], [round(s['main_kernel_ms'],4) for s in d['knn_detail']['stage_ms'] if s['call']=='k16'], d['roofline']['dist_evals_per_query']))"
done >> $O/${T}_ab.txt 2>&1
cat $O/${T}_ab.txt
