#!/bin/bash
# A/B sweep of the selection kernel's ring feeders and geometry on one GPU (env knobs of csrc/selection.cu).
O=gpurun_out
mkdir -p $O
{
for which in fps32 kc32 fps256 kc256 fps64 fps128; do
  echo "== $which feed=cp.async";  SSDR_SEL_FEED=0 python tools/prof_select.py $which 1500 | tail -1
  for st in 2 3 4 6; do for w in 8 12 16; do
    echo "== $which feed=tma stages=$st warps=$w"; SSDR_SEL_STAGES=$st SSDR_SEL_WARPS=$w python tools/prof_select.py $which 1500 | tail -1
  done; done
done
} > $O/sweep_select.txt 2>&1
