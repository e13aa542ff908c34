#!/bin/bash
# Run on the GPU box (gpurun -- 'bash tools/capture_profiles.sh rNN'): bench lines, ncu launch lists and one
# `ncu --set full` capture per hot kernel, all into gpurun_out/.  Summaries for profiles/ are made afterwards on the
# CPU box with tools/ncu_summary.py (see profiles/README.md).
R=${1:-r02}
O=gpurun_out
mkdir -p $O
NCU_FULL="ncu --set full --clock-control none --import-source on -f"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
python bench.py --steps 10 --warmup 3 > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2> /dev/null
# launch lists (the summaries skip nothing: compare shares)
$LIST --log-file $O/${R}_launches_pyramid.csv python tools/prof_knn.py 3 fused > /dev/null 2>&1
$LIST --log-file $O/${R}_launches_grid.csv python tools/prof_select.py grid > /dev/null 2>&1
$LIST --log-file $O/${R}_launches_fps32.csv python tools/prof_select.py fps32 50 > /dev/null 2>&1
# full captures.  query: level 0, k=16, 245760 queries = the first query16_kernel launch of the 3rd pyramid (3 per pyramid:
# levels 0-2; the two lowest levels run the brute-force kernel)
$NCU_FULL -k regex:^query16_kernel -s 6 -c 1 -o $O/${R}_ncu_knn_query python tools/prof_knn.py 3 calls > /dev/null 2>&1
$NCU_FULL -k regex:^exact_query_kernel -s 0 -c 1 -o $O/${R}_ncu_knn_exact python tools/prof_knn.py 1 calls > /dev/null 2>&1
$NCU_FULL -k regex:^build_kernel -s 2 -c 1 -o $O/${R}_ncu_knn_build python tools/prof_tree.py 6 40960 > /dev/null 2>&1
$NCU_FULL -k regex:select32_kernel -s 1 -c 1 -o $O/${R}_ncu_fps32 python tools/prof_select.py fps32 30 > /dev/null 2>&1
$NCU_FULL -k regex:select32_kernel -s 1 -c 1 -o $O/${R}_ncu_kc32 python tools/prof_select.py kc32 30 > /dev/null 2>&1
$NCU_FULL -k regex:select_kernel -s 1 -c 1 -o $O/${R}_ncu_fps256 python tools/prof_select.py fps256 30 > /dev/null 2>&1
$NCU_FULL -k regex:sort_kernel -s 2 -c 1 -o $O/${R}_ncu_grid_sort python tools/prof_select.py grid > /dev/null 2>&1
$NCU_FULL -k regex:reduce_small_kernel -s 2 -c 1 -o $O/${R}_ncu_grid_reduce python tools/prof_select.py grid > /dev/null 2>&1
$NCU_FULL -k regex:pack_rows_kernel -s 2 -c 1 -o $O/${R}_ncu_grid_pack python tools/prof_select.py grid > /dev/null 2>&1
$NCU_FULL -k regex:reduce_rec_kernel -s 2 -c 1 -o $O/${R}_ncu_grid_reduce_groups python tools/prof_select.py grid > /dev/null 2>&1
# the same two kernels on the 80 M-point scan of config 3 (third call)
$LIST --log-file $O/${R}_launches_grid_80m.csv python tools/prof_grid_scan.py 80000000 2 > /dev/null 2>&1
$NCU_FULL -k regex:sort_kernel -s 2 -c 1 -o $O/${R}_ncu_grid_sort_80m python tools/prof_grid_scan.py 80000000 3 > /dev/null 2>&1
$NCU_FULL -k regex:reduce_small_kernel -s 2 -c 1 -o $O/${R}_ncu_grid_reduce_80m python tools/prof_grid_scan.py 80000000 3 > /dev/null 2>&1
# text summaries next to the reports; gpurun copies back at most 64 MiB, so only a few reports travel
for f in $O/${R}_ncu_*.ncu-rep; do python tools/ncu_summary.py rep $f > ${f%.ncu-rep}.txt 2>&1; done
for f in $O/${R}_launches_*.csv; do python tools/ncu_summary.py launches $f > ${f%.csv}.txt 2>&1; done
python tools/ncu_summary.py traffic $O/${R}_ncu_knn_query.ncu-rep knn_query_kernel_level0 $O/${R}_traffic.json 245760
python tools/ncu_summary.py traffic $O/${R}_ncu_fps32.ncu-rep fps_d32 $O/${R}_traffic.json 29
python tools/ncu_summary.py traffic $O/${R}_ncu_fps256.ncu-rep fps_d256 $O/${R}_traffic.json 29
python tools/ncu_summary.py traffic $O/${R}_ncu_kc32.ncu-rep kcenter_d32 $O/${R}_traffic.json 45
python tools/ncu_summary.py traffic $O/${R}_ncu_grid_sort.ncu-rep grid_sort_1m $O/${R}_traffic.json 1000000
python tools/ncu_summary.py traffic $O/${R}_ncu_grid_reduce.ncu-rep grid_reduce_1m $O/${R}_traffic.json 1000000
python tools/ncu_summary.py traffic $O/${R}_ncu_grid_pack.ncu-rep grid_pack_1m $O/${R}_traffic.json 1000000
python tools/ncu_summary.py traffic $O/${R}_ncu_grid_reduce_groups.ncu-rep grid_reduce_groups_1m $O/${R}_traffic.json 1000000
python tools/ncu_summary.py traffic $O/${R}_ncu_grid_sort_80m.ncu-rep grid_sort_80m $O/${R}_traffic.json 80000000
python tools/ncu_summary.py traffic $O/${R}_ncu_grid_reduce_80m.ncu-rep grid_reduce_small_80m $O/${R}_traffic.json 80000000
rm -f $O/${R}_ncu_fps256.ncu-rep $O/${R}_ncu_kc32.ncu-rep $O/${R}_ncu_knn_build.ncu-rep $O/${R}_ncu_grid_reduce.ncu-rep $O/${R}_ncu_knn_exact.ncu-rep $O/${R}_ncu_grid_reduce_80m.ncu-rep $O/${R}_ncu_grid_sort_80m.ncu-rep $O/${R}_ncu_grid_pack.ncu-rep $O/${R}_ncu_grid_reduce_groups.ncu-rep
python tools/prof_pyramid_host.py > $O/${R}_pyramid_host_call.txt 2>&1
python tools/prof_e2e.py > $O/${R}_e2e_per_call.txt 2>&1
python tools/prof_grid.py 80000000 > $O/${R}_grid_phases.txt 2>&1
python tools/prof_cfg3_knn.py > $O/${R}_cfg3_knn_stages.txt 2>&1
for n in 160 640 2560 10240 40960; do python tools/prof_tree.py 6 $n; done > $O/${R}_tree_phases.txt 2>&1
ls -la $O | tail -30
