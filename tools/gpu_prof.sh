#!/bin/bash
# ncu --set full captures of one configs[2] step (eager launches) and of the .depth.gz encoder kernels.
#   gpurun --timeout 1200 -- "bash tools/gpu_prof.sh r02d"
TAG="${1:-run}"
O=gpurun_out
mkdir -p $O
step() { echo "== $1" >> $O/${TAG}_steps.log; shift; local t0=$SECONDS; "$@"; echo "   exit $? after $((SECONDS - t0)) s" >> $O/${TAG}_steps.log; }
STEP_K='paf_mark_kernel|paf_fill_kernel|paf_elect|cigar_lane_kernel|join_kernel|bucket_fill_kernel|depth_tile_tma_kernel|runs_count_kernel|runs_scan_kernel|runs_unstage_kernel|runs_write_kernel|tile_reduce_kernel|tile_apply_kernel'
step "ncu full step" timeout 700 env GCI_GRAPH=0 ncu --set full --clock-control none --import-source on \
    -k regex:"$STEP_K" -s 42 -c 14 -o $O/${TAG}_prof_step -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_step.log 2>&1
step "ncu full gz" timeout 700 env GCI_GRAPH=0 ncu --set full --clock-control none --import-source on \
    -k regex:'gz_' -s 5 -c 5 -o $O/${TAG}_prof_gz -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_gz.log 2>&1
ls -la $O | grep ${TAG}
cat $O/${TAG}_steps.log
