#!/bin/bash
# Evidence call of a round on ONE GPU: tests, both bench arms, the ncu launch list of the step (the recipe's cold,
# serialised pass and a hot-cache one), one `ncu --set full` capture of the step and of the .depth.gz encoder.
#   gpurun --timeout 2400 -- "bash tools/gpu_final.sh r02z"
TAG="${1:-run}"
O=gpurun_out
mkdir -p $O
bash tools/gpu_r2.sh $TAG
step() { echo "== $1" >> $O/${TAG}_steps.log; shift; local t0=$SECONDS; "$@"; echo "   exit $? after $((SECONDS - t0)) s" >> $O/${TAG}_steps.log; }
step "ncu launches hot" timeout 400 env GCI_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv \
    --log-file $O/${TAG}_launches_hot.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_launches_hot.log 2>&1
bash tools/gpu_prof.sh $TAG
