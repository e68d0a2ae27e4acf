#!/bin/bash
# Round-2 evidence call: box facts, parity tests, bench (both arms), ncu launch list of the configs[2] step.
#   gpurun --timeout 1500 -- "bash tools/gpu_r2.sh r02a"
TAG="${1:-run}"
O=gpurun_out
mkdir -p $O
step() { echo "== $1" >> $O/${TAG}_steps.log; shift; local t0=$SECONDS; "$@"; echo "   exit $? after $((SECONDS - t0)) s" >> $O/${TAG}_steps.log; }
{ nproc; free -g | head -2; nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -12; } > $O/${TAG}_box.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
step "pytest gpu" timeout 900 python -m pytest tests -q -m gpu --durations=8 > $O/${TAG}_pytest.log 2>&1
tail -15 $O/${TAG}_pytest.log
fi
step "bench" timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -3 $O/${TAG}_bench.err
if [ -z "$SKIP_REF" ]; then
step "bench reference arm" timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err
fi
if [ -z "$SKIP_NCU" ]; then
step "ncu launches" timeout 500 env GCI_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_launches.log 2>&1
fi
cat $O/${TAG}_steps.log
