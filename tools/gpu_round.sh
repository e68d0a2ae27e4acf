#!/bin/bash
# One gpurun call: parity tests, bench (graph replay and eager), stage-level scale checks, ncu evidence.
# Every step has its own timeout and writes under gpurun_out/; a failing step does not stop the next one.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01g'
TAG="${1:-run}"
O=gpurun_out
mkdir -p $O
step() { echo "== $1" | tee -a $O/${TAG}_steps.log; shift; "$@"; echo "   exit $?" | tee -a $O/${TAG}_steps.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.csv 2>&1
step "pytest gpu" timeout 600 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest.log 2>&1
tail -5 $O/${TAG}_pytest.log
step "bench graph" timeout 400 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
step "bench eager" timeout 300 env GCI_GRAPH=0 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_eager.json 2> $O/${TAG}_bench_eager.err
step "bench reference arm" timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err
step "cigar ont" timeout 300 python tools/cigar_bench.py --records 60000 --mean-ops 2400 > $O/${TAG}_cigar_ont.json 2> $O/${TAG}_cigar_ont.err
step "cigar hifi" timeout 300 python tools/cigar_bench.py --records 3000000 --mean-ops 31 --sigma 0.3 > $O/${TAG}_cigar_hifi.json 2> $O/${TAG}_cigar_hifi.err
step "scale 3.1Gbp" timeout 400 python tools/scale_check.py --gbp 3.1 --contigs 24 --coverage 30 --second-aligner > $O/${TAG}_scale_3gbp_hifi_2files.json 2> $O/${TAG}_scale3.err
# ncu: launch list of a short eager bench (per-launch device times, cold cache, serialised)
step "ncu launches" timeout 400 env GCI_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
# ncu --set full: dominant kernel + the reworked ones
step "ncu full bench" timeout 500 env GCI_GRAPH=0 ncu --set full --clock-control none --import-source on \
    -k regex:'depth_tile_kernel|cigar_stats_kernel|runs_kernel|join_kernel|bucket_fill_kernel|tile_apply_kernel' -s 60 -c 12 \
    -o $O/${TAG}_prof_bench -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
step "ncu full cigar stream" timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:'cigar_stream_kernel' -s 2 -c 2 -o $O/${TAG}_prof_cigar_stream -f \
    python tools/cigar_bench.py --records 30000 --mean-ops 2400 --steps 2 > $O/${TAG}_ncu_cigar.log 2>&1
ls -la $O | tail -30
