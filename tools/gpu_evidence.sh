#!/bin/bash
# One gpurun call that produces the evidence set copied to profiles/ by tools/collect_profiles.py: parity tests, bench
# + reference arm, CIGAR stage benches, 3.1 Gbp stage table, sanitizer, ncu launch list and --set full captures.
#   gpurun --timeout 1000 -- "bash tools/gpu_evidence.sh r02a"
TAG="${1:-run}"   # tag of the output files under gpurun_out/
O=gpurun_out
mkdir -p $O
step() { echo "== $1" >> $O/${TAG}_steps.log; shift; "$@"; echo "   exit $?" >> $O/${TAG}_steps.log; }
step "pytest gpu" timeout 600 python -m pytest tests -q -m gpu > $O/${TAG}_pytest.log 2>&1
tail -5 $O/${TAG}_pytest.log
step "bench graph" timeout 400 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
step "bench reference arm" timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err
step "cigar ont" timeout 300 python tools/cigar_bench.py --records 60000 --mean-ops 2400 > $O/${TAG}_cigar_ont.json 2> $O/${TAG}_cigar_ont.err
step "cigar hifi" timeout 300 python tools/cigar_bench.py --records 3000000 --mean-ops 31 --sigma 0.3 > $O/${TAG}_cigar_hifi.json 2> $O/${TAG}_cigar_hifi.err
step "scale 3.1Gbp" timeout 400 python tools/scale_check.py --gbp 3.1 --contigs 24 --coverage 30 --second-aligner > $O/${TAG}_scale_3gbp_hifi_2files.json 2> $O/${TAG}_scale3.err
step "memcheck" timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $O/${TAG}_memcheck.log 2>&1
step "racecheck" timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > $O/${TAG}_racecheck.log 2>&1
step "ncu launches" timeout 400 env GCI_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
step "ncu full bench" timeout 500 env GCI_GRAPH=0 ncu --set full --clock-control none --import-source on \
    -k regex:'depth_tile_kernel|cigar_stats_kernel|runs_kernel|join_kernel|bucket_fill_kernel|tile_apply_kernel|gate_span' -s 60 -c 14 \
    -o $O/${TAG}_prof_bench -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
step "ncu full cigar stream" timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:'cigar_stream_kernel' -s 2 -c 2 -o $O/${TAG}_prof_cigar_stream -f \
    python tools/cigar_bench.py --records 30000 --mean-ops 2400 --steps 2 > $O/${TAG}_ncu_cigar.log 2>&1
step "ncu full cigar staged" timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:'cigar_stats_kernel' -s 2 -c 2 -o $O/${TAG}_prof_cigar_staged -f \
    python tools/cigar_bench.py --records 1000000 --mean-ops 31 --sigma 0.3 --steps 2 > $O/${TAG}_ncu_cigar2.log 2>&1
cat $O/${TAG}_steps.log
