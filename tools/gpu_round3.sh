#!/bin/bash
# A/B of the depth kernel's constant-run shortcut + parity tests.
TAG="${1:-run}"
O=gpurun_out
mkdir -p $O
step() { echo "== $1" >> $O/${TAG}_steps.log; shift; "$@"; echo "   exit $?" >> $O/${TAG}_steps.log; }
step "pytest gpu" timeout 600 python -m pytest tests -q -m gpu > $O/${TAG}_pytest.log 2>&1
tail -3 $O/${TAG}_pytest.log
step "bench flat" timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_flat1.json 2> $O/${TAG}_bench_flat1.err
step "bench noflat" timeout 400 env GCI_DEPTH_FLAT=0 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_flat0.json 2> $O/${TAG}_bench_flat0.err
step "bench flat again" timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_flat1b.json 2> $O/${TAG}_bench_flat1b.err
step "scale 1Gbp flat" timeout 400 python tools/scale_check.py --gbp 1 --contigs 12 --coverage 30 > $O/${TAG}_scale_1gbp_flat1.json 2> $O/${TAG}_scale1.err
step "scale 1Gbp noflat" timeout 400 env GCI_DEPTH_FLAT=0 python tools/scale_check.py --gbp 1 --contigs 12 --coverage 30 > $O/${TAG}_scale_1gbp_flat0.json 2> $O/${TAG}_scale0.err
step "ncu depth flat" timeout 300 env GCI_GRAPH=0 ncu --set full --clock-control none --import-source on \
    -k regex:'depth_tile_kernel' -s 4 -c 2 -o $O/${TAG}_prof_depth_flat1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_depth.log 2>&1
cat $O/${TAG}_steps.log
