#!/bin/bash
# Multi-GPU call: whole GPU test suite (incl. the multi-process tests), depth-kernel A/B, bench at N = 1 and N = GPUs.
#   gpurun --gpus 2 --timeout 1500 -- "bash tools/gpu_r2_multi.sh r02e 2"
TAG="${1:-run}"
N="${2:-2}"
O=gpurun_out
mkdir -p $O
step() { echo "== $1" >> $O/${TAG}_steps.log; shift; local t0=$SECONDS; "$@"; echo "   exit $? after $((SECONDS - t0)) s" >> $O/${TAG}_steps.log; }
{ nproc; free -g | head -2; nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -14; } > $O/${TAG}_box.txt 2>&1
step "pytest gpu" timeout 1200 python -m pytest tests -q -m gpu --durations=6 > $O/${TAG}_pytest.log 2>&1
tail -12 $O/${TAG}_pytest.log
step "bench n1 tma" timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
step "bench n1 stg" timeout 600 env GCI_DEPTH_TMA=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_n1_stg.json 2> $O/${TAG}_bench_n1_stg.err
step "bench n$N" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
tail -5 $O/${TAG}_bench_n$N.err
cat $O/${TAG}_steps.log
