#!/bin/bash
# N-GPU call, lean: box facts and bench.py at N ranks (optionally the 2-process tests first).
#   gpurun --gpus 8 --timeout 1500 -- "bash tools/gpu_r2_n.sh r02v 8 [tests]"
TAG="${1:-run}"
N="${2:-8}"
O=gpurun_out
mkdir -p $O
step() { echo "== $1" >> $O/${TAG}_steps.log; shift; local t0=$SECONDS; "$@"; echo "   exit $? after $((SECONDS - t0)) s" >> $O/${TAG}_steps.log; }
{ nproc; free -g | head -2; nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -14; } > $O/${TAG}_box.txt 2>&1
if [ "$3" = "tests" ]; then
  step "pytest multi" timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > $O/${TAG}_pytest_multi.log 2>&1
  tail -3 $O/${TAG}_pytest_multi.log
fi
step "bench n$N" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
tail -5 $O/${TAG}_bench_n$N.err
cat $O/${TAG}_bench_n$N.json | cut -c1-600
cat $O/${TAG}_steps.log
