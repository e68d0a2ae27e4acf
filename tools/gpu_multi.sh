#!/bin/bash
# N-GPU call: the bench through torchrun with the peer-memory row exchange and with ncclAllGather (GCI_P2P=0),
# the single-GPU line of the same box, the two-rank GPU test.   gpurun --gpus 2 --timeout 600 -- "bash tools/gpu_multi.sh r02b 2"
TAG="${1:-run}"; N="${2:-2}"
O=gpurun_out
mkdir -p $O
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
echo "== bench graph x$N" >> $O/${TAG}_steps.log
run bench.py --gpus $N --steps 100 --warmup 5 > $O/${TAG}_bench_x$N.json 2> $O/${TAG}_bench_x$N.err; echo "   exit $?" >> $O/${TAG}_steps.log
echo "== bench nccl x$N" >> $O/${TAG}_steps.log
GCI_P2P=0 run bench.py --gpus $N --steps 100 --warmup 5 > $O/${TAG}_bench_nccl_x$N.json 2> $O/${TAG}_bench_nccl_x$N.err; echo "   exit $?" >> $O/${TAG}_steps.log
echo "== bench graph x1" >> $O/${TAG}_steps.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_x1.json 2> $O/${TAG}_bench_x1.err; echo "   exit $?" >> $O/${TAG}_steps.log
echo "== pytest two ranks" >> $O/${TAG}_steps.log
timeout 300 python -m pytest tests/test_gpu_random.py -q -m gpu -k "two_ranks" > $O/${TAG}_pytest2.log 2>&1; echo "   exit $?" >> $O/${TAG}_steps.log
cat $O/${TAG}_steps.log; tail -3 $O/${TAG}_pytest2.log; tail -c 600 $O/${TAG}_bench_x$N.err
