#!/bin/bash
# N-GPU sanity of the peer-memory row exchange + graph replay (short): bench only.
TAG="${1:-run}"; N="${2:-4}"
O=gpurun_out
mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 100 --warmup 5 > $O/${TAG}_bench_x$N.json 2> $O/${TAG}_bench_x$N.err
echo "exit $?"; tail -c 400 $O/${TAG}_bench_x$N.json
