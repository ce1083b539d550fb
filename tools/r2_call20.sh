#!/bin/bash
# Round-2 GPU call 20 (2 GPUs): the final bench.py (incl. the input-pipeline extra) at world 2.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29691 bench.py --gpus 2 --no-cpu-baseline ) > $O/c20_bench_w2.log 2>&1; echo "exit $?"; tail -c 300 $O/c20_bench_w2.log
