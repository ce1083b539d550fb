#!/bin/bash
# Round-2 GPU call 15 (2 GPUs): world-2 suite and N=2 bench (parity leg) on the final kernels.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c15_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c15_$name.log" | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TMO=500 run gpu_multi python -m pytest tests/test_gpu_multi.py -m gpu -q
TMO=300 run bench_w2 $TR --master-port 29681 bench.py --gpus 2 --no-cpu-baseline
