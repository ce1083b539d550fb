#!/bin/bash
# Round-2 GPU call 6 (2 GPUs): world-2 suite on ABI v7 (bf16 gradients on the wire), 2-GPU train.py CLI, N=2 bench with parity leg
# and ride-along extras (multi-scale, sweep, SyncBN exchange cost, torch DDP + SyncBatchNorm arm).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c6_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c6_$name.log" | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TMO=700 run gpu_multi python -m pytest tests/test_gpu_multi.py tests/test_gpu_train_cli.py -m gpu -q
TMO=400 run bench_w2 $TR --master-port 29621 bench.py --gpus 2 --no-cpu-baseline
