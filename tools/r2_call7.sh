#!/bin/bash
# Round-2 GPU call 7 (2 GPUs): N=2 bench with the bf16-wire parity leg after the exclusive-detection fix; loss-kernel tests.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c7_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c7_$name.log" | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TMO=400 run bench_w2 $TR --master-port 29631 bench.py --gpus 2 --no-cpu-baseline
TMO=300 run loss_tests python -m pytest tests/test_gpu_loss.py tests/test_gpu_step.py -m gpu -q
TMO=300 run bench_ref_w2 $TR --master-port 29632 bench.py --gpus 2 --impl reference --steps 4 --warmup 1
