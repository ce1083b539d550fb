#!/bin/bash
# Round-2 GPU call 12 (8 GPUs): all-reduce variants at world 8 after the auto-selection change, and the final N=8 bench line.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c12_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c12_$name.log" | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
SOD_TEST_WORLD=8 TMO=200 run gpu_multi_w8 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "allreduce_variants"
TMO=300 run bench_w8 $TR --master-port 29671 bench.py --gpus 8 --no-cpu-baseline
