#!/bin/bash
# Round-2 GPU call 3 (2 GPUs): changed single-GPU tests, the world-2 suite, the N=2 bench with its parity leg, the stock-torch
# DDP+SyncBatchNorm arm, the all-reduce sweep, compute-sanitizer passes.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c3_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c3_$name.log" | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TMO=420 run gpu_tests_1 python -m pytest tests/test_gpu_syncbn.py tests/test_gpu_step.py tests/test_gpu_pipeline.py tests/test_gpu_sgd.py tests/test_gpu_train_cli.py -m gpu -q
TMO=600 run gpu_multi python -m pytest tests/test_gpu_multi.py -m gpu -q
TMO=240 run bench_w2 $TR --master-port 29611 bench.py --gpus 2 --no-cpu-baseline
TMO=240 run bench_w2_torch $TR --master-port 29612 bench.py --gpus 2 --impl torch
TMO=200 run sweep_w2 $TR --master-port 29613 bench.py --gpus 2 --sweep
TMO=200 run san_racecheck compute-sanitizer --tool racecheck python tools/sanitize_small.py
TMO=150 run san_synccheck compute-sanitizer --tool synccheck python tools/sanitize_small.py
TMO=240 run san_memcheck_w2 compute-sanitizer --tool memcheck --target-processes all $TR --master-port 29614 tools/sanitize_small.py
