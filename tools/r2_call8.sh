#!/bin/bash
# Round-2 GPU call 8 (8 GPUs): world-8 parity suite (subset) and the N=8 bench line with its parity leg and ride-along extras.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c8_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c8_$name.log" | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TMO=300 run bench_w8 $TR --master-port 29641 bench.py --gpus 8 --no-cpu-baseline
SOD_TEST_WORLD=8 TMO=420 run gpu_multi_w8 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "allreduce_sgd or syncbn_world2 or training_step or graph_replay"
TMO=200 run bench_ref_w8 $TR --master-port 29642 bench.py --gpus 8 --impl reference --steps 3 --warmup 1
