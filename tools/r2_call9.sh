#!/bin/bash
# Round-2 GPU call 9 (4 GPUs): the N=4 bench line (parity leg + extras) and the reference arm's 4-rank gloo run.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c9_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c9_$name.log" | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
TMO=300 run bench_w4 $TR --master-port 29651 bench.py --gpus 4 --no-cpu-baseline
