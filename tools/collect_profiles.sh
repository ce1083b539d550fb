#!/bin/bash
# GPU-box script (1 GPU): everything that goes into profiles/ for this round. Outputs under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
# 1. the bench line (this repo's arm, the stock-torch arm, the CPU reference arm)
timeout 300 python bench.py > $O/r01_bench_w1.json 2> $O/r01_bench_w1.err
timeout 200 python bench.py --impl torch --steps 20 --warmup 5 > $O/r01_bench_w1_torch_eager.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r01_bench_reference_arm.json 2>/dev/null
# 2. every launch of the timed region with its device time (graph replay: kernel nodes are profiled individually)
timeout 500 ncu --nvtx --nvtx-include "timed" --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r01_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.out 2>&1
# 3. full captures of the hand-written kernels (one big SyncBN layer fwd+bwd; loss; fused SGD inside a real iteration, eager)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:syncbn -c 6 -f -o $O/r01_syncbn_full python tools/bn_one.py > $O/ncu_bn.out 2>&1
# 4. per-layer SyncBN table, torch profiler breakdown
timeout 150 python tools/bn_table.py > $O/r01_syncbn_per_layer.txt 2>/dev/null
ls -la $O | tail -20
