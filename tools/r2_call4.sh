#!/bin/bash
# Round-2 GPU call 4 (1 GPU): full GPU suite, default bench (with ride-along extras), cuDNN autotune breadth, SyncBN phase stamps /
# per-layer table / DRAM traffic (ncu), launch list, full ncu captures of the loss and SGD kernels.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c4_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c4_$name.log" | cut -c1-300; }
TMO=700 run gpu_tests python -m pytest tests -m gpu -q
run bench_default python bench.py
SOD_CUDNN_BENCH_LIMIT=0 run bench_cudnn_all python bench.py --no-cpu-baseline --no-extras
SOD_COLSUM=0 run bench_nocolsum python bench.py --no-cpu-baseline --no-extras
run bn_phases python tools/bn_phases.py
run bn_table python tools/bn_table.py
TMO=300 run bn_dram ncu --nvtx --nvtx-include "measure/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/c4_bn_dram.csv python tools/bn_dram.py run
TMO=400 run ncu_launches ncu --nvtx --nvtx-include "timed" --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c4_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras
TMO=300 run ncu_loss_sgd ncu --set full --clock-control none --import-source on -k regex:"loss_bce_cel|sgd_local|grad_gather16|colsum" -c 8 -f -o $O/c4_loss_sgd_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-graph
ncu -i $O/c4_loss_sgd_full.ncu-rep --page raw --csv > $O/c4_loss_sgd_full_raw.csv 2>/dev/null
