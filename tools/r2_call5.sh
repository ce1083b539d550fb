#!/bin/bash
# Round-2 GPU call 5 (1 GPU): ABI v7 (bf16-only gradient segments), batched re-poll in the SyncBN exchange, unrolled colsum.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c5_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c5_$name.log" | cut -c1-300; }
TMO=700 run gpu_tests python -m pytest tests -m gpu -q
run bench_default python bench.py --no-cpu-baseline
run bn_phases python tools/bn_phases.py
run bn_table python tools/bn_table.py
TMO=300 run bn_dram ncu --nvtx --nvtx-include "measure/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/c5_bn_dram.csv python tools/bn_dram.py run
