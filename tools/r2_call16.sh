#!/bin/bash
# Round-2 GPU call 16 (1 GPU): launch list and SyncBN DRAM traffic of the FINAL code (for the "where the iteration goes" table).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c16_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c16_$name.log" | cut -c1-300; }
TMO=400 run ncu_launches ncu --nvtx --nvtx-include "timed" --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c16_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras
TMO=300 run bn_dram ncu --nvtx --nvtx-include "measure/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/c16_bn_dram.csv python tools/bn_dram.py run
