#!/bin/bash
# Round-2 GPU call 19 (1 GPU): default bench with the measured uint8 input-pipeline extra.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 300 python bench.py ) > $O/c19_bench_default.log 2>&1; echo "exit $?"; tail -c 600 $O/c19_bench_default.log
