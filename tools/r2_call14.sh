#!/bin/bash
# Round-2 GPU call 14 (1 GPU): SyncBN per-channel side effects spread over the grid — full suite, bench, phase stamps, per-layer table.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c14_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c14_$name.log" | cut -c1-300; }
TMO=800 run gpu_tests python -m pytest tests -m gpu -x -q
run bench_default python bench.py --no-cpu-baseline
run bn_phases python tools/bn_phases.py
run bn_table python tools/bn_table.py
