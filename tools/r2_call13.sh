#!/bin/bash
# Round-2 GPU call 13 (1 GPU): CLI tests incl. test.py, pipeline/metrics tests after the last host-side edits.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c13_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c13_$name.log" | cut -c1-300; }
TMO=500 run gpu_tests python -m pytest tests/test_gpu_train_cli.py tests/test_gpu_pipeline.py tests/test_gpu_sgd.py -m gpu -q
