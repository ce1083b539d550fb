#!/bin/bash
# Round-2 GPU call 21 (1 GPU): the driver's own sequence on the final tree: full GPU suite with -x, then smoke().
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 330 python -m pytest tests -m gpu -x -q ) > $O/c21_gpu_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/c21_gpu_tests.log
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/c21_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/c21_smoke.log
