#!/bin/bash
# Round-2 GPU call 18 (2 GPUs): the default train.py configuration (cp_res50, graph, uint8 pipeline, val + test) at world 2.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 400 python -m pytest tests/test_gpu_train_cli.py -m gpu -q -k "two_gpus" ) > $O/c18_cli_w2.log 2>&1; echo "exit $?"; tail -3 $O/c18_cli_w2.log
