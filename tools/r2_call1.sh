#!/bin/bash
# Round-2 first GPU call (trimmed from tools/round2_first_call.sh to ≈8 minutes of box time).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; ( time timeout 600 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   exit $?"; tail -3 "gpurun_out/r2_$name.log"; }
SOD_EXPERIMENTAL=1 run gpu_experimental python -m pytest tests/test_gpu_syncbn.py tests/test_gpu_step.py tests/test_gpu_resample.py -m gpu -q -k "mask_from_x or step_from_host or maxpool"
SOD_BN_MASK_FROM_X=1 SOD_BN_L2_HINTS=1 SOD_MAXPOOL=1 run gpu_all_on python -m pytest tests/test_gpu_syncbn.py tests/test_gpu_step.py tests/test_gpu_train_cli.py -m gpu -x -q
run ab_bn_bwd          python tools/ab_bn_bwd_variants.py
run bench_default      python bench.py --no-cpu-baseline
SOD_E2E_PREFETCH=1 SOD_MAXPOOL=1 run bench_prefetch_maxpool python bench.py --no-cpu-baseline
SOD_BN_MASK_FROM_X=1 SOD_BN_L2_HINTS=1 SOD_E2E_PREFETCH=1 SOD_MAXPOOL=1 run bench_all python bench.py --no-cpu-baseline
