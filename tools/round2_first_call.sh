#!/bin/bash
# One gpurun call that validates and measures everything that was prepared without a GPU at the end of round 1
# (DESIGN §4.8).  Usage from the build container:
#     gpurun --timeout 1800 -- 'bash tools/round2_first_call.sh'        (≈20 minutes of box time)
# Results land in gpurun_out/r2_*.  Every step has its own timeout so that a hang costs minutes, not the box.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; ( time timeout 900 "$@" ) > "gpurun_out/r2_$name.log" 2>&1; echo "   exit $?"; tail -3 "gpurun_out/r2_$name.log"; }

SOD_EXPERIMENTAL=1 run gpu_experimental python -m pytest tests/test_gpu_syncbn.py tests/test_gpu_step.py -m gpu -q -k "mask_from_x or step_from_host"
SOD_EXPERIMENTAL=1 run gpu_maxpool      python -m pytest tests/test_gpu_resample.py -m gpu -q -k maxpool
SOD_MAXPOOL=1 run gpu_step_maxpool      python -m pytest tests/test_gpu_step.py tests/test_gpu_train_cli.py -m gpu -x -q
SOD_BN_MASK_FROM_X=1 run gpu_xmask      python -m pytest tests -m gpu -x -q
SOD_BN_MASK_FROM_X=1 SOD_BN_L2_HINTS=1 run gpu_xmask_hints python -m pytest tests/test_gpu_syncbn.py tests/test_gpu_step.py -m gpu -x -q
run ab_bn_bwd          python tools/ab_bn_bwd_variants.py
run bench_default      python bench.py --no-cpu-baseline
SOD_E2E_PREFETCH=1 run bench_prefetch   python bench.py --no-cpu-baseline
SOD_BN_MASK_FROM_X=1 run bench_xmask    python bench.py --no-cpu-baseline
SOD_MAXPOOL=1 run bench_maxpool          python bench.py --no-cpu-baseline
SOD_CUDNN_BENCH_LIMIT=0 run bench_cudnn_all_engines python bench.py --no-cpu-baseline
SOD_BN_MASK_FROM_X=1 SOD_BN_L2_HINTS=1 SOD_E2E_PREFETCH=1 SOD_MAXPOOL=1 run bench_all python bench.py --no-cpu-baseline
grep -h '"metric"' gpurun_out/r2_bench_*.log | python -c '
import json, sys
for l in sys.stdin:
    d = json.loads(l)
    print(round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["roofline"]["frac"], 3), d.get("extras", {}).keys())
' || true
