#!/bin/bash
# Round-2 GPU call 2 (1 GPU): full GPU test suite on the ABI-v6 library, bench A/Bs, multi-scale, stock-torch arm, launch list.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c2_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c2_$name.log" | cut -c1-300; }
TMO=600 run gpu_tests python -m pytest tests -m gpu -x -q
run bench_default python bench.py --no-cpu-baseline
SOD_STEAL_WGRADS=0 run bench_nosteal python bench.py --no-cpu-baseline
SOD_BN_LAUNCH=coop run bench_coop python bench.py --no-cpu-baseline
SOD_BN_LAUNCH=pdl run bench_pdl python bench.py --no-cpu-baseline
run bench_multiscale python bench.py --no-cpu-baseline --multiscale
run bench_torch python bench.py --impl torch
run bench_torch_multiscale python bench.py --impl torch --multiscale
run bn_phases python tools/bn_phases.py
run bn_table python tools/bn_table.py
TMO=400 run ncu_launches ncu --nvtx --nvtx-include "timed" --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c2_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
grep -h '"metric"' $O/c2_bench_*.log | python -c '
import json, sys
for l in sys.stdin:
    d = json.loads(l)
    print(d.get("impl","b200"), round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["ms_per_step"],3), d.get("roofline", {}).get("frac"), d.get("gpu_launches"))
' || true
