#!/bin/bash
# Round-2 GPU call 10 (1 GPU): final full GPU suite, smoke(), default bench incl. CPU baseline, reference arm, cp_res50 line,
# full ncu captures of the SyncBN kernels on a large and a medium layer.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; echo "== $name"; ( time timeout ${TMO:-300} "$@" ) > "$O/c10_$name.log" 2>&1; echo "   exit $?"; tail -2 "$O/c10_$name.log" | cut -c1-300; }
TMO=800 run gpu_tests python -m pytest tests -m gpu -x -q
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_default python bench.py
run bench_reference python bench.py --impl reference --steps 6 --warmup 1
run bench_cp_res50 python bench.py --model cp_res50 --no-cpu-baseline --no-extras
TMO=300 run ncu_bn_big ncu --set full --clock-control none --import-source on -k regex:syncbn -c 6 -f -o $O/c10_syncbn_big python tools/bn_one.py
SHAPE=16,256,20,20 TMO=300 run ncu_bn_mid ncu --set full --clock-control none --import-source on -k regex:syncbn -c 6 -f -o $O/c10_syncbn_mid python tools/bn_one.py
ncu -i $O/c10_syncbn_big.ncu-rep --page raw --csv > $O/c10_syncbn_big_raw.csv 2>/dev/null
ncu -i $O/c10_syncbn_mid.ncu-rep --page raw --csv > $O/c10_syncbn_mid_raw.csv 2>/dev/null
