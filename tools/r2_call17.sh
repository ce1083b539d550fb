#!/bin/bash
# Round-2 GPU call 17 (1 GPU): how many strips may a wide-channel SyncBN layer use?  (hop-1 packet traffic cap, SOD_BN_TRAFFIC_DIV)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for d in 4 2 1; do
  echo "== traffic_div $d"
  SOD_BN_TRAFFIC_DIV=$d timeout 200 python tools/bn_table.py > $O/c17_bn_table_div$d.log 2>&1
  tail -1 $O/c17_bn_table_div$d.log
done
SOD_BN_TRAFFIC_DIV=2 timeout 200 python bench.py --no-cpu-baseline --no-extras > $O/c17_bench_div2.log 2>&1; grep -o '"value": [0-9.]*' $O/c17_bench_div2.log | head -1
SOD_BN_TRAFFIC_DIV=1 timeout 200 python bench.py --no-cpu-baseline --no-extras > $O/c17_bench_div1.log 2>&1; grep -o '"value": [0-9.]*' $O/c17_bench_div1.log | head -1
