#!/usr/bin/env python
"""GPU-box tool: DRAM traffic of the SyncBN backward on the TestModel's layer shapes, for `roofline.traffic` of the bench line.

    ncu --nvtx --nvtx-include "measure/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \\
        --clock-control none --csv --log-file gpurun_out/bn_dram.csv python tools/bn_dram.py run
    python tools/bn_dram.py parse gpurun_out/bn_dram.csv > profiles/r02_syncbn_bwd_dram.json

`run` launches every unique (shape, fusion) backward of one iteration once inside the NVTX range (inputs cold: L2 flushed
before each), in a fixed order that `parse` relies on; `parse` weights each by its count per iteration."""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (n, c, h, w, has_pre, has_res, relu) → launches per iteration; bs 16 at 320x320 (trace of network.res50, see tools/bn_table.py)
LAYERS = [
    ((16, 32, 320, 320, 0, 0, 1), 1), ((16, 64, 160, 160, 0, 0, 1), 2), ((16, 256, 80, 80, 0, 0, 0), 1), ((16, 256, 80, 80, 0, 1, 1), 3),
    ((16, 64, 160, 160, 1, 0, 1), 2), ((16, 128, 80, 80, 0, 0, 1), 1), ((16, 512, 40, 40, 0, 0, 0), 1), ((16, 512, 40, 40, 0, 1, 1), 4),
    ((16, 32, 160, 160, 0, 0, 1), 1), ((16, 64, 80, 80, 0, 0, 1), 8), ((16, 256, 40, 40, 0, 0, 1), 1), ((16, 1024, 20, 20, 0, 0, 0), 1),
    ((16, 1024, 20, 20, 0, 1, 1), 6), ((16, 64, 80, 80, 1, 0, 1), 2), ((16, 128, 40, 40, 0, 0, 1), 7), ((16, 512, 20, 20, 0, 0, 1), 1),
    ((16, 2048, 10, 10, 0, 0, 0), 1), ((16, 2048, 10, 10, 0, 1, 1), 3), ((16, 32, 80, 80, 0, 0, 1), 1), ((16, 32, 80, 80, 1, 0, 1), 1),
    ((16, 256, 20, 20, 0, 0, 1), 11), ((16, 64, 40, 40, 0, 0, 1), 2), ((16, 64, 40, 40, 1, 0, 1), 2), ((16, 512, 10, 10, 0, 0, 1), 5),
    ((16, 32, 40, 40, 0, 0, 1), 1), ((16, 32, 40, 40, 1, 0, 1), 1), ((16, 64, 20, 20, 0, 0, 1), 2), ((16, 64, 20, 20, 1, 0, 1), 2),
    ((16, 32, 20, 20, 0, 0, 1), 1), ((16, 32, 20, 20, 1, 0, 1), 1), ((16, 64, 10, 10, 0, 0, 1), 2), ((16, 64, 10, 10, 1, 0, 1), 2),
    ((16, 32, 10, 10, 0, 0, 1), 1), ((16, 32, 10, 10, 1, 0, 1), 1), ((16, 32, 5, 5, 0, 0, 1), 1), ((16, 32, 5, 5, 1, 0, 1), 1),
]
assert sum(c for _, c in LAYERS) == 84


def run():
    import torch
    from distributed_sod_project_b200 import syncbn
    from distributed_sod_project_b200.syncbn import raw_backward
    dtype = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (n, c, h, w, has_pre, has_res, relu), _ in LAYERS:
        mk = lambda: torch.randn((n, c, h, w), device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)  # noqa: E731
        x, dy = mk(), mk()
        pre = mk() if has_pre else None
        y = mk() if relu else None
        weight, bias = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        mean, invstd = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        args = (dy, x, pre, y, weight, mean, invstd, bool(relu), bool(has_res))
        raw_backward(*args, bias=bias)                      # warm-up (not profiled)
        flush.zero_()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("measure")
        raw_backward(*args, bias=bias)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    print("ran", len(LAYERS), "mask_from_x", syncbn.MASK_FROM_X, "l2_hints", syncbn.L2_HINTS)


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    recs = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r[0] != "ID"]
    by_id: dict = {}
    for r in recs:
        if "syncbn_bwd" not in r["Kernel Name"]:
            continue
        by_id.setdefault(int(r["ID"]), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    ids = sorted(by_id)
    assert len(ids) == len(LAYERS), (len(ids), len(LAYERS))
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1}
    out, tot_b, tot_alg, tot_ns = [], 0.0, 0.0, 0.0
    for i, ((n, c, h, w, has_pre, has_res, relu), count) in zip(ids, LAYERS):
        m = by_id[i]
        rd = m["dram__bytes_read.sum"][0] * unit[m["dram__bytes_read.sum"][1]]
        wr = m["dram__bytes_write.sum"][0] * unit[m["dram__bytes_write.sum"][1]]
        ns = m["gpu__time_duration.sum"][0] * unit[m["gpu__time_duration.sum"][1]]
        elems = n * c * h * w
        from_x = bool(relu) and not has_res
        alg = (2 + has_pre + (1 if relu and not from_x else 0) + 1 + has_res) * 2 * elems
        out.append({"shape": [n, c, h, w], "pre": has_pre, "res": has_res, "relu": relu, "count": count, "dram_read": rd, "dram_write": wr,
                    "us": ns / 1e3, "algorithmic": alg, "traffic_over_algorithmic": (rd + wr) / alg})
        tot_b += (rd + wr) * count; tot_alg += alg * count; tot_ns += ns * count
    print(json.dumps({"kernel": "syncbn_bwd_kernel (mask-from-x + L2 hints), bf16, bs16 320x320 layer set, each launch cold (L2 flushed)",
                      "launches_per_iteration": 84, "avg_dram_bytes_per_launch": tot_b / 84, "avg_algorithmic_bytes_per_launch": tot_alg / 84,
                      "traffic_over_algorithmic": tot_b / tot_alg, "ncu_us_per_iteration": tot_ns / 1e3, "layers": out}, indent=1))


if __name__ == "__main__":
    run() if sys.argv[1] == "run" else parse(sys.argv[2])
