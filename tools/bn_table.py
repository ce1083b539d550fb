#!/usr/bin/env python
"""GPU-box tool: per-layer time / bandwidth of the SyncBN forward and backward kernels on the TestModel's
84 layer shapes (bs 16, 320x320, bf16, channels-last), each launch alone with L2 flushed (CUDA events)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sod_project_b200 import network, syncbn
from distributed_sod_project_b200.syncbn import SyncBatchNorm, convert_syncbn_model, raw_backward

bs = int(os.environ.get("BS", 16)); size = int(os.environ.get("SIZE", 320))
dtype = torch.bfloat16
model = convert_syncbn_model(network.res50().cuda().to(memory_format=torch.channels_last)).train()
syncbn.TRACE = []
with torch.autocast("cuda", dtype=dtype):
    model(torch.randn(bs, 3, size, size, device="cuda").contiguous(memory_format=torch.channels_last))
trace, syncbn.TRACE = syncbn.TRACE, None
uniq = {}
for t in trace:
    uniq[t] = uniq.get(t, 0) + 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
tot = {"fwd_us": 0.0, "bwd_us": 0.0, "fwd_bytes": 0, "bwd_bytes": 0}
for (n, c, h, w, has_pre, has_res, relu), count in sorted(uniq.items(), key=lambda kv: -kv[0][0] * kv[0][1] * kv[0][2] * kv[0][3]):
    mk = lambda: torch.randn((n, c, h, w), device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    x, dy = mk(), mk(); pre = mk() if has_pre else None; res = mk() if has_res else None
    bn = SyncBatchNorm(c).cuda()
    elems = n * c * h * w
    fb = (1 + has_pre + has_res + 1) * 2 * elems
    bb = (2 + has_pre + relu + 1 + has_res) * 2 * elems
    def timeit(fn, iters=6):
        ts = []
        for i in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        return sorted(ts)[len(ts) // 2]
    with torch.no_grad():
        y = bn.fused_forward(x, pre_add=pre, residual=res, relu=relu)
        f_us = timeit(lambda: bn.fused_forward(x, pre_add=pre, residual=res, relu=relu))
    weight = bn.weight.detach(); mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda")
    b_us = timeit(lambda: raw_backward(dy, x, pre, y if relu else None, weight, mean, invstd, relu, has_res))
    rows.append(dict(shape=[n, c, h, w], pre=has_pre, res=has_res, relu=relu, count=count, MB=round(2 * elems / 1e6, 2),
                     fwd_us=round(f_us, 1), fwd_gbs=round(fb / f_us / 1e3), bwd_us=round(b_us, 1), bwd_gbs=round(bb / b_us / 1e3)))
    tot["fwd_us"] += f_us * count; tot["bwd_us"] += b_us * count; tot["fwd_bytes"] += fb * count; tot["bwd_bytes"] += bb * count
for r in rows:
    print(json.dumps(r))
print(json.dumps({"total_fwd_us": round(tot["fwd_us"]), "fwd_gbs": round(tot["fwd_bytes"] / tot["fwd_us"] / 1e3),
                  "total_bwd_us": round(tot["bwd_us"]), "bwd_gbs": round(tot["bwd_bytes"] / tot["bwd_us"] / 1e3)}))
