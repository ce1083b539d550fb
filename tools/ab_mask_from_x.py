#!/usr/bin/env python
"""GPU-box tool (round-2 first call): A/B of the experimental backward variant SOD_BN_BWD_MASK_FROM_X against the
default on every BN+ReLU-without-residual layer shape of the TestModel (bs 16, 320x320, bf16, channels-last).
Each launch alone, L2 flushed (CUDA events).  Also checks the two variants against each other: the ReLU mask is
re-derived from x with the forward's arithmetic, so dz may differ only by the summation order of the statistics.

    python tools/ab_mask_from_x.py            → one JSON line per shape + totals
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distributed_sod_project_b200 import network, syncbn
from distributed_sod_project_b200.syncbn import SyncBatchNorm, convert_syncbn_model, raw_backward

bs, size, dtype = int(os.environ.get("BS", 16)), int(os.environ.get("SIZE", 320)), torch.bfloat16
model = convert_syncbn_model(network.res50().cuda().to(memory_format=torch.channels_last)).train()
syncbn.TRACE = []
with torch.autocast("cuda", dtype=dtype):
    model(torch.randn(bs, 3, size, size, device="cuda").contiguous(memory_format=torch.channels_last))
trace, syncbn.TRACE = syncbn.TRACE, None
uniq = {}
for t in trace:
    uniq[t] = uniq.get(t, 0) + 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=7):
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


tot = {"default_us": 0.0, "xmask_us": 0.0, "other_us": 0.0}
for (n, c, h, w, has_pre, has_res, relu), count in sorted(uniq.items(), key=lambda kv: -kv[0][0] * kv[0][1] * kv[0][2] * kv[0][3]):
    mk = lambda: torch.randn((n, c, h, w), device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)  # noqa: E731
    x, dy = mk(), mk()
    pre = mk() if has_pre else None
    res = mk() if has_res else None
    bn = SyncBatchNorm(c).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.linspace(0.5, 1.5, c)); bn.bias.copy_(torch.linspace(-0.3, 0.3, c))
    # a real forward, so that y, mean, invstd are the ones the backward's recomputation must agree with
    xr = x.clone().requires_grad_(True)
    y = bn.fused_forward(xr, pre_add=pre, residual=res, relu=relu)
    _, _, _, _, mean, invstd = y.grad_fn.saved_tensors          # (x, pre, y|None, weight, mean, invstd) of _SyncBNFn
    weight, bias = bn.weight.detach(), bn.bias.detach()
    args = (dy, x, pre, y.detach() if relu else None, weight, mean, invstd, relu, has_res)
    syncbn.MASK_FROM_X = False
    d_us = timeit(lambda: raw_backward(*args, bias=bias))
    dz0, _, dg0, db0 = raw_backward(*args, bias=bias)
    rec = dict(shape=[n, c, h, w], pre=has_pre, res=has_res, relu=relu, count=count, MB=round(2 * n * c * h * w / 1e6, 2),
               default_us=round(d_us, 1))
    if relu and not has_res:
        syncbn.MASK_FROM_X = True
        x_us = timeit(lambda: raw_backward(*args, bias=bias))
        dz1, _, dg1, db1 = raw_backward(*args, bias=bias)
        syncbn.MASK_FROM_X = False
        scale = float(dz0.float().abs().max())
        rec.update(xmask_us=round(x_us, 1), speedup=round(d_us / x_us, 3),
                   dz_max_rel=float((dz0.float() - dz1.float()).abs().max()) / max(scale, 1e-12),
                   dz_mismatch_frac=float((dz0 != dz1).float().mean()),
                   dgamma_max_rel=float(((dg0 - dg1).abs() / (dg0.abs() + 1e-3)).max()),
                   dbeta_max_rel=float(((db0 - db1).abs() / (db0.abs() + 1e-3)).max()))
        tot["default_us"] += d_us * count
        tot["xmask_us"] += x_us * count
    else:
        tot["other_us"] += d_us * count
    print(json.dumps(rec), flush=True)
print(json.dumps({"eligible_layers_default_us": round(tot["default_us"]), "eligible_layers_xmask_us": round(tot["xmask_us"]),
                  "other_layers_us": round(tot["other_us"]),
                  "bwd_total_default_us": round(tot["default_us"] + tot["other_us"]),
                  "bwd_total_with_xmask_us": round(tot["xmask_us"] + tot["other_us"])}))
