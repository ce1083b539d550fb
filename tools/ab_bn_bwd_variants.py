#!/usr/bin/env python
"""GPU-box tool (round-2 first call): A/B of the experimental backward variant SOD_BN_BWD_MASK_FROM_X against the
default on every BN+ReLU-without-residual layer shape of the TestModel (bs 16, 320x320, bf16, channels-last).
Each launch alone, L2 flushed (CUDA events).  Also checks the two variants against each other: the ReLU mask is
re-derived from x with the forward's arithmetic, so dz may differ only by the summation order of the statistics.

    python tools/ab_bn_bwd_variants.py            → one JSON line per shape + totals
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


VARIANTS = [("default", False, False), ("xmask", True, False), ("hints", False, True), ("xmask_hints", True, True)]
tot = {name: 0.0 for name, _, _ in VARIANTS}
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
    eligible = relu and not has_res
    rec = dict(shape=[n, c, h, w], pre=has_pre, res=has_res, relu=relu, count=count, MB=round(2 * n * c * h * w / 1e6, 2),
               xmask_eligible=eligible)
    ref = None
    for name, xm, hints in VARIANTS:
        syncbn.MASK_FROM_X, syncbn.L2_HINTS = xm and eligible, hints
        try:
            us = timeit(lambda: raw_backward(*args, bias=bias))
            dz, _, dg, db = raw_backward(*args, bias=bias)
            torch.cuda.synchronize()
        finally:
            syncbn.MASK_FROM_X, syncbn.L2_HINTS = False, False
        rec[name + "_us"] = round(us, 1)
        tot[name] += us * count
        if ref is None:
            ref = (dz.float(), dg.clone(), db.clone())
        else:
            scale = max(float(ref[0].abs().max()), 1e-12)
            rec[name + "_dz_max_rel"] = float((ref[0] - dz.float()).abs().max()) / scale
            rec[name + "_dgamma_max_rel"] = float(((ref[1] - dg).abs() / (ref[1].abs() + 1e-3)).max())
    print(json.dumps(rec), flush=True)
print(json.dumps({"bwd_total_us": {k: round(v) for k, v in tot.items()},
                  "note": "84 launches per iteration, each alone with L2 flushed; xmask applies to the eligible layers only"}))
