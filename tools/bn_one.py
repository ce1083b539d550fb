#!/usr/bin/env python
"""GPU-box tool for ncu: a few launches of the SyncBN forward/backward kernels on one layer shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sod_project_b200.syncbn import SyncBatchNorm, raw_backward
n, c, h, w = (int(v) for v in os.environ.get("SHAPE", "16,64,160,160").split(","))
has_pre = bool(int(os.environ.get("PRE", 0))); has_res = bool(int(os.environ.get("RES", 0))); relu = bool(int(os.environ.get("RELU", 1)))
dtype = torch.bfloat16
mk = lambda: torch.randn((n, c, h, w), device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
x, dy = mk(), mk(); pre = mk() if has_pre else None; res = mk() if has_res else None
bn = SyncBatchNorm(c).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        flush.zero_()
        y = bn.fused_forward(x, pre_add=pre, residual=res, relu=relu)
    weight = bn.weight.detach(); mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda")
    for _ in range(3):
        flush.zero_()
        raw_backward(dy, x, pre, y if relu else None, weight, mean, invstd, relu, has_res, bias=bn.bias.detach())   # default variant: mask from x
torch.cuda.synchronize()
print("done")
