#!/usr/bin/env python
"""GPU-box diagnostic: eager vs CUDA-graph replay, step by step (fp32), plus eager vs eager(cudnn.benchmark) as the
yardstick for run-to-run reassociation drift on this model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_sod_project_b200.engine import Trainer
from distributed_sod_project_b200.synthetic import synth_batch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False

def run(graph, bench=False, steps=4):
    torch.backends.cudnn.benchmark = bench
    tr = Trainer("res50", dtype=torch.float32, channels_last=True, use_graph=graph, report_items=False)
    out = []
    for it in range(steps):
        x, m = synth_batch(1234 + 1000 * it, 4, 128)
        red, items, _ = tr.forward_backward_update(x.cuda(), m.cuda())
        torch.cuda.synchronize()
        out.append((float(red), tr.optimizer.flat.param.clone(), tr.optimizer.flat.mom.clone()))
    return out

a = run(False); b = run(True); c = run(False, bench=True); d = run(False)
for it in range(len(a)):
    def rel(u, v): return ((u - v).abs().max() / (u.abs().max() + 1e-30)).item()
    print(f"it{it}: loss eager {a[it][0]:.6f} graph {b[it][0]:.6f} eager-bench {c[it][0]:.6f} eager2 {d[it][0]:.6f} | "
          f"param rel diff graph {rel(a[it][1], b[it][1]):.2e} bench {rel(a[it][1], c[it][1]):.2e} rerun {rel(a[it][1], d[it][1]):.2e} | "
          f"mom rel diff graph {rel(a[it][2], b[it][2]):.2e} bench {rel(a[it][2], c[it][2]):.2e}")
