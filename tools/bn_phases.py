#!/usr/bin/env python
"""GPU-box tool: where does a SyncBN kernel spend its time?  Per-CTA globaltimer stamps (start, end of phase 1,
end of exchange, end) for one layer shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distributed_sod_project_b200 import _lib, syncbn
from distributed_sod_project_b200.syncbn import SyncBatchNorm, raw_backward
syncbn.DEBUG_FLAGS = 4
for shape in os.environ.get("SHAPES", "16,64,160,160;16,64,80,80;16,256,80,80;16,32,320,320;16,256,20,20;16,2048,10,10").split(";"):
    n, c, h, w = (int(v) for v in shape.split(","))
    mk = lambda: torch.randn((n, c, h, w), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x, dy = mk(), mk()
    bn = SyncBatchNorm(c).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = syncbn._dev_state(x.device)
    off = int(_lib.lib().sod_syncbn_workspace_bytes(1, c))
    def stamps():
        raw = st["ws"][off:off + 160 * 64].view(torch.int64).cpu().numpy().reshape(160, 8)
        raw = raw[raw[:, 0] > 0]
        t0 = raw[:, 0].min()
        return (raw - t0) / 1e3, len(raw)
    with torch.no_grad():
        for name, fn in (("fwd", lambda: bn.fused_forward(x, relu=True)),):
            for _ in range(3):
                flush.zero_(); st["ws"][off:off + 160 * 64].zero_(); y = fn(); torch.cuda.synchronize()
            t, n_cta = stamps()
            print(f"{shape} fwd ctas={n_cta}: start max {t[:,0].max():.1f}us | phase1 end med {np.median(t[:,1]):.1f} max {t[:,1].max():.1f} | "
                  f"reduced med {np.median(t[:,4]):.1f} max {t[:,4].max():.1f} | slice published med {np.median(t[:,5]):.1f} max {t[:,5].max():.1f} | "
                  f"exchange end med {np.median(t[:,2]):.1f} max {t[:,2].max():.1f} | end med {np.median(t[:,3]):.1f} max {t[:,3].max():.1f}")
        weight = bn.weight.detach(); mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda")
        for _ in range(3):
            flush.zero_(); st["ws"][off:off + 160 * 64].zero_(); raw_backward(dy, x, None, y, weight, mean, invstd, True, False); torch.cuda.synchronize()
        t, n_cta = stamps()
        print(f"{shape} bwd ctas={n_cta}: start max {t[:,0].max():.1f}us | phase1 end med {np.median(t[:,1]):.1f} max {t[:,1].max():.1f} | "
              f"reduced med {np.median(t[:,4]):.1f} max {t[:,4].max():.1f} | slice published med {np.median(t[:,5]):.1f} max {t[:,5].max():.1f} | "
              f"exchange end med {np.median(t[:,2]):.1f} max {t[:,2].max():.1f} | end med {np.median(t[:,3]):.1f} max {t[:,3].max():.1f}")
