"""Synthetic image/mask batches with the dataloader's output contract (reference utils/dataset.py:86-116:
image float32 [N,3,S,S] after Normalize, mask float32 [N,1,S,S] in [0,1] on the k/255 grid that
`ToTensor` produces from an 8-bit PNG).  Shared by golden generation, tests and bench."""
from __future__ import annotations

import torch


def synth_batch(seed: int, n: int, size: int):
    """Deterministic (seeded, CPU generator) batch: image ~ N(0,1); mask = soft-edged disk."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, size, size, generator=g)
    cy = torch.rand(n, generator=g) * size
    cx = torch.rand(n, generator=g) * size
    rad = (0.15 + 0.25 * torch.rand(n, generator=g)) * size
    ar = torch.arange(size, dtype=torch.float32)
    yy, xx = torch.meshgrid(ar, ar, indexing="ij")
    d = ((yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2).sqrt()
    m = ((rad[:, None, None] - d) / 2.0 + 0.5).clamp(0, 1)
    m = (m * 255).round() / 255
    return x, m[:, None].contiguous()


def synth_eval_set(name: str, indices, batch_size: int, size: int):
    """Deterministic synthetic evaluation set: yields (image [n,3,S,S] f32 CUDA, gt_u8 [n,S,S] uint8 CUDA) for the given
    image indices (each index is one image, the same on whichever rank evaluates it)."""
    import zlib
    base = zlib.crc32(name.encode()) % 100000
    idx = list(indices)
    for i in range(0, len(idx), batch_size):
        xs, gs = [], []
        for j in idx[i:i + batch_size]:
            x, m = synth_batch(base + j, 1, size)
            xs.append(x)
            gs.append((m[:, 0] * 255).round().to(torch.uint8))
        yield torch.cat(xs).cuda(non_blocking=True), torch.cat(gs).cuda(non_blocking=True)
