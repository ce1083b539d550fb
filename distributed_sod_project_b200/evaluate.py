"""Evaluation seam: `test()` / `_test_process()` of the reference (train.py:338-417) — distributed, with the
per-pixel work on the GPU.

Reference: rank 0 alone runs every test set through the model, moves each sigmoid map to the CPU, turns it into an
8-bit PIL image, resizes it to the ground truth's size and feeds `CalTotalMetric` (≈40 numpy passes per image).
Here every rank evaluates its shard (`indices[rank::world]`, the order `DistributedSampler(shuffle=False)` would
give), the sigmoid + 8-bit quantisation + normalisation + histogramming run as four small kernels per batch
(`metrics.SaliencyMetrics`), and `show()` combines the ranks with one all-reduce.  Predictions whose size differs from
the ground truth's are resized on the GPU with `F.interpolate(bilinear)` — the reference uses PIL's resize there
(`to_pil(out).resize(gt.size)`, train.py:392), a documented deviation that only matters for datasets with non-square
originals; the synthetic evaluation sets keep both sizes equal.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import comm
from .metrics import SaliencyMetrics


@torch.no_grad()
def test_process(model, batches, length: int | None = None, wfm: str | None = None) -> dict:
    """`_test_process` (train.py:372-417).  `batches`: iterable of (image [N,3,S,S] float CUDA, gt_u8 [N,H,W] uint8 CUDA)
    holding THIS rank's shard; `length`: total number of images over all ranks (checked in show(), like the reference's
    `num`).  Returns {"MaxF","MeanF","WFM","MAE","SM","EM"}."""
    was_training = model.training
    model.eval()
    cal = SaliencyMetrics(num=length, wfm=wfm)
    try:
        for x, gt in batches:
            logits = model(x)                                              # train.py:386-390
            pred_u8 = SaliencyMetrics.quantize(logits.float(), apply_sigmoid=True)[:, 0]    # .sigmoid() → ToPILImage
            if pred_u8.shape[-2:] != gt.shape[-2:]:
                p = F.interpolate(pred_u8[:, None].float(), size=gt.shape[-2:], mode="bilinear", align_corners=False)
                pred_u8 = p.round().clamp_(0, 255).to(torch.uint8)[:, 0]
            cal.update_batch(pred_u8, gt)
        return cal.show()
    finally:
        model.train(was_training)


def shard(n_items: int) -> range:
    """this rank's indices of an n-item evaluation set"""
    return range(comm.rank(), n_items, comm.world_size())
