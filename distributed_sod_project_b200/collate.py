"""Multi-scale batch collation — the loader-side neighbour of the hot path.

The reference's `_collate_fn` (utils/dataset.py:125-132) draws ONE size per batch from `size_list` with the shared
`random` module (every rank is seeded alike, utils/misc.py:38-43, so all ranks draw the same size), stacks the samples
and resizes — image bilinear `align_corners=False`, mask nearest.  As written it unpacks 2-tuples while
`ImageFolder.__getitem__` yields `(img, mask, name)` (utils/dataset.py:116) and the training loop expects names
(train.py:284), so the `size_list` path raises in the reference; this is the evident intent, with the names kept.

`resize_batch` is the same resize applied to an already stacked batch — on whatever device the batch lives on, so a
loader can ship fixed-size pinned batches and resize after the H2D copy (one bilinear + one nearest kernel per batch
instead of per-sample CPU work in the loader processes).
"""
from __future__ import annotations

import random
from typing import Sequence

import torch
from torch.nn.functional import interpolate


def resize_batch(img: torch.Tensor, mask: torch.Tensor, size: int):
    """img [N,3,H,W] → bilinear (align_corners=False); mask [N,1,H,W] → nearest; both to (size, size)."""
    if img.shape[-2:] == (size, size) and mask.shape[-2:] == (size, size):
        return img, mask
    img = interpolate(img, size=(size, size), mode="bilinear", align_corners=False)
    mask = interpolate(mask, size=(size, size), mode="nearest")
    return img, mask


def multiscale_collate(batch: Sequence[tuple], size_list: Sequence[int], rng=random):
    """`collate_fn` for `DataLoader`: list of (img[3,H,W], mask[1,H,W][, name]) → (img, mask[, names]) at one size drawn
    from `size_list` by `rng.choice` (the module-level `random` by default, as in the reference)."""
    size = rng.choice(list(size_list))
    cols = list(zip(*batch))
    img, mask = resize_batch(torch.stack(list(cols[0]), dim=0), torch.stack(list(cols[1]), dim=0), size)
    if len(cols) > 2:
        return img, mask, list(cols[2])
    return img, mask
