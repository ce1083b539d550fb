"""Input-pipeline seam: what the reference's `ImageFolder` + `create_loader` hand to the training loop
(utils/dataset.py:72-156, consumed at train.py:284-292), with the tensor-side work on the GPU.

Reference, per sample on DataLoader workers: PIL decode → joint resize / flip / rotate → ColorJitter → ToTensor →
Normalize; then the collate function stacks fp32 tensors and — for multi-scale training — resizes the stacked batch
with `F.interpolate` (bilinear image, nearest mask) on the CPU; the fp32 batch (16 B per pixel) is pinned and copied.

Here the PIL part (decode and the geometric / colour transforms that produce 8-bit images) stays on the workers and the
batch stays uint8 until it is on the device: `preprocess_batch` is ONE kernel (csrc/pipeline.cu) doing ToTensor,
Normalize, the optional horizontal flip and the collate's resize, and writing the image channels-last.  4 B per pixel
cross PCIe instead of 16.  `DevicePrefetcher` is the `BackgroundGenerator(tr_loader, max_prefetch=2)` of train.py:285
with a copy stream: batch k+1 is uploaded and pre-processed while iteration k runs.
"""
from __future__ import annotations

import ctypes as C
import random

import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)     # utils/dataset.py:92
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_batch(img_u8: torch.Tensor, mask_u8: torch.Tensor | None = None, size: int | tuple[int, int] | None = None,
                     flip: torch.Tensor | None = None, dtype: torch.dtype = torch.float32,
                     mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """img_u8 [N,H,W,3] uint8 (HWC, as PIL/numpy produce it), mask_u8 [N,H,W] uint8 or None, both on the GPU.
    Returns (image [N,3,S,S] `dtype`, channels-last; mask [N,1,S,S] fp32 or None): exactly
    `Normalize(ToTensor(img))` / `ToTensor(mask)` followed by the multi-scale collate's `interpolate` to `size`
    (utils/dataset.py:86-96,110-116,125-132).  `flip` (uint8 [N] on the GPU): mirror sample i left-right first."""
    if not img_u8.is_cuda:
        raise _lib.SodError("preprocess_batch needs CUDA tensors (no CPU fallback; keep torchvision transforms for CPU runs)")
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[-1] != 3:
        raise ValueError("img_u8 must be uint8 [N,H,W,3]")
    img_u8 = img_u8.contiguous()
    n, hs, ws, _ = img_u8.shape
    ho, wo = (hs, ws) if size is None else ((size, size) if isinstance(size, int) else tuple(size))
    out = torch.empty((n, ho, wo, 3), dtype=dtype, device=img_u8.device)
    m_out = None
    if mask_u8 is not None:
        if mask_u8.dtype != torch.uint8 or tuple(mask_u8.shape[-2:]) != (hs, ws) or mask_u8.numel() != n * hs * ws:
            raise ValueError("mask_u8 must be uint8 [N,H,W] matching the images")
        mask_u8 = mask_u8.contiguous()
        m_out = torch.empty((n, 1, ho, wo), dtype=torch.float32, device=img_u8.device)
    if flip is not None:
        flip = flip.to(device=img_u8.device, dtype=torch.uint8).contiguous()
        if flip.numel() != n:
            raise ValueError("flip must hold one flag per sample")
    rc = _lib.lib().sod_preprocess_batch(
        img_u8.data_ptr(), mask_u8.data_ptr() if mask_u8 is not None else None, flip.data_ptr() if flip is not None else None,
        out.data_ptr(), _lib.dtype_code(dtype), m_out.data_ptr() if m_out is not None else None, n, hs, ws, ho, wo,
        (C.c_float * 3)(*mean), (C.c_float * 3)(*std), _lib.stream_ptr())
    _lib.check(rc, "sod_preprocess_batch")
    _lib.count_launch()
    return out.permute(0, 3, 1, 2), m_out          # logical NCHW over channels-last storage


class DevicePrefetcher:
    """Wraps an iterable of host batches `(img_u8 [N,H,W,3], mask_u8 [N,H,W], names)` (pinned or not) and yields
    device batches `(image, mask, names)` in the training format, one batch ahead: upload + `preprocess_batch` of batch
    k+1 run on a side stream while the caller trains on batch k (the reference prefetches on a host thread,
    train.py:285; dead `DataPrefetcher`, utils/useful_but_dont_use.py:61-105, is the same idea).

    `size_list`: the multi-scale training sizes (config.py `size_list`); one size per batch, drawn from a generator
    seeded identically on every rank — the reference's collate draws with the worker's `random` module
    (utils/dataset.py:126), which is only consistent across ranks by accident of seeding."""

    def __init__(self, loader, size_list=None, dtype: torch.dtype = torch.float32, seed: int = 0, flip_prob: float = 0.0,
                 device: torch.device | None = None):
        self.loader, self.size_list, self.dtype, self.flip_prob = loader, size_list, dtype, flip_prob
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.rng = random.Random(seed)
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        img, mask, names = batch
        size = self.rng.choice(self.size_list) if self.size_list else None
        with torch.cuda.stream(self.stream):
            img_d = img.to(self.device, non_blocking=True)
            mask_d = mask.to(self.device, non_blocking=True)
            flip = None
            if self.flip_prob > 0:
                flip = torch.tensor([self.rng.random() < self.flip_prob for _ in range(img.shape[0])], dtype=torch.uint8).to(self.device, non_blocking=True)
            x, m = preprocess_batch(img_d, mask_d, size=size, flip=flip, dtype=self.dtype)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return x, m, names, ready, (img_d, mask_d)

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            x, m, names, ready, keep = nxt
            try:
                nxt = self._stage(next(it))          # queue the next upload before handing this batch out
            except StopIteration:
                nxt = None
            torch.cuda.current_stream().wait_event(ready)
            for t in (x, m, *keep):
                t.record_stream(torch.cuda.current_stream())
            yield x, m, names
