"""`res50` / `cp_res50`: the TestModel plugin pair of the reference (network/TestModel.py:17-185).

Zero-arg factories looked up by name from `user_config["model"]` (reference train.py:141):
ResNet-50 encoder → five 1×1 "trans" convs to 64 ch → top-down decoder of SIM + ConvBNReLU with
bilinear upsample-add → 1×1 classifier → logits [N,1,H,W].

`cp_res50` wraps the same twelve segments in `torch.utils.checkpoint` (reference
network/TestModel.py:51-66): activations are recomputed in backward, so every BN forward — and,
distributed, its statistics exchange and running-stat update — runs twice per iteration.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint

from .blocks import SIM, ConvBNReLU, bilinear, resnet50_stages, upsample_add

_TAPS = (32, 16, 8, 4, 2)                       # encoder strides tapped, deepest first
_TAP_CHANNELS = {32: 2048, 16: 1024, 8: 512, 4: 256, 2: 64}


class _TestModel(nn.Module):
    recompute = False  # cp_res50 flips this

    def __init__(self):
        super().__init__()
        self.div_2, self.div_4, self.div_8, self.div_16, self.div_32 = resnet50_stages()
        for s in _TAPS:
            setattr(self, f"trans{s}", nn.Conv2d(_TAP_CHANNELS[s], 64, 1, 1))
        for s in _TAPS:
            setattr(self, f"sim{s}", SIM(64, 32))
        for s in _TAPS:
            setattr(self, f"upconv{s}", ConvBNReLU(64, 32 if s == 2 else 64, 3, 1, 1))
        self.upconv1 = ConvBNReLU(32, 32, 3, 1, 1)
        self.classifier = nn.Conv2d(32, 1, 1)

    # -- segments (each is one checkpoint unit in cp_res50) -------------------------------------
    def _encode(self, stride: int, x):
        return getattr(self, f"div_{stride}")(x)

    def _project(self, f2, f4, f8, f16, f32):
        return self.trans2(f2), self.trans4(f4), self.trans8(f8), self.trans16(f16), self.trans32(f32)

    def _decode(self, stride: int, lateral, coarser=None):
        x = lateral if coarser is None else upsample_add(coarser, lateral)
        sim, up = getattr(self, f"sim{stride}"), getattr(self, f"upconv{stride}")
        return up(sim(x) + x)

    def _head(self, d2):
        return self.classifier(self.upconv1(bilinear(d2, scale_factor=2)))

    def _run(self, fn, *args):
        if self.recompute and torch.is_grad_enabled():
            # preserve_rng_state=False: the segments hold no dropout / RNG op, and saving the CUDA RNG state is not
            # allowed while a CUDA graph is being captured (the default config captures cp_res50 iterations)
            return checkpoint(fn, *args, use_reentrant=False, preserve_rng_state=False)
        return fn(*args)

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        feats, x = {}, image
        for s in (2, 4, 8, 16, 32):
            x = self._run(lambda t, s=s: self._encode(s, t), x)
            feats[s] = x
        lat = dict(zip((2, 4, 8, 16, 32), self._run(self._project, *(feats[s] for s in (2, 4, 8, 16, 32)))))
        d = None
        for s in _TAPS:
            d = self._run(lambda a, b=None, s=s: self._decode(s, a, b), lat[s], *(() if d is None else (d,)))
        return self._run(self._head, d)


class res50(_TestModel):
    """Plain variant (reference network/TestModel.py:129-185)."""


class cp_res50(_TestModel):
    """Activation-checkpointed variant (reference network/TestModel.py:17-127)."""

    recompute = True


def load_pretrained_backbone(model: _TestModel, path: str) -> list[str]:
    """Load torchvision-format ResNet-50 weights from a local file into `div_*`
    (the reference downloads them, backbone/origin/resnet.py:199-217; no network here).
    Keeps the reference's filter-then-update semantics: unknown keys are ignored. Returns loaded keys."""
    src = torch.load(path, map_location="cpu")
    remap = {"conv1.": "div_2.0.", "bn1.": "div_2.1.", "layer1.": "div_4.1.", "layer2.": "div_8.",
             "layer3.": "div_16.", "layer4.": "div_32."}
    own = model.state_dict()
    picked = {}
    for k, v in src.items():
        for old, new in remap.items():
            if k.startswith(old):
                nk = new + k[len(old):]
                if nk in own and own[nk].shape == v.shape:
                    picked[nk] = v
                break
    own.update(picked)
    model.load_state_dict(own)
    return sorted(picked)
