"""Multi-scale collate: same draw and same pixels as the reference's `_collate_fn` on the 2-tuples it can
handle, plus the 3-tuple form the training loop actually needs."""
import os
import random
import subprocess
import sys

import torch

from distributed_sod_project_b200.collate import multiscale_collate, resize_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _samples(n=5, h=40, w=40, names=True):
    g = torch.Generator().manual_seed(3)
    out = []
    for i in range(n):
        img = torch.randn(3, h, w, generator=g)
        mask = (torch.rand(1, h, w, generator=g) * 255).round() / 255
        out.append((img, mask, f"img{i}") if names else (img, mask))
    return out


def test_three_tuples_keep_names_and_one_size_per_batch():
    random.seed(11)
    sizes = set()
    for _ in range(12):
        img, mask, names = multiscale_collate(_samples(), [24, 40, 56])
        assert img.shape[0] == 5 and img.shape[1] == 3 and mask.shape[1] == 1 and names == [f"img{i}" for i in range(5)]
        assert img.shape[-2:] == mask.shape[-2:] and img.shape[-1] == img.shape[-2]
        sizes.add(img.shape[-1])
        assert set(mask.unique().tolist()) <= set((torch.arange(256) / 255).tolist())     # nearest: no new mask values
    assert sizes == {24, 40, 56}


def test_same_seed_same_sizes_on_every_rank():
    draws = []
    for _rank in range(2):
        random.seed(0)                                   # utils/misc.py:38-43 seeds `random` alike on all ranks
        draws.append([multiscale_collate(_samples(2), [256, 320, 384])[0].shape[-1] for _ in range(6)])
    assert draws[0] == draws[1]


def test_native_size_is_a_no_op():
    img, mask = torch.randn(2, 3, 32, 32), torch.rand(2, 1, 32, 32)
    a, b = resize_batch(img, mask, 32)
    assert a is img and b is mask


def test_pixels_match_reference_collate():
    code = r'''
import sys, types, random, torch
sys.path.insert(0, "/root/reference"); sys.path.insert(0, %r); sys.path.insert(0, %r)
for n in ("openpyxl", "thop", "prefetch_generator"):
    mm = types.ModuleType(n); mm.load_workbook = mm.Workbook = mm.profile = mm.BackgroundGenerator = None; sys.modules[n] = mm
from utils.dataset import _collate_fn as ref_collate
from distributed_sod_project_b200.collate import multiscale_collate
from test_collate_cpu import _samples
batch = _samples(names=False)
for seed in range(4):
    random.seed(seed); a = ref_collate(batch, [24, 40, 56])
    random.seed(seed); b = multiscale_collate(batch, [24, 40, 56])
    assert len(a) == len(b) == 2 and all(torch.equal(x, y) for x, y in zip(a, b))
print("SAME")
''' % (ROOT, os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "SAME" in out.stdout, out.stderr[-2000:]
