"""Call-site marshalling of the SyncBN entry points, checked on CPU: the library call is replaced by a recorder and
every recorded positional argument is matched BY PARAMETER NAME (parsed from include/sod_b200.h) against the tensor it
must point to.  28+ positional pointers are easy to transpose and nothing else on a GPU-less machine would notice."""
import os
import re

import pytest
import torch

from distributed_sod_project_b200 import _lib, syncbn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _param_names(fn: str):
    text = open(os.path.join(ROOT, "include", "sod_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    args = re.search(r"\b" + fn + r"\s*\(([^)]*)\)\s*;", text).group(1)
    return [re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", a.strip()).group(1) for a in args.split(",")]


class _Recorder:
    def __init__(self, real):
        self._real, self.calls = real, {}

    def __getattr__(self, name):
        if name in ("sod_syncbn_fwd", "sod_syncbn_bwd"):
            def rec(*a):
                self.calls.setdefault(name, []).append(dict(zip(_param_names(name), a)))
                assert len(a) == len(_param_names(name)), (name, len(a))
                return 0
            return rec
        return getattr(self._real, name)


@pytest.fixture
def recorder(monkeypatch):
    rec = _Recorder(_lib.lib())
    monkeypatch.setattr(_lib, "lib", lambda: rec)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0x5151)
    monkeypatch.setattr(syncbn, "FORCE_LOCAL", True)
    syncbn._state.clear()
    yield rec
    syncbn._state.clear()


def _ptr(t):
    return None if t is None else t.data_ptr()


@pytest.mark.parametrize("mask_from_x", [False, True])
@pytest.mark.parametrize("with_res", [False, True])
def test_fwd_bwd_arguments_by_name(recorder, monkeypatch, mask_from_x, with_res):
    monkeypatch.setattr(syncbn, "MASK_FROM_X", mask_from_x)
    n, c, h, w = 2, 16, 3, 5
    cl = torch.channels_last
    x = torch.randn(n, c, h, w).contiguous(memory_format=cl).requires_grad_()
    pre = torch.randn(n, c, h, w).contiguous(memory_format=cl).requires_grad_()
    res = torch.randn(n, c, h, w).contiguous(memory_format=cl).requires_grad_() if with_res else None
    weight, bias = torch.ones(c, requires_grad=True), torch.zeros(c, requires_grad=True)
    rm, rv, nbt = torch.zeros(c), torch.ones(c), torch.zeros((), dtype=torch.int64)
    cb1 = torch.randn(c, requires_grad=True)
    y = syncbn._SyncBNFn.apply(x, pre, res, weight, bias, rm, rv, nbt, 0.1, 1e-5, True, True, cb1, None)
    f = recorder.calls["sod_syncbn_fwd"][0]
    assert (f["x"], f["pre_add"], f["residual"], f["y"]) == (_ptr(x), _ptr(pre), _ptr(res), _ptr(y))
    assert (f["gamma"], f["beta"], f["running_mean"], f["running_var"]) == (_ptr(weight), _ptr(bias), _ptr(rm), _ptr(rv))
    assert (f["rows"], f["channels"], f["relu"], f["training"], f["dtype"]) == (n * h * w, c, 1, 1, _lib.SOD_F32)
    assert f["momentum"] == pytest.approx(0.1) and f["eps"] == pytest.approx(1e-5)
    assert f["num_batches_tracked"] == _ptr(nbt) and f["comm"] is None and f["stats_off"] == 0 and f["epoch"] is None
    assert (f["conv_bias1"], f["conv_bias2"], f["conv_bias_dtype"]) == (_ptr(cb1), None, _lib.SOD_F32)
    ws = syncbn._state[None]["ws"]
    assert (f["workspace"], f["workspace_bytes"], f["stream"]) == (_ptr(ws), ws.numel(), 0x5151)
    assert f["save_mean"] and f["save_invstd"] and f["save_mean"] != f["save_invstd"] and f["seq"] >= 1

    dy = torch.randn_like(y)
    y.backward(dy)
    b = recorder.calls["sod_syncbn_bwd"][0]
    assert (b["x"], b["pre_add"], b["y"]) == (_ptr(x), _ptr(pre), _ptr(y))
    assert b["dy"] not in (None, b["x"], b["y"], b["dz"]) and b["dz"] not in (None, b["x"], b["y"])
    assert (b["dres"] is not None) == with_res
    assert (b["gamma"], b["beta"]) == (_ptr(weight), _ptr(bias))
    assert (b["save_mean"], b["save_invstd"]) == (f["save_mean"], f["save_invstd"])
    assert b["dgamma"] and b["dbeta"] and b["dgamma"] != b["dbeta"]
    assert (b["rows"], b["channels"], b["relu"], b["dtype"]) == (n * h * w, c, 1, _lib.SOD_F32)
    assert b["seq"] == f["seq"] + 1 and b["comm"] is None and b["epoch"] is None
    assert (b["conv_bias1"], b["conv_bias2"], b["dconv_bias2"], b["conv_bias_dtype"]) == (_ptr(cb1), None, None, _lib.SOD_F32)
    assert b["dconv_bias1"] is not None
    assert (b["workspace"], b["workspace_bytes"], b["stream"]) == (_ptr(ws), ws.numel(), 0x5151)
    # the experimental mask-from-x variant is requested only where it is defined: ReLU, no residual, and only when enabled
    assert bool(b["flags"] & _lib.SOD_BN_BWD_MASK_FROM_X) == (mask_from_x and not with_res)
    assert not (b["flags"] & _lib.SOD_BN_ACCUMULATE_PARAM_GRADS)        # γ/β have no bound fp32 .grad here


def test_bound_grads_are_written_directly(recorder):
    """FusedSGD binds fp32 .grad views of the flat buffer to γ/β: the kernel must get exactly those and the accumulate flag"""
    c = 16
    x = torch.randn(2, c, 4, 4).contiguous(memory_format=torch.channels_last).requires_grad_()
    weight, bias = torch.ones(c, requires_grad=True), torch.zeros(c, requires_grad=True)
    weight.grad, bias.grad = torch.zeros(c), torch.zeros(c)
    y = syncbn._SyncBNFn.apply(x, None, None, weight, bias, None, None, None, 0.1, 1e-5, False, True)
    y.sum().backward()
    b = recorder.calls["sod_syncbn_bwd"][0]
    assert (b["dgamma"], b["dbeta"]) == (weight.grad.data_ptr(), bias.grad.data_ptr())
    assert b["flags"] & _lib.SOD_BN_ACCUMULATE_PARAM_GRADS
    assert b["y"] is None and b["relu"] == 0 and b["pre_add"] is None
