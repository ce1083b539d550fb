"""The oracle (oracle/*.py) against the golden vectors produced by the UNMODIFIED reference
(tools/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import loss as oloss
from oracle import sgd as osgd

CASES = ["n1", "n7", "n1000", "n4097", "img", "mask_zero", "mask_one", "binary", "extreme"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("red", ["mean", "sum"])
def test_loss_value_and_grad(golden, case, red):
    g = golden("loss_kat.npz")
    x, t = g[f"{case}/{red}/x"], g[f"{case}/{red}/t"]
    out = oloss.bce_cel_fwd_bwd(x, t, reduction=red)
    assert out["bce"] == pytest.approx(float(g[f"{case}/{red}/bce"]), rel=1e-12, abs=1e-14)
    assert out["cel"] == pytest.approx(float(g[f"{case}/{red}/cel"]), rel=1e-12, abs=1e-14)
    np.testing.assert_allclose(out["grad"], g[f"{case}/{red}/grad"], rtol=1e-10, atol=1e-15)
    assert oloss.bce_with_logits(x, t, red) == pytest.approx(out["bce"], rel=1e-14)
    assert oloss.cel(x, t) == pytest.approx(out["cel"], rel=1e-14)


def test_total_loss_strings(golden):
    g = golden("loss_kat.npz")
    x, t = g["img/mean/x"].astype(np.float32), g["img/mean/t"].astype(np.float32)
    out = oloss.bce_cel_fwd_bwd(x, t)
    assert out["total"] == pytest.approx(float(g["total_loss/value"]), rel=1e-6)
    assert oloss.get_total_loss_strings([out["bce"], out["cel"]]) == list(g["total_loss/strings"])


def _segments(g, kind, it):
    """one Segment per parameter tensor, lr/wd from the reference optimizer's group"""
    names = list(g[f"{kind}/names"])
    group_of = g[f"{kind}/group_of"]
    sizes = {"div_2.weight": 35, "div_2.bias": 7, "div_4.weight": 42, "div_4.bias": 6, "div_16.weight": 18,
             "head.weight": 6, "head.bias": 2, "classifier.weight": 2, "classifier.bias": 1}
    segs, off = [], 0
    for n, gi in zip(names, group_of):
        ln = sizes[str(n)]
        if gi < 0:
            segs.append(osgd.Segment(off, off + ln, 0.0, 0.0, frozen=True))
        else:
            segs.append(osgd.Segment(off, off + ln, float(g[f"{kind}/lr{it}"][gi]), float(g[f"{kind}/group_wd"][gi])))
        off += ln
    return segs


@pytest.mark.parametrize("kind", ["f3_trick", "sgd_trick", "sgd_all"])
def test_sgd_matches_reference_optimizer(golden, kind):
    g = golden("sgd_kat.npz")
    p = g[f"{kind}/p0"].astype(np.float32).copy()
    v = np.zeros_like(p)
    for it in range(4):
        assert osgd.sgd_step(p, v, g[f"{kind}/g{it}"].astype(np.float32), _segments(g, kind, it))
        np.testing.assert_allclose(p, g[f"{kind}/p{it + 1}"], rtol=2e-6, atol=1e-7)
    if kind == "f3_trick":   # div_2.* never moves
        np.testing.assert_array_equal(p[:42], g[f"{kind}/p0"][:42])


def test_sgd_skips_on_overflow():
    p = np.ones(8, np.float32); v = np.zeros(8, np.float32); gr = np.ones(8, np.float32); gr[3] = np.inf
    assert not osgd.sgd_step(p, v, gr, [osgd.Segment(0, 8, 0.1, 0.0)])
    assert np.all(p == 1) and np.all(v == 0)


@pytest.mark.parametrize("kind", ["poly", "poly_warmup", "cosine_warmup", "f3_sche"])
def test_scheduler_table(golden, kind):
    g = golden("sgd_kat.npz")
    rows = g[f"sched/{kind}"]
    st = osgd.SchedulerState(30, kind, lr_decay=0.9, warmup_epoch=3)
    base = np.array([0.005, 0.05])
    for e, row in enumerate(rows):
        # the reference evaluates the coefficient once PER PARAM GROUP (utils/pipeline_ops.py:226-229),
        # and the warmup branches mutate total_num on every evaluation
        cs = [st.coefficient(e) for _ in base]
        if any(isinstance(c, complex) for c in cs) or np.isnan(row).any():
            continue   # reference goes complex-valued once (1 - e/total) < 0; nothing to pin
        np.testing.assert_allclose(base * np.array(cs), row, rtol=1e-12)


def test_allreduce_mean():
    a = np.arange(6, dtype=np.float32); b = np.ones(6, np.float32)
    np.testing.assert_allclose(osgd.allreduce_mean([a, b]), (a + b) / 2)
