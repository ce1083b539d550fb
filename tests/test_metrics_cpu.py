"""Evaluation metrics without a GPU: (1) the oracle restatement against golden vectors from the UNMODIFIED reference
classes (tools/make_golden_metrics.py); (2) the product's host-side formulas (`metrics.metrics_from_hist`) fed by a numpy
emulation of the two integer kernels of csrc/pipeline.cu — the kernels only count, everything numeric lives in these
formulas, so this pins the arithmetic of the GPU path on the CPU."""
import numpy as np
import pytest

from oracle import metrics as om


def _load(golden):
    g = golden("metrics_kat.npz")
    n = int(g["n"])
    return g, [(g[f"pred{i}"], g[f"gt{i}"]) for i in range(n)]


emulate_kernels = om.emulate_kernels


def test_oracle_matches_the_reference_classes(golden):
    g, cases = _load(golden)
    tot = om.TotalMetric(len(cases))
    for i, (p8, g8) in enumerate(cases):
        pred, gt = om.normalise(p8, g8)
        tot.update(pred, gt)
    for name, mine in (("mae", tot.mae), ("meanf", tot.meanf), ("sm", tot.sm), ("em", tot.em), ("wfm", tot.wfm)):
        np.testing.assert_allclose(mine, g[name], rtol=1e-12, atol=1e-14, err_msg=name)
    np.testing.assert_allclose(tot.precision, g["precision"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(tot.recall, g["recall"], rtol=1e-12, atol=1e-14)
    res = tot.show()
    for key, val in zip(g["show_keys"], g["show_vals"]):
        assert res[str(key)] == pytest.approx(float(val), rel=1e-12), key


def test_histogram_formulas_match_the_reference(golden):
    from distributed_sod_project_b200.metrics import _wfm_host, metrics_from_hist
    g, cases = _load(golden)
    for i, (p8, g8) in enumerate(cases):
        head, hist = emulate_kernels(p8, g8)
        m = metrics_from_hist(hist, head, *p8.shape)
        # integer counts are exact; what differs from numpy's per-pixel float64 passes is only the order of summation
        assert m["mae"] == pytest.approx(float(g["mae"][i]), rel=1e-11, abs=1e-13), i
        assert m["meanf"] == pytest.approx(float(g["meanf"][i]), rel=1e-11, abs=1e-13), i
        # S-measure: the reference takes the ground truth's quadrant mean / variance in float32 (np.float32 cast,
        # utils/saliency_metric.py:159, pairwise float32 summation inside np.var); exact counts agree to float32 noise
        assert m["sm"] == pytest.approx(float(g["sm"][i]), rel=2e-7, abs=1e-9), i
        assert m["em"] == pytest.approx(float(g["em"][i]), rel=1e-11, abs=1e-13), i
        np.testing.assert_allclose(m["precision"], g["precision"][i], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(m["recall"], g["recall"][i], rtol=1e-12, atol=1e-14)
        assert _wfm_host(p8, g8) == pytest.approx(float(g["wfm"][i]), rel=1e-12, abs=1e-14), i


def test_histogram_formulas_on_random_images_vs_oracle():
    from distributed_sod_project_b200.metrics import metrics_from_hist
    rng = np.random.default_rng(0)
    for trial in range(40):
        h, w = int(rng.integers(3, 70)), int(rng.integers(3, 70))
        p8 = rng.integers(0, 256, (h, w)).astype(np.uint8)
        if trial % 3 == 0:
            p8 = (p8 // 32 * 32).astype(np.uint8)                       # few grey levels
        g8 = ((rng.random((h, w)) < rng.uniform(0.1, 0.9)) * int(rng.integers(1, 256))).astype(np.uint8)
        pred, gt = om.normalise(p8, g8)
        head, hist = emulate_kernels(p8, g8)
        m = metrics_from_hist(hist, head, h, w)
        assert m["mae"] == pytest.approx(om.mae(pred, gt), rel=1e-11)
        assert m["em"] == pytest.approx(om.emeasure(pred, gt), rel=1e-11)
        t = om.fmeasure_terms(pred, gt)
        if t is not None:
            np.testing.assert_allclose(m["precision"], t[0], rtol=1e-12, atol=1e-14)
            assert m["meanf"] == pytest.approx(t[2], rel=1e-11, abs=1e-13)
        ys, xs = np.nonzero(gt)
        cy, cx = int(round(ys.mean())) + 1, int(round(xs.mean())) + 1
        if 0 < gt.mean() < 1 and min(cy, h - cy, cx, w - cx) * min(cx, w - cx) > 1 and min(cy, h - cy) > 0:
            assert m["sm"] == pytest.approx(om.smeasure(pred, gt), rel=2e-6, abs=1e-8)
