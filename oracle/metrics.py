"""Oracle (TEST INFRASTRUCTURE, never imported by the product path): numpy restatement of the reference's evaluation
metrics and of the per-image normalisation its test loop applies before them.

Follows, formula by formula (paths in the reference repository):
  normalise        train.py:396-409          (gt / (max + 1e-8) > 0.5; prediction min-max stretched, /255 if constant)
  MAE              utils/saliency_metric.py:57-73
  F-measure        utils/saliency_metric.py:8-54     (adaptive-threshold meanF; 255-bin PR curves → maxF)
  S-measure        utils/saliency_metric.py:76-177   (object + region terms, α = 0.5)
  E-measure        utils/saliency_metric.py:180-229
The weighted F-measure (utils/saliency_metric.py:232-303: Euclidean distance transform + 7x7 Gaussian) is restated here
as well (scipy), the product computes it on the host from the same arrays (see metrics.py).
Pinned against the UNMODIFIED reference classes by tools/make_golden_metrics.py → tests/golden/metrics_kat.npz.
"""
from __future__ import annotations

import numpy as np


def normalise(pred_u8: np.ndarray, gt_u8: np.ndarray):
    """train.py:396-409 → (pred float64 in [0,1], gt int {0,1})"""
    gt = gt_u8 / (gt_u8.max() + 1e-8)
    gt = np.where(gt > 0.5, 1, 0)
    mx, mn = pred_u8.max(), pred_u8.min()
    pred = pred_u8 / 255 if mx == mn else (pred_u8 - mn) / (mx - mn)
    return pred, gt


def mae(pred, gt):
    return np.mean(np.abs(pred - gt))


def fmeasure_terms(pred, gt):
    """→ (precision[255], recall[255], meanF) of one image, or None when the ground truth is empty"""
    if gt.max() == 0:
        return None
    th = min(2 * pred.mean(), 1)
    binary = (pred >= th).astype(np.float64)
    hard = (gt > 0.5).astype(np.float64)
    tp = (binary * hard).sum()
    if tp == 0:
        mf = 0.0
    else:
        pre, rec = tp / binary.sum(), tp / hard.sum()
        mf = 1.3 * pre * rec / (0.3 * pre + rec)
    p8 = np.uint8(pred * 255)
    t_hist, _ = np.histogram(p8[gt > 0.5], bins=range(256))
    n_hist, _ = np.histogram(p8[gt <= 0.5], bins=range(256))
    t_cum, n_cum = np.cumsum(np.flip(t_hist)), np.cumsum(np.flip(n_hist))
    return t_cum / (t_cum + n_cum + 1e-8), t_cum / np.sum(gt), mf


def _s_object(vals):
    x, s = np.mean(vals), np.std(vals)
    return 2 * x / (x * x + 1 + s + 1e-8)


def _ssim(p, g):
    g = np.float32(g)
    n = p.size
    x, y = np.mean(p), np.mean(g)
    sx, sy = np.var(p), np.var(g)
    sxy = np.sum((p - x) * (g - y)) / (n - 1)
    alpha, beta = 4 * x * y * sxy, (x * x + y * y) * (sx + sy)
    if alpha != 0:
        return alpha / (beta + 1e-8)
    return 1 if beta == 0 else 0


def smeasure(pred, gt, alpha=0.5):
    gt = gt > 0.5
    y = np.mean(gt)
    if y == 0:
        return 1 - np.mean(pred)
    if y == 1:
        return np.mean(pred)
    obj = y * _s_object((pred * gt)[gt]) + (1 - y) * _s_object(((1 - pred) * (1 - gt))[~gt])
    h, w = gt.shape
    ys, xs = np.nonzero(gt)
    cy, cx = int(round(ys.mean())) + 1, int(round(xs.mean())) + 1      # scipy center_of_mass of a binary map
    area = h * w
    reg = 0.0
    for (r0, r1, c0, c1) in ((0, cy, 0, cx), (0, cy, cx, w), (cy, h, 0, cx), (cy, h, cx, w)):
        wgt = (r1 - r0) * (c1 - c0) / area
        reg = reg + wgt * _ssim(pred[r0:r1, c0:c1], gt[r0:r1, c0:c1])
    return alpha * obj + (1 - alpha) * reg


def emeasure(pred, gt):
    th = min(2 * pred.mean(), 1)
    fm = (pred >= th).astype(np.float64)
    g = gt.astype(bool)
    if g.sum() == 0:
        enhanced = 1.0 - fm
    elif (~g).sum() == 0:
        enhanced = fm
    else:
        dg = g.astype(np.float64)
        a_f, a_g = fm - fm.mean(), dg - dg.mean()
        align = 2.0 * (a_g * a_f) / (a_g * a_g + a_f * a_f + 1e-8)
        enhanced = (align + 1) ** 2 / 4
    return enhanced.sum() / (gt.size - 1 + 1e-8)


def wfmeasure(pred, gt, beta=1, eps=1e-6):
    from scipy.ndimage import convolve, distance_transform_edt
    gt = gt > 0.5
    if gt.max() == 0:
        return 0.0
    dst, idx = distance_transform_edt(gt == 0, return_indices=True)
    e = np.abs(pred - gt)
    et = np.copy(e)
    et[gt == 0] = et[idx[0][gt == 0], idx[1][gt == 0]]
    yy, xx = np.ogrid[-3:4, -3:4]
    k = np.exp(-(xx * xx + yy * yy) / 50.0)
    k[k < np.finfo(k.dtype).eps * k.max()] = 0
    k /= k.sum()
    ea = convolve(et, weights=k, mode="constant", cval=0)
    min_e = np.where(gt & (ea < e), ea, e)
    b = np.where(gt == 0, 2 - np.exp(np.log(0.5) / 5 * dst), np.ones_like(gt, dtype=np.float64))
    ew = min_e * b
    tpw = np.sum(gt) - np.sum(ew[gt == 1])
    fpw = np.sum(ew[gt == 0])
    r = 1 - np.mean(ew[gt])
    p = tpw / (eps + tpw + fpw)
    return (1 + beta) * r * p / (eps + r + beta * p)


class TotalMetric:
    """CalTotalMetric (utils/saliency_metric.py:306-341): per-image update, dataset means in show()"""

    def __init__(self, num: int, with_wfm: bool = True):
        self.num, self.idx, self.with_wfm = num, 0, with_wfm
        self.precision, self.recall = np.zeros((num, 255)), np.zeros((num, 255))
        self.meanf, self.mae, self.sm, self.em, self.wfm = (np.zeros(num) for _ in range(5))

    def update(self, pred, gt):
        i = self.idx
        self.mae[i] = mae(pred, gt)
        t = fmeasure_terms(pred, gt)
        if t is not None:
            self.precision[i], self.recall[i], self.meanf[i] = t
        self.sm[i] = smeasure(pred, gt)
        self.em[i] = emeasure(pred, gt)
        if self.with_wfm:
            self.wfm[i] = wfmeasure(pred, gt)
        self.idx += 1

    def show(self):
        assert self.num == self.idx
        p, r = self.precision.mean(axis=0), self.recall.mean(axis=0)
        f = 1.3 * p * r / (0.3 * p + r + 1e-8)
        return {"MaxF": f.max(), "MeanF": self.meanf.mean(), "WFM": self.wfm.mean() if self.with_wfm else None,
                "MAE": self.mae.mean(), "SM": self.sm.mean(), "EM": self.em.mean()}


def emulate_kernels(p8, g8):
    """what sod_saliency_head / sod_saliency_hist write for one image"""
    h, w = p8.shape
    gmax = int(g8.max())
    gb = (2 * g8.astype(np.int64) > gmax) & (gmax > 0)
    ys, xs = np.nonzero(gb)
    head = np.array([p8.min(), p8.max(), gmax, 0, gb.sum(), ys.sum(), xs.sum(), 0], dtype=np.int64)
    n_fg = max(int(gb.sum()), 1)
    cy, cx = int(round(int(ys.sum()) / n_fg)) + 1, int(round(int(xs.sum()) / n_fg)) + 1
    yy, xx = np.mgrid[0:h, 0:w]
    q = (yy >= cy) * 2 + (xx >= cx)
    k = p8.astype(np.int64) - int(p8.min())
    hist = np.zeros((4, 2, 256), dtype=np.int64)
    np.add.at(hist, (q.ravel(), gb.astype(np.int64).ravel(), k.ravel()), 1)
    return head, hist
