"""ORACLE (test infrastructure only): BCE-with-logits + CEL, value and gradient, numpy fp64.

Follows
* `torch.nn.BCEWithLogitsLoss(reduction=user_config["reduction"])` — reference train.py:203, config.py:61;
* `CEL.forward` — reference loss/CEL.py:15-20 (eps 1e-6 at :13): sums over the WHOLE batch tensor;
* `get_total_loss` — reference utils/pipeline_ops.py:37-42: total = sum of the listed losses, each also
  reported as f"{item:.5f}".
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x: np.ndarray) -> np.ndarray:
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def bce_with_logits(x, t, reduction: str = "mean") -> float:
    """max(x,0) - x*t + log(1+exp(-|x|)), mean or sum over all elements."""
    x = np.asarray(x, np.float64).ravel()
    t = np.asarray(t, np.float64).ravel()
    per = np.maximum(x, 0.0) - x * t + np.log1p(np.exp(-np.abs(x)))
    if reduction == "mean":
        return float(per.mean())
    if reduction == "sum":
        return float(per.sum())
    raise ValueError(reduction)


def cel(x, t, eps: float = 1e-6) -> float:
    """loss/CEL.py:15-20: ((p - p t).sum() + (t - p t).sum()) / (p.sum() + t.sum() + eps)."""
    x = np.asarray(x, np.float64).ravel()
    t = np.asarray(t, np.float64).ravel()
    p = _sigmoid(x)
    inter = p * t
    num = (p - inter).sum() + (t - inter).sum()
    den = p.sum() + t.sum()
    return float(num / (den + eps))


def bce_cel_fwd_bwd(x, t, reduction: str = "mean", w_bce: float = 1.0, w_cel: float = 1.0,
                    grad_scale: float = 1.0, eps: float = 1e-6):
    """Value and d(total)/d(logits) of total = w_bce*BCE + w_cel*CEL, scaled by grad_scale.

    Returns dict(bce, cel, total, sum_p, sum_t, sum_pt, grad[fp64, shape of x]).
    Gradient derivation: dBCE/dx = (p - t)/N (mean) or (p - t) (sum);
    CEL = num/(den+eps), num = Σp + Σt - 2Σpt, den = Σp + Σt  ⇒
    dCEL/dp_i = ((1-2 t_i)(den+eps) - num)/(den+eps)^2,  dp/dx = p(1-p).
    """
    xs = np.asarray(x, np.float64)
    shape = xs.shape
    xs = xs.ravel()
    ts = np.asarray(t, np.float64).ravel()
    n = xs.size
    p = _sigmoid(xs)
    per = np.maximum(xs, 0.0) - xs * ts + np.log1p(np.exp(-np.abs(xs)))
    bce = per.mean() if reduction == "mean" else per.sum()
    sp, st, spt = p.sum(), ts.sum(), (p * ts).sum()
    num, den = sp + st - 2.0 * spt, sp + st
    celv = num / (den + eps)
    g_bce = (p - ts) * ((1.0 / n) if reduction == "mean" else 1.0)
    g_cel = ((1.0 - 2.0 * ts) * (den + eps) - num) / (den + eps) ** 2 * p * (1.0 - p)
    grad = grad_scale * (w_bce * g_bce + w_cel * g_cel)
    return dict(bce=float(bce), cel=float(celv), total=float(w_bce * bce + w_cel * celv),
                sum_p=float(sp), sum_t=float(st), sum_pt=float(spt), grad=grad.reshape(shape))


def get_total_loss_strings(values) -> list[str]:
    """The per-loss report strings of utils/pipeline_ops.py:40."""
    return [f"{v:.5f}" for v in values]
