"""SiLog loss and depth metrics, CPU restatement (numpy float64 + torch).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np
import torch


def silog_np(pred: np.ndarray, target: np.ndarray, lambd: float = 0.5) -> float:
    """SiLogLoss.forward (utils/criterion.py:15-21) in float64:
    m = target>0; d = log(target[m]) - log(pred[m]); sqrt(mean(d^2) - lambd*mean(d)^2).
    The mean runs over ALL valid pixels of the whole batch; no valid pixel -> NaN."""
    p = np.asarray(pred, dtype=np.float64).reshape(-1)
    t = np.asarray(target, dtype=np.float64).reshape(-1)
    m = t > 0
    if not m.any():
        return float("nan")
    d = np.log(t[m]) - np.log(p[m])
    return float(np.sqrt((d * d).mean() - lambd * d.mean() ** 2))


def silog_grad_np(pred: np.ndarray, target: np.ndarray, lambd: float = 0.5, gout: float = 1.0) -> np.ndarray:
    """Closed-form dL/dpred (checked against autograd):
    -(d_i - lambd*mean(d)) / (n * L * pred_i) on valid pixels, 0 elsewhere."""
    p = np.asarray(pred, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    m = t > 0
    d = np.zeros_like(p)
    d[m] = np.log(t[m]) - np.log(p[m])
    n = m.sum()
    mean = d[m].mean()
    L = np.sqrt((d[m] ** 2).mean() - lambd * mean ** 2)
    g = np.zeros_like(p)
    g[m] = -(d[m] - lambd * mean) / (n * L * p[m]) * gout
    return g


def silog_torch(pred: torch.Tensor, target: torch.Tensor, lambd: float = 0.5) -> torch.Tensor:
    """Same loss through torch ops (differentiable), following utils/criterion.py:15-21."""
    m = (target > 0).detach()
    d = torch.log(target[m]) - torch.log(pred[m])
    return torch.sqrt((d ** 2).mean() - lambd * d.mean() ** 2)


def eval_depth_np(pred: np.ndarray, target: np.ndarray) -> dict:
    """eval_depth (utils/metrics.py:9-32) on 1-D arrays of valid pixels, float64.
    d1/d2/d3 divide by len(thresh) = number of pixels."""
    p = np.asarray(pred, dtype=np.float64).reshape(-1)
    t = np.asarray(target, dtype=np.float64).reshape(-1)
    assert p.shape == t.shape
    thresh = np.maximum(t / p, p / t)
    n = float(len(thresh))
    diff = p - t
    dlog = np.log(p) - np.log(t)
    return dict(
        d1=float((thresh < 1.25).sum() / n),
        d2=float((thresh < 1.25 ** 2).sum() / n),
        d3=float((thresh < 1.25 ** 3).sum() / n),
        abs_rel=float(np.mean(np.abs(diff) / t)),
        sq_rel=float(np.mean(diff ** 2 / t)),
        rmse=float(np.sqrt(np.mean(diff ** 2))),
        rmse_log=float(np.sqrt(np.mean(dlog ** 2))),
        log10=float(np.mean(np.abs(np.log10(p) - np.log10(t)))),
        silog=float(np.sqrt(np.mean(dlog ** 2) - 0.5 * np.mean(dlog) ** 2)),
    )
