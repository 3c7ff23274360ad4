"""Drop-in for the reference's ``utils/criterion.py``: ``SiLogLoss`` backed by the sm_100a kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class SiLogLoss(nn.Module):
    """Scale-invariant log loss over ``target > 0`` (reference: utils/criterion.py:10-21).

    Same constructor and ``forward(pred, target) -> 0-dim tensor`` as the reference.  One fused
    masked reduction on the GPU (no boolean-index compaction, no host sync) with a closed-form
    backward; float32 statistics.  All-invalid targets give NaN like the reference.
    """

    def __init__(self, lambd: float = 0.5):
        super().__init__()
        self.lambd = lambd

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return ops.silog_loss(pred, target, self.lambd)
