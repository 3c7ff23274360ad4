"""timm.models.layers: DropPath (per-sample stochastic depth, scale by keep), to_2tuple, trunc_normal_."""
import torch.nn as nn


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


trunc_normal_ = nn.init.trunc_normal_
