"""mmcv.runner: the optimizer-constructor registry models/optimizer.py:9-11 builds on.

``build_optimizer(model, cfg)``: pops ``constructor`` (default ``DefaultOptimizerConstructor``) and ``paramwise_cfg``
from ``cfg``, instantiates the registered constructor with the remaining optimizer config and calls it on the model --
the mmcv 1.x behaviour the reference relies on (train.py:113-115)."""
import copy

import torch


class _Registry(dict):
    def register_module(self, name=None):
        def deco(cls):
            self[name or cls.__name__] = cls
            return cls
        return deco


OPTIMIZER_BUILDERS = _Registry()


def get_dist_info():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


@OPTIMIZER_BUILDERS.register_module()
class DefaultOptimizerConstructor:
    def __init__(self, optimizer_cfg, paramwise_cfg=None):
        self.optimizer_cfg = optimizer_cfg
        self.paramwise_cfg = {} if paramwise_cfg is None else paramwise_cfg
        self.base_lr = optimizer_cfg.get("lr", None)
        self.base_wd = optimizer_cfg.get("weight_decay", None)

    def add_params(self, params, module, prefix="", is_dcn_module=None):
        params.append({"params": [p for p in module.parameters() if p.requires_grad]})

    def __call__(self, model):
        if hasattr(model, "module"):
            model = model.module
        cfg = copy.deepcopy(self.optimizer_cfg)
        typ = cfg.pop("type")
        params = []
        self.add_params(params, model)
        return getattr(torch.optim, typ)(params, **cfg)


def build_optimizer(model, cfg):
    cfg = copy.deepcopy(cfg)
    constructor = cfg.pop("constructor", "DefaultOptimizerConstructor")
    paramwise_cfg = cfg.pop("paramwise_cfg", None)
    return OPTIMIZER_BUILDERS[constructor](cfg, paramwise_cfg)(model)
