"""Import the UNMODIFIED reference modules from /root/reference (build container only).

Used only by ``make_golden.py`` (fixture generation) and by the optional
``tests/test_oracle_vs_reference.py`` cross-check, both of which skip when
``/root/reference`` is absent (it does not exist on the GPU box).  No reference source is
copied into the repository: the module source is read from where it lies, one line is
patched IN MEMORY (the hard-coded ``.to('cuda:0')`` of ``swin_transformer_v2.py:294``, which
breaks CPU execution), and it is exec'd against import shims for the two packages the image
does not have (``timm`` -> DropPath/to_2tuple/trunc_normal_; ``models.checkpoint`` -> stubs).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("B200SWIN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "swin_transformer_v2.py"))


class _DropPath(nn.Module):
    """timm.models.layers.DropPath semantics (per-sample stochastic depth, scale by keep)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _install_shims():
    if "timm.models.layers" not in sys.modules:
        timm = types.ModuleType("timm")
        tm = types.ModuleType("timm.models")
        tl = types.ModuleType("timm.models.layers")
        tl.DropPath = _DropPath
        tl.to_2tuple = _to_2tuple
        tl.trunc_normal_ = nn.init.trunc_normal_
        timm.models = tm
        tm.layers = tl
        sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    if "_refpkg_models" not in sys.modules:
        pkg = types.ModuleType("_refpkg_models")
        pkg.__path__ = []
        ck = types.ModuleType("_refpkg_models.checkpoint")
        ck.load_checkpoint_swin = lambda *a, **k: None
        ck.get_root_logger = lambda *a, **k: None
        sys.modules.update({"_refpkg_models": pkg, "_refpkg_models.checkpoint": ck})


_CACHE: dict = {}


def load_swin():
    """Return the reference ``models.swin_transformer_v2`` module object."""
    if "swin" in _CACHE:
        return _CACHE["swin"]
    _install_shims()
    path = os.path.join(REF_ROOT, "models", "swin_transformer_v2.py")
    with open(path, "r") as f:
        src = f.read()
    bad = ".to('cuda:0')"
    assert src.count(bad) == 1, "reference changed: expected exactly one hard-coded cuda:0"
    src = src.replace(bad, ".to(self.logit_scale.device)")
    mod = types.ModuleType("_refpkg_models.swin_transformer_v2")
    mod.__package__ = "_refpkg_models"
    mod.__file__ = path
    sys.modules[mod.__name__] = mod
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, path, "exec"), mod.__dict__)
    _CACHE["swin"] = mod
    return mod


def load_cnn_transformer():
    """Return the reference ``models.cnn_transformer`` module object (imports torchvision, which the image has)."""
    if "cnn" not in _CACHE:
        spec = importlib.util.spec_from_file_location("_ref_cnn_transformer",
                                                      os.path.join(REF_ROOT, "models", "cnn_transformer.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _CACHE["cnn"] = m
    return _CACHE["cnn"]


def load_criterion():
    spec = importlib.util.spec_from_file_location("_ref_criterion", os.path.join(REF_ROOT, "utils", "criterion.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_metrics():
    spec = importlib.util.spec_from_file_location("_ref_metrics", os.path.join(REF_ROOT, "utils", "metrics.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def quiet(fn, *a, **k):
    """Call ``fn`` with stdout silenced (the reference prints in constructors)."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
