"""The reference arm: the unmodified reference staged under ``baseline/_ref`` (see ``stage_reference.py``).

``load()`` returns the reference's modules imported from that copy (``models.*`` / ``utils.*`` resolve there, ``timm``
and ``mmcv`` resolve to the shims).  Only ``bench.py`` (reference arms), ``tests/`` and ``tools/`` import this package;
the product (``b200swin``) never does.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SHIMS = os.path.join(HERE, "shims")


def available() -> bool:
    return os.path.exists(os.path.join(REF, ".staged"))


def why_unavailable() -> str:
    return (f"{REF} is not staged: run `python baseline/stage_reference.py` where /root/reference exists "
            "(the build container); the directory is git-ignored but travels with gpurun")


_mods = None


def load() -> types.SimpleNamespace:
    """Import the staged reference.  Names: swin, decoder_v1, decoder_v2, model, cnn_transformer, optimizer, criterion,
    metrics, util."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(why_unavailable())
    for p in (SHIMS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("models", "utils"):              # an unrelated top-level module of that name must not shadow the copy
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(REF):
            del sys.modules[name]
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints on import / construction
        ns.swin = importlib.import_module("models.swin_transformer_v2")
        ns.decoder_v1 = importlib.import_module("models.decoder_v1")
        ns.decoder_v2 = importlib.import_module("models.decoder_v2")
        ns.model = importlib.import_module("models.model")
        ns.cnn_transformer = importlib.import_module("models.cnn_transformer")
        ns.optimizer = importlib.import_module("models.optimizer")
        ns.criterion = importlib.import_module("utils.criterion")
        ns.metrics = importlib.import_module("utils.metrics")
        ns.util = importlib.import_module("utils.util")
    _mods = ns
    return ns


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
