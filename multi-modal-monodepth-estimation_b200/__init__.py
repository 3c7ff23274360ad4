"""b200swin - B200 (sm_100a) drop-in for the Swin-V2 shifted-window attention block and the
SiLog depth loss of junnyfilm/multi-modal-monodepth-estimation.

Import as ``b200swin`` (the repository root holds a shim that maps that name onto this
directory, whose on-disk name is fixed by the build contract).  The package mirrors the
reference's module layout for the path it replaces:

    b200swin.swin_transformer_v2   <->  models/swin_transformer_v2.py
    b200swin.criterion             <->  utils/criterion.py
    b200swin.cnn_transformer       <->  the global-attention encoder layer of models/cnn_transformer.py

Everything executes through libb200swin.so (hand-written CUDA behind the C-ABI of
include/b200swin.h); there is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .criterion import SiLogLoss  # noqa: F401
