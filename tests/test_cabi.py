"""CPU checks of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/b200swin.h declares, the ctypes table covers them all, and the product package never
touches oracle/ or a CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-modal-monodepth-estimation_b200")
HEADER = os.path.join(ROOT, "include", "b200swin.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200swin_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    path = os.path.join(PKG, "lib", "libb200swin.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(path)


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ["b200swin_version", "b200swin_last_error", "b200swin_silog_fwd", "b200swin_silog_bwd",
                 "b200swin_window_gather", "b200swin_window_scatter", "b200swin_shift_mask", "b200swin_patch_merge", "b200swin_patchify",
                 "b200swin_cpb_fwd", "b200swin_cpb_bwd",
                 "b200swin_ln_fwd", "b200swin_ln_bwd"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/b200swin.h but not exported"
    lib.b200swin_version.restype = ctypes.c_int
    assert lib.b200swin_version() >= 100


def test_ctypes_table_matches_header():
    import b200swin._lib as L
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call -> usable on a CPU-only box
    lib.b200swin_last_error.restype = ctypes.c_char_p
    lib.b200swin_shift_mask.restype = ctypes.c_int
    rc = lib.b200swin_shift_mask(None, 8, 8, 4, 2, None)
    assert rc == -1 and b"null" in lib.b200swin_last_error()
    lib.b200swin_window_gather.restype = ctypes.c_int
    rc = lib.b200swin_window_gather(ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 8, 8, 8, 4, 4, 4, None)
    assert rc == -1 and b"shift" in lib.b200swin_last_error()


def test_product_never_imports_oracle_or_reference():
    bad = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b|/root/reference", re.M)
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f"{f} references oracle/tests/reference"


def test_ops_refuse_cpu_tensors():
    import torch
    import b200swin
    with pytest.raises(RuntimeError, match="CUDA"):
        b200swin.SiLogLoss()(torch.ones(4), torch.ones(4))
