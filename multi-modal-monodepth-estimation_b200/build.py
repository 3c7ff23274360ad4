"""Build libb200swin.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python multi-modal-monodepth-estimation_b200/build.py [--force] [--verbose]

Objects go to <pkg>/build/ (git-ignored), the library to <pkg>/lib/libb200swin.so (git-ignored but
shipped to the GPU box by gpurun).  nvcc cross-compiles without a GPU.  The library links cudart
statically and resolves the one driver symbol it needs (cuTensorMapEncodeTiled) at run time through
cudaGetDriverEntryPoint, so it loads on a CPU-only box for the symbol-export test.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libb200swin.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the b200swin library cannot be built")
    return exe


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_library(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "b200swin.h"))
    srcs = sources()
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def compile_one(pair):
        src, obj = pair
        if not force and _newer(obj, [src] + headers):
            return None
        extra = os.environ.get("B200SWIN_EXTRA_NVCC", "").split()          # e.g. -DB200SWIN_TRACE for debug builds
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return (src, r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, zip(srcs, objs)))
    rebuilt = [r for r in results if r is not None]
    if verbose or ptxas_info:
        for src, log in rebuilt:
            print(f"[nvcc] {os.path.basename(src)}")
            if log.strip():
                print(log)
    if rebuilt or force or not _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                        ptxas_info="--ptxas" in sys.argv)
    print(lib)
