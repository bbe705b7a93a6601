"""Build ``libprb200.so`` in-tree with nvcc for sm_100a (no JIT cache: the built .so travels with the repo snapshot).

    python -m bo_pr_b200.build            # or: python -c "import __graft_entry__ as g; g.build()"
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# PRB_LIB_NAME / PRB_NVCC_EXTRA: build a tuning variant next to the default library (experiments only)
OUT = os.path.join(HERE, "_lib", os.environ.get("PRB_LIB_NAME", "libprb200.so"))
SOURCES = ["prb_gemm_tma.cu", "prb_small.cu", "prb_kernels.cu", "prb_finalize.cu", "prb_chol.cu", "prb_api.cu"]
HEADERS = ["prb_internal.cuh", "prb_device.cuh", "prb_kernels.cuh", "prb_ptx.cuh", os.path.join(ROOT, "include", "prb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wno-format-truncation", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libprb200.so cannot be built (bo_pr_b200 has no CPU fallback)")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    cmd += os.environ.get("PRB_NVCC_EXTRA", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libprb200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
