"""Build libvad_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library is a plain C ABI)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["vad_api.cu", "vad_conv_umma.cu", "vad_model.cu"]
OUT = os.path.join(HERE, "libvad_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "vad_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, timeline: bool = False) -> str:
    """timeline=True compiles the per-tile clock stamps in (tools/timeline*.py need them; they slow the kernels)."""
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DVAD_TIMELINE"] if timeline else []) + \
          ["-o", OUT] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--timeline" in sys.argv, verbose="-v" in sys.argv,
                timeline="--timeline" in sys.argv))
