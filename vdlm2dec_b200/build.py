"""Build libvdl2gpu.so (sm_100a only) in-tree with nvcc.  `python -m vdlm2dec_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvdl2gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr"]
SOURCES = ["vdl2_kernel.cu", "vdl2_link.cu", "vdl2_avlc.cu", "vdl2_host.cu", "vdl2_multi.cu"]
HEADERS = ["vdl2_kernel.h", "vdl2_link.h", "vdl2_common.h", "vdl2_demod.cuh", "vdl2_avlc.cuh", "vdl2_tables.h", "vdl2_mma_tables.h", os.path.join("..", "..", "include", "vdl2gpu.h")]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: str | None = None) -> str:
    """defines/out build an experimental variant next to the product library (tools/ab_probe.py)."""
    lib = LIB if out is None else os.path.join(HERE, out)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(lib, deps):
        return lib
    objs = []
    tag = "" if out is None else "." + os.path.splitext(out)[0]
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", tag + ".o"))
        cmd = [NVCC, *ARCH, *FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(o)
    cmd = [NVCC, *ARCH, "-shared", "-o", lib, *objs]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
