"""In-tree build of libstellar_rhmc.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m hmc_stellar_toy_model_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libstellar_rhmc.so")
SOURCES = [os.path.join(CSRC, "stellar_rhmc.cu")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "stellar_rhmc.h"))
    return deps


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found; cannot build libstellar_rhmc.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT + ".tmp"] + SOURCES
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (see %s):\n%s" % (log, proc.stdout[-4000:]))
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
