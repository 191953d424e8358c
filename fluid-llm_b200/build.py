"""Build libfluidgrid.so in-tree with nvcc for sm_100a (B200).  No other architecture, no JIT.

    python fluid-llm_b200/build.py [--force] [--verbose]

The .so lands next to this file (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfluidgrid.so")
SOURCES = ["fl_api.cu", "fl_locate.cu", "fl_interp.cu", "fl_tiled.cu", "fl_ring.cu", "fl_dynamic.cu", "fl_patch.cu", "fl_stats.cu", "fl_embed.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "--use_fast_math=false", "-fmad=true", "-Xcompiler", "-fvisibility=default"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    deps = sources() + [os.path.join(CSRC, "fl_common.cuh"), os.path.join(CSRC, "fl_geom.cuh"), os.path.join(CSRC, "fl_interp.cuh"), os.path.join(HERE, "..", "include", "fluidgrid.h"),
                        os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [_nvcc(), "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
