"""Build libb2y.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m bs_yolo_b200.build            (or __graft_entry__.build())

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun snapshots."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb2y.so")
SOURCES = ["plan.cu", "conv_simt.cu", "conv_tc.cu", "conv_pair.cu", "misc_ops.cu", "pmsfa_dw.cu", "msca.cu", "dw3_tile.cu", "attention.cu", "preprocess.cu", "decode.cu", "nms.cu", "val_match.cu", "extras.cu", "seg.cu"]
EXTRA = {"nms.cu": ["-fmad=false"], "preprocess.cu": ["-fmad=false"], "val_match.cu": ["-fmad=false"], "extras.cu": ["-fmad=false"]}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libb2y.so cannot be built (there is no CPU fallback)")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "dw_regs.cuh"), os.path.join(HERE, "..", "include", "b2y.h")]
    exe = nvcc()

    def compile_one(src: str) -> str:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [exe, *ARCH, *COMMON, *EXTRA.get(src, []), "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [exe, *ARCH, "-shared", "-o", LIB, *objs, "-Xlinker", "--exclude-libs,ALL"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
