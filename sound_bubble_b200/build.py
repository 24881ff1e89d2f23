"""Builds libsoundbubble_sm100a.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree).

    python -m sound_bubble_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(PKG, "csrc", "build")
LIB = os.path.join(PKG, "libsoundbubble_sm100a.so")
SOURCES = ["sb_host.cu", "sb_lstm.cu", "sb_lstm_tc.cu", "sb_lstm_tcp.cu", "sb_frontend.cu", "sb_frontend_tc.cu", "sb_backend.cu", "sb_convlstm.cu", "sb_attn.cu", "sb_attn_tc.cu", "sb_net.cu", "sb_pipe.cu", "sb_prepare.cu", "sb_train.cu", "sb_train_tc.cu"]
HEADERS = [os.path.join(CSRC, "sb_common.cuh"), os.path.join(CSRC, "sb_lstm.cuh"), os.path.join(INCLUDE, "soundbubble.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-I" + INCLUDE, "-I" + CSRC]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsoundbubble_sm100a.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = find_nvcc()
    os.makedirs(BUILD, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str):
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
