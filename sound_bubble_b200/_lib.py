"""Loader of the product library libsoundbubble_sm100a.so (hand-written sm_100a CUDA kernels behind a C ABI).

There is NO CPU path and no fallback: if the library has not been built, or no CUDA device is present when a forward
pass is requested, this raises.  Build with ``python -m sound_bubble_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import _abi as abi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsoundbubble_sm100a.so")
_lock = threading.Lock()
_cdll = None


def load():
    """dlopen + bind prototypes + verify the ABI (works without a GPU; no kernel is launched)."""
    global _cdll
    with _lock:
        if _cdll is None:
            if not os.path.exists(LIB_PATH):
                raise abi.SoundBubbleError(
                    "%s is missing: build it with `python -m sound_bubble_b200.build` "
                    "(the separator has no CPU or PyTorch fallback)" % LIB_PATH)
            _cdll = abi.bind(ctypes.CDLL(LIB_PATH))
        return _cdll


def require_cuda(t):
    if not t.is_cuda:
        raise abi.SoundBubbleError(
            "sound_bubble_b200 runs on CUDA tensors only (got a %s tensor); move the model and its inputs to a B200 "
            "— there is no CPU path" % t.device.type)


def launch_count() -> int:
    return int(load().sb_launch_count())


def set_pdl(enabled: bool):
    lib = load()
    abi.check(lib, lib.sb_set_option(abi.SB_OPT_PDL, int(bool(enabled))), "sb_set_option")
