"""Loader of the host-emulated TEST build (tests/emu/libsoundbubble_emu.so).  TEST INFRASTRUCTURE ONLY — nothing in
``sound_bubble_b200`` imports this; the product library is the sm_100a CUDA build and has no CPU path."""
import ctypes
import os
import subprocess

from sound_bubble_b200 import _abi as abi

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        if build:
            subprocess.run(["sh", os.path.join(HERE, "build_emu.sh")], check=True, capture_output=True)
        _lib = abi.bind(ctypes.CDLL(os.path.join(HERE, "libsoundbubble_emu.so")))
    return _lib
