"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/soundbubble.h
declares, and its struct layouts match the ctypes mirror.  No kernel is launched (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from sound_bubble_b200 import _abi as abi
from sound_bubble_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    from sound_bubble_b200.build import build
    build()
    return _lib.load()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "soundbubble.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(abi.PROTOTYPES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.sb_version() == abi.SB_VERSION


def test_struct_layouts_match(lib):
    for which, st in abi.ABI_STRUCTS.items():
        assert lib.sb_abi_sizeof(which) == ctypes.sizeof(st), st.__name__
    assert lib.sb_abi_sizeof(99) == -1


def test_argument_errors_are_reported_without_a_gpu(lib):
    assert lib.sb_net_forward(None, None, None) == -1
    assert b"null" in lib.sb_last_error_string()
    a = abi.InterArgs()
    assert lib.sb_inter_lstm_fwd(ctypes.byref(a), None) == -1
    assert lib.sb_set_option(12345, 1) == -1
    assert lib.sb_set_option(abi.SB_OPT_PDL, 0) == 0
    assert lib.sb_set_option(abi.SB_OPT_ATTN_TC, 1) == 0


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: a forward pass on CPU tensors must fail loudly, not route through the oracle."""
    from oracle.cases import SYN
    from sound_bubble_b200 import Net, SoundBubbleError
    m = Net(**SYN).eval()
    x = {"mixture": torch.zeros(1, 6, 192 * 2 + 96), "dis_embed": torch.tensor([[0., 0., 1.]])}
    with pytest.raises(SoundBubbleError):
        m(x, pad=False)


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sound_bubble_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "emu_lib" not in text and "libsoundbubble_emu" not in text, f


def test_integration_doc_shows_the_current_lstm_dir_layout():
    """INTEGRATION.md's reference-side ctypes stub must name the fields of sb_lstm_dir in header order."""
    hdr = open(os.path.join(ROOT, "include", "soundbubble.h")).read()
    body = hdr[hdr.index("typedef struct sb_lstm_dir"):hdr.index("} sb_lstm_dir;")]
    fields = re.findall(r"const float\*\s+(\w+);", body)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = doc[doc.index("class sb_lstm_dir"):doc.index("class sb_inter_args")]
    assert re.findall(r'"(\w+)"', stub) == fields == [n for n, _ in abi.LstmDir._fields_]
