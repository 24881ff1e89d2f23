"""Streaming-state wire format against the reference's edge/flatbuf.py + edge/edge_utils.py layout."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle.cases import RPI, SYN
from sound_bubble_b200 import Net, NetOptim
from sound_bubble_b200 import state_io as sio

REF_FLATBUF = "/root/reference/edge/flatbuf.py"


def _state(kw=SYN, cls=Net, batch=2):
    st = cls(**kw).init_buffers(batch, "cpu")
    g = torch.Generator().manual_seed(3)
    names, bufs = sio.flatten_state(st)
    for b in bufs:
        b.copy_(torch.randn(b.shape, generator=g))
    return st


def test_names_follow_the_reference_order():
    names, bufs = sio.flatten_state(_state())
    expect = ["conv_buf", "deconv_buf"]
    for i in range(SYN["B"]):
        expect += [f"gridnet_bufs::buf{i}::c0", f"gridnet_bufs::buf{i}::h0"]
    expect += ["istft_buf"]
    assert names == expect
    assert [tuple(b.shape) for b in bufs[:2]] == [(2, 27, 2, 145), (2, 32, 2, 145)]
    names_attn, _ = sio.flatten_state(_state(dict(SYN, use_attn=True)))
    assert names_attn[2:6] == ["gridnet_bufs::buf0::K_buf", "gridnet_bufs::buf0::V_buf", "gridnet_bufs::buf0::c0",
                               "gridnet_bufs::buf0::h0"]


def test_round_trip_and_arena(tmp_path):
    st = _state(RPI, NetOptim, batch=3)
    names, bufs = sio.flatten_state(st, clone=True)
    back = sio.unflatten_state(names, bufs)
    n2, b2 = sio.flatten_state(back)
    assert n2 == names and all(torch.equal(a, b) for a, b in zip(bufs, b2))
    arena = sio.StateArena(st)
    n3, b3 = sio.flatten_state(arena.state)
    assert n3 == names and all(torch.equal(a, b) for a, b in zip(bufs, b3))
    assert all(v.data_ptr() % 256 == arena.flat.data_ptr() % 256 for v in arena.views)        # 256-byte aligned views
    host = arena.to_host()
    assert all(np.array_equal(host[n], b.numpy()) for n, b in zip(names, bufs))
    arena.flat.zero_()
    arena.load(st)
    assert all(torch.equal(a, b) for a, b in zip(bufs, sio.flatten_state(arena.state)[1]))
    with pytest.raises(KeyError):
        arena.load({"conv_buf": st["conv_buf"]})
    # the directory layout the ONNX tools exchange (edge/edge_utils.py:5-17)
    mix = torch.randn(3, 6, 288)
    sio.save_vectors(str(tmp_path), mix, st)
    lines = open(os.path.join(tmp_path, "input_names.txt")).read().split()
    assert lines == ["mixture"] + names
    mix2, st2 = sio.load_vectors(str(tmp_path))
    assert torch.equal(mix, mix2)
    assert all(torch.equal(a, b) for a, b in zip(bufs, sio.flatten_state(st2)[1]))


def test_bad_inputs():
    with pytest.raises(TypeError):
        sio.flatten_state({"a": 1})
    with pytest.raises(ValueError):
        sio.unflatten_state(["a", "a"], [torch.zeros(1), torch.zeros(1)])
    with pytest.raises(ValueError):
        sio.unflatten_state(["a", "a::b"], [torch.zeros(1), torch.zeros(1)])


@pytest.mark.skipif(not os.path.exists(REF_FLATBUF), reason="the reference tree is only mounted in the build container")
def test_same_as_the_reference_flatbuf():
    spec = importlib.util.spec_from_file_location("ref_flatbuf", REF_FLATBUF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for kw, cls in ((SYN, Net), (dict(SYN, use_attn=True), Net), (RPI, NetOptim)):
        st = _state(kw, cls)
        rn, rb = ref.flatten_state_buffers(st)
        n, b = sio.flatten_state(st)
        assert rn == n and all(torch.equal(x, y) for x, y in zip(rb, b))
        back_ref = ref.unflatten_state_buffers(rn, rb)
        back = sio.unflatten_state(n, b)
        assert ref.flatten_state_buffers(back_ref)[0] == sio.flatten_state(back)[0]
        assert all(torch.equal(x, y) for x, y in zip(ref.flatten_state_buffers(back_ref)[1], sio.flatten_state(back)[1]))
