"""The host-side mirror of the reference module API: constructor kwargs, checkpoint keys, state schema."""
import inspect

import pytest
import torch

from oracle.cases import CASES, RPI, SYN
from oracle.tfgridnet_oracle import OracleConfig, init_state as oracle_init_state, param_shapes
from oracle.weights import make_state_dict
from sound_bubble_b200 import ModelConfig, Net, NetOptim
from sound_bubble_b200.engine import init_state
from sound_bubble_b200.filterbank import stft_filters
from sound_bubble_b200.packing import PackedWeights

# Net.__init__ signatures of the reference (DE3/net.py:21-26, OPT/net.py:21-26)
DE3_KWARGS = dict(stft_chunk_size=160, stft_pad_size=120, stft_back_pad=0, num_ch=2, D=64, B=6, I=1, J=1, L=0, H=128,
                  use_attn=False, lookahead=True, local_atten_len=100, E=4, chunk_causal=False, num_src=1,
                  spectral_masking=False, use_first_ln=False, merge_method="None", directional=False, conv_lstm=True,
                  fb_type='stft', dis_type="conv3")
OPT_KWARGS = dict(stft_chunk_size=160, stft_pad_size=120, stft_back_pad=0, num_ch=2, D=64, B=6, I=1, J=1, L=0, H=128,
                  use_attn=False, lookahead=True, local_atten_len=100, E=4, chunk_causal=False, num_src=1,
                  spectral_masking=False, use_first_ln=False, merge_method="None", directional=False, conv_lstm=True,
                  lstm_down=5, fb_type='stft')


def _defaults(cls):
    sig = inspect.signature(cls.__init__)
    return {k: v.default for k, v in sig.parameters.items() if k != "self"}


def test_constructor_signatures_match_reference():
    assert _defaults(Net) == DE3_KWARGS
    assert list(_defaults(Net)) == list(DE3_KWARGS)
    assert _defaults(NetOptim) == OPT_KWARGS
    with pytest.raises(TypeError):
        Net(not_a_kwarg=1)
    assert list(inspect.signature(Net.forward).parameters) == ["self", "inputs", "input_state", "pad"]
    assert list(inspect.signature(Net.predict).parameters) == ["self", "x", "dis_embed", "input_state", "pad"]
    assert list(inspect.signature(NetOptim.predict).parameters) == ["self", "x", "input_state", "pad"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dict_layout_matches_reference_checkpoint(name):
    case = CASES[name]
    cls = Net if case["variant"] == "dis_embed" else NetOptim
    m = cls(**case["kwargs"])
    ocfg = OracleConfig.from_kwargs(case["variant"], **case["kwargs"])
    want = param_shapes(ocfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    # asteroid_filterbanks.STFTFB also registers its analysis window (`torch_window`); nothing on the path reads it
    windows = {k: got.pop(k) for k in list(got) if k.endswith("filterbank.torch_window")}
    assert windows == {"tfgridnet.enc.filterbank.torch_window": (ocfg.n_fft,), "tfgridnet.dec.filterbank.torch_window": (ocfg.n_fft,)}
    assert list(got) == list(want)
    assert got == want
    sd = make_state_dict(ocfg, 0)
    m.load_state_dict(sd, strict=True)                      # a checkpoint written without torch_window
    with_window = dict(sd)
    for k in windows:
        with_window[k] = torch.hann_window(ocfg.n_fft, periodic=True, dtype=torch.float64).sqrt()
    m.load_state_dict(with_window, strict=True)             # ... and one written with it (float64 in some releases)
    assert m.state_dict()["tfgridnet.enc.filterbank.torch_window"].dtype == torch.float32
    assert torch.allclose(m.state_dict()["tfgridnet.enc.filterbank.torch_window"], with_window[list(windows)[0]].float(), atol=1e-6)


def test_unsupported_sizes_are_rejected_in_the_constructor():
    """The reference's own constructor defaults (D=64, H=128; DE3/net.py:21-26) are outside what the kernels are built
    for: that must be a clear error at construction, not a shape error at the first forward."""
    for cls in (Net, NetOptim):
        with pytest.raises(NotImplementedError, match="H=64"):
            cls()
    for bad in (dict(D=64), dict(D=24), dict(H=128), dict(num_ch=9), dict(stft_back_pad=32)):
        with pytest.raises(NotImplementedError):
            Net(**dict(SYN, **bad))


def test_weights_are_found_on_data_parallel_replicas():
    """PLModule wraps the model in nn.DataParallel when use_dp=true (hl_module.py:34-35); replicas hold their weights as
    plain attributes, so state_dict() on them is nearly empty.  The drop-in gathers weights by name instead."""
    m = Net(**SYN)
    names = list(m.state_dict().keys())
    # what torch.nn.parallel.replicate does, on one device
    mods, reps = list(m.modules()), [x._replicate_for_data_parallel() for x in m.modules()]
    idx = {mod: i for i, mod in enumerate(mods)}
    for i, mod in enumerate(mods):
        r = reps[i]
        for key, child in mod._modules.items():
            r._modules[key] = reps[idx[child]] if child is not None else None
        for key, p in mod._parameters.items():
            setattr(r, key, p.detach().clone().requires_grad_(p.requires_grad) if p is not None else None)
        for key, b in mod._buffers.items():
            r._buffers[key] = b
    rep = reps[0]
    assert len(rep.state_dict()) < len(names)               # the situation the walk exists for
    got = rep._named_weights()
    assert list(got) == names
    assert all(torch.equal(got[k], v) for k, v in m.state_dict().items())
    assert rep._wants_grad(None) == m._wants_grad(None)


def test_filterbank_buffer_is_bit_identical_to_the_oracle_basis():
    from oracle.tfgridnet_oracle import stft_basis
    assert torch.equal(stft_filters(288, 192), stft_basis(288, 192))


def test_state_schema_matches_reference():
    for variant, kw in (("dis_embed", SYN), ("optim", RPI), ("dis_embed", dict(SYN, use_attn=True))):
        a = init_state(ModelConfig(variant=variant, **kw), 3, "cpu")
        b = oracle_init_state(OracleConfig.from_kwargs(variant, **kw), 3)

        def walk(x, y, path=""):
            assert list(x) == list(y), path
            for k in x:
                if isinstance(x[k], dict):
                    walk(x[k], y[k], path + k + "::")
                else:
                    assert x[k].shape == y[k].shape and x[k].dtype == y[k].dtype, path + k
        walk(a, b)


def test_packing_layouts():
    """w_tile / w_lane orders documented in include/soundbubble.h (sb_lstm_dir)."""
    ocfg = OracleConfig.from_kwargs("dis_embed", **SYN)
    sd = make_state_dict(ocfg, 0)
    pk = PackedWeights(sd, ModelConfig(variant="dis_embed", **SYN), "cpu")
    H, C = 64, 32
    K = C + H
    w = torch.cat([sd["tfgridnet.blocks.2.inter_rnn.weight_ih_l0"], sd["tfgridnet.blocks.2.inter_rnn.weight_hh_l0"]], 1)
    bias = sd["tfgridnet.blocks.2.inter_rnn.bias_ih_l0"] + sd["tfgridnet.blocks.2.inter_rnn.bias_hh_l0"]
    o = pk.offsets["b2.inter.w_tile"]
    w_tile = pk.flat[o:o + K * 4 * H].view(K, 4 * H)
    o = pk.offsets["b2.inter.w_lane"]
    w_lane = pk.flat[o:o + K * 4 * H].view(K // 4, 4 * H, 4)
    o = pk.offsets["b2.inter.b_lane"]
    b_lane = pk.flat[o:o + 4 * H]
    for g, u, k in ((0, 0, 0), (1, 5, 17), (2, 63, 95), (3, 30, 40)):
        col = (g // 2) * 2 * H + 4 * (u // 2) + 2 * (g % 2) + u % 2
        assert w_tile[k, col] == w[g * H + u, k]
        assert w_lane[k // 4, 4 * u + g, k % 4] == w[g * H + u, k]
        assert b_lane[4 * u + g] == bias[g * H + u]
    o = pk.offsets["conv_w_pack"]
    wp = pk.flat[o:o + 3 * 27 * 3 * 32].view(3, 27, 3, 32)
    assert wp[2, 11, 1, 7] == sd["tfgridnet.conv.0.weight"][7, 11, 2, 1]
    assert pk.flat.data_ptr() % 16 == 0 and all(v % 64 == 0 for v in pk.offsets.values())
