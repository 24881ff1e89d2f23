"""Training path without a GPU: (1) autograd through the oracle is pinned to the gradients of the UNMODIFIED reference
(tests/golden/grad_*.npz); (2) the backward kernels are checked on the host-emulated TEST build of the same .cu sources
(tests/emu) against those fixtures and against oracle autograd.  The product library has no CPU path."""
import pytest
import torch

import train_cases as tc
from oracle import tfgridnet_oracle as orc
from oracle.cases import GRAD_CASES, OPI, RPI, SYN

GRAD_TOL = 2e-5          # relative to the largest entry of each gradient tensor


@pytest.fixture(scope="module")
def lib():
    from emu.emu_lib import load
    return load()


def _ok(errs, tol=GRAD_TOL):
    assert all(v <= tol for v in errs.values()), {k: v for k, v in errs.items() if v > tol}


@pytest.mark.parametrize("name", sorted(GRAD_CASES))
def test_oracle_autograd_matches_reference_gradients(name):
    meta, mix, dis, ref_out, ref_g = tc.load_grad_fixture(name)
    ocfg, sd, cfg = tc._model(meta["variant"], meta["kwargs"], meta["seed"])
    leaf = {k: (v.clone().requires_grad_(True) if k in ref_g else v) for k, v in sd.items()}
    out = orc.net_forward(leaf, ocfg, {"mixture": mix, "dis_embed": dis})["output"]
    (out * tc.loss_weights(out.shape, meta["loss_seed"])).sum().backward()
    assert tc.relerr(out, ref_out) <= 2e-6
    for k, g in ref_g.items():
        assert tc.relerr(leaf[k].grad, g) <= GRAD_TOL, k


def test_fixture_covers_every_parameter():
    meta, _, _, _, ref_g = tc.load_grad_fixture("grad_syn_b2")
    shapes = orc.param_shapes(orc.OracleConfig.from_kwargs(meta["variant"], **meta["kwargs"]))
    assert set(ref_g) == {k for k in shapes if "_filters" not in k}


@pytest.mark.parametrize("inter", [False, True])
def test_recurrent_path_gradients(lib, inter):
    _ok(tc.check_path(lib, "cpu", "dis_embed", SYN, inter, B=1, T=2))
    _ok(tc.check_path(lib, "cpu", "optim", dict(OPI, D=16), inter, B=1, T=3))      # ragged last CTA (3 / 435 rows)


def test_whole_path_gradients_against_the_reference_fixtures(lib):
    _ok(tc.check_golden_grads(lib, "cpu", "grad_opi_d16"))
    _ok(tc.check_golden_grads(lib, "cpu", "grad_syn_b2"))
    _ok(tc.check_golden_grads(lib, "cpu", "grad_tfg_s"))           # the benchmark architecture, all 149 tensors


def test_conv_lstm_gradients_against_the_reference_fixtures(lib):
    """a9': the Raspberry-Pi model (OPT, k = 5, D = 16, output_padding tail) and the DE3 variant (k = 4, D = 32, pad-and-crop)"""
    _ok(tc.check_golden_grads(lib, "cpu", "grad_rpi"))
    _ok(tc.check_golden_grads(lib, "cpu", "grad_syn_convlstm"))


def test_variants_against_oracle_autograd(lib):
    """two sources + spectral masking, and the plain front-end (no spatial features, no first LayerNorm)"""
    _ok(tc.check_net(lib, "cpu", "dis_embed", dict(SYN, B=1, num_src=2, spectral_masking=True), B=1, T=2))
    _ok(tc.check_net(lib, "cpu", "dis_embed", dict(SYN, B=2, merge_method="None", use_first_ln=False), B=1, T=2))


def test_untrainable_configurations_raise():
    from sound_bubble_b200.packing import ModelConfig
    from sound_bubble_b200.training import check_trainable
    check_trainable(ModelConfig(variant="dis_embed", **SYN))
    check_trainable(ModelConfig(variant="dis_embed", **dict(SYN, conv_lstm=True)))
    check_trainable(ModelConfig(variant="dis_embed", **dict(SYN, dis_type="linear2")))
    check_trainable(ModelConfig(variant="dis_embed", **dict(SYN, use_attn=True)))           # L * E = 8: backward kernels exist
    for kw in (dict(SYN, use_attn=True, E=3),):                                                # L * E = 12: none
        with pytest.raises(NotImplementedError):
            check_trainable(ModelConfig(variant="dis_embed", **kw))


def test_first_version_lstm_training_kernels(lib):
    """SB_OPT_TRAIN_ONE_ROW: the one-gate-row / one-column-per-thread kernels stay selectable and correct"""
    from sound_bubble_b200 import _abi as abi
    assert lib.sb_set_option(abi.SB_OPT_TRAIN_ONE_ROW, 1) == 0
    try:
        _ok(tc.check_path(lib, "cpu", "dis_embed", SYN, False, B=1, T=2))
        _ok(tc.check_path(lib, "cpu", "dis_embed", SYN, True, B=1, T=2))
    finally:
        lib.sb_set_option(abi.SB_OPT_TRAIN_ONE_ROW, 0)


def test_other_geometries(lib):
    """the constructor defaults' STFT sizes (n_fft 280, F = 141), 2-3 microphones, single frames, ragged row counts"""
    kw = dict(stft_chunk_size=160, stft_pad_size=120, num_ch=2, D=16, B=2, H=64, L=4, E=2, use_attn=False, lookahead=True,
              use_first_ln=False, merge_method="None", conv_lstm=False, dis_type="conv3")
    _ok(tc.check_net(lib, "cpu", "dis_embed", kw, B=3, T=1))
    _ok(tc.check_net(lib, "cpu", "dis_embed", dict(kw, conv_lstm=True, use_first_ln=True, merge_method="early_cat"), B=1, T=2))
    _ok(tc.check_net(lib, "cpu", "optim", dict(RPI, lstm_down=4, B=1), B=5, T=1))
    _ok(tc.check_net(lib, "cpu", "dis_embed", dict(SYN, num_ch=3, B=1), B=1, T=2))


def test_packed_fma_variants_of_the_gemm_kernels(lib):
    """SB_OPT_TRAIN_FFMA2: same arithmetic (two fmaf per packed instruction), other instruction stream"""
    from sound_bubble_b200 import _abi as abi
    assert lib.sb_set_option(abi.SB_OPT_TRAIN_FFMA2, 0) == 0          # the default is 1: this is the unpacked variant
    try:
        _ok(tc.check_path(lib, "cpu", "dis_embed", SYN, False, B=1, T=2))
        _ok(tc.check_path(lib, "cpu", "optim", dict(OPI, D=16), True, B=1, T=3))
    finally:
        lib.sb_set_option(abi.SB_OPT_TRAIN_FFMA2, 1)


def test_every_distance_embedding_type(lib):
    """Dis_Embed_Linear (LayerNorm over the whole F*Din vector, DE3:114-147) and the other Dis_Embed_Conv widths"""
    for dt in ("linear1", "linear2", "conv4"):       # conv1 / conv2 normalise 1 - 2 values: embedding gradients vanish
        _ok(tc.check_net(lib, "cpu", "dis_embed", dict(SYN, B=2, dis_type=dt), B=2, T=2))


def test_training_call_returns_the_next_state(lib):
    _ok(tc.check_next_state(lib, "cpu", "dis_embed", dict(SYN, B=2), B=2, T=3))
    _ok(tc.check_next_state(lib, "cpu", "optim", dict(RPI, B=1), B=1, T=1))          # single frame: conv_buf keeps a zero frame


def test_attention_backward(lib):
    """a11 under autograd: the reference's gradients for a windowed-attention model, the K / V history a training call hands
    back, the D = 16 / E = 4 instantiations against oracle autograd, and the unit alone with more frames than the window"""
    _ok(tc.check_golden_grads(lib, "cpu", "grad_syn_attn"))
    _ok(tc.check_next_state(lib, "cpu", "dis_embed", dict(SYN, B=1, use_attn=True, local_atten_len=3), B=1, T=5))
    _ok(tc.check_net(lib, "cpu", "optim", dict(RPI, B=1, conv_lstm=False, use_attn=True, local_atten_len=3), B=1, T=4))
    _ok(tc.check_net(lib, "cpu", "dis_embed", dict(SYN, B=1, E=4, use_attn=True, local_atten_len=9), B=1, T=3))
    _ok(tc.check_attn_stage(lib, "cpu", "dis_embed", dict(SYN, B=1, use_attn=True, local_atten_len=4), B=2, T=9))
