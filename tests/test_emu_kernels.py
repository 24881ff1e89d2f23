"""Kernel logic checked on the CPU through the host-emulated TEST build of the same .cu sources (tests/emu): indexing,
layouts and barrier structure of the SIMT kernels against the oracle / golden fixtures before they go to a B200.
This build is test infrastructure; the product library has no CPU path (see tests/test_abi.py)."""
import pytest

import kernel_cases as kc
import parity_cases as pc
from oracle.cases import OPI, RPI, SYN
from sound_bubble_b200 import _abi as abi

TOL = 2e-5


@pytest.fixture(scope="module")
def lib():
    from emu.emu_lib import load
    return load()


def _ok(errs, tol=TOL):
    assert all(v <= tol for v in errs.values()), errs


def test_frontend_backend_kernels(lib):
    _ok(kc.check_stft_features(lib, "cpu", "dis_embed", dict(SYN, spectral_masking=True), B=1, T=9, with_spec=True))
    _ok(kc.check_stft_features(lib, "cpu", "dis_embed", dict(SYN, directional=True), B=1, T=1))
    _ok(kc.check_conv_in(lib, "cpu", "dis_embed", SYN, B=2, T=5))
    _ok(kc.check_conv_in(lib, "cpu", "optim", RPI, B=1, T=1))
    _ok(kc.check_film(lib, "cpu", "dis_embed", SYN, B=4))
    _ok(kc.check_backend(lib, "cpu", "dis_embed", SYN, B=1, T=11, with_mask=True))
    _ok(kc.check_backend(lib, "cpu", "optim", RPI, B=2, T=1))
    _ok(kc.check_backend(lib, "cpu", "dis_embed", dict(SYN, num_src=2), B=1, T=3, with_mask=True))
    _ok(kc.check_backend(lib, "cpu", "dis_embed", dict(SYN, num_src=2), B=1, T=10))


@pytest.mark.parametrize("algo", [abi.SB_ALGO_TILE, abi.SB_ALGO_LANE1, abi.SB_ALGO_LANE4, abi.SB_ALGO_WS, abi.SB_ALGO_WS2])
def test_lstm_kernels(lib, algo):
    _ok(kc.check_inter(lib, "cpu", "dis_embed", SYN, algo, B=1, T=3, alias_state=True))
    _ok(kc.check_intra(lib, "cpu", "dis_embed", SYN, algo, B=1, T=2))
    _ok(kc.check_intra(lib, "cpu", "optim", dict(OPI, D=16), algo, B=1, T=2))


def test_two_sequences_per_cta_with_a_ragged_last_cta(lib):
    """SB_ALGO_WS2 pairs rows (2i, 2i+1): an odd number of rows leaves the last CTA with one real sequence."""
    _ok(kc.check_intra(lib, "cpu", "dis_embed", SYN, abi.SB_ALGO_WS2, B=1, T=3))


def test_streaming_attention_path(lib):
    """T == 1: scores / weighted sum streamed over the K / V history, which moves up one row on the way."""
    _ok(kc.check_attn(lib, "cpu", "dis_embed", dict(SYN, use_attn=True, local_atten_len=7), B=2, T=1))
    _ok(kc.check_attn(lib, "cpu", "optim", dict(RPI, use_attn=True, local_atten_len=10, conv_lstm=False), B=1, T=1, block=1))


def test_whole_path_golden_plain(lib):
    pc.assert_parity(pc.run_golden(lib, "cpu", "syn_plain"))


def test_summed_directions_are_refused_where_the_tensor_core_kernel_cannot_run(lib):
    """y_bwd == y_fwd (both intra directions summed into one buffer) exists only on lstm_tcr_kernel: the host-emulated
    build has no tensor-core kernels, so the query answers 0, the call fails with SB_E_UNSUPP instead of racing two
    directions into one buffer, and sb_net_forward keeps the two-buffer path (the whole-path goldens above)."""
    with pytest.raises(AssertionError):
        kc.check_intra(lib, "cpu", "dis_embed", SYN, abi.SB_ALGO_AUTO, B=1, T=2, block=1, summed=True)
    import ctypes
    import torch
    x = torch.zeros(1, 2, 145, 32)
    y = torch.zeros_like(x)
    a = abi.IntraArgs()
    a.x, a.y_fwd, a.y_bwd = x.data_ptr(), y.data_ptr(), y.data_ptr()
    a.B, a.T, a.F, a.C, a.H, a.algo = 1, 2, 145, 32, 64, abi.SB_ALGO_AUTO
    assert lib.sb_intra_sum_supported(ctypes.byref(a)) == 0
    assert lib.sb_intra_lstm_fwd(ctypes.byref(a), None) == -2          # SB_E_UNSUPP (include/soundbubble.h)
