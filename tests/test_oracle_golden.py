"""The CPU oracle must reproduce the outputs of the unmodified reference stored in tests/golden."""
import pytest
import torch

from conftest import Golden, flatten_state, golden_names
from oracle.tfgridnet_oracle import (OracleConfig, init_state, net_forward, param_shapes, rms, si_sdr, stft_basis,
                                     stft_basis_closed_form, streaming_forward)
from oracle.weights import make_state_dict, state_dict_digest

TOL = 2e-5      # max-abs; the oracle uses the same aten ops as the reference, only attention reorders a sum


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_output(name):
    g = Golden(name)
    cfg = OracleConfig.from_kwargs(g.variant, **g.kwargs)
    sd = make_state_dict(cfg, g.meta["seed"])
    assert state_dict_digest(sd) == g.meta["weights_digest"], "deterministic weights drifted (torch RNG change?)"
    with torch.no_grad():
        r = net_forward(sd, cfg, g.inputs(), None, pad=g.pad)
    assert r["output"].shape == g.output.shape
    assert (r["output"] - g.output).abs().max().item() <= TOL
    got = flatten_state(r["next_state"])
    assert set(got) == set(g.state)
    for k, v in g.state.items():
        assert got[k].shape == v.shape, k
        assert (got[k] - v).abs().max().item() <= TOL, k
    if g.mixture2 is not None:
        with torch.no_grad():
            r2 = net_forward(sd, cfg, {"mixture": g.mixture2, "dis_embed": g.dis_embed}, r["next_state"], pad=False)
        assert (r2["output"] - g.output2).abs().max().item() <= TOL


def test_streaming_equals_offline():
    """edge/causal_infer.py:49-86 (atol 1e-3 there); chunk-by-chunk with carried state == one offline call."""
    g = Golden("syn_nopad")
    cfg = OracleConfig.from_kwargs(g.variant, **g.kwargs)
    sd = make_state_dict(cfg, 0)
    with torch.no_grad():
        y = streaming_forward(sd, cfg, g.mixture, g.dis_embed)
    assert (y - g.output).abs().max().item() <= TOL


def test_prefix_causality():
    """OPT/net.py:94-140 self-check: the output on a prefix equals the prefix of the output (atol 1e-2 there)."""
    g = Golden("rpi_offline")
    cfg = OracleConfig.from_kwargs(g.variant, **g.kwargs)
    sd = make_state_dict(cfg, 0)
    x = g.mixture[..., : 192 * 8 + 96]
    with torch.no_grad():
        full = net_forward(sd, cfg, {"mixture": x}, None, pad=False)["output"]
        part = net_forward(sd, cfg, {"mixture": x[..., : 192 * 3 + 96]}, None, pad=False)["output"]
    assert torch.allclose(full[..., : 192 * 3], part, atol=1e-6)


def test_basis_closed_form():
    a, b = stft_basis(288, 192), stft_basis_closed_form(288, 192)
    assert a.shape == (290, 1, 288)
    assert (a - b).abs().max().item() < 1e-7
    # rows are rfft(x * sqrt(hann)) / 10.3923 with DC / Nyquist real rows / sqrt(2)
    x = torch.randn(288, dtype=torch.float64)
    w = torch.hann_window(288, periodic=True, dtype=torch.float64).sqrt()
    spec = torch.fft.rfft(x * w) / (0.5 * (288 * 288 / 192) ** 0.5)
    got = a[:, 0].double() @ x
    ref = torch.cat([spec.real, spec.imag])
    ref[0] /= 2 ** 0.5
    ref[144] /= 2 ** 0.5
    assert (got - ref).abs().max().item() < 1e-5


def test_param_inventory_counts():
    from oracle.cases import RPI, SYN
    n = lambda cfg: sum(torch.Size(s).numel() for k, s in param_shapes(cfg).items() if not k.endswith("_filters"))
    assert n(OracleConfig.from_kwargs("dis_embed", **SYN)) == 501398        # SURVEY §2a / BASELINE.md
    assert n(OracleConfig.from_kwargs("optim", **RPI)) == 231125


def test_si_sdr_definition():
    t = torch.randn(2, 1000)
    assert torch.all(si_sdr(3.0 * t, t) > 60)
    noisy = t + 0.1 * torch.randn(2, 1000)
    assert torch.all((si_sdr(noisy, t) - 20).abs() < 1.5)
    assert rms(torch.ones(4)) == 1.0


def test_reference_copy_in_oracle_ref_equals_the_port():
    """oracle/_ref (oracle/build_ref.py: byte-for-byte copy of the reference's model files, present where the recipe has
    run) through the reference's own streaming protocol == the oracle port's whole-clip call; and the headline checker
    reports a perfect score for the reference's own output."""
    from oracle import ref_runner
    from oracle.headline import compare_with_oracle
    from oracle.cases import SYN
    from oracle.weights import radius_one_hot, synthetic_mixture
    if not ref_runner.available():
        pytest.skip("oracle/_ref has not been built (python oracle/build_ref.py in the build container)")
    cfg = OracleConfig.from_kwargs("dis_embed", **SYN)
    sd = make_state_dict(cfg, 0)
    net = ref_runner.reference_net(SYN, sd)
    mix = synthetic_mixture(3, 6, 192 * 12 + 96)
    mix[..., 192 * 12:] = 0.0                                 # the look-ahead of the last chunk = the offline call's zero pad
    dis = radius_one_hot(3)
    _, y = ref_runner.streaming_sample(net, mix, dis, 192, 96, 11, warm=1)
    r = compare_with_oracle(sd, SYN, mix[..., : 192 * 12], dis, y, [0, 2], target=mix[:, 0, : 192 * 12])
    assert y.shape == (3, 1, 192 * 12)
    assert r["rms"] <= 1e-6 and r["si_sdr_delta_db"] <= 1e-3 and r["ok"], r
