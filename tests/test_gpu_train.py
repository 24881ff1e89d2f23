"""Training path on a B200 (run with `-m gpu`): gradients of the hand-written backward kernels, through the C ABI and
through the drop-in ``Net`` module in train() mode, against the gradients of the UNMODIFIED reference (fixtures) and
against autograd through the oracle; size-independent properties at larger sizes."""
import pytest
import torch

import train_cases as tc
from oracle.cases import GRAD_CASES, OPI, SYN
from oracle.weights import make_state_dict, radius_one_hot, synthetic_mixture
from oracle import tfgridnet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_TOL = 1e-4          # relative to the largest entry of each gradient tensor (fp32 both sides; atomics reorder the sums,
                         # sigmoid / tanh are ex2.approx + rcp.approx on the device)
C16 = dict(OPI, D=16)


@pytest.fixture(scope="module")
def lib():
    from sound_bubble_b200 import _lib
    return _lib.load()


def _ok(errs, tol=GRAD_TOL):
    assert all(v <= tol for v in errs.values()), {k: v for k, v in errs.items() if v > tol}


@pytest.mark.parametrize("inter", [False, True])
@pytest.mark.parametrize("variant,kw,B,T", [("dis_embed", SYN, 2, 5), ("optim", C16, 1, 3), ("dis_embed", SYN, 1, 31)])
def test_recurrent_path_gradients(lib, inter, variant, kw, B, T):
    _ok(tc.check_path(lib, DEV, variant, kw, inter, B=B, T=T))


@pytest.mark.parametrize("name", sorted(n for n in GRAD_CASES if not GRAD_CASES[n]["kwargs"].get("use_attn")))
def test_module_gradients_match_the_reference(lib, name):
    _ok(tc.check_golden_grads(lib, DEV, name))


def test_attention_gradients(lib):
    """a11 under autograd on the B200 (DE3:856-898): the reference's own gradients for a windowed-attention model through
    ``Net`` in train() mode, the attention unit alone at 7 250 positions with five windows' worth of frames, and a small
    whole network with more frames than the window."""
    _ok(tc.check_golden_grads(lib, DEV, "grad_syn_attn"))
    _ok(tc.check_attn_stage(lib, DEV, "dis_embed", dict(SYN, B=1, use_attn=True, local_atten_len=10), B=2, T=25))
    _ok(tc.check_attn_stage(lib, DEV, "optim", dict(C16, B=1, use_attn=True, local_atten_len=100), B=1, T=40))
    _ok(tc.check_net(lib, DEV, "dis_embed", dict(SYN, B=1, use_attn=True, local_atten_len=5), B=2, T=12))


def test_tfg_s_gradients_against_oracle_autograd(lib):
    """the benchmark architecture (6 blocks, FiLM on 5 of them), every parameter"""
    _ok(tc.check_net(lib, DEV, "dis_embed", SYN, B=3, T=6))


def test_long_clip_gradients_against_oracle_autograd(lib):
    """T >= 64 frames: the FiLM backward runs as eight frame slices that meet by atomics, the conv-in / deconv weight
    gradients walk thousands of positions per CTA with carried (b, t, f) counters across frame and utterance boundaries"""
    _ok(tc.check_net(lib, DEV, "dis_embed", dict(SYN, B=2), B=2, T=70))


def test_variants_against_oracle_autograd(lib):
    _ok(tc.check_net(lib, DEV, "dis_embed", dict(SYN, B=1, num_src=2, spectral_masking=True), B=2, T=4))
    _ok(tc.check_net(lib, DEV, "dis_embed", dict(SYN, B=2, merge_method="None", use_first_ln=False), B=1, T=9))
    _ok(tc.check_net(lib, DEV, "optim", dict(OPI, B=2), B=2, T=5))


def _net(kw=SYN, seed=0):
    from sound_bubble_b200 import Net
    net = Net(**kw)
    net.load_state_dict(make_state_dict(orc.OracleConfig.from_kwargs("dis_embed", **kw), seed), strict=True)
    return net.to(DEV)


def _grads(net, mix, dis, R):
    net.zero_grad(set_to_none=True)
    out = net({"mixture": mix, "dis_embed": dis})["output"]
    (out * R).sum().backward()
    return out.detach(), {k: p.grad.clone() for k, p in net.named_parameters()}


def test_gradients_add_over_utterances_at_one_second_clips():
    """size-independent property: for a sum loss, grad(batch of 4) = grad(first 2) + grad(last 2); 1 s clips (125 frames)"""
    net = _net().train()
    mix = synthetic_mixture(4, 6, 24000, seed=5).to(DEV)
    dis = radius_one_hot(4).to(DEV)
    R = torch.randn(4, 1, 24000, generator=torch.Generator().manual_seed(6)).to(DEV)
    out, g_all = _grads(net, mix, dis, R)
    out_a, g_a = _grads(net, mix[:2], dis[:2], R[:2])
    out_b, g_b = _grads(net, mix[2:], dis[2:], R[2:])
    assert float((out[:2] - out_a).abs().max()) <= 1e-6 and float((out[2:] - out_b).abs().max()) <= 1e-6   # utterances never mix
    for k in g_all:
        assert tc.relerr(g_a[k] + g_b[k], g_all[k]) <= 1e-4, k
    # and the inference kernels agree with the training forward
    with torch.no_grad():
        ref = net.eval()({"mixture": mix, "dis_embed": dis})["output"]
    assert float((ref - out).abs().max()) <= 3e-4


def test_mode_switches():
    net = _net(dict(SYN, B=2))
    mix, dis = synthetic_mixture(1, 6, 192 * 4, seed=1).to(DEV), radius_one_hot(1).to(DEV)
    out = net.train()({"mixture": mix, "dis_embed": dis})
    assert out["output"].requires_grad
    from conftest import flatten_state
    assert list(flatten_state(out["next_state"])) == list(flatten_state(net.init_buffers(1, DEV)))
    out = net.eval()({"mixture": mix, "dis_embed": dis})
    assert not out["output"].requires_grad and out["next_state"] is not None
    with torch.no_grad():
        assert not net.train()({"mixture": mix, "dis_embed": dis})["output"].requires_grad
    st = net.init_buffers(1, DEV)                      # carried state: the inference kernels, no autograd graph
    assert not net.train()({"mixture": mix, "dis_embed": dis}, st)["output"].requires_grad
    from sound_bubble_b200 import Net
    rpi = Net(**dict(SYN, B=1, use_attn=True, E=3)).to(DEV).train()  # L*E = 12: no backward kernels, forward-only, backward fails loudly
    y = rpi({"mixture": mix, "dis_embed": dis})["output"]
    with pytest.raises(RuntimeError):
        y.sum().backward()


def test_a_few_optimizer_steps_reduce_the_loss():
    """train_pt.py's inner loop in miniature: negative SNR loss, clip_grad_norm_(1), Adam (hl_module.py:321, 437-441)"""
    net = _net(dict(SYN, B=2)).train()
    mix = synthetic_mixture(2, 6, 192 * 20, seed=9).to(DEV)
    dis = radius_one_hot(2).to(DEV)
    target = mix[:, :1] * 0.5
    opt = torch.optim.Adam(net.parameters(), lr=2e-3)
    losses = []
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        est = net({"mixture": mix, "dis_embed": dis})["output"]
        loss = -(10 * torch.log10(target.pow(2).sum(-1) / ((est - target).pow(2).sum(-1) + 1e-8))).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        losses.append(float(loss))
    # the same loop through the oracle port on the CPU gives 16.543, 11.697, ..., 4.959
    assert abs(losses[0] - 16.5427) <= 2e-3 and losses[-1] < 7.0, losses


def test_gradients_with_programmatic_dependent_launch(lib):
    """bench.py runs the library with SB_OPT_PDL on: every training kernel must wait before it reads a predecessor's output"""
    from sound_bubble_b200 import _lib
    _lib.set_pdl(True)
    try:
        _ok(tc.check_golden_grads(lib, DEV, "grad_syn_b2"))
        _ok(tc.check_net(lib, DEV, "dis_embed", SYN, B=2, T=20))
    finally:
        _lib.set_pdl(False)


def test_first_version_lstm_training_kernels(lib):
    from sound_bubble_b200 import _abi as abi
    assert lib.sb_set_option(abi.SB_OPT_TRAIN_ONE_ROW, 1) == 0
    try:
        for inter in (False, True):
            _ok(tc.check_path(lib, DEV, "dis_embed", SYN, inter, B=2, T=7))
    finally:
        lib.sb_set_option(abi.SB_OPT_TRAIN_ONE_ROW, 0)


@pytest.mark.parametrize("on", [0, 1])
def test_packed_fma_variants_of_the_gemm_kernels(lib, on):
    from sound_bubble_b200 import _abi as abi
    assert lib.sb_set_option(abi.SB_OPT_TRAIN_FFMA2, on) == 0
    try:
        _ok(tc.check_golden_grads(lib, DEV, "grad_syn_b2"))
        _ok(tc.check_path(lib, DEV, "optim", C16, True, B=2, T=9))
    finally:
        lib.sb_set_option(abi.SB_OPT_TRAIN_FFMA2, 1)


@pytest.mark.parametrize("dt", ["linear1", "linear2", "conv4"])      # conv1 / conv2: LayerNorm over 1 - 2 values, gradients ~ 0
def test_every_distance_embedding_type(lib, dt):
    _ok(tc.check_net(lib, DEV, "dis_embed", dict(SYN, B=2, dis_type=dt), B=3, T=4))


def test_training_call_returns_the_next_state(lib):
    _ok(tc.check_next_state(lib, DEV, "dis_embed", dict(SYN, B=2), B=2, T=7))
    _ok(tc.check_next_state(lib, DEV, "optim", dict(OPI, B=1), B=1, T=1))


def test_pretrain_stage_three_optimizer_steps_reproduce_the_oracle_loss_curve():
    """syn_experiments/pretrain_stage.json as src/train_pt.py runs it (PLModule args from the fixture: TFG_S model, SNRLPLoss
    ('snr', neg_weight 100), Adam lr 1.2e-3, grad_clip 1), driven through train_dist.TrainModule on the B200 for three
    optimizer steps with train_epoch's call order, against the same three steps of oracle autograd on the CPU (same
    initial weights, same loss restatement, torch Adam + clip_grad_norm_).  One clip of the batch has a silent target, so
    both branches of the loss are on the path."""
    import json
    import os
    from sound_bubble_b200.losses import SNRLPLoss
    from sound_bubble_b200.train_dist import TrainModule
    exp = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "experiments.json")))["syn_experiments/pretrain_stage.json"]
    args = dict(exp["pl_module_args"], grad_clip=exp["grad_clip"])
    ocfg = orc.OracleConfig.from_kwargs("dis_embed", **args["model_params"])
    sd = make_state_dict(ocfg, 0)
    B, n = 3, 192 * 40
    mix = synthetic_mixture(B, 6, n, seed=21)
    tgt = 0.6 * mix[:, :1].clone()
    tgt[1] = 0.0
    dis = radius_one_hot(B)
    nspk = torch.tensor([1, 0, 2])

    # oracle: the reference's arithmetic under torch autograd
    leaf = {k: v.clone().requires_grad_("_filters" not in k) for k, v in sd.items()}
    params = [v for v in leaf.values() if v.requires_grad]
    opt = torch.optim.Adam(params, **args["optimizer_params"])
    loss_fn = SNRLPLoss(**args["loss_params"])
    want = []
    for _ in range(3):
        opt.zero_grad()
        est = orc.net_forward(leaf, ocfg, {"mixture": mix, "dis_embed": dis})["output"]
        loss = loss_fn(est=est, gt=tgt).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, args["grad_clip"])
        opt.step()
        want.append(float(loss.detach()))

    m = TrainModule(**args)
    m.model.load_state_dict(sd, strict=True)
    m.model.to(DEV)
    m.optimizer = torch.optim.Adam(m.model.parameters(), **args["optimizer_params"])      # after .to(), as PLModule.load_state re-creates it
    m.train()
    batch = ({"mixture": mix.to(DEV), "dis_embed": dis.to(DEV)}, {"target": tgt.to(DEV), "num_target_speakers": nspk})
    got = []
    for i in range(3):
        m.reset_grad()
        loss, nb = m.training_step(batch, i)
        loss.backward()
        m.backprop()
        got.append(float(loss.detach()))
    print("loss curve  ours %s  oracle %s" % (got, want))
    assert nb == B and m.get_avg_metric_at_epoch("train/loss") == pytest.approx(sum(got) / 3, rel=1e-5)
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-3 * max(1.0, abs(b)), (got, want)
    assert want[2] < want[0]                                           # and the curve goes down
