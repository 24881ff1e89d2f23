"""Gradient parity checks of the training entry points (sb_*_train_fwd / sb_*_bwd) against autograd through the CPU
oracle's restatement of the same reference spans.  TEST INFRASTRUCTURE (see kernel_cases.py): every function takes a
bound CDLL and a device; the GPU tests pass the sm_100a library, the emu tests the host-emulated test build.

Errors are reported relative to the largest entry of the reference tensor (gradients span several orders of magnitude).
"""
from __future__ import annotations

import ctypes

import torch

from oracle import tfgridnet_oracle as orc
from oracle.weights import make_state_dict, synthetic_mixture
from sound_bubble_b200 import _abi as abi
from sound_bubble_b200.packing import ModelConfig
from sound_bubble_b200.training import TrainGraph, differentiable_forward

from kernel_cases import _stream, _sync


def relerr(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _model(variant, kwargs, seed=0):
    ocfg = orc.OracleConfig.from_kwargs(variant, **kwargs)
    sd = make_state_dict(ocfg, seed)
    cfg = ModelConfig(variant=variant, **kwargs)
    return ocfg, sd, cfg


def _leaf_sd(sd):
    out = {}
    for k, v in sd.items():
        t = v.clone().float()
        if "_filters" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def check_path(lib, device, variant, kwargs, inter, B=1, T=2, block=1, seed=7):
    """One recurrent path (a9 / a10): forward value, dL/dx and every parameter gradient for L = sum(y * R)."""
    ocfg, sd, cfg = _model(variant, kwargs)
    lsd = _leaf_sd(sd)
    g = torch.Generator().manual_seed(seed)
    Fq, C, H = cfg.n_freqs, cfg.D, cfg.H
    x = torch.randn(B, T, Fq, C, generator=g).requires_grad_(True)
    R = torch.randn(B, T, Fq, C, generator=g)
    if inter:
        ref, _, _ = orc.inter_path(lsd, ocfg, block, x, torch.zeros(1, B * Fq, H), torch.zeros(1, B * Fq, H))
    else:
        ref = orc.intra_path(lsd, ocfg, block, x)
    (ref * R).sum().backward()

    P = {k: v.detach().to(device).contiguous() for k, v in sd.items()}
    tg = TrainGraph(lib, cfg)
    xd, Rd = x.detach().to(device), R.to(device)
    y = torch.full_like(xd, float("nan"))
    saved = torch.empty(int(lib.sb_path_train_saved_floats(B, T, Fq, C, H, int(inter))), device=device)
    pa = tg._path_args(P, block, inter, B, T)
    pa.x, pa.y, pa.saved = xd.data_ptr(), y.data_ptr(), saved.data_ptr()
    fwd = lib.sb_inter_lstm_train_fwd if inter else lib.sb_intra_lstm_train_fwd
    abi.check(lib, fwd(ctypes.byref(pa), _stream(device)), "path train fwd")
    _sync(device)
    errs = {"y": relerr(y, ref)}

    kind = "inter" if inter else "intra"
    b = f"tfgridnet.blocks.{block}."
    names = [b + kind + "_norm.norm.weight", b + kind + "_norm.norm.bias", b + kind + "_linear.weight", b + kind + "_linear.bias"]
    sfxs = ("",) if inter else ("", "_reverse")
    for sfx in sfxs:
        names += [b + kind + "_rnn." + n + sfx for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    G = {n: torch.zeros_like(P[n]) for n in names}
    pb = abi.PathBwdArgs()
    pb.f = pa
    gx = torch.full_like(xd, float("nan"))
    ws = torch.empty(int(lib.sb_path_bwd_workspace_floats(B, T, Fq, C, H, int(inter))), device=device)
    pb.gy, pb.gx, pb.ws = Rd.data_ptr(), gx.data_ptr(), ws.data_ptr()
    pb.g_ln_g, pb.g_ln_b, pb.g_lin_w, pb.g_lin_b = (G[n].data_ptr() for n in names[:4])
    for d, sfx in enumerate(sfxs):
        r = b + kind + "_rnn."
        pb.g_w_ih[d], pb.g_w_hh[d] = G[r + "weight_ih_l0" + sfx].data_ptr(), G[r + "weight_hh_l0" + sfx].data_ptr()
        pb.g_b_ih[d], pb.g_b_hh[d] = G[r + "bias_ih_l0" + sfx].data_ptr(), G[r + "bias_hh_l0" + sfx].data_ptr()
    bwd = lib.sb_inter_lstm_bwd if inter else lib.sb_intra_lstm_bwd
    abi.check(lib, bwd(ctypes.byref(pb), _stream(device)), "path bwd")
    _sync(device)
    errs["gx"] = relerr(gx, x.grad)
    for n in names:
        errs[n.split("blocks.%d." % block)[1]] = relerr(G[n], lsd[n].grad)
    return errs


def oracle_gradients(variant, kwargs, wave, dis, R, seed=0):
    """autograd through the oracle's whole path: L = sum(output * R) -> (output, {name: grad})."""
    ocfg, sd, cfg = _model(variant, kwargs, seed)
    lsd = _leaf_sd(sd)
    out, _ = orc.core_forward(lsd, ocfg, wave, dis, orc.init_state(ocfg, wave.shape[0]))
    (out * R).sum().backward()
    return out.detach(), {k: v.grad for k, v in lsd.items() if v.requires_grad}      # None = unused parameter


def check_net(lib, device, variant, kwargs, B=1, T=3, seed=11, golden=None):
    """The whole differentiable forward (SeparatorFunction) against the oracle: output and every parameter gradient."""
    ocfg, sd, cfg = _model(variant, kwargs)
    n = cfg.stft_chunk_size * T + cfg.n_fft - cfg.stft_chunk_size
    wave = synthetic_mixture(B, cfg.num_ch, n, seed=seed)
    dis = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])[torch.arange(B) % 3] if variant == "dis_embed" else None
    R = torch.randn(B, cfg.num_src, cfg.stft_chunk_size * T, generator=torch.Generator().manual_seed(seed + 1))
    ref_out, ref_g = oracle_gradients(variant, kwargs, wave, dis, R) if golden is None else golden

    named = {}
    for k, v in sd.items():
        t = v.detach().clone().float().to(device)
        if "_filters" not in k:
            t.requires_grad_(True)
        named[k] = t
    out = differentiable_forward(lib, cfg, named, wave.to(device), None if dis is None else dis.to(device))
    (out * R.to(device)).sum().backward()
    _sync(device)
    errs = {"output": relerr(out, ref_out)}
    for k, g in ref_g.items():
        if g is None:                       # autograd leaves unused parameters without a gradient; so does the node
            assert named[k].grad is None, k
            continue
        assert named[k].grad is not None, k
        errs[k.replace("tfgridnet.", "")] = relerr(named[k].grad, g)
    return errs


def check_attn_stage(lib, device, variant, kwargs, B=1, T=12, block=0, seed=5):
    """a11 under autograd at stage level (sb_attn_train_fwd / sb_attn_bwd through the C ABI) on the SAME input and output
    gradient as autograd through the oracle's attention_path (zero K / V history, as every training call has).  A stage
    check is robust where a whole-network one is not: four PReLUs sit in the attention unit, and with ~10^5 pre-activations
    one of them can land within rounding distance of the kink, where the two sides' 1e-6 different inputs pick different
    slopes and a whole weight gradient moves by ~1/sqrt(N) (observed: 3e-3 at 7 250 positions, kernels correct)."""
    ocfg, sd, cfg = _model(variant, kwargs)
    Fq, C, W = cfg.n_freqs, cfg.D, cfg.local_atten_len
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, Fq, C, generator=g)
    gy = torch.randn(B, T, Fq, C, generator=g)
    pre = "tfgridnet.blocks.%d." % block
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(pre + "attn")}
    xs = x.clone().requires_grad_(True)
    full = dict(sd)
    full.update(leaf)
    Kb = torch.zeros(B * cfg.L, W - 1, cfg.attn_E * Fq)
    Vb = torch.zeros(B * cfg.L, W - 1, (C // cfg.L) * Fq)
    y, newK, newV = orc.attention_path(full, ocfg, block, xs, Kb, Vb)
    (y * gy).sum().backward()

    tg = TrainGraph(lib, cfg)
    P = {k: v.clone().float().to(device) for k, v in sd.items()}
    xd, gx = x.to(device), gy.clone().to(device)
    nan = lambda n: torch.full((int(n),), float("nan"), device=device)      # uninitialised reads would surface
    aa = tg._attn_args(P, block, B, T)
    saved, yo = nan(lib.sb_attn_train_saved_floats(ctypes.byref(aa))), nan(B * T * Fq * C).view(B, T, Fq, C)
    aa.x, aa.y, aa.saved = xd.data_ptr(), yo.data_ptr(), saved.data_ptr()
    abi.check(lib, lib.sb_attn_train_fwd(ctypes.byref(aa), _stream(device)), "sb_attn_train_fwd")
    ab = abi.AttnBwdArgs()
    ab.f = tg._attn_args(P, block, B, T)
    ab.f.saved, ab.f.x = saved.data_ptr(), xd.data_ptr()
    wsa = nan(lib.sb_attn_bwd_workspace_floats(ctypes.byref(ab.f)))
    ab.gy, ab.gx, ab.ws = gx.data_ptr(), gx.data_ptr(), wsa.data_ptr()           # in place, as TrainGraph.backward calls it
    G = {}
    for field, mod in tg._ATTN:
        gp = getattr(ab, "g" + field)
        for f2, name in tg._ATTN_P:
            G[mod + name] = torch.zeros_like(P[pre + mod + name])
            setattr(gp, f2, G[mod + name].data_ptr())
    abi.check(lib, lib.sb_attn_bwd(ctypes.byref(ab), _stream(device)), "sb_attn_bwd")
    _sync(device)
    errs = {"y": relerr(yo, y), "gx": relerr(gx, xs.grad)}
    for k, v in G.items():
        errs[k] = relerr(v, leaf[pre + k].grad)
    return errs


def load_grad_fixture(name):
    import json
    import os

    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    meta = json.loads(str(z["meta"]))
    grads = {k[len("grad::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad::")}
    return meta, torch.from_numpy(z["mixture"]), torch.from_numpy(z["dis_embed"]), torch.from_numpy(z["output"]), grads


def loss_weights(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def check_golden_grads(lib, device, name):
    """Gradients of the UNMODIFIED reference (tests/golden/grad_*.npz, oracle/make_golden_grads.py) against this library.
    On a CUDA device the call goes through the drop-in ``Net`` module in train() mode, exactly as PLModule._step would
    (hl_module.py:303-330); with the host-emulated test build the same padding + autograd node are driven directly."""
    import torch.nn.functional as F
    meta, mix, dis, ref_out, ref_g = load_grad_fixture(name)
    variant, kw = meta["variant"], meta["kwargs"]
    ocfg, sd, cfg = _model(variant, kw, meta["seed"])
    R = loss_weights(ref_out.shape, meta["loss_seed"])
    if torch.device(device).type == "cuda":
        if variant == "dis_embed":
            from sound_bubble_b200.tfgridnet_realtime_clean_dis_embd3.net import Net
        else:
            from sound_bubble_b200.tfgridnet_realtime_clean_optim.net import Net
        net = Net(**kw)
        net.load_state_dict(sd, strict=True)
        net = net.to(device).train()
        res = net({"mixture": mix.to(device), "dis_embed": dis.to(device)})
        out = res["output"]
        (out * R.to(device)).sum().backward()
        got = {k: p.grad for k, p in net.named_parameters()}
    else:
        named = {k: (v.clone().float().requires_grad_("_filters" not in k)) for k, v in sd.items()}
        chunk = cfg.stft_chunk_size
        mod = (chunk - mix.shape[-1] % chunk) % chunk
        x = F.pad(F.pad(mix, (0, mod)), (0, cfg.stft_pad_size))
        out = differentiable_forward(lib, cfg, named, x, dis if variant == "dis_embed" else None)
        if mod:
            out = out[:, :, :-mod]
        (out * R).sum().backward()
        got = {k: v.grad for k, v in named.items() if v.requires_grad}
    _sync(device)
    errs = {"output": relerr(out, ref_out)}
    assert set(got) == set(ref_g), set(got) ^ set(ref_g)
    for k, g in ref_g.items():
        assert got[k] is not None, k
        errs[k.replace("tfgridnet.", "")] = relerr(got[k], g)
    return errs


def check_next_state(lib, device, variant, kwargs, B=2, T=3, seed=13):
    """The state a training call returns (same schema as init_buffers) against the oracle's after the same call."""
    from sound_bubble_b200.training import differentiable_forward_with_state
    ocfg, sd, cfg = _model(variant, kwargs)
    n = cfg.stft_chunk_size * T + cfg.n_fft - cfg.stft_chunk_size
    wave = synthetic_mixture(B, cfg.num_ch, n, seed=seed)
    dis = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])[torch.arange(B) % 3] if variant == "dis_embed" else None
    with torch.no_grad():
        ref_out, ref_state = orc.core_forward(sd, ocfg, wave, dis, orc.init_state(ocfg, B))
    named = {k: v.detach().clone().float().to(device).requires_grad_("_filters" not in k) for k, v in sd.items()}
    out, state = differentiable_forward_with_state(lib, cfg, named, wave.to(device), None if dis is None else dis.to(device))
    _sync(device)

    def flat(st, prefix=""):
        o = {}
        for k, v in st.items():
            o.update(flat(v, prefix + k + "::") if isinstance(v, dict) else {prefix + k: v})
        return o
    got, ref = flat(state), flat(ref_state)
    assert list(got) == list(ref), (list(got), list(ref))                      # same keys in the same order
    errs = {"output": relerr(out, ref_out)}
    for k in ref:
        assert tuple(got[k].shape) == tuple(ref[k].shape), k
        assert not got[k].requires_grad
        errs[k] = relerr(got[k], ref[k])
    return errs
