"""Stage-by-stage parity checks of the C-ABI entry points against the CPU oracle.  TEST INFRASTRUCTURE.

Every function takes a bound CDLL and a torch device, runs ONE entry point of include/soundbubble.h on seeded inputs
and returns the max-abs error against the oracle's restatement of the same reference span.  `tests/test_gpu_*.py`
call them with the sm_100a library on cuda:0; `tests/test_emu_kernels.py` with the host-emulated test build.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn.functional as F

from oracle import tfgridnet_oracle as orc
from oracle.weights import make_state_dict, synthetic_mixture
from sound_bubble_b200 import _abi as abi
from sound_bubble_b200.packing import ModelConfig, PackedWeights


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream if torch.device(device).type == "cuda" else 0


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def make_model(variant, kwargs, device, seed=0):
    ocfg = orc.OracleConfig.from_kwargs(variant, **kwargs)
    sd = make_state_dict(ocfg, seed)
    cfg = ModelConfig(variant=variant, **kwargs)
    packed = PackedWeights(sd, cfg, device)
    return ocfg, sd, cfg, packed


def maxerr(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def check_stft_features(lib, device, variant, kwargs, B=2, T=5, seed=3, with_spec=False):
    ocfg, sd, cfg, pk = make_model(variant, kwargs, device)
    n = cfg.stft_chunk_size * T + cfg.n_fft - cfg.stft_chunk_size
    wave = synthetic_mixture(B, cfg.num_ch, n, seed=seed)
    re, im = orc.stft_frames(wave, sd["tfgridnet.enc.filterbank._filters"], cfg.stft_chunk_size)
    chans = [re, im]
    if cfg.merge_method == "early_cat":
        chans.append(orc.spatial_features(re, im, cfg.directional))
    ref = torch.cat(chans, dim=1).permute(0, 3, 2, 1).contiguous()            # [B, T, F, Cin]
    a = abi.StftArgs()
    wd = wave.to(device)
    feats = torch.full((B, T, cfg.n_freqs, cfg.conv_in_ch), float("nan"), device=device)
    spec = torch.full((B, T, cfg.num_src, 2 * cfg.n_freqs), float("nan"), device=device) if with_spec else None
    a.wave, a.filt, a.feats = wd.data_ptr(), pk.ptr("enc_filt"), feats.data_ptr()
    a.spec = spec.data_ptr() if with_spec else None
    a.B, a.M, a.n_samples, a.T = B, cfg.num_ch, n, T
    a.n_fft, a.stride, a.F = cfg.n_fft, cfg.stft_chunk_size, cfg.n_freqs
    a.feat_mode, a.Cin, a.n_src = pk.desc.feat_mode, cfg.conv_in_ch, cfg.num_src
    abi.check(lib, lib.sb_stft_features_fwd(ctypes.byref(a), _stream(device)), "sb_stft_features_fwd")
    _sync(device)
    errs = {"feats": maxerr(feats, ref)}
    if with_spec:
        ref_spec = torch.cat([re, im], dim=2)[:, : cfg.num_src].permute(0, 3, 1, 2)   # [B, T, S, 2F]
        errs["spec"] = maxerr(spec, ref_spec)
    return errs


def check_conv_in(lib, device, variant, kwargs, B=2, T=5, seed=4):
    ocfg, sd, cfg, pk = make_model(variant, kwargs, device)
    g = torch.Generator().manual_seed(seed)
    feats_tf = torch.randn(B, cfg.conv_in_ch, T, cfg.n_freqs, generator=g)
    conv_buf = torch.randn(B, cfg.conv_in_ch, 2, cfg.n_freqs, generator=g)
    ref_x, ref_buf = orc.conv_in(sd, ocfg, feats_tf, conv_buf)
    a = abi.ConvInArgs()
    feats = feats_tf.permute(0, 2, 3, 1).contiguous().to(device)
    cb_in = conv_buf.to(device)
    cb_out = torch.full_like(cb_in, float("nan"))
    x = torch.full((B, T, cfg.n_freqs, cfg.D), float("nan"), device=device)
    a.feats, a.conv_buf_in, a.conv_buf_out = feats.data_ptr(), cb_in.data_ptr(), cb_out.data_ptr()
    a.w_pack, a.bias, a.ln_g, a.ln_b = pk.ptr("conv_w_pack"), pk.ptr("conv_bias"), pk.ptr("conv_ln_g"), pk.ptr("conv_ln_b")
    a.x = x.data_ptr()
    a.B, a.T, a.F, a.Cin, a.C = B, T, cfg.n_freqs, cfg.conv_in_ch, cfg.D
    abi.check(lib, lib.sb_conv_in_fwd(ctypes.byref(a), _stream(device)), "sb_conv_in_fwd")
    _sync(device)
    return {"x": maxerr(x, ref_x), "conv_buf": maxerr(cb_out, ref_buf)}


def run_film(lib, device, cfg, pk, dis):
    B = dis.shape[0]
    film = torch.full((cfg.B - 1, 2, B, cfg.n_freqs, cfg.D), float("nan"), device=device)
    a = abi.FilmArgs()
    dd = dis.to(device)
    a.dis, a.emb_w, a.emb_ln_g, a.emb_ln_b = dd.data_ptr(), pk.ptr("emb_w"), pk.ptr("emb_ln_g"), pk.ptr("emb_ln_b")
    a.w_w, a.w_b, a.b_w, a.b_b = pk.ptr("film_w_w"), pk.ptr("film_w_b"), pk.ptr("film_b_w"), pk.ptr("film_b_b")
    a.film = film.data_ptr()
    a.B, a.F, a.C, a.Din, a.n_layers, a.emb_mode = B, cfg.n_freqs, cfg.D, cfg.film_in, cfg.B - 1, pk.desc.emb_mode
    abi.check(lib, lib.sb_film_params_fwd(ctypes.byref(a), _stream(device)), "sb_film_params_fwd")
    _sync(device)
    return film


def check_film(lib, device, variant, kwargs, B=3):
    ocfg, sd, cfg, pk = make_model(variant, kwargs, device)
    dis = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.], [0.3, 0.2, 0.5]])[:B]
    film = run_film(lib, device, cfg, pk, dis)
    emb = orc.distance_embedding(sd, ocfg, dis)
    err = 0.0
    for j in range(cfg.B - 1):
        p = f"tfgridnet.embeds.{j}."
        w = F.conv1d(emb, sd[p + "weight.weight"], sd[p + "weight.bias"]).transpose(1, 2)     # [B, F, C]
        b = F.conv1d(emb, sd[p + "bias.weight"], sd[p + "bias.bias"]).transpose(1, 2)
        err = max(err, maxerr(film[j, 0], w), maxerr(film[j, 1], b))
    return {"film": err}


def check_intra(lib, device, variant, kwargs, algo, B=2, T=3, block=1, seed=5, use_film=True, summed=False):
    ocfg, sd, cfg, pk = make_model(variant, kwargs, device)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, cfg.n_freqs, cfg.D, generator=g)
    xin = x
    fs = fb = None
    if use_film and cfg.variant == "dis_embed" and block > 0:
        dis = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])[torch.arange(B) % 3]
        film = run_film(lib, device, cfg, pk, dis)
        fs, fb = film[block - 1, 0].contiguous(), film[block - 1, 1].contiguous()
        xin = orc.film(sd, block - 1, x, orc.distance_embedding(sd, ocfg, dis))
    ref = orc.intra_path(sd, ocfg, block, xin)
    xd = x.to(device)
    if cfg.conv_lstm:
        y = torch.full_like(xd, float("nan"))
        J = cfg.lstm_steps
        ws = torch.empty(B * T * J * (cfg.D + 2 * cfg.H), device=device)
        a = abi.IntraConvArgs()
        a.x, a.y, a.ws = xd.data_ptr(), y.data_ptr(), ws.data_ptr()
        a.film_scale = fs.data_ptr() if fs is not None else None
        a.film_shift = fb.data_ptr() if fb is not None else None
        for f in ("cl_conv_w", "cl_conv_b", "cl_prelu", "cl_deconv_w", "cl_deconv_b"):
            setattr(a, f[3:], pk.ptr(f"b{block}.{f}"))
        a.dir[0], a.dir[1] = pk.lstm_dir(block, "intra0"), pk.lstm_dir(block, "intra1")
        a.B, a.T, a.F, a.C, a.H = B, T, cfg.n_freqs, cfg.D, cfg.H
        a.down, a.tail_mode, a.algo = cfg.lstm_down, pk.desc.tail_mode, algo
        abi.check(lib, lib.sb_intra_convlstm_fwd(ctypes.byref(a), _stream(device)), "sb_intra_convlstm_fwd")
        _sync(device)
        return {"y": maxerr(y, ref)}
    yf = torch.full_like(xd, float("nan"))
    yb = torch.full_like(xd, float("nan"))
    a = abi.IntraArgs()
    a.x, a.y_fwd, a.y_bwd = xd.data_ptr(), yf.data_ptr(), yb.data_ptr()
    a.film_scale = fs.data_ptr() if fs is not None else None
    a.film_shift = fb.data_ptr() if fb is not None else None
    a.dir[0], a.dir[1] = pk.lstm_dir(block, "intra0"), pk.lstm_dir(block, "intra1")
    a.B, a.T, a.F, a.C, a.H, a.algo = B, T, cfg.n_freqs, cfg.D, cfg.H, algo
    abi.check(lib, lib.sb_intra_lstm_fwd(ctypes.byref(a), _stream(device)), "sb_intra_lstm_fwd")
    _sync(device)
    if summed:
        # y_bwd == y_fwd: both directions add into one buffer; bitwise the sum the consumer would have formed on load
        ys = torch.full_like(xd, float("nan"))
        a.y_fwd = a.y_bwd = ys.data_ptr()
        assert lib.sb_intra_sum_supported(ctypes.byref(a)) == 1
        abi.check(lib, lib.sb_intra_lstm_fwd(ctypes.byref(a), _stream(device)), "sb_intra_lstm_fwd")
        _sync(device)
        assert torch.equal(ys, yf + yb)
    return {"y": maxerr(yf + yb, ref)}


def check_inter(lib, device, variant, kwargs, algo, B=2, T=4, block=0, seed=6, two_inputs=True, alias_state=False):
    ocfg, sd, cfg, pk = make_model(variant, kwargs, device)
    g = torch.Generator().manual_seed(seed)
    Fq = cfg.n_freqs
    x = torch.randn(B, T, Fq, cfg.D, generator=g)
    h0 = 0.5 * torch.randn(1, B * Fq, cfg.H, generator=g)
    c0 = 0.5 * torch.randn(1, B * Fq, cfg.H, generator=g)
    ref, rh, rc = orc.inter_path(sd, ocfg, block, x, h0, c0)
    x0 = x.to(device)
    x1 = None
    if two_inputs:
        part = torch.randn(B, T, Fq, cfg.D, generator=g)
        x0, x1 = (x - part).to(device), part.to(device)
    y = torch.full((B, T, Fq, cfg.D), float("nan"), device=device)
    hd, cd = h0.to(device), c0.to(device)
    hN = hd if alias_state else torch.full_like(hd, float("nan"))
    cN = cd if alias_state else torch.full_like(cd, float("nan"))
    a = abi.InterArgs()
    a.x0, a.x1, a.y = x0.data_ptr(), (x1.data_ptr() if x1 is not None else None), y.data_ptr()
    a.h0, a.c0, a.hN, a.cN = hd.data_ptr(), cd.data_ptr(), hN.data_ptr(), cN.data_ptr()
    a.dir = pk.lstm_dir(block, "inter")
    a.B, a.T, a.F, a.C, a.H, a.algo = B, T, Fq, cfg.D, cfg.H, algo
    abi.check(lib, lib.sb_inter_lstm_fwd(ctypes.byref(a), _stream(device)), "sb_inter_lstm_fwd")
    _sync(device)
    return {"y": maxerr(y, ref), "h": maxerr(hN, rh), "c": maxerr(cN, rc)}


def check_attn(lib, device, variant, kwargs, B=2, T=6, block=0, seed=8):
    ocfg, sd, cfg, pk = make_model(variant, kwargs, device)
    g = torch.Generator().manual_seed(seed)
    Fq, W, L = cfg.n_freqs, cfg.local_atten_len, cfg.L
    x = torch.randn(B, T, Fq, cfg.D, generator=g)
    Kb = torch.randn(B * L, W - 1, cfg.attn_E * Fq, generator=g)
    Vb = torch.randn(B * L, W - 1, (cfg.D // L) * Fq, generator=g)
    ref, rK, rV = orc.attention_path(sd, ocfg, block, x, Kb, Vb)
    xd, Kd, Vd = x.to(device), Kb.to(device), Vb.to(device)
    Ko, Vo = torch.full_like(Kd, float("nan")), torch.full_like(Vd, float("nan"))
    n = lib.sb_attn_workspace_floats(B, T, Fq, cfg.D, L, cfg.attn_E, W)
    ws = torch.empty(max(int(n), 1), device=device)
    a = abi.AttnArgs()
    a.x, a.y, a.ws = xd.data_ptr(), xd.data_ptr(), ws.data_ptr()
    bd = pk.desc.blocks[block]
    a.q, a.k, a.v, a.o = bd.attn_q, bd.attn_k, bd.attn_v, bd.attn_o
    a.K_buf_in, a.K_buf_out, a.V_buf_in, a.V_buf_out = Kd.data_ptr(), Ko.data_ptr(), Vd.data_ptr(), Vo.data_ptr()
    a.B, a.T, a.F, a.C, a.L, a.E, a.W = B, T, Fq, cfg.D, L, cfg.attn_E, W
    abi.check(lib, lib.sb_attn_fwd(ctypes.byref(a), _stream(device)), "sb_attn_fwd")
    _sync(device)
    return {"y": maxerr(xd, ref), "K": maxerr(Ko, rK), "V": maxerr(Vo, rV)}


def check_backend(lib, device, variant, kwargs, B=2, T=5, seed=7, with_mask=False):
    ocfg, sd, cfg, pk = make_model(variant, kwargs, device)
    g = torch.Generator().manual_seed(seed)
    Fq, S = cfg.n_freqs, cfg.num_src
    x = torch.randn(B, T, Fq, cfg.D, generator=g)
    dbuf = torch.randn(B, cfg.D, 2, Fq, generator=g)
    ibuf = torch.randn(B, S, 2 * Fq, 1, generator=g)
    spec, ref_dbuf = orc.deconv_out(sd, ocfg, x, dbuf)
    mask = None
    if with_mask:
        mask = torch.randn(B, S, 2 * Fq, T, generator=g)
        spec = spec * mask
    ref_wave, ref_ibuf = orc.istft_ola(spec, ibuf, sd["tfgridnet.dec.filterbank._filters"], cfg.stft_chunk_size,
                                       cfg.n_fft - cfg.stft_chunk_size)
    a = abi.BackendArgs()
    xd, dbd, ibd = x.to(device), dbuf.to(device), ibuf.to(device)
    dbo, ibo = torch.full_like(dbd, float("nan")), torch.full_like(ibd, float("nan"))
    wave = torch.full((B, S, cfg.stft_chunk_size * T), float("nan"), device=device)
    ws = torch.empty(B * S * T * 2 * Fq, device=device)
    md = mask.permute(0, 3, 1, 2).contiguous().to(device) if with_mask else None      # [B, T, S, 2F]
    a.x, a.deconv_buf_in, a.deconv_buf_out = xd.data_ptr(), dbd.data_ptr(), dbo.data_ptr()
    a.istft_buf_in, a.istft_buf_out = ibd.data_ptr(), ibo.data_ptr()
    a.w, a.bias, a.filt = pk.ptr("deconv_w"), pk.ptr("deconv_bias"), pk.ptr("dec_filt")
    a.mask_spec = md.data_ptr() if with_mask else None
    a.wave_out, a.ws = wave.data_ptr(), ws.data_ptr()
    a.B, a.T, a.F, a.C, a.n_src, a.n_fft, a.stride = B, T, Fq, cfg.D, S, cfg.n_fft, cfg.stft_chunk_size
    abi.check(lib, lib.sb_backend_fwd(ctypes.byref(a), _stream(device)), "sb_backend_fwd")
    _sync(device)
    return {"wave": maxerr(wave, ref_wave), "deconv_buf": maxerr(dbo, ref_dbuf), "istft_buf": maxerr(ibo, ref_ibuf)}
