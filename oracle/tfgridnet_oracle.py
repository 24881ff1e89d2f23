"""CPU oracle for the Sound Bubble separator forward pass.  TEST INFRASTRUCTURE — NOT A PRODUCT PATH.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  ``sound_bubble_b200`` never does: the product path is the sm_100a CUDA library and it raises
when that library is missing.

What it is: a from-scratch, functional restatement (plain fp32 PyTorch ops on CPU, channels-last ``[B, T, F, C]``
throughout, weights read straight from a reference-layout ``state_dict``) of the algorithm in

  /root/reference/src/models/tfgridnet_realtime_clean_dis_embd3/{net.py, tfgridnet_causal.py}   ("DE3")
  /root/reference/src/models/tfgridnet_realtime_clean_optim/{net.py, tfgridnet_causal.py}       ("OPT")

Parity pin: the reference ships no golden vectors.  The pin is therefore "outputs of the
reference itself run here": ``oracle/make_golden.py`` imports the UNMODIFIED reference modules (with the
third-party stand-ins under ``oracle/shims``), loads the deterministic weights of ``oracle/weights.py`` with
``load_state_dict(strict=True)``, and stores its outputs under ``tests/golden``.  ``tests/test_oracle_golden.py``
checks this restatement against those files, so the oracle is pinned to the reference's own arithmetic.

Each function cites the reference lines it restates (paths relative to the DE3 directory unless noted).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------------------
# configuration (Net.__init__ kwargs, DE3/net.py:21-26 and OPT/net.py:21-26)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    variant: str = "dis_embed"          # "dis_embed" (DE3) | "optim" (OPT)
    stft_chunk_size: int = 160
    stft_pad_size: int = 120
    stft_back_pad: int = 0
    num_ch: int = 2
    D: int = 64
    B: int = 6
    I: int = 1
    J: int = 1
    L: int = 0
    H: int = 128
    use_attn: bool = False
    lookahead: bool = True
    local_atten_len: int = 100
    E: int = 4
    chunk_causal: bool = False
    num_src: int = 1
    spectral_masking: bool = False
    use_first_ln: bool = False
    merge_method: str = "None"
    directional: bool = False
    conv_lstm: bool = True
    fb_type: str = "stft"
    dis_type: str = "conv3"             # DE3 only
    lstm_down: Optional[int] = None     # OPT exposes it (default 5); DE3's core default is 4 (tfgridnet_causal.py:282)
    eps: float = 1e-5

    def __post_init__(self):
        if self.lstm_down is None:
            self.lstm_down = 5 if self.variant == "optim" else 4
        assert self.variant in ("dis_embed", "optim")
        assert self.stft_back_pad == 0, "stft_back_pad > 0 (causal_decoder, :423-431) is not restated"
        assert self.fb_type == "stft"
        assert self.merge_method in ("None", "early_cat")

    # derived sizes -------------------------------------------------------------------------------------------
    @property
    def n_fft(self):          # DE3/net.py:38
        return self.stft_back_pad + self.stft_chunk_size + self.stft_pad_size

    @property
    def n_freqs(self):        # tfgridnet_causal.py:308
        return self.n_fft // 2 + 1

    @property
    def feat_num(self):       # :335-338
        if self.merge_method != "early_cat":
            return 0
        return (self.num_ch - 1) * 3 - (1 if self.directional else 0)

    @property
    def conv_in_ch(self):     # :342-347
        return 2 * self.num_ch + self.feat_num

    @property
    def film_in(self):        # :356-374
        return {"linear1": 1, "linear2": self.D, "conv1": 1, "conv2": 2, "conv3": 4, "conv4": 8}[self.dis_type]

    @property
    def attn_E(self):         # :591-593 with approx_qk_dim = E * n_freqs (net.py:53)
        return math.ceil(self.E * self.n_freqs * 1.0 / self.n_freqs)

    @classmethod
    def from_kwargs(cls, variant: str, **kw) -> "OracleConfig":
        return cls(variant=variant, **kw)


# --------------------------------------------------------------------------------------------------------------
# STFT basis (third-party asteroid_filterbanks.STFTFB; algorithm restated in oracle/shims/asteroid_filterbanks)
# --------------------------------------------------------------------------------------------------------------
def stft_basis(n_fft: int, stride: int) -> Tensor:
    """``_filters`` buffer ``[n_fft + 2, 1, n_fft]`` built at tfgridnet_causal.py:326-330.

    Same construction as the published STFTFB (DFT of the identity, so the float32 rounding is bit-identical to
    what a real checkpoint stores); the closed form is ``sqrt(hann_periodic[n]) * {cos, -sin}(2 pi k n / N) /
    (0.5 sqrt(N*N/stride))`` with the DC and Nyquist real rows divided by sqrt(2) (checked in the tests).
    """
    window = np.hanning(n_fft + 1)[:-1] ** 0.5
    basis = np.fft.fft(np.eye(n_fft)) / (0.5 * np.sqrt(n_fft * n_fft / stride))
    cut = n_fft // 2 + 1
    basis = np.vstack([np.real(basis[:cut]), np.imag(basis[:cut])])
    basis[0] /= np.sqrt(2.0)
    basis[n_fft // 2] /= np.sqrt(2.0)
    return torch.from_numpy(basis * window).unsqueeze(1).float()


def stft_basis_closed_form(n_fft: int, stride: int) -> Tensor:
    n = np.arange(n_fft)
    k = np.arange(n_fft // 2 + 1)
    window = np.sqrt(0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft))
    ang = 2.0 * np.pi * np.outer(k, n) / n_fft
    scale = 0.5 * np.sqrt(n_fft * n_fft / stride)
    re, im = np.cos(ang) / scale, -np.sin(ang) / scale
    re[0] /= np.sqrt(2.0)
    re[n_fft // 2] /= np.sqrt(2.0)
    return torch.from_numpy(np.concatenate([re, im], axis=0) * window[None, :]).float().unsqueeze(1)


# --------------------------------------------------------------------------------------------------------------
# parameter inventory: reference ``state_dict`` key -> shape  (verified against the reference by
# oracle/make_golden.py, which loads these with strict=True)
# --------------------------------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig) -> "Dict[str, Tuple[int, ...]]":
    C, H, Fq, nfft = cfg.D, cfg.H, cfg.n_freqs, cfg.n_fft
    s: Dict[str, Tuple[int, ...]] = {}
    p = "tfgridnet."
    s[p + "enc.filterbank._filters"] = (nfft + 2, 1, nfft)
    s[p + "dec.filterbank._filters"] = (nfft + 2, 1, nfft)
    s[p + "conv.0.weight"] = (C, cfg.conv_in_ch, 3, 3)
    s[p + "conv.0.bias"] = (C,)
    if cfg.use_first_ln:
        s[p + "conv.1.weight"] = (C,)
        s[p + "conv.1.bias"] = (C,)
    if cfg.variant == "dis_embed":
        if cfg.dis_type.startswith("conv"):
            s[p + "embed_net.dis_embedding.0.weight"] = (Fq * cfg.film_in, 3)
            s[p + "embed_net.dis_norm.weight"] = (cfg.film_in,)
            s[p + "embed_net.dis_norm.bias"] = (cfg.film_in,)
        else:
            n = Fq if cfg.dis_type == "linear1" else Fq * C
            s[p + "embed_net.dis_embedding.0.weight"] = (n, 3)
            s[p + "embed_net.dis_embedding.1.weight"] = (n,)
            s[p + "embed_net.dis_embedding.1.bias"] = (n,)
    for i in range(cfg.B):
        b = f"{p}blocks.{i}."
        if cfg.conv_lstm:
            s[b + "conv.weight"] = (C, C, cfg.lstm_down)
            s[b + "conv.bias"] = (C,)
            s[b + "act.weight"] = (1,)
            s[b + "norm.norm.weight"] = (C,)
            s[b + "norm.norm.bias"] = (C,)
        else:
            s[b + "intra_norm.norm.weight"] = (C,)
            s[b + "intra_norm.norm.bias"] = (C,)
        for sfx in ("", "_reverse"):
            s[b + "intra_rnn.weight_ih_l0" + sfx] = (4 * H, C)
            s[b + "intra_rnn.weight_hh_l0" + sfx] = (4 * H, H)
            s[b + "intra_rnn.bias_ih_l0" + sfx] = (4 * H,)
            s[b + "intra_rnn.bias_hh_l0" + sfx] = (4 * H,)
        if cfg.conv_lstm:
            s[b + "deconv.weight"] = (2 * H, C, cfg.lstm_down)
            s[b + "deconv.bias"] = (C,)
        else:
            s[b + "intra_linear.weight"] = (C, 2 * H)
            s[b + "intra_linear.bias"] = (C,)
        s[b + "inter_norm.norm.weight"] = (C,)
        s[b + "inter_norm.norm.bias"] = (C,)
        s[b + "inter_rnn.weight_ih_l0"] = (4 * H, C)
        s[b + "inter_rnn.weight_hh_l0"] = (4 * H, H)
        s[b + "inter_rnn.bias_ih_l0"] = (4 * H,)
        s[b + "inter_rnn.bias_hh_l0"] = (4 * H,)
        s[b + "inter_linear.weight"] = (C, H)
        s[b + "inter_linear.bias"] = (C,)
        if cfg.use_attn:
            LE, Vd = cfg.L * cfg.attn_E, C // cfg.L
            for nm, out, width in (("Q", LE, Fq * cfg.attn_E), ("K", LE, Fq * cfg.attn_E), ("V", Vd * cfg.L, Fq * Vd)):
                s[b + f"attn_conv_{nm}.0.weight"] = (out, C)
                s[b + f"attn_conv_{nm}.0.bias"] = (out,)
                s[b + f"attn_conv_{nm}.1.weight"] = (1,)
                s[b + f"attn_conv_{nm}.3.norm.weight"] = (width,)
                s[b + f"attn_conv_{nm}.3.norm.bias"] = (width,)
            s[b + "attn_concat_proj.0.weight"] = (C, C)
            s[b + "attn_concat_proj.0.bias"] = (C,)
            s[b + "attn_concat_proj.1.weight"] = (1,)
            s[b + "attn_concat_proj.3.norm.weight"] = (Fq * C,)
            s[b + "attn_concat_proj.3.norm.bias"] = (Fq * C,)
    if cfg.variant == "dis_embed":
        for j in range(cfg.B - 1):
            e = f"{p}embeds.{j}."
            s[e + "weight.weight"] = (C, cfg.film_in, 1)
            s[e + "weight.bias"] = (C,)
            s[e + "bias.weight"] = (C, cfg.film_in, 1)
            s[e + "bias.bias"] = (C,)
    s[p + "deconv.weight"] = (C, 2 * cfg.num_src, 3, 3)
    s[p + "deconv.bias"] = (2 * cfg.num_src,)
    return s


# --------------------------------------------------------------------------------------------------------------
# state (init_buffers: tfgridnet_causal.py:403-421, 696-720)
# --------------------------------------------------------------------------------------------------------------
def init_state(cfg: OracleConfig, batch: int) -> dict:
    Fq = cfg.n_freqs
    st = {
        "conv_buf": torch.zeros(batch, cfg.conv_in_ch, 2, Fq),
        "deconv_buf": torch.zeros(batch, cfg.D, 2, Fq),
        "istft_buf": torch.zeros(batch, cfg.num_src, 2 * Fq, 1),
        "gridnet_bufs": {},
    }
    for i in range(cfg.B):
        buf = {}
        if cfg.use_attn:
            W = cfg.local_atten_len
            buf["K_buf"] = torch.zeros(batch * cfg.L, W - 1, cfg.attn_E * Fq)
            buf["V_buf"] = torch.zeros(batch * cfg.L, W - 1, (cfg.D // cfg.L) * Fq)
        buf["c0"] = torch.zeros(1, batch * Fq, cfg.H)
        buf["h0"] = torch.zeros(1, batch * Fq, cfg.H)
        st["gridnet_bufs"][f"buf{i}"] = buf
    return st


# --------------------------------------------------------------------------------------------------------------
# stages
# --------------------------------------------------------------------------------------------------------------
def stft_frames(wave: Tensor, basis: Tensor, stride: int) -> Tuple[Tensor, Tensor]:
    """a3 + a4: ``self.enc(input)`` (:475) and the re/im split (:482-483).  wave [B, M, N] -> re, im [B, M, F, T]."""
    Bn, M, N = wave.shape
    spec = F.conv1d(wave.reshape(Bn * M, 1, N), basis, stride=stride)
    spec = spec.reshape(Bn, M, spec.shape[-2], spec.shape[-1])
    nf = basis.shape[0] // 2
    return spec[:, :, :nf], spec[:, :, nf:]


def spatial_features(re: Tensor, im: Tensor, directional: bool, eps: float = 1e-6) -> Tensor:
    """a5 / a5': MC_features_OMNX (:72-93) + IPD_OMNX (:32-48); MC_features_direct (:176-207).

    re, im [B, M, F, T] -> [B, 3(M-1) (-1 if directional), F, T]; channel order ILD..., then (sin_m, cos_m) pairs.
    """
    mag = torch.sqrt(re * re + im * im)
    ref_mag, ref_re, ref_im = mag[:, :1], re[:, :1], im[:, :1]
    o_mag, o_re, o_im = mag[:, 1:], re[:, 1:], im[:, 1:]
    den = o_mag * ref_mag + eps
    cos = (o_re * ref_re + o_im * ref_im) / den
    sin = (ref_re * o_im - ref_im * o_re) / den
    ipd = torch.stack([sin, cos], dim=2).flatten(1, 2)           # sin1, cos1, sin2, cos2, ...
    if not directional:
        ild = torch.log10((o_mag + eps) / (ref_mag + eps))
        return torch.cat([ild, ipd], dim=1)
    ild_d = torch.log10((mag[:, 2:3] + eps) / (mag[:, 3:4] + eps))
    ild_m = torch.log10((mag[:, [1, 4, 5]] + eps) / (ref_mag + eps))
    return torch.cat([ild_d, ild_m, ipd], dim=1)


def conv_in(sd, cfg: OracleConfig, feats_tf: Tensor, conv_buf: Tensor) -> Tuple[Tensor, Tensor]:
    """a6: history cat (:504-505), Conv2d k=(3,3) pad (0,1) (:332-347), LayerNormPermuted (:219-231).

    feats_tf [B, Cin, T, F] (time-major), conv_buf [B, Cin, 2, F] -> x [B, T, F, C], new conv_buf.
    """
    p = "tfgridnet."
    full = torch.cat([conv_buf, feats_tf], dim=2)
    new_buf = full[:, :, -2:, :]
    y = F.conv2d(full, sd[p + "conv.0.weight"], sd[p + "conv.0.bias"], padding=(0, 1))     # [B, C, T, F]
    y = y.permute(0, 2, 3, 1)
    if cfg.use_first_ln:
        y = F.layer_norm(y, (cfg.D,), sd[p + "conv.1.weight"], sd[p + "conv.1.bias"], 1e-5)
    return y.contiguous(), new_buf


def distance_embedding(sd, cfg: OracleConfig, dis_embed: Tensor) -> Tensor:
    """a7: Dis_Embed_Conv (:150-173) / Dis_Embed_Linear (:114-147).  [B, 3] -> [B, film_in, F]."""
    p = "tfgridnet.embed_net."
    Fq = cfg.n_freqs
    e = dis_embed @ sd[p + "dis_embedding.0.weight"].t()
    if cfg.dis_type.startswith("conv"):
        e = e.view(e.shape[0], Fq, cfg.film_in)
        e = F.layer_norm(e, (cfg.film_in,), sd[p + "dis_norm.weight"], sd[p + "dis_norm.bias"], 1e-5)
        return e.transpose(1, 2)
    e = F.layer_norm(e, (e.shape[-1],), sd[p + "dis_embedding.1.weight"], sd[p + "dis_embedding.1.bias"], 1e-5)
    if cfg.dis_type == "linear1":
        return e.unsqueeze(1)
    return e.view(e.shape[0], cfg.D, Fq)


def film(sd, j: int, x: Tensor, emb: Tensor) -> Tensor:
    """a8: FilmLayer (:51-68) applied at :509-513.  x [B, T, F, C], emb [B, Din, F] -> x * w + b."""
    p = f"tfgridnet.embeds.{j}."
    w = F.conv1d(emb, sd[p + "weight.weight"], sd[p + "weight.bias"])          # [B, C, F]
    b = F.conv1d(emb, sd[p + "bias.weight"], sd[p + "bias.bias"])
    return x * w.transpose(1, 2).unsqueeze(1) + b.transpose(1, 2).unsqueeze(1)


def _lstm(x: Tensor, hc: Tuple[Tensor, Tensor], sd, prefix: str, bidirectional: bool):
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    flat = [sd[prefix + n] for n in names]
    if bidirectional:
        flat += [sd[prefix + n + "_reverse"] for n in names]
    out, h, c = torch._VF.lstm(x, hc, flat, True, 1, 0.0, False, bidirectional, True)
    return out, h, c


def intra_path(sd, cfg: OracleConfig, i: int, x: Tensor) -> Tensor:
    """a9 / a9': intra-frame BiLSTM across frequency (:794-827; OPT :684-707).  x [B, T, F, C] -> same."""
    b = f"tfgridnet.blocks.{i}."
    Bn, T, Fq, C = x.shape
    H = cfg.H
    zeros = (torch.zeros(2, Bn * T, H), torch.zeros(2, Bn * T, H))
    if not cfg.conv_lstm:
        y = F.layer_norm(x, (C,), sd[b + "intra_norm.norm.weight"], sd[b + "intra_norm.norm.bias"], cfg.eps)
        y, _, _ = _lstm(y.reshape(Bn * T, Fq, C), zeros, sd, b + "intra_rnn.", True)
        y = F.linear(y, sd[b + "intra_linear.weight"], sd[b + "intra_linear.bias"])
    else:
        k = cfg.lstm_down
        y = x.reshape(Bn * T, Fq, C).transpose(1, 2)
        y = F.conv1d(y, sd[b + "conv.weight"], sd[b + "conv.bias"], stride=k)            # [BT, C, F//k]
        y = F.prelu(y, sd[b + "act.weight"])
        y = F.layer_norm(y.transpose(1, 2), (C,), sd[b + "norm.norm.weight"], sd[b + "norm.norm.bias"], 1e-5)
        y, _, _ = _lstm(y, zeros, sd, b + "intra_rnn.", True)
        if cfg.variant == "optim":      # OPT :506-510 output_padding
            op = Fq - (Fq // k) * k
            y = F.conv_transpose1d(y.transpose(1, 2), sd[b + "deconv.weight"], sd[b + "deconv.bias"], stride=k,
                                   output_padding=op)
        else:                            # DE3 :810-813 pad 3 zeros then crop
            y = F.conv_transpose1d(y.transpose(1, 2), sd[b + "deconv.weight"], sd[b + "deconv.bias"], stride=k)
            y = F.pad(y, (0, 3))[..., :Fq]
        y = y.transpose(1, 2)
    return y.reshape(Bn, T, Fq, C) + x


def inter_path(sd, cfg: OracleConfig, i: int, x: Tensor, h0: Tensor, c0: Tensor):
    """a10: inter-frame LSTM across time with carried (h0, c0) (:829-849).  Rows are b*F + f (:833)."""
    b = f"tfgridnet.blocks.{i}."
    Bn, T, Fq, C = x.shape
    y = F.layer_norm(x, (C,), sd[b + "inter_norm.norm.weight"], sd[b + "inter_norm.norm.bias"], cfg.eps)
    y = y.transpose(1, 2).reshape(Bn * Fq, T, C)
    y, h, c = _lstm(y, (h0, c0), sd, b + "inter_rnn.", False)
    y = F.linear(y, sd[b + "inter_linear.weight"], sd[b + "inter_linear.bias"])
    y = y.view(Bn, Fq, T, C).transpose(1, 2)
    return y + x, h, c


def _attn_branch(sd, prefix: str, x: Tensor, heads: int, eps: float) -> Tensor:
    """Linear -> PReLU -> head split -> LayerNorm(F*E) (:642-675).  x [B, T, F, C] -> [B*L, T, F*E]."""
    Bn, T, Fq, _ = x.shape
    y = F.prelu(F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"]), sd[prefix + "1.weight"])
    E = y.shape[-1] // heads
    y = y.view(Bn, T, Fq, heads, E).permute(0, 3, 1, 2, 4).reshape(Bn * heads, T, Fq * E)
    return F.layer_norm(y, (Fq * E,), sd[prefix + "3.norm.weight"], sd[prefix + "3.norm.bias"], eps)


def attention_path(sd, cfg: OracleConfig, i: int, x: Tensor, K_buf: Tensor, V_buf: Tensor):
    """a11: sliding-window single-query attention with zero-initialised, unmasked history (:856-898, :722-744)."""
    b = f"tfgridnet.blocks.{i}."
    Bn, T, Fq, C = x.shape
    L, W = cfg.L, cfg.local_atten_len
    Q = _attn_branch(sd, b + "attn_conv_Q.", x, L, cfg.eps)
    K = torch.cat([K_buf, _attn_branch(sd, b + "attn_conv_K.", x, L, cfg.eps)], dim=1)
    V = torch.cat([V_buf, _attn_branch(sd, b + "attn_conv_V.", x, L, cfg.eps)], dim=1)
    new_K, new_V = K[:, -(W - 1):], V[:, -(W - 1):]
    Kw = K.unfold(1, W, 1)                                    # [BL, T, FE, W]
    Vw = V.unfold(1, W, 1)                                    # [BL, T, FV, W]
    logits = torch.einsum("btd,btdw->btw", Q, Kw) / math.sqrt(Q.shape[-1])
    att = torch.softmax(logits, dim=-1)
    o = torch.einsum("btw,btdw->btd", att, Vw)                # [BL, T, F*Vd]
    Vd = C // L
    o = o.view(Bn, L, T, Fq, Vd).permute(0, 2, 3, 1, 4).reshape(Bn, T, Fq, C)
    p = b + "attn_concat_proj."
    o = F.prelu(F.linear(o, sd[p + "0.weight"], sd[p + "0.bias"]), sd[p + "1.weight"])
    o = F.layer_norm(o.reshape(Bn, T, Fq * C), (Fq * C,), sd[p + "3.norm.weight"], sd[p + "3.norm.bias"], cfg.eps)
    return x + o.view(Bn, T, Fq, C), new_K, new_V


def deconv_out(sd, cfg: OracleConfig, x: Tensor, deconv_buf: Tensor) -> Tuple[Tensor, Tensor]:
    """a13: history cat (:517-518), ConvTranspose2d k=(3,3) pad (2,1) (:401,520), re/im concat (:521-526).

    x [B, T, F, C], deconv_buf [B, C, 2, F] -> spec [B, S, 2F, T], new deconv_buf.
    """
    Bn, T, Fq, _ = x.shape
    full = torch.cat([deconv_buf, x.permute(0, 3, 1, 2)], dim=2)
    new_buf = full[:, :, -2:, :]
    y = F.conv_transpose2d(full, sd["tfgridnet.deconv.weight"], sd["tfgridnet.deconv.bias"], padding=(2, 1))
    y = y.view(Bn, cfg.num_src, 2, T, Fq).transpose(3, 4)      # [B, S, 2, F, T]
    return y.reshape(Bn, cfg.num_src, 2 * Fq, T), new_buf


def istft_ola(spec: Tensor, istft_buf: Tensor, basis: Tensor, stride: int, lookahead: int):
    """a15: previous-frame cat (:533-534), ``self.dec`` (:537), crops (:538,542).  -> wave [B, S, stride*T]."""
    full = torch.cat([istft_buf, spec], dim=3)
    new_buf = full[..., -1:]
    Bn, S, F2, T1 = full.shape
    y = F.conv_transpose1d(full.reshape(Bn * S, F2, T1), basis, stride=stride).view(Bn, S, -1)
    y = y[..., :-lookahead][..., stride:]
    return y, new_buf


# --------------------------------------------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------------------------------------------
def core_forward(sd, cfg: OracleConfig, wave: Tensor, dis_embed: Optional[Tensor], state: dict):
    """TFGridNet.forward (:433-552).  wave [B, M, stride*T + (n_fft - stride)] -> ([B, S, stride*T], state)."""
    p = "tfgridnet."
    stride = cfg.stft_chunk_size
    re, im = stft_frames(wave, sd[p + "enc.filterbank._filters"], stride)
    chans = [re, im]
    if cfg.merge_method == "early_cat":
        chans.append(spatial_features(re, im, cfg.directional))
    feats = torch.cat(chans, dim=1).transpose(2, 3)            # [B, Cin, T, F]
    x, conv_buf = conv_in(sd, cfg, feats, state["conv_buf"])
    emb = distance_embedding(sd, cfg, dis_embed) if cfg.variant == "dis_embed" else None
    bufs = state["gridnet_bufs"]
    for i in range(cfg.B):
        if i > 0 and emb is not None:
            x = film(sd, i - 1, x, emb)
        buf = bufs[f"buf{i}"]
        x = intra_path(sd, cfg, i, x)
        x, buf["h0"], buf["c0"] = inter_path(sd, cfg, i, x, buf["h0"], buf["c0"])
        if cfg.use_attn:
            x, buf["K_buf"], buf["V_buf"] = attention_path(sd, cfg, i, x, buf["K_buf"], buf["V_buf"])
    spec, deconv_buf = deconv_out(sd, cfg, x, state["deconv_buf"])
    if cfg.spectral_masking:                                   # :529-530
        re_im = torch.cat([re, im], dim=2)
        spec = spec * re_im[:, : cfg.num_src]
    wave_out, istft_buf = istft_ola(spec, state["istft_buf"], sd[p + "dec.filterbank._filters"], stride,
                                    cfg.n_fft - stride)
    state["conv_buf"], state["deconv_buf"], state["istft_buf"] = conv_buf, deconv_buf, istft_buf
    return wave_out, state


def net_forward(sd, cfg: OracleConfig, inputs: dict, input_state: Optional[dict] = None, pad: bool = True) -> dict:
    """Net.forward / Net.predict / mod_pad (DE3/net.py:8-18, 70-93)."""
    x = inputs["mixture"]
    dis = inputs["dis_embed"] if cfg.variant == "dis_embed" else None
    if input_state is None:
        input_state = init_state(cfg, x.shape[0])
    mod = 0
    if pad:
        chunk = cfg.stft_chunk_size
        if x.shape[-1] % chunk:
            mod = chunk - x.shape[-1] % chunk
        x = F.pad(x, (0, mod))
        if cfg.lookahead:
            x = F.pad(x, (cfg.stft_back_pad, cfg.stft_pad_size))
    y, st = core_forward(sd, cfg, x, dis, input_state)
    if mod:
        y = y[:, :, :-mod]
    return {"output": y, "next_state": st}


def streaming_forward(sd, cfg: OracleConfig, wave: Tensor, dis_embed: Optional[Tensor]) -> Tensor:
    """Chunk-by-chunk protocol of /root/reference/edge/causal_infer.py:28-47 (window rolls by one chunk)."""
    chunk, padlen = cfg.stft_chunk_size, cfg.stft_pad_size
    Bn, M, N = wave.shape
    state = init_state(cfg, Bn)
    frame = torch.zeros(Bn, M, chunk + padlen)
    frame[..., -padlen:] = wave[..., :padlen]
    outs = []
    for i in range(padlen, N - padlen + 1, chunk):
        frame = torch.roll(frame, shifts=-chunk, dims=-1)
        frame[..., -chunk:] = wave[..., i:i + chunk]
        r = net_forward(sd, cfg, {"mixture": frame, "dis_embed": dis_embed}, state, pad=False)
        state = r["next_state"]
        outs.append(r["output"])
    return torch.cat(outs, dim=-1)


# --------------------------------------------------------------------------------------------------------------
# parity metrics (torchmetrics semantics used at /root/reference/src/metrics/metrics.py:6-9,52-55)
# --------------------------------------------------------------------------------------------------------------
def si_sdr(est: Tensor, ref: Tensor) -> Tensor:
    eps = torch.finfo(est.dtype).eps
    alpha = ((est * ref).sum(-1, keepdim=True) + eps) / ((ref * ref).sum(-1, keepdim=True) + eps)
    tgt = alpha * ref
    return 10.0 * torch.log10(((tgt * tgt).sum(-1) + eps) / (((tgt - est) ** 2).sum(-1) + eps))


def rms(x: Tensor) -> float:
    return float(x.double().pow(2).mean().sqrt())
