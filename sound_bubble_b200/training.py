"""Training path of the separator: a differentiable forward + hand-written backward kernels behind one autograd node.

What the reference differentiates is ``self.model(inputs)`` inside ``PLModule._step`` followed by ``loss.backward()``
(src/hl_modules/distance_based_hl_module.py:303-330, 437-441): TFGridNet.forward from zero state
(DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py:433-552).  ``TrainGraph`` runs that launch sequence
through the training entry points of include/soundbubble.h (``*_train_fwd`` keep what back-propagation needs, ``*_bwd`` are
the twins); ``SeparatorFunction`` is the ``torch.autograd.Function`` that puts it behind ``Net.forward`` when the module is
in training mode with gradients enabled.  Parameters are read in their checkpoint layouts - nothing is re-packed per step.

Supported: plain BiLSTM and conv-LSTM intra-frame paths, the inter-frame LSTM, the windowed self-attention unit
(use_attn=True, DE3:856-898; L*E in {8, 16}), FiLM with either distance embedding (dis_type conv* / linear*), 1-2 sources,
optional spectral masking and first LayerNorm - i.e. every shipped training config and the attention the constructor offers.
Anything else raises ``NotImplementedError`` rather than training a different model.  No CPU path: the library handed in
is the sm_100a build (the tests' host-emulated build goes through the same code on tiny shapes).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _abi as abi
from .packing import ModelConfig

_LSTM = ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()




def check_trainable(cfg: ModelConfig):
    if cfg.conv_lstm and cfg.lstm_down * cfg.D > 256:
        raise NotImplementedError("training: conv-LSTM backward kernels need lstm_down * D <= 256")
    if cfg.use_attn and cfg.L * cfg.attn_E not in (8, 16):
        raise NotImplementedError("training: attention backward kernels need L * E in {8, 16}")
    if cfg.H != 64 or cfg.D not in (16, 32):
        raise NotImplementedError("training: kernels are instantiated for H = 64 and D in {16, 32}")


class TrainGraph:
    """One forward of the separator that can be back-propagated.  ``params``: state_dict names -> float32 tensors."""

    def __init__(self, lib, cfg: ModelConfig):
        check_trainable(cfg)
        self.lib, self.cfg = lib, cfg
        self.has_film = cfg.variant == "dis_embed" and cfg.B > 1

    # -- small helpers ---------------------------------------------------------------------------------------
    def _call(self, fn, args, like, what):
        abi.check(self.lib, fn(ctypes.byref(args), _stream_ptr(like)), what)

    def _path_args(self, P, i: int, inter: bool, B: int, T: int) -> abi.PathTrainArgs:
        cfg = self.cfg
        b = f"tfgridnet.blocks.{i}."
        kind = "inter" if inter else "intra"
        a = abi.PathTrainArgs()
        a.ln_g, a.ln_b = P[b + kind + "_norm.norm.weight"].data_ptr(), P[b + kind + "_norm.norm.bias"].data_ptr()
        for d, sfx in enumerate(("",) if inter else ("", "_reverse")):
            w_ih, w_hh, b_ih, b_hh = (P[b + kind + "_rnn." + n + sfx] for n in _LSTM)
            a.w_ih[d], a.w_hh[d], a.b_ih[d], a.b_hh[d] = w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr()
        a.lin_w, a.lin_b = P[b + kind + "_linear.weight"].data_ptr(), P[b + kind + "_linear.bias"].data_ptr()
        a.B, a.T, a.F, a.C, a.H, a.inter = B, T, cfg.n_freqs, cfg.D, cfg.H, int(inter)
        return a

    def _convpath_args(self, P, i: int, B: int, T: int) -> abi.ConvPathTrainArgs:
        cfg = self.cfg
        b = f"tfgridnet.blocks.{i}."
        a = abi.ConvPathTrainArgs()
        a.conv_w, a.conv_b, a.prelu = P[b + "conv.weight"].data_ptr(), P[b + "conv.bias"].data_ptr(), P[b + "act.weight"].data_ptr()
        a.ln_g, a.ln_b = P[b + "norm.norm.weight"].data_ptr(), P[b + "norm.norm.bias"].data_ptr()
        for d, sfx in enumerate(("", "_reverse")):
            w_ih, w_hh, b_ih, b_hh = (P[b + "intra_rnn." + n + sfx] for n in _LSTM)
            a.w_ih[d], a.w_hh[d], a.b_ih[d], a.b_hh[d] = w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr()
        a.deconv_w, a.deconv_b = P[b + "deconv.weight"].data_ptr(), P[b + "deconv.bias"].data_ptr()
        a.B, a.T, a.F, a.C, a.H = B, T, cfg.n_freqs, cfg.D, cfg.H
        a.down = cfg.lstm_down
        a.tail_mode = abi.SB_CONVLSTM_OUTPAD if cfg.variant == "optim" else abi.SB_CONVLSTM_PADCROP
        return a

    _ATTN = (("q", "attn_conv_Q."), ("k", "attn_conv_K."), ("v", "attn_conv_V."), ("o", "attn_concat_proj."))
    _ATTN_P = (("w", "0.weight"), ("b", "0.bias"), ("prelu", "1.weight"), ("ln_g", "3.norm.weight"), ("ln_b", "3.norm.bias"))

    def _attn_args(self, P, i: int, B: int, T: int) -> abi.AttnTrainArgs:
        cfg = self.cfg
        a = abi.AttnTrainArgs()
        for field, mod in self._ATTN:
            pr = getattr(a, field)
            for f2, name in self._ATTN_P:
                setattr(pr, f2, P[f"tfgridnet.blocks.{i}." + mod + name].data_ptr())
        a.B, a.T, a.F, a.C, a.L, a.E, a.W = B, T, cfg.n_freqs, cfg.D, cfg.L, cfg.attn_E, cfg.local_atten_len
        return a

    def _film_args(self, P, dis: torch.Tensor, stacks: Dict[str, torch.Tensor]) -> abi.FilmArgs:
        cfg = self.cfg
        a = abi.FilmArgs()
        a.dis = dis.data_ptr()
        a.emb_w = P["tfgridnet.embed_net.dis_embedding.0.weight"].data_ptr()
        a.emb_ln_g, a.emb_ln_b = (P[n].data_ptr() for n in self._emb_ln_names())
        a.w_w, a.w_b, a.b_w, a.b_b = (stacks[k].data_ptr() for k in ("w_w", "w_b", "b_w", "b_b"))
        a.B, a.F, a.C, a.Din, a.n_layers = dis.shape[0], cfg.n_freqs, cfg.D, cfg.film_in, cfg.B - 1
        a.emb_mode = abi.SB_EMB_CONV if cfg.dis_type.startswith("conv") else abi.SB_EMB_LINEAR
        return a

    def _emb_ln_names(self):
        """Dis_Embed_Conv keeps its LayerNorm(Din) as dis_norm (DE3:150-173), Dis_Embed_Linear as dis_embedding.1 (:114-147)"""
        e = "tfgridnet.embed_net."
        if self.cfg.dis_type.startswith("conv"):
            return e + "dis_norm.weight", e + "dis_norm.bias"
        return e + "dis_embedding.1.weight", e + "dis_embedding.1.bias"

    def _film_stacks(self, P) -> Dict[str, torch.Tensor]:
        L = self.cfg.B - 1
        e = "tfgridnet.embeds.%d."
        return {"w_w": torch.stack([P[e % j + "weight.weight"][:, :, 0] for j in range(L)]).contiguous(),
                "w_b": torch.stack([P[e % j + "weight.bias"] for j in range(L)]).contiguous(),
                "b_w": torch.stack([P[e % j + "bias.weight"][:, :, 0] for j in range(L)]).contiguous(),
                "b_b": torch.stack([P[e % j + "bias.bias"] for j in range(L)]).contiguous()}

    # -- forward ---------------------------------------------------------------------------------------------
    def forward(self, P: Dict[str, torch.Tensor], wave: torch.Tensor, dis: Optional[torch.Tensor]):
        """wave [B, M, stride*T + n_fft - stride] (already padded) -> (output [B, S, stride*T], ctx for backward)."""
        cfg, lib = self.cfg, self.lib
        dev = wave.device
        B, M, N = wave.shape
        if M != cfg.num_ch:
            raise ValueError("mixture has %d channels, the model was built for num_ch=%d" % (M, cfg.num_ch))
        hop, Fq, C, S = cfg.stft_chunk_size, cfg.n_freqs, cfg.D, cfg.num_src
        T = (N - cfg.n_fft) // hop + 1
        if T < 1:
            raise ValueError("input of %d samples is shorter than one window (%d)" % (N, cfg.n_fft))
        wave = wave[..., : hop * T + cfg.n_fft - hop].contiguous().float()
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        ctx = {"B": B, "T": T, "P": P}

        # a3-a5 (no parameters, no gradient)
        feats = new(B, T, Fq, cfg.conv_in_ch)
        spec = new(B, T, S, 2 * Fq) if cfg.spectral_masking else None
        a = abi.StftArgs()
        a.wave, a.filt, a.feats, a.spec = wave.data_ptr(), P["tfgridnet.enc.filterbank._filters"].data_ptr(), feats.data_ptr(), _ptr(spec)
        a.B, a.M, a.n_samples, a.T = B, M, wave.shape[-1], T
        a.n_fft, a.stride, a.F = cfg.n_fft, hop, Fq
        a.feat_mode = (abi.SB_FEAT_DIRECTIONAL if cfg.directional else abi.SB_FEAT_OMNI) if cfg.merge_method == "early_cat" else abi.SB_FEAT_NONE
        a.Cin, a.n_src = cfg.conv_in_ch, S
        self._call(lib.sb_stft_features_fwd, a, wave, "sb_stft_features_fwd")

        # a6
        NP = B * T * Fq
        x = new(B, T, Fq, C)
        ci = abi.ConvInTrainArgs()
        ci.feats, ci.w, ci.bias = feats.data_ptr(), P["tfgridnet.conv.0.weight"].data_ptr(), P["tfgridnet.conv.0.bias"].data_ptr()
        conv_saved = None
        if cfg.use_first_ln:
            conv_saved = new(NP * (2 * C + 1))
            ci.ln_g, ci.ln_b, ci.saved = P["tfgridnet.conv.1.weight"].data_ptr(), P["tfgridnet.conv.1.bias"].data_ptr(), conv_saved.data_ptr()
        ci.x = x.data_ptr()
        ci.B, ci.T, ci.F, ci.Cin, ci.C = B, T, Fq, cfg.conv_in_ch, C
        self._call(lib.sb_conv_in_train_fwd, ci, wave, "sb_conv_in_train_fwd")
        ctx.update(feats=feats, spec=spec, conv_saved=conv_saved)

        # a7 + a8 (parameter part)
        film = None
        if self.has_film:
            if dis is None:
                raise KeyError("dis_embed")
            dis = dis.to(dev).float().contiguous()
            stacks = self._film_stacks(P)
            film = new(cfg.B - 1, 2, B, Fq, C)
            fa = self._film_args(P, dis, stacks)
            fa.film = film.data_ptr()
            self._call(lib.sb_film_params_fwd, fa, wave, "sb_film_params_fwd")
            ctx.update(dis=dis, film=film, film_stacks=stacks, film_in=[])

        # blocks
        n_saved = [int(lib.sb_path_train_saved_floats(B, T, Fq, C, cfg.H, k)) for k in (0, 1)]
        if cfg.conv_lstm:
            n_saved[0] = int(lib.sb_convpath_train_saved_floats(B, T, Fq, C, cfg.H, cfg.lstm_down))
        ctx["saved"] = []
        ctx["intra_in"] = []
        for i in range(cfg.B):
            if i > 0 and film is not None:
                y = new(B, T, Fq, C)
                f = abi.FilmApplyArgs()
                f.x, f.film_scale, f.film_shift, f.y = x.data_ptr(), film[i - 1, 0].data_ptr(), film[i - 1, 1].data_ptr(), y.data_ptr()
                f.B, f.T, f.F, f.C = B, T, Fq, C
                self._call(lib.sb_film_apply_fwd, f, wave, "sb_film_apply_fwd")
                ctx["film_in"].append(x)
                x = y
            per_block = []
            for inter in (False, True):
                saved = new(n_saved[int(inter)])
                y = new(B, T, Fq, C)
                if cfg.conv_lstm and not inter:
                    ca = self._convpath_args(P, i, B, T)
                    ca.x, ca.y, ca.saved = x.data_ptr(), y.data_ptr(), saved.data_ptr()
                    self._call(lib.sb_intra_convlstm_train_fwd, ca, wave, "sb_intra_convlstm_train_fwd")
                    ctx["intra_in"].append(x)           # the conv weight gradient needs the path input
                else:
                    pa = self._path_args(P, i, inter, B, T)
                    pa.x, pa.y, pa.saved = x.data_ptr(), y.data_ptr(), saved.data_ptr()
                    fn = lib.sb_inter_lstm_train_fwd if inter else lib.sb_intra_lstm_train_fwd
                    self._call(fn, pa, wave, "sb_%s_lstm_train_fwd" % ("inter" if inter else "intra"))
                per_block.append(saved)
                x = y
            if cfg.use_attn:
                aa = self._attn_args(P, i, B, T)
                saved = new(int(lib.sb_attn_train_saved_floats(ctypes.byref(aa))))
                y = new(B, T, Fq, C)
                aa.x, aa.y, aa.saved = x.data_ptr(), y.data_ptr(), saved.data_ptr()
                self._call(lib.sb_attn_train_fwd, aa, wave, "sb_attn_train_fwd")
                per_block.append(saved)
                ctx.setdefault("attn_in", []).append(x)
                x = y
            ctx["saved"].append(per_block)

        # a13-a15 from zero history (the inference entry point; nothing has to be kept but its input)
        out = new(B, S, hop * T)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        ba = abi.BackendArgs()
        dbi, dbo, ibi, ibo = z(B, C, 2, Fq), new(B, C, 2, Fq), z(B, S, 2 * Fq), new(B, S, 2 * Fq)
        ws = new(B * (T + 1) * S * 2 * Fq)
        ba.x, ba.deconv_buf_in, ba.deconv_buf_out, ba.istft_buf_in, ba.istft_buf_out = (t.data_ptr() for t in (x, dbi, dbo, ibi, ibo))
        ba.w, ba.bias = P["tfgridnet.deconv.weight"].data_ptr(), P["tfgridnet.deconv.bias"].data_ptr()
        ba.filt, ba.mask_spec = P["tfgridnet.dec.filterbank._filters"].data_ptr(), _ptr(spec)
        ba.wave_out, ba.ws = out.data_ptr(), ws.data_ptr()
        ba.B, ba.T, ba.F, ba.C, ba.n_src, ba.n_fft, ba.stride = B, T, Fq, C, S, cfg.n_fft, hop
        self._call(lib.sb_backend_fwd, ba, wave, "sb_backend_fwd")
        ctx["x_last"] = x

        # the state a streaming call would carry on from (TFGridNet.forward returns it in training too, DE3:547-552): same
        # keys / shapes / order as init_buffers; detached copies (the backward pass consumes the saved buffers)
        NI, H = B * Fq * T, cfg.H
        feats2 = feats if T >= 2 else torch.cat([torch.zeros(B, 2 - T, Fq, cfg.conv_in_ch, device=dev), feats], dim=1)
        state = {"conv_buf": feats2[:, -2:].permute(0, 3, 1, 2).contiguous(), "deconv_buf": dbo, "istft_buf": ibo.view(B, S, 2 * Fq, 1),
                 "gridnet_bufs": {}}
        off_c = 2 * NI * C + (NI + 3) // 4 * 4 + NI * 4 * H            # layout of sb_path_train_args.saved (inter: one direction)
        for i in range(cfg.B):
            sv = ctx["saved"][i][1]
            last = lambda o: sv[o:o + NI * H].view(B * Fq, T, H)[:, -1].clone().unsqueeze(0)
            buf = {}
            if cfg.use_attn:                        # layout of sb_attn_train_args.saved: zq, zk, zv, qn, kn, vn, ... (each padded to 4 floats)
                al4 = lambda n: (n + 3) // 4 * 4
                LE, W1, sa = cfg.L * cfg.attn_E, cfg.local_atten_len - 1, ctx["saved"][i][2]
                o_kn = 2 * al4(NP * LE) + al4(NP * C) + al4(NP * LE)
                o_vn = o_kn + al4(NP * LE)
                for key, o, width in (("K_buf", o_kn, Fq * cfg.attn_E), ("V_buf", o_vn, Fq * (C // cfg.L))):
                    rows = sa[o:o + B * cfg.L * T * width].view(B * cfg.L, T, width)
                    buf[key] = torch.cat([torch.zeros(B * cfg.L, W1, width, device=dev), rows], dim=1)[:, -W1:].contiguous() if W1 > 0 \
                        else rows[:, :0].contiguous()
            buf["c0"], buf["h0"] = last(off_c), last(off_c + NI * H)
            state["gridnet_bufs"][f"buf{i}"] = buf
        ctx["next_state"] = state
        return out, ctx

    # -- backward --------------------------------------------------------------------------------------------
    def backward(self, ctx, g_out: torch.Tensor) -> Dict[str, torch.Tensor]:
        """dL/d(output) [B, S, stride*T] -> {parameter name: gradient}.  Consumes ctx (one backward per forward)."""
        cfg, lib, P = self.cfg, self.lib, ctx["P"]
        B, T = ctx["B"], ctx["T"]
        hop, Fq, C, S = cfg.stft_chunk_size, cfg.n_freqs, cfg.D, cfg.num_src
        dev = g_out.device
        g_out = g_out.contiguous().float()
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        G: Dict[str, torch.Tensor] = {}

        def grad(name):
            G[name] = torch.zeros_like(P[name])
            return G[name]

        # back-end
        gx = new(B, T, Fq, C)
        bb = abi.BackendBwdArgs()
        ws = new(B * T * S * 2 * Fq)
        bb.x, bb.g_wave, bb.w, bb.filt = ctx["x_last"].data_ptr(), g_out.data_ptr(), P["tfgridnet.deconv.weight"].data_ptr(), P["tfgridnet.dec.filterbank._filters"].data_ptr()
        bb.mask_spec, bb.gx, bb.ws = _ptr(ctx["spec"]), gx.data_ptr(), ws.data_ptr()
        bb.g_w, bb.g_bias = grad("tfgridnet.deconv.weight").data_ptr(), grad("tfgridnet.deconv.bias").data_ptr()
        bb.B, bb.T, bb.F, bb.C, bb.n_src, bb.n_fft, bb.stride = B, T, Fq, C, S, cfg.n_fft, hop
        self._call(lib.sb_backend_bwd, bb, g_out, "sb_backend_bwd")
        del ws

        # blocks, last to first
        n_ws = [int(lib.sb_path_bwd_workspace_floats(B, T, Fq, C, cfg.H, k)) for k in (0, 1)]
        n_ws.append(B * T * Fq * C)
        if cfg.conv_lstm:
            n_ws.append(int(lib.sb_convpath_bwd_workspace_floats(B, T, Fq, C, cfg.H, cfg.lstm_down)))
        wsp = new(max(n_ws))
        g_film = torch.zeros(cfg.B - 1, 2, B, Fq, C, dtype=torch.float32, device=dev) if self.has_film else None
        for i in reversed(range(cfg.B)):
            b = f"tfgridnet.blocks.{i}."
            if cfg.use_attn:
                ab = abi.AttnBwdArgs()
                ab.f = self._attn_args(P, i, B, T)
                saved = ctx["saved"][i][2]
                ab.f.saved, ab.f.x = saved.data_ptr(), ctx["attn_in"][i].data_ptr()
                wsa = new(int(lib.sb_attn_bwd_workspace_floats(ctypes.byref(ab.f))))
                ab.gy, ab.gx, ab.ws = gx.data_ptr(), gx.data_ptr(), wsa.data_ptr()
                for field, mod in self._ATTN:
                    gp = getattr(ab, "g" + field)
                    for f2, name in self._ATTN_P:
                        setattr(gp, f2, grad(b + mod + name).data_ptr())
                self._call(lib.sb_attn_bwd, ab, g_out, "sb_attn_bwd")
                ctx["saved"][i][2] = None
                ctx["attn_in"][i] = None
                del saved, wsa
            for inter in (True, False):
                kind = "inter" if inter else "intra"
                if cfg.conv_lstm and not inter:
                    cb = abi.ConvPathBwdArgs()
                    cb.f = self._convpath_args(P, i, B, T)
                    saved = ctx["saved"][i][0]
                    cb.f.saved, cb.f.x = saved.data_ptr(), ctx["intra_in"][i].data_ptr()
                    cb.gy, cb.gx, cb.ws = gx.data_ptr(), gx.data_ptr(), wsp.data_ptr()
                    cb.g_conv_w, cb.g_conv_b, cb.g_prelu = (grad(b + n).data_ptr() for n in ("conv.weight", "conv.bias", "act.weight"))
                    cb.g_ln_g, cb.g_ln_b = grad(b + "norm.norm.weight").data_ptr(), grad(b + "norm.norm.bias").data_ptr()
                    for d, sfx in enumerate(("", "_reverse")):
                        g4 = [grad(b + "intra_rnn." + n + sfx) for n in _LSTM]
                        cb.g_w_ih[d], cb.g_w_hh[d], cb.g_b_ih[d], cb.g_b_hh[d] = (t.data_ptr() for t in g4)
                    cb.g_deconv_w, cb.g_deconv_b = grad(b + "deconv.weight").data_ptr(), grad(b + "deconv.bias").data_ptr()
                    self._call(lib.sb_intra_convlstm_bwd, cb, g_out, "sb_intra_convlstm_bwd")
                    ctx["saved"][i][0] = None
                    ctx["intra_in"][i] = None
                    del saved
                    continue
                pb = abi.PathBwdArgs()
                pb.f = self._path_args(P, i, inter, B, T)
                saved = ctx["saved"][i][int(inter)]
                pb.f.saved = saved.data_ptr()
                pb.f.x = gx.data_ptr()              # not read by the backward pass (any valid pointer)
                pb.gy, pb.gx, pb.ws = gx.data_ptr(), gx.data_ptr(), wsp.data_ptr()
                pb.g_ln_g, pb.g_ln_b = grad(b + kind + "_norm.norm.weight").data_ptr(), grad(b + kind + "_norm.norm.bias").data_ptr()
                for d, sfx in enumerate(("",) if inter else ("", "_reverse")):
                    g4 = [grad(b + kind + "_rnn." + n + sfx) for n in _LSTM]
                    pb.g_w_ih[d], pb.g_w_hh[d], pb.g_b_ih[d], pb.g_b_hh[d] = (t.data_ptr() for t in g4)
                pb.g_lin_w, pb.g_lin_b = grad(b + kind + "_linear.weight").data_ptr(), grad(b + kind + "_linear.bias").data_ptr()
                fn = lib.sb_inter_lstm_bwd if inter else lib.sb_intra_lstm_bwd
                self._call(fn, pb, g_out, "sb_%s_lstm_bwd" % kind)
                ctx["saved"][i][int(inter)] = None
                del saved
            if i > 0 and g_film is not None:
                f = abi.FilmApplyArgs()
                x_in = ctx["film_in"][i - 1]
                f.x, f.film_scale, f.film_shift = x_in.data_ptr(), ctx["film"][i - 1, 0].data_ptr(), ctx["film"][i - 1, 1].data_ptr()
                f.gy, f.gx = gx.data_ptr(), gx.data_ptr()
                f.g_scale, f.g_shift = g_film[i - 1, 0].data_ptr(), g_film[i - 1, 1].data_ptr()
                f.B, f.T, f.F, f.C = B, T, Fq, C
                self._call(lib.sb_film_apply_bwd, f, g_out, "sb_film_apply_bwd")
                ctx["film_in"][i - 1] = None

        # distance embedding + FiLM parameter nets
        if g_film is not None:
            st = ctx["film_stacks"]
            gs = {k: torch.zeros_like(v) for k, v in st.items()}
            fb = abi.FilmBwdArgs()
            fb.f = self._film_args(P, ctx["dis"], st)
            fb.g_film = g_film.data_ptr()
            fb.g_emb_w = grad("tfgridnet.embed_net.dis_embedding.0.weight").data_ptr()
            fb.g_emb_ln_g, fb.g_emb_ln_b = (grad(n).data_ptr() for n in self._emb_ln_names())
            fb.g_w_w, fb.g_w_b, fb.g_b_w, fb.g_b_b = (gs[k].data_ptr() for k in ("w_w", "w_b", "b_w", "b_b"))
            self._call(lib.sb_film_params_bwd, fb, g_out, "sb_film_params_bwd")
            e = "tfgridnet.embeds.%d."
            for j in range(cfg.B - 1):
                G[e % j + "weight.weight"] = gs["w_w"][j].unsqueeze(-1)
                G[e % j + "weight.bias"] = gs["w_b"][j]
                G[e % j + "bias.weight"] = gs["b_w"][j].unsqueeze(-1)
                G[e % j + "bias.bias"] = gs["b_b"][j]

        # conv-in
        ci = abi.ConvInTrainArgs()
        ci.feats, ci.w, ci.bias = ctx["feats"].data_ptr(), P["tfgridnet.conv.0.weight"].data_ptr(), P["tfgridnet.conv.0.bias"].data_ptr()
        ci.gx, ci.ws = gx.data_ptr(), wsp.data_ptr()
        if wsp.numel() < B * T * Fq * C:
            raise RuntimeError("workspace too small")
        ci.g_w, ci.g_bias = grad("tfgridnet.conv.0.weight").data_ptr(), grad("tfgridnet.conv.0.bias").data_ptr()
        if cfg.use_first_ln:
            ci.ln_g, ci.ln_b, ci.saved = P["tfgridnet.conv.1.weight"].data_ptr(), P["tfgridnet.conv.1.bias"].data_ptr(), ctx["conv_saved"].data_ptr()
            ci.g_ln_g, ci.g_ln_b = grad("tfgridnet.conv.1.weight").data_ptr(), grad("tfgridnet.conv.1.bias").data_ptr()
        ci.B, ci.T, ci.F, ci.Cin, ci.C = B, T, Fq, cfg.conv_in_ch, C
        self._call(lib.sb_conv_in_bwd, ci, g_out, "sb_conv_in_bwd")
        ctx.clear()
        return G


class SeparatorFunction(torch.autograd.Function):
    """autograd node around TrainGraph: inputs (mixture, dis_embed) carry no gradient, the parameters do."""

    @staticmethod
    def forward(ctx, graph: TrainGraph, names: List[str], wave, dis, *tensors):
        P = {n: (t.detach() if t.is_contiguous() else t.detach().contiguous()) for n, t in zip(names, tensors)}
        out, saved = graph.forward(P, wave.detach(), None if dis is None else dis.detach())
        graph.last_state = saved.pop("next_state")
        ctx.graph, ctx.names, ctx.saved = graph, names, saved
        ctx.needs = [t.requires_grad for t in tensors]
        return out

    @staticmethod
    def backward(ctx, g_out):
        if not ctx.saved:
            raise RuntimeError("the training graph of this forward call was already back-propagated (retain_graph is not supported)")
        G = ctx.graph.backward(ctx.saved, g_out)
        grads = [G.get(n) if need else None for n, need in zip(ctx.names, ctx.needs)]
        return (None, None, None, None, *grads)


def differentiable_forward_with_state(lib, cfg: ModelConfig, named_tensors: Dict[str, torch.Tensor], wave, dis):
    """``named_tensors``: every state_dict entry (parameters with requires_grad, buffers without) by reference name.
    Returns (output with an autograd graph, next_state as detached tensors in the reference's schema)."""
    names = list(named_tensors.keys())
    graph = TrainGraph(lib, cfg)
    out = SeparatorFunction.apply(graph, names, wave, dis, *[named_tensors[n] for n in names])
    return out, graph.last_state


def differentiable_forward(lib, cfg: ModelConfig, named_tensors: Dict[str, torch.Tensor], wave, dis):
    return differentiable_forward_with_state(lib, cfg, named_tensors, wave, dis)[0]
