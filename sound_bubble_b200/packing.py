"""Checkpoint tensors (reference ``state_dict`` layout) -> the packed device buffers the kernels read.

Everything here is layout plumbing (permute / cat of the ~0.5 M parameters); no arithmetic of the forward pass
happens on this side.  Reference files: DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/{net.py,
tfgridnet_causal.py}, OPT = src/models/tfgridnet_realtime_clean_optim/{net.py, tfgridnet_causal.py}.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _abi as abi


@dataclass
class ModelConfig:
    """Net.__init__ kwargs (DE3/net.py:21-26, OPT/net.py:21-26) and the sizes TFGridNet.__init__ derives from them."""
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
    lstm_down: Optional[int] = None     # OPT exposes it (default 5); DE3's core default is 4 (DE3:282)

    def __post_init__(self):
        if self.variant not in ("dis_embed", "optim"):
            raise ValueError("variant must be 'dis_embed' or 'optim'")
        if self.lstm_down is None:
            self.lstm_down = 5 if self.variant == "optim" else 4
        if self.fb_type != "stft":
            raise NotImplementedError("only fb_type='stft' is implemented (the only one the reference configs use)")
        if self.stft_back_pad != 0:
            raise NotImplementedError("stft_back_pad > 0 (causal_decoder, DE3:423-431) is not implemented")
        if self.merge_method not in ("None", "early_cat"):
            raise NotImplementedError("merge_method %r (DE3:334-347 knows 'None' and 'early_cat')" % self.merge_method)
        if self.H != 64 or self.D not in (16, 32):
            raise NotImplementedError(
                "the sm_100a kernels are built for H=64 and D in {16, 32} (every shipped reference config: "
                "syn_experiments/*.json D=32, real_experiments/*.json D=16/32, all H=64); got D=%d, H=%d" % (self.D, self.H))
        if not (1 <= self.num_ch <= 8) or not (1 <= self.num_src <= 2):
            raise NotImplementedError("1..8 microphones and 1..2 sources are supported, got num_ch=%d, num_src=%d"
                                      % (self.num_ch, self.num_src))
        assert self.n_fft % 2 == 0                                  # DE3:307
        if self.use_attn:
            assert self.D % self.L == 0                             # DE3:641

    @property
    def n_fft(self):                    # DE3/net.py:38
        return self.stft_back_pad + self.stft_chunk_size + self.stft_pad_size

    @property
    def n_freqs(self):                  # DE3:308
        return self.n_fft // 2 + 1

    @property
    def feat_num(self):                 # DE3:335-338
        if self.merge_method != "early_cat":
            return 0
        return (self.num_ch - 1) * 3 - (1 if self.directional else 0)

    @property
    def conv_in_ch(self):               # DE3:342-347
        return 2 * self.num_ch + self.feat_num

    @property
    def film_in(self):                  # DE3:356-374
        if self.variant != "dis_embed":
            return 0
        return {"linear1": 1, "linear2": self.D, "conv1": 1, "conv2": 2, "conv3": 4, "conv4": 8}[self.dis_type]

    @property
    def attn_E(self):                   # DE3:591-593 with approx_qk_dim = E * n_freqs (net.py:53)
        return math.ceil(self.E * self.n_freqs * 1.0 / self.n_freqs)

    @property
    def lstm_steps(self):               # conv-LSTM sequence length (DE3:800-803)
        return (self.n_freqs - self.lstm_down) // self.lstm_down + 1


def _lstm_dir(sd, prefix: str, sfx: str, H: int) -> Dict[str, torch.Tensor]:
    """[W_ih | W_hh] in the two operand layouts of sb_lstm.cu (see sb_lstm_dir in include/soundbubble.h)."""
    w = torch.cat([sd[prefix + "weight_ih_l0" + sfx], sd[prefix + "weight_hh_l0" + sfx]], dim=1).float()   # [4H, K]
    b = (sd[prefix + "bias_ih_l0" + sfx] + sd[prefix + "bias_hh_l0" + sfx]).float()                       # [4H]
    K = w.shape[1]
    assert w.shape[0] == 4 * H and K % 4 == 0
    g = torch.arange(4).view(4, 1).expand(4, H).reshape(-1)
    u = torch.arange(H).view(1, H).expand(4, H).reshape(-1)
    col = (g // 2) * (2 * H) + 4 * (u // 2) + 2 * (g % 2) + (u % 2)          # row g*H+u -> tile column
    w_tile = torch.empty(K, 4 * H, dtype=torch.float32, device=w.device)
    w_tile[:, col.to(w.device)] = w.t()
    b_tile = torch.empty(4 * H, dtype=torch.float32, device=w.device)
    b_tile[col.to(w.device)] = b
    slot = (4 * u + g).to(w.device)                                          # row g*H+u -> lane slot
    w_slot = torch.empty(4 * H, K, dtype=torch.float32, device=w.device)
    w_slot[slot] = w
    w_lane = w_slot.view(4 * H, K // 4, 4).permute(1, 0, 2).contiguous()     # [K/4][4H][4]
    b_lane = torch.empty(4 * H, dtype=torch.float32, device=w.device)
    b_lane[slot] = b
    # warp-specialised kernel: thread t = 4*ur + kq owns the four gates of units (ur, ur + 32) over a quarter of K
    C = K - H
    w_ih, w_hh = w[:, :C], w[:, C:]

    def two_unit(m, kpt):                                   # m [4H, Kx] -> [kpt][2][128][4] = (k, A|B, thread, gate)
        v = m.reshape(4, 2, H // 2, 4, kpt)                 # (g, ab, ur, kq, k): unit = ab*32 + ur, column = kq*kpt + k
        return v.permute(4, 1, 2, 3, 0).reshape(kpt, 2, 2 * H, 4).contiguous()
    return {"w_tile": w_tile, "b_tile": b_tile, "w_lane": w_lane, "b_lane": b_lane,
            "w_rec": two_unit(w_hh, H // 4), "w_xp": two_unit(w_ih, C // 4)}


def _bf16_split(m: torch.Tensor):
    """fp32 -> (hi, lo) bfloat16 pair with hi + lo == m to ~2^-17 relative (round-to-nearest both times)."""
    hi = m.to(torch.bfloat16)
    lo = (m - hi.float()).to(torch.bfloat16)
    return hi, lo


def _umma_kmajor(m: torch.Tensor) -> torch.Tensor:
    """[R, K] bf16 -> the canonical no-swizzle K-major UMMA shared-memory image [K/8][R/8][8 rows][8 k] (one 128-byte
    core matrix = 8 rows x 16 bytes; SBO = 128 B between 8-row groups, LBO = (R/8)*128 B between 8-element k chunks),
    returned as raw 16-bit words."""
    R, K = m.shape
    assert R % 8 == 0 and K % 8 == 0
    return m.view(R // 8, 8, K // 8, 8).permute(2, 0, 1, 3).contiguous().view(torch.int16).reshape(-1)


def _tc_operands(w: torch.Tensor, b: torch.Tensor, lin: torch.Tensor, H: int) -> Dict[str, torch.Tensor]:
    """Operands of lstm_tc_kernel (tcgen05): gate matrix [4H, K] with rows re-ordered unit-major (n = 4u + g), the
    projection [C, H], both split into bf16 hi/lo and laid out for UMMA; biases fp32 in the same n order.
    16-bit payloads are carried in the float32 weight buffer as raw bits."""
    g = torch.arange(4).view(1, 4).expand(H, 4).reshape(-1)
    u = torch.arange(H).view(H, 1).expand(H, 4).reshape(-1)
    rows = g * H + u                                             # n = 4u + g  <-  weight row g*H + u
    wn = w[rows].contiguous()
    whi, wlo = _bf16_split(wn)
    phi, plo = _bf16_split(lin.float().contiguous())
    words = torch.cat([_umma_kmajor(whi), _umma_kmajor(wlo), _umma_kmajor(phi), _umma_kmajor(plo)])
    return {"tc_w": words.view(torch.float32).clone(), "tc_b": b[rows].contiguous()}


def _proj_ws(lin: torch.Tensor, H: int) -> torch.Tensor:
    """lin [C, H] -> w_prj [4][128][4]: recurrence thread t = 4*ur + kq holds 16 weights for the h slice it has in
    registers (k = 16kq + j): lin[ur % C][16kq + j] where j belongs to its plane ur // C, zero elsewhere (the 32/C
    planes of an output channel are summed when the block is stored; sb_lstm.cu, lstm_ws_kernel)."""
    C = lin.shape[0]
    npl = 32 // C
    kpl = 16 // npl
    t = torch.arange(2 * H)
    ur, kq = t // 4, t % 4
    j = torch.arange(16)
    col = (16 * kq).view(-1, 1) + j.view(1, -1)                                   # [128, 16]
    p = lin.float()[(ur % C).view(-1, 1).expand(-1, 16), col]
    mask = (j.view(1, -1) // kpl) == (ur // C).view(-1, 1)
    p = p * mask.to(p.dtype)
    return p.view(2 * H, 4, 4).permute(1, 0, 2).contiguous()


class PackedWeights:
    """One flat float32 device buffer + the sb_net_desc that points into it."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: ModelConfig, device):
        self.cfg = cfg
        pieces: List[Tuple[str, torch.Tensor]] = []
        # re-pack on the host (0.5 M parameters, index plumbing only), then ship one flat buffer to the device
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}

        def add(name: str, t: torch.Tensor):
            pieces.append((name, t.detach().to(torch.float32).contiguous().reshape(-1)))

        p = "tfgridnet."
        C_, H, Fq = cfg.D, cfg.H, cfg.n_freqs
        nfft = cfg.n_fft
        add("enc_filt", sd[p + "enc.filterbank._filters"].reshape(nfft + 2, nfft))
        add("dec_filt", sd[p + "dec.filterbank._filters"].reshape(nfft + 2, nfft))
        add("conv_w_pack", sd[p + "conv.0.weight"].permute(2, 1, 3, 0))            # [o][c][kt][kf] -> (kt, c, kf, o)
        add("conv_bias", sd[p + "conv.0.bias"])
        if cfg.use_first_ln:
            add("conv_ln_g", sd[p + "conv.1.weight"])
            add("conv_ln_b", sd[p + "conv.1.bias"])
        if cfg.variant == "dis_embed":
            add("emb_w", sd[p + "embed_net.dis_embedding.0.weight"])
            if cfg.dis_type.startswith("conv"):
                add("emb_ln_g", sd[p + "embed_net.dis_norm.weight"])
                add("emb_ln_b", sd[p + "embed_net.dis_norm.bias"])
            else:
                add("emb_ln_g", sd[p + "embed_net.dis_embedding.1.weight"])
                add("emb_ln_b", sd[p + "embed_net.dis_embedding.1.bias"])
            if cfg.B > 1:
                add("film_w_w", torch.stack([sd[f"{p}embeds.{j}.weight.weight"][:, :, 0] for j in range(cfg.B - 1)]))
                add("film_w_b", torch.stack([sd[f"{p}embeds.{j}.weight.bias"] for j in range(cfg.B - 1)]))
                add("film_b_w", torch.stack([sd[f"{p}embeds.{j}.bias.weight"][:, :, 0] for j in range(cfg.B - 1)]))
                add("film_b_b", torch.stack([sd[f"{p}embeds.{j}.bias.bias"] for j in range(cfg.B - 1)]))
        add("deconv_w", sd[p + "deconv.weight"])
        add("deconv_bias", sd[p + "deconv.bias"])
        for i in range(cfg.B):
            b = f"{p}blocks.{i}."
            norm = "norm.norm." if cfg.conv_lstm else "intra_norm.norm."
            for d, sfx in enumerate(("", "_reverse")):
                for k, v in _lstm_dir(sd, b + "intra_rnn.", sfx, H).items():
                    add(f"b{i}.intra{d}.{k}", v)
                if not cfg.conv_lstm:
                    lin = sd[b + "intra_linear.weight"][:, d * H:(d + 1) * H]       # [C, H]
                    add(f"b{i}.intra{d}.lin_n", lin)
                    add(f"b{i}.intra{d}.lin_t", lin.t())
                    add(f"b{i}.intra{d}.w_prj", _proj_ws(lin, H))
                    for k, v in _tc_operands(*self._cat(sd, b + "intra_rnn.", sfx), lin, H).items():
                        add(f"b{i}.intra{d}.{k}", v)
            if not cfg.conv_lstm:
                add(f"b{i}.intra.lin_b", sd[b + "intra_linear.bias"])
            add(f"b{i}.intra.ln_g", sd[b + norm + "weight"])
            add(f"b{i}.intra.ln_b", sd[b + norm + "bias"])
            for k, v in _lstm_dir(sd, b + "inter_rnn.", "", H).items():
                add(f"b{i}.inter.{k}", v)
            add(f"b{i}.inter.lin_n", sd[b + "inter_linear.weight"])
            add(f"b{i}.inter.lin_t", sd[b + "inter_linear.weight"].t())
            add(f"b{i}.inter.w_prj", _proj_ws(sd[b + "inter_linear.weight"], H))
            for k, v in _tc_operands(*self._cat(sd, b + "inter_rnn.", ""), sd[b + "inter_linear.weight"], H).items():
                add(f"b{i}.inter.{k}", v)
            add(f"b{i}.inter.lin_b", sd[b + "inter_linear.bias"])
            add(f"b{i}.inter.ln_g", sd[b + "inter_norm.norm.weight"])
            add(f"b{i}.inter.ln_b", sd[b + "inter_norm.norm.bias"])
            if cfg.conv_lstm:
                add(f"b{i}.cl_conv_w", sd[b + "conv.weight"].permute(2, 1, 0))     # [o][c][j] -> (j, c, o)
                add(f"b{i}.cl_conv_b", sd[b + "conv.bias"])
                add(f"b{i}.cl_prelu", sd[b + "act.weight"])
                k = cfg.lstm_down
                dw = sd[b + "deconv.weight"].view(2, H, C_, k).permute(0, 3, 1, 2)  # [d*H+u][c][j] -> (d, j, u, c)
                add(f"b{i}.cl_deconv_w", dw)
                add(f"b{i}.cl_deconv_b", sd[b + "deconv.bias"])
            if cfg.use_attn:
                for nm in ("Q", "K", "V"):
                    a = b + f"attn_conv_{nm}."
                    add(f"b{i}.attn_{nm}.w", sd[a + "0.weight"])
                    add(f"b{i}.attn_{nm}.b", sd[a + "0.bias"])
                    add(f"b{i}.attn_{nm}.prelu", sd[a + "1.weight"])
                    add(f"b{i}.attn_{nm}.ln_g", sd[a + "3.norm.weight"])
                    add(f"b{i}.attn_{nm}.ln_b", sd[a + "3.norm.bias"])
                a = b + "attn_concat_proj."
                add(f"b{i}.attn_O.w", sd[a + "0.weight"])
                add(f"b{i}.attn_O.b", sd[a + "0.bias"])
                add(f"b{i}.attn_O.prelu", sd[a + "1.weight"])
                add(f"b{i}.attn_O.ln_g", sd[a + "3.norm.weight"])
                add(f"b{i}.attn_O.ln_b", sd[a + "3.norm.bias"])

        # lay the pieces out at 256-byte granules in one buffer
        offsets, total = {}, 0
        for name, t in pieces:
            offsets[name] = total
            total += (t.numel() + 63) // 64 * 64
        flat = torch.zeros(total, dtype=torch.float32, device=pieces[0][1].device)
        for name, t in pieces:
            flat[offsets[name]: offsets[name] + t.numel()] = t
        self.flat = flat.to(device)
        self.offsets = offsets
        base = self.flat.data_ptr()
        assert base % 16 == 0

        def ptr(name: str):
            return base + 4 * offsets[name] if name in offsets else None

        self.ptr = ptr
        d = abi.NetDesc()
        d.M, d.n_fft, d.stride, d.F = cfg.num_ch, nfft, cfg.stft_chunk_size, Fq
        d.C, d.H, d.n_blocks, d.n_src = C_, H, cfg.B, cfg.num_src
        if cfg.merge_method == "early_cat":
            d.feat_mode = abi.SB_FEAT_DIRECTIONAL if cfg.directional else abi.SB_FEAT_OMNI
        else:
            d.feat_mode = abi.SB_FEAT_NONE
        d.Cin = cfg.conv_in_ch
        d.film_din = cfg.film_in
        d.emb_mode = abi.SB_EMB_CONV if cfg.dis_type.startswith("conv") else abi.SB_EMB_LINEAR
        d.spectral_masking = int(cfg.spectral_masking)
        d.conv_lstm, d.lstm_down = int(cfg.conv_lstm), cfg.lstm_down
        d.tail_mode = abi.SB_CONVLSTM_OUTPAD if cfg.variant == "optim" else abi.SB_CONVLSTM_PADCROP
        d.use_attn, d.L, d.E, d.W = int(cfg.use_attn), cfg.L, cfg.attn_E, cfg.local_atten_len
        for f in ("enc_filt", "dec_filt", "conv_w_pack", "conv_bias", "conv_ln_g", "conv_ln_b", "emb_w", "emb_ln_g",
                  "emb_ln_b", "film_w_w", "film_w_b", "film_b_w", "film_b_b", "deconv_w", "deconv_bias"):
            setattr(d, f, ptr(f))
        for i in range(cfg.B):
            bd = d.blocks[i]
            for dd in range(2):
                self._fill_dir(bd.intra[dd], f"b{i}.intra{dd}.", f"b{i}.intra.")
            self._fill_dir(bd.inter, f"b{i}.inter.", f"b{i}.inter.")
            for f in ("cl_conv_w", "cl_conv_b", "cl_prelu", "cl_deconv_w", "cl_deconv_b"):
                setattr(bd, f, ptr(f"b{i}.{f}"))
            for nm, fld in (("Q", bd.attn_q), ("K", bd.attn_k), ("V", bd.attn_v), ("O", bd.attn_o)):
                for f in ("w", "b", "prelu", "ln_g", "ln_b"):
                    setattr(fld, f, ptr(f"b{i}.attn_{nm}.{f}"))
        self.desc = d

    @staticmethod
    def _cat(sd, prefix: str, sfx: str):
        w = torch.cat([sd[prefix + "weight_ih_l0" + sfx], sd[prefix + "weight_hh_l0" + sfx]], dim=1).float()
        b = (sd[prefix + "bias_ih_l0" + sfx] + sd[prefix + "bias_hh_l0" + sfx]).float()
        return w, b

    def _fill_dir(self, dst, own: str, shared: str):
        for f in ("w_tile", "b_tile", "w_lane", "b_lane", "w_rec", "w_xp", "w_prj", "tc_w", "tc_b", "lin_t", "lin_n"):
            setattr(dst, f, self.ptr(own + f))
        for f in ("lin_b", "ln_g", "ln_b"):
            v = self.ptr(own + f)
            setattr(dst, f, v if v is not None else self.ptr(shared + f))

    def lstm_dir(self, block: int, which: str) -> abi.LstmDir:
        bd = self.desc.blocks[block]
        return {"intra0": bd.intra[0], "intra1": bd.intra[1], "inter": bd.inter}[which]

    def desc_ref(self):
        return ctypes.byref(self.desc)
