"""Host driver of the C-ABI library: owns workspaces and state tensors, fills sb_net_io, calls sb_net_forward.

``Engine`` is handed a bound CDLL (see _lib.py) and works on whatever device its tensors live on; the product path
(``Net``) only ever gives it the sm_100a CUDA library and CUDA tensors.  Mirrors TFGridNet.forward / init_buffers
(DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py:403-421, 433-552, 696-720).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _abi as abi
from .packing import ModelConfig, PackedWeights


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0


def init_state(cfg: ModelConfig, batch: int, device) -> dict:
    """TFGridNet.init_buffers (DE3:403-421) + GridNetBlock.init_buffers (DE3:696-720): same keys, shapes, order."""
    Fq = cfg.n_freqs
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
    st = {"conv_buf": z(batch, cfg.conv_in_ch, 2, Fq),
          "deconv_buf": z(batch, cfg.D, 2, Fq),
          "istft_buf": z(batch, cfg.num_src, 2 * Fq, 1),
          "gridnet_bufs": {}}
    for i in range(cfg.B):
        buf = {}
        if cfg.use_attn:
            W = cfg.local_atten_len
            buf["K_buf"] = z(batch * cfg.L, W - 1, cfg.attn_E * Fq)
            buf["V_buf"] = z(batch * cfg.L, W - 1, (cfg.D // cfg.L) * Fq)
        buf["c0"] = z(1, batch * Fq, cfg.H)
        buf["h0"] = z(1, batch * Fq, cfg.H)
        st["gridnet_bufs"][f"buf{i}"] = buf
    return st


class Engine:
    def __init__(self, cdll, cfg: ModelConfig, packed: PackedWeights):
        self.lib = cdll
        self.cfg = cfg
        self.packed = packed
        self.intra_algo = abi.SB_ALGO_AUTO
        self.inter_algo = abi.SB_ALGO_AUTO
        self._ws: Dict[Tuple[int, int, str], torch.Tensor] = {}

    # -- helpers ---------------------------------------------------------------------------------------------
    def workspace(self, B: int, T: int, device) -> torch.Tensor:
        key = (B, T, str(device))
        ws = self._ws.get(key)
        if ws is None:
            n = self.lib.sb_workspace_floats(self.packed.desc_ref(), B, T)
            if len(self._ws) > 4:                           # eager calls only: sessions own their workspaces (a captured
                self._ws.clear()                            # graph keeps raw pointers, so it must never read this cache)
            ws = torch.empty(max(int(n), 1), dtype=torch.float32, device=device)
            self._ws[key] = ws
        return ws

    def n_frames(self, n_samples: int) -> int:
        cfg = self.cfg
        return (n_samples - cfg.n_fft) // cfg.stft_chunk_size + 1

    @staticmethod
    def _f32c(t: torch.Tensor, shape=None) -> torch.Tensor:
        if t.dtype != torch.float32:
            t = t.float()
        if not t.is_contiguous():
            t = t.contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError("state tensor has shape %s, expected %s" % (tuple(t.shape), tuple(shape)))
        return t

    def film_table(self, dis_embed: torch.Tensor) -> Optional[torch.Tensor]:
        """Dis_Embed_* + every FilmLayer's 1x1 convs (DE3:51-68,114-173) for these rows: [n_blocks-1, 2, B, F, C].
        Time-invariant, so a streaming session computes it once and hands it to every chunk."""
        cfg, pk = self.cfg, self.packed
        if cfg.variant != "dis_embed" or cfg.B < 2:
            return None
        dis = self._f32c(dis_embed)
        B = dis.shape[0]
        film = torch.empty(cfg.B - 1, 2, B, cfg.n_freqs, cfg.D, dtype=torch.float32, device=dis.device)
        a = abi.FilmArgs()
        a.dis, a.emb_w, a.emb_ln_g, a.emb_ln_b = dis.data_ptr(), pk.ptr("emb_w"), pk.ptr("emb_ln_g"), pk.ptr("emb_ln_b")
        a.w_w, a.w_b, a.b_w, a.b_b = pk.ptr("film_w_w"), pk.ptr("film_w_b"), pk.ptr("film_b_w"), pk.ptr("film_b_b")
        a.film = film.data_ptr()
        a.B, a.F, a.C, a.Din, a.n_layers, a.emb_mode = B, cfg.n_freqs, cfg.D, cfg.film_in, cfg.B - 1, pk.desc.emb_mode
        abi.check(self.lib, self.lib.sb_film_params_fwd(ctypes.byref(a), _stream_ptr(dis)), "sb_film_params_fwd")
        return film

    # -- the forward pass ------------------------------------------------------------------------------------
    def prepare(self, wave: torch.Tensor, dis_embed: Optional[torch.Tensor], state: dict,
                out: Optional[torch.Tensor] = None, new_state: Optional[dict] = None,
                film: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                intra_algo: Optional[int] = None, inter_algo: Optional[int] = None) -> "PreparedCall":
        """Validates one call and fills its sb_net_io (no launch).  wave [B, M, stride*T + n_fft - stride]; the result
        goes to `out` [B, S, stride*T]; the next state is written into fresh tensors unless `new_state` supplies them."""
        cfg = self.cfg
        B, M, N = wave.shape
        if M != cfg.num_ch:
            raise ValueError("mixture has %d channels, the model was built for num_ch=%d" % (M, cfg.num_ch))
        T = self.n_frames(N)
        if T < 1:
            raise ValueError("input of %d samples is shorter than one window (%d)" % (N, cfg.n_fft))
        need = cfg.stft_chunk_size * T + cfg.n_fft - cfg.stft_chunk_size
        dev = wave.device
        wave = self._f32c(wave if N == need else wave[..., :need])
        Fq, S = cfg.n_freqs, cfg.num_src

        io = abi.NetIO()
        io.B, io.T = B, T
        io.intra_algo = self.intra_algo if intra_algo is None else intra_algo
        io.inter_algo = self.inter_algo if inter_algo is None else inter_algo
        io.wave = wave.data_ptr()
        keep = [wave]
        if cfg.variant == "dis_embed":
            if dis_embed is None:
                raise KeyError("dis_embed")
            dis = self._f32c(dis_embed.to(dev), (B, 3))
            io.dis_embed = dis.data_ptr()
            keep.append(dis)
            if film is not None:
                io.film = film.data_ptr()
                keep.append(film)
        if out is None:
            out = torch.empty(B, S, cfg.stft_chunk_size * T, dtype=torch.float32, device=dev)
        io.wave_out = out.data_ptr()

        ns = new_state if new_state is not None else {}
        def fresh(name, like):
            t = ns.get(name)
            if t is None:
                t = torch.empty_like(like)
            return t
        conv_in = self._f32c(state["conv_buf"], (B, cfg.conv_in_ch, 2, Fq))
        deconv_in = self._f32c(state["deconv_buf"], (B, cfg.D, 2, Fq))
        istft_in = self._f32c(state["istft_buf"], (B, S, 2 * Fq, 1))
        conv_out, deconv_out, istft_out = fresh("conv_buf", conv_in), fresh("deconv_buf", deconv_in), fresh("istft_buf", istft_in)
        io.conv_buf_in, io.conv_buf_out = conv_in.data_ptr(), conv_out.data_ptr()
        io.deconv_buf_in, io.deconv_buf_out = deconv_in.data_ptr(), deconv_out.data_ptr()
        io.istft_buf_in, io.istft_buf_out = istft_in.data_ptr(), istft_out.data_ptr()
        keep += [conv_in, deconv_in, istft_in]
        bufs = state["gridnet_bufs"]
        nbufs = ns.get("gridnet_bufs", {})
        outs = {}
        for i in range(cfg.B):
            buf = bufs[f"buf{i}"]
            nb = nbufs.get(f"buf{i}", {})
            h_in = self._f32c(buf["h0"], (1, B * Fq, cfg.H))
            c_in = self._f32c(buf["c0"], (1, B * Fq, cfg.H))
            h_out = nb.get("h0") if nb.get("h0") is not None else torch.empty_like(h_in)
            c_out = nb.get("c0") if nb.get("c0") is not None else torch.empty_like(c_in)
            io.h_in[i], io.h_out[i] = h_in.data_ptr(), h_out.data_ptr()
            io.c_in[i], io.c_out[i] = c_in.data_ptr(), c_out.data_ptr()
            keep += [h_in, c_in]
            o = {"h0": h_out, "c0": c_out}
            if cfg.use_attn:
                W = cfg.local_atten_len
                k_in = self._f32c(buf["K_buf"], (B * cfg.L, W - 1, cfg.attn_E * Fq))
                v_in = self._f32c(buf["V_buf"], (B * cfg.L, W - 1, (cfg.D // cfg.L) * Fq))
                k_out = nb.get("K_buf") if nb.get("K_buf") is not None else torch.empty_like(k_in)
                v_out = nb.get("V_buf") if nb.get("V_buf") is not None else torch.empty_like(v_in)
                io.K_in[i], io.K_out[i] = k_in.data_ptr(), k_out.data_ptr()
                io.V_in[i], io.V_out[i] = v_in.data_ptr(), v_out.data_ptr()
                keep += [k_in, v_in]
                o["K_buf"], o["V_buf"] = k_out, v_out
            outs[i] = o
        ws = workspace if workspace is not None else self.workspace(B, T, dev)
        need_ws = int(self.lib.sb_workspace_floats(self.packed.desc_ref(), B, T))
        if ws.numel() < need_ws or ws.dtype != torch.float32 or ws.device != dev:
            raise ValueError("workspace must be %d float32 on %s" % (need_ws, dev))
        io.workspace = ws.data_ptr()
        keep.append(ws)
        return PreparedCall(self, io, keep, out, state, {"conv_buf": conv_out, "deconv_buf": deconv_out,
                                                          "istft_buf": istft_out}, outs)

    def forward(self, wave: torch.Tensor, dis_embed: Optional[torch.Tensor], state: dict,
                out: Optional[torch.Tensor] = None, new_state: Optional[dict] = None,
                film: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
        """wave [B, M, stride*T + n_fft - stride] -> ([B, S, stride*T], state).  `state` is updated in place (the
        dict, as the reference does, DE3:547-552) with freshly written tensors unless `new_state` supplies them."""
        call = self.prepare(wave, dis_embed, state, out=out, new_state=new_state, film=film, workspace=workspace)
        call.launch()
        return call.out, call.commit()


class PreparedCall:
    """One validated call of the forward pass: its sb_net_io and the tensors it points at.  `launch()` runs the whole
    launch sequence (sb_net_forward) or a range of its units (sb_net_forward_range) on the current stream."""

    def __init__(self, engine: Engine, io, keep, out, state, top, blocks):
        self.engine, self.io, self.keep, self.out = engine, io, keep, out
        self._state, self._top, self._blocks = state, top, blocks
        self.n_units = 2 * engine.cfg.B + 2           # front-end, (intra, inter) per block, back-end

    def launch(self, first_unit: int = 0, last_unit: Optional[int] = None):
        eng = self.engine
        last = self.n_units - 1 if last_unit is None else last_unit
        stream = _stream_ptr(self.keep[0])
        if first_unit == 0 and last == self.n_units - 1:
            rc = eng.lib.sb_net_forward(eng.packed.desc_ref(), ctypes.byref(self.io), stream)
        else:
            rc = eng.lib.sb_net_forward_range(eng.packed.desc_ref(), ctypes.byref(self.io), first_unit, last, stream)
        abi.check(eng.lib, rc, "sb_net_forward")

    def commit(self) -> dict:
        """Point the caller's state dict at the tensors this call wrote (the reference mutates the dict it was given)."""
        state = self._state
        for k, v in self._top.items():
            state[k] = v
        bufs = state["gridnet_bufs"]
        for i, o in self._blocks.items():
            for k, v in o.items():
                bufs[f"buf{i}"][k] = v
        return state
