"""Chunk-by-chunk streaming inference with device-resident state and CUDA-graph replay.

Protocol of the reference's edge/causal_infer.py:15-47: every call feeds one window ``[B, M, chunk + lookahead]``
(the previous window rolled left by one chunk with ``chunk`` new samples appended), ``pad=False``, and the state dict is
threaded through.  Here the state lives in two device arenas that alternate (the conv / deconv / iSTFT / attention
histories must not be updated in place), the per-chunk launch sequence of sb_net_forward is captured once per parity
into a CUDA graph, and a chunk costs one H2D copy, one graph launch and one D2H copy.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .engine import init_state


def _clone_state(st):
    return {k: (_clone_state(v) if isinstance(v, dict) else v.clone()) for k, v in st.items()}


class StreamingSession:
    def __init__(self, net, batch_size: int, dis_embed: Optional[torch.Tensor] = None, use_graph: bool = True,
                 frames_per_call: int = 1):
        self.net = net
        self.cfg = cfg = net.cfg
        self.engine = net.engine()
        dev = self.engine.packed.flat.device
        self.device = dev
        self.B = batch_size
        self.frames = frames_per_call
        n_in = cfg.stft_chunk_size * frames_per_call + cfg.n_fft - cfg.stft_chunk_size
        self.x = torch.zeros(batch_size, cfg.num_ch, n_in, dtype=torch.float32, device=dev)
        self.y = torch.zeros(batch_size, cfg.num_src, cfg.stft_chunk_size * frames_per_call, dtype=torch.float32, device=dev)
        self.dis = None
        if cfg.variant == "dis_embed":
            if dis_embed is None:
                raise KeyError("dis_embed")
            self.dis = dis_embed.to(dev, torch.float32).contiguous().clone()
        self.film = self.engine.film_table(self.dis) if self.dis is not None else None      # time-invariant: once
        self.states = [init_state(cfg, batch_size, dev), init_state(cfg, batch_size, dev)]
        self.parity = 0
        self.graphs = None
        self.n_calls = 0
        if use_graph:
            self._capture()

    # ------------------------------------------------------------------------------------------------------
    def _step_eager(self, p: int):
        src, dst = self.states[p], self.states[p ^ 1]
        work = {k: (dict((kk, dict(vv)) for kk, vv in v.items()) if k == "gridnet_bufs" else v) for k, v in src.items()}
        self.engine.forward(self.x, self.dis, work, out=self.y, new_state=dst, film=self.film)

    def _capture(self):
        saved = [_clone_state(s) for s in self.states]
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                       # warm-up: workspaces, smem opt-ins, both parities
            self._step_eager(0)
            self._step_eager(1)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graphs = []
        for p in (0, 1):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_eager(p)
            self.graphs.append(g)
        torch.cuda.synchronize(self.device)
        for s, z in zip(self.states, saved):                # undo the warm-up / capture side effects
            self._copy_state(s, z)
        self.parity = 0

    @staticmethod
    def _copy_state(dst, src):
        for k, v in src.items():
            if isinstance(v, dict):
                StreamingSession._copy_state(dst[k], v)
            else:
                dst[k].copy_(v)

    # ------------------------------------------------------------------------------------------------------
    def reset(self):
        for s in self.states:
            for k, v in s.items():
                if isinstance(v, dict):
                    for b in v.values():
                        for t in b.values():
                            t.zero_()
                else:
                    v.zero_()
        self.parity = 0

    def load_state(self, state: dict):
        """Adopt a reference-layout state dict (e.g. one produced by Net.forward or by the reference module)."""
        self._copy_state(self.states[self.parity], state)

    @property
    def state(self) -> dict:
        """The current state in the reference's schema (views of the live arena)."""
        return self.states[self.parity]

    def step(self):
        """Run one call on whatever is in ``self.x``; result in ``self.y`` (both device-resident)."""
        if self.graphs is not None:
            self.graphs[self.parity].replay()
        else:
            self._step_eager(self.parity)
        self.parity ^= 1
        self.n_calls += 1
        return self.y

    def feed(self, window: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """window [B, M, chunk*frames + lookahead] (host or device) -> [B, S, chunk*frames].  With a host ``out``
        tensor the result is copied back and the call returns after it has landed."""
        self.x.copy_(window, non_blocking=True)
        self.step()
        if out is None:
            return self.y
        out.copy_(self.y, non_blocking=True)
        if not out.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()
        return out

    def launches_per_step(self) -> int:
        before = _lib.launch_count()
        self._step_eager(self.parity)       # eager twin of the captured step, then restore parity bookkeeping
        torch.cuda.synchronize(self.device)
        self.parity ^= 1
        return _lib.launch_count() - before
