"""Chunk-by-chunk streaming inference with device-resident state and CUDA-graph replay.

Protocol of the reference's edge/causal_infer.py:15-47: every call feeds one window ``[B, M, chunk + lookahead]``
(the previous window rolled left by one chunk with ``chunk`` new samples appended), ``pad=False``, and the state dict is
threaded through.  Here the state lives in two device arenas that alternate (the conv / deconv / iSTFT / attention
histories must not be updated in place), the per-chunk launch sequence of sb_net_forward is captured once per parity
into a CUDA graph, and a chunk costs one H2D copy, one graph launch and one D2H copy.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _abi as abi
from . import _lib
from .engine import init_state
from .state_io import StateArena


def _clone_state(st):
    return {k: (_clone_state(v) if isinstance(v, dict) else v.clone()) for k, v in st.items()}


def _shallow(st):
    """A throw-away copy of the dict structure (Engine.forward re-points the dict it is given at the new tensors)."""
    return {k: (dict((kk, dict(vv)) for kk, vv in v.items()) if k == "gridnet_bufs" else v) for k, v in st.items()}


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
        # the session's own workspace: the captured graphs hold its raw pointer, so it must not come from (and be evicted
        # with) the engine's cache of eager-call workspaces
        n_ws = max(int(self.engine.lib.sb_workspace_floats(self.engine.packed.desc_ref(), batch_size, frames_per_call)), 1)
        self.ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
        # two state arenas that alternate; each is ONE flat device buffer with the reference-layout dict as views into it
        self.arenas = [StateArena(init_state(cfg, batch_size, dev)) for _ in (0, 1)]
        self.states = [a.state for a in self.arenas]
        self.parity = 0
        self.graphs = None
        self.n_calls = 0
        if use_graph:
            self._capture()

    # ------------------------------------------------------------------------------------------------------
    def _step_eager(self, p: int):
        src, dst = self.states[p], self.states[p ^ 1]
        self.engine.forward(self.x, self.dis, _shallow(src), out=self.y, new_state=dst, film=self.film, workspace=self.ws)

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
        for a in self.arenas:
            a.flat.zero_()
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


class PipelinedSession:
    """Throughput mode of the same protocol: one call per chunk, state carried, but the calls are ASYNCHRONOUS and
    consecutive chunks overlap on the GPU.

    A chunk at batch 32 fills 64 of the 148 SMs for most of its life (one CTA per (utterance, direction) of the
    intra-frame BiLSTM, 145 dependent steps), so a single in-order stream leaves half of the GPU idle.  Every state
    tensor of the reference belongs to exactly one unit of the launch sequence (conv_buf: front-end; h0/c0 [+K/V]: one
    GridNet block; deconv_buf/istft_buf: back-end; DE3:403-421, 696-720), hence chunk t+1 depends on chunk t PER UNIT
    only.  The session owns `depth` slots (window, result, workspace; chunk t uses slot t % depth, reads state arena
    t % 2 and writes the other) and hands them to the library's native pipe (csrc/sb_pipe.cu), which runs chunk t on
    stream t % depth as one CUDA graph per unit range, range j waiting for range j of chunk t-1.  Results are identical
    to StreamingSession's (same kernels, same order per unit).

    Use: ``begin()`` once after the caller's stream has produced the windows, ``feed(window, out)`` per chunk (returns
    immediately; `out` may be a pinned host tensor), ``end()`` to make the caller's stream wait for everything fed."""

    def __init__(self, net, batch_size: int, dis_embed: Optional[torch.Tensor] = None, ranges=None, depth: int = 8,
                 intra_algo: Optional[int] = None, inter_algo: Optional[int] = None, frames_per_call: int = 1,
                 group: int = 1):
        """group = G > 1: grouped throughput mode.  feed() still takes ONE 8 ms window per call, the native pipe gathers G
        consecutive windows and launches them as one G-frame call (`depth` groups in flight), so the intra-frame
        recurrences of G chunks share 128-row tcgen05 tiles.  The result of a chunk lands in its `out` once its group has
        run; a window must stay untouched until then (G calls later, flush() or end())."""
        if group < 1 or (group > 1 and frames_per_call != 1):
            raise ValueError("group must be >= 1 and excludes frames_per_call > 1")
        self.group = group
        if group > 1:
            frames_per_call = group
        self.net = net
        self.cfg = cfg = net.cfg
        self.engine = eng = net.engine()
        dev = eng.packed.flat.device
        self.device = dev
        self.B = batch_size
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.depth = depth
        mk = lambda *shape: [torch.zeros(*shape, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.frames = frames_per_call
        self.x = mk(batch_size, cfg.num_ch, cfg.stft_chunk_size * frames_per_call + cfg.n_fft - cfg.stft_chunk_size)
        self.y = mk(batch_size, cfg.num_src, cfg.stft_chunk_size * frames_per_call)
        n_ws = max(int(eng.lib.sb_workspace_floats(eng.packed.desc_ref(), batch_size, frames_per_call)), 1)
        self.ws = [torch.empty(n_ws, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.dis = None
        if cfg.variant == "dis_embed":
            if dis_embed is None:
                raise KeyError("dis_embed")
            self.dis = dis_embed.to(dev, torch.float32).contiguous().clone()
        self.film = eng.film_table(self.dis) if self.dis is not None else None
        self.arenas = [StateArena(init_state(cfg, batch_size, dev)) for _ in (0, 1)]
        self.states = [a.state for a in self.arenas]
        # units of sb_net_forward_range: 0 front-end, 1 + 2i / 2 + 2i intra / inter path of block i, 2B + 1 back-end
        n_units = 2 * cfg.B + 2
        if ranges is None:                                  # one range per unit: measured best on B200 with depth >= 8
            ranges = n_units                                # (profiles/r01_pipeline_sweep.txt)
        if isinstance(ranges, int) and ranges > cfg.B:
            if ranges >= n_units:                           # every unit on its own
                ranges = [(u, u) for u in range(n_units)]
            else:                                           # front-end | one range per block | back-end
                ranges = [(0, 0)] + [(1 + 2 * i, 2 + 2 * i) for i in range(cfg.B)] + [(n_units - 1, n_units - 1)]
        if isinstance(ranges, int):                         # `ranges` near-equal groups of blocks
            k = max(1, min(int(ranges), cfg.B))
            cuts = [round(i * cfg.B / k) for i in range(k + 1)]
            ranges = [(1 + 2 * cuts[i], 2 * cuts[i + 1]) for i in range(k)]
            ranges[0] = (0, ranges[0][1])
            ranges[-1] = (ranges[-1][0], n_units - 1)
        flat = [u for lo, hi in ranges for u in range(lo, hi + 1)]
        if flat != list(range(n_units)):
            raise ValueError("ranges must cover units 0..%d in order, got %r" % (n_units - 1, ranges))
        self.ranges = [tuple(r) for r in ranges]
        # Throughput mode pays in SM-time, not latency.  With every SM busy (14 unit ranges, 8 chunks in flight) the
        # intra-frame recurrence with two sequences per CTA wins: 0.81x the SM-time per sequence at 1.6x the latency
        # (212k instead of 192k frames/s at batch 32).
        if intra_algo is None and eng.intra_algo == abi.SB_ALGO_AUTO and batch_size >= 8 and frames_per_call == 1:
            intra_algo = abi.SB_ALGO_WS2
        # grouped mode: G x B rows per direction on the tcgen05 kernel (a fifth of the SM-time of the SIMT recurrence per
        # sequence) once they fill most of a 128-row tile
        if intra_algo is None and eng.intra_algo == abi.SB_ALGO_AUTO and group > 1 and cfg.D == 32 and not cfg.conv_lstm \
                and batch_size * group >= 96:
            intra_algo = abi.SB_ALGO_TC
        self.intra_algo = intra_algo          # None = the engine's choice (SB_ALGO_AUTO unless the caller forced one)
        # Throughput mode pays in SM-time, not latency: the one-step inter-frame call as a tcgen05 GEMM occupies a quarter
        # of the SMs the SIMT tile kernel needs (128-row tiles), which leaves room for the other chunks' recurrences.
        if inter_algo is None and eng.inter_algo == abi.SB_ALGO_AUTO and cfg.D == 32 and batch_size * cfg.n_freqs >= 1024 \
                and (frames_per_call == 1 or group > 1):
            inter_algo = abi.SB_ALGO_TC
        self.inter_algo = inter_algo
        self.n_calls = 0
        self.n_chunks = 0
        self._pending = []
        self._pipe = None
        self._build()

    def _call(self, p: int, slot: int):
        return self.engine.prepare(self.x[slot], self.dis, _shallow(self.states[p]), out=self.y[slot],
                                   new_state=self.states[p ^ 1], film=self.film, workspace=self.ws[slot],
                                   intra_algo=self.intra_algo, inter_algo=self.inter_algo)

    def _build(self):
        """Warm up eagerly (shared-memory opt-ins, lazy module loading), then hand the per-(arena, slot) sb_net_io
        blocks to the native pipe, which captures one CUDA graph per unit range on its own streams."""
        lib = self.engine.lib
        saved = [_clone_state(s) for s in self.states]
        with torch.cuda.device(self.device):
            for p in (0, 1):
                self._call(p, 0).launch()
            torch.cuda.synchronize(self.device)
            period = self.depth if self.depth % 2 == 0 else 2 * self.depth
            self._calls = [self._call(t % 2, t % self.depth) for t in range(period)]     # keeps the tensors alive
            ios = (abi.NetIO * period)(*[c.io for c in self._calls])
            n = len(self.ranges)
            first = (ctypes.c_int * n)(*[lo for lo, _ in self.ranges])
            last = (ctypes.c_int * n)(*[hi for _, hi in self.ranges])
            handle = ctypes.c_void_p()
            rc = lib.sb_pipe_create(self.engine.packed.desc_ref(), ios, period, self.depth, first, last, n,
                                    ctypes.byref(handle))
            abi.check(lib, rc, "sb_pipe_create")
            self._pipe = handle
            torch.cuda.synchronize(self.device)
        for s, z in zip(self.states, saved):
            StreamingSession._copy_state(s, z)

    def close(self):
        if self._pipe is not None:
            self.engine.lib.sb_pipe_destroy(self._pipe)
            self._pipe = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _cur(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------------------------------------------
    @property
    def parity(self) -> int:
        return self.n_calls % 2

    def reset(self):
        """Zero the state on the caller's stream (after an end(), so that nothing fed earlier is still running); call
        begin() afterwards."""
        for a in self.arenas:
            a.flat.zero_()
        self.n_calls = 0
        self.n_chunks = 0
        self._pending.clear()
        lib = self.engine.lib
        abi.check(lib, lib.sb_pipe_reset(self._pipe), "sb_pipe_reset")

    def begin(self):
        """Order every stream of the pipe after what the caller's current stream has enqueued so far (the windows, a
        reset(), a load_state())."""
        lib = self.engine.lib
        abi.check(lib, lib.sb_pipe_begin(self._pipe, self._cur()), "sb_pipe_begin")

    def feed(self, window: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Enqueue one chunk: window [B, M, chunk + lookahead] (pinned host or device) -> `out` [B, S, chunk] (or the
        slot's device buffer, valid until `depth` calls later).  Returns without waiting."""
        if self.group > 1:
            return self._feed_chunk(window, out)
        slot = self.n_calls % self.depth
        if window.shape != self.x[slot].shape or window.dtype != torch.float32 or not window.is_contiguous():
            raise ValueError("window must be a contiguous float32 tensor of shape %s" % (tuple(self.x[slot].shape),))
        optr = None
        if out is not None:
            if out.shape != self.y[slot].shape or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError("out must be a contiguous float32 tensor of shape %s" % (tuple(self.y[slot].shape),))
            optr = out.data_ptr()
        lib = self.engine.lib
        abi.check(lib, lib.sb_pipe_feed(self._pipe, window.data_ptr(), optr), "sb_pipe_feed")
        self.n_calls += 1
        return out if out is not None else self.y[slot]

    def _feed_chunk(self, window: torch.Tensor, out: Optional[torch.Tensor]):
        cfg = self.cfg
        wshape, oshape = (self.B, cfg.num_ch, cfg.n_fft), (self.B, cfg.num_src, cfg.stft_chunk_size)
        if tuple(window.shape) != wshape or window.dtype != torch.float32 or not window.is_contiguous():
            raise ValueError("window must be a contiguous float32 tensor of shape %s" % (wshape,))
        if out is None:
            raise ValueError("grouped mode needs an `out` tensor per chunk (the slot's buffer holds the whole group)")
        if tuple(out.shape) != oshape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 tensor of shape %s" % (oshape,))
        lib = self.engine.lib
        abi.check(lib, lib.sb_pipe_feed_chunk(self._pipe, window.data_ptr(), out.data_ptr()), "sb_pipe_feed_chunk")
        self._pending.append((window, out))                 # keep the tensors alive until their group is launched
        self.n_chunks += 1
        if len(self._pending) == self.group:
            self._pending.clear()
            self.n_calls += 1
        return out

    def flush(self):
        """Grouped mode: run the chunks of a partial group now."""
        lib = self.engine.lib
        abi.check(lib, lib.sb_pipe_flush(self._pipe), "sb_pipe_flush")
        if self._pending:
            self._pending.clear()
            self.n_calls += 1

    def end(self):
        """Make the caller's current stream wait for every chunk fed so far (a partial group is run first)."""
        if self._pending:
            self.flush()
        lib = self.engine.lib
        abi.check(lib, lib.sb_pipe_end(self._pipe, self._cur()), "sb_pipe_end")

    @property
    def state(self) -> dict:
        """The current state in the reference's schema (valid after end() + a synchronisation of the caller's stream)."""
        return self.states[self.parity]

    def load_state(self, state: dict):
        """Adopt a reference-layout state dict as the state the next chunk starts from (on the caller's stream)."""
        StreamingSession._copy_state(self.states[self.parity], state)

    def set_dis_embed(self, dis_embed: torch.Tensor):
        """New bubble radii for the rows of this session (on the caller's stream, before begin()): the FiLM table the
        captured graphs read is recomputed in place."""
        if self.dis is None:
            return
        self.dis.copy_(dis_embed.to(self.device, torch.float32))
        self.film.copy_(self.engine.film_table(self.dis))

    def launches_per_step(self) -> int:
        before = _lib.launch_count()
        saved = [_clone_state(s) for s in self.states]
        self._call(0, 0).launch()
        torch.cuda.synchronize(self.device)
        for s, z in zip(self.states, saved):
            StreamingSession._copy_state(s, z)
        return _lib.launch_count() - before
