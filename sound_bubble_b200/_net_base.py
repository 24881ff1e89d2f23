"""Shared implementation of the two drop-in ``Net`` classes (reference: src/models/*/net.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _abi as abi
from . import _lib
from ._modules import TFGridNetParams
from .engine import Engine, init_state
from .packing import ModelConfig, PackedWeights


def mod_pad(x, chunk_size, pad):
    """DE3/net.py:8-18: right-pad to a whole number of chunks, then the look-back / look-ahead pad."""
    mod = 0
    if (x.shape[-1] % chunk_size) != 0:
        mod = chunk_size - (x.shape[-1] % chunk_size)
    x = F.pad(x, (0, mod))
    x = F.pad(x, pad)
    return x, mod


class NetBase(nn.Module):
    variant = "dis_embed"

    def _setup(self, cfg: ModelConfig):
        self.cfg = cfg
        self.stft_chunk_size = cfg.stft_chunk_size
        self.stft_pad_size = cfg.stft_pad_size
        self.num_ch = cfg.num_ch
        self.lookahead = cfg.lookahead
        self.stft_back_pad = cfg.stft_back_pad
        self.embed_dim = cfg.D
        self.E = cfg.E
        self.nfft = cfg.n_fft
        self.tfgridnet = TFGridNetParams(cfg)
        # state_dict() keys, recorded once: nn.DataParallel replicas (the reference harness wraps the model when
        # use_dp=true, hl_module.py:34-35) keep their weights as plain attributes, not in _parameters, so state_dict() on a
        # replica is nearly empty - _named_weights() walks these names instead
        self._weight_names = list(self.state_dict().keys())
        self._engine = None
        self._engine_key = None
        # Whole-utterance calls on large batches run as a pipeline of time slices (streaming.PipelinedSession with
        # `offline_slice_frames` frames per call): the serial inter-frame recurrence of one slice overlaps the
        # intra-frame work of the next.  Same kernels per frame, so results agree with the single call to fp32 rounding.
        self.pipeline_offline = True
        self.offline_slice_frames = 125
        self.offline_min_rows = 12000             # batch * frames below which the single call is used (measured crossover:
                                                  # batch 16 x 625 frames, profiles/r01_offline_slices.txt)
        self._offline_pipes = {}
        self.offline_intra_algo = None            # None = SB_ALGO_AUTO per slice
        # the inter-frame path of a slice on the tcgen05 kernel: 37 CTAs instead of one per SM, so that the next slice's
        # intra-frame work finds free SMs (27.5 ms instead of 45.5 ms per batch of 32 x 5 s; the single call takes 48.9 ms)
        self.offline_inter_algo = abi.SB_ALGO_TC if cfg.D == 32 else None

    # -- engine (packed weights on the device of the parameters) ---------------------------------------------
    def _named_weights(self) -> dict:
        """name -> tensor for every state_dict() key, by attribute walk (works on nn.DataParallel replicas too, whose
        tensors are autograd-connected broadcast copies held as plain attributes).  The walk down to the module that holds
        each tensor is done once per instance: a replica's __dict__ is a shallow copy of the original's, so the cache
        records whose it is and a replica rebuilds it against its own submodules."""
        holders = self.__dict__.get("_weight_holders")
        if holders is None or holders[0] is not self:
            pairs = []
            for name in self._weight_names:
                obj, parts = self, name.split(".")
                for part in parts[:-1]:
                    obj = getattr(obj, part)
                pairs.append((name, obj, parts[-1]))
            holders = (self, pairs)
            self.__dict__["_weight_holders"] = holders
        out = {}
        for name, mod, leaf in holders[1]:
            t = mod._parameters.get(leaf)
            if t is None:
                t = mod._buffers.get(leaf)
            if t is None:
                t = mod.__dict__[leaf]                   # replica: plain attribute
            out[name] = t
        return out

    def _weights_key(self, named):
        return tuple((t.data_ptr(), t._version) for t in named.values())

    def engine(self) -> Engine:
        named = self._named_weights()
        key = self._weights_key(named)
        if self._engine is None or key != self._engine_key:
            sd = {k: v.detach() for k, v in named.items()}
            dev = next(iter(sd.values())).device
            _lib.require_cuda(next(iter(sd.values())))
            self._engine = Engine(_lib.load(), self.cfg, PackedWeights(sd, self.cfg, dev))
            self._engine_key = key
        return self._engine

    # -- reference API -----------------------------------------------------------------------------------------
    def init_buffers(self, batch_size, device):
        return init_state(self.cfg, batch_size, device)

    def _predict(self, x, dis_embed, input_state, pad=True):
        _lib.require_cuda(x)
        mod = 0
        if pad:
            pad_size = (self.stft_back_pad, self.stft_pad_size) if self.lookahead else (0, 0)
            x, mod = mod_pad(x, chunk_size=self.stft_chunk_size, pad=pad_size)
        with torch.no_grad():
            y = self._forward_sliced(x, dis_embed, input_state)
            if y is None:
                y, _ = self.engine().forward(x, dis_embed, input_state)
            next_state = input_state
        if mod != 0:
            y = y[:, :, :-mod]
        return y, next_state

    # -- training (autograd) ------------------------------------------------------------------------------------
    def _wants_grad(self, input_state) -> bool:
        """The differentiable path is taken exactly where the reference trains: module in train() mode, autograd on, no
        carried state (PLModule._step calls self.model(inputs), hl_module.py:309), for the configurations that have
        backward kernels.  Everything else runs the inference kernels under no_grad, so a backward() on their output
        fails loudly instead of training nothing."""
        if not (self.training and torch.is_grad_enabled() and input_state is None):
            return False
        from .training import check_trainable
        try:
            check_trainable(self.cfg)
        except NotImplementedError:
            return False
        return any(t.requires_grad for t in self._named_weights().values())

    def _train_forward(self, x, dis_embed, pad=True):
        """Net.forward for a training step: same padding / cropping as _predict, gradients to every parameter through
        SeparatorFunction (training.py).  'next_state' has the reference's schema; its tensors are detached (the reference's
        carry the graph, which nothing in its training loop uses)."""
        from .training import differentiable_forward_with_state
        _lib.require_cuda(x)
        mod = 0
        if pad:
            pad_size = (self.stft_back_pad, self.stft_pad_size) if self.lookahead else (0, 0)
            x, mod = mod_pad(x, chunk_size=self.stft_chunk_size, pad=pad_size)
        named = self._named_weights()
        y, next_state = differentiable_forward_with_state(_lib.load(), self.cfg, named, x, dis_embed)
        if mod != 0:
            y = y[:, :, :-mod]
        return {'output': y, 'next_state': next_state}

    def _forward_sliced(self, x, dis_embed, state):
        """The whole-utterance call as K time slices through the native pipe (None = not applicable, use the single
        call).  Updates `state` in place like Engine.forward does (the reference mutates the dict it was given)."""
        cfg = self.cfg
        eng = self.engine()
        if not self.pipeline_offline or x.dim() != 3 or x.dtype != torch.float32 or getattr(self, "_is_replica", False):
            return None                                     # (DataParallel replicas live for one call: no cached pipes)
        B, M, N = x.shape
        hop, look = cfg.stft_chunk_size, cfg.n_fft - cfg.stft_chunk_size
        T = eng.n_frames(N)
        Tc = int(self.offline_slice_frames)
        if M != cfg.num_ch or T < 2 * Tc or B * T < self.offline_min_rows or (cfg.variant == "dis_embed" and dis_embed is None):
            return None
        K = max(2, round(T / Tc))
        Tc = T // K
        rem = T - K * Tc
        from .state_io import StateArena, flatten_state
        from .streaming import PipelinedSession
        try:                                                # anything unusual about the state: let the single call report it
            names, _ = flatten_state(state)
        except (TypeError, AttributeError):
            return None
        if names != flatten_state(self.init_buffers(1, "meta"))[0]:
            return None
        key = (B, Tc, str(x.device), id(eng), self.offline_intra_algo, self.offline_inter_algo)
        pipe = self._offline_pipes.get(key)
        if pipe is None:
            if len(self._offline_pipes) > 2:
                for p in self._offline_pipes.values():
                    p.close()
                self._offline_pipes.clear()
            dis0 = torch.zeros(B, 3, device=x.device) if cfg.variant == "dis_embed" else None
            pipe = PipelinedSession(self, B, dis0, depth=min(K, 6), frames_per_call=Tc, intra_algo=self.offline_intra_algo,
                                    inter_algo=self.offline_inter_algo)
            self._offline_pipes[key] = pipe
        n_in = hop * Tc + look
        windows = x.unfold(-1, n_in, hop * Tc)[:, :, :K].permute(2, 0, 1, 3).contiguous()       # [K, B, M, n_in]
        out = torch.empty(K, B, cfg.num_src, hop * Tc, dtype=torch.float32, device=x.device)
        pipe.reset()
        if cfg.variant == "dis_embed":
            pipe.set_dis_embed(dis_embed)
        pipe.load_state(state)
        pipe.begin()
        for c in range(K):
            pipe.feed(windows[c], out[c])
        pipe.end()
        y = out.permute(1, 2, 0, 3).reshape(B, cfg.num_src, K * hop * Tc)
        final = StateArena(pipe.state)                      # a private copy: the pipe's arenas are reused by the next call
        for k, v in final.state.items():
            state[k] = v
        if rem > 0:                                         # the last few frames: one ordinary call on the carried state
            tail = x[..., K * hop * Tc:].contiguous()
            y_tail, _ = eng.forward(tail, dis_embed, state)
            y = torch.cat([y, y_tail], dim=-1)
        return y

    def streaming(self, batch_size: int, dis_embed=None, use_graph: bool = True, pipelined: bool = False, ranges=None, depth: int = 8,
                  intra_algo=None, inter_algo=None, frames_per_call: int = 1, group: int = 1):
        """A chunk-by-chunk session with device-resident state and captured CUDA graphs (see streaming.py).
        pipelined=True: asynchronous feed() with consecutive chunks overlapping on two streams (throughput mode)."""
        from .streaming import PipelinedSession, StreamingSession
        if pipelined:
            return PipelinedSession(self, batch_size, dis_embed, ranges=ranges, depth=depth, intra_algo=intra_algo,
                                    inter_algo=inter_algo, frames_per_call=frames_per_call, group=group)
        return StreamingSession(self, batch_size, dis_embed, use_graph=use_graph, frames_per_call=frames_per_call)
