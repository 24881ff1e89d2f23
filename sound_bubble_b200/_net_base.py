"""Shared implementation of the two drop-in ``Net`` classes (reference: src/models/*/net.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

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
        self._engine = None
        self._engine_key = None

    # -- engine (packed weights on the device of the parameters) ---------------------------------------------
    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def engine(self) -> Engine:
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            sd = {k: v.detach() for k, v in self.state_dict(keep_vars=True).items()}
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
            y, next_state = self.engine().forward(x, dis_embed, input_state)
        if mod != 0:
            y = y[:, :, :-mod]
        return y, next_state

    def streaming(self, batch_size: int, dis_embed=None, use_graph: bool = True, pipelined: bool = False, ranges=None, depth: int = 8,
                  intra_algo=None, inter_algo=None):
        """A chunk-by-chunk session with device-resident state and captured CUDA graphs (see streaming.py).
        pipelined=True: asynchronous feed() with consecutive chunks overlapping on two streams (throughput mode)."""
        from .streaming import PipelinedSession, StreamingSession
        if pipelined:
            return PipelinedSession(self, batch_size, dis_embed, ranges=ranges, depth=depth, intra_algo=intra_algo,
                                    inter_algo=inter_algo)
        return StreamingSession(self, batch_size, dis_embed, use_graph=use_graph)
