"""Parameter tree of the separator with the reference's exact ``state_dict`` key names and shapes, so
``Net.load_state_dict(torch.load('best.pt')['model'])`` is strict-compatible with reference checkpoints.

The torch.nn modules below are used as *parameter containers only* (registration, reference-identical initialisation
in the reference's construction order, serialisation).  Their ``forward`` is never called: the arithmetic runs in the
sm_100a kernels of libsoundbubble_sm100a.so.  Reference: DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/
tfgridnet_causal.py (:271-401 TFGridNet.__init__, :566-684 GridNetBlock.__init__), OPT = the _optim twin.
"""
import torch
import torch.nn as nn

from .filterbank import stft_filters, stft_window
from .packing import ModelConfig


class _Bag(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the forward pass runs in libsoundbubble_sm100a.so")


def _wrapped_ln(n: int) -> nn.Module:          # LayerNormalization4D / 4DCF add a `.norm.` level (DE3:906-932)
    m = _Bag()
    m.norm = nn.LayerNorm(n)
    return m


class _STFTFBParams(_Bag):
    """Buffers of asteroid_filterbanks.STFTFB.  Published releases register ``_filters`` and (0.3.x onwards) the analysis
    window as ``torch_window``; the package is absent here, so a checkpoint may or may not carry ``torch_window``.  Both
    load strictly: a missing window is filled with the closed form (nothing reads it - the kernels take ``_filters``)."""

    def __init__(self, n_fft: int, hop: int):
        super().__init__()
        self.register_buffer("_filters", stft_filters(n_fft, hop))
        self.register_buffer("torch_window", stft_window(n_fft))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + "torch_window"
        if key not in state_dict:
            state_dict[key] = self.torch_window.detach().clone()
        elif state_dict[key].dtype != self.torch_window.dtype:          # some releases stored the numpy float64 window
            state_dict[key] = state_dict[key].to(self.torch_window.dtype)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


def _filterbank(n_fft: int, hop: int) -> nn.Module:
    m = _Bag()
    m.filterbank = _STFTFBParams(n_fft, hop)
    return m


def _seq(*mods) -> nn.Sequential:
    return nn.Sequential(*mods)


class BlockParams(_Bag):
    def __init__(self, cfg: ModelConfig):
        super().__init__()
        D, H, Fq = cfg.D, cfg.H, cfg.n_freqs
        if cfg.conv_lstm:
            self.conv = nn.Conv1d(D, D, cfg.lstm_down, stride=cfg.lstm_down)
            self.act = nn.PReLU()
            self.norm = _wrapped_ln(D)
            self.intra_rnn = nn.LSTM(D, H, 1, batch_first=True, bidirectional=True)
            self.deconv = nn.ConvTranspose1d(2 * H, D, cfg.lstm_down, stride=cfg.lstm_down)
        else:
            self.intra_norm = _wrapped_ln(D)
            self.intra_rnn = nn.LSTM(D, H, 1, batch_first=True, bidirectional=True)
            self.intra_linear = nn.Linear(2 * H, D)
        self.inter_norm = _wrapped_ln(D)
        self.inter_rnn = nn.LSTM(D, H, 1, batch_first=True, bidirectional=False)
        self.inter_linear = nn.Linear(H, D)
        if cfg.use_attn:
            E, L = cfg.attn_E, cfg.L
            self.attn_conv_Q = _seq(nn.Linear(D, E * L), nn.PReLU(), nn.Identity(), _wrapped_ln(Fq * E))
            self.attn_conv_K = _seq(nn.Linear(D, E * L), nn.PReLU(), nn.Identity(), _wrapped_ln(Fq * E))
            self.attn_conv_V = _seq(nn.Linear(D, (D // L) * L), nn.PReLU(), nn.Identity(), _wrapped_ln(Fq * (D // L)))
            self.attn_concat_proj = _seq(nn.Linear(D, D), nn.PReLU(), nn.Identity(), _wrapped_ln(Fq * D))


class FilmParams(_Bag):
    def __init__(self, d_in: int, D: int):
        super().__init__()
        self.weight = nn.Conv1d(d_in, D, 1)
        self.bias = nn.Conv1d(d_in, D, 1)


class TFGridNetParams(_Bag):
    def __init__(self, cfg: ModelConfig):
        super().__init__()
        D, Fq = cfg.D, cfg.n_freqs
        self.enc = _filterbank(cfg.n_fft, cfg.stft_chunk_size)
        self.dec = _filterbank(cfg.n_fft, cfg.stft_chunk_size)
        mods = [nn.Conv2d(cfg.conv_in_ch, D, (3, 3), padding=(0, 1))]
        if cfg.use_first_ln:
            mods.append(nn.LayerNorm(D))
        self.conv = _seq(*mods)
        if cfg.variant == "dis_embed":
            en = _Bag()
            if cfg.dis_type.startswith("conv"):
                en.dis_embedding = _seq(nn.Linear(3, Fq * cfg.film_in, bias=False))
                en.dis_norm = nn.LayerNorm(cfg.film_in)
            else:
                n = Fq if cfg.dis_type == "linear1" else Fq * D
                en.dis_embedding = _seq(nn.Linear(3, n, bias=False), nn.LayerNorm(n))
            self.embed_net = en
        self.blocks = nn.ModuleList([])
        if cfg.variant == "dis_embed":
            self.embeds = nn.ModuleList([])
        for i in range(cfg.B):
            self.blocks.append(BlockParams(cfg))
            if i > 0 and cfg.variant == "dis_embed":
                self.embeds.append(FilmParams(cfg.film_in, D))
        self.deconv = nn.ConvTranspose2d(D, 2 * cfg.num_src, (3, 3), padding=(2, 1))
