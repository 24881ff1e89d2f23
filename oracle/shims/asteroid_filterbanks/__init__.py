"""Stand-in for the third-party ``asteroid_filterbanks`` package (NOT part of the reference repo).

TEST INFRASTRUCTURE ONLY.  The reference imports ``make_enc_dec`` from ``asteroid_filterbanks``
(/root/reference/src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py:13, used at :326-330).
That dependency is unpinned in /root/reference/requirements2.txt:15 (``asteroid``) and is absent from this
image, so this file restates the published behaviour of its ``STFTFB`` / ``Encoder`` / ``Decoder`` for the one
call the hot path makes: ``make_enc_dec('stft', n_filters=N, kernel_size=N, stride=S, window_type=...)``.

Published algorithm restated here (asteroid-filterbanks ``stft_fb.STFTFB``):
  * ``window_type`` is not an ``STFTFB`` argument, it is swallowed by ``**kwargs`` -> window =
    ``sqrt(hanning(N+1)[:-1])`` (periodic Hann, square-rooted);
  * basis = ``fft(eye(N)) / (0.5*sqrt(kernel*N/stride))``; rows ``[Re 0..N/2 ; Im 0..N/2]``; the DC and Nyquist
    real rows are additionally divided by ``sqrt(2)``; multiplied by the window; stored as the float32 buffer
    ``_filters`` of shape ``[N+2, 1, N]``;
  * ``Encoder``: >=3-D input is viewed ``[-1, 1, time]`` -> ``conv1d(stride)`` -> viewed back ``[B, M, N+2, T]``;
  * ``Decoder``: >=4-D input is viewed ``[-1, N+2, T]`` -> ``conv_transpose1d(stride)`` -> ``[B, S, time]``.
The encoder and the decoder each own a copy of the basis (``enc.filterbank._filters`` and
``dec.filterbank._filters`` in a checkpoint).

This shim is only ever imported by oracle/make_golden.py (in the build container, to run the UNMODIFIED
reference model) and by tests that check the oracle's STFT basis against it.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class STFTFB(nn.Module):
    def __init__(self, n_filters, kernel_size, stride=None, window=None, sample_rate=8000.0, **kwargs):
        super().__init__()
        assert n_filters >= kernel_size
        self.n_filters = n_filters
        self.kernel_size = kernel_size
        self.stride = stride if stride else kernel_size // 2
        self.sample_rate = sample_rate
        self.cutoff = int(n_filters / 2 + 1)
        self.n_feats_out = 2 * self.cutoff
        if window is None:
            self.window = np.hanning(kernel_size + 1)[:-1] ** 0.5
        else:
            assert len(window) == kernel_size
            self.window = np.asarray(window)
        basis = np.fft.fft(np.eye(n_filters))
        basis /= 0.5 * np.sqrt(kernel_size * n_filters / self.stride)
        lpad = int((n_filters - kernel_size) // 2)
        rpad = int(n_filters - kernel_size - lpad)
        idx = list(range(lpad, n_filters - rpad))
        basis = np.vstack([np.real(basis[: self.cutoff, idx]), np.imag(basis[: self.cutoff, idx])])
        basis[0, :] /= np.sqrt(2)
        basis[n_filters // 2, :] /= np.sqrt(2)
        self.register_buffer("_filters", torch.from_numpy(basis * self.window).unsqueeze(1).float())

    def filters(self):
        return self._filters


class _Coder(nn.Module):
    def __init__(self, filterbank):
        super().__init__()
        self.filterbank = filterbank
        self.stride = filterbank.stride


class Encoder(_Coder):
    def forward(self, waveform):
        filters = self.filterbank.filters()
        if waveform.ndim == 1:
            return F.conv1d(waveform[None, None], filters, stride=self.stride).squeeze()
        if waveform.ndim == 2:
            return F.conv1d(waveform.unsqueeze(1), filters, stride=self.stride)
        if waveform.ndim == 3:
            batch, channels, time_len = waveform.shape
            if channels == 1:
                return F.conv1d(waveform, filters, stride=self.stride)
            out = F.conv1d(waveform.reshape(-1, 1, time_len), filters, stride=self.stride)
            return out.view(batch, channels, out.shape[-2], out.shape[-1])
        out = F.conv1d(waveform.reshape(-1, 1, waveform.shape[-1]), filters, stride=self.stride)
        return out.view(waveform.shape[:-1] + out.shape[-2:])


class Decoder(_Coder):
    def forward(self, spec):
        filters = self.filterbank.filters()
        if spec.ndim == 2:
            return F.conv_transpose1d(spec.unsqueeze(0), filters, stride=self.stride).squeeze()
        if spec.ndim == 3:
            out = F.conv_transpose1d(spec, filters, stride=self.stride)
            return out.squeeze(1) if out.shape[1] == 1 else out
        out = F.conv_transpose1d(spec.reshape((-1,) + spec.shape[-2:]), filters, stride=self.stride)
        return out.view(spec.shape[:-2] + (-1,))


def make_enc_dec(fb_name, n_filters, kernel_size, stride=None, sample_rate=8000.0, who_is_pinv=None,
                 padding=0, output_padding=0, **kwargs):
    assert fb_name == "stft", "only the STFT filterbank is on the Sound Bubble hot path"
    assert who_is_pinv is None and padding == 0 and output_padding == 0
    enc = Encoder(STFTFB(n_filters, kernel_size, stride=stride, sample_rate=sample_rate, **kwargs))
    dec = Decoder(STFTFB(n_filters, kernel_size, stride=stride, sample_rate=sample_rate, **kwargs))
    return enc, dec
