"""The STFT analysis/synthesis basis the reference gets from the third-party ``asteroid_filterbanks`` package
(``make_enc_dec('stft', n_filters=n_fft, kernel_size=n_fft, stride=hop, window_type='hann')`` at
src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py:326-330; asteroid is unpinned in requirements2.txt:15).

Restated from the published STFTFB construction: the window argument is swallowed by **kwargs, so the window is
``sqrt(hanning(n_fft + 1)[:-1])``; the basis is the DFT of the identity scaled by ``0.5 * sqrt(n_fft * n_fft / hop)``,
real rows 0..n_fft/2 stacked over imaginary rows, DC and Nyquist real rows divided by sqrt(2), times the window.  It is
stored as the non-trainable buffer ``filterbank._filters`` of shape [n_fft + 2, 1, n_fft] in both encoder and decoder;
a checkpoint's own buffer always wins (the kernels read the buffer, not a closed form).
"""
import numpy as np
import torch


def stft_window(n_fft: int) -> torch.Tensor:
    """STFTFB's default analysis window, ``sqrt(hanning(n_fft + 1)[:-1])``, as its ``torch_window`` buffer holds it."""
    return torch.from_numpy(np.hanning(n_fft + 1)[:-1] ** 0.5).float()


def stft_filters(n_fft: int, hop: int) -> torch.Tensor:
    window = np.hanning(n_fft + 1)[:-1] ** 0.5
    basis = np.fft.fft(np.eye(n_fft)) / (0.5 * np.sqrt(n_fft * n_fft / hop))
    cut = n_fft // 2 + 1
    basis = np.vstack([np.real(basis[:cut]), np.imag(basis[:cut])])
    basis[0] /= np.sqrt(2.0)
    basis[n_fft // 2] /= np.sqrt(2.0)
    return torch.from_numpy(basis * window).unsqueeze(1).float()
