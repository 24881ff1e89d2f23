"""The training losses of the reference's shipped experiments, importable where `asteroid` / `auraloss` are not installed.

The reference's loss classes are thin wrappers over two third-party packages that are absent from this image:
  * src/losses/SNRLosses.py:3,6-30  -> ``asteroid.losses.sdr.SingleSrcNegSDR``      (pretrain stages, via SNRLP.py:9-42)
  * src/losses/MultiResoLoss.py:3,6-31 -> ``auraloss.freq.MultiResolutionSTFTLoss`` (finetune stages, + l1_ratio * L1)
This module restates the published algorithms of those two classes in plain torch (autograd does the backward; the losses
see only the separated waveform, they are not part of the kernel path) and mirrors the reference's own wrappers
(``SNRLosses``, ``SNRLPLoss``, ``MultiResoFuseLoss``: same constructor arguments, same ``forward(est, gt)`` contract,
same per-item / scalar return shapes), so that ``train_dist.TrainModule`` can run every shipped experiment JSON.
Where the third-party packages ARE installed the reference's own classes work unchanged on this package's ``Net``.

PINNING: `SingleSrcNegSDR` follows asteroid's published definition (zero-mean, EPS = 1e-8, 10 log10) and is checked in
tests against the closed forms; `MultiResolutionSTFTLoss` (+ the A-weighting FIR prefilter the finetune JSONs switch on)
is restated from the published auraloss 0.4 source from memory - neither package is available here to compare against,
so these two are "unpinned" in the sense of DESIGN.md section 3.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class SingleSrcNegSDR(nn.Module):
    """asteroid.losses.sdr.SingleSrcNegSDR: negative SI-SDR / SD-SDR / SNR of [batch, time] signals, one value per item."""

    def __init__(self, sdr_type: str, zero_mean: bool = True, take_log: bool = True, reduction: str = "none", EPS: float = 1e-8):
        super().__init__()
        assert sdr_type in ("snr", "sisdr", "sdsdr") and reduction in ("none", "mean")
        self.sdr_type, self.zero_mean, self.take_log, self.reduction, self.EPS = sdr_type, zero_mean, take_log, reduction, EPS

    def forward(self, est_target: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.size() != est_target.size() or target.ndim != 2:
            raise TypeError("Inputs must be of shape [batch, time], got %s and %s" % (tuple(target.size()), tuple(est_target.size())))
        if self.zero_mean:
            target = target - target.mean(dim=1, keepdim=True)
            est_target = est_target - est_target.mean(dim=1, keepdim=True)
        if self.sdr_type in ("sisdr", "sdsdr"):
            dot = (est_target * target).sum(dim=1, keepdim=True)
            energy = (target ** 2).sum(dim=1, keepdim=True) + self.EPS
            scaled = dot * target / energy
        else:
            scaled = target
        noise = est_target - (target if self.sdr_type in ("sdsdr", "snr") else scaled)
        val = (scaled ** 2).sum(dim=1) / ((noise ** 2).sum(dim=1) + self.EPS)
        if self.take_log:
            val = 10 * torch.log10(val + self.EPS)
        val = val.mean() if self.reduction == "mean" else val
        return -val


class SNRLosses(nn.Module):
    """src/losses/SNRLosses.py:6-60 (every `name` it accepts)."""

    def __init__(self, name: str, **kwargs):
        super().__init__()
        self.name = name
        if name in ("sisdr", "snr"):
            self.loss_fn = SingleSrcNegSDR(name)
        elif name in ("fused", "max_fused"):
            self.loss1, self.loss2 = SingleSrcNegSDR("sisdr"), SingleSrcNegSDR("snr")
        elif name == "sdsdr":
            self.loss1, self.loss2 = SingleSrcNegSDR("snr"), SingleSrcNegSDR("sdsdr")
        elif name == "full":
            self.loss1, self.loss2, self.loss3 = SingleSrcNegSDR("snr"), SingleSrcNegSDR("sdsdr"), SingleSrcNegSDR("sisdr")
        else:
            raise AssertionError("Invalid loss function used: Loss %s not found" % name)

    def forward(self, est: torch.Tensor, gt: torch.Tensor, **kwargs) -> torch.Tensor:
        B, C, T = est.shape
        est, gt = est.reshape(B * C, T), gt.reshape(B * C, T)
        if self.name == "fused":
            return 0.5 * self.loss1(est, gt) + 0.5 * self.loss2(est, gt)
        if self.name in ("max_fused", "sdsdr"):
            return torch.maximum(self.loss1(est, gt), self.loss2(est, gt))
        if self.name == "full":
            return 0.5 * self.loss3(est, gt) + 0.5 * torch.maximum(self.loss1(est, gt), self.loss2(est, gt))
        return self.loss_fn(est, gt)


class SNRLPLoss(nn.Module):
    """src/losses/SNRLP.py:9-42: negative SNR on clips with a target, `neg_weight` x L1 on clips whose target is silence
    (nobody inside the bubble).  Returns one value per item; PLModule takes the mean (hl_module.py:321)."""

    def __init__(self, snr_loss_name: str = "snr", neg_weight: float = 1):
        super().__init__()
        self.snr_loss = SNRLosses(snr_loss_name)
        self.lp_loss = nn.L1Loss()
        self.neg_weight = neg_weight

    def forward(self, est: torch.Tensor, gt: torch.Tensor, **kwargs) -> torch.Tensor:
        comp = torch.zeros(est.shape[0], device=est.device, dtype=est.dtype)
        mask = gt.abs().amax(dim=(1, 2)) == 0
        if bool(mask.any()):
            comp[mask] = self.lp_loss(est[mask], gt[mask]) * self.neg_weight
        if bool((~mask).any()):
            comp[~mask] = self.snr_loss(est[~mask], gt[~mask])
        return comp


# ---------------------------------------------------------------------------------------------------------------------
# auraloss.freq (0.4) restated: STFTLoss / MultiResolutionSTFTLoss with the options the reference's JSONs use
# ---------------------------------------------------------------------------------------------------------------------
def a_weighting_fir(fs: float, ntaps: int = 101) -> torch.Tensor:
    """auraloss.perceptual.FIRFilter(filter_type="aw"): analog A-weighting (IEC/CD 1672) -> bilinear -> 512-point response
    -> least-squares FIR with `ntaps` taps."""
    import scipy.signal
    f1, f2, f3, f4, a1000 = 20.598997, 107.65265, 737.86223, 12194.217, 1.9997
    nums = [(2 * np.pi * f4) ** 2 * (10 ** (a1000 / 20)), 0, 0, 0, 0]
    dens = np.polymul([1, 4 * np.pi * f4, (2 * np.pi * f4) ** 2], [1, 4 * np.pi * f1, (2 * np.pi * f1) ** 2])
    dens = np.polymul(np.polymul(dens, [1, 2 * np.pi * f3]), [1, 2 * np.pi * f2])
    b, a = scipy.signal.bilinear(nums, dens, fs=fs)
    w, h = scipy.signal.freqz(b, a, worN=512, fs=fs)
    taps = scipy.signal.firls(ntaps, w, np.abs(h), fs=fs)
    return torch.tensor(taps.astype("float32")).view(1, 1, -1)


class STFTLoss(nn.Module):
    def __init__(self, fft_size=1024, hop_size=256, win_length=1024, window="hann_window", w_sc=1.0, w_log_mag=1.0,
                 w_lin_mag=0.0, w_phs=0.0, sample_rate=None, perceptual_weighting=False, eps=1e-8, reduction="mean"):
        super().__init__()
        if w_phs:
            raise NotImplementedError("phase term of auraloss.freq.STFTLoss (unused by the reference's experiments)")
        self.fft_size, self.hop_size, self.win_length = fft_size, hop_size, win_length
        self.register_buffer("window", getattr(torch, window)(win_length), persistent=False)
        self.w_sc, self.w_log_mag, self.w_lin_mag, self.eps, self.reduction = w_sc, w_log_mag, w_lin_mag, eps, reduction
        self.perceptual_weighting = perceptual_weighting
        if perceptual_weighting:
            if sample_rate is None:
                raise ValueError("`sample_rate` must be supplied when `perceptual_weighting = True`.")
            self.register_buffer("fir", a_weighting_fir(sample_rate), persistent=False)

    def _mag(self, x: torch.Tensor) -> torch.Tensor:
        s = torch.stft(x, self.fft_size, self.hop_size, self.win_length, self.window.to(x.device, x.dtype), return_complex=True)
        return torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=self.eps))

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        bs, chs, n = x.shape
        if self.perceptual_weighting:
            k = self.fir.to(x.device, x.dtype)
            x = F.conv1d(x.reshape(bs * chs, 1, n), k, padding=k.shape[-1] // 2).view(bs, chs, -1)
            y = F.conv1d(y.reshape(bs * chs, 1, n), k, padding=k.shape[-1] // 2).view(bs, chs, -1)
        xm, ym = self._mag(x.reshape(-1, x.shape[-1])), self._mag(y.reshape(-1, y.shape[-1]))
        red = torch.mean if self.reduction == "mean" else torch.sum
        loss = x.new_zeros(())
        if self.w_sc:
            loss = loss + self.w_sc * torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")
        if self.w_log_mag:
            loss = loss + self.w_log_mag * red((torch.log(xm) - torch.log(ym)).abs())
        if self.w_lin_mag:
            loss = loss + self.w_lin_mag * red((xm - ym).abs())
        return loss


class MultiResolutionSTFTLoss(nn.Module):
    def __init__(self, fft_sizes: Optional[List[int]] = None, hop_sizes: Optional[List[int]] = None,
                 win_lengths: Optional[List[int]] = None, window="hann_window", w_sc=1.0, w_log_mag=1.0, w_lin_mag=0.0,
                 w_phs=0.0, sample_rate=None, perceptual_weighting=False, **kwargs):
        super().__init__()
        fft_sizes = fft_sizes or [1024, 2048, 512]
        hop_sizes = hop_sizes or [120, 240, 50]
        win_lengths = win_lengths or [600, 1200, 240]
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = nn.ModuleList(
            STFTLoss(fs, ss, wl, window, w_sc, w_log_mag, w_lin_mag, w_phs, sample_rate, perceptual_weighting, **kwargs)
            for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths))

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return sum(f(x, y) for f in self.stft_losses) / len(self.stft_losses)


class MultiResoFuseLoss(nn.Module):
    """src/losses/MultiResoLoss.py:6-31: multi-resolution STFT loss + l1_ratio x L1 (scalar)."""

    def __init__(self, l1_ratio=0, **kwargs):
        super().__init__()
        self.l1_ratio = l1_ratio
        self.l1 = nn.L1Loss()
        self.loss_fn = MultiResolutionSTFTLoss(**kwargs)

    def forward(self, est: torch.Tensor, gt: torch.Tensor, **kwargs) -> torch.Tensor:
        if self.l1_ratio > 0:
            return self.loss_fn(est, gt) + self.l1_ratio * self.l1(est, gt)
        return self.loss_fn(est, gt)


# the reference's dotted paths (experiment JSONs) -> the classes above
REFERENCE_LOSSES = {
    "src.losses.SNRLP.SNRLPLoss": SNRLPLoss,
    "src.losses.SNRLosses.SNRLosses": SNRLosses,
    "src.losses.MultiResoLoss.MultiResoFuseLoss": MultiResoFuseLoss,
}
