"""Headline-configuration parity check.  TEST INFRASTRUCTURE ONLY (used by tests/, smoke and bench.py's checker leg).

BASELINE.json's north_star bar: "outputs match the reference PyTorch model on the same inputs to within 1e-3 RMS on
the separated waveform (SI-SDR within 0.05 dB)".  `compare_with_oracle` runs the CPU oracle (whole-clip call of
oracle.tfgridnet_oracle.net_forward = DE3/net.py:70-93; the reference's own streaming protocol edge/causal_infer.py:28-47
yields the same samples, the model is causal) on a few utterances of a batch and reports every figure of that bar for the
separated waveforms the GPU path produced for the SAME utterances:

  rms / rel_rms / maxabs           error of the waveform against the oracle's (absolute, relative to the oracle's RMS)
  si_sdr_vs_oracle_db              SI-SDR(ours, oracle), the worst utterance
  si_sdr_target_{ours,oracle}_db   SI-SDR against a synthetic target (the clean source at the reference microphone)
  si_sdr_delta_db                  max |difference| of the two over the utterances  (bar: <= 0.05 dB)
  worst_second_rms                 largest RMS error over 1 s windows (error growth along the 625 recurrent steps)
"""
from __future__ import annotations

import time
from typing import Optional, Sequence

import torch

from . import tfgridnet_oracle as orc

RMS_BAR = 1e-3              # north_star
SI_SDR_DELTA_BAR = 0.05     # dB, north_star


def compare_with_oracle(sd, kwargs: dict, mixture: torch.Tensor, dis_embed: torch.Tensor, ours: torch.Tensor,
                        rows: Sequence[int], target: Optional[torch.Tensor] = None, variant: str = "dis_embed",
                        sample_rate: int = 24000) -> dict:
    """sd: CPU state dict (reference layout); mixture [B, M, N], dis_embed [B, 3], ours [B, S, N] (any device);
    rows: the utterances to check; target [B, N] optional clean source for the SI-SDR delta."""
    ocfg = orc.OracleConfig.from_kwargs(variant, **kwargs)
    rows = list(rows)
    sd = {k: v.detach().cpu().float() for k, v in sd.items()}
    x = mixture[rows].detach().cpu().float()
    d = dis_embed[rows].detach().cpu().float()
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = orc.net_forward(sd, ocfg, {"mixture": x, "dis_embed": d})["output"]
    cpu_s = time.perf_counter() - t0
    got = ours[rows].detach().cpu().float()
    n = min(got.shape[-1], ref.shape[-1])
    got, ref = got[..., :n], ref[..., :n]
    err = got - ref
    out = {"rows": rows, "frames_per_row": n // ocfg.stft_chunk_size, "oracle_cpu_s": cpu_s,
           "rms": orc.rms(err), "ref_rms": orc.rms(ref), "maxabs": float(err.abs().max()),
           "si_sdr_vs_oracle_db": float(orc.si_sdr(got.reshape(-1, n), ref.reshape(-1, n)).min())}
    out["rel_rms"] = out["rms"] / max(out["ref_rms"], 1e-30)
    sec = err[..., : n // sample_rate * sample_rate].reshape(len(rows), -1, sample_rate) if n >= sample_rate else err
    out["worst_second_rms"] = float(sec.double().pow(2).mean(-1).sqrt().max())
    if target is not None:
        tgt = target[rows].detach().cpu().float()[..., :n]
        a = orc.si_sdr(got[:, 0], tgt)
        b = orc.si_sdr(ref[:, 0], tgt)
        out["si_sdr_target_ours_db"] = [float(v) for v in a]
        out["si_sdr_target_oracle_db"] = [float(v) for v in b]
        out["si_sdr_delta_db"] = float((a - b).abs().max())
    out["ok"] = bool(out["rms"] <= RMS_BAR and out.get("si_sdr_delta_db", 0.0) <= SI_SDR_DELTA_BAR)
    return out
