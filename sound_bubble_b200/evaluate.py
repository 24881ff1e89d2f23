"""Evaluation plumbing around the drop-in ``Net``: what src/test_samples.py does per clip, without its
third-party imports (librosa, soundfile, torchmetrics, pandas) and with the clips of one radius batched into one call.

    python -m sound_bubble_b200.evaluate <test_dir> <run_dir> [--distance_threshold 1] [--csv out.csv]

  * run directory contract (src/utils.py:60-73, 112-135; hl_module.py:115-156): ``config.json`` with
    ``pl_module_args.{model, model_params}`` + ``checkpoints/best.pt`` = ``{'model': state_dict, ...}``, loaded strict.
    The reference's dotted model path is mapped onto this package's modules.
  * sample directory contract (src/test_samples.py:35-88): ``mixture.wav`` (M channels, int16), ``metadata.json`` with
    ``voiceNN.{dis, angle}`` (+ ``real``: distances in cm), ``mic00_voiceNN.wav`` for every voice; the target is the sum
    of the voices with ``dis <= distance_threshold``.
  * metrics follow torchmetrics' definitions used by src/metrics/metrics.py:6-9, 38-55 (no mean removal for SNR /
    SI-SDR, zero-mean for SI-SNR, eps = float32 machine epsilon) and ``compute_decay`` (:20-36).
The separator itself runs on the GPU through the C-ABI library; there is no CPU path here either.
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import wave
from typing import Dict, List, Optional

import numpy as np
import torch

MODEL_PATHS = {
    "src.models.tfgridnet_realtime_clean_dis_embd3.net.Net": "sound_bubble_b200.tfgridnet_realtime_clean_dis_embd3.net.Net",
    "src.models.tfgridnet_realtime_clean_optim.net.Net": "sound_bubble_b200.tfgridnet_realtime_clean_optim.net.Net",
}
RADIUS_ONE_HOT = {1.0: [0.0, 0.0, 1.0], 1.5: [0.0, 1.0, 0.0], 2.0: [1.0, 0.0, 0.0]}     # src/test_samples.py:96-102


# ---------------------------------------------------------------------------------------------------------------
# audio files: 16-bit PCM wav <-> float32 [channels, samples]  (librosa.load(mono=False, sr=native) / sf.write PCM_16)
# ---------------------------------------------------------------------------------------------------------------
def read_wav(path: str, sr: Optional[int] = None) -> np.ndarray:
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError("%s: only 16-bit PCM is supported (got %d bytes per sample)" % (path, w.getsampwidth()))
        if sr is not None and w.getframerate() != sr:
            raise ValueError("%s: sampling rate %d != project rate %d (resampling is not implemented)"
                             % (path, w.getframerate(), sr))
        ch = w.getnchannels()
        data = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, ch)
    x = (data.astype(np.float32) / 32768.0).T
    return x[0] if ch == 1 else np.ascontiguousarray(x)


def write_wav(path: str, data: np.ndarray, sr: int):
    """data [channels, samples] or [samples], float in [-1, 1) -> PCM_16 (src/utils.py:144-151)."""
    x = np.atleast_2d(np.asarray(data, dtype=np.float32))
    pcm = np.clip(np.round(x.T * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(x.shape[0])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(pcm.tobytes())


# ---------------------------------------------------------------------------------------------------------------
# metrics (torchmetrics.functional semantics), inputs [..., T]
# ---------------------------------------------------------------------------------------------------------------
def _eps(x: torch.Tensor) -> float:
    return torch.finfo(x.dtype).eps


def snr(preds: torch.Tensor, target: torch.Tensor, zero_mean: bool = False) -> torch.Tensor:
    if zero_mean:
        target = target - target.mean(dim=-1, keepdim=True)
        preds = preds - preds.mean(dim=-1, keepdim=True)
    eps = _eps(preds)
    noise = target - preds
    return 10 * torch.log10((torch.sum(target ** 2, dim=-1) + eps) / (torch.sum(noise ** 2, dim=-1) + eps))


def si_sdr(preds: torch.Tensor, target: torch.Tensor, zero_mean: bool = False) -> torch.Tensor:
    if zero_mean:
        target = target - target.mean(dim=-1, keepdim=True)
        preds = preds - preds.mean(dim=-1, keepdim=True)
    eps = _eps(preds)
    alpha = (torch.sum(preds * target, dim=-1, keepdim=True) + eps) / (torch.sum(target ** 2, dim=-1, keepdim=True) + eps)
    scaled = alpha * target
    noise = scaled - preds
    return 10 * torch.log10((torch.sum(scaled ** 2, dim=-1) + eps) / (torch.sum(noise ** 2, dim=-1) + eps))


def si_snr(preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return si_sdr(preds, target, zero_mean=True)


def compute_decay(est: torch.Tensor, mix: torch.Tensor) -> torch.Tensor:
    """src/metrics/metrics.py:20-36: attenuation of the mixture's energy, dB, averaged over channels."""
    p_est = 10 * torch.log10(torch.sum(est ** 2, dim=-1))
    p_mix = 10 * torch.log10(torch.sum(mix ** 2, dim=-1))
    return (p_mix - p_est).mean(dim=-1)


def clip_metrics(est: torch.Tensor, gt: torch.Tensor, mix0: torch.Tensor, n_tgt: int) -> Dict[str, float]:
    """The row src/test_samples.py:175-206 builds for one clip (without STOI / PESQ, which need pystoi / pesq).
    est, gt, mix0: [1, T] (separated, target, reference-microphone mixture)."""
    row: Dict[str, float] = {"n_tgt_speakers": n_tgt}
    if n_tgt == 0:
        row["decay"] = float(compute_decay(est, mix0))
        return row
    m = lambda f, p: float(f(p, gt).mean(dim=-1))
    row["input_snr"] = m(snr, mix0)
    row["snri"] = m(snr, est) - row["input_snr"]
    row["input_sisnr"] = m(si_snr, mix0)
    row["sisnri"] = m(si_snr, est) - row["input_sisnr"]
    row["input_sisdr"] = m(si_sdr, mix0)
    row["sisdri"] = m(si_sdr, est) - row["input_sisdr"]
    return row


# ---------------------------------------------------------------------------------------------------------------
# run directory and sample directory
# ---------------------------------------------------------------------------------------------------------------
def load_run_dir(run_dir: str, device="cuda:0", checkpoint: str = "checkpoints/best.pt"):
    """src/utils.py:112-135 + hl_module.py:115-139 for the model only: config.json -> Net(**model_params), strict load."""
    import importlib
    with open(os.path.join(run_dir, "config.json")) as f:
        params = json.load(f)
    args = params["pl_module_args"]
    dotted = MODEL_PATHS.get(args["model"], args["model"])
    mod, cls = dotted.rsplit(".", 1)
    net = getattr(importlib.import_module(mod), cls)(**args["model_params"])
    ckpt = os.path.join(run_dir, checkpoint)
    if not os.path.exists(ckpt):
        raise FileNotFoundError("Given run (%s) doesn't have any pretrained checkpoints!" % run_dir)
    state = torch.load(ckpt, map_location="cpu", weights_only=False)
    net.load_state_dict(state["model"], strict=True)
    return net.to(device).eval(), params


def load_testcase(sample_dir: str, distance_threshold: float, sr: Optional[int] = 24000):
    """src/test_samples.py:35-88."""
    with open(os.path.join(sample_dir, "metadata.json"), "rb") as f:
        metadata = json.load(f)
    mixture = read_wav(os.path.join(sample_dir, "mixture.wav"), sr)
    gt = np.zeros((1, mixture.shape[-1]), dtype=np.float32)
    tgt, near, far = [], [], []
    for speaker in [k for k in metadata if k.startswith("voice")]:
        dis = metadata[speaker]["dis"] / 100 if metadata.get("real") else metadata[speaker]["dis"]
        if dis <= distance_threshold:
            gt += read_wav(os.path.join(sample_dir, "mic00_%s.wav" % speaker), sr)
            tgt.append(metadata[speaker])
            near.append(dis)
        else:
            far.append(dis)
    return metadata, mixture, gt, tgt, {"dis_near": near, "dis_far": far}


def run_testcases(net, sample_dirs: List[str], distance_threshold: float = 1.0, sr: Optional[int] = 24000,
                  max_batch: int = 32, device="cuda:0") -> List[Dict[str, float]]:
    """All clips through the separator, equal-length clips batched (the reference runs them one by one, :90-112),
    then the per-clip metric rows."""
    if float(distance_threshold) not in RADIUS_ONE_HOT:
        raise ValueError("Invalid distance threshold")
    one_hot = torch.tensor(RADIUS_ONE_HOT[float(distance_threshold)])
    cases = [load_testcase(d, distance_threshold, sr) for d in sample_dirs]
    rows: List[Optional[Dict[str, float]]] = [None] * len(cases)
    by_len: Dict[int, List[int]] = {}
    for i, c in enumerate(cases):
        by_len.setdefault(c[1].shape[-1], []).append(i)
    for _, idx in sorted(by_len.items()):
        for lo in range(0, len(idx), max_batch):
            part = idx[lo:lo + max_batch]
            mix = torch.from_numpy(np.stack([cases[i][1] for i in part])).to(device)
            inputs = {"mixture": mix, "dis_embed": one_hot.repeat(len(part), 1).to(device)}
            with torch.no_grad():
                out = net(inputs)["output"].cpu()
            for j, i in enumerate(part):
                _, mixture, gt, tgt, spatial = cases[i]
                row = {"sample": os.path.basename(sample_dirs[i])}
                row.update(clip_metrics(out[j], torch.from_numpy(gt), torch.from_numpy(mixture[0:1]), len(tgt)))
                rows[i] = row
    return rows          # type: ignore[return-value]


def summarize(rows: List[Dict[str, float]]) -> Dict[str, float]:
    keys = ("decay", "input_snr", "snri", "input_sisdr", "sisdri")
    return {k: float(np.mean([r[k] for r in rows if k in r])) for k in keys if any(k in r for r in rows)}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("test_dir", type=str, help="Path to test dataset")
    ap.add_argument("run_dir", type=str, help="Path to model run")
    ap.add_argument("--distance_threshold", type=float, default=1.0)
    ap.add_argument("--sr", type=int, default=24000)
    ap.add_argument("--csv", type=str, default=None)
    ap.add_argument("--device", type=str, default="cuda:0")
    args = ap.parse_args(argv)
    net, _ = load_run_dir(args.run_dir, device=args.device)
    sample_dirs = sorted(d for d in glob.glob(os.path.join(args.test_dir, "*")) if os.path.isdir(d))
    rows = run_testcases(net, sample_dirs, args.distance_threshold, args.sr, device=args.device)
    for r in rows:
        print(r)
    print(summarize(rows))
    if args.csv:
        keys = sorted({k for r in rows for k in r})
        with open(args.csv, "w", newline="") as f:
            wr = csv.DictWriter(f, fieldnames=keys)
            wr.writeheader()
            wr.writerows(rows)


if __name__ == "__main__":
    main()
