#!/usr/bin/env python
"""HBM roofline of the batch-assembly kernel (sb_prepare_batch_fwd): algorithmic bytes / CUDA-event time vs the measured
copy bandwidth in MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sound_bubble_b200.batching import prepare_batch  # noqa: E402

dev = torch.device("cuda", 0)
B, M, V, N = 256, 6, 4, 120000
g = torch.Generator(device=dev).manual_seed(0)
mix = torch.randint(-20000, 20000, (B, M, N), generator=g, device=dev, dtype=torch.int16)
voices = torch.randint(-8000, 8000, (B, V, N), generator=g, device=dev, dtype=torch.int16)
inside = torch.zeros(B, V, dtype=torch.uint8, device=dev)
inside[:, :2] = 1                                         # two voices inside the bubble per clip
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for name, kw in (("decode + target", {}),
                 ("+ gain, drop, peak-norm", {"gain": torch.rand(B, M, device=dev) + 0.5, "drop": torch.zeros(B, M, dtype=torch.uint8, device=dev),
                                              "peak_scale": torch.full((B,), 0.5, device=dev)}),
                 ("+ shift (unaligned gather)", {"shift": torch.randint(-8, 9, (B, M), device=dev, dtype=torch.int32)})):
    ts = []
    for it in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        prepare_batch(mix, voices, inside, **kw)
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts[1:])[len(ts[1:]) // 2]
    nbytes = B * N * ((M + 2) * 2 + (M + 1) * 4) + (B * M * N * 2 if "peak_scale" in kw else 0)
    print("%-28s %7.3f ms  %7.1f GB/s algorithmic  %.2f of the measured HBM peak (%.0f GB/s)" % (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak))
