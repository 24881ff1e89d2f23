#!/usr/bin/env python
"""Benchmark of the Sound Bubble separator hot path on B200 (BASELINE.json: "6-mic 24kHz frames/sec (RTF)").

Workload (BASELINE.json configs[1]): batch 32 synthetic 6-mic 5 s clips @ 24 kHz per GPU, TFG_S model
(syn_experiments/finetune_stage.json), STREAMING inference in 8 ms chunks: 625 calls of the forward pass per step, each
on one [32, 6, 288] window with the state carried (protocol of the reference's edge/causal_infer.py:28-47).
One "step" = one pass over the batch = 32 x 625 = 20 000 frames (1 frame = 192 samples = 8 ms of one utterance).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  value      : frames/s, inputs already resident in HBM when the timed region starts (device-side windows).
  e2e        : frames/s through the public API with HOST buffers: per chunk one H2D copy of the window from pinned
               memory, one CUDA-graph launch, one D2H copy of the separated chunk; synchronised once per step.
  roofline   : the dominant kernel of the timed region, timed with CUDA events on the launching stream.
  cpu_baseline / --impl reference : the oracle port of the reference (oracle/, PyTorch CPU = the reference's own CPU
               path: same aten LSTM/conv kernels) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# /root/reference/syn_experiments/{pretrain,finetune}_stage.json:8-27 (the "TFG_S" configuration)
SYN = dict(stft_chunk_size=192, stft_pad_size=96, num_ch=6, D=32, L=4, I=1, J=1, B=6, H=64, E=2,
           local_atten_len=100, use_attn=False, lookahead=True, chunk_causal=True, use_first_ln=True,
           merge_method="early_cat", conv_lstm=False, dis_type="conv3")
BATCH = 32                  # utterances per GPU
N_SAMPLES = 120000          # 5 s @ 24 kHz
CHUNK, LOOK, NFFT, MICS = 192, 96, 288, 6
T_FRAMES = N_SAMPLES // CHUNK       # 625
WORKLOAD = "batch32_6mic_5s_24kHz_TFG_S_streaming_8ms_chunks"
METRIC, UNIT = "6mic_24kHz_frames_per_sec", "frames/s"


def synthetic_clips(batch, seed):
    """0.1*randn source with per-microphone delays and gains + independent noise."""
    g = torch.Generator().manual_seed(seed)
    src = 0.1 * torch.randn(batch, N_SAMPLES + 8, generator=g)
    noise = 0.02 * torch.randn(batch, MICS, N_SAMPLES, generator=g)
    gains = 0.5 + torch.rand(batch, MICS, generator=g)
    out = torch.empty(batch, MICS, N_SAMPLES)
    for m in range(MICS):
        d = (3 * m + 1) % 9
        out[:, m] = gains[:, m:m + 1] * src[:, d:d + N_SAMPLES]
    return out + noise


def radius_one_hot(batch):
    table = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])      # 1 m / 1.5 m / 2 m
    return table[torch.arange(batch) % 3].clone()


def windows_of(mix):
    """[B, M, N] -> [T, B, M, 288]: the window every streaming call sees (lookahead pad of 96 zeros at the end)."""
    padded = torch.nn.functional.pad(mix, (0, LOOK))
    return padded.unfold(-1, NFFT, CHUNK).permute(2, 0, 1, 3).contiguous()


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: oracle port of the reference on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_streaming_sample(sd, n_chunks, warm=2, seed=1234):
    """Times `n_chunks` streaming calls at batch 32 through the oracle (PyTorch CPU).  Returns seconds."""
    from oracle import tfgridnet_oracle as orc
    ocfg = orc.OracleConfig.from_kwargs("dis_embed", **SYN)
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(BATCH, MICS, CHUNK * (n_chunks + warm) + LOOK, generator=g)
    dis = radius_one_hot(BATCH)
    st = orc.init_state(ocfg, BATCH)
    t0 = None
    with torch.no_grad():
        for t in range(n_chunks + warm):
            if t == warm:
                t0 = time.perf_counter()
            r = orc.net_forward(sd, ocfg, {"mixture": x[..., t * CHUNK: t * CHUNK + NFFT], "dis_embed": dis}, st, pad=False)
            st = r["next_state"]
    return time.perf_counter() - t0


def cpu_training_sample(sd, seconds=1.0, seed=77):
    """One training step (forward + backward under torch autograd) of ONE clip of `seconds` through the oracle port on the
    host cores: the CPU figure beside the "train" leg.  Returns (seconds of wall time, frames)."""
    from oracle import tfgridnet_oracle as orc
    ocfg = orc.OracleConfig.from_kwargs("dis_embed", **SYN)
    torch.set_num_threads(os.cpu_count() or 1)
    leaf = {k: v.detach().clone().requires_grad_("_filters" not in k) for k, v in sd.items()}
    n = int(seconds * 24000) // CHUNK * CHUNK
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(1, MICS, n, generator=g)
    tgt = 0.1 * torch.randn(1, 1, n, generator=g)
    t0 = time.perf_counter()
    est = orc.net_forward(leaf, ocfg, {"mixture": x, "dis_embed": radius_one_hot(1)})["output"]
    (-(10 * torch.log10(tgt.pow(2).sum(-1) / ((est - tgt).pow(2).sum(-1) + 1e-8))).mean()).backward()
    return time.perf_counter() - t0, n // CHUNK


def reference_weights():
    from oracle import tfgridnet_oracle as orc
    from oracle.weights import make_state_dict
    return make_state_dict(orc.OracleConfig.from_kwargs("dis_embed", **SYN), 0)


def run_reference(args, rank):
    if rank != 0:
        return
    sd = reference_weights()
    cores = os.cpu_count() or 1
    per_chunk = cpu_streaming_sample(sd, 3, warm=2) / 3.0
    budget = 150.0 / max(args.steps + args.warmup, 1)             # whole run within a few minutes
    n = int(max(2, min(T_FRAMES, budget / max(per_chunk, 1e-6))))
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_streaming_sample(sd, n, warm=1, seed=1234 + i)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = BATCH * n * len(times) / total
    sample = "first %d of %d chunks of the batch-32 streaming pass per step" % (n, T_FRAMES)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": (total / len(times)) / (BATCH * n * CHUNK / 24000.0),
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "clip_seconds": 5.0, "frames_per_step": BATCH * n,
                   "sample": sample, "engine": "oracle port of the reference on PyTorch CPU (torch %s)" % torch.__version__},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def stage_profile(lib, fn):
    """Runs fn() with the library's per-stage CUDA-event timing armed; returns {stage: (total_ms, calls)}."""
    from sound_bubble_b200 import _abi as abi
    ms = (ctypes.c_double * len(abi.SB_STAGES))()
    calls = (ctypes.c_int64 * len(abi.SB_STAGES))()
    lib.sb_profile_begin()
    fn()
    torch.cuda.synchronize()
    abi.check(lib, lib.sb_profile_end(ms, calls), "sb_profile_end")
    return {name: (ms[i], calls[i]) for i, name in enumerate(abi.SB_STAGES) if calls[i]}


def ncu_dram_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed summary of its
    `ncu --set full` capture (profiles/rNN_prof_*.txt, written by tools/summarize_profiles.py).  None if there is none."""
    import glob
    import re
    vals = []
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_*.txt"))):
        cur = None
        for line in open(path):
            if line.startswith("== "):
                cur = {"name": line, "r": None, "w": None}
            m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m and cur is not None and kernel_substr in cur["name"]:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 1.0)
                cur["r" if m.group(1) == "read" else "w"] = float(m.group(2)) * scale
                if cur["r"] is not None and cur["w"] is not None:
                    vals.append(cur["r"] + cur["w"])
                    cur = None
    return sum(vals) / len(vals) if vals else None


def training_leg(dev, rank, world, batch=8, steps=3):
    """BASELINE config 4's shape, for context next to the headline: one data-parallel training step of the same TFG_S
    model (forward + hand-written backward kernels + ONE NCCL all-reduce of the flat gradient + clip + Adam) on `batch`
    5 s clips per GPU, fp32.  Device-timed, max over ranks.  A failure is reported in the line, never hidden."""
    import torch.distributed as dist
    from sound_bubble_b200 import Net
    from sound_bubble_b200.train_dist import FlatGradReducer, backprop
    ok, err, step = 1, "", None
    try:
        torch.manual_seed(0)
        tnet = Net(**SYN).to(dev).train()
        g = torch.Generator().manual_seed(99 + rank)
        mix = (0.1 * torch.randn(batch, MICS, N_SAMPLES, generator=g)).to(dev)
        tgt = (0.1 * torch.randn(batch, 1, N_SAMPLES, generator=g)).to(dev)
        dis = radius_one_hot(batch).to(dev)
        red = FlatGradReducer(tnet.parameters())
        opt = torch.optim.Adam(tnet.parameters(), lr=1e-3)

        def local_grads():
            red.zero_grad()
            est = tnet({"mixture": mix, "dis_embed": dis})["output"]
            loss = -(10 * torch.log10(tgt.pow(2).sum(-1) / ((est - tgt).pow(2).sum(-1) + 1e-8))).mean()
            loss.backward()
            return loss

        def step():
            loss = local_grads()
            backprop(red, opt, grad_clip=1.0)                      # all-reduce (mean) -> clip -> Adam
            return loss
        local_grads()                                              # warm-up without a collective
        torch.cuda.synchronize(dev)
    except Exception as e:                                         # noqa: BLE001 - reported in the JSON line
        ok, err = 0, "%s: %s" % (type(e).__name__, str(e)[:200])
    flag = torch.tensor([ok], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)                # every rank takes the same branch below
    if int(flag.item()) == 0:
        return {"error": err or "a peer rank failed"}
    step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    b.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    frames = batch * T_FRAMES * world
    return {"value": frames / (ms * 1e-3), "unit": "training frames/s", "ms_per_step": ms, "global_batch": batch * world,
            "clip_seconds": 5.0, "dtype": "f32", "steps": steps, "loss": float(loss.detach()),
            "collective": "one all-reduce of %d fp32 gradients per step (NCCL)" % red.numel if world > 1 else "none (1 GPU)",
            "note": "forward + backward through csrc/sb_train.cu, clip after the reduction, Adam; see tools/train_bench.py"}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from sound_bubble_b200 import Net, _lib
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    if args.pdl:
        _lib.set_pdl(True)

    torch.manual_seed(0)
    net = Net(**SYN).to(dev).eval()                                # random-init weights of the TFG_S architecture
    mix = synthetic_clips(BATCH, 1234 + rank)
    dis = radius_one_hot(BATCH).to(dev)
    win_host = windows_of(mix).pin_memory()                        # [T, B, M, 288]
    win_dev = win_host.to(dev)
    out_dev = torch.empty(T_FRAMES, BATCH, 1, CHUNK, device=dev)
    out_host = torch.empty(T_FRAMES, BATCH, 1, CHUNK).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)        # > 126 MB L2
    sess = net.streaming(BATCH, dis, use_graph=not args.no_graph)
    launches_per_chunk = sess.launches_per_step()
    sess.reset()
    stream = torch.cuda.current_stream(dev)
    # throughput mode: the same one-call-per-chunk protocol, calls asynchronous, consecutive chunks overlapping on two
    # streams with per-unit dependencies (sound_bubble_b200/streaming.py::PipelinedSession)
    pipe = net.streaming(BATCH, dis, pipelined=True, ranges=args.ranges or None, depth=args.depth,
                         intra_algo=args.pipe_intra_algo or None, inter_algo=args.pipe_inter_algo or None) if args.pipeline else None

    def pass_in_order(win, out):
        sess.reset()
        for t in range(T_FRAMES):
            sess.x.copy_(win[t], non_blocking=True)
            sess.step()
            out[t].copy_(sess.y, non_blocking=True)

    enqueue_s = []

    def pass_pipelined(win, out):
        pipe.reset()
        pipe.begin()
        t0 = time.perf_counter()
        for t in range(T_FRAMES):
            pipe.feed(win[t], out=out[t])
        enqueue_s.append(time.perf_counter() - t0)
        pipe.end()

    one_pass = pass_pipelined if pipe is not None else pass_in_order

    def pass_device():
        one_pass(win_dev, out_dev)
        return out_dev.permute(1, 2, 0, 3).reshape(BATCH, 1, N_SAMPLES)

    def pass_host():
        one_pass(win_host, out_host)
        stream.synchronize()
        return out_host

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        times = []
        barrier()
        for _ in range(steps):
            flush.zero_()                                          # flush L2 between timed iterations
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            times.append(a.elapsed_time(b))
        barrier()
        t = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)               # max over ranks, device-timed
        return float(t.item()) / steps                             # ms per step

    with ClockSampler(local_rank) as clk:
        ms_dev = timed(pass_device, args.steps, args.warmup)
    clocks = clk.summary()
    ms_e2e = timed(pass_host, max(2, args.steps // 2), 1)
    ms_in_order = timed(lambda: pass_in_order(win_dev, out_dev), 2, 1) if pipe is not None else ms_dev
    frames = BATCH * T_FRAMES * world
    value = frames / (ms_dev * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)

    train_info = None if args.no_train else training_leg(dev, rank, world)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel, timed live with CUDA events (eager replay of the same per-chunk sequence) ----
    n_prof = 100
    eng = net.engine()
    saved_algos = (eng.intra_algo, eng.inter_algo)
    if pipe is not None:                                           # the kernel families the timed region ran
        eng.intra_algo = pipe.intra_algo if pipe.intra_algo is not None else eng.intra_algo
        eng.inter_algo = pipe.inter_algo if pipe.inter_algo is not None else eng.inter_algo
    sess_e = net.streaming(BATCH, dis, use_graph=False)

    def eager_chunks():
        for t in range(n_prof):
            sess_e.x.copy_(win_dev[t], non_blocking=True)
            sess_e.step()
    eager_chunks()
    prof = stage_profile(lib, eager_chunks)
    eng.intra_algo, eng.inter_algo = saved_algos
    tot = sum(v[0] for v in prof.values())
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms = prof[dom][0] / prof[dom][1]
    F, C, H = NFFT // 2 + 1, SYN["D"], SYN["H"]
    act = F * C * 4                                                # one [F][C] fp32 slab = 18 560 B
    alg_bytes = {"intra": 2 * act, "inter": 2 * act + 2 * 2 * F * H * 4, "stft_features": MICS * CHUNK * 4 + F * 27 * 4,
                 "conv_in": F * 27 * 4 + act, "backend": act + CHUNK * 4, "film_params": 0}
    alg_flops = {"intra": 2 * (2 * F * (4 * H * (C + H)) + F * 2 * H * C), "inter": 2 * (F * 4 * H * (C + H) + F * H * C)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_bw = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes.get(dom, 0) * BATCH / (dom_ms * 1e-3) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak_bw, "unit": "GB/s",
                "frac": achieved / peak_bw,
                "traffic": ncu_dram_traffic({"intra": "lstm_ws_kernel<32, 0, 2>" if pipe is not None and pipe.intra_algo == 8
                                             else "lstm_ws_kernel<32, 0, 1>", "inter": "lstm_t"}.get(dom, dom)),
                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                "avg_launch_us": dom_ms * 1e3, "bytes_per_launch": alg_bytes.get(dom, 0) * BATCH,
                "share_of_step": prof[dom][0] / tot,
                "fp32_tflops": alg_flops.get(dom, 0) * BATCH / (dom_ms * 1e-3) / 1e12,
                "note": "latency-bound: 145 dependent LSTM steps per launch (two sequences per CTA in the pipelined session); see DESIGN.md",
                "stage_us_per_chunk": {k: 1e3 * v[0] / n_prof for k, v in prof.items()}}

    # ---- offline (whole-utterance) pass of the same clips, for context ----
    x_dev = mix.to(dev)
    inputs = {"mixture": x_dev, "dis_embed": dis}

    def offline():
        return net(inputs)["output"]
    y_off = offline()
    y_str = pass_device()
    stream_vs_offline = float((y_str - y_off).abs().max())
    ms_off = timed(offline, 3, 1) if world == 1 else None
    offline_info = None
    if ms_off is not None:
        net.pipeline_offline = False                               # the single whole-utterance call, with its stage breakdown
        ms_single = timed(offline, 3, 1)
        oprof = stage_profile(lib, offline)
        net.pipeline_offline = True
        offline_info = {"value": BATCH * T_FRAMES / (ms_off * 1e-3), "unit": UNIT, "ms_per_step": ms_off,
                        "rtf": ms_off * 1e-3 / (BATCH * 5.0),
                        "note": "Net.forward on whole clips = %d-frame time slices through the native pipe" % net.offline_slice_frames,
                        "single_call": {"value": BATCH * T_FRAMES / (ms_single * 1e-3), "ms_per_step": ms_single,
                                        "stage_ms": {k: v[0] for k, v in oprof.items()}}}

    cpu = None
    if world == 1 and not args.no_cpu:
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        n_cpu = 300
        dt = cpu_streaming_sample(sd, n_cpu, warm=2)
        cpu = {"value": BATCH * n_cpu / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "first %d of %d chunks of the batch-32 streaming pass (%.1f s of CPU work)" % (n_cpu, T_FRAMES, dt)}
        if train_info is not None and "error" not in train_info:
            try:
                dt_t, fr_t = cpu_training_sample(sd)
                cpu["train"] = {"value": fr_t / dt_t, "unit": "training frames/s", "kind": "port", "cores": os.cpu_count() or 1,
                                "sample": "forward + backward of 1 clip x 1 s under torch autograd (%.2f s of CPU work)" % dt_t}
            except Exception as e:                                 # noqa: BLE001 - an extra figure must not cost the headline line
                cpu["train"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "rtf": ms_dev * 1e-3 / (BATCH * 5.0),
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "clip_seconds": 5.0, "frames_per_step": frames,
                   "chunks_per_step": T_FRAMES, "weights": "random init (seed 0) of the TFG_S architecture",
                   "l2": "256 MB buffer written between timed steps", "cuda_graph": not args.no_graph, "pdl": bool(args.pdl),
                   "pipelined": pipe is not None, "unit_ranges": pipe.ranges if pipe is not None else None,
                   "pipeline_depth": pipe.depth if pipe is not None else 1,
                   "pipeline_intra_algo": pipe.intra_algo if pipe is not None else None,
                   "pipeline_inter_algo": pipe.inter_algo if pipe is not None else None,
                   "host_enqueue_ms_per_step": 1e3 * statistics.median(enqueue_s) if enqueue_s else None,
                   "parallelism": "dp%d (utterances sharded, no collective on the data path)" % world},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(win_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4)},
        "gpu_launches": int(launches_per_chunk * T_FRAMES * args.steps),
        "launches_per_chunk": int(launches_per_chunk),
        "in_order": {"value": frames / (ms_in_order * 1e-3), "unit": UNIT, "ms_per_step": ms_in_order,
                     "note": "one stream, chunk t+1 starts when chunk t has finished"},
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "offline": offline_info,
        "streaming_vs_offline_maxabs": stream_vs_offline, "train": train_info,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pdl", type=int, default=int(os.environ.get("SB_PDL", "1")))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pipeline", type=int, default=int(os.environ.get("SB_PIPELINE", "1")))
    ap.add_argument("--ranges", type=int, default=0, help="pipelined session: number of unit ranges (0 = default)")
    ap.add_argument("--pipe-intra-algo", type=int, default=0, help="pipelined session: force an SB_ALGO_* for the intra path")
    ap.add_argument("--pipe-inter-algo", type=int, default=0, help="pipelined session: force an SB_ALGO_* for the inter path")
    ap.add_argument("--depth", type=int, default=8, help="pipelined session: chunks in flight")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"               # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
