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

# 16-32 streams of the pipelined session must not alias onto the default 8 hardware work queues (read at context creation)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402

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


def synthetic_target(batch, seed):
    """The clean source as the reference microphone (mic 0, delay 1) sees it: the synthetic target of the SI-SDR figures."""
    g = torch.Generator().manual_seed(seed)
    src = 0.1 * torch.randn(batch, N_SAMPLES + 8, generator=g)
    return src[:, 1:1 + N_SAMPLES].clone()


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
def cpu_engine():
    """("reference", description) when oracle/_ref holds the unmodified reference module (oracle/build_ref.py), else
    ("port", description): the oracle restatement, which dispatches to the same aten CPU kernels."""
    from oracle import ref_runner
    if ref_runner.available():
        return "reference", ("UNMODIFIED reference module (oracle/_ref copy of src/models/tfgridnet_realtime_clean_dis_embd3, "
                             "asteroid/espnet stand-ins from oracle/shims) on PyTorch CPU (torch %s)" % torch.__version__)
    return "port", "oracle port of the reference on PyTorch CPU (torch %s)" % torch.__version__


def cpu_streaming_sample(sd, n_chunks, warm=2, seed=1234):
    """Times `n_chunks` streaming calls at batch 32 (edge/causal_infer.py:28-47 protocol) on the host cores through the
    unmodified reference module when oracle/_ref is present, else through the oracle port.  Returns seconds."""
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(BATCH, MICS, CHUNK * (n_chunks + warm) + LOOK, generator=g)
    dis = radius_one_hot(BATCH)
    from oracle import ref_runner
    if ref_runner.available():
        net = ref_runner.reference_net(SYN, sd)
        return ref_runner.streaming_sample(net, x, dis, CHUNK, LOOK, n_chunks, warm)[0]
    from oracle import tfgridnet_oracle as orc
    ocfg = orc.OracleConfig.from_kwargs("dis_embed", **SYN)
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
    kind, engine = cpu_engine()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": (total / len(times)) / (BATCH * n * CHUNK / 24000.0),
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "clip_seconds": 5.0, "frames_per_step": BATCH * n,
                   "sample": sample, "engine": engine},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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


def ncu_dram_traffic(kernel_substr, grid_substr=None):
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
            if m and cur is not None and kernel_substr in cur["name"] and (grid_substr is None or grid_substr in cur["name"]):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 1.0)
                cur["r" if m.group(1) == "read" else "w"] = float(m.group(2)) * scale
                if cur["r"] is not None and cur["w"] is not None:
                    vals.append(cur["r"] + cur["w"])
                    cur = None
    return sum(vals) / len(vals) if vals else None


def training_leg(dev, rank, world, batch=8, steps=5):
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
    # the reduced gradient is the gradient of the GLOBAL-batch mean loss: 2 clips x 1 s per rank through the kernels and the
    # NCCL all-reduce, against the same world x 2 clips in one batch on rank 0
    grad_relerr = None
    if world > 1:
        n1 = 24000 // CHUNK * CHUNK
        gg = torch.Generator().manual_seed(4242)
        mix_g = 0.1 * torch.randn(2 * world, MICS, n1, generator=gg)
        tgt_g = 0.1 * torch.randn(2 * world, 1, n1, generator=gg)
        dis_g = radius_one_hot(2 * world)

        def snr_grads(lo, hi):
            red.zero_grad()
            est = tnet({"mixture": mix_g[lo:hi].to(dev), "dis_embed": dis_g[lo:hi].to(dev)})["output"]
            t = tgt_g[lo:hi].to(dev)
            (-(10 * torch.log10(t.pow(2).sum(-1) / ((est - t).pow(2).sum(-1) + 1e-8))).mean()).backward()
        snr_grads(2 * rank, 2 * rank + 2)
        red.all_reduce_mean(2)
        reduced = red.flat.clone()
        if rank == 0:
            snr_grads(0, 2 * world)
            grad_relerr = float((red.flat - reduced).abs().max() / red.flat.abs().max())
        dist.barrier()
    for _ in range(2):                                             # the caching allocator settles on the step's 22 GB working set
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    per_step = []
    for _ in range(steps):                                         # every step device-timed on its own; the median is reported
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # (a step that has to go back to
        a.record()                                                 # cudaMalloc after the other legs' allocations costs 1.5x)
        loss = step()
        b.record()
        b.synchronize()
        per_step.append(a.elapsed_time(b))
    t = torch.tensor([statistics.median(per_step)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    frames = batch * T_FRAMES * world
    return {"value": frames / (ms * 1e-3), "unit": "training frames/s", "ms_per_step": ms, "global_batch": batch * world,
            "clip_seconds": 5.0, "dtype": "f32", "steps": steps, "ms_each_step": per_step, "loss": float(loss.detach()),
            "grad_relerr_vs_global_batch": grad_relerr,
            "collective": "one all-reduce of %d fp32 gradients per step (NCCL)" % red.numel if world > 1 else "none (1 GPU)",
            "note": "forward + backward through csrc/sb_train.cu, clip after the reduction, Adam; see tools/train_bench.py"}


def library_baseline(net, mix, dis, dev, n_chunks=40):
    """Context, never the product path: the UNMODIFIED reference module (oracle/_ref; else the oracle port) on the SAME
    B200 through stock PyTorch (cuDNN LSTM, cuBLAS, aten) - the library bar SURVEY.md section 2b names.  Whole-clip call at
    batch 32 and a sample of the chunk-by-chunk protocol; device-timed."""
    from oracle import ref_runner
    out = {"engine": "stock PyTorch %s / cuDNN %s on the same GPU" % (torch.__version__, torch.backends.cudnn.version())}
    try:
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        if ref_runner.available():
            ref = ref_runner.reference_net(SYN, sd).to(dev).eval()
            out["kind"] = "reference"
            fwd = lambda x, st=None, pad=True: ref({"mixture": x, "dis_embed": dis}, st, pad=pad)       # noqa: E731
            init = lambda: ref.init_buffers(BATCH, dev)                                                # noqa: E731
        else:
            from oracle import tfgridnet_oracle as orc
            ocfg = orc.OracleConfig.from_kwargs("dis_embed", **SYN)
            sdd = {k: v.to(dev) for k, v in sd.items()}
            out["kind"] = "port"
            fwd = lambda x, st=None, pad=True: orc.net_forward(sdd, ocfg, {"mixture": x, "dis_embed": dis}, st, pad=pad)   # noqa: E731

            def init():
                st = orc.init_state(ocfg, BATCH)
                mv = lambda d: {k: (mv(v) if isinstance(v, dict) else v.to(dev)) for k, v in d.items()}  # noqa: E731
                return mv(st)
        x = mix.to(dev)
        ev = lambda: torch.cuda.Event(enable_timing=True)      # noqa: E731
        with torch.no_grad():
            y = fwd(x)["output"]
            torch.cuda.synchronize(dev)
            t0, t1 = ev(), ev()
            t0.record()
            for _ in range(2):
                y = fwd(x)["output"]
            t1.record(); t1.synchronize()
            ms_off = t0.elapsed_time(t1) / 2
            xp = torch.nn.functional.pad(x, (0, LOOK))
            st = init()
            for t in range(3):
                st = fwd(xp[..., t * CHUNK: t * CHUNK + NFFT], st, pad=False)["next_state"]
            torch.cuda.synchronize(dev)
            t0, t1 = ev(), ev()
            t0.record()
            for t in range(3, 3 + n_chunks):
                st = fwd(xp[..., t * CHUNK: t * CHUNK + NFFT], st, pad=False)["next_state"]
            t1.record(); t1.synchronize()
            ms_chunk = t0.elapsed_time(t1) / n_chunks
        out.update({"offline": {"value": BATCH * T_FRAMES / (ms_off * 1e-3), "unit": UNIT, "ms_per_step": ms_off},
                    "streaming": {"value": BATCH / (ms_chunk * 1e-3), "unit": UNIT, "us_per_chunk": 1e3 * ms_chunk,
                                  "sample": "%d chunks of the batch-32 streaming pass" % n_chunks},
                    "output_rms": float(y.float().pow(2).mean().sqrt())})
        out["_y"] = y
    except Exception as e:                                         # noqa: BLE001 - context figure, reported not hidden
        out["error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
    return out


def strong_leg(net, dev, rank, world, global_batch=256):
    """BASELINE config 3: a FIXED job of 256 synthetic 6-mic 5 s clips (radii 1 / 1.5 / 2 m), sharded over the ranks
    (256 / N per GPU, whole-clip calls, no collective).  Device-timed, max over ranks: the driver's per-N lines give the
    strong-scaling curve."""
    import torch.distributed as dist
    nb = global_batch // world
    try:
        g = torch.Generator().manual_seed(4321 + rank)
        x = (0.1 * torch.randn(nb, MICS, N_SAMPLES, generator=g)).to(dev)
        dis = radius_one_hot(global_batch)[rank * nb:(rank + 1) * nb].to(dev).contiguous()
        inputs = {"mixture": x, "dis_embed": dis}
        with torch.no_grad():
            for _ in range(2):
                y = net(inputs)["output"]
        torch.cuda.synchronize(dev)
        ok = 1
    except Exception as e:                                         # noqa: BLE001
        ok, err = 0, "%s: %s" % (type(e).__name__, str(e)[:200])
    flag = torch.tensor([ok], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return {"error": err if not ok else "a peer rank failed"}
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    a.record()
    with torch.no_grad():
        for _ in range(steps):
            y = net(inputs)["output"]
    b.record(); b.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"global_batch": global_batch, "batch_per_gpu": nb, "ms_per_job": ms, "value": global_batch * T_FRAMES / (ms * 1e-3),
            "unit": UNIT, "scaling": "strong", "finite": bool(torch.isfinite(y).all()),
            "note": "Net.forward on whole 5 s clips, %d per GPU; speed-up over the N=1 line of the same key = strong scaling" % nb}


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
    sess.reset()
    stream = torch.cuda.current_stream(dev)
    # throughput mode: the same one-call-per-8-ms-chunk protocol, calls asynchronous; the native pipe gathers `group`
    # consecutive chunks per launch and keeps `depth` groups in flight with per-unit dependencies
    # (sound_bubble_b200/streaming.py::PipelinedSession, csrc/sb_pipe.cu)
    pipe = net.streaming(BATCH, dis, pipelined=True, ranges=args.ranges or None, depth=args.depth, group=args.group,
                         intra_algo=args.pipe_intra_algo or None, inter_algo=args.pipe_inter_algo or None) if args.pipeline else None
    G = pipe.group if pipe is not None else 1
    launches_per_call = pipe.launches_per_step() if pipe is not None else sess.launches_per_step()
    sess.reset()

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

    def pass_host():
        one_pass(win_host, out_host)
        stream.synchronize()

    def as_wave(o):
        return o.permute(1, 2, 0, 3).reshape(BATCH, 1, N_SAMPLES)

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
    y_timed = as_wave(out_dev).clone()                             # the output of the LAST TIMED step (parity below)
    ms_e2e = timed(pass_host, max(2, args.steps // 2), 1)
    y_e2e = as_wave(out_host.to(dev))
    ms_in_order = timed(lambda: pass_in_order(win_dev, out_dev), 2, 1) if pipe is not None else ms_dev
    frames = BATCH * T_FRAMES * world
    value = frames / (ms_dev * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)

    train_info = None if args.no_train else training_leg(dev, rank, world)
    strong_info = None if args.no_strong else strong_leg(net, dev, rank, world)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    # ---- parity of the timed run's own output against the CPU oracle (checker only) -------------------------------
    parity = None
    if not args.no_parity:
        from oracle.headline import compare_with_oracle
        rows = [0, 13, 29]                                         # one utterance per bubble radius
        sd_cpu = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        parity = compare_with_oracle(sd_cpu, SYN, mix, dis.cpu(), y_timed, rows, target=synthetic_target(BATCH, 1234 + rank))
        parity["e2e_vs_device_maxabs"] = float((y_e2e - y_timed).abs().max())
        parity["path"] = "output of the last timed step (pipelined session, group %d, depth %d)" % (G, pipe.depth if pipe else 1)
        parity["bar"] = "north_star: rms <= 1e-3 and |SI-SDR delta| <= 0.05 dB"

    # ---- roofline: the launch sequence of ONE group exactly as the timed region runs it (same T = G call, same kernel
    # families), replayed eagerly in order on one stream with CUDA events around every stage (sb_profile_*) ----------
    n_prof = 12
    F, C, H, NB = NFFT // 2 + 1, SYN["D"], SYN["H"], SYN["B"]
    if pipe is not None:
        call = pipe._call(0, 0)
        arena_saved = [a.flat.clone() for a in pipe.arenas]

        def replay():
            for _ in range(n_prof):
                call.launch()
    else:
        arena_saved = None

        def replay():
            for t in range(n_prof):
                sess.x.copy_(win_dev[t], non_blocking=True)
                sess._step_eager(sess.parity)
    replay()
    prof = stage_profile(lib, replay)
    if arena_saved is not None:
        for a, z in zip(pipe.arenas, arena_saved):
            a.flat.copy_(z)
    calls_per_step = T_FRAMES / G                                   # launches of each stage's sequence per step
    per_call_ms = {k: v[0] / n_prof for k, v in prof.items()}       # one group's stage time (all 6 blocks)
    tot = sum(per_call_ms.values())
    dom = max(per_call_ms, key=lambda k: per_call_ms[k])
    dom_ms = prof[dom][0] / prof[dom][1]                            # average duration of ONE launch of the dominant stage
    act = F * C * 4                                                # one [F][C] fp32 slab = 18 560 B
    alg_bytes = {"intra": 2 * act, "inter": 2 * act + 2 * 2 * F * H * 4 / G, "stft_features": MICS * CHUNK * 4 + F * 27 * 4,
                 "conv_in": F * 27 * 4 + act, "backend": act + CHUNK * 4, "film_params": 0}       # per (utterance, frame)
    alg_flops = {"intra": 2 * (2 * F * (4 * H * (C + H)) + F * 2 * H * C), "inter": 2 * (F * 4 * H * (C + H) + F * H * C)}
    units = BATCH * G                                              # (utterance, frame) pairs one launch processes
    fl, by = alg_flops.get(dom, 0) * units, alg_bytes.get(dom, 0) * units
    t_s = dom_ms * 1e-3
    tc_family = pipe is not None and dom in ("intra", "inter") and \
        (pipe.intra_algo if dom == "intra" else pipe.inter_algo) in (7, 9)
    bw_peak = float(peaks.get("hbm_gbs", 6650.0))
    tensor_peak = float(peaks.get("bf16_tflops", 1684.0))           # burst figure: the kernel is timed alone
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12                      # 74.4 TFLOP/s fp32 FMA (tools/ubench: 72.7 measured)
    fracs = {"hbm": by / t_s / 1e9 / bw_peak, "tensor": (3 * fl / t_s / 1e12 / tensor_peak) if tc_family else 0.0,
             "fp32_fma": 0.0 if tc_family else fl / t_s / 1e12 / fp32_peak}
    kernel_name = {"intra": "lstm_tcr_kernel" if tc_family else "lstm_ws_kernel", "inter": "lstm_tcr_kernel" if tc_family else "lstm_tile_kernel"}.get(dom, dom)
    # the roof this kernel actually leans on: the XU (MUFU) pipe, 16 lanes per clock and SM; 7 ex2 / rcp per LSTM cell
    cells = {"intra": 2 * F * H, "inter": F * H}.get(dom, 0) * units
    n_ctas = {"intra": 2 * -(-units // 128), "inter": BATCH * (F // 128) + -(-BATCH // (128 // (F % 128)))}.get(dom, 148)
    fracs["xu"] = (7 * cells / t_s) / (min(n_ctas, 148) * 16 * 1.965e9) if tc_family else 0.0      # of the SMs the launch occupies
    if tc_family:
        bound, achieved, peak, unit = "tensor", fl / t_s / 1e12, tensor_peak, "TFLOP/s"
    elif dom in alg_flops:
        bound, achieved, peak, unit = "fp32_fma", fl / t_s / 1e12, fp32_peak, "TFLOP/s"
    else:
        bound, achieved, peak, unit = "hbm", by / t_s / 1e9, bw_peak, "GB/s"
    n_tiles = {"intra": 2 * -(-units // 128), "inter": BATCH * (F // 128) + -(-BATCH // (128 // (F % 128)))} if tc_family else {}
    step_flops = (sum(alg_flops.values())) * NB * BATCH * T_FRAMES
    step_bytes = (alg_bytes["stft_features"] + alg_bytes["conv_in"] + alg_bytes["backend"] + NB * (alg_bytes["intra"] + alg_bytes["inter"])) * BATCH * T_FRAMES
    sm_ms = sum(per_call_ms[k] * (min(n_tiles.get(k, 148), 148) / 148.0) for k in per_call_ms) * calls_per_step
    roofline = {
        "kernel": "%s (%s stage)" % (kernel_name, dom), "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
        "frac": achieved / peak,
        # the capture of the same launch shape (profiles/r02_prof_tcr.txt: 32 utterances x G frames, tools/gpu_r2_final.sh)
        "traffic": ncu_dram_traffic("lstm_tcr_kernel", "grid (%d, %d, 1)" % ((n_ctas // 2, 2) if dom == "intra" else (n_ctas, 1))) if tc_family
        else ncu_dram_traffic({"intra": "lstm_ws_kernel<32, 0, 2>", "inter": "lstm_t"}.get(dom, dom)),
        "peak_source": ("MEASURED_PEAKS.json (measured)" if peaks else "fallback of B200_PROFILING.md") + (
            "; fp32 FMA peak = 148 SMs x 128 FMA/clk x 1.965 GHz (tools/ubench measured 72.7)" if bound == "fp32_fma" else ""),
        "avg_launch_us": dom_ms * 1e3, "launch_units": "%d utterances x %d frames" % (BATCH, G),
        "flops_per_launch": fl, "bytes_per_launch": by,
        "timing": "CUDA events on the launching stream around every stage (sb_profile_*), eager in-order replay of the SAME "
                  "T = %d launch sequence the timed region runs as graphs; share checked against profiles/r02_launches_bench.txt" % G,
        "share_of_group": per_call_ms[dom] / tot,
        "all_roofs": {"hbm_frac": fracs["hbm"], "tensor_frac_executed_3term": fracs["tensor"], "fp32_fma_frac": fracs["fp32_fma"],
                      "xu_frac": fracs["xu"], "max": max(fracs.values()),
                      "note": "tcgen05 path: algorithmic 2*MAC FLOPs in `achieved`; the bf16 hi/lo three-term split executes 3x that on "
                              "the tensor pipe.  The recurrence is serial in the step: what bounds a step is the XU pipe (7 MUFU per "
                              "cell; xu_frac = 7 x cells / launch time / (CTAs of the launch x 16 lanes x 1.965 GHz), one CTA per SM) "
                              "plus the three MMAs that cannot overlap the cell update.  ncu "
                              "(profiles/r02_prof_tcr.txt): XU 58 %, tensor pipe 40 %, issue 53 % of active cycles"},
        "stage_us_per_group": {k: 1e3 * v for k, v in per_call_ms.items()},
        "step": {"alg_tflop": step_flops / 1e12, "alg_gbyte": step_bytes / 1e9, "tflops": step_flops / (ms_dev * 1e-3) / 1e12,
                 "frac_of_fp32_fma_peak": step_flops / (ms_dev * 1e-3) / 1e12 / fp32_peak,
                 "gbs": step_bytes / (ms_dev * 1e-3) / 1e9, "frac_of_hbm_peak": step_bytes / (ms_dev * 1e-3) / 1e9 / bw_peak,
                 "sm_time_ms": sm_ms, "packing": sm_ms / ms_dev,
                 "note": "sm_time = sum over stages of (serial duration x CTAs / 148) x %.1f groups: the share of ms_per_step "
                         "the kernels occupy when perfectly packed; packing = sm_time / ms_per_step" % calls_per_step},
    }

    # ---- offline (whole-utterance) pass of the same clips, for context ----
    x_dev = mix.to(dev)
    inputs = {"mixture": x_dev, "dis_embed": dis}

    def offline():
        return net(inputs)["output"]
    y_off = offline()
    stream_vs_offline = float((y_timed - y_off).abs().max())
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

    lib_base = None
    if world == 1 and not args.no_library:
        lib_base = library_baseline(net, mix, dis, dev)
        y_lib = lib_base.pop("_y", None)
        if y_lib is not None:
            lib_base["ours_vs_library_maxabs"] = float((y_off - y_lib).abs().max())
            lib_base["speedup_streaming_value"] = value / lib_base["streaming"]["value"]
            if offline_info is not None:
                lib_base["speedup_offline"] = offline_info["value"] / lib_base["offline"]["value"]

    cpu = None
    if world == 1 and not args.no_cpu:
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        n_cpu = 300
        dt = cpu_streaming_sample(sd, n_cpu, warm=2)
        kind, engine = cpu_engine()
        cpu = {"value": BATCH * n_cpu / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind, "engine": engine,
               "sample": "first %d of %d chunks of the batch-32 streaming pass (%.1f s of CPU work)" % (n_cpu, T_FRAMES, dt)}
        if train_info is not None and "error" not in train_info:
            try:
                dt_t, fr_t = cpu_training_sample(sd)
                cpu["train"] = {"value": fr_t / dt_t, "unit": "training frames/s", "kind": "port", "cores": os.cpu_count() or 1,
                                "sample": "forward + backward of 1 clip x 1 s under torch autograd (%.2f s of CPU work)" % dt_t}
            except Exception as e:                                 # noqa: BLE001 - an extra figure must not cost the headline line
                cpu["train"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    calls = -(-T_FRAMES // G)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "rtf": ms_dev * 1e-3 / (BATCH * 5.0),
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "clip_seconds": 5.0, "frames_per_step": frames,
                   "chunks_per_step": T_FRAMES, "weights": "random init (seed 0) of the TFG_S architecture",
                   "l2": "256 MB buffer written between timed steps", "cuda_graph": not args.no_graph, "pdl": bool(args.pdl),
                   "pipelined": pipe is not None, "unit_ranges": pipe.ranges if pipe is not None else None,
                   "pipeline_group": G, "pipeline_depth": pipe.depth if pipe is not None else 1,
                   "protocol": "one feed() per 8 ms chunk with the state carried; the native pipe launches %d consecutive chunks "
                               "as one call (buffering %d ms) and keeps %d groups in flight" % (G, 8 * G, pipe.depth if pipe else 1),
                   "pipeline_intra_algo": pipe.intra_algo if pipe is not None else None,
                   "pipeline_inter_algo": pipe.inter_algo if pipe is not None else None,
                   "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
                   "host_enqueue_ms_per_step": 1e3 * statistics.median(enqueue_s) if enqueue_s else None,
                   "parallelism": "dp%d (utterances sharded, no collective on the data path)" % world},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(win_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4)},
        # + the gather / scatter kernel of a device-resident group (sb_pipe.cu), which the eager twin of a call does not launch
        "gpu_launches": int((launches_per_call + (2 if pipe is not None and G > 1 else 0)) * calls * args.steps),
        "launches_per_call": int(launches_per_call), "calls_per_step": calls,
        "in_order": {"value": frames / (ms_in_order * 1e-3), "unit": UNIT, "ms_per_step": ms_in_order,
                     "note": "latency mode: one stream, one chunk per launch sequence, chunk t+1 starts when chunk t has finished"},
        "clocks": clocks, "parity": parity, "roofline": roofline, "cpu_baseline": cpu, "gpu_library_baseline": lib_base,
        "offline": offline_info, "streaming_vs_offline_maxabs": stream_vs_offline, "strong": strong_info, "train": train_info,
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
    ap.add_argument("--group", type=int, default=int(os.environ.get("SB_GROUP", "64")), help="pipelined session: chunks per launch")
    ap.add_argument("--depth", type=int, default=int(os.environ.get("SB_DEPTH", "16")), help="pipelined session: groups in flight")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed output")
    ap.add_argument("--no-library", action="store_true", help="skip the stock-PyTorch-on-GPU context baseline")
    ap.add_argument("--no-strong", action="store_true", help="skip the 256-clip strong-scaling leg")
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
