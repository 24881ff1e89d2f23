"""Generate tests/golden/*.npz by running the UNMODIFIED reference model.  Build-container only.

    python oracle/make_golden.py            # needs /root/reference (read-only) — absent on the GPU box

For every case it (1) instantiates the reference ``Net`` from /root/reference/src/models/... (third-party
``asteroid_filterbanks`` / ``espnet2`` replaced by the stand-ins in oracle/shims), (2) loads the deterministic
weights of oracle/weights.py with ``load_state_dict(strict=True)`` — which also pins the checkpoint key layout —,
(3) runs it on seeded inputs, and (4) stores inputs, outputs and the final streaming state.  It also asserts that
the oracle restatement reproduces the reference on each case before writing the file.
"""
import json
import os
import sys
import wave as wavmod

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shims"), REF]

from oracle.tfgridnet_oracle import OracleConfig, net_forward, streaming_forward, rms      # noqa: E402
from oracle.weights import make_state_dict, state_dict_digest, synthetic_mixture, radius_one_hot  # noqa: E402
from oracle.cases import CASES, SYN, RPI                                                   # noqa: E402


STATE_STRIDE = 7      # recurrent / attention state is stored as every 7th element of the flattened tensor


def flatten_state(st, prefix="state"):
    """Names follow /root/reference/edge/flatbuf.py:10-25 (sorted keys joined by '::')."""
    out = {}
    for k in sorted(st):
        if isinstance(st[k], dict):
            out.update(flatten_state(st[k], f"{prefix}::{k}"))
        elif k in ("h0", "c0", "K_buf", "V_buf"):
            out[f"{prefix}::{k}"] = st[k].detach().reshape(-1)[::STATE_STRIDE].numpy().copy()
        else:
            out[f"{prefix}::{k}"] = st[k].detach().numpy()
    return out


def read_wav_int16(path):
    with wavmod.open(path, "rb") as w:
        assert w.getsampwidth() == 2
        data = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
        return data.reshape(-1, w.getnchannels()).T.copy(), w.getframerate()


def main():
    from src.models.tfgridnet_realtime_clean_dis_embd3.net import Net as NetDE3
    from src.models.tfgridnet_realtime_clean_optim.net import Net as NetOPT
    from asteroid_filterbanks import STFTFB

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_grad_enabled(False)
    for name, case in CASES.items():
        variant, kw = case["variant"], case["kwargs"]
        cfg = OracleConfig.from_kwargs(variant, **kw)
        sd = make_state_dict(cfg, case.get("seed", 0))
        ref = (NetDE3 if variant == "dis_embed" else NetOPT)(**kw).eval()
        ref.load_state_dict(sd, strict=True)
        # the closed-form basis must equal what the third-party constructor builds
        fb = STFTFB(cfg.n_fft, cfg.n_fft, stride=cfg.stft_chunk_size, window_type="hann")
        assert torch.equal(fb._filters, sd["tfgridnet.enc.filterbank._filters"])

        if case.get("wav"):
            pcm, sr = read_wav_int16(os.path.join(REF, case["wav"]))
            assert sr == 24000
            pcm = pcm[:, : case["n_samples"]]
            mix = torch.from_numpy(pcm.astype(np.float32) / 32768.0).unsqueeze(0)
            extra = {"mixture_int16": pcm}
        else:
            mix = synthetic_mixture(case["batch"], kw["num_ch"], case["n_samples"], seed=case.get("input_seed", 1234))
            extra = {"mixture": mix.numpy()}
        dis = radius_one_hot(mix.shape[0]) if not case.get("radius") else torch.tensor([case["radius"]])
        inputs = {"mixture": mix, "dis_embed": dis}
        pad = case.get("pad", True)
        r = ref(dict(inputs), None, pad=pad)
        o = net_forward(sd, cfg, dict(inputs), None, pad=pad)
        err = (r["output"] - o["output"]).abs().max().item()
        assert err <= 2e-5, (name, err)
        rec = {
            "output": r["output"].numpy(),
            "dis_embed": dis.numpy(),
            "meta": np.array(json.dumps({"variant": variant, "kwargs": kw, "pad": pad, "seed": case.get("seed", 0),
                                         "weights_digest": state_dict_digest(sd),
                                         "oracle_vs_reference_maxabs": err})),
        }
        rec.update(extra)
        rec.update(flatten_state(r["next_state"]))
        if case.get("second_call"):           # continue from the returned state with another segment (streaming)
            mix2 = synthetic_mixture(mix.shape[0], kw["num_ch"], case["second_call"], seed=4321)
            r2 = ref({"mixture": mix2, "dis_embed": dis}, r["next_state"], pad=False)
            rec["mixture2"] = mix2.numpy()
            rec["output2"] = r2["output"].numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(f"{name:16s} out {tuple(r['output'].shape)} rms {rms(r['output']):.4f} oracle-vs-ref {err:.2e}")


if __name__ == "__main__":
    main()
