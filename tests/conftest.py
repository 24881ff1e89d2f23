import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
STATE_STRIDE = 7      # oracle/make_golden.py stores h0/c0/K_buf/V_buf as every 7th element


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/*.npz file: inputs/outputs of the UNMODIFIED reference (see oracle/make_golden.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.meta = json.loads(str(z["meta"]))
        self.variant = self.meta["variant"]
        self.kwargs = self.meta["kwargs"]
        self.pad = self.meta["pad"]
        if "mixture_int16" in z.files:      # read as int16/32768 like librosa.load (reference src/utils.py:137-141)
            self.mixture = torch.from_numpy(z["mixture_int16"].astype(np.float32) / 32768.0).unsqueeze(0)
        else:
            self.mixture = torch.from_numpy(z["mixture"])
        self.dis_embed = torch.from_numpy(z["dis_embed"])
        self.output = torch.from_numpy(z["output"])
        self.state = {k: torch.from_numpy(z[k]) for k in z.files if k.startswith("state::")}
        self.mixture2 = torch.from_numpy(z["mixture2"]) if "mixture2" in z.files else None
        self.output2 = torch.from_numpy(z["output2"]) if "output2" in z.files else None

    def inputs(self, device="cpu"):
        return {"mixture": self.mixture.to(device), "dis_embed": self.dis_embed.to(device)}


def flatten_state(st, prefix="state"):
    out = {}
    for k in sorted(st):
        if isinstance(st[k], dict):
            out.update(flatten_state(st[k], f"{prefix}::{k}"))
        elif k in ("h0", "c0", "K_buf", "V_buf"):
            out[f"{prefix}::{k}"] = st[k].detach().cpu().reshape(-1)[::STATE_STRIDE]
        else:
            out[f"{prefix}::{k}"] = st[k].detach().cpu()
    return out


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("grad_"))   # grad_*: train_cases.py


@pytest.fixture(scope="session")
def golden_cache():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return get
