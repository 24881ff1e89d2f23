"""Streaming-state wire format: the nested state dict of ``Net.init_buffers`` <-> named flat tensors.

Mirrors the reference's edge tooling so that a state produced here can be fed to the reference runtimes and back:
  * names are the keys along the path joined with ``::``, visited in sorted order at every level
    (edge/flatbuf.py:8-25 ``flatten_state_buffers``), e.g. ``gridnet_bufs::buf0::h0``;
  * ``unflatten_state`` rebuilds the nested dict from (names, tensors) (edge/flatbuf.py:27-71);
  * ``save_vectors`` / ``load_vectors`` use the directory layout the ONNX tools exchange (edge/edge_utils.py:5-17,
    edge/to_onnx.py:94-136): ``input_names.txt`` (one name per line, ``mixture`` first), ``mixture.npy``, ``<name>.npy``.
``StateArena`` additionally packs one state into ONE contiguous device buffer (every tensor a 256-byte aligned view),
so that a whole streaming state moves with a single copy and a captured CUDA graph sees fixed addresses.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

DELIMITER = "::"


def flatten_state(state: dict, prefix: str = "", clone: bool = False) -> Tuple[List[str], List[torch.Tensor]]:
    """edge/flatbuf.py:10-25.  The reference clones every tensor; here that is opt-in (views by default)."""
    names: List[str] = []
    bufs: List[torch.Tensor] = []
    for k in sorted(state.keys()):
        v = state[k]
        if isinstance(v, dict):
            n, b = flatten_state(v, prefix=f"{prefix}{k}{DELIMITER}", clone=clone)
            names.extend(n)
            bufs.extend(b)
        else:
            if not torch.is_tensor(v):
                raise TypeError(f"Expected torch.Tensor, found {type(v)}")
            names.append(f"{prefix}{k}")
            bufs.append(v.clone() if clone else v)
    return names, bufs


def unflatten_state(names: Sequence[str], bufs: Sequence[torch.Tensor], clone: bool = False) -> dict:
    """edge/flatbuf.py:27-71: inverse of flatten_state (children keep the order in which their names arrive)."""
    if len(names) != len(bufs):
        raise ValueError("names and buffers differ in length")
    root: dict = {}
    for name, buf in zip(names, bufs):
        path = name.split(DELIMITER)
        node = root
        for key in path[:-1]:
            nxt = node.setdefault(key, {})
            if not isinstance(nxt, dict):
                raise ValueError(f"'{name}': '{key}' is both a tensor and a sub-dictionary")
            node = nxt
        if path[-1] in node:
            raise ValueError(f"duplicate state name '{name}'")
        node[path[-1]] = buf.clone() if clone else buf
    return root


class StateArena:
    """One streaming state as a single flat float32 buffer plus the nested dict of views into it."""

    ALIGN = 64                                              # floats (256 bytes)

    def __init__(self, template: dict, device=None):
        names, bufs = flatten_state(template)
        self.names = names
        self.shapes = [tuple(b.shape) for b in bufs]
        self.offsets: List[int] = []
        off = 0
        for b in bufs:
            if b.dtype != torch.float32:
                raise TypeError("state tensors are float32")
            self.offsets.append(off)
            off += (b.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        dev = device if device is not None else (bufs[0].device if bufs else "cpu")
        self.flat = torch.zeros(max(off, 1), dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + int(np.prod(s))].view(s) for o, s in zip(self.offsets, self.shapes)]
        self.state = unflatten_state(names, self.views)
        for v, b in zip(self.views, bufs):
            v.copy_(b)

    def load(self, state: dict):
        """Copy a reference-layout state dict (same names and shapes) into the arena."""
        names, bufs = flatten_state(state)
        if names != self.names:
            raise KeyError("state names differ: %r vs %r" % (names[:4], self.names[:4]))
        for v, b, s in zip(self.views, bufs, self.shapes):
            if tuple(b.shape) != s:
                raise ValueError("state tensor has shape %s, expected %s" % (tuple(b.shape), s))
            v.copy_(b)

    def to_host(self) -> Dict[str, np.ndarray]:
        """name -> array, one device-to-host copy for the whole state."""
        host = self.flat.detach().cpu().numpy()
        return {n: host[o:o + int(np.prod(s))].reshape(s).copy() for n, o, s in zip(self.names, self.offsets, self.shapes)}


def save_vectors(path: str, mixture: torch.Tensor, state: dict):
    """Write ``mixture`` + the flattened state in the layout edge/edge_utils.py::load_inputs reads."""
    os.makedirs(path, exist_ok=True)
    names, bufs = flatten_state(state)
    with open(os.path.join(path, "input_names.txt"), "w") as f:
        f.write("\n".join(["mixture"] + names) + "\n")
    np.save(os.path.join(path, "mixture.npy"), mixture.detach().cpu().numpy())
    for n, b in zip(names, bufs):
        np.save(os.path.join(path, f"{n}.npy"), b.detach().cpu().numpy())


def load_vectors(path: str, device="cpu") -> Tuple[torch.Tensor, dict]:
    """edge/edge_utils.py:5-17, returning torch tensors and the nested state dict."""
    with open(os.path.join(path, "input_names.txt")) as f:
        names = [x.strip() for x in f.readlines() if x.strip()]
    mixture = torch.from_numpy(np.load(os.path.join(path, "mixture.npy"))).to(device)
    names.remove("mixture")
    bufs = [torch.from_numpy(np.load(os.path.join(path, f"{n}.npy"))).to(device) for n in names]
    return mixture, unflatten_state(names, bufs)
