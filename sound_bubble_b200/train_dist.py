"""Data-parallel training plumbing: process-per-GPU replacement for the reference's single-process
``nn.DataParallel`` (src/hl_modules/distance_based_hl_module.py:34-35) around ANY autograd module.

  * ``FlatGradReducer``: every parameter's ``.grad`` is a view into ONE flat fp32 buffer (2.0 MB for the TFG_S model), so
    the gradient exchange of a step is a single ``all_reduce`` (NCCL over NVLink on a GPU box, Gloo in the CPU tests),
    followed by the division by the world size: the reference's loss is the mean over the GLOBAL batch (:321), so
    per-rank mean-loss gradients are averaged, not summed.  With ragged shards (a last partial batch, fewer items than
    ranks) pass each rank's item count: the gradients are then weighted by it and the count travels in the same
    collective (one extra float), which reproduces the global-batch mean exactly; an empty shard contributes zero.
  * ``clip_grad_norm_`` on the reduced flat buffer: the reference clips the already-reduced DataParallel gradients
    (:433-441), so the order is reduce -> clip -> optimizer step.
  * ``dump_state / load_state``: the checkpoint layout of PLModule (:115-156) with the model saved WITHOUT a wrapper
    prefix, so run directories stay interchangeable with the reference's.
It serves any autograd module: the drop-in ``Net`` in train() mode (whose backward pass is the hand-written kernels of
csrc/sb_train.cu behind training.SeparatorFunction; ``tools/train_ddp.py`` runs that over NCCL) as well as the reference
``Net`` (``tests/test_train_dist.py``, Gloo).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


class FlatGradReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self._buf = torch.zeros(self.numel + 1, dtype=dt, device=dev)      # + one slot for the item count
        self.flat = self._buf[:self.numel]
        off = 0
        for p in self.params:
            if p.dtype != dt or p.device != dev:
                raise TypeError("all parameters must be float32 on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)      # autograd accumulates into this view in place
            off += p.numel()

    def zero_grad(self):
        """Replaces optimizer.zero_grad(): keeps the views (set_to_none would detach them from the flat buffer)."""
        self.flat.zero_()

    def check_views(self):
        off = 0
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + 4 * off:
                raise RuntimeError("a parameter's .grad no longer aliases the flat buffer (zero_grad(set_to_none=True)?)")
            off += p.numel()

    def all_reduce_mean(self, n_local: Optional[int] = None):
        """ONE collective for the whole model, then the mean over ranks (equal shards), or - with `n_local` = the number of
        items this rank's mean loss was taken over - the item-weighted mean, i.e. the gradient of the global-batch mean
        loss whatever the shard sizes.  A rank with n_local == 0 contributes nothing (its local loss is NaN: 0 / 0)."""
        self.check_views()
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.group)
            if world > 1:
                if n_local is None:
                    dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                    self.flat.div_(world)
                    return
                if n_local > 0:
                    self.flat.mul_(float(n_local))
                else:
                    self.flat.zero_()
                self._buf[self.numel] = float(n_local)
                dist.all_reduce(self._buf, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(self._buf[self.numel].clamp_min(1.0))

    def clip_grad_norm_(self, max_norm: Optional[float]) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ semantics (2-norm, coefficient clamped to 1) on the flat buffer."""
        total = torch.linalg.vector_norm(self.flat, 2)
        if max_norm is not None:
            coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
            self.flat.mul_(coef)
        return total


def backprop(reducer: FlatGradReducer, optimizer, grad_clip: Optional[float] = None,
             n_local: Optional[int] = None) -> torch.Tensor:
    """PLModule.backprop (:433-441) for one rank of a data-parallel job: reduce -> clip -> step.  `n_local`: see
    FlatGradReducer.all_reduce_mean (needed when the ranks' shards differ in size)."""
    reducer.all_reduce_mean(n_local)
    norm = reducer.clip_grad_norm_(grad_clip)
    optimizer.step()
    return norm


def dump_state(path: str, model, optimizer, epoch: int, metric_values=None, statistics=None, scheduler=None):
    """PLModule.dump_state (:141-156); `model` may be wrapped (DataParallel / DDP): the bare module is saved."""
    bare = model.module if hasattr(model, "module") else model
    state = dict(model=bare.state_dict(), optimizer=optimizer.state_dict(), current_epoch=epoch,
                 metric_values=metric_values if metric_values is not None else {},
                 statistics=statistics if statistics is not None else {})
    if scheduler is not None:
        state["scheduler"] = scheduler.state_dict()
    torch.save(state, path)


def load_state(path: str, model, optimizer=None, scheduler=None, map_location=None) -> dict:
    """PLModule.load_state (:115-139): strict model load, then optimizer / scheduler; returns the rest of the record."""
    state = torch.load(path, map_location=map_location, weights_only=False)
    bare = model.module if hasattr(model, "module") else model
    bare.load_state_dict(state["model"])
    if optimizer is not None:
        optimizer.load_state_dict(state["optimizer"])
    if scheduler is not None and "scheduler" in state:
        scheduler.load_state_dict(state["scheduler"])
    return {"current_epoch": state["current_epoch"], "metric_values": state["metric_values"],
            "statistics": state.get("statistics", {})}
