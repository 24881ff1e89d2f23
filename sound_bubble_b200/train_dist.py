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


# ---------------------------------------------------------------------------------------------------------------------
# PLModule for one rank of a process-per-GPU job
# ---------------------------------------------------------------------------------------------------------------------
_MODEL_PATHS = {
    "src.models.tfgridnet_realtime_clean_dis_embd3.net.Net": "sound_bubble_b200.tfgridnet_realtime_clean_dis_embd3.net.Net",
    "src.models.tfgridnet_realtime_clean_optim.net.Net": "sound_bubble_b200.tfgridnet_realtime_clean_optim.net.Net",
}


def import_attr(path):
    """src/utils.py:10-12, with the reference's dotted paths of the separator mapped onto this package and its loss paths
    falling back to `sound_bubble_b200.losses` where the reference tree (or asteroid / auraloss) is not importable."""
    import importlib
    if not isinstance(path, str):
        return path                                   # a class / factory passed directly
    path = _MODEL_PATHS.get(path, path)
    module, attr = path.rsplit(".", 1)
    try:
        return getattr(importlib.import_module(module), attr)
    except ImportError:
        from .losses import REFERENCE_LOSSES
        if path in REFERENCE_LOSSES:
            return REFERENCE_LOSSES[path]
        raise


class TrainModule:
    """Per-rank counterpart of the reference's ``PLModule`` (src/hl_modules/distance_based_hl_module.py:22-113, 303-441):
    same constructor arguments (an experiment JSON's ``pl_module_args`` go in unchanged), same ``train / eval /
    reset_grad / training_step / validation_step / backprop / dump_state / load_state / get_current_lr`` contract, so
    the epoch loops of src/training (train_epoch: reset_grad -> training_step -> loss.backward() -> backprop) drive it
    as they drive PLModule.  Differences, all of them the DDP-isation of SURVEY.md section 8f-1:

      * no ``nn.DataParallel`` (``use_dp`` is accepted and ignored): every process owns one GPU and a shard of the batch;
      * ``reset_grad`` / ``backprop`` work on ONE flat gradient buffer: zero -> (autograd) -> one all-reduce, weighted by
        the items of the rank's shard = the gradient of the GLOBAL-batch mean loss (:321) -> clip (:433-438) -> step;
      * wandb logging, audio samples and the dataset statistics are left to the caller (rank 0); the metric bookkeeping
        (``log_metric``, ``get_avg_metric_at_epoch``) and the best-checkpoint rule of ``on_epoch_end`` are kept.
    """

    def __init__(self, model, model_params, sr, optimizer, optimizer_params, scheduler=None, scheduler_params=None,
                 loss=None, loss_params=None, metrics=(), init_ckpt=None, grad_clip=None, use_dp=False,
                 val_log_interval=10, samples_per_speaker_number=3, group=None):
        self.model = import_attr(model)(**model_params)
        self.use_dp = False
        self.sr = sr
        self.samples_per_speaker_number = samples_per_speaker_number
        self.metrics = list(metrics)
        self.metric_values, self.statistics = {}, {}
        self.monitor, self.monitor_mode = "val/loss", "min"
        self.mode = None
        self.loss_fn = import_attr(loss)(**(loss_params or {}))
        if init_ckpt is not None:
            state = torch.load(init_ckpt, map_location="cpu", weights_only=False)
            self.model.load_state_dict(state["model"] if "model" in state else
                                       {k[len("model."):]: v for k, v in state["state_dict"].items()})
        self.optim_name, self.opt_params = optimizer, dict(optimizer_params)
        self.optimizer = import_attr(optimizer)(self.model.parameters(), **optimizer_params)
        self.grad_clip = grad_clip
        self.scheduler_name, self.scheduler_params = scheduler, scheduler_params
        self.scheduler = self.init_scheduler(scheduler, scheduler_params)
        self.epoch = 0
        self.group = group
        self._reducer = None
        self._n_local = None

    # -- plumbing -----------------------------------------------------------------------------------------------------
    def init_scheduler(self, scheduler, scheduler_params):
        if scheduler is None:
            return None
        return import_attr(scheduler)(self.optimizer, **(scheduler_params or {}))

    def reducer(self) -> FlatGradReducer:
        """Created on first use, i.e. after the caller's ``hl_module.model.to(device)`` (src/train_pt.py:86)."""
        p0 = next(self.model.parameters())
        if self._reducer is None or self._reducer.flat.device != p0.device:
            self._reducer = FlatGradReducer(self.model.parameters(), group=self.group)
        return self._reducer

    def train(self):
        self.model.train()
        self.mode = "train"

    def eval(self):
        self.model.eval()
        self.mode = "val"

    def get_current_lr(self):
        for g in self.optimizer.param_groups:
            return g["lr"]

    def reset_grad(self):
        self.reducer().zero_grad()

    def backprop(self):
        return backprop(self.reducer(), self.optimizer, self.grad_clip, n_local=self._n_local)

    # -- steps (:303-330, :379-418) --------------------------------------------------------------------------------------
    def log_metric(self, name, value, batch_size=1, on_step=False, on_epoch=True, **kwargs):
        rec = self.metric_values.setdefault(self.epoch, {}).setdefault(name, dict(step=None, epoch=None))
        value = value.item() if isinstance(value, torch.Tensor) else value
        if on_step:
            rec["step"] = (rec["step"] or []) + [value]
        if on_epoch:
            rec["epoch"] = (rec["epoch"] or 0) + value * batch_size
            rec["num_elements"] = rec.get("num_elements", 0) + batch_size

    def get_avg_metric_at_epoch(self, metric, epoch=None):
        rec = self.metric_values[self.epoch if epoch is None else epoch][metric]
        return rec["epoch"] / rec["num_elements"]

    def _step(self, batch, batch_idx, step="train"):
        from . import evaluate as ev
        inputs, targets = batch
        batch_size = inputs["mixture"].shape[0]
        outputs = self.model(inputs)
        mix = inputs["mixture"][:, 0:1]
        est, gt = outputs["output"], targets["target"]
        loss = self.loss_fn(est=est, gt=gt).mean()
        self._n_local = batch_size
        with torch.no_grad():
            self.log_metric(f"{step}/loss", loss.item(), batch_size=batch_size, on_step=(step == "train"))
            n_spk = targets.get("num_target_speakers")
            fns = {"snr_i": lambda e, g, m: ev.snr(e, g) - ev.snr(m, g), "si_snr_i": lambda e, g, m: ev.si_snr(e, g) - ev.si_snr(m, g),
                   "si_sdr_i": lambda e, g, m: ev.si_sdr(e, g) - ev.si_sdr(m, g)}
            for name in self.metrics:
                if name not in fns:
                    continue
                val = fns[name](est.detach(), gt, mix).reshape(batch_size, -1).mean(dim=1)
                for i in range(batch_size):
                    if n_spk is None or int(n_spk[i]) > 0:
                        self.log_metric(f"{step}/{name}", val[i].item())
        sample = {"mixture": mix, "output": est.detach(), "target": gt, "n_tgt_speakers": n_spk}
        return loss, sample

    def training_step(self, batch, batch_idx):
        loss, _ = self._step(batch, batch_idx, step="train")
        return loss, self._n_local

    def validation_step(self, batch, batch_idx):
        with torch.no_grad():
            loss, _ = self._step(batch, batch_idx, step="val")
        return loss, self._n_local

    # -- epochs and checkpoints (:115-156, :171-204, :279-287) ------------------------------------------------------------
    def on_epoch_start(self):
        pass

    def on_epoch_end(self, best_path=None, wandb_run=None):
        last = self.get_avg_metric_at_epoch(self.monitor)
        best = all(not (last > self.get_avg_metric_at_epoch(self.monitor, e)) for e in range(len(self.metric_values) - 1))
        if best and best_path is not None:
            self.dump_state(best_path)
        if self.scheduler is not None:
            if isinstance(self.scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau):
                self.scheduler.step(last)
            else:
                self.scheduler.step()
        self.epoch += 1

    def dump_state(self, path):
        dump_state(path, self.model, self.optimizer, self.epoch, self.metric_values, self.statistics, self.scheduler)

    def load_state(self, path, map_location=None):
        rest = load_state(path, self.model, self.optimizer, self.scheduler, map_location=map_location)
        self.epoch, self.metric_values = rest["current_epoch"], rest["metric_values"]
