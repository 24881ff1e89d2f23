"""Multi-GPU inference: utterances are independent, so a batch shards over ranks with NO collective on the data path
(SURVEY.md §8e).  One process per GPU (torchrun); ``torch.distributed`` is only used to agree on shapes and, when the
caller wants the full result on one rank, for a final gather of the separated waveforms.

The reference's multi-GPU story is single-process ``nn.DataParallel`` (src/hl_modules/distance_based_hl_module.py:34-35),
which scatters dim 0 the same way: contiguous, near-equal shards in rank order."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal shard [lo, hi) of `n_items` for `rank` (same split as torch.chunk / DataParallel.scatter)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def shard_inputs(inputs: dict, world_size: int, rank: int) -> dict:
    """Slice every batched tensor of a reference-style input dict ({'mixture', 'dis_embed', ...}) to this rank's rows."""
    n = inputs["mixture"].shape[0]
    lo, hi = shard_bounds(n, world_size, rank)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in inputs.items()}


def gather_outputs(local: torch.Tensor, n_items: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Optional epilogue: reassemble [n_items, ...] from per-rank shards (all ranks, or only `dst`).  Off the timed data
    path; uses all_gather on padded shards so ragged splits work on both NCCL and Gloo."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_items + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    if dst is not None and rank != dst:
        return None
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_bounds(n_items, world, r)
        out.append(p[: hi - lo])
    return torch.cat(out, dim=0)


def sharded_forward(net, inputs: dict, world_size: int, rank: int, **kw) -> dict:
    """Run this rank's shard through `net` (the drop-in Net).  No communication."""
    return net(shard_inputs(inputs, world_size, rank), **kw)
