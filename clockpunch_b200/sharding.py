"""Target sharding across the GPUs of one box.

Every target's truth propagation, UKF and vis-map row are independent of every other target
(punchclock/environment/env.py:493-506, common/visibility.py:102-106), so a catalog is split
into contiguous target ranges, one process per GPU; the sensor network is replicated.  The only
per-step exchange is what the scheduler reads from every shard -- the bit-packed estimated vis
map and four per-target summaries -- gathered with one `all_gather_into_tensor` per buffer over
NCCL (NVLink/NVSwitch).  With the `gloo` backend the same code runs on CPU tensors, which is how
tests/test_sharding.py covers it at world_size 2 without a GPU.

`torch` / `torch.distributed` are used for buffers and the collective only.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class TargetPartition:
    """Contiguous split of `n_total` targets over `world` ranks; the first `n_total % world`
    ranks own one extra target.  `cap` (the largest shard) is the padded per-rank row count
    of every gathered buffer, so the collective needs no variable-size variant."""
    n_total: int
    world: int

    def __post_init__(self):
        if self.world < 1 or self.n_total < 0:
            raise ValueError("need world >= 1 and n_total >= 0")

    @property
    def cap(self) -> int:
        return -(-self.n_total // self.world) if self.n_total else 0

    def count(self, rank: int) -> int:
        base, rem = divmod(self.n_total, self.world)
        return base + (1 if rank < rem else 0)

    def start(self, rank: int) -> int:
        base, rem = divmod(self.n_total, self.world)
        return rank * base + min(rank, rem)

    def bounds(self, rank: int) -> tuple[int, int]:
        s = self.start(rank)
        return s, s + self.count(rank)

    def owner(self, target: int) -> int:
        base, rem = divmod(self.n_total, self.world)
        split = rem * (base + 1)
        if target < split:
            return target // (base + 1)
        return rem + (target - split) // base if base else self.world - 1

    def localize_actions(self, actions: np.ndarray, rank: int) -> np.ndarray:
        """Global MultiDiscrete actions (target index in [0, n_total], n_total = inaction,
        common/utilities.py:147-172) -> this rank's local actions (index in [0, count],
        count = inaction): a sensor that tasks another rank's target is idle here."""
        a = np.asarray(actions, dtype=np.int64)
        lo, hi = self.bounds(rank)
        return np.where((a >= lo) & (a < hi), a - lo, hi - lo)

    def unpad(self, gathered, trailing_shape=()):
        """(world * cap, ...) gathered buffer -> (n_total, ...) in global target order."""
        cap = self.cap
        parts = [gathered[r * cap: r * cap + self.count(r)] for r in range(self.world)]
        import torch
        return torch.cat(parts, dim=0) if parts else gathered[:0]


SUMMARY_FIELDS = ("vis_est", "tr_pos", "tr_vel", "last_time_tasked", "num_tasked")


class StepGather:
    """Preallocated send/receive buffers for the per-step all-gather of one shard's
    scheduler summaries.  `local` maps field -> the shard's own (count, ...) tensor (the
    buffers the kernels write); `gather()` returns field -> (n_total, ...) global tensors."""

    def __init__(self, part: TargetPartition, rank: int, local: dict, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.part, self.rank, self.group = part, rank, group
        self.local = {k: local[k] for k in SUMMARY_FIELDS if k in local}
        cap, n = part.cap, part.count(rank)
        self.send, self.recv = {}, {}
        for k, t in self.local.items():
            if t.shape[0] != n:
                raise ValueError(f"{k}: expected {n} rows on rank {rank}, got {t.shape[0]}")
            tail = tuple(t.shape[1:])
            # equal shards (the bench's weak-scaling layout): send straight from the kernel's buffer
            self.send[k] = t if n == cap else torch.zeros((cap,) + tail, dtype=t.dtype, device=t.device)
            self.recv[k] = torch.empty((part.world * cap,) + tail, dtype=t.dtype, device=t.device)

    @property
    def bytes_per_step(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.recv.values())

    def gather(self, unpad: bool = True) -> dict:
        n = self.part.count(self.rank)
        for k, t in self.local.items():
            if self.send[k] is not t:
                self.send[k][:n].copy_(t)
            if self.part.world == 1:
                self.recv[k].copy_(self.send[k])
            else:
                self.dist.all_gather_into_tensor(self.recv[k], self.send[k], group=self.group)
        if not unpad or self.part.cap * self.part.world == self.part.n_total:
            return dict(self.recv)
        return {k: self.part.unpad(v) for k, v in self.recv.items()}


class PackedGather:
    """One all-gather per step for equal shards: every rank contributes its engine's `summary`
    buffer (num_tasked | packed estimated vis map | tr_pos | tr_vel | last_time_tasked, one
    contiguous uint8 tensor written in place by the step kernels) and receives (world, nbytes).
    `fields(r)` gives typed views of rank r's block.  The collective can be captured into the
    step's CUDA graph (NCCL supports stream capture), which removes its host launch cost."""

    def __init__(self, world: int, rank: int, summary, n_local: int, wpr: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world, self.rank, self.group = world, rank, group
        self.send, self.n_local, self.wpr = summary, n_local, wpr
        self.recv = torch.empty((world, summary.numel()), dtype=summary.dtype, device=summary.device)

    @property
    def bytes_per_step(self) -> int:
        return self.recv.numel() * self.recv.element_size()

    def gather(self):
        if self.world == 1:
            self.recv[0].copy_(self.send)
        else:
            self.dist.all_gather_into_tensor(self.recv.view(-1), self.send, group=self.group)
        return self.recv

    def fields(self, r: int) -> dict:
        from .engine import CatalogEngine
        return CatalogEngine.summary_views(self.recv[r], self.n_local, self.wpr)


def shard_catalog(cat, part: TargetPartition, rank: int):
    """This rank's slice of a synthetic catalog (sensors replicated)."""
    lo, hi = part.bounds(rank)
    return dict(sensors=cat.sensors, sensor_kinds=cat.sensor_kinds, targets=cat.targets[:, lo:hi],
                est_x=cat.est_x[:, lo:hi], est_p=cat.est_p[lo:hi], Q=cat.Q, R=cat.R)
