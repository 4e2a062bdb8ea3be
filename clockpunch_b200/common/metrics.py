"""Tasking metrics (mirror of punchclock/common/metrics.py:21-291).

`TaskingMetricTracker` keeps the reference's attributes, `reset()` and
`update(actions, vis_mask_est, vis_mask_truth)` signature and return tuple.  The reference builds
two dense (N, M) action arrays per call and multiplies them with the masks
(`_calcNonVisTaskings`, metrics.py:252-291): O(N M).  An action vector names at most M pairs,
so everything here is O(M): the mask entry of each (action, sensor) pair is looked up directly,
either in a dense (N, M) array (reference signature) or in the bit-packed (N, ceil(M/32)) uint32
maps the CUDA step produces (`update_packed`), which is how the env calls it at catalog sizes.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
from numpy import mean, ndarray, stack, sum, trace, zeros


def meanVarUncertainty(covs: list[ndarray], pos_vel: str = "position") -> float:
    """mean_i tr(cov_i[block]) over (6, 6) matrices or (6,) diagonals (metrics.py:21-90)."""
    if type(covs) != list:  # noqa: E721  (reference behaviour)
        raise TypeError("`covs` must be a list of arrays.")
    if covs[0].ndim == 2:
        sl = slice(0, 3) if pos_vel == "position" else slice(3, 6)
        covs_traces = trace(stack([cov[sl, sl] for cov in covs], axis=0), axis1=1, axis2=2)
    elif covs[0].ndim == 1:
        covs_traces = sum([cov[:3] if pos_vel == "position" else cov[3:] for cov in covs], axis=1)
    return mean(covs_traces)


class TaskingMetricTracker:
    """Persistent tasking metrics (metrics.py:93-250)."""

    def __init__(self, sensor_ids: list, target_ids: list):
        self.sensor_ids = sensor_ids
        self.target_ids = target_ids
        self.num_sensors = len(sensor_ids)
        self.num_targets = len(target_ids)
        self.reset()

    def reset(self):
        self._target_set = set()
        self.targets_tasked = []
        self.unique_tasks = 0
        self.non_vis_tasked_est = 0
        self.non_vis_tasked_truth = 0
        self.non_vis_by_sensor_est = zeros([self.num_sensors], dtype=int)
        self.non_vis_by_sensor_truth = zeros([self.num_sensors], dtype=int)
        self.multiple_taskings = 0

    # -- O(M) core -------------------------------------------------------------------------
    def _update(self, actions: ndarray, lookup_est, lookup_truth):
        actions = np.asarray(actions)
        assert len(actions) == self.num_sensors
        sens = np.nonzero(actions < self.num_targets)[0]
        targ = actions[sens].astype(np.int64)
        new = {self.target_ids[i] for i in targ}.difference(self._target_set)
        self.unique_tasks += len(new)
        self._target_set = self._target_set.union(new)
        self.targets_tasked = list(self._target_set)
        for lookup, tot, by in ((lookup_est, "non_vis_tasked_est", self.non_vis_by_sensor_est),
                                (lookup_truth, "non_vis_tasked_truth", self.non_vis_by_sensor_truth)):
            non_vis = lookup(targ, sens) == 0            # tasked pair whose mask entry is 0
            setattr(self, tot, getattr(self, tot) + int(np.count_nonzero(non_vis)))
            by[sens[non_vis]] += 1
        _, counts = np.unique(targ, return_counts=True)
        self.multiple_taskings += int(np.count_nonzero(counts > 1))
        return (self.unique_tasks, self.targets_tasked, self.non_vis_tasked_est, self.non_vis_tasked_truth,
                self.multiple_taskings, self.non_vis_by_sensor_est, self.non_vis_by_sensor_truth)

    def _calcNonVisTaskings(self, vis_mask: ndarray, actions: ndarray) -> Tuple:
        """(number of taskings of a non-visible target, the same per sensor (M,)) for one (N, M) mask and one
        (M,) action vector (metrics.py:252-291) -- the O(M) form: a tasking is non-visible iff the mask entry
        of its (target, sensor) pair is not 1."""
        actions, vis_mask = np.asarray(actions), np.asarray(vis_mask)
        assert len(actions) == self.num_sensors
        assert vis_mask.shape == (self.num_targets, self.num_sensors)
        sens = np.nonzero(actions < self.num_targets)[0]
        # the reference forms (actions - actions * mask) and counts its ones
        miss = (1 - vis_mask[actions[sens].astype(np.int64), sens]) == 1
        by_sensor = zeros([self.num_sensors], dtype=int)
        by_sensor[sens] = (1 - vis_mask[actions[sens].astype(np.int64), sens]).astype(int)
        return int(np.count_nonzero(miss)), by_sensor

    def update(self, actions: ndarray, vis_mask_est: ndarray, vis_mask_truth: ndarray) -> Tuple:
        """Reference signature: dense (N, M) 0/1 masks (metrics.py:161-250)."""
        for m in (vis_mask_est, vis_mask_truth):
            assert np.shape(m) == (self.num_targets, self.num_sensors)
        est, truth = np.asarray(vis_mask_est), np.asarray(vis_mask_truth)
        return self._update(actions, lambda t, s: est[t, s], lambda t, s: truth[t, s])

    def update_packed(self, actions: ndarray, words_est: ndarray, words_truth: ndarray) -> Tuple:
        """Same update from bit-packed (N, ceil(M/32)) uint32 maps (bit j % 32 of word j / 32)."""
        we, wt = np.asarray(words_est).view(np.uint32), np.asarray(words_truth).view(np.uint32)

        def look(words):
            return lambda t, s: (words[t, s >> 5] >> (s & 31).astype(np.uint32)) & np.uint32(1)
        return self._update(actions, look(we), look(wt))
