"""The one helper of punchclock/common/utilities.py that sits on the step path."""
from __future__ import annotations

import numpy as np
from numpy import ndarray


def actionSpace2Array(actions: ndarray, num_sensors: int, num_targets: int) -> ndarray:
    """(M,) MultiDiscrete actions valued 0..N (N = inaction) -> (N + 1, M) one-hot int array
    (utilities.py:147-172).  The CUDA step never builds this array -- the tasking kernel reads the
    M actions directly (pc_step_fused / pc_task_from_actions) -- it exists for callers that want
    the reference's dense form."""
    actions = np.asarray(actions)
    out = np.zeros([num_targets + 1, num_sensors], dtype=int)
    out[actions.astype(int), np.arange(len(actions))] = 1
    return out
