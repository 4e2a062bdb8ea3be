"""Call-sequence regressions of CatalogEngine / pc_step_host (round-1 advisor findings):
  * pc_step_host runs on the library's own stream when torch's current stream is the legacy default one; work the
    engine enqueued on torch's stream before (reset, eager steps, sigma generation) must be ordered before it;
  * a tasking mask left in the workspace by step(tasked=mask) must not leak into steps that build the mask from
    actions (their update warps own exactly the targets pc_step_args.task_of names);
  * several engines share one library context (deep-copied environments do): alternating step_host calls must
    neither disturb each other's results nor re-capture the step graph every call."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("truth_x", "est_x", "est_p", "pred_p", "vis_est", "vis_truth", "num_tasked", "last_time_tasked", "flags")


def _engine(n=600, m=24, seed=3, eseed=7):
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    cat = synth_catalog(n, m, seed=seed, space_sensor_frac=0.25)
    return cat, CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=eseed)


def _state(eng):
    return {k: eng.host(k) for k in KEYS}


def _actions(eng, rng):
    vis = eng.host("vis_est")
    acts = np.full(eng.M, eng.N, np.int64)
    for j in range(eng.M):
        seen = np.flatnonzero(vis[:, j])
        if len(seen):
            acts[j] = seen[rng.integers(0, len(seen))]
    return acts


def test_step_host_is_ordered_behind_work_on_torchs_stream():
    import torch
    cat, a = _engine()
    _, b = _engine()
    hba, hbb = a.host_buffers(), b.host_buffers()
    rng = np.random.default_rng(1)
    for rep in range(4):
        acts = [_actions(a, rng) for _ in range(3)]
        for eng, hb, careful in ((a, hba, False), (b, hbb, True)):
            # an eager masked step and a reset (both on torch's stream), then step_host straight away
            mask = np.zeros(eng.N, np.uint8)
            mask[:: 7 + rep] = 1
            eng.step(dt=100.0, tasked=mask, z=cat.targets)
            if careful:
                torch.cuda.synchronize()
            eng.reset(rewind_rng=True)
            if careful:
                torch.cuda.synchronize()
            for k in range(3):
                hb["actions"][: eng.M].copy_(torch.from_numpy(acts[k]))
                eng.step_host(hb, 100.0)
                if careful:
                    torch.cuda.synchronize()
        sa, sb = _state(a), _state(b)
        for k in KEYS:
            np.testing.assert_array_equal(sa[k], sb[k], err_msg=f"{k} (repetition {rep})")
        np.testing.assert_array_equal(hba["obs"].numpy(), hbb["obs"].numpy())
    assert a.host("num_tasked").sum() > 0


def test_leftover_tasking_mask_does_not_leak_into_action_steps():
    import torch
    cat, a = _engine(seed=5)
    _, b = _engine(seed=5)
    rng = np.random.default_rng(2)
    mask = (rng.random(a.N) < 0.3).astype(np.uint8)
    z = cat.targets + rng.normal(size=cat.targets.shape)
    for eng in (a, b):
        eng.step(dt=100.0, tasked=mask, z=z)             # leaves `mask` in the engine's tasking workspace
    acts = _actions(a, rng)
    hb = a.host_buffers()
    hb["actions"][: a.M].copy_(torch.from_numpy(acts))
    a.step_host(hb, 100.0)                               # same dt as the masked step: must start from a zero mask
    b.tasked.zero_()
    b.step(dt=100.0, actions=acts)                       # reference: explicitly clean workspace
    sa, sb = _state(a), _state(b)
    for k in KEYS:
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)
    assert sa["num_tasked"].sum() <= mask.sum() + a.M     # only the sensors' picks were measured in the second step
    g = a.build_graph(100.0)
    a.step(dt=100.0, tasked=mask, z=z)
    a.actions[: a.M].copy_(torch.from_numpy(acts))
    a.replay(g)                                          # the same through a replayed graph
    b.step(dt=100.0, tasked=mask, z=z)
    b.tasked.zero_()
    b.step(dt=100.0, actions=acts)
    sa, sb = _state(a), _state(b)
    for k in ("truth_x", "vis_truth", "num_tasked", "last_time_tasked", "vis_est"):
        np.testing.assert_array_equal(sa[k], sb[k], err_msg=k)


def test_engines_sharing_a_context_keep_their_own_step_graphs():
    import torch
    from copy import deepcopy
    cat, a = _engine(n=900, seed=8)
    b = deepcopy(a)                                      # what the reference's wrappers do with the env every step
    _, c = _engine(n=900, seed=8)                        # never interleaved: the expected results
    assert a.ctx is b.ctx
    hba, hbb, hbc = a.host_buffers(), b.host_buffers(), c.host_buffers()
    rng = np.random.default_rng(3)
    acts = [_actions(a, rng) for _ in range(6)]
    for k in range(6):
        for eng, hb in ((a, hba), (b, hbb), (c, hbc)):
            hb["actions"][: eng.M].copy_(torch.from_numpy(acts[k]))
        a.step_host(hba, 100.0)
        b.step_host(hbb, 100.0)                          # alternating on the shared context
        c.step_host(hbc, 100.0)
    for k in KEYS:
        np.testing.assert_array_equal(a.host(k), c.host(k), err_msg=f"a vs c: {k}")
        np.testing.assert_array_equal(b.host(k), c.host(k), err_msg=f"b vs c: {k}")
    np.testing.assert_array_equal(hba["obs"].numpy(), hbc["obs"].numpy())
    # the per-engine graphs stay cached: alternating calls cost about what back-to-back calls of one engine cost
    def per_call(fn, n=60):
        for _ in range(5):
            fn()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t) / n
    solo = per_call(lambda: a.step_host(hba, 100.0))
    both = per_call(lambda: (a.step_host(hba, 100.0), b.step_host(hbb, 100.0))) / 2
    assert both < 3 * solo + 50e-6, (solo, both)          # a re-capture per call costs milliseconds
