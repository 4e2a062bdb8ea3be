"""Parity of the per-target kernels that run the LARGE configurations (BASELINE.json configs[2..4]):
ukf_finish_kernel<64> (32 769 .. 131 072 targets) and ukf_finish_kernel<128> (above), the in-line
measurement update they carry, their shared-memory covariance stores and the bulk-copied sensor tile
(clockpunch_b200/csrc/kernels_ukf.cu).  They evaluate UnscentedKalmanFilter.predictAndUpdate
(punchclock/estimation/ukf_v2.py:185-375) + Target.updateNonPhysical (common/agents.py:144-171) + one row
of both vis maps (common/visibility.py:56-116) with a different thread mapping than the four-lanes-per-
target kernel that the golden vectors of tests/test_gpu_parity.py exercise, so they get

  * an ORACLE check: a random sample of >= 256 targets (every measured one, the eccentric ones and the
    one with a non-PD covariance among them) of a >= 40 000 / >= 200 000 target catalog stepped through
    pc_step_fused, against oracle.OracleUKF at the stated gates (covariances 1e-8
    relative, est_x max(1e-5 km, 1e-9 |r|) after an update and 1e-9 km without, truth 1e-9 km, vis bits
    outside 1e-9 rad of grazing), and
  * a CROSS-KERNEL check on EVERY target: the same catalog, measurements and tasking through each of the
    other mappings (pc_set_option PC_OPT_FINISH_KERNEL); the two thread-per-target instantiations must
    agree bit for bit, the quad kernel to round-off (its sums over the sigma pairs are taken in a
    different order; the +-2e6 sigma weights turn 1 ulp of that into ~1e-11 relative of pred_p).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POS_TOL = 1e-9
COV_RTOL = 1e-8
GRAZE = 1e-9
MU = 398600.0
RE = 6378.1363


def estx_tol(x):
    return np.maximum(1e-5, 1e-9 * np.linalg.norm(np.asarray(x)[:3], axis=0))


def cov_err(a, b):
    scale = np.sqrt(np.abs(np.einsum("...ii,...jj->...ij", b, b)))
    return np.max(np.abs(a - b) / scale)


def _v_values(sens, targ):
    rs, rt = sens[:3], targ[:3]
    ns, nt = np.linalg.norm(rs, axis=0), np.linalg.norm(rt, axis=0)
    cosphi = np.clip((rt.T @ rs) / np.outer(nt, ns), -1, 1)
    return np.arccos(np.clip(RE / nt, -1, 1))[:, None] + np.arccos(np.clip(RE / ns, -1, 1))[None, :] - np.arccos(cosphi)


def _catalog(n, m, seed, n_heo=12):
    """Mixed LEO/MEO/GEO catalog with a few eccentric orbits (e up to 0.7, started near perigee: several
    DOP853 substeps per step) and one target whose covariance is not positive definite."""
    from clockpunch_b200.simulation.catalog import synth_catalog
    cat = synth_catalog(n, m, seed=seed, space_sensor_frac=0.2)
    rng = np.random.default_rng(seed + 1)
    heo = rng.choice(n, n_heo, replace=False)
    for k, i in enumerate(heo):
        rp, e = RE + 450.0 + 60.0 * k, 0.1 + 0.05 * k
        inc, nu = rng.uniform(0.1, 1.4), rng.uniform(-0.4, 0.4)        # near perigee
        p = rp * (1 + e)
        r = p / (1 + e * np.cos(nu))
        rv = np.array([r * np.cos(nu), r * np.sin(nu), 0.0])
        vv = np.sqrt(MU / p) * np.array([-np.sin(nu), e + np.cos(nu), 0.0])
        rot = np.array([[1, 0, 0], [0, np.cos(inc), -np.sin(inc)], [0, np.sin(inc), np.cos(inc)]])
        cat.targets[:, i] = np.concatenate([rot @ rv, rot @ vv])
        cat.est_x[:, i] = cat.targets[:, i] + rng.normal(size=6) * np.array([1, 1, 1, 1e-3, 1e-3, 1e-3])
    npd = int(rng.choice(np.setdiff1d(np.arange(n), heo)))
    cat.est_p[npd, 0, 1] = cat.est_p[npd, 1, 0] = 10.5            # the golden non-PD matrix (tests/golden/make_golden.py)
    return cat, heo, npd


def _engine(cat, variant=0, seed=5, num_envs=1):
    from clockpunch_b200 import _lib
    from clockpunch_b200.engine import CatalogEngine
    ctx = _lib.get_ctx()
    old = ctx.set_option(_lib.OPT_FINISH_KERNEL, variant)
    try:
        eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R,
                            seed=seed, num_envs=num_envs)
    finally:
        ctx.set_option(_lib.OPT_FINISH_KERNEL, old)
    return eng


def _step(eng, variant, **kw):
    from clockpunch_b200 import _lib
    old = eng.ctx.set_option(_lib.OPT_FINISH_KERNEL, variant)
    try:
        eng.step(**kw)
    finally:
        eng.ctx.set_option(_lib.OPT_FINISH_KERNEL, old)


OUT = ("truth_x", "est_x", "est_p", "pred_p", "nis", "flags", "num_tasked", "last_time_tasked", "vis_truth", "vis_est",
       "tr_pos", "tr_vel")


def _outputs(eng):
    return {k: eng.host(k) for k in OUT}


def _assert_same_bits(a, b, what):
    for k in OUT:
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"{what}: {k}")


def _assert_roundoff(a, b, what, steps=1):
    """Two mappings of the same equations.  First step (bit-identical inputs): everything that is not a
    sum over sigma pairs is identical, the rest agrees to round-off (1e-10 relative in the covariances,
    1e-9 km in est_x).  Later steps start from states that differ in their last bits, and the reference's
    own sigma weights (-2e6 / +1.7e5, ukf_v2.py:120-128) turn 1 ulp of a sigma point (1e-12 km) into up to
    1e-6 km of pred_x (the spread between two runs of the reference itself), so from
    the second step on the gates are the stated one-step PARITY gates: covariances 1e-8 relative per
    step, est_x max(1e-5 km, 1e-9 |r|) per step for targets that have been measured, 1e-9 km otherwise."""
    for k in ("truth_x", "num_tasked", "last_time_tasked", "vis_truth", "flags"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"{what}: {k}")
    ctol = 1e-10 if steps == 1 else COV_RTOL * steps
    assert cov_err(a["pred_p"], b["pred_p"]) <= ctol, (what, "pred_p", cov_err(a["pred_p"], b["pred_p"]))
    assert cov_err(a["est_p"], b["est_p"]) <= ctol, (what, "est_p", cov_err(a["est_p"], b["est_p"]))
    dx = np.abs(a["est_x"] - b["est_x"]).max(axis=0)
    scale = np.maximum(1.0, np.linalg.norm(b["est_x"][:3], axis=0) / 7000.0)
    if steps == 1:
        assert (dx <= 1e-9 * scale).all(), (what, "est_x", dx.max())
    else:
        measured = b["num_tasked"] > 0
        assert (dx[measured] <= estx_tol(b["est_x"][:, measured]) * steps).all(), (what, "est_x (measured)", dx[measured].max())
        assert (dx[~measured] <= POS_TOL * steps * scale[~measured]).all(), (what, "est_x (never measured)", dx[~measured].max())
    np.testing.assert_allclose(a["nis"], b["nis"], rtol=1e-7 if steps == 1 else 1e-4, atol=1e-12, equal_nan=True,
                               err_msg=f"{what}: nis")
    np.testing.assert_allclose(a["tr_pos"], b["tr_pos"], rtol=1e-6, err_msg=f"{what}: tr_pos")
    # the estimated map may differ only where the estimate moved a pair across the horizon within round-off
    diff = a["vis_est"] != b["vis_est"]
    assert diff.mean() < 1e-6, (what, "vis_est", diff.sum())


def _oracle_check(cat, sample, tasked, z, dt, got, steps_done, orc_filters, orc_truth, sens_now):
    """One more step of the sampled targets through the oracle, compared with `got` (the engine's
    outputs after the same step)."""
    from oracle.dynamics import propagate_satellite
    from oracle.visibility import calc_vis_map
    t0, t1 = dt * steps_done, dt * (steps_done + 1)
    for k, i in enumerate(sample):
        orc_truth[:, k] = propagate_satellite(orc_truth[:, k], t0, t1)
        f = orc_filters[k]
        f.predict_and_update(t1, z[:, i] if tasked[i] else None)
    s = steps_done + 1
    ex = np.stack([f.est_x for f in orc_filters], axis=1)
    ep = np.stack([f.est_p for f in orc_filters])
    pp = np.stack([f.pred_p for f in orc_filters])
    assert np.abs(got["truth_x"][:, sample] - orc_truth)[:3].max() <= POS_TOL * s
    upd = tasked[sample]
    dx = np.abs(got["est_x"][:, sample] - ex).max(axis=0)
    seen = np.array([f.last_measurement_time is not None for f in orc_filters])
    assert (dx[seen] <= estx_tol(ex[:, seen]) * s).all(), ("est_x (measured)", dx[seen].max())
    assert (dx[~seen] <= POS_TOL * s).all(), ("est_x (never measured)", dx[~seen].max())
    nonpd = np.array([f.used_nearest_pd for f in orc_filters])
    ok = ~nonpd
    assert cov_err(got["est_p"][sample][ok], ep[ok]) <= COV_RTOL * s, cov_err(got["est_p"][sample][ok], ep[ok])
    assert cov_err(got["pred_p"][sample][ok], pp[ok]) <= COV_RTOL * s, cov_err(got["pred_p"][sample][ok], pp[ok])
    if nonpd.any():       # repaired factor: LAPACK's and the kernel's shift sequences differ (tests/test_gpu_parity.py)
        assert cov_err(got["pred_p"][sample][nonpd], pp[nonpd]) <= 1e-6
    nis_o = np.array([float(f.nis) if (u and np.size(f.nis)) else np.nan for f, u in zip(orc_filters, upd)])
    np.testing.assert_allclose(got["nis"][sample][upd], nis_o[upd], rtol=1e-6)
    assert np.isnan(got["nis"][sample][~upd]).all()
    # visibility rows of the sample against the oracle restatement, outside the grazing band
    vt, ve = calc_vis_map(sens_now, orc_truth), calc_vis_map(sens_now, ex)
    Vt, Ve = _v_values(sens_now, orc_truth), _v_values(sens_now, ex)
    assert not ((got["vis_truth"][sample] != vt) & (np.abs(Vt) > GRAZE)).any()
    assert not ((got["vis_est"][sample] != ve) & (np.abs(Ve) > 1e-7)).any()
    return ex, ep


@pytest.mark.parametrize("n, kernel", [(40_000, "ukf_finish_kernel<64>"), (200_000, "ukf_finish_kernel<128>")])
def test_thread_per_target_kernels_match_oracle_and_other_mappings(n, kernel):
    from clockpunch_b200 import _lib
    from clockpunch_b200.dynamics.propagator import propagate_batch
    from oracle.ukf import OracleUKF
    m, dt = 70, 100.0
    cat, heo, npd = _catalog(n, m, seed=4100 + n % 97)
    rng = np.random.default_rng(n)
    eng = _engine(cat)                                   # size rule: n > 32768 -> one thread per target
    others = {v: _engine(cat, v) for v in (_lib.FINISH_QUAD_INLINE, _lib.FINISH_THREAD_64, _lib.FINISH_THREAD_128)}
    # sample: every measured target of both steps, the eccentric ones, the non-PD one, and random others
    tasked_steps = []
    for s in range(2):
        tk = np.zeros(n, bool)
        tk[rng.choice(n, 90, replace=False)] = True
        tk[heo[: 4 + 2 * s]] = True
        if s == 1:
            tk[npd] = True
        tasked_steps.append(tk)
    sample = np.unique(np.concatenate([np.flatnonzero(tasked_steps[0] | tasked_steps[1]), heo, [npd],
                                       rng.choice(n, 120, replace=False)]))
    assert len(sample) >= 256
    orc_filters = [OracleUKF(0.0, cat.est_x[:, i], cat.est_p[i], cat.Q, cat.R) for i in sample]
    orc_truth = cat.targets[:, sample].copy()
    for s in range(2):
        tk = tasked_steps[s]
        # measurements: the truth at the END of the step + noise, drawn on the host (injected on both sides)
        z = propagate_batch(eng.host("truth_x"), dt * s, dt * (s + 1)) + rng.normal(size=(6, n)) * np.sqrt(np.diag(cat.R))[:, None]
        eng.step(dt=dt, tasked=tk.astype(np.uint8), z=z)
        got = _outputs(eng)
        _oracle_check(cat, sample, tk, z, dt, got, s, orc_filters, orc_truth, eng.host("sens_x"))
        assert got["flags"][npd] & _lib.FLAG_NONPD_EST if s == 0 else True
        np.testing.assert_array_equal(got["num_tasked"], tasked_steps[0].astype(np.int64) + (tasked_steps[1] if s else 0))
        np.testing.assert_array_equal(got["last_time_tasked"][tk], np.float32(dt * (s + 1)))
        # every target, every other mapping
        res = {}
        for v, e2 in others.items():
            _step(e2, v, dt=dt, tasked=tk.astype(np.uint8), z=z)
            res[v] = _outputs(e2)
        _assert_same_bits(res[_lib.FINISH_THREAD_64], res[_lib.FINISH_THREAD_128], "64- vs 128-thread CTAs")
        same = _lib.FINISH_THREAD_64 if n <= 128 * 1024 else _lib.FINISH_THREAD_128
        _assert_same_bits(got, res[same], f"size rule picks {kernel}")
        _assert_roundoff(got, res[_lib.FINISH_QUAD_INLINE], f"{kernel} vs quad kernel, step {s}", steps=s + 1)
    fl = eng.host("flags")
    assert (np.delete(fl, npd) & ~np.uint32(_lib.FLAG_SUBSTEP_CLAMP) == 0).all()


def test_actions_path_of_large_catalog_equals_mask_path_and_update_warps():
    """The same large catalog stepped with an ACTIONS buffer (tasking built by the sensor kernel, in-line
    update in the thread-per-target kernel) equals the mask path bit for bit, and the four-lane kernel
    with update warps (what a <= 32 k catalog would run) to round-off -- measurements drawn on device,
    so the Philox keying (seed, step, target) is compared as well."""
    from clockpunch_b200 import _lib
    n, m, dt = 40_000, 70, 100.0
    cat, heo, npd = _catalog(n, m, seed=77)
    a, b, c = _engine(cat, seed=11), _engine(cat, seed=11), _engine(cat, _lib.FINISH_QUAD_UPDATE_WARPS, seed=11)
    rng = np.random.default_rng(1)
    for s in range(3):
        vis = a.host("vis_est")
        acts = np.full(m, n, np.int64)
        for j in range(m):
            seen = np.flatnonzero(vis[:, j])
            if len(seen):
                acts[j] = seen[rng.integers(0, len(seen))]
        acts[::7] = acts[1]                                    # several sensors on one target
        acts[3] = int(np.flatnonzero(vis[:, 3] == 0)[0])       # an invisible pick: masked out
        a.step(dt=dt, actions=acts)
        b.task_from_actions(acts)
        b.step(dt=dt, tasked=b.tasked)
        _step(c, _lib.FINISH_QUAD_UPDATE_WARPS, dt=dt, actions=acts)
        ra, rb, rc = _outputs(a), _outputs(b), _outputs(c)
        _assert_same_bits(ra, rb, f"actions vs mask, step {s}")
        _assert_roundoff(ra, rc, f"thread-per-target vs update warps, step {s}", steps=s + 1)
        assert (a.host("tasked") == 0).all()                   # the step leaves its tasking workspace zeroed
    assert 3 * m >= a.host("num_tasked").sum() > m


def test_batched_envs_large_kernel_equals_independent_engines():
    """BASELINE.json configs[2] shape: 1024 environments x 100 targets x 10 sensors = 102 400 targets in one
    catalog (thread-per-target kernel, per-environment sensor rows).  Sampled environments must equal
    INDEPENDENT single-environment engines (which run the four-lane kernel) -- bookkeeping, truth and
    visibility exactly, filter outputs to round-off -- and the whole batch must equal the four-lane
    kernel forced onto the same batch."""
    from clockpunch_b200 import _lib
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    E, Ne, Me, dt = 1024, 100, 10, 100.0
    cat = synth_catalog(E * Ne, E * Me, seed=303, space_sensor_frac=0.0)
    big = _engine(cat, num_envs=E, seed=3)
    quad = _engine(cat, _lib.FINISH_QUAD_INLINE, num_envs=E, seed=3)
    envs = [0, 1, 511, 1023]
    solo = []
    for e in envs:
        ts, ss = slice(e * Ne, (e + 1) * Ne), slice(e * Me, (e + 1) * Me)
        solo.append(CatalogEngine(cat.sensors[:, ss], cat.sensor_kinds[ss], cat.targets[:, ts], cat.est_x[:, ts],
                                  cat.est_p[ts], cat.Q, cat.R, seed=3))
    rng = np.random.default_rng(9)
    for s in range(3):
        vis = big.host("vis_est").reshape(E, Ne, Me)
        acts = np.full((E, Me), Ne, np.int64)
        for e in range(E):
            for j in range(Me):
                seen = np.flatnonzero(vis[e, :, j])
                if len(seen) and rng.random() < 0.8:
                    acts[e, j] = seen[rng.integers(0, len(seen))]
        z = cat.targets + rng.normal(size=cat.targets.shape)
        big.step(dt=dt, actions=acts.reshape(-1), z=z)
        _step(quad, _lib.FINISH_QUAD_INLINE, dt=dt, actions=acts.reshape(-1), z=z)
        rb = _outputs(big)
        _assert_roundoff(rb, _outputs(quad), f"batched envs, thread-per-target vs quad, step {s}", steps=s + 1)
        for e, eng in zip(envs, solo):
            ts = slice(e * Ne, (e + 1) * Ne)
            eng.step(dt=dt, actions=acts[e], z=z[:, ts])
            rs = _outputs(eng)
            sub = {k: (v[:, ts] if k in ("truth_x", "est_x") else v[ts]) for k, v in rb.items()}
            _assert_roundoff(sub, rs, f"env {e} of the batch vs its own engine, step {s}", steps=s + 1)
    assert big.host("num_tasked").sum() > E


def test_full_size_oracle_sample():
    """BASELINE.json configs[1] (100 sensors x 10 000 mixed targets, the bench workload): 64 random targets
    of that very catalog through the oracle, measured ones included, at the stated gates."""
    from clockpunch_b200.dynamics.propagator import propagate_batch
    from clockpunch_b200.simulation.catalog import synth_catalog
    from oracle.ukf import OracleUKF
    n, m, dt = 10_000, 100, 100.0
    cat = synth_catalog(n, m, seed=20240627)
    eng = _engine(cat, seed=9)
    rng = np.random.default_rng(2)
    tk = np.zeros(n, bool)
    tk[::97] = True
    sample = np.unique(np.concatenate([np.flatnonzero(tk)[:20], rng.choice(n, 48, replace=False)]))
    assert len(sample) >= 64
    z = propagate_batch(cat.targets, 0.0, dt) + rng.normal(size=(6, n)) * np.sqrt(np.diag(cat.R))[:, None]
    eng.step(dt=dt, tasked=tk.astype(np.uint8), z=z)
    orc_filters = [OracleUKF(0.0, cat.est_x[:, i], cat.est_p[i], cat.Q, cat.R) for i in sample]
    orc_truth = cat.targets[:, sample].copy()
    _oracle_check(cat, sample, tk, z, dt, _outputs(eng), 0, orc_filters, orc_truth, eng.host("sens_x"))


@pytest.mark.parametrize("n, m", [(10_000, 100), (40_000, 70), (140_000, 50)])
def test_step_host_zero_copy_at_bench_size_equals_device_path(n, m):
    """configs[1] size through the reference-facing call (pc_step_host, zero-copy output block: the HOST instantiation
    of the four-lane kernel stores the float32 observation and the summaries into mapped host memory while it runs),
    and catalogs that take the thread-per-target kernels <64, 6> / <128, 3> (one device-to-host copy node behind the
    kernels), against the device-resident graph replay of the same steps: the host block must be the float32 /
    integer image of the device state, bit for bit, every step."""
    import torch
    from clockpunch_b200.simulation.catalog import synth_catalog
    dt = 100.0
    cat = synth_catalog(n, m, seed=20240627)
    a, b = _engine(cat, seed=9), _engine(cat, seed=9)
    rng = np.random.default_rng(0)
    g = b.build_graph(dt)
    hb = a.host_buffers(pinned=True)
    A = n + m
    for s in range(4):
        vis = b.host("vis_est")
        acts = np.full(m, n, np.int64)
        for j in range(m):
            seen = np.flatnonzero(vis[:, j])
            if len(seen):
                acts[j] = seen[rng.integers(0, len(seen))]
        hb["actions"][:m].copy_(torch.from_numpy(acts))
        a.step_host(hb, dt)
        b.actions[:m].copy_(torch.from_numpy(acts))
        b.replay(g)
        obs = hb["obs"].numpy()
        np.testing.assert_array_equal(obs[: 6 * A].reshape(6, A)[:, m:], b.host("est_x").astype(np.float32), err_msg=f"eci_state step {s}")
        np.testing.assert_array_equal(obs[: 6 * A].reshape(6, A)[:, :m], b.host("sens_x").astype(np.float32))
        np.testing.assert_array_equal(obs[6 * A: 6 * A + 36 * n].reshape(n, 6, 6), b.host("est_p").astype(np.float32), err_msg=f"est_cov step {s}")
        np.testing.assert_array_equal(obs[6 * A + 36 * n:], (b.time - b.host("last_time_tasked").astype(np.float64)).astype(np.float32))
        np.testing.assert_array_equal(hb["vis_est"].numpy(), b.vis_est.cpu().numpy(), err_msg=f"vis_est step {s}")
        np.testing.assert_array_equal(hb["num_tasked"].numpy(), b.host("num_tasked"))
        np.testing.assert_array_equal(hb["tr_pos"].numpy(), b.host("tr_pos"))
        for k in ("truth_x", "est_x", "est_p", "pred_p", "vis_truth", "flags"):
            np.testing.assert_array_equal(a.host(k), b.host(k), err_msg=f"{k} step {s}")
    assert b.host("num_tasked").sum() > 2 * m
    np.testing.assert_array_equal(hb["last_time_tasked"].numpy(), b.host("last_time_tasked"))


def test_day_long_rollout_without_reset_stays_finite():
    """BASELINE.json configs[3] asks for 24 h at 60 s steps: 1 440 replays of the step graph with no reset in between
    (stand-in policy, measurements drawn on device; most targets are never measured, so their covariances grow all
    day).  A 50 000-target shard: every estimate and covariance stays finite, no non-finite flag, the truth orbits keep
    their energy and angular momentum, the bookkeeping adds up."""
    import torch
    from clockpunch_b200 import _lib
    from clockpunch_b200.simulation.catalog import synth_catalog, MU
    n, m, steps = 50_000, 100, 1440
    cat = synth_catalog(n, m, seed=20240628)
    eng = _engine(cat, seed=1)
    g = eng.build_graph(60.0, policy_probes=64, policy_seed=3)
    x0 = eng.host("truth_x")
    for _ in range(steps):
        eng.replay(g)
    torch.cuda.synchronize()
    assert eng.step_count == steps and eng.time == 60.0 * steps
    x1, ex, ep = eng.host("truth_x"), eng.host("est_x"), eng.host("est_p")
    assert np.isfinite(x1).all() and np.isfinite(ex).all() and np.isfinite(ep).all()
    en = lambda x: 0.5 * (x[3:] ** 2).sum(0) - MU / np.linalg.norm(x[:3], axis=0)
    hv = lambda x: np.cross(x[:3].T, x[3:].T)
    assert np.abs(en(x1) / en(x0) - 1).max() < 1e-12
    assert np.abs(hv(x1) - hv(x0)).max() / np.abs(hv(x0)).max() < 1e-12
    fl = eng.host("flags")
    assert not (fl & _lib.FLAG_NONFINITE).any()
    assert (np.diagonal(ep, axis1=1, axis2=2) > 0).all()
    nt, ltt = eng.host("num_tasked"), eng.host("last_time_tasked")
    assert 0 < nt.sum() <= m * steps and (ltt[nt > 0] > 0).all() and (ltt[nt == 0] == 0).all()
    assert (ltt <= 60.0 * steps).all()
