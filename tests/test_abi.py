"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/punch_b200.h declares; host-side argument logic; no CPU fallback."""
import os
import re

import numpy as np
import pytest

import clockpunch_b200 as cpb
from clockpunch_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "punch_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(pc_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in punch_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.pc_version() == 100


def test_struct_layouts_match_header():
    """ctypes mirrors vs the C structs, measured by compiling the header with the host compiler."""
    import ctypes as C
    import subprocess
    import tempfile
    src = ('#include "%s"\n#include <cstdio>\n#include <cstddef>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu",'
           'sizeof(pc_step_args), sizeof(pc_ukf_params), offsetof(pc_step_args, rng_seed),'
           'offsetof(pc_step_args, words_per_row), offsetof(pc_step_args, tr_vel), offsetof(pc_ukf_params, R),'
           'sizeof(pc_step_io), offsetof(pc_step_io, zero_copy), offsetof(pc_step_args, task_of),'
           'offsetof(pc_step_args, obs_cov_host));}'
           % os.path.join(ROOT, "include", "punch_b200.h"))
    with tempfile.TemporaryDirectory() as d:
        cpp, exe = os.path.join(d, "a.cpp"), os.path.join(d, "a.out")
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", cpp, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    S, U, IO = _lib.StepArgs, _lib.UkfParams, _lib.StepIO
    assert got == [C.sizeof(S), C.sizeof(U), S.rng_seed.offset, S.words_per_row.offset, S.tr_vel.offset,
                   U.R.offset, C.sizeof(IO), IO.zero_copy.offset, S.task_of.offset, S.obs_cov_host.offset]


def test_reference_weights_bit_exact(golden):
    g = golden("ukf")
    gamma, wm0, wi, wc0, eps = _lib.reference_ukf_weights()
    assert gamma == float(g["w_gamma"])
    assert wm0 == g["w_mean"][0] and wi == g["w_mean"][1] and wc0 == g["w_cvr_diag"][0]
    assert abs(eps) < 1e-9 and eps != 0.0
    p = _lib.make_ukf_params(np.eye(6), 4 * np.eye(6))
    assert p.Lr[0] == 2.0 and p.Lr[1] == 0.0 and p.Q[35] == 1.0


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    """Without a GPU every physics entry point must raise, never compute on the host."""
    from clockpunch_b200.common.visibility import calcVisMap
    from clockpunch_b200.dynamics.dynamics_classes import SatDynamicsModel
    with pytest.raises(_lib.PunchError):
        SatDynamicsModel().propagate(np.array([7000.0, 0, 0, 0, 7.5, 0]), 0, 10)
    with pytest.raises(_lib.PunchError):
        calcVisMap(np.ones((6, 2)) * 7000, np.ones((6, 3)) * 8000)


def test_argument_errors_before_device():
    from clockpunch_b200.common.transforms import coe2eci, ecef2eci
    from clockpunch_b200.common.visibility import calcVisMap
    from clockpunch_b200.dynamics.propagator import simplePropagate
    from clockpunch_b200.dynamics.dynamics import satelliteDynamics
    with pytest.raises(ValueError):
        simplePropagate(satelliteDynamics, np.zeros(6), 1.0, 1.0)
    with pytest.raises(NotImplementedError):
        simplePropagate(lambda t, x: x, np.zeros(6), 0.0, 1.0)
    with pytest.raises(ValueError):
        calcVisMap(np.zeros((5, 2)), np.zeros((6, 2)))
    with pytest.raises(ValueError):
        ecef2eci(np.zeros((5, 3)))
    with pytest.raises(ValueError):
        coe2eci(6000.0, 0, 0, 0, 0, 0)


def test_catalog_synthesis_matches_oracle_transforms():
    from clockpunch_b200.simulation.catalog import coe_to_eci, lla_to_eci, synth_catalog
    from oracle import dynamics as od
    rng = np.random.default_rng(1)
    coe = np.stack([rng.uniform(7000, 42000, 8), rng.uniform(0, 0.7, 8), rng.uniform(0, np.pi, 8),
                    *rng.uniform(0, 2 * np.pi, (3, 8))])
    got = coe_to_eci(coe)
    for i in range(8):
        np.testing.assert_allclose(got[:, i], od.coe2eci(*coe[:, i]), rtol=1e-14, atol=1e-11)
    lla = np.stack([rng.uniform(-1.5, 1.5, 5), rng.uniform(-3, 3, 5), np.full(5, 1e-6)])
    got = lla_to_eci(lla, 50.0)
    for i in range(5):
        np.testing.assert_allclose(got[:, i], od.ecef2eci(od.lla2ecef(lla[:, i]), 50.0), rtol=1e-14, atol=1e-11)
    cat = synth_catalog(100, 7, seed=3)
    assert cat.targets.shape == (6, 100) and cat.sensors.shape == (6, 7) and cat.est_p.shape == (100, 6, 6)
    r = np.linalg.norm(cat.targets[:3], axis=0)
    assert r.min() > 6378.1363 + 300 and set(np.unique(cat.regime)) <= {0, 1, 2}
    cat2 = synth_catalog(100, 7, seed=3)
    np.testing.assert_array_equal(cat.targets, cat2.targets)


def test_install_as_punchclock():
    cpb.install_as_punchclock()
    import punchclock.dynamics.dynamics_classes as dc
    from punchclock.estimation.ez_ukf import ezUKF  # noqa: F401
    assert dc.SatDynamicsModel.__module__.startswith("clockpunch_b200")


def test_host_policy_matches_restatement():
    """pc_policy_random_visible_host (no device needed): first visible of `probes` hashed probes per sensor,
    else inaction -- against a plain restatement of the same counter-based hash."""
    import ctypes as C
    import numpy as np
    from clockpunch_b200 import _lib
    lib = _lib.load_library()
    M64 = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M64
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
        return x ^ (x >> 31)

    rng = np.random.default_rng(4)
    for N, M, probes, dens in ((50, 7, 8, 0.2), (1000, 100, 64, 0.05), (33, 40, 4, 0.0), (9, 3, 16, 1.0)):
        wpr = (M + 31) // 32
        dense = rng.random((N, M)) < dens
        packed = np.zeros((N, wpr), np.uint32)
        for j in range(M):
            packed[:, j >> 5] |= dense[:, j].astype(np.uint32) << np.uint32(j & 31)
        out = np.full(M, -7, np.int64)
        seed, step = 20240627, 11
        rc = lib.pc_policy_random_visible_host(packed.ctypes.data_as(C.c_void_p), N, M, wpr, seed, step, probes,
                                               out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        for j in range(M):
            base = sm(seed ^ sm((step * 0x100000001B3 + j) & M64))
            want = N
            for k in range(probes):
                i = sm((base + k) & M64) % N
                if dense[i, j]:
                    want = i
                    break
            assert out[j] == want, (N, M, j)
    assert lib.pc_policy_random_visible_host(None, 10, 3, 5, 0, 0, 4, None) == -1      # wrong words_per_row


def test_host_policy_of_a_batch_of_environments_is_the_same_on_one_core_and_on_many():
    """pc_policy_random_visible_envs_host splits the sensors of a large batch over the cores the process may use
    (10 240 sensors at BASELINE.json configs[2]); every sensor's pick depends on its own counters only, so the
    threaded result must equal the one-core result, and both the restatement of the hash for sampled sensors."""
    import os
    import numpy as np
    from clockpunch_b200 import _lib
    lib = _lib.load_library()
    M64 = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M64
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
        return x ^ (x >> 31)

    E, Ne, Me, probes, seed, step = 700, 100, 10, 32, 5, 9
    rng = np.random.default_rng(1)
    dense = rng.random((E * Ne, Me)) < 0.06
    packed = np.zeros((E * Ne, 1), np.uint32)
    for j in range(Me):
        packed[:, 0] |= dense[:, j].astype(np.uint32) << np.uint32(j)
    many = np.full(E * Me, -7, np.int64)
    assert lib.pc_policy_random_visible_envs_host(packed.ctypes.data, E, Ne, Me, 1, seed, step, probes, many.ctypes.data) == 0
    allowed = os.sched_getaffinity(0)
    one = np.full(E * Me, -7, np.int64)
    try:
        os.sched_setaffinity(0, {min(allowed)})
        assert lib.pc_policy_random_visible_envs_host(packed.ctypes.data, E, Ne, Me, 1, seed, step, probes, one.ctypes.data) == 0
    finally:
        os.sched_setaffinity(0, allowed)
    np.testing.assert_array_equal(many, one)
    assert 0.3 < (many < Ne).mean() < 1.0
    for j in list(range(0, E * Me, 397)) + [E * Me - 1]:
        e, jl = divmod(j, Me)
        base = sm(seed ^ sm((step * 0x100000001B3 + j) & M64))
        want = Ne
        for k in range(probes):
            i = sm((base + k) & M64) % Ne
            if dense[e * Ne + i, jl]:
                want = i
                break
        assert many[j] == want, j


def test_unpack_vis_host_matches_numpy_on_ragged_shapes():
    """pc_unpack_vis_host (no device needed): (N, ceil(M/32)) packed words -> the (N, M) int64 map of the reference's
    observation, for sensor counts around the word boundaries, empty maps, and a map large enough to be split over
    threads; a words_per_row that does not fit M is a bad argument."""
    import numpy as np
    from clockpunch_b200 import _lib
    from clockpunch_b200.common.visibility import unpack_vis_words
    lib = _lib.load_library()
    rng = np.random.default_rng(3)
    for N, M in ((0, 5), (3, 1), (7, 31), (7, 32), (7, 33), (5, 64), (9, 100), (120_000, 10), (4000, 300)):
        wpr = (M + 31) // 32
        dense = rng.random((N, M)) < 0.4
        words = np.zeros((N, wpr), np.uint32)
        for j in range(M):
            words[:, j >> 5] |= dense[:, j].astype(np.uint32) << np.uint32(j & 31)
        words[:, -1:] |= np.uint32(0) if M % 32 == 0 else np.uint32(1) << np.uint32(31)   # a stray padding bit is ignored
        if M % 32 == 31:
            words[:, -1:] &= ~(np.uint32(1) << np.uint32(31))
        out = unpack_vis_words(words, M)
        assert out.dtype == np.int64 and out.shape == (N, M)
        np.testing.assert_array_equal(out, dense.astype(np.int64))
    out = np.zeros((4, 40), np.int64)
    assert lib.pc_unpack_vis_host(np.zeros((4, 1), np.uint32).ctypes.data, 4, 1, 40, out.ctypes.data) == -1


def test_only_tests_smoke_and_the_cpu_baseline_leg_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under clockpunch_b200/ or tools/ may import it (the product has no CPU
    path), and outside tests/ only __graft_entry__ (build / smoke) and bench.py (cpu_baseline legs) do."""
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    users = []
    for base, dirs, files in os.walk(root):
        dirs[:] = [d for d in dirs if d not in (".git", "gpurun_out", "__pycache__", "variants", "baseline")]
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(base, f)
            rel = os.path.relpath(path, root)
            if rel.startswith(("tests" + os.sep, "oracle" + os.sep)):
                continue
            for node in ast.walk(ast.parse(open(path, encoding="utf-8").read())):
                mods = ([a.name for a in node.names] if isinstance(node, ast.Import) else
                        [node.module or ""] if isinstance(node, ast.ImportFrom) else [])
                if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                    users.append(rel)
    assert sorted(set(users)) == ["__graft_entry__.py", "bench.py"], sorted(set(users))
