"""ctypes binding of libpunch_b200.so (C-ABI in include/punch_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import threading
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (PUNCH_B200_LIB: variant builds of tools/build_variant.sh, e.g. the phase-stamp library of tools/debug)
LIB_PATH = os.environ.get("PUNCH_B200_LIB") or os.path.join(_HERE, "libpunch_b200.so")

PC_OK, PC_ERR_BAD_ARG, PC_ERR_CUDA, PC_ERR_EQUAL_TIMES, PC_ERR_NO_DEVICE = 0, -1, -2, -3, -4
KIND_SATELLITE, KIND_TERRESTRIAL = 0, 1
FORCE_TWOBODY, FORCE_TWOBODY_J2 = 0, 1
FLAG_NONPD_EST, FLAG_NONPD_PRED, FLAG_NONPD_INNOV, FLAG_NONFINITE, FLAG_SUBSTEP_CLAMP = 1, 2, 4, 8, 16
OPT_FINISH_KERNEL, OPT_PROPAGATE_KERNEL = 1, 2          # pc_option
FINISH_AUTO, FINISH_QUAD_UPDATE_WARPS, FINISH_QUAD_INLINE, FINISH_THREAD_64, FINISH_THREAD_128 = 0, 1, 2, 3, 4
COMM_HANDLE_BYTES = 64

c_dp = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_u32p = C.POINTER(C.c_uint32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)


class UkfParams(C.Structure):
    """pc_ukf_params."""
    _fields_ = [("gamma", C.c_double), ("wm0", C.c_double), ("wi", C.c_double), ("wc0", C.c_double),
                ("eps_w", C.c_double), ("Q", C.c_double * 36), ("R", C.c_double * 36),
                ("Lr", C.c_double * 36)]


class StepArgs(C.Structure):
    """pc_step_args."""
    _fields_ = [
        ("sens_x", C.c_void_p), ("sens_kind", C.c_void_p), ("M", C.c_int64), ("ld_s", C.c_int64),
        ("truth_x", C.c_void_p), ("est_x", C.c_void_p), ("est_p", C.c_void_p), ("z", C.c_void_p),
        ("tasked", C.c_void_p), ("N", C.c_int64), ("ld", C.c_int64),
        ("t0", C.c_double), ("tf", C.c_double), ("substeps", C.c_int), ("force_model", C.c_int),
        ("rng_seed", C.c_uint64), ("rng_step", C.c_uint64),
        ("pred_p", C.c_void_p), ("nis", C.c_void_p), ("flags", C.c_void_p),
        ("vis_truth", C.c_void_p), ("vis_est", C.c_void_p), ("words_per_row", C.c_int64),
        ("body_radius", C.c_double), ("vis_tol", C.c_double),
        ("num_tasked", C.c_void_p), ("last_time_tasked", C.c_void_p), ("tr_pos", C.c_void_p),
        ("tr_vel", C.c_void_p), ("clock", C.c_void_p),
        ("actions", C.c_void_p), ("policy_probes", C.c_int), ("sigma_valid", C.c_int),
        ("policy_seed", C.c_uint64), ("sigma_ws", C.c_void_p), ("sig_flags", C.c_void_p),
        ("sens_q", C.c_void_p), ("env_targets", C.c_int64), ("env_sensors", C.c_int64),
        ("peer_table", C.c_void_p), ("peer_world", C.c_int32), ("peer_rank", C.c_int32),
        ("summary_base", C.c_void_p), ("summary_bytes", C.c_int64),
        ("task_of", C.c_void_p),
        ("mirror_dev", C.c_void_p), ("mirror_host", C.c_void_p), ("mirror_bytes", C.c_int64),
        ("obs_cov_host", C.c_void_p),
    ]


class StepIO(C.Structure):
    """pc_step_io."""
    _fields_ = [("h_actions", C.c_void_p), ("d_obs", C.c_void_p), ("h_obs", C.c_void_p),
                ("h_vis_est", C.c_void_p), ("h_num_tasked", C.c_void_p),
                ("d_block", C.c_void_p), ("h_block", C.c_void_p), ("block_bytes", C.c_int64),
                ("zero_copy", C.c_int32)]


_vp, _i64, _dbl, _int, _u64 = C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_uint64
_SIGNATURES = {
    "pc_version": ([], C.c_int),
    "pc_ctx_create": ([_int, C.POINTER(_vp)], _int),
    "pc_ctx_destroy": ([_vp], _int),
    "pc_last_error": ([_vp], C.c_char_p),
    "pc_launch_count": ([_vp], _i64),
    "pc_set_option": ([_vp, _int, _int], _int),
    "pc_get_option": ([_vp, _int, C.POINTER(_int)], _int),
    "pc_propagate_twobody": ([_vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _int, _int, _vp], _int),
    "pc_rotate_terrestrial": ([_vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp], _int),
    "pc_propagate_agents": ([_vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _int, _int, _vp], _int),
    "pc_propagate_agents_host": ([_vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _int, _int], _int),
    "pc_frame_change": ([_vp, _vp, _vp, _i64, _i64, _dbl, _int, _vp], _int),
    "pc_coe2eci": ([_vp, _vp, _vp, _i64, _i64, _dbl, _vp], _int),
    "pc_lla2eci": ([_vp, _vp, _vp, _i64, _i64, _dbl, _vp], _int),
    "pc_vismap_packed": ([_vp, _vp, _i64, _i64, _vp, _dbl, _vp, _vp, _i64, _i64, _dbl, _dbl, _vp, _vp,
                          _i64, _vp], _int),
    "pc_vismap_value": ([_vp, _vp, _i64, _i64, _vp, _i64, _i64, _dbl, _dbl, _vp, _vp], _int),
    "pc_vismap_host": ([_vp, _vp, _i64, _vp, _i64, _dbl, _dbl, _int, _vp, _vp], _int),
    "pc_vismap_packed_envs": ([_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl, _vp, _vp, _vp], _int),
    "pc_obs_pack_envs": ([_vp, _vp, _i64, _vp, _vp, _i64, _vp, _dbl, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp], _int),
    "pc_vismap_der": ([_vp, _vp, _i64, _i64, _vp, _i64, _i64, _dbl, _dbl, _vp, _vp, _vp], _int),
    "pc_vismap_der_host": ([_vp, _vp, _i64, _vp, _i64, _dbl, _dbl, _vp, _vp], _int),
    "pc_vis_history": ([_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _dbl, _vp, _i64, _int, _int, _dbl, _dbl,
                        _vp, _vp], _int),
    "pc_vis_history_host": ([_vp, _vp, _vp, _i64, _vp, _vp, _i64, _dbl, _vp, _i64, _int, _int, _dbl, _dbl, _vp], _int),
    "pc_ukf_predict_update": ([_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _int, _int,
                               C.POINTER(UkfParams), _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp], _int),
    "pc_sigma_points": ([_vp, _vp, _vp, _i64, _i64, _dbl, _vp, _vp, _vp], _int),
    "pc_ukf_predict_update_host": ([_vp, _vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _int, _int,
                                    C.POINTER(UkfParams), _vp, _vp, _vp, _vp], _int),
    "pc_task_from_actions": ([_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp], _int),
    "pc_obs_pack": ([_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _dbl, _vp, _vp, _vp, _vp, _vp], _int),
    "pc_policy_random_visible": ([_vp, _vp, _i64, _i64, _i64, _u64, _u64, _vp, _int, _vp, _vp], _int),
    "pc_policy_random_visible_host": ([_vp, _i64, _i64, _i64, _u64, _u64, _int, _vp], _int),
    "pc_policy_random_visible_envs_host": ([_vp, _i64, _i64, _i64, _i64, _u64, _u64, _int, _vp], _int),
    "pc_unpack_vis_host": ([_vp, _i64, _i64, _i64, _vp], _int),
    "pc_set_timing": ([_vp, _int], _int),
    "pc_get_timing": ([_vp, C.POINTER(_dbl)], _int),
    "pc_get_kernel_timing": ([_vp, C.POINTER(_dbl), C.POINTER(_dbl)], _int),
    "pc_fp64_peak": ([_vp, _int, C.POINTER(_dbl)], _int),
    "pc_step_fused": ([_vp, C.POINTER(StepArgs), C.POINTER(UkfParams), _vp], _int),
    "pc_peer_status": ([_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)], _int),
    "pc_peer_exchange": ([_vp, _vp, C.c_int32, C.c_int32, _vp, _i64, _vp], _int),
    "pc_comm_init": ([_vp, C.c_int32, C.c_int32, _i64, _vp], _int),
    "pc_comm_attach": ([_vp, _vp], _int),
    "pc_comm_table": ([_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                       C.POINTER(_i64)], _int),
    "pc_allgather_step": ([_vp, _vp, _vp], _int),
    "pc_comm_detach": ([_vp], _int),
    "pc_comm_destroy": ([_vp], _int),
    "pc_step_host": ([_vp, C.POINTER(StepArgs), C.POINTER(UkfParams), C.POINTER(StepIO), _vp], _int),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lib_lock = threading.Lock()
_ctxs: dict[int, "Context"] = {}


def load_library() -> C.CDLL:
    """dlopen libpunch_b200.so.  Raises (never falls back) if it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -m clockpunch_b200.build` "
                    "(nvcc, sm_100a).  clockpunch_b200 has no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (argtypes, restype) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = argtypes
                fn.restype = restype
            _lib = lib
    return _lib


class PunchError(RuntimeError):
    pass


class Context:
    """One pc_ctx per (process, device)."""

    def __init__(self, device: int):
        self.lib = load_library()
        self.device = device
        h = C.c_void_p()
        rc = self.lib.pc_ctx_create(device, C.byref(h))
        if rc == PC_ERR_NO_DEVICE:
            raise PunchError("no CUDA device available: clockpunch_b200 needs a GPU (no CPU fallback)")
        if rc != PC_OK:
            raise PunchError(f"pc_ctx_create(device={device}) failed with status {rc}")
        self.handle = h

    def check(self, rc: int) -> None:
        if rc == PC_OK:
            return
        msg = (self.lib.pc_last_error(self.handle) or b"").decode()
        if rc == PC_ERR_EQUAL_TIMES:
            raise ValueError("Woah buddy, t0 and tf are equal.")      # propagator.py:40-41
        if rc == PC_ERR_BAD_ARG:
            raise ValueError(msg)
        raise PunchError(f"libpunch_b200 status {rc}: {msg}")

    @property
    def launch_count(self) -> int:
        return int(self.lib.pc_launch_count(self.handle))

    def set_option(self, option: int, value: int) -> int:
        """pc_set_option; returns the previous value (kernel-mapping overrides: tests, tuning)."""
        old = C.c_int(0)
        self.check(self.lib.pc_get_option(self.handle, option, C.byref(old)))
        self.check(self.lib.pc_set_option(self.handle, option, value))
        return old.value

    def close(self):
        if self.handle:
            self.lib.pc_ctx_destroy(self.handle)
            self.handle = None


def default_device() -> int:
    """PUNCHCLOCK_DEVICE, else LOCAL_RANK (torchrun), else 0.  Ray never grants GPUs to
    the reference's workers (build_tuner.py:91-95), so selection cannot rely on it."""
    for var in ("PUNCHCLOCK_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v not in (None, ""):
            return int(v)
    return 0


def get_ctx(device: int | None = None) -> Context:
    dev = default_device() if device is None else int(device)
    ctx = _ctxs.get(dev)
    if ctx is None:
        ctx = _ctxs[dev] = Context(dev)
    return ctx


# --------------------------------------------------------------------------
def reference_ukf_weights(n: int = 6, alpha: float = 0.001, beta: float = 2.0, kappa=None):
    """gamma, wm0, wi, wc0 in the reference's float64 operation order (ukf_v2.py:113-128),
    plus eps_w = (wm0 + 2n wi) - 1 evaluated exactly on those float64 values."""
    if not kappa:
        kappa = 3 - n
    lam = (alpha**2) * (n + kappa) - n
    gamma = float(np.sqrt(n + lam))
    wm0 = lam / (n + lam)
    wi = 1 / (2.0 * (lam + n))
    wc0 = wm0 + (1 - alpha**2.0 + beta)
    eps_w = float(Fraction(wm0) + 2 * n * Fraction(wi) - 1)
    return gamma, wm0, wi, wc0, eps_w


def make_ukf_params(Q, R, alpha: float = 0.001, beta: float = 2.0, kappa=None) -> UkfParams:
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    if Q.shape != (6, 6) or R.shape != (6, 6):
        raise ValueError("Q and R must be (6, 6)")
    p = UkfParams()
    p.gamma, p.wm0, p.wi, p.wc0, p.eps_w = reference_ukf_weights(6, alpha, beta, kappa)
    p.Q[:] = Q.ravel().tolist()
    p.R[:] = R.ravel().tolist()
    try:
        Lr = np.linalg.cholesky(R)
    except np.linalg.LinAlgError:
        Lr = np.zeros((6, 6))
    p.Lr[:] = Lr.ravel().tolist()
    return p


def np_ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data
