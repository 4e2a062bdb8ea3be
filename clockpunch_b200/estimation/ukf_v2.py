"""UnscentedKalmanFilter (mirror of punchclock/estimation/ukf_v2.py:26-396).

Same constructor, methods and public attributes as the reference.  The arithmetic of
predict / update runs in the K4 CUDA kernel (csrc/kernels_ukf.cu) for the filter the
reference actually builds on the hot path -- 6-D state, identity measurement model,
satellite dynamics (ez_ukf.py:17-95).  Other configurations (n != 6, custom Python
measurement or dynamics models; the reference's own tests build such filters,
tests/estimation/test_ukf_v2.py:54-170) are not hot-path work: constructing one returns an
instance of the REFERENCE's own UnscentedKalmanFilter, loaded from the reference tree when
one is available (clockpunch_b200._reference), and raises NotImplementedError otherwise.
The hot-path configuration never takes that route and has no CPU fallback.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Callable

import numpy as np
from numpy import array, diagflat, full, ndarray

from .. import _lib
from ..dynamics.dynamics_classes import DynamicsModel, SatDynamicsModel
from .filter_base_class import Filter


def _is_hot_path(est_x, dynamics_model, measurement_model) -> bool:
    """6-D state, identity measurement model (ez_ukf.py:58-59), SatDynamicsModel."""
    probe = np.asarray(est_x, dtype=np.float64).ravel()
    if probe.size != 6 or not isinstance(dynamics_model, SatDynamicsModel):
        return False
    try:
        probe2 = probe + np.arange(1.0, 7.0)
        y1, y2 = np.asarray(measurement_model(probe)), np.asarray(measurement_model(probe2))
    except Exception:
        return False
    return y1.shape == probe.shape and np.array_equal(y1, probe) and np.array_equal(y2, probe2)


def ukf_batch_host(est_x: ndarray, est_p: ndarray, z, tasked, t0: float, tf: float,
                   params: "_lib.UkfParams", substeps: int = 0, force_model: int = _lib.FORCE_TWOBODY):
    """predictAndUpdate for N independent filters through pc_ukf_predict_update_host.

    est_x (6,N), est_p (N,6,6), z (6,N) or None, tasked (N,) bool or None.
    Returns dict(est_x, est_p, pred_x, pred_p, nis, flags)."""
    est_x = np.array(est_x, dtype=np.float64, order="C")
    est_p = np.array(est_p, dtype=np.float64, order="C")
    n = est_x.shape[1]
    zp = kp = None
    if z is not None:
        z = np.ascontiguousarray(z, dtype=np.float64)
        zp = z.ctypes.data
    if tasked is not None:
        tasked = np.ascontiguousarray(tasked, dtype=np.uint8)
        kp = tasked.ctypes.data
    pred_x = np.empty((6, n))
    pred_p = np.empty((n, 6, 6))
    nis = np.empty(n)
    flags = np.empty(n, dtype=np.uint32)
    ctx = _lib.get_ctx()
    ctx.check(ctx.lib.pc_ukf_predict_update_host(
        ctx.handle, est_x.ctypes.data, est_p.ctypes.data, zp, kp, n, float(t0), float(tf),
        int(substeps), int(force_model), params, pred_x.ctypes.data, pred_p.ctypes.data,
        nis.ctypes.data, flags.ctypes.data))
    return dict(est_x=est_x, est_p=est_p, pred_x=pred_x, pred_p=pred_p, nis=nis, flags=flags)


class UnscentedKalmanFilter(Filter):
    """See module docstring; attribute names follow ukf_v2.py:29-55."""

    def __new__(cls, time=None, est_x=None, est_p=None, dynamics_model=None, measurement_model=None, *args, **kwargs):
        if cls is UnscentedKalmanFilter and est_x is not None and not _is_hot_path(est_x, dynamics_model, measurement_model):
            from .. import _reference
            ref_cls = _reference.load("estimation.ukf_v2").UnscentedKalmanFilter
            return ref_cls(time, est_x, est_p, dynamics_model, measurement_model, *args, **kwargs)
        return super().__new__(cls)

    def __init__(self, time: float, est_x: ndarray, est_p: ndarray, dynamics_model: DynamicsModel,
                 measurement_model: Callable, q_matrix: ndarray, r_matrix: ndarray,
                 alpha: float = 0.001, beta: float = 2.0, kappa: float | None = None):
        super().__init__(time, est_x, est_p)
        self.reset_params = {"time": time, "est_x": est_x, "est_p": est_p}
        self.reset()
        self.dynamics = dynamics_model
        self.measurement_model = measurement_model
        self.x_dim = len(est_x)
        self.q_matrix = q_matrix
        self.r_matrix = r_matrix
        self.y_dim = 6
        if not _is_hot_path(est_x, dynamics_model, measurement_model):      # subclasses: __new__ did not divert
            raise NotImplementedError(
                "clockpunch_b200 implements the hot-path filter of the reference (ezUKF: 6-D state, "
                "identity measurement, SatDynamicsModel) as a CUDA kernel; other configurations are served "
                "by the reference's own class (see clockpunch_b200._reference), not by a subclass of this one.")
        gamma, wm0, wi, wc0, _ = _lib.reference_ukf_weights(self.x_dim, alpha, beta, kappa)
        self.gamma = gamma
        self.mean_weight = full(self.num_sigmas, wi)
        self.mean_weight[0] = wm0
        self.cvr_weight = diagflat(self.mean_weight)
        self.cvr_weight[0, 0] = wc0
        self._abk = (alpha, beta, kappa)
        self._pending = None

    # ukf_v2.py:130-151
    def reset(self):
        self.time = deepcopy(self.reset_params["time"])
        self.last_measurement_time = None
        self.est_x = deepcopy(self.reset_params["est_x"])
        self.est_p = deepcopy(self.reset_params["est_p"])
        self.pred_x = deepcopy(self.reset_params["est_x"])
        self.pred_p = deepcopy(self.reset_params["est_p"])
        # Intermediates of the reference that the fused kernel keeps on chip; they stay
        # empty here (no hot-path caller reads them).
        self.nis = array([])
        self.innov_cvr = array([])
        self.cross_cvr = array([])
        self.kalman_gain = array([])
        self.mean_pred_y = array([])
        self.innovation = array([])
        self.sigma_points = array([])
        self.sigma_x_res = array([])
        self.sigma_y_res = array([])
        self._pending = None

    @property
    def num_sigmas(self):
        return 2 * self.x_dim + 1

    def _params(self):
        return _lib.make_ukf_params(self.q_matrix, self.r_matrix, *self._abk)

    def _run(self, t0, tf, x, p, measurement):
        x = np.asarray(x, dtype=np.float64).reshape(6, 1)
        p = np.asarray(p, dtype=np.float64).reshape(1, 6, 6)
        if measurement is None:
            return ukf_batch_host(x, p, None, None, t0, tf, self._params())
        z = np.asarray(measurement, dtype=np.float64).reshape(6, 1)
        return ukf_batch_host(x, p, z, np.ones(1, dtype=np.uint8), t0, tf, self._params())

    def _absorb(self, out, measured: bool, final_time):
        self.pred_x = out["pred_x"][:, 0].copy()
        self.pred_p = out["pred_p"][0].copy()
        self.est_x = out["est_x"][:, 0].copy()
        self.est_p = out["est_p"][0].copy()
        self.time = final_time
        if measured:
            self.nis = out["nis"][0]
            self.last_measurement_time = self.time
        self.flags = int(out["flags"][0])
        self._pending = None

    # ukf_v2.py:185-196
    def predictAndUpdate(self, final_time: float, measurement: ndarray | None):
        if measurement is not None and np.asarray(measurement).ndim > 1:
            raise ValueError("Measurement must be a (M,) array, but has ndim > 1")
        out = self._run(self.time, final_time, self.est_x, self.est_p, measurement)
        self._absorb(out, measurement is not None, final_time)

    # ukf_v2.py:198-209
    def predict(self, final_time: float):
        """pred_x, pred_p at final_time; est_x / est_p untouched until update()."""
        t0, x0, p0 = self.time, self.est_x, self.est_p
        out = self._run(t0, final_time, x0, p0, None)
        self.pred_x = out["pred_x"][:, 0].copy()
        self.pred_p = out["pred_p"][0].copy()
        self.time = final_time
        self._pending = (t0, x0, p0, out)

    # ukf_v2.py:230-259
    def update(self, measurement: ndarray | None):
        if self._pending is None:
            raise RuntimeError("update() must follow predict()")
        t0, x0, p0, out = self._pending
        if measurement is None:
            self._absorb(out, False, self.time)
            return
        if np.asarray(measurement).ndim > 1:
            raise ValueError("Measurement must be a (M,) array, but has ndim > 1")
        # the fused kernel redoes the (deterministic) prediction together with the update
        self._absorb(self._run(t0, self.time, x0, p0, measurement), True, self.time)

    def forecast(self, measurement: ndarray | None):
        """Covariance part of the update (ukf_v2.py:211-228); folded into update()."""
        self.update(measurement)
