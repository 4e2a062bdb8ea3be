"""Oracle: unscented Kalman filter of the reference (test infrastructure).

Follows punchclock/estimation/ukf_v2.py:58-396 and estimation/ez_ukf.py:17-95
step for step, including its quirks:
  * kappa = 3 - n whenever ``not kappa``            (ukf_v2.py:113-114)
  * sigma offsets are COLUMNS of the UPPER Cholesky factor (ukf_v2.py:171,179-183)
  * update() with a measurement re-samples the sigma points around
    (pred_x, pred_p) but keeps the state residuals of the propagated points
    (ukf_v2.py:211-228 vs 278-292, 330-332)
  * without a measurement est_x is the propagated centre point, not the
    weighted mean                                        (ukf_v2.py:238-244)
The non-PD fallback (ukf_utils.py:155-198,242-256) is restated in
``nearest_pd_cholesky``.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import LinAlgError, cholesky, inv

from .dynamics import propagate_satellite, propagate_terrestrial


def ukf_weights(n=6, alpha=0.001, beta=2.0, kappa=None):
    """gamma, mean weights (2n+1,), covariance weights (2n+1,) in the reference's
    own float64 evaluation order (ukf_v2.py:113-128)."""
    if not kappa:
        kappa = 3 - n
    lam = (alpha**2) * (n + kappa) - n
    gamma = np.sqrt(n + lam)
    w0 = lam / (n + lam)
    w = 1 / (2.0 * (lam + n))
    wm = np.full(2 * n + 1, w)
    wm[0] = w0
    wc = wm.copy()
    wc[0] += 1 - alpha**2.0 + beta
    return gamma, wm, wc


def nearest_pd_cholesky(cov):
    """Upper Cholesky factor of the nearest PD matrix (ukf_utils.py:155-198,242-256)."""
    sym = (cov + cov.T) / 2
    _, s, vt = np.linalg.svd(sym)
    pol = np.linalg.multi_dot((vt.T, np.diag(s), vt))
    pd = (sym + pol) / 2
    pd = (pd + pd.T) / 2

    def is_pd(m):
        try:
            np.linalg.cholesky(m)
            return True
        except np.linalg.LinAlgError:
            return False

    if not is_pd(pd):
        spacing = np.spacing(np.linalg.norm(cov))
        eye = np.eye(cov.shape[0])
        k = 1
        while not is_pd(pd):
            min_eig = np.amin(np.real(np.linalg.eigvals(pd)))
            pd += eye * (-min_eig * k**2 + spacing)
            k += 1
    return cholesky(pd)


class OracleUKF:
    """State container + predict/update mirroring UnscentedKalmanFilter."""

    def __init__(self, time, est_x, est_p, q_matrix, r_matrix, dynamics="satellite",
                 measurement_model=None, alpha=0.001, beta=2.0, kappa=None):
        self.time = time
        self.last_measurement_time = None
        self.est_x = np.array(est_x, dtype=float)
        self.est_p = np.array(est_p, dtype=float)
        self.pred_x = self.est_x.copy()
        self.pred_p = self.est_p.copy()
        self.q_matrix = np.asarray(q_matrix, dtype=float)
        self.r_matrix = np.asarray(r_matrix, dtype=float)
        self.h = measurement_model or (lambda s: s)          # ez_ukf.py:58-59
        self.propagate = {"satellite": propagate_satellite,
                          "terrestrial": propagate_terrestrial}.get(dynamics, dynamics)
        self.x_dim = len(self.est_x)
        self.y_dim = len(self.h(self.est_x))
        self.gamma, self.mean_weight, wc = ukf_weights(self.x_dim, alpha, beta, kappa)
        self.cvr_weight = np.diagflat(wc)
        self.nis = np.array([])
        self.used_nearest_pd = False

    # ukf_v2.py:158-183
    def sigma_points_of(self, mean, cov):
        try:
            u = cholesky(cov)                    # upper: cov = u.T @ u
        except LinAlgError:
            self.used_nearest_pd = True
            u = nearest_pd_cholesky(cov)
        n = self.x_dim
        return mean.reshape(n, 1) @ np.ones((1, 2 * n + 1)) + self.gamma * np.concatenate(
            (np.zeros((n, 1)), u, -u), axis=1)

    # ukf_v2.py:198-209, 261-292
    def predict(self, final_time):
        x_k = self.sigma_points_of(self.est_x, self.est_p)
        self.sigma_points = self.propagate(x_k, self.time, final_time)
        self.pred_x = self.sigma_points.dot(self.mean_weight)
        self.sigma_x_res = self.sigma_points - self.pred_x.reshape(-1, 1)
        self.pred_p = self.sigma_x_res.dot(self.cvr_weight.dot(self.sigma_x_res.T)) + self.q_matrix
        self.time = final_time

    # ukf_v2.py:211-259, 294-375
    def update(self, measurement):
        if measurement is None:
            self.est_x = self.sigma_points[:, 0].copy()
            self.est_p = self.pred_p.copy()
            return
        measurement = np.asarray(measurement, dtype=float)
        if measurement.ndim > 1:
            raise ValueError("Measurement must be a (M,) array, but has ndim > 1")
        self.sigma_points = self.sigma_points_of(self.pred_x, self.pred_p)
        ns = 2 * self.x_dim + 1
        sigma_obs = np.zeros((self.y_dim, ns))
        for i in range(ns):
            sigma_obs[:, i] = np.asarray(self.h(self.sigma_points[:, i])).squeeze()
        self.mean_pred_y = np.array([row.dot(self.mean_weight) for row in sigma_obs])
        self.sigma_y_res = sigma_obs - self.mean_pred_y.reshape(-1, 1)
        self.innov_cvr = self.sigma_y_res.dot(self.cvr_weight.dot(self.sigma_y_res.T)) + self.r_matrix
        self.cross_cvr = self.sigma_x_res.dot(self.cvr_weight.dot(self.sigma_y_res.T))
        self.kalman_gain = self.cross_cvr.dot(inv(self.innov_cvr))
        self.est_p = self.pred_p - self.kalman_gain.dot(self.innov_cvr.dot(self.kalman_gain.T))
        self.innovation = measurement - self.mean_pred_y
        self.nis = self.innovation.T.dot(inv(self.innov_cvr).dot(self.innovation))
        self.est_x = self.pred_x + self.kalman_gain.dot(self.innovation)
        self.last_measurement_time = self.time

    def predict_and_update(self, final_time, measurement):
        self.predict(final_time)
        self.update(measurement)


def ez_ukf(x_init, p_init, Q, R, dynamics="satellite", time=0):
    """6-D fully observable shortcut (ez_ukf.py:17-95); scalars become v*I6."""
    def mat(v):
        return v * np.eye(6) if np.isscalar(v) else np.asarray(v, dtype=float)
    return OracleUKF(time, x_init, mat(p_init), mat(Q), mat(R), dynamics=dynamics)
