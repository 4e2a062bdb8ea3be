"""TEST-ONLY stand-in for the HOST-buffer entry points of libpunch_b200.so, implemented on the oracle.

The mirrored modules marshal their arguments for the C ABI (include/punch_b200.h) and call
`_lib.get_ctx().lib.pc_*_host(...)`.  On the build container there is no GPU, and the GPU box has no reference
tree, so the seam tests (tests/test_seam.py: the reference's own wrappers / runners on top of the mirrors) plug
this object in place of the CDLL: same entry-point names, same argument order and meaning, same status codes,
arithmetic by oracle/*.  It is never imported by the package; the product raises without the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import dynamics as od
from oracle.ukf import OracleUKF
from oracle.visibility import calc_vis_map, calc_vis_map_and_derivative

PC_OK, PC_ERR_BAD_ARG, PC_ERR_EQUAL_TIMES = 0, -1, -3        # include/punch_b200.h
KIND_SATELLITE, KIND_TERRESTRIAL = 0, 1


def _view(ptr, shape, dtype):
    """numpy view (no copy) of host memory handed over as a raw address."""
    if ptr is None:
        return None
    ptr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    n = int(np.prod(shape))
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def _advance(x, kinds, t0, tf):
    out = np.empty_like(x)
    for i in range(x.shape[1]):
        f = od.propagate_terrestrial if kinds is not None and int(kinds[i]) == KIND_TERRESTRIAL else od.propagate_satellite
        out[:, i] = np.asarray(f(x[:, i], t0, tf)).reshape(6)
    return out


class OracleLib:
    """See module docstring.  `calls` counts the entry points used (the tests assert the path went through here)."""

    def __init__(self):
        self.calls = {}
        self._err = b""

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def pc_last_error(self, handle):
        return self._err

    def pc_launch_count(self, handle):
        return 0

    # int pc_propagate_agents_host(ctx, h_x_in, h_x_out, h_kind, n, ld, t0, tf, substeps, force_model)
    def pc_propagate_agents_host(self, handle, x_in, x_out, kind, n, ld, t0, tf, substeps, force_model):
        self._count("pc_propagate_agents_host")
        if t0 == tf:
            self._err = b"t0 == tf"
            return PC_ERR_EQUAL_TIMES
        x = _view(x_in, (6, ld), np.float64)[:, :n]
        out = _view(x_out, (6, ld), np.float64)
        kinds = _view(kind, (n,), np.uint8)
        out[:, :n] = _advance(x.copy(), kinds, t0, tf)
        return PC_OK

    # int pc_ukf_predict_update_host(ctx, h_est_x, h_est_p, h_z, h_tasked, n, t0, tf, substeps, force_model, params,
    #                                h_pred_x, h_pred_p, h_nis, h_flags)
    def pc_ukf_predict_update_host(self, handle, est_x, est_p, z, tasked, n, t0, tf, substeps, force_model, params,
                                   pred_x, pred_p, nis, flags):
        self._count("pc_ukf_predict_update_host")
        if t0 == tf:
            self._err = b"t0 == tf"
            return PC_ERR_EQUAL_TIMES
        ex, ep = _view(est_x, (6, n), np.float64), _view(est_p, (n, 6, 6), np.float64)
        zz, tk = _view(z, (6, n), np.float64), _view(tasked, (n,), np.uint8)
        px, pp = _view(pred_x, (6, n), np.float64), _view(pred_p, (n, 6, 6), np.float64)
        nn, fl = _view(nis, (n,), np.float64), _view(flags, (n,), np.uint32)
        Q, R = np.array(params.Q[:]).reshape(6, 6), np.array(params.R[:]).reshape(6, 6)
        for i in range(n):
            f = OracleUKF(t0, ex[:, i].copy(), ep[i].copy(), Q, R)
            meas = zz[:, i].copy() if (zz is not None and (tk is None or tk[i])) else None
            f.predict_and_update(tf, meas)
            ex[:, i], ep[i], px[:, i], pp[i] = f.est_x, f.est_p, f.pred_x, f.pred_p
            nn[i] = float(f.nis) if meas is not None else np.nan
            fl[i] = 0
        return PC_OK

    # int pc_vismap_host(ctx, h_sens, M, h_targ, N, body_radius, tol, binary, out_i64, out_f64)
    def pc_vismap_host(self, handle, sens, M, targ, N, body_radius, tol, binary, out_i64, out_f64):
        self._count("pc_vismap_host")
        s, t = _view(sens, (6, M), np.float64), _view(targ, (6, N), np.float64)
        vm = calc_vis_map(s, t, body_radius=body_radius, binary=bool(binary))
        if binary:
            _view(out_i64, (N, M), np.int64)[:] = vm
        else:
            _view(out_f64, (N, M), np.float64)[:] = vm
        return PC_OK

    # int pc_vismap_der_host(ctx, h_sens, M, h_targ, N, body_radius, tol, out_v, out_dv)
    def pc_vismap_der_host(self, handle, sens, M, targ, N, body_radius, tol, out_v, out_dv):
        self._count("pc_vismap_der_host")
        s, t = _view(sens, (6, M), np.float64), _view(targ, (6, N), np.float64)
        vm, dm = calc_vis_map_and_derivative(s, t, body_radius=body_radius, binary=False)
        _view(out_v, (N, M), np.float64)[:] = vm
        _view(out_dv, (N, M), np.float64)[:] = dm
        return PC_OK

    # int pc_vis_history_host(ctx, h_sens, h_sens_kind, M, h_targ, h_targ_kind, N, t_now, h_times, T, substeps,
    #                         force_model, body_radius, tol, h_out_words)
    def pc_vis_history_host(self, handle, sens, sens_kind, M, targ, targ_kind, N, t_now, times, T, substeps,
                            force_model, body_radius, tol, out_words):
        self._count("pc_vis_history_host")
        s, t = _view(sens, (6, M), np.float64).copy(), _view(targ, (6, N), np.float64).copy()
        sk, tk = _view(sens_kind, (M,), np.uint8), _view(targ_kind, (N,), np.uint8)
        epochs = _view(times, (T,), np.float64)
        wpr = (M + 31) // 32
        words = _view(out_words, (T, N, wpr), np.uint32)
        words[:] = 0
        now = float(t_now)
        for k in range(T):
            if epochs[k] != now:                      # an epoch equal to the current time: no propagation
                s, t = _advance(s, sk, now, float(epochs[k])), _advance(t, tk, now, float(epochs[k]))
                now = float(epochs[k])
            vm = calc_vis_map(s, t, body_radius=body_radius, binary=True)
            for j in range(M):
                words[k, :, j // 32] |= (vm[:, j].astype(np.uint32) & 1) << np.uint32(j % 32)
        return PC_OK


class OracleContext:
    """Shape of clockpunch_b200._lib.Context around an OracleLib (no device, no handle)."""

    def __init__(self):
        self.lib = OracleLib()
        self.handle = None
        self.device = -1

    def check(self, rc: int) -> None:
        if rc == PC_OK:
            return
        msg = self.lib.pc_last_error(None).decode()
        if rc == PC_ERR_EQUAL_TIMES:
            raise ValueError("Woah buddy, t0 and tf are equal.")      # propagator.py:40-41
        if rc == PC_ERR_BAD_ARG:
            raise ValueError(msg)
        raise RuntimeError(f"oracle double status {rc}: {msg}")

    @property
    def launch_count(self) -> int:
        return 0
