"""The ctypes binding a maintainer of the reference adds in place of the Cython helpers
`flps_lgss` / `bpf_lgss` (python/state/particle_methods/cython_lgss_helper.pyx:35,144).
INTEGRATION.md section 2 quotes this file verbatim; tests/test_gpu_configs.py executes it."""
# --- snippet begin
import ctypes as C
import os

import numpy as np

LIB = os.environ.get("PFB200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "..", "qnmh_sysid2018_b200", "csrc", "libpfb200.so")
lib = C.CDLL(LIB)
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
lib.pfb200_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
lib.pfb200_run.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_double, C.c_int, C.c_uint, dp, dp,
                           C.c_uint64, C.POINTER(C.c_uint64), dp, dp, dp, dp, dp, dp, dp, dp, ip, dp]
lib.pfb200_last_error.restype = C.c_char_p
PFB200_MODEL_LGSS, PFB200_F64, PFB200_FLAG_STORE_STATES = 0, 0, 16

_handle = None
_calls = 0


def _run(obs, theta, do_smoother, no_particles, fixed_lag, seed):
    global _handle, _calls
    if _handle is None:
        _handle = C.c_void_p()
        if lib.pfb200_create(-1, PFB200_MODEL_LGSS, PFB200_F64, C.byref(_handle)):
            raise RuntimeError(lib.pfb200_last_error().decode())
    T = len(obs) - 1
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    ll, filt, smo = np.empty(1), np.empty(T + 1), np.empty(T + 1)
    grad, traj, idx = np.empty((4, T + 1)), np.empty(T + 1), np.empty(1, dtype=np.int32)
    eval_id = (C.c_uint64 * 1)(_calls)
    _calls += 1
    rc = lib.pfb200_run(_handle, 1, no_particles, T, fixed_lag, do_smoother, 1, 0.0, 0,
                        PFB200_FLAG_STORE_STATES,
                        obs.ctypes.data_as(dp), theta.ctypes.data_as(dp), seed, eval_id,
                        None, None, None, None,                          # Philox noise
                        ll.ctypes.data_as(dp), filt.ctypes.data_as(dp), smo.ctypes.data_as(dp),
                        grad.ctypes.data_as(dp), idx.ctypes.data_as(ip), traj.ctypes.data_as(dp))
    if rc:
        raise RuntimeError(lib.pfb200_last_error().decode())
    return filt, smo, float(ll[0]), grad, traj


def flps_lgss(obs, mu, phi, sigmav, sigmae, no_particles=1000, fixed_lag=10, seed=87655678):
    """-> (filt, smo, ll, gradient[4 * NoObs], traj) like cython_lgss_helper.pyx:335"""
    filt, smo, ll, grad, traj = _run(obs, [mu, phi, sigmav, sigmae], 1, no_particles, fixed_lag, seed)
    return filt, smo, ll, grad.reshape(-1), traj


def bpf_lgss(obs, mu, phi, sigmav, sigmae, no_particles=1000, seed=87655678):
    """-> (filt, ll, traj) like cython_lgss_helper.pyx:140 (filter only: do_smoother = 0)"""
    filt, _, ll, _, traj = _run(obs, [mu, phi, sigmav, sigmae], 0, no_particles, 0, seed)
    return filt, ll, traj
# --- snippet end
