"""Thin NumPy-facing wrapper of one libpfb200 handle (device workspaces live inside the handle)."""
import ctypes as C

import numpy as np

from . import _lib as L

MODEL_IDS = {"lgss": L.MODEL_LGSS, "sv": L.MODEL_SV, "sv_leverage": L.MODEL_SVL,
             "lgss_fully_adapted": L.MODEL_LGSS_FA}
DTYPES = {"float64": L.F64, "fp64": L.F64, "f64": L.F64, "float32": L.F32, "fp32": L.F32, "f32": L.F32}
RESAMPLING = {"systematic": L.RESAMPLE_SYSTEMATIC, "stratified": L.RESAMPLE_STRATIFIED,
              "multinomial": L.RESAMPLE_MULTINOMIAL}


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


class Engine(object):
    """One estimator handle bound to a CUDA device: configure -> execute -> fetch."""

    def __init__(self, model="lgss", dtype="float64", device=-1):
        self._lib = L.load()
        self.model = model
        self.model_id = MODEL_IDS[model]
        self.dtype = dtype
        self.no_params = self._lib.pfb200_model_no_params(self.model_id)
        h = C.c_void_p()
        L.check(self._lib.pfb200_create(int(device), self.model_id, DTYPES[dtype], C.byref(h)))
        self._h = h
        self._cfg = None
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pfb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        L.check(self._lib.pfb200_set_option(self._h, name.encode(), int(value)))
        self._cfg = None

    def set_stream(self, cuda_stream_ptr):
        L.check(self._lib.pfb200_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def info(self, name):
        v = C.c_longlong()
        L.check(self._lib.pfb200_get_info(self._h, name.encode(), C.byref(v)))
        return int(v.value)

    # ------------------------------------------------------------------------------------------
    def configure(self, obs, theta, no_particles, fixed_lag=0, do_smoother=True,
                  generate_initial_state=True, initial_state=0.0, resampling="systematic",
                  seed=0, eval_ids=None, draws=None, store_history=False, svl_obs_model=False,
                  store_states=False):
        """obs: (T+1,) shared or (B, T+1); theta: (P,) or (B, P).  `draws`: injected randomness,
        an object with z0 (N,), u (T+1,) [(T+1,N) stratified], z (T+1,N), u_final -- or a list of B
        such objects (one per filter)."""
        theta = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
        B = theta.shape[0]
        if theta.shape[1] != self.no_params:
            raise ValueError("theta must have %d columns" % self.no_params)
        obs = np.ascontiguousarray(np.asarray(obs, dtype=np.float64))
        if obs.ndim == 2 and obs.shape[1] == 1:
            obs = obs.reshape(-1)  # the reference stores obs as (T+1, 1)
        flags = 0
        if obs.ndim == 2:
            if obs.shape[0] != B:
                raise ValueError("per-filter obs must have B rows")
            flags |= L.FLAG_OBS_PER_FILTER
        T = obs.shape[-1] - 1
        N = int(no_particles)
        if store_history:
            flags |= L.FLAG_STORE_HISTORY
        elif store_states:  # all generations of the particles only (state_trajectory rows)
            flags |= L.FLAG_STORE_STATES
        if svl_obs_model:
            flags |= L.FLAG_SVL_OBS_MODEL
        z0 = u = z = uf = None
        if draws is not None:
            strat = resampling in ("stratified", "multinomial")  # one uniform per particle
            ushape = (T + 1, N) if strat else (T + 1,)
            if isinstance(draws, (list, tuple)):
                if len(draws) != B:
                    raise ValueError("need one Draws per filter")
                flags |= L.FLAG_INJECT_PER_FILTER
                z0 = _f64(np.stack([d.z0 for d in draws]), (B, N))
                u = _f64(np.stack([d.u for d in draws]), (B,) + ushape)
                z = _f64(np.stack([d.z for d in draws]), (B, T + 1, N))
                uf = _f64(np.array([d.u_final for d in draws]), (B,))
            else:
                z0, u, z = _f64(draws.z0, (N,)), _f64(draws.u, ushape), _f64(draws.z, (T + 1, N))
                uf = _f64(np.array([draws.u_final]), (1,))
        ev = None
        if eval_ids is not None:
            ev = np.ascontiguousarray(np.asarray(eval_ids, dtype=np.uint64))
            if ev.shape != (B,):
                raise ValueError("eval_ids must have shape (B,)")
        self._keep = [obs, theta, z0, u, z, uf, ev]
        L.check(self._lib.pfb200_configure(
            self._h, B, N, T, int(fixed_lag), int(bool(do_smoother)), int(bool(generate_initial_state)),
            float(initial_state), RESAMPLING[resampling], flags, _dptr(obs), _dptr(theta),
            C.c_uint64(int(seed)), None if ev is None else ev.ctypes.data_as(C.POINTER(C.c_uint64)),
            _dptr(z0), _dptr(u), _dptr(z), _dptr(uf)))
        self._cfg = dict(B=B, N=N, T=T, L=int(fixed_lag), do_smoother=bool(do_smoother),
                         history=bool(store_history), states=bool(store_states))
        return self

    def update_params(self, theta, seed=0, eval_ids=None):
        theta = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
        if theta.shape != (self._cfg["B"], self.no_params):
            raise ValueError("theta shape mismatch")
        ev = None
        if eval_ids is not None:
            ev = np.ascontiguousarray(np.asarray(eval_ids, dtype=np.uint64))
            if ev.shape != (self._cfg["B"],):
                raise ValueError("eval_ids must have shape (B,)")
        L.check(self._lib.pfb200_update_params(
            self._h, _dptr(theta), C.c_uint64(int(seed)),
            None if ev is None else ev.ctypes.data_as(C.POINTER(C.c_uint64))))

    def execute(self):
        L.check(self._lib.pfb200_execute(self._h))

    def synchronize(self):
        L.check(self._lib.pfb200_synchronize(self._h))

    def last_execute_ms(self):
        ms = C.c_float()
        L.check(self._lib.pfb200_last_execute_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def last_kernel_ms(self):
        ms = C.c_float()
        L.check(self._lib.pfb200_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def fetch(self, out=None):
        """Returns dict(log_like (B,), filt (B,T+1), smo (B,T+1), grad_terms (B,P,T+1),
        traj_index (B,), traj (B,T+1)).  `out` may hold preallocated (e.g. pinned) arrays."""
        c = self._cfg
        B, T1, P = c["B"], c["T"] + 1, self.no_params
        out = out or {}
        res = dict(log_like=out.get("log_like", np.empty(B)),
                   filt=out.get("filt", np.empty((B, T1))),
                   smo=out.get("smo", np.empty((B, T1))),
                   grad_terms=out.get("grad_terms", np.empty((B, P, T1))),
                   traj_index=out.get("traj_index", np.empty(B, dtype=np.int32)),
                   traj=out.get("traj", np.empty((B, T1))))
        L.check(self._lib.pfb200_fetch(
            self._h, _dptr(res["log_like"]), _dptr(res["filt"]), _dptr(res["smo"]),
            _dptr(res["grad_terms"]), res["traj_index"].ctypes.data_as(C.POINTER(C.c_int32)),
            _dptr(res["traj"])))
        return res

    def fetch_log_like(self, out=None):
        B = self._cfg["B"]
        ll = out if out is not None else np.empty(B)
        L.check(self._lib.pfb200_fetch(self._h, _dptr(ll), None, None, None, None, None))
        return ll

    def fetch_steps(self):
        c = self._cfg
        shp = (c["B"], c["T"] + 1)
        ll, m, s = np.empty(shp), np.empty(shp), np.empty(shp)
        L.check(self._lib.pfb200_fetch_steps(self._h, _dptr(ll), _dptr(m), _dptr(s)))
        return dict(ll_terms=ll, max_logw=m, sum_w=s)

    def fetch_history(self):
        c = self._cfg
        shp = (c["B"], c["T"] + 1, c["N"])
        x, lw, a = np.empty(shp), np.empty(shp), np.empty(shp, dtype=np.int32)
        L.check(self._lib.pfb200_fetch_history(self._h, _dptr(x), _dptr(lw),
                                               a.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(particles=x, log_weights=lw, ancestors=a)

    def run(self, *args, **kwargs):
        self.configure(*args, **kwargs)
        self.execute()
        return self.fetch()

    # ---- one filter sharded over several GPUs (see include/pfb200.h, pfb200_dist_*)
    def dist_export(self):
        n = self._lib.pfb200_dist_handle_bytes()
        buf = C.create_string_buffer(n)
        L.check(self._lib.pfb200_dist_export(self._h, buf, n))
        return buf.raw

    def dist_connect(self, all_handles, world):
        buf = C.create_string_buffer(all_handles, len(all_handles))
        L.check(self._lib.pfb200_dist_connect(self._h, buf, int(world)))

    def dist_disconnect(self):
        L.check(self._lib.pfb200_dist_disconnect(self._h))

    @staticmethod
    def dist_connect_local(engines):
        """Connects the rank handles of ONE process (engines[r] configured as rank r)."""
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        L.check(engines[0]._lib.pfb200_dist_connect_local(arr, len(engines)))

    def dist_traj_count(self):
        v = C.c_int32()
        L.check(self._lib.pfb200_dist_traj_count(self._h, C.byref(v)))
        return int(v.value)

    def dist_traj_lookup(self, k):
        v = C.c_int32()
        L.check(self._lib.pfb200_dist_traj_lookup(self._h, int(k), C.byref(v)))
        return int(v.value)

    def debug_profile(self, max_ctas=4096):
        """(n_ctas, 12) cycle counters per phase; all zero unless built with -DPFB_PROFILE."""
        buf = np.zeros((max_ctas, 16), dtype=np.int64)
        n = self._lib.pfb200_debug_profile(self._h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), max_ctas)
        if n < 0:
            L.check(n)
        return buf[:n]

    def philox_normals(self, seed, eval_id, t, n):
        z = np.empty(int(n))
        L.check(self._lib.pfb200_philox_normals(self._h, C.c_uint64(int(seed)), C.c_uint64(int(eval_id)),
                                                int(t), int(n), _dptr(z)))
        return z

    def philox_uniforms(self, seed, eval_id, t, count=1):
        u = np.empty(int(count))
        L.check(self._lib.pfb200_philox_uniforms(self._h, C.c_uint64(int(seed)), C.c_uint64(int(eval_id)),
                                                 int(t), int(count), _dptr(u)))
        return u
