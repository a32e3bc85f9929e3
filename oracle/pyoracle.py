"""ctypes access to the CPU checkers (TEST INFRASTRUCTURE ONLY).

    Port  -> oracle/_build/libc3sc_oracle.so   the C restatement (oracle/c3sc_oracle.c)
    Ref   -> oracle/_ref/libc3sc_ref.so        the reference's own objects (built here only)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "_build", "libc3sc_oracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libc3sc_ref.so")

vp = C.c_void_p
sz = C.c_size_t
f64p = C.POINTER(C.c_double)
szp = C.POINTER(C.c_size_t)
i32p = C.POINTER(C.c_int)


def build_port() -> str:
    subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)
    return PORT_PATH


def build_ref() -> str | None:
    if not os.path.isdir("/root/reference/src"):
        return REF_PATH if os.path.exists(REF_PATH) else None
    subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return REF_PATH


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


class OrcProblem(C.Structure):
    _fields_ = [
        ("dx", sz), ("du", sz), ("dw", sz),
        ("ngrid", szp), ("xgrid", C.POINTER(f64p)),
        ("h2", C.c_double), ("t", f64p), ("bc", i32p),
        ("nobs", sz), ("obs_lb", f64p), ("obs_ub", f64p),
        ("beta", C.c_double), ("nu", sz), ("utab", f64p),
        ("drift", vp), ("drift_arg", vp), ("diff", vp), ("diff_arg", vp),
        ("stage", vp), ("boundcost", vp), ("obscost", vp),
    ]


class OrcFT(C.Structure):
    _fields_ = [("d", sz), ("n", szp), ("ranks", szp), ("cores", C.POINTER(f64p))]


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


class FT:
    """Host-side nodal cores in the layout of valuef_precompute_cores."""

    def __init__(self, n, ranks, cores):
        self.d = len(cores)
        self.n = np.ascontiguousarray(n, dtype=np.uintp)
        self.ranks = np.ascontiguousarray(ranks, dtype=np.uintp)
        self.cores = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores]
        self._arr = (f64p * self.d)(*[c.ctypes.data_as(f64p) for c in self.cores])
        self.c = OrcFT(self.d, self.n.ctypes.data_as(szp), self.ranks.ctypes.data_as(szp), C.cast(self._arr, C.POINTER(f64p)))
        self.flat = np.concatenate(self.cores)


class Port:
    """The oracle restatement, driven on the same inputs as the GPU path."""

    def __init__(self, cfg, xgrid, h2, t, obs_lb, obs_ub):
        if not os.path.exists(PORT_PATH):
            build_port()
        L = self.L = C.CDLL(PORT_PATH)
        for name in ("orc_model_drift", "orc_model_diff", "orc_model_stage", "orc_model_boundcost",
                     "orc_model_obscost", "orc_model_diff_arg"):
            getattr(L, name).restype = vp
        L.orc_rhs.restype = C.c_double
        L.orc_rhs.argtypes = [sz, C.c_double, C.c_double, vp, C.c_double, vp]
        L.orc_ft_eval_linear.restype = C.c_double
        L.orc_model_select.argtypes = [C.c_int, sz, vp, sz]
        self.cfg = cfg
        self.dx = cfg.dx
        params = np.ascontiguousarray(cfg.params, dtype=np.float64)
        # model 0 = geometry only (flags, neighbour indices, neighbour values): no dynamics callbacks
        if cfg.model and L.orc_model_select(cfg.model, cfg.dx, _p(params) if params.size else None, params.size):
            raise ValueError("oracle: unknown model")
        self.ngrid = np.ascontiguousarray(cfg.ngrid, dtype=np.uintp)
        self.nmax = int(self.ngrid.max())
        self.xgrid = [np.ascontiguousarray(g, dtype=np.float64) for g in xgrid]
        self._xg = (f64p * cfg.dx)(*[g.ctypes.data_as(f64p) for g in self.xgrid])
        self.t = np.ascontiguousarray(t, dtype=np.float64)
        self.bc = np.ascontiguousarray(cfg.bc, dtype=np.intc)
        self.obs_lb = np.ascontiguousarray(obs_lb, dtype=np.float64).reshape(-1)
        self.obs_ub = np.ascontiguousarray(obs_ub, dtype=np.float64).reshape(-1)
        self.utab = np.ascontiguousarray(cfg.controls, dtype=np.float64).reshape(-1)
        p = self.p = OrcProblem()
        p.dx, p.du, p.dw = cfg.dx, cfg.du, cfg.dw
        p.ngrid = self.ngrid.ctypes.data_as(szp)
        p.xgrid = C.cast(self._xg, C.POINTER(f64p))
        p.h2 = h2
        p.t = self.t.ctypes.data_as(f64p)
        p.bc = self.bc.ctypes.data_as(i32p)
        p.nobs = self.obs_lb.size // cfg.dx
        p.obs_lb = self.obs_lb.ctypes.data_as(f64p)
        p.obs_ub = self.obs_ub.ctypes.data_as(f64p)
        p.beta = cfg.beta
        p.nu = cfg.nu
        p.utab = self.utab.ctypes.data_as(f64p)
        if cfg.model:
            p.drift = L.orc_model_drift(); p.diff = L.orc_model_diff(); p.diff_arg = L.orc_model_diff_arg()
            p.stage = L.orc_model_stage(); p.boundcost = L.orc_model_boundcost(); p.obscost = L.orc_model_obscost()

    def fiber_points(self, k, fixed):
        fixed = np.ascontiguousarray(fixed, dtype=np.intc)
        x = np.empty((int(self.ngrid[k]), self.dx))
        self.L.orc_fiber_points(C.byref(self.p), sz(k), _p(fixed), _p(x))
        return x

    def fiber_neighbors(self, k, fixed):
        N = int(self.ngrid[k])
        x = self.fiber_points(k, fixed)
        fi = np.ascontiguousarray(fixed, dtype=np.uintp)
        ab = np.empty(N, np.intc); nv = np.empty(2 * N, np.uintp); nf = np.zeros(max(2 * (self.dx - 1), 1), np.uintp)
        rc = self.L.orc_fiber_neighbors(C.byref(self.p), _p(fi), sz(k), _p(x), _p(ab), _p(nv), _p(nf))
        assert rc == 0
        return ab, nv.reshape(N, 2).astype(np.int64), nf.reshape(-1, 2).astype(np.int64)

    def neighbor_costs(self, ft: FT, k, fixed):
        N = int(self.ngrid[k])
        x = self.fiber_points(k, fixed)
        fi = np.zeros(self.dx, np.uintp); kk = sz()
        ab = np.empty(N, np.intc); costs = np.empty((N, 2 * self.dx + 1))
        rc = self.L.orc_neighbor_costs(C.byref(self.p), C.byref(ft.c), sz(N), _p(x), _p(fi), C.byref(kk), _p(ab), _p(costs))
        assert rc == 0 and kk.value == k
        return ab, costs

    def transition(self, drift, sigma_diag):
        drift = np.ascontiguousarray(drift, np.float64); sigma_diag = np.ascontiguousarray(sigma_diag, np.float64)
        n, dx = drift.shape
        prob = np.zeros((n, 2 * dx + 1)); dt = np.zeros(n); st = np.zeros(n, np.int32)
        dd = np.zeros(dx * dx)
        for e in range(n):
            dd[:] = 0.0
            dd[np.arange(dx) * dx + np.arange(dx)] = sigma_diag[e]
            d1 = C.c_double()
            st[e] = self.L.orc_transition(sz(dx), sz(dx), C.c_double(self.p.h2), _p(self.t), _p(drift[e]), _p(dd),
                                          _p(prob[e]), C.byref(d1))
            dt[e] = d1.value
        return prob, dt, st

    def model_eval(self, x, u):
        x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
        n, dx = x.shape
        DYN = C.CFUNCTYPE(C.c_int, C.c_double, vp, vp, vp, vp, vp)
        STG = C.CFUNCTYPE(C.c_int, C.c_double, vp, vp, vp, vp)
        BND = C.CFUNCTYPE(C.c_int, C.c_double, vp, vp)
        OBS = C.CFUNCTYPE(C.c_int, vp, vp)
        fdr, fdf = DYN(self.p.drift), DYN(self.p.diff)
        fst, fbd, fob = STG(self.p.stage), BND(self.p.boundcost), OBS(self.p.obscost)
        drift = np.zeros((n, dx)); sig = np.zeros((n, dx)); stage = np.zeros(n); bound = np.zeros(n); obs = np.zeros(n)
        dd = np.zeros(dx * dx + 64)
        s1 = C.c_double()
        for e in range(n):
            fdr(0.0, _p(x[e]), _p(u[e]), _p(drift[e]), None, None)
            fdf(0.0, _p(x[e]), _p(u[e]), _p(dd), None, self.p.diff_arg)
            sig[e] = dd[np.arange(dx) * dx + np.arange(dx)]
            fst(0.0, _p(x[e]), _p(u[e]), C.byref(s1), None); stage[e] = s1.value
            fbd(0.0, _p(x[e]), C.byref(s1)); bound[e] = s1.value
            fob(_p(x[e]), C.byref(s1)); obs[e] = s1.value
        return drift, sig, stage, bound, obs

    def vi_batch(self, ft: FT, dim_vary, fixed_ind, nthreads=0):
        dv = np.ascontiguousarray(dim_vary, np.intc).reshape(-1)
        fi = np.ascontiguousarray(fixed_ind, np.intc).reshape(-1, self.dx)
        F = dv.size
        out = np.zeros((F, self.nmax)); ub = np.full((F, self.nmax), -1, np.intc)
        rc = self.L.orc_vi_batch(C.byref(self.p), C.byref(ft.c), sz(F), _p(dv), _p(fi), sz(self.nmax), _p(out), _p(ub), C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"oracle vi_batch rc={rc}")
        return out, ub

    def vi_fiber_full(self, ft: FT, k, fixed):
        """one fiber with every intermediate (absorbed, costs)"""
        N = int(self.ngrid[k]); x = self.fiber_points(k, fixed)
        out = np.zeros(N); ub = np.zeros(N, np.intc); ab = np.zeros(N, np.intc); costs = np.zeros((N, 2 * self.dx + 1))
        rc = self.L.orc_vi_fiber(C.byref(self.p), C.byref(ft.c), sz(N), _p(x), _p(out), _p(ub), _p(ab), _p(costs))
        assert rc == 0
        return out, ub, ab, costs

    def pi_batch(self, ft_pol: FT, ft_it: FT, dim_vary, fixed_ind, rows=None, nthreads=0):
        dv = np.ascontiguousarray(dim_vary, np.intc).reshape(-1)
        fi = np.ascontiguousarray(fixed_ind, np.intc).reshape(-1, self.dx)
        F = dv.size
        have = rows is not None
        if not have:
            rows = np.zeros((F, self.nmax, 2 * self.dx + 3))
        ub = np.full((F, self.nmax), -1, np.intc)
        out = np.zeros((F, self.nmax))
        rc = self.L.orc_pi_batch(C.byref(self.p), C.byref(ft_pol.c), C.byref(ft_it.c), sz(F), _p(dv), _p(fi), sz(self.nmax),
                                 C.c_int(int(have)), _p(rows), _p(ub), _p(out), C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"oracle pi_batch rc={rc}")
        return out, rows, ub

    def neighbor_node_costs(self, ft: FT, x):
        x = np.ascontiguousarray(x, np.float64)
        ab = C.c_int(); out = np.zeros(2 * self.dx + 1)
        rc = self.L.orc_neighbor_node_costs(C.byref(self.p), C.byref(ft.c), _p(x), C.byref(ab), _p(out))
        assert rc == 0
        return ab.value, out

    def policy_eval(self, ft: FT, x):
        x = np.ascontiguousarray(x, np.float64)
        u = np.zeros(self.cfg.du); val = C.c_double(); ab = C.c_int(); costs = np.zeros(2 * self.dx + 1)
        rc = self.L.orc_policy_eval(C.byref(self.p), C.byref(ft.c), _p(x), _p(u), C.byref(val), C.byref(ab), _p(costs))
        assert rc == 0
        return u, val.value, ab.value, costs

    def ft_eval_linear(self, ft: FT, x):
        x = np.ascontiguousarray(x, np.float64)
        return self.L.orc_ft_eval_linear(C.byref(ft.c), C.cast(self._xg, C.POINTER(f64p)), _p(x))


class Ref:
    """The reference's own object code (oracle/_ref), one bellman_vi / bellman_pi call per fiber."""

    def __init__(self, cfg):
        if not os.path.exists(REF_PATH):
            raise FileNotFoundError(REF_PATH)
        L = self.L = C.CDLL(REF_PATH)
        L.ref_create.restype = vp
        L.ref_valuef_create.restype = vp
        L.ref_get_hmin.restype = C.c_double
        L.ref_rhs.restype = C.c_double
        L.ref_vi_fibers.restype = C.c_double
        L.ref_pi_fibers.restype = C.c_double
        L.ref_create.argtypes = [C.c_int, sz, vp, sz, vp, vp, vp, C.c_double, vp, sz, vp, vp, sz, vp]
        L.ref_valuef_create.argtypes = [sz, vp, vp, vp]
        L.valuef_destroy.argtypes = [vp]
        L.ref_destroy.argtypes = [vp]
        L.ref_get_grid.argtypes = [vp, sz, vp]
        L.ref_get_hmin.argtypes = [vp]
        L.ref_get_h.argtypes = [vp, vp]
        L.ref_get_h2_t.argtypes = [vp, vp, vp]
        L.ref_get_obstacle.argtypes = [vp, sz, vp, vp]
        L.ref_fiber_neighbors.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.ref_neighbor_costs.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        L.ref_transition.argtypes = [vp, vp, vp, vp, vp]
        L.ref_rhs.argtypes = [vp, C.c_double, C.c_double, vp, C.c_double, vp]
        L.ref_vi_fibers.argtypes = [vp, vp, sz, vp, vp, sz, vp, C.c_int]
        L.ref_pi_begin.argtypes = [vp, vp]
        L.ref_pi_fibers.argtypes = [vp, vp, sz, vp, vp, sz, vp]
        self.cfg = cfg
        self.dx = cfg.dx
        self.ngrid = np.ascontiguousarray(cfg.ngrid, dtype=np.uintp)
        self.nmax = int(self.ngrid.max())
        params = np.ascontiguousarray(cfg.params, dtype=np.float64)
        lb = np.ascontiguousarray(cfg.lb, np.float64); ub = np.ascontiguousarray(cfg.ub, np.float64)
        bc = np.ascontiguousarray(cfg.bc, np.intc)
        nobs = int(cfg.obs_center.shape[0]) if cfg.obs_center.size else 0
        oc = np.ascontiguousarray(cfg.obs_center, np.float64); ow = np.ascontiguousarray(cfg.obs_width, np.float64)
        utab = np.ascontiguousarray(cfg.controls, np.float64)
        self.h = L.ref_create(cfg.model, cfg.dx, _p(params) if params.size else None, params.size, _p(lb), _p(ub),
                              _p(self.ngrid), cfg.beta, _p(bc), nobs, _p(oc) if nobs else None, _p(ow) if nobs else None,
                              cfg.nu, _p(utab))
        if not self.h:
            raise ValueError("reference driver: bad config")
        self.nobs = nobs
        self._vfs = []

    def close(self):
        if getattr(self, "h", None):
            for v in self._vfs:
                self.L.valuef_destroy(v)
            self.L.ref_destroy(self.h)
            self.h = None

    def grids(self):
        out = []
        for i in range(self.dx):
            g = np.empty(int(self.ngrid[i])); self.L.ref_get_grid(self.h, i, _p(g)); out.append(g)
        return out

    def constants(self):
        h = np.empty(self.dx); self.L.ref_get_h(self.h, _p(h))
        h2 = C.c_double(); t = np.empty(2 * self.dx)
        self.L.ref_get_h2_t(self.h, C.byref(h2), _p(t))
        return h, self.L.ref_get_hmin(self.h), h2.value, t

    def obstacles(self):
        lb = np.zeros((self.nobs, self.dx)); ub = np.zeros((self.nobs, self.dx))
        for o in range(self.nobs):
            self.L.ref_get_obstacle(self.h, o, _p(lb[o]), _p(ub[o]))
        return lb, ub

    def valuef(self, ft: FT):
        v = self.L.ref_valuef_create(ft.d, _p(ft.n), _p(ft.ranks), _p(ft.flat))
        self._vfs.append(v)
        return v

    def fiber_neighbors(self, k, fixed):
        N = int(self.ngrid[k]); fixed = np.ascontiguousarray(fixed, np.intc)
        ab = np.empty(N, np.intc); nv = np.empty(2 * N, np.uintp); nf = np.zeros(max(2 * (self.dx - 1), 1), np.uintp)
        rc = self.L.ref_fiber_neighbors(self.h, int(k), _p(fixed), _p(ab), _p(nv), _p(nf))
        assert rc == 0
        return ab, nv.reshape(N, 2).astype(np.int64), nf.reshape(-1, 2).astype(np.int64)

    def neighbor_costs(self, vf, k, fixed):
        N = int(self.ngrid[k]); fixed = np.ascontiguousarray(fixed, np.intc)
        ab = np.empty(N, np.intc); costs = np.empty((N, 2 * self.dx + 1))
        rc = self.L.ref_neighbor_costs(self.h, vf, int(k), _p(fixed), _p(ab), _p(costs))
        assert rc == 0
        return ab, costs

    def neighbor_node_costs(self, vf, x):
        x = np.ascontiguousarray(x, np.float64)
        ab = C.c_int(); out = np.zeros(2 * self.dx + 1)
        rc = self.L.ref_neighbor_node_costs(self.h, vf, _p(x), C.byref(ab), _p(out))
        assert rc == 0
        return ab.value, out

    def transition(self, drift, sigma_diag):
        drift = np.ascontiguousarray(drift, np.float64); sigma_diag = np.ascontiguousarray(sigma_diag, np.float64)
        n, dx = drift.shape
        prob = np.zeros((n, 2 * dx + 1)); dt = np.zeros(n); st = np.zeros(n, np.int32)
        dd = np.zeros(dx * dx)
        for e in range(n):
            dd[:] = 0.0
            dd[np.arange(dx) * dx + np.arange(dx)] = sigma_diag[e]
            d1 = C.c_double()
            st[e] = self.L.ref_transition(self.h, _p(drift[e]), _p(dd), _p(prob[e]), C.byref(d1))
            dt[e] = d1.value
        return prob, dt, st

    def rhs(self, stage, beta, prob, dt, cost):
        prob = np.ascontiguousarray(prob, np.float64); cost = np.ascontiguousarray(cost, np.float64)
        return self.L.ref_rhs(self.h, stage, beta, _p(prob), dt, _p(cost))

    def vi_fibers(self, vf, dim_vary, fixed_ind, fresh_per_fiber=True):
        dv = np.ascontiguousarray(dim_vary, np.intc).reshape(-1)
        fi = np.ascontiguousarray(fixed_ind, np.intc).reshape(-1, self.dx)
        out = np.zeros((dv.size, self.nmax))
        secs = self.L.ref_vi_fibers(self.h, vf, dv.size, _p(dv), _p(fi), self.nmax, _p(out), int(fresh_per_fiber))
        if secs < 0:
            raise RuntimeError("reference bellman_vi failed")
        return out, secs

    def pi_begin(self, vf_policy):
        self.L.ref_pi_begin(self.h, vf_policy)

    def pi_fibers(self, vf_iter, dim_vary, fixed_ind):
        dv = np.ascontiguousarray(dim_vary, np.intc).reshape(-1)
        fi = np.ascontiguousarray(fixed_ind, np.intc).reshape(-1, self.dx)
        out = np.zeros((dv.size, self.nmax))
        secs = self.L.ref_pi_fibers(self.h, vf_iter, dv.size, _p(dv), _p(fi), self.nmax, _p(out))
        if secs < 0:
            raise RuntimeError("reference bellman_pi failed")
        return out, secs

    def omp_threads(self):
        return int(self.L.ref_omp_threads())

    def set_omp_threads(self, n: int):
        self.L.ref_set_omp_threads(int(n))
