"""ctypes binding of the C-ABI in include/c3sc_b200.h (libc3sc_b200.so).

Plumbing for tests / bench only: the product is the shared library.  Loading
fails loudly when the library has not been built -- there is no Python or CPU
implementation of the backup to fall back to.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("C3SC_LIB") or os.path.join(_HERE, "lib", "libc3sc_b200.so")   # C3SC_LIB: kernel experiments only

c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_u64p = C.POINTER(C.c_uint64)


class ProblemDesc(C.Structure):
    _fields_ = [
        ("dx", C.c_uint32), ("du", C.c_uint32), ("dw", C.c_uint32),
        ("ngrid", c_u64p), ("xgrid", C.POINTER(c_f64p)),
        ("h2", C.c_double), ("t", c_f64p), ("bc", c_i32p),
        ("nobs", C.c_uint32), ("obs_lb", c_f64p), ("obs_ub", c_f64p),
        ("discount", C.c_double),
        ("nu", C.c_uint32), ("controls", c_f64p),
        ("model", C.c_int32), ("model_params", c_f64p), ("n_model_params", C.c_uint32),
        ("arith", C.c_int32),
    ]


class BatchOut(C.Structure):
    _fields_ = [
        ("value", C.c_void_p), ("argmin", C.c_void_p), ("absorbed", C.c_void_p),
        ("costs", C.c_void_p), ("rows", C.c_void_p), ("nbr_vary", C.c_void_p),
        ("nbr_fixed", C.c_void_p),
        ("value_peers", C.c_void_p * 8), ("n_peers", C.c_uint32), ("peer_offset", C.c_uint64), ("peer_mode", C.c_uint32),
    ]


class CrossOpts(C.Structure):
    _fields_ = [("maxiter", C.c_uint32), ("tol", C.c_double), ("verbose", C.c_int)]


class AdaptOpts(C.Structure):
    _fields_ = [("kickrank", C.c_uint32), ("maxrank", C.c_uint32), ("round_tol", C.c_double), ("maxiter_adapt", C.c_uint32)]


# int f(size_t F, const int32_t *dim_vary, const int32_t *fixed_ind, size_t ldo, double *out, void *arg)
FIBER_FN = C.CFUNCTYPE(C.c_int, C.c_size_t, c_i32p, c_i32p, C.c_size_t, c_f64p, C.c_void_p)


EXPORTS = [
    "c3sc_cuda_init", "c3sc_cuda_device_count", "c3sc_last_error", "c3sc_version", "c3sc_launch_count",
    "c3sc_problem_create", "c3sc_problem_destroy", "c3sc_problem_check", "c3sc_problem_control_path",
    "c3sc_valuef_create", "c3sc_valuef_update", "c3sc_valuef_device_buffer", "c3sc_valuef_destroy",
    "c3sc_vi_batch_dev", "c3sc_pi_batch_dev", "c3sc_vi_batch", "c3sc_vi_batch_debug", "c3sc_pi_batch",
    "c3sc_transition_batch", "c3sc_model_eval", "c3sc_measure_fp64_peak", "c3sc_measure_fp64_tensor_peak",
    "c3sc_neighbor_costs_batch", "c3sc_node_backup_batch", "c3sc_control_value_batch", "c3sc_rhs_batch",
    "c3sc_transition_raw", "c3sc_ft_fiber_nn_batch", "c3sc_valuef_commit",
    "c3sc_cross_create", "c3sc_cross_copy", "c3sc_cross_destroy", "c3sc_cross_pin_buffers", "c3sc_cross_ranks", "c3sc_cross_run", "c3sc_cross_run_vi", "c3sc_cross_run_pi",
    "c3sc_vi_solve", "c3sc_cores_dot", "c3sc_cores_norm", "c3sc_cores_norm2diff", "c3sc_cores_dot_l2", "c3sc_cores_norm_l2", "c3sc_cores_norm2diff_l2",
    "c3sc_valuef_eval_batch", "c3sc_policy_eval_batch",
    "c3sc_cross_index_sets", "c3sc_cores_round", "c3sc_cross_adapt_capacity", "c3sc_cross_set_ranks", "c3sc_cross_run_adapt", "c3sc_cross_run_vi_adapt",
    "c3sc_peer_buffer_create", "c3sc_peer_buffer_open", "c3sc_peer_buffer_close",
    "c3sc_fibers_check", "c3sc_fiber_flags_batch", "c3sc_neighbor_node_costs_batch", "c3sc_stage1_batch_dev", "c3sc_pi_batch_resident", "c3sc_rowstore_create", "c3sc_rowstore_reserve", "c3sc_rowstore_destroy", "c3sc_pi_batch_store",
    "c3sc_multi_create", "c3sc_multi_destroy", "c3sc_multi_device_count", "c3sc_multi_uses_nccl", "c3sc_multi_problem",
    "c3sc_multi_valuef_create", "c3sc_multi_valuef_update", "c3sc_multi_valuef_destroy", "c3sc_multi_valuef_get", "c3sc_multi_shard",
    "c3sc_multi_vi_batch", "c3sc_multi_pi_batch", "c3sc_multi_pi_reset", "c3sc_multi_gathered_count", "c3sc_multi_vi_batch_gathered",
    "c3sc_cross_run_vi_multi", "c3sc_cross_run_pi_multi", "c3sc_host_alloc", "c3sc_host_free", "c3sc_vi_batch_peers", "c3sc_guard_check",
    "c3sc_valuef_dot_l2", "c3sc_valuef_norm_l2", "c3sc_valuef_norm2diff_l2",
    "c3sc_cross_dim", "c3sc_cross_uses_memo", "c3sc_fiber_memo_create", "c3sc_fiber_memo_call", "c3sc_fiber_memo_stats", "c3sc_fiber_memo_clear", "c3sc_fiber_memo_destroy",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C c3sc_b200/csrc).  The Bellman backup has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.c3sc_last_error.restype = C.c_char_p
        L.c3sc_version.restype = C.c_char_p
        L.c3sc_launch_count.restype = C.c_uint64
        L.c3sc_problem_destroy.restype = None
        L.c3sc_valuef_destroy.restype = None
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L.c3sc_cuda_init.argtypes = [i32]
        L.c3sc_problem_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp)]
        L.c3sc_problem_destroy.argtypes = [vp]
        L.c3sc_problem_check.argtypes = [vp]
        L.c3sc_problem_control_path.argtypes = [vp]
        L.c3sc_valuef_create.argtypes = [C.c_uint32, c_u64p, c_u64p, C.POINTER(c_f64p), C.POINTER(vp)]
        L.c3sc_valuef_update.argtypes = [vp, C.POINTER(c_f64p)]
        L.c3sc_valuef_device_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
        L.c3sc_valuef_destroy.argtypes = [vp]
        L.c3sc_valuef_commit.argtypes = [vp, vp]
        L.c3sc_vi_batch_dev.argtypes = [vp, vp, sz, vp, vp, sz, C.POINTER(BatchOut), vp]
        L.c3sc_pi_batch_dev.argtypes = [vp, vp, vp, sz, vp, vp, sz, i32, vp, vp, vp, vp]
        L.c3sc_vi_batch.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp]
        L.c3sc_vi_batch_peers.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp, C.POINTER(BatchOut)]
        L.c3sc_vi_batch_debug.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp, vp, vp, vp, vp, vp]
        L.c3sc_fibers_check.argtypes = [vp, sz, vp, vp]
        L.c3sc_stage1_batch_dev.argtypes = [vp, vp, sz, vp, vp, sz, vp]
        L.c3sc_pi_batch_resident.argtypes = [vp, vp, vp, sz, vp, vp, sz, i32, vp, vp]
        L.c3sc_rowstore_create.argtypes = [C.c_uint32, sz, C.POINTER(vp)]
        L.c3sc_rowstore_reserve.argtypes = [vp, sz]
        L.c3sc_rowstore_destroy.argtypes = [vp]; L.c3sc_rowstore_destroy.restype = None
        L.c3sc_multi_create.argtypes = [C.POINTER(ProblemDesc), i32, vp, C.POINTER(vp)]
        L.c3sc_multi_destroy.argtypes = [vp]; L.c3sc_multi_destroy.restype = None
        L.c3sc_multi_device_count.argtypes = [vp]; L.c3sc_multi_uses_nccl.argtypes = [vp]
        L.c3sc_multi_problem.argtypes = [vp, i32]; L.c3sc_multi_problem.restype = vp
        L.c3sc_multi_valuef_create.argtypes = [vp, C.c_uint32, c_u64p, c_u64p, C.POINTER(c_f64p), C.POINTER(vp)]
        L.c3sc_multi_valuef_update.argtypes = [vp, C.POINTER(c_f64p)]
        L.c3sc_multi_valuef_destroy.argtypes = [vp]; L.c3sc_multi_valuef_destroy.restype = None
        L.c3sc_multi_valuef_get.argtypes = [vp, i32]; L.c3sc_multi_valuef_get.restype = vp
        L.c3sc_multi_shard.argtypes = [sz, i32, i32, C.POINTER(sz), C.POINTER(sz)]; L.c3sc_multi_shard.restype = None
        L.c3sc_multi_vi_batch.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp]
        L.c3sc_multi_pi_batch.argtypes = [vp, vp, vp, C.c_uint32, sz, vp, vp, sz, i32, vp]
        L.c3sc_multi_pi_reset.argtypes = [vp]
        L.c3sc_host_alloc.argtypes = [sz, C.POINTER(vp)]; L.c3sc_host_free.argtypes = [vp]
        L.c3sc_multi_gathered_count.argtypes = [sz, i32, sz]; L.c3sc_multi_gathered_count.restype = sz
        L.c3sc_multi_vi_batch_gathered.argtypes = [vp, vp, sz, vp, vp, sz, vp]
        L.c3sc_cross_run_vi_multi.argtypes = [vp, vp, vp, C.POINTER(CrossOpts), C.POINTER(c_f64p), c_u64p, c_f64p]
        L.c3sc_cross_run_pi_multi.argtypes = [vp, vp, vp, vp, C.POINTER(CrossOpts), C.POINTER(c_f64p), c_u64p, c_f64p]
        L.c3sc_pi_batch_store.argtypes = [vp, vp, vp, sz, vp, vp, sz, i32, vp, vp, vp]
        L.c3sc_fiber_flags_batch.argtypes = [vp, sz, vp, vp, sz, vp, vp, vp]
        L.c3sc_neighbor_node_costs_batch.argtypes = [vp, vp, sz, vp, vp, vp]
        L.c3sc_pi_batch.argtypes = [vp, vp, vp, sz, vp, vp, sz, i32, vp, vp, vp]
        L.c3sc_transition_batch.argtypes = [vp, sz, vp, vp, vp, vp, vp]
        L.c3sc_model_eval.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp]
        L.c3sc_neighbor_costs_batch.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp, vp, vp]
        L.c3sc_node_backup_batch.argtypes = [vp, sz, vp, vp, vp, vp, vp]
        L.c3sc_control_value_batch.argtypes = [vp, sz, vp, vp, vp, vp]
        L.c3sc_rhs_batch.argtypes = [i32, C.c_uint32, C.c_double, sz, vp, vp, vp, vp, vp]
        L.c3sc_transition_raw.argtypes = [i32, C.c_uint32, C.c_double, vp, sz, vp, vp, vp, vp, vp]
        L.c3sc_ft_fiber_nn_batch.argtypes = [vp, sz, vp, vp, vp, vp, sz, vp]
        L.c3sc_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), i32, i32]
        L.c3sc_measure_fp64_tensor_peak.argtypes = [C.POINTER(C.c_double), i32, i32]
        L.c3sc_peer_buffer_create.argtypes = [sz, C.POINTER(vp), C.c_char_p]
        L.c3sc_peer_buffer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.c3sc_peer_buffer_close.argtypes = [vp, i32]
        L.c3sc_valuef_eval_batch.argtypes = [vp, vp, sz, vp, vp]
        L.c3sc_policy_eval_batch.argtypes = [vp, vp, sz, vp, vp, vp, vp, vp]
        L.c3sc_cross_create.argtypes = [C.c_uint32, c_u64p, c_u64p, C.POINTER(vp)]
        L.c3sc_cross_destroy.argtypes = [vp]
        L.c3sc_cross_pin_buffers.argtypes = [vp, C.c_int]
        L.c3sc_cross_copy.argtypes = [vp, C.POINTER(vp)]
        L.c3sc_cross_destroy.restype = None
        L.c3sc_cross_ranks.argtypes = [vp, c_u64p]
        L.c3sc_valuef_dot_l2.argtypes = [vp, vp, C.POINTER(c_f64p), c_f64p]
        L.c3sc_valuef_norm_l2.argtypes = [vp, C.POINTER(c_f64p), c_f64p]
        L.c3sc_valuef_norm2diff_l2.argtypes = [vp, vp, C.POINTER(c_f64p), c_f64p]
        L.c3sc_cross_dim.argtypes = [vp]
        L.c3sc_cross_dim.restype = C.c_uint32
        L.c3sc_fiber_memo_create.argtypes = [C.c_uint32, FIBER_FN, vp, C.POINTER(vp)]
        L.c3sc_fiber_memo_call.argtypes = [C.c_size_t, c_i32p, c_i32p, C.c_size_t, c_f64p, vp]
        L.c3sc_fiber_memo_stats.argtypes = [vp, c_u64p, c_u64p]
        L.c3sc_fiber_memo_stats.restype = None
        L.c3sc_fiber_memo_clear.argtypes = [vp]
        L.c3sc_fiber_memo_clear.restype = None
        L.c3sc_fiber_memo_destroy.argtypes = [vp]
        L.c3sc_fiber_memo_destroy.restype = None
        L.c3sc_cross_run.argtypes = [vp, FIBER_FN, vp, C.POINTER(CrossOpts), C.POINTER(c_f64p), c_u64p, c_f64p]
        L.c3sc_cross_run_vi.argtypes = [vp, vp, vp, C.POINTER(CrossOpts), C.POINTER(c_f64p), c_u64p, c_f64p]
        L.c3sc_cross_run_pi.argtypes = [vp, vp, vp, vp, C.c_uint32, C.POINTER(CrossOpts), C.POINTER(c_f64p), c_u64p, c_f64p]
        L.c3sc_vi_solve.argtypes = [vp, vp, c_u64p, C.POINTER(c_f64p), C.c_uint32, C.c_double, C.POINTER(CrossOpts), C.POINTER(c_f64p),
                                    C.POINTER(C.c_uint32), c_f64p, c_u64p]
        L.c3sc_cores_round.argtypes = [C.c_uint32, c_u64p, c_u64p, C.POINTER(c_f64p), C.c_double, c_u64p, C.POINTER(c_f64p)]
        L.c3sc_cross_adapt_capacity.argtypes = [vp, C.POINTER(AdaptOpts), c_u64p]
        L.c3sc_cross_set_ranks.argtypes = [vp, c_u64p]
        L.c3sc_cross_index_sets.argtypes = [vp, C.c_uint32, vp, vp]
        L.c3sc_cross_run_adapt.argtypes = [vp, FIBER_FN, vp, C.POINTER(CrossOpts), C.POINTER(AdaptOpts), c_u64p, C.POINTER(c_f64p),
                                           c_u64p, c_f64p]
        L.c3sc_cross_run_vi_adapt.argtypes = [vp, vp, vp, C.POINTER(CrossOpts), C.POINTER(AdaptOpts), c_u64p, C.POINTER(c_f64p),
                                              c_u64p, c_f64p]
        for fn in (L.c3sc_cores_dot, L.c3sc_cores_norm, L.c3sc_cores_norm2diff, L.c3sc_cores_dot_l2, L.c3sc_cores_norm_l2, L.c3sc_cores_norm2diff_l2):
            fn.restype = C.c_double
        L.c3sc_cores_dot_l2.argtypes = [C.c_uint32, c_u64p, C.POINTER(c_f64p), c_u64p, C.POINTER(c_f64p), c_u64p, C.POINTER(c_f64p)]
        L.c3sc_cores_norm_l2.argtypes = [C.c_uint32, c_u64p, C.POINTER(c_f64p), c_u64p, C.POINTER(c_f64p)]
        L.c3sc_cores_norm2diff_l2.argtypes = [C.c_uint32, c_u64p, C.POINTER(c_f64p), c_u64p, C.POINTER(c_f64p), c_u64p, C.POINTER(c_f64p)]
        L.c3sc_cores_dot.argtypes = [C.c_uint32, c_u64p, c_u64p, C.POINTER(c_f64p), c_u64p, C.POINTER(c_f64p)]
        L.c3sc_cores_norm.argtypes = [C.c_uint32, c_u64p, c_u64p, C.POINTER(c_f64p)]
        L.c3sc_cores_norm2diff.argtypes = [C.c_uint32, c_u64p, c_u64p, C.POINTER(c_f64p), c_u64p, C.POINTER(c_f64p)]
        _lib = L
    return _lib


def measure_fp64_peak(iters: int = 8192, repeats: int = 5) -> float:
    """TFLOP/s of a pure DFMA loop on the current device (roofline denominator)."""
    v = C.c_double()
    check(lib().c3sc_measure_fp64_peak(C.byref(v), iters, repeats))
    return v.value


def measure_fp64_tensor_peak(iters: int = 4096, repeats: int = 5) -> float:
    """TFLOP/s of a pure DMMA (mma.sync m8n8k4 f64) loop on the current device."""
    v = C.c_double()
    check(lib().c3sc_measure_fp64_tensor_peak(C.byref(v), iters, repeats))
    return v.value


class C3scError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise C3scError(f"c3sc error {rc}: {lib().c3sc_last_error().decode()}")


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def grid_constants(xgrid: list[np.ndarray], lb: np.ndarray, ub: np.ndarray):
    """h, hmin, h2, t of c3control_create + mca_add_grid_refs
    (reference src/bellman.c:1975-1986, :181-186), same operation order."""
    dx = len(xgrid)
    h = np.array([xgrid[i][1] - xgrid[i][0] for i in range(dx)], dtype=np.float64)
    hmin = np.float64(ub[0]) - np.float64(lb[0])
    for i in range(dx):
        if h[i] < hmin:
            hmin = h[i]
    h2 = np.float64(hmin) * np.float64(hmin)
    t = np.empty(2 * dx, dtype=np.float64)
    for i in range(dx):
        t[2 * i] = h2 / h[i]
        t[2 * i + 1] = t[2 * i] / h[i]
    return h, float(hmin), float(h2), t


class Problem:
    """Device mirror of the reference's MCAparam + DPparam + Boundary + brute-force table."""

    def __init__(self, cfg, arith: int = 1, xgrid: list[np.ndarray] | None = None):
        from .configs import c3_linspace
        self.cfg = cfg
        self.dx, self.du, self.dw = cfg.dx, cfg.du, cfg.dw
        self.ngrid = np.ascontiguousarray(cfg.ngrid, dtype=np.uint64)
        self.xgrid = xgrid if xgrid is not None else [c3_linspace(cfg.lb[i], cfg.ub[i], int(self.ngrid[i])) for i in range(cfg.dx)]
        self.xgrid = [np.ascontiguousarray(g, dtype=np.float64) for g in self.xgrid]
        self.h, self.hmin, self.h2, self.t = grid_constants(self.xgrid, cfg.lb, cfg.ub)
        self.bc = np.ascontiguousarray(cfg.bc, dtype=np.int32)
        nobs = int(cfg.obs_center.shape[0]) if cfg.obs_center.size else 0
        self.nobs = nobs
        if nobs:
            # BoundRect: lb = center - len/2, ub = center + len/2 (reference src/boundary.c:264-267)
            self.obs_lb = np.ascontiguousarray(cfg.obs_center - cfg.obs_width / 2.0, dtype=np.float64)
            self.obs_ub = np.ascontiguousarray(cfg.obs_center + cfg.obs_width / 2.0, dtype=np.float64)
        else:
            self.obs_lb = np.zeros((0, cfg.dx)); self.obs_ub = np.zeros((0, cfg.dx))
        self.controls = np.ascontiguousarray(cfg.controls, dtype=np.float64)
        self.params = np.ascontiguousarray(cfg.params, dtype=np.float64)
        self.nmax = int(self.ngrid.max())
        self.arith = arith
        d = ProblemDesc()
        d.dx, d.du, d.dw = cfg.dx, cfg.du, cfg.dw
        d.ngrid = self.ngrid.ctypes.data_as(c_u64p)
        self._xg = (c_f64p * cfg.dx)(*[g.ctypes.data_as(c_f64p) for g in self.xgrid])
        d.xgrid = C.cast(self._xg, C.POINTER(c_f64p))
        d.h2 = self.h2
        d.t = self.t.ctypes.data_as(c_f64p)
        d.bc = self.bc.ctypes.data_as(c_i32p)
        d.nobs = nobs
        d.obs_lb = self.obs_lb.ctypes.data_as(c_f64p)
        d.obs_ub = self.obs_ub.ctypes.data_as(c_f64p)
        d.discount = cfg.beta
        d.nu = cfg.nu
        d.controls = self.controls.ctypes.data_as(c_f64p)
        d.model = cfg.model
        d.model_params = self.params.ctypes.data_as(c_f64p) if self.params.size else None
        d.n_model_params = int(self.params.size)
        d.arith = arith
        self.handle = C.c_void_p()
        check(lib().c3sc_problem_create(C.byref(d), C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None) and lib is not None:
            lib().c3sc_problem_destroy(self.handle)
            self.handle = None

    __del__ = close

    def check(self):
        check(lib().c3sc_problem_check(self.handle))

    def fibers_check(self, dim_vary, fixed_ind):
        """c3sc_fibers_check: raises C3scError naming the first descriptor outside the grid."""
        dv, fi, F = _fibers(dim_vary, fixed_ind, self.dx)
        check(lib().c3sc_fibers_check(self.handle, F, _ptr(dv), _ptr(fi)))

    # ---- host-buffer entry points -----------------------------------------------
    def vi_batch(self, vf: "ValueF", dim_vary, fixed_ind, want_argmin=True):
        dv, fi, F = _fibers(dim_vary, fixed_ind, self.dx)
        val = np.empty((F, self.nmax), dtype=np.float64)
        arg = np.empty((F, self.nmax), dtype=np.int32) if want_argmin else None
        check(lib().c3sc_vi_batch(self.handle, vf.handle, F, _ptr(dv), _ptr(fi), self.nmax, _ptr(val), _ptr(arg)))
        return (val, arg) if want_argmin else val

    def vi_batch_debug(self, vf: "ValueF", dim_vary, fixed_ind):
        dv, fi, F = _fibers(dim_vary, fixed_ind, self.dx)
        n, dx = self.nmax, self.dx
        out = dict(
            value=np.empty((F, n)), argmin=np.empty((F, n), np.int32), absorbed=np.empty((F, n), np.int32),
            costs=np.empty((F, n, 2 * dx + 1)), rows=np.empty((F, n, 2 * dx + 3)),
            nbr_vary=np.empty((F, n, 2), np.int32), nbr_fixed=np.zeros((F, max(dx - 1, 1), 2), np.int32))
        check(lib().c3sc_vi_batch_debug(self.handle, vf.handle, F, _ptr(dv), _ptr(fi), n,
                                        *[_ptr(out[k]) for k in ("value", "argmin", "absorbed", "costs", "rows", "nbr_vary", "nbr_fixed")]))
        return out

    def pi_batch(self, vf_policy: "ValueF", vf_iter: "ValueF", dim_vary, fixed_ind, rows=None):
        dv, fi, F = _fibers(dim_vary, fixed_ind, self.dx)
        n, dx = self.nmax, self.dx
        have = rows is not None
        if not have:
            rows = np.empty((F, n, 2 * dx + 3), dtype=np.float64)
        arg = np.empty((F, n), dtype=np.int32)
        val = np.empty((F, n), dtype=np.float64)
        check(lib().c3sc_pi_batch(self.handle, vf_policy.handle if vf_policy is not None else None, vf_iter.handle,
                                  F, _ptr(dv), _ptr(fi), n, int(have), _ptr(rows), _ptr(arg), _ptr(val)))
        return val, rows, (None if have else arg)

    def neighbor_costs(self, vf: "ValueF", dim_vary, fixed_ind):
        dv, fi, F = _fibers(dim_vary, fixed_ind, self.dx)
        n, dx = self.nmax, self.dx
        ab = np.empty((F, n), np.int32); costs = np.empty((F, n, 2 * dx + 1))
        nv = np.empty((F, n, 2), np.int32); nf = np.zeros((F, max(dx - 1, 1), 2), np.int32)
        check(lib().c3sc_neighbor_costs_batch(self.handle, vf.handle, F, _ptr(dv), _ptr(fi), n, _ptr(ab), _ptr(costs), _ptr(nv), _ptr(nf)))
        return ab, costs, nv, nf

    def node_backup(self, x, costs, absorbed=None):
        x = np.ascontiguousarray(x, np.float64); costs = np.ascontiguousarray(costs, np.float64)
        n = x.shape[0]
        ab = None if absorbed is None else np.ascontiguousarray(absorbed, np.int32)
        val = np.empty(n); arg = np.empty(n, np.int32)
        check(lib().c3sc_node_backup_batch(self.handle, n, _ptr(x), _ptr(costs), _ptr(ab), _ptr(val), _ptr(arg)))
        return val, arg

    def transition(self, drift: np.ndarray, sigma: np.ndarray):
        drift = np.ascontiguousarray(drift, np.float64); sigma = np.ascontiguousarray(sigma, np.float64)
        n = drift.shape[0]
        prob = np.empty((n, 2 * self.dx + 1)); dt = np.empty(n); st = np.empty(n, np.int32)
        check(lib().c3sc_transition_batch(self.handle, n, _ptr(drift), _ptr(sigma), _ptr(prob), _ptr(dt), _ptr(st)))
        return prob, dt, st

    def model_eval(self, x: np.ndarray, u: np.ndarray):
        x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
        n = x.shape[0]
        drift = np.empty((n, self.dx)); sig = np.empty((n, self.dx))
        stage = np.empty(n); bound = np.empty(n); obs = np.empty(n)
        check(lib().c3sc_model_eval(self.handle, n, _ptr(x), _ptr(u), _ptr(drift), _ptr(sig), _ptr(stage), _ptr(bound), _ptr(obs)))
        return drift, sig, stage, bound, obs

    def valuef_eval(self, vf: "ValueF", x):
        """valuef_eval at arbitrary points (piecewise-linear FT interpolation)"""
        x = np.ascontiguousarray(x, np.float64).reshape(-1, self.dx)
        out = np.empty(x.shape[0])
        check(lib().c3sc_valuef_eval_batch(self.handle, vf.handle, x.shape[0], _ptr(x), _ptr(out)))
        return out

    def policy_eval(self, vf: "ValueF", x):
        """c3control_policy_eval at n states: (u [n,du], value [n], absorbed [n], costs [n,2dx+1])"""
        x = np.ascontiguousarray(x, np.float64).reshape(-1, self.dx)
        n = x.shape[0]
        u = np.empty((n, self.du)); val = np.empty(n); ab = np.empty(n, np.int32); costs = np.empty((n, 2 * self.dx + 1))
        check(lib().c3sc_policy_eval_batch(self.handle, vf.handle, n, _ptr(x), _ptr(u), _ptr(val), _ptr(ab), _ptr(costs)))
        return u, val, ab, costs

    # ---- device-pointer entry points (ints are raw device addresses) ---------------
    def vi_batch_dev(self, vf: "ValueF", F: int, d_dim_vary: int, d_fixed_ind: int, ldo: int, value: int,
                     argmin: int = 0, stream: int = 0, rows: int = 0, peers=None, peer_offset: int = 0, costs: int = 0,
                     peer_mode: int = 0):
        """peers: device addresses of every rank's gathered buffer (peer-mapped) for the fused all-gather;
        costs alone (no value / argmin / rows): stage 1 only, node-major neighbour values"""
        o = BatchOut(value or None, argmin or None, None, costs or None, rows or None, None, None)
        if peers:
            o.n_peers = len(peers)
            for g, ptr in enumerate(peers):
                o.value_peers[g] = ptr
            o.peer_offset = peer_offset
            o.peer_mode = peer_mode
        check(lib().c3sc_vi_batch_dev(self.handle, vf.handle, F, d_dim_vary, d_fixed_ind, ldo, C.byref(o), stream or None))

    def stage1_batch_dev(self, vf: "ValueF", F: int, d_dim_vary: int, d_fixed_ind: int, ldo: int, stream: int = 0):
        """stage 1 of the pipeline alone, into the library's own scratch (measurement entry)"""
        check(lib().c3sc_stage1_batch_dev(self.handle, vf.handle, F, d_dim_vary, d_fixed_ind, ldo, stream or None))

    def pi_batch_dev(self, vf_policy, vf_iter, F, d_dim_vary, d_fixed_ind, ldo, have_rows, rows, argmin, value, stream=0):
        check(lib().c3sc_pi_batch_dev(self.handle, vf_policy.handle if vf_policy is not None else None, vf_iter.handle, F,
                                      d_dim_vary, d_fixed_ind, ldo, int(have_rows), rows, argmin or None, value, stream or None))


def _fibers(dim_vary, fixed_ind, dx):
    dv = np.ascontiguousarray(dim_vary, dtype=np.int32).reshape(-1)
    fi = np.ascontiguousarray(fixed_ind, dtype=np.int32).reshape(-1, dx)
    assert fi.shape[0] == dv.shape[0]
    return dv, fi, int(dv.shape[0])


class ValueF:
    """Device mirror of ValueF::cores (reference src/valuefunc.c:62-78,165-189)."""

    def __init__(self, n, ranks, cores: list[np.ndarray]):
        self.d = len(cores)
        self.n = np.ascontiguousarray(n, dtype=np.uint64)
        self.ranks = np.ascontiguousarray(ranks, dtype=np.uint64)
        self.cores = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores]
        for k in range(self.d):
            assert self.cores[k].size == int(self.n[k] * self.ranks[k] * self.ranks[k + 1])
        self.handle = C.c_void_p()
        arr = (c_f64p * self.d)(*[c.ctypes.data_as(c_f64p) for c in self.cores])
        check(lib().c3sc_valuef_create(self.d, self.n.ctypes.data_as(c_u64p), self.ranks.ctypes.data_as(c_u64p),
                                       C.cast(arr, C.POINTER(c_f64p)), C.byref(self.handle)))

    def update(self, cores: list[np.ndarray]):
        self.cores = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores]
        arr = (c_f64p * self.d)(*[c.ctypes.data_as(c_f64p) for c in self.cores])
        check(lib().c3sc_valuef_update(self.handle, C.cast(arr, C.POINTER(c_f64p))))

    def device_buffer(self):
        p = C.c_void_p(); n = C.c_size_t()
        check(lib().c3sc_valuef_device_buffer(self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    # valuef_norm / valuef_norm2diff on the device-resident cores (continuous L2 of the piecewise-linear trains)
    def _xg(self, xgrid):
        xg = [np.ascontiguousarray(g, dtype=np.float64) for g in xgrid]
        return xg, (c_f64p * self.d)(*[g.ctypes.data_as(c_f64p) for g in xg])

    def dot_l2(self, other: "ValueF", xgrid) -> float:
        keep, ax = self._xg(xgrid); out = C.c_double()
        check(lib().c3sc_valuef_dot_l2(self.handle, other.handle, ax, C.byref(out)))
        return float(out.value)

    def norm_l2(self, xgrid) -> float:
        keep, ax = self._xg(xgrid); out = C.c_double()
        check(lib().c3sc_valuef_norm_l2(self.handle, ax, C.byref(out)))
        return float(out.value)

    def norm2diff_l2(self, other: "ValueF", xgrid) -> float:
        keep, ax = self._xg(xgrid); out = C.c_double()
        check(lib().c3sc_valuef_norm2diff_l2(self.handle, other.handle, ax, C.byref(out)))
        return float(out.value)

    def commit(self, stream: int = 0):
        """Rebuild the transposed core copy after the device buffer was written directly."""
        check(lib().c3sc_valuef_commit(self.handle, stream or None))

    def close(self):
        if getattr(self, "handle", None) and lib is not None:
            lib().c3sc_valuef_destroy(self.handle)
            self.handle = None

    __del__ = close


class Cross:
    """Host cross-approximation driver with batched fiber requests (include/c3sc_cross.h):
    the stand-in for what valuef_interp gets from C3 (reference src/valuefunc.c:603-767)."""

    def __init__(self, n, ranks):
        self.n = np.ascontiguousarray(n, dtype=np.uint64)
        self.d = int(self.n.size)
        r = np.ascontiguousarray(ranks, dtype=np.uint64)
        self.handle = C.c_void_p()
        check(lib().c3sc_cross_create(self.d, self.n.ctypes.data_as(c_u64p), r.ctypes.data_as(c_u64p), C.byref(self.handle)))
        self.ranks = np.zeros(self.d + 1, dtype=np.uint64)
        check(lib().c3sc_cross_ranks(self.handle, self.ranks.ctypes.data_as(c_u64p)))

    def _cores(self):
        cores = [np.zeros(int(self.n[k] * self.ranks[k] * self.ranks[k + 1])) for k in range(self.d)]
        arr = (c_f64p * self.d)(*[c.ctypes.data_as(c_f64p) for c in cores])
        return cores, arr

    def run_vi(self, prob: "Problem", vf: "ValueF", maxiter=5, tol=0.0, verbose=0):
        """one c3control_step_vi on the GPU path: cores of cross(bellman_vi(.; vf))"""
        cores, arr = self._cores()
        o = CrossOpts(maxiter, tol, verbose)
        nf = C.c_uint64(); ch = C.c_double()
        check(lib().c3sc_cross_run_vi(self.handle, prob.handle, vf.handle, C.byref(o), arr, C.byref(nf), C.byref(ch)))
        return cores, int(nf.value), float(ch.value)

    def run_pi(self, prob: "Problem", vf_policy: "ValueF", vf_iter: "ValueF", maxiter=5, tol=0.0, verbose=0):
        cores, arr = self._cores()
        o = CrossOpts(maxiter, tol, verbose)
        nf = C.c_uint64(); ch = C.c_double()
        check(lib().c3sc_cross_run_pi(self.handle, prob.handle, vf_policy.handle, vf_iter.handle, prob.dx, C.byref(o), arr,
                                      C.byref(nf), C.byref(ch)))
        return cores, int(nf.value), float(ch.value)

    def vi_solve(self, prob: "Problem", ranks0, cores0, maxiter, abs_conv_tol=0.0, sweeps=5, verbose=0):
        """c3control_vi_solve on the GPU path; returns (cores, iterations, last l2 difference, fibers)"""
        r0 = np.ascontiguousarray(ranks0, dtype=np.uint64)
        c0 = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores0]
        a0 = (c_f64p * self.d)(*[c.ctypes.data_as(c_f64p) for c in c0])
        cores, arr = self._cores()
        o = CrossOpts(sweeps, 0.0, verbose)
        it = C.c_uint32(); diff = C.c_double(); nf = C.c_uint64()
        check(lib().c3sc_vi_solve(self.handle, prob.handle, r0.ctypes.data_as(c_u64p), a0, maxiter, abs_conv_tol, C.byref(o), arr,
                                  C.byref(it), C.byref(diff), C.byref(nf)))
        return cores, int(it.value), float(diff.value), int(nf.value)

    def run(self, fn, maxiter=5, tol=0.0, verbose=0, memo=False):
        """same driver, operator = Python callable fn(dim_vary[F], fixed_ind[F,d]) -> values[F, nmax];
        memo: behind a c3sc_fiber_memo (every distinct fiber reaches fn once); returns its (requested, computed) as self.memo_stats"""
        nmax = int(self.n.max())
        d = self.d

        def _cb(F, dv, fi, ldo, out, _arg):
            dvn = np.ctypeslib.as_array(dv, shape=(F,)).copy()
            fin = np.ctypeslib.as_array(fi, shape=(F * d,)).reshape(F, d).copy()
            vals = np.asarray(fn(dvn, fin), dtype=np.float64).reshape(F, -1)
            dst = np.ctypeslib.as_array(out, shape=(F * ldo,)).reshape(F, ldo)
            dst[:, :vals.shape[1]] = vals[:, :ldo]
            return 0
        cb = FIBER_FN(_cb)
        cores, arr = self._cores()
        o = CrossOpts(maxiter, tol, verbose)
        nf = C.c_uint64(); ch = C.c_double()
        if memo:
            mh = C.c_void_p()
            check(lib().c3sc_fiber_memo_create(d, cb, None, C.byref(mh)))
            try:
                fptr = C.cast(lib().c3sc_fiber_memo_call, FIBER_FN)
                check(lib().c3sc_cross_run(self.handle, fptr, mh, C.byref(o), arr, C.byref(nf), C.byref(ch)))
                rq = C.c_uint64(); cp = C.c_uint64()
                lib().c3sc_fiber_memo_stats(mh, C.byref(rq), C.byref(cp))
                self.memo_stats = (int(rq.value), int(cp.value))
            finally:
                lib().c3sc_fiber_memo_destroy(mh)
        else:
            check(lib().c3sc_cross_run(self.handle, cb, None, C.byref(o), arr, C.byref(nf), C.byref(ch)))
        assert nmax >= 1
        return cores, int(nf.value), float(ch.value)

    # ---- rank adaptation (the adapt == 1 branch of valuef_interp) ----
    def _adapt_buffers(self, a):
        cap = np.zeros(self.d + 1, dtype=np.uint64)
        check(lib().c3sc_cross_adapt_capacity(self.handle, C.byref(a), cap.ctypes.data_as(c_u64p)))
        cores = [np.zeros(int(self.n[k] * cap[k] * cap[k + 1])) for k in range(self.d)]
        arr = (c_f64p * self.d)(*[c.ctypes.data_as(c_f64p) for c in cores])
        return cores, arr

    def _adapt_result(self, cores, rout):
        check(lib().c3sc_cross_ranks(self.handle, self.ranks.ctypes.data_as(c_u64p)))
        return [cores[k][:int(self.n[k] * rout[k] * rout[k + 1])].copy() for k in range(self.d)], rout

    def index_sets(self, k):
        """(left, right) multi-index sets at bond k, each [r_k, d]"""
        check(lib().c3sc_cross_ranks(self.handle, self.ranks.ctypes.data_as(c_u64p)))
        r = int(self.ranks[k])
        left = np.zeros((r, self.d), dtype=np.int32); right = np.zeros((r, self.d), dtype=np.int32)
        check(lib().c3sc_cross_index_sets(self.handle, k, left.ctypes.data, right.ctypes.data))
        return left, right

    def set_ranks(self, ranks):
        r = np.ascontiguousarray(ranks, dtype=np.uint64)
        check(lib().c3sc_cross_set_ranks(self.handle, r.ctypes.data_as(c_u64p)))
        check(lib().c3sc_cross_ranks(self.handle, self.ranks.ctypes.data_as(c_u64p)))

    def run_vi_adapt(self, prob, vf, kickrank=2, maxrank=0, round_tol=1e-8, maxiter_adapt=0, maxiter=5, tol=0.0, verbose=0):
        """cross -> round -> kick -> cross ... of bellman_vi(.; vf); returns (cores, ranks, fibers, change)"""
        a = AdaptOpts(kickrank, maxrank, round_tol, maxiter_adapt)
        cores, arr = self._adapt_buffers(a)
        o = CrossOpts(maxiter, tol, verbose)
        rout = np.zeros(self.d + 1, dtype=np.uint64)
        nf = C.c_uint64(); ch = C.c_double()
        check(lib().c3sc_cross_run_vi_adapt(self.handle, prob.handle, vf.handle, C.byref(o), C.byref(a),
                                            rout.ctypes.data_as(c_u64p), arr, C.byref(nf), C.byref(ch)))
        cores, rout = self._adapt_result(cores, rout)
        return cores, rout, int(nf.value), float(ch.value)

    def run_adapt(self, fn, kickrank=2, maxrank=0, round_tol=1e-8, maxiter_adapt=0, maxiter=5, tol=0.0, verbose=0):
        """same with a Python operator fn(dim_vary[F], fixed_ind[F,d]) -> values[F, nmax]"""
        d = self.d

        def _cb(F, dv, fi, ldo, out, _arg):
            dvn = np.ctypeslib.as_array(dv, shape=(F,)).copy()
            fin = np.ctypeslib.as_array(fi, shape=(F * d,)).reshape(F, d).copy()
            vals = np.asarray(fn(dvn, fin), dtype=np.float64).reshape(F, -1)
            dst = np.ctypeslib.as_array(out, shape=(F * ldo,)).reshape(F, ldo)
            dst[:, :vals.shape[1]] = vals[:, :ldo]
            return 0
        cb = FIBER_FN(_cb)
        a = AdaptOpts(kickrank, maxrank, round_tol, maxiter_adapt)
        cores, arr = self._adapt_buffers(a)
        o = CrossOpts(maxiter, tol, verbose)
        rout = np.zeros(self.d + 1, dtype=np.uint64)
        nf = C.c_uint64(); ch = C.c_double()
        check(lib().c3sc_cross_run_adapt(self.handle, cb, None, C.byref(o), C.byref(a), rout.ctypes.data_as(c_u64p), arr,
                                         C.byref(nf), C.byref(ch)))
        cores, rout = self._adapt_result(cores, rout)
        return cores, rout, int(nf.value), float(ch.value)

    def close(self):
        if getattr(self, "handle", None) and lib is not None:
            lib().c3sc_cross_destroy(self.handle)
            self.handle = None

    __del__ = close


def cores_norm2diff(n, ranks_a, cores_a, ranks_b, cores_b) -> float:
    """valuef_norm2diff on nodal cores (discrete l2 over the grid nodes)"""
    n = np.ascontiguousarray(n, dtype=np.uint64); d = int(n.size)
    ra = np.ascontiguousarray(ranks_a, dtype=np.uint64); rb = np.ascontiguousarray(ranks_b, dtype=np.uint64)
    ca = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores_a]
    cb = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores_b]
    aa = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in ca]); ab = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in cb])
    return float(lib().c3sc_cores_norm2diff(d, n.ctypes.data_as(c_u64p), ra.ctypes.data_as(c_u64p), aa, rb.ctypes.data_as(c_u64p), ab))


def cores_norm(n, ranks, cores) -> float:
    n = np.ascontiguousarray(n, dtype=np.uint64); d = int(n.size)
    r = np.ascontiguousarray(ranks, dtype=np.uint64)
    ca = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores]
    aa = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in ca])
    return float(lib().c3sc_cores_norm(d, n.ctypes.data_as(c_u64p), r.ctypes.data_as(c_u64p), aa))


def cores_dot_l2(n, xgrid, ranks_a, cores_a, ranks_b, cores_b) -> float:
    """continuous L2 inner product of two piecewise-linear trains over the box (the reference's function_train_inner on
    LINELM cores): c3sc_cores_dot_l2"""
    n = np.ascontiguousarray(n, dtype=np.uint64); d = int(n.size)
    ra = np.ascontiguousarray(ranks_a, dtype=np.uint64); rb = np.ascontiguousarray(ranks_b, dtype=np.uint64)
    xg = [np.ascontiguousarray(g, dtype=np.float64) for g in xgrid]
    ca = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores_a]
    cb = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores_b]
    ax = (c_f64p * d)(*[g.ctypes.data_as(c_f64p) for g in xg])
    aa = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in ca]); ab = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in cb])
    return float(lib().c3sc_cores_dot_l2(d, n.ctypes.data_as(c_u64p), ax, ra.ctypes.data_as(c_u64p), aa, rb.ctypes.data_as(c_u64p), ab))


def cores_norm2diff_l2(n, xgrid, ranks_a, cores_a, ranks_b, cores_b) -> float:
    n = np.ascontiguousarray(n, dtype=np.uint64); d = int(n.size)
    ra = np.ascontiguousarray(ranks_a, dtype=np.uint64); rb = np.ascontiguousarray(ranks_b, dtype=np.uint64)
    xg = [np.ascontiguousarray(g, dtype=np.float64) for g in xgrid]
    ca = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores_a]
    cb = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores_b]
    ax = (c_f64p * d)(*[g.ctypes.data_as(c_f64p) for g in xg])
    aa = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in ca]); ab = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in cb])
    return float(lib().c3sc_cores_norm2diff_l2(d, n.ctypes.data_as(c_u64p), ax, ra.ctypes.data_as(c_u64p), aa, rb.ctypes.data_as(c_u64p), ab))


class PeerBuffers:
    """Every rank's gathered buffer, peer-mapped into this process (fused all-gather).
    `exchange(handle_bytes) -> list of every rank's handle` is the out-of-band step
    (torch.distributed.all_gather_object in bench.py)."""

    def __init__(self, nbytes: int, rank: int, world: int, exchange):
        self.rank, self.world = rank, world
        own = C.c_void_p()
        hbuf = C.create_string_buffer(64)
        check(lib().c3sc_peer_buffer_create(nbytes, C.byref(own), hbuf))
        self.own = own.value
        handles = exchange(hbuf.raw)
        self.ptrs = []
        for g in range(world):
            if g == rank:
                self.ptrs.append(self.own)
            else:
                p = C.c_void_p()
                check(lib().c3sc_peer_buffer_open(C.create_string_buffer(handles[g], 64), C.byref(p)))
                self.ptrs.append(p.value)

    def close(self):
        for g, p in enumerate(getattr(self, "ptrs", [])):
            lib().c3sc_peer_buffer_close(p, int(g != self.rank))
        self.ptrs = []


class McastPeers:
    """The gathered buffers of all ranks behind ONE NVSwitch multicast address (bench.py, C3SC_GATHER=mcast): to the
    library it is a single `peer` -- a store to it lands in every rank's copy."""

    def __init__(self, mc_ptr: int):
        self.ptrs = [int(mc_ptr)]

    def close(self):
        self.ptrs = []


def cores_round(n, ranks, cores, eps):
    """function_train_round on nodal cores (c3sc_cores_round): returns (ranks, cores)"""
    n = np.ascontiguousarray(n, dtype=np.uint64); d = int(n.size)
    r = np.ascontiguousarray(ranks, dtype=np.uint64)
    cin = [np.ascontiguousarray(c, dtype=np.float64).reshape(-1) for c in cores]
    ain = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in cin])
    cout = [np.zeros_like(c) for c in cin]
    aout = (c_f64p * d)(*[c.ctypes.data_as(c_f64p) for c in cout])
    rout = np.zeros(d + 1, dtype=np.uint64)
    check(lib().c3sc_cores_round(d, n.ctypes.data_as(c_u64p), r.ctypes.data_as(c_u64p), ain, float(eps),
                                 rout.ctypes.data_as(c_u64p), aout))
    return rout, [cout[k][:int(n[k] * rout[k] * rout[k + 1])].copy() for k in range(d)]
