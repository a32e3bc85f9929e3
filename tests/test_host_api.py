"""The C host mirror (include/c3sc_host.h): reference-named entry points, driven the way the
reference's own tests drive them (tprob_test.c: build the structs, call bellman_vi / bellman_pi
one fiber of points at a time), compared with the oracle."""
import ctypes as C

import numpy as np
import pytest

from c3sc_b200 import capi, configs, synthetic
from oracle import pyoracle as po
from helpers import make_ft, make_port, rel_err

vp, sz, dbl = C.c_void_p, C.c_size_t, C.c_double
BC_NAME = {configs.ABSORB: b"absorb", configs.PERIODIC: b"periodic", configs.REFLECT: b"reflect"}


def host_lib():
    L = capi.lib()
    for name in ("c3control_create", "c3control_get_dp", "c3control_get_mca", "c3control_get_work", "c3control_get_xgrid",
                 "c3control_get_boundary", "c3opt_alloc", "valuef_from_cores", "control_params_create", "vi_param_create",
                 "pi_param_create", "boundary_obstacle_get_lb", "valuef_get_ranks", "valuef_copy"):
        getattr(L, name).restype = vp
    L.c3control_create.argtypes = [sz, sz, sz, vp, vp, vp, dbl]
    L.c3control_set_external_boundary.argtypes = [vp, sz, C.c_char_p]
    L.c3control_add_obstacle.argtypes = [vp, vp, vp]
    L.c3control_set_device_model.argtypes = [vp, C.c_int, vp, sz]
    L.c3opt_alloc.argtypes = [C.c_int, sz]
    L.c3opt_set_brute_force_vals.argtypes = [vp, sz, vp]
    L.valuef_from_cores.argtypes = [sz, vp, vp, vp]
    L.control_params_create.argtypes = [sz, sz, vp, vp, vp, vp]
    L.vi_param_create.argtypes = [dbl]
    L.pi_param_create.argtypes = [dbl, vp]
    for n in ("vi_param_add_cp", "vi_param_add_value", "pi_param_add_cp", "pi_param_add_value"):
        getattr(L, n).argtypes = [vp, vp]
    L.bellman_vi.argtypes = [sz, vp, vp, vp]
    L.bellman_pi.argtypes = [sz, vp, vp, vp]
    L.bellman_vi_batch.argtypes = [sz, vp, vp, vp]
    L.bellman_optimal.argtypes = [sz, vp, vp, vp]
    L.bellman_control.argtypes = [sz, vp, vp, vp]
    L.bellman_control.restype = dbl
    L.bellmanrhs.restype = dbl
    L.bellmanrhs.argtypes = [sz, sz, dbl, vp, dbl, vp, vp, dbl, vp, vp, vp]
    L.transition_assemble.argtypes = [sz, sz, sz, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.mca_get_neighbor_costs.argtypes = [sz, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.valuef_eval_fiber_ind_nn.argtypes = [vp, vp, sz, vp, vp, vp]
    L.control_params_add_time_and_states.argtypes = [vp, dbl, sz, vp]
    L.workspace_get_costs.restype = vp
    L.workspace_get_absorbed.restype = vp
    L.workspace_get_costs.argtypes = [vp, sz]
    L.workspace_get_absorbed.argtypes = [vp, sz]
    for n in ("workspace_increment_pi_iter", "workspace_increment_pi_subiter", "workspace_increment_vi_iter",
              "workspace_reset_pi_prob_htable", "control_params_destroy", "vi_param_destroy", "pi_param_destroy",
              "valuef_destroy", "c3opt_free", "c3control_destroy"):
        getattr(L, n).argtypes = [vp]
    return L


class HostProblem:
    """what an examples/*.c main() does up to the solver loop"""

    def __init__(self, L, cfg, arith=1):
        self.L, self.cfg = L, cfg
        self.lb = np.ascontiguousarray(cfg.lb, np.float64); self.ub = np.ascontiguousarray(cfg.ub, np.float64)
        self.ngrid = np.ascontiguousarray(cfg.ngrid, np.uintp)
        self.c3c = L.c3control_create(cfg.dx, cfg.du, cfg.dw, po._p(self.lb), po._p(self.ub), po._p(self.ngrid), cfg.beta)
        for i in range(cfg.dx):
            if cfg.bc[i] != configs.ABSORB:
                L.c3control_set_external_boundary(self.c3c, i, BC_NAME[int(cfg.bc[i])])
        for o in range(cfg.obs_center.shape[0] if cfg.obs_center.size else 0):
            c = np.ascontiguousarray(cfg.obs_center[o]); w = np.ascontiguousarray(cfg.obs_width[o])
            L.c3control_add_obstacle(self.c3c, po._p(c), po._p(w))
        L.c3control_set_device_model(self.c3c, cfg.model, None, 0)
        L.dp_param_set_arith.argtypes = [vp, C.c_int]
        L.dp_param_set_arith(L.c3control_get_dp(self.c3c), arith)
        self.opt = L.c3opt_alloc(3, cfg.du)                        # BRUTEFORCE
        self.utab = np.ascontiguousarray(cfg.controls, np.float64)
        L.c3opt_set_brute_force_vals(self.opt, cfg.nu, po._p(self.utab))
        self.cp = L.control_params_create(cfg.dx, cfg.dw, L.c3control_get_dp(self.c3c), L.c3control_get_mca(self.c3c),
                                          L.c3control_get_work(self.c3c), self.opt)

    def valuef(self, ranks, cores):
        n = np.ascontiguousarray(self.cfg.ngrid, np.uintp); r = np.ascontiguousarray(ranks, np.uintp)
        cs = [np.ascontiguousarray(c, np.float64) for c in cores]
        arr = (vp * len(cs))(*[c.ctypes.data for c in cs])
        return self.L.valuef_from_cores(len(cs), po._p(n), po._p(r), arr)

    def close(self):
        self.L.control_params_destroy(self.cp); self.L.c3opt_free(self.opt); self.L.c3control_destroy(self.c3c)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,rank,dx", [("double_int", 24, 5, None), ("dubinscar_new", 16, 4, None),
                                             ("skidding5d", 10, 3, None), ("lqgnd_reflect", 8, 3, 6), ("user_vdp", 20, 3, None)])
def test_bellman_vi_and_pi_one_fiber_per_call(gpu, name, n, rank, dx):
    L = host_lib()
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    hp = HostProblem(L, cfg, arith=0)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    _, cores2, ft2 = make_ft(cfg, seed=0xABCD00)
    vf, vf2 = hp.valuef(ranks, cores), hp.valuef(ranks, cores2)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 24, face_frac=0.2)
    oval, _ = port.vi_batch(ft, dv, fi)
    vi = L.vi_param_create(1e-10)
    L.vi_param_add_cp(vi, hp.cp); L.vi_param_add_value(vi, vf)
    L.workspace_increment_vi_iter(L.c3control_get_work(hp.c3c))
    xs = []
    for f in range(len(dv)):
        x = port.fiber_points(dv[f], fi[f]); xs.append(x)
        out = np.zeros(cfg.n)
        assert L.bellman_vi(cfg.n, po._p(x), po._p(out), vi) == 0          # the reference's operator signature
        assert rel_err(out, oval[f, :cfg.n], scale=np.abs(oval).max()) <= 1e-12
    # the batched form: all fibers' points back to back, one launch
    xall = np.ascontiguousarray(np.concatenate(xs)); outall = np.zeros(len(dv) * cfg.n)
    assert L.bellman_vi_batch(len(dv), po._p(xall), po._p(outall), vi) == 0
    assert rel_err(outall.reshape(len(dv), cfg.n), oval[:, :cfg.n], scale=np.abs(oval).max()) <= 1e-12
    # policy evaluation: pi_solve protocol (bellman.c:2343-2356): new policy -> pi_iter++, rows dropped
    work = L.c3control_get_work(hp.c3c)
    poli = L.pi_param_create(1e-10, vf)
    L.workspace_increment_pi_iter(work); L.workspace_reset_pi_prob_htable(work)
    o1, rows, _ = port.pi_batch(ft, ft2, dv, fi)
    o2, _, _ = port.pi_batch(ft, ft, dv, fi, rows=rows)
    for vf_it, oref in ((vf2, o1), (vf, o2)):
        L.pi_param_add_cp(poli, hp.cp); L.pi_param_add_value(poli, vf_it); L.workspace_increment_pi_subiter(work)
        for f in range(len(dv)):
            out = np.zeros(cfg.n)
            assert L.bellman_pi(cfg.n, po._p(xs[f]), po._p(out), poli) == 0
            assert rel_err(out, oref[f, :cfg.n], scale=np.abs(oref).max()) <= 1e-12
    L.pi_param_destroy(poli); L.vi_param_destroy(vi); L.valuef_destroy(vf); L.valuef_destroy(vf2)
    hp.close()


@pytest.mark.gpu
def test_scalar_reference_entry_points(gpu):
    """transition_assemble, bellmanrhs, mca_get_neighbor_costs, valuef_eval_fiber_ind_nn,
    bellman_optimal, bellman_control with the reference's own signatures."""
    L = host_lib()
    cfg = configs.get_config("skidding5d", n=10, rank=3)
    hp = HostProblem(L, cfg, arith=0)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    vf = hp.valuef(ranks, cores)
    dx = cfg.dx
    # transition_assemble(dx,du,dw,h2,t,drift,NULL,ddiff,NULL,prob,NULL,&dt,NULL,NULL)
    drift = np.array([[1.0, -2.0, 3e-15, 0.5, -0.25]]); sig = np.array([[0.3, 0.0, 1.0, 0.2, 0.7]])
    dd = np.zeros(dx * dx); dd[np.arange(dx) * dx + np.arange(dx)] = sig[0]
    prob = np.zeros(2 * dx + 1); dt = dbl()
    rc = L.transition_assemble(dx, cfg.du, cfg.dw, port.p.h2, po._p(port.t), po._p(drift[0]), None, po._p(dd), None,
                               po._p(prob), None, C.byref(dt), None, None)
    op, odt, ost = port.transition(drift, sig)
    assert rc == ost[0] == 0 and np.array_equal(prob, op[0]) and dt.value == odt[0]
    cost = synthetic.uniform01(9, 2 * dx + 1)
    v = L.bellmanrhs(dx, cfg.du, 1.3, None, cfg.beta, po._p(prob), None, dt.value, None, po._p(cost), None)
    assert abs(v - port.L.orc_rhs(dx, 1.3, cfg.beta, po._p(prob), dt.value, po._p(cost))) <= 1e-15 * abs(v)
    # mca_get_neighbor_costs + valuef_eval_fiber_ind_nn
    k, fixed = 2, np.array([3, 0, 0, 4, 9], np.int32)
    x = port.fiber_points(k, fixed)
    xg = L.c3control_get_xgrid(hp.c3c)
    fi_out = np.zeros(dx, np.uintp); kk = sz(); ab = np.zeros(cfg.n, np.intc); costs = np.zeros((cfg.n, 2 * dx + 1))
    rc = L.mca_get_neighbor_costs(dx, cfg.n, po._p(x), L.c3control_get_boundary(hp.c3c), vf, po._p(hp.ngrid), xg,
                                  po._p(fi_out), C.byref(kk), po._p(ab), po._p(costs))
    oab, ocosts = port.neighbor_costs(ft, k, fixed)
    assert rc == 0 and kk.value == k and np.array_equal(ab, oab)
    assert rel_err(costs, ocosts, scale=np.abs(ocosts).max()) <= 1e-12
    _, nv, nf = port.fiber_neighbors(k, fixed)
    out = np.zeros((cfg.n, 2 * dx + 1))
    fiz = np.ascontiguousarray(fixed, np.uintp); nvz = np.ascontiguousarray(nv.reshape(-1), np.uintp); nfz = np.ascontiguousarray(nf.reshape(-1), np.uintp)
    assert L.valuef_eval_fiber_ind_nn(vf, po._p(fiz), k, po._p(nfz), po._p(nvz), po._p(out)) == 0
    assert rel_err(out, ocosts, scale=np.abs(ocosts).max()) <= 1e-12
    # bellman_optimal / bellman_control through struct Memory{shared, private}
    work = L.c3control_get_work(hp.c3c)
    L.control_params_add_time_and_states(hp.cp, 0.0, cfg.n, po._p(x))
    wc = np.ctypeslib.as_array(C.cast(L.workspace_get_costs(work, 0), C.POINTER(dbl)), shape=(cfg.n, 2 * dx + 1))
    wa = np.ctypeslib.as_array(C.cast(L.workspace_get_absorbed(work, 0), C.POINTER(C.c_int)), shape=(cfg.n,))
    wc[:] = ocosts; wa[:] = oab
    oval, oub, _, _ = port.vi_fiber_full(ft, k, fixed)

    class Mem(C.Structure):
        _fields_ = [("shared", vp), ("private_", sz)]
    for j in range(cfg.n):
        mem = Mem(hp.cp, j); u = np.zeros(cfg.du); val = dbl()
        assert L.bellman_optimal(cfg.du, po._p(u), C.byref(val), C.byref(mem)) == 0
        assert abs(val.value - oval[j]) <= 1e-12 * max(abs(oval[j]), np.abs(oval).max())
        if oub[j] >= 0:
            assert np.array_equal(u, cfg.controls[oub[j]])
            vv = L.bellman_control(cfg.du, po._p(np.ascontiguousarray(cfg.controls[oub[j]])), None, C.byref(mem))
            assert abs(vv - oval[j]) <= 1e-12 * max(abs(oval[j]), np.abs(oval).max())
    L.valuef_destroy(vf); hp.close()


def test_host_containers_without_gpu():
    """host logic only: grid constants, boundary boxes, fiber decode -- no compute calls."""
    L = host_lib()
    cfg = configs.get_config("dubinscar_new", n=11, rank=2)
    lb = np.ascontiguousarray(cfg.lb); ub = np.ascontiguousarray(cfg.ub); ng = np.ascontiguousarray(cfg.ngrid, np.uintp)
    c3c = L.c3control_create(cfg.dx, cfg.du, cfg.dw, po._p(lb), po._p(ub), po._p(ng), cfg.beta)
    xg = C.cast(L.c3control_get_xgrid(c3c), C.POINTER(C.POINTER(dbl)))
    for i in range(cfg.dx):
        g = np.ctypeslib.as_array(xg[i], shape=(11,))
        assert np.array_equal(g, configs.c3_linspace(cfg.lb[i], cfg.ub[i], 11))
    L.c3control_set_external_boundary(c3c, 2, b"periodic")
    b = L.c3control_get_boundary(c3c)
    L.boundary_type_dim.argtypes = [vp, sz, C.c_int]
    assert [L.boundary_type_dim(b, i, 0) for i in range(3)] == [1, 1, 2]
    c = np.zeros(3); w = np.array([0.5, 0.5, 2 * np.pi])
    L.c3control_add_obstacle(c3c, po._p(c), po._p(w))
    L.boundary_obstacle_get_lb.argtypes = [vp, sz]
    olb = np.ctypeslib.as_array(C.cast(L.boundary_obstacle_get_lb(b, 0), C.POINTER(dbl)), shape=(3,))
    assert np.array_equal(olb, c - w / 2.0)
    L.boundary_in_obstacle.argtypes = [vp, vp]
    assert L.boundary_in_obstacle(b, po._p(np.array([0.25, -0.25, 3.0]))) == 1
    assert L.boundary_in_obstacle(b, po._p(np.array([0.26, 0.0, 0.0]))) == 0
    # convert_fiber_to_ind
    port = make_port(cfg)
    x = port.fiber_points(1, np.array([4, 0, 7], np.int32))
    fi = np.zeros(3, np.uintp); kk = sz()
    L.convert_fiber_to_ind.argtypes = [sz, sz, vp, vp, vp, vp, vp]
    assert L.convert_fiber_to_ind(3, 11, po._p(x), po._p(ng), xg, po._p(fi), C.byref(kk)) == 0
    assert kk.value == 1 and list(fi) == [4, 0, 7]
    assert L.convert_fiber_to_ind(3, 10, po._p(x), po._p(ng), xg, po._p(fi), C.byref(kk)) == 2
    x[0, 0] += 1e-9
    assert L.convert_fiber_to_ind(3, 11, po._p(x), po._p(ng), xg, po._p(fi), C.byref(kk)) == 1
    # outer_bound_dim (src/boundary.c:577-597): only the periodic dimension maps, left face first
    L.outer_bound_dim.restype = dbl
    L.outer_bound_dim.argtypes = [vp, sz, dbl, C.POINTER(C.c_int)]
    mp = C.c_int(7)
    assert L.outer_bound_dim(b, 2, cfg.lb[2], C.byref(mp)) == cfg.ub[2] and mp.value == 1
    assert L.outer_bound_dim(b, 2, cfg.ub[2] + 0.5, C.byref(mp)) == cfg.lb[2] and mp.value == 2
    assert L.outer_bound_dim(b, 2, 0.25, C.byref(mp)) == 0.25 and mp.value == 0
    assert L.outer_bound_dim(b, 0, cfg.lb[0] - 1.0, C.byref(mp)) == cfg.lb[0] - 1.0 and mp.value == 0   # absorbing: no map
    # uniform_stride (src/util.c:995-1006)
    L.uniform_stride.restype = sz
    L.uniform_stride.argtypes = [sz, sz]
    assert [L.uniform_stride(100, m) for m in (2, 5, 20, 100)] == [98, 24, 5, 0]
    assert L.uniform_stride(60, 3) == 29 and L.uniform_stride(7, 1) == 0
    # dyn_copy_deep / dyn_init_ref (src/dynamics.c:279-312)
    for name in ("drift_alloc", "diff_alloc", "dyn_alloc", "dyn_copy_deep"):
        getattr(L, name).restype = vp
    L.drift_alloc.argtypes = [sz, sz]; L.diff_alloc.argtypes = [sz, sz, sz]; L.dyn_alloc.argtypes = [vp, vp]
    L.dyn_copy_deep.argtypes = [vp]; L.dyn_init_ref.argtypes = [vp, vp, vp]; L.dyn_free_deep.argtypes = [vp]; L.dyn_free.argtypes = [vp]
    for name in ("dyn_get_dx", "dyn_get_du", "dyn_get_dw"):
        getattr(L, name).restype = sz; getattr(L, name).argtypes = [vp]
    dyn = L.dyn_alloc(L.drift_alloc(3, 1), L.diff_alloc(3, 1, 3))
    cp = L.dyn_copy_deep(dyn)
    assert cp and cp != dyn and (L.dyn_get_dx(cp), L.dyn_get_du(cp), L.dyn_get_dw(cp)) == (3, 1, 3)
    assert L.dyn_copy_deep(None) is None
    shell = L.dyn_alloc(None, None)
    d5, f5 = L.drift_alloc(5, 2), L.diff_alloc(5, 2, 4)
    L.dyn_init_ref(shell, d5, f5)
    assert (L.dyn_get_dx(shell), L.dyn_get_du(shell), L.dyn_get_dw(shell)) == (5, 2, 4)
    L.dyn_free_deep(shell); L.dyn_free_deep(cp); L.dyn_free_deep(dyn)
    L.c3control_destroy(c3c)


# ---- the solver facade: ApproxArgs, valuef_interp, c3control_step_* / *_solve, policy_eval, Diag --------
def solver_lib():
    L = host_lib()
    for name in ("approx_args_init", "c3control_step_vi", "c3control_step_pi", "c3control_init_value", "c3control_vi_solve",
                 "c3control_pi_solve", "valuef_interp"):
        getattr(L, name).restype = vp
    for name in ("valuef_norm", "valuef_norm2diff", "valuef_eval", "approx_args_get_cross_tol", "approx_args_get_round_tol"):
        getattr(L, name).restype = dbl
    for name in ("approx_args_get_kickrank", "approx_args_get_maxrank", "approx_args_get_startrank"):
        getattr(L, name).restype = sz
        getattr(L, name).argtypes = [vp]
    L.approx_args_free.argtypes = [vp]
    L.approx_args_get_cross_tol.argtypes = [vp]; L.approx_args_get_round_tol.argtypes = [vp]; L.approx_args_get_adapt.argtypes = [vp]
    L.approx_args_set_cross_tol.argtypes = [vp, dbl]; L.approx_args_set_round_tol.argtypes = [vp, dbl]
    for name in ("approx_args_set_kickrank", "approx_args_set_maxrank", "approx_args_set_startrank"):
        getattr(L, name).argtypes = [vp, sz]
    L.approx_args_set_adapt.argtypes = [vp, C.c_int]
    L.valuef_norm.argtypes = [vp]; L.valuef_norm2diff.argtypes = [vp, vp]; L.valuef_eval.argtypes = [vp, vp]
    L.valuef_set_grid.argtypes = [vp, vp]
    L.c3control_step_vi.argtypes = [vp, vp, vp, vp, C.c_int, vp]
    L.c3control_step_pi.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    L.c3control_vi_solve.argtypes = [vp, sz, dbl, vp, vp, vp, C.c_int, vp]
    L.c3control_pi_solve.argtypes = [vp, sz, dbl, vp, vp, vp, C.c_int, vp]
    L.c3control_init_value.argtypes = [vp, vp, vp, vp, C.c_int]
    L.c3control_add_policy_sim.argtypes = [vp, vp, vp, vp]
    L.c3control_policy_eval.argtypes = [vp, dbl, vp, vp]
    L.c3control_controller.argtypes = [dbl, vp, vp, vp]
    L.diag_destroy.argtypes = [vp]; L.diag_save.argtypes = [vp, C.c_char_p]
    L.diag_append.argtypes = [vp, sz, C.c_int, dbl, dbl, sz, vp, dbl]
    return L


def _host_cores(L, vf, n):
    """(ranks, cores) of a host-mirror ValueF (struct: d, N*, ranks*, cores**)"""
    class VF(C.Structure):
        _fields_ = [("d", sz), ("N", C.POINTER(sz)), ("ranks", C.POINTER(sz)), ("cores", C.POINTER(C.POINTER(dbl)))]
    v = C.cast(vf, C.POINTER(VF)).contents
    d = v.d
    ranks = np.array([v.ranks[k] for k in range(d + 1)], dtype=np.uint64)
    cores = [np.ctypeslib.as_array(v.cores[k], shape=(int(n[k] * ranks[k] * ranks[k + 1]),)).copy() for k in range(d)]
    return ranks, cores


def test_approx_args_defaults_and_diag(tmp_path):
    """src/util.c:116-132 defaults; Diag list round trip (no device needed)"""
    L = solver_lib()
    a = L.approx_args_init()
    assert L.approx_args_get_cross_tol(a) == 1e-10 and L.approx_args_get_round_tol(a) == 1e-10
    assert (L.approx_args_get_kickrank(a), L.approx_args_get_startrank(a), L.approx_args_get_maxrank(a)) == (10, 5, 40)
    assert L.approx_args_get_adapt(a) == 1
    L.approx_args_set_kickrank(a, 3); L.approx_args_set_adapt(a, 0); L.approx_args_set_round_tol(a, 1e-6)
    assert L.approx_args_get_kickrank(a) == 3 and L.approx_args_get_adapt(a) == 0 and L.approx_args_get_round_tol(a) == 1e-6
    L.approx_args_free(a)
    head = vp()
    ranks = np.array([1, 4, 6, 1], dtype=np.uintp)
    for it in range(3):
        L.diag_append(C.byref(head), it, 1, 10.0 + it, 0.5 / (it + 1), 3, po._p(ranks), 0.25)
    f = tmp_path / "diag.txt"
    assert L.diag_save(head, str(f).encode()) == 0
    rows = f.read_text().strip().splitlines()
    assert len(rows) == 4 and rows[1].split()[:2] == ["1", "0"] and float(rows[3].split()[2]) == 12.0
    assert float(rows[1].split()[-1]) == 5.0                                  # mean interior rank
    L.diag_destroy(C.byref(head))
    assert not head.value


@pytest.mark.gpu
def test_vi_solve_fixed_rank_equals_the_cross_driver(gpu):
    """c3control_vi_solve with adapt = 0 is a fresh startrank cross of bellman_vi per iteration
    (src/valuefunc.c:713-733): same cores as driving include/c3sc_cross.h by hand"""
    L = solver_lib()
    cfg = configs.get_config("lqg2d_reflect", n=20, rank=4)
    hp = HostProblem(L, cfg, arith=1)
    prob = capi.Problem(cfg, arith=1)
    r0, c0 = synthetic.quadratic_cores(prob.xgrid)
    v0 = hp.valuef(r0, c0)
    a = L.approx_args_init()
    L.approx_args_set_adapt(a, 0); L.approx_args_set_startrank(a, 4); L.approx_args_set_cross_tol(a, 1e-12)
    head = vp()
    out = L.c3control_vi_solve(hp.c3c, 3, 0.0, v0, a, hp.opt, 0, C.byref(head))
    ranks, cores = _host_cores(L, out, cfg.ngrid)
    cg, rg = c0, np.asarray(r0, dtype=np.uint64)
    for _ in range(3):
        vf = capi.ValueF(cfg.ngrid, rg, cg)
        cr = capi.Cross(cfg.ngrid, [1, 4, 1])
        cg, _, _ = cr.run_vi(prob, vf, maxiter=5, tol=1e-12)
        rg = cr.ranks.copy(); vf.close(); cr.close()
    assert list(ranks) == list(rg)
    for x, y in zip(cores, cg):
        assert np.array_equal(x, y)
    L.valuef_norm_nodal.restype = dbl; L.valuef_norm_nodal.argtypes = [vp]
    assert abs(L.valuef_norm_nodal(out) - capi.cores_norm(cfg.ngrid, rg, cg)) <= 1e-12 * L.valuef_norm_nodal(out)
    # valuef_norm is the reference's quantity: the continuous L2 norm of the piecewise-linear train
    want = np.sqrt(capi.cores_dot_l2(cfg.ngrid, prob.xgrid, rg, cg, rg, cg))
    assert abs(L.valuef_norm(out) - want) <= 1e-12 * want
    assert L.valuef_norm2diff(out, v0) > 0
    f = C.create_string_buffer(b"/dev/null")
    assert L.diag_save(head, f) == 0
    L.diag_destroy(C.byref(head)); L.valuef_destroy(out); L.valuef_destroy(v0); L.approx_args_free(a)
    hp.close(); prob.close()


@pytest.mark.gpu
def test_adaptive_vi_solve_pi_solve_and_controller(gpu):
    """the example main() flow (examples/lqgnd/lqgnd.c:300-420): init value from a host start cost,
    value iteration with rank adaptation, a policy-iteration solve on top, then the online controller"""
    L = solver_lib()
    cfg = configs.get_config("lqgnd", n=12, rank=6, dx=4)
    hp = HostProblem(L, cfg, arith=1)
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    START = C.CFUNCTYPE(C.c_int, sz, C.POINTER(dbl), C.POINTER(dbl), vp)

    def _start(N, x, out, _):                                                # startcost of lqgnd.c:200-215: sum x_i^2
        xs = np.ctypeslib.as_array(x, shape=(N * cfg.dx,)).reshape(N, cfg.dx)
        np.ctypeslib.as_array(out, shape=(N,))[:] = (xs ** 2).sum(axis=1)
        return 0
    start = START(_start)
    a = L.approx_args_init()
    L.approx_args_set_startrank(a, 3); L.approx_args_set_kickrank(a, 2); L.approx_args_set_maxrank(a, 10)
    L.approx_args_set_round_tol(a, 1e-6); L.approx_args_set_cross_tol(a, 1e-8)
    v0 = L.c3control_init_value(hp.c3c, start, None, a, 0)
    r0, c0 = _host_cores(L, v0, cfg.ngrid)
    assert list(r0) == [1, 2, 2, 2, 1]                                        # sum of squares has TT rank 2
    pt = np.array([0.3, -0.7, 1.1, 0.05])
    assert abs(L.valuef_eval(v0, po._p(pt)) - port.ft_eval_linear(po.FT(cfg.ngrid, r0, c0), pt)) <= 1e-12
    head = vp()
    v1 = L.c3control_vi_solve(hp.c3c, 4, 1e-12, v0, a, hp.opt, 0, C.byref(head))
    r1, c1 = _host_cores(L, v1, cfg.ngrid)
    assert max(int(x) for x in r1) <= 10 and L.valuef_norm(v1) > 0
    # one more step by hand must move the function less than the first one did (contraction)
    nev = sz(0)
    v2 = L.c3control_step_vi(hp.c3c, v1, a, hp.opt, 0, C.byref(nev))
    assert nev.value > 0
    assert L.valuef_norm2diff(v2, v1) < L.valuef_norm2diff(v1, v0)
    # policy iteration on the policy of v2
    v3 = L.c3control_pi_solve(hp.c3c, 2, 1e-12, v2, a, hp.opt, 0, C.byref(head))
    assert np.isfinite(L.valuef_norm(v3)) and L.valuef_norm2diff(v3, v2) < 0.5 * L.valuef_norm(v2)
    # online controller: same control as the batched device entry at a few states
    L.c3control_add_policy_sim(hp.c3c, v3, hp.opt, None)
    r3, c3 = _host_cores(L, v3, cfg.ngrid)
    vf3 = capi.ValueF(cfg.ngrid, r3, c3)
    xs = np.array([[0.3, -0.7, 1.1, 0.05], [-1.2, 0.4, 0.0, 0.9], [1.9, 1.9, -1.9, 0.2]])
    want = prob.policy_eval(vf3, xs)
    for i, x in enumerate(xs):
        u = np.zeros(cfg.du)
        assert L.c3control_controller(0.0, po._p(np.ascontiguousarray(x)), po._p(u), hp.c3c) == 0
        assert np.array_equal(u, np.asarray(want[0])[i])
    for v in (v0, v1, v2, v3):
        L.valuef_destroy(v)
    L.diag_destroy(C.byref(head)); L.approx_args_free(a); vf3.close(); hp.close(); prob.close()


@pytest.mark.gpu
def test_pi_regression_norm_from_the_paper(gpu):
    """Test_bellman_pi_25 (tprob_test.c:1996-2357, "FROM PAPER!! This is a regression test"): 2-D LQG,
    reflecting boundaries, discount 0.1, 25 x 25 nodes, start value 0.2, outer loop pi_solve(10) + vi_solve(1)
    until the iterates differ by < 1e-5; the L2 norm of the value function is 100 within 10 %.  The
    reference minimises over u in [-1, 1] with BFGS; here the control set is {-1, 0, 1}."""
    L = solver_lib()
    N = 25
    cfg = configs.get_config("lqg2d_reflect", n=N, rank=3)
    hp = HostProblem(L, cfg, arith=1)
    START = C.CFUNCTYPE(C.c_int, sz, C.POINTER(dbl), C.POINTER(dbl), vp)

    def _quad2d(n, x, out, _):                                    # tprob_test.c:1389-1398
        np.ctypeslib.as_array(out, shape=(n,))[:] = 0.2
        return 0
    quad2d = START(_quad2d)
    a = L.approx_args_init()                                      # tprob_test.c:2025-2031
    L.approx_args_set_cross_tol(a, 1e-8); L.approx_args_set_round_tol(a, 1e-7); L.approx_args_set_kickrank(a, 5)
    L.approx_args_set_adapt(a, 0); L.approx_args_set_startrank(a, 3); L.approx_args_set_maxrank(a, 20)
    cost = L.c3control_init_value(hp.c3c, quad2d, None, a, 0)
    diff = 1.0
    for it in range(2000):
        nxt = L.c3control_pi_solve(hp.c3c, 10, 1e-5, cost, a, hp.opt, 0, None)
        tmp = L.c3control_vi_solve(hp.c3c, 1, 1e-5, nxt, a, hp.opt, 0, None)
        diff = L.valuef_norm2diff(nxt, tmp)
        L.valuef_destroy(nxt); L.valuef_destroy(cost)
        cost = tmp
        if diff < 1e-5:
            break
    assert diff < 1e-5, (it, diff)
    ranks, cores = _host_cores(L, cost, cfg.ngrid)
    g0 = np.asarray(cores[0]).reshape(N, int(ranks[1]))            # [j][b] (r_0 = 1)
    g1 = np.asarray(cores[1]).reshape(N, int(ranks[1]))            # [j][a] (r_2 = 1)
    V = g0 @ g1.T
    w = np.full(N, 4.0 / (N - 1)); w[0] = w[-1] = 2.0 / (N - 1)    # trapezoid rule on [-2, 2]
    l2 = float(np.sqrt(np.einsum("i,j,ij->", w, w, V * V)))
    assert abs(100.0 - l2) / 100.0 < 0.1, l2
    # valuef_norm IS that norm now (continuous L2 of the piecewise-linear train, as in tprob_test.c:2346-2348)
    assert abs(100.0 - L.valuef_norm(cost)) / 100.0 < 0.1 and abs(L.valuef_norm(cost) - l2) < 0.02 * l2
    assert V.min() > 0 and abs(V[N // 2, N // 2] - V.min()) < 0.05 * V.max()      # bowl centred at the origin
    L.valuef_destroy(cost); L.approx_args_free(a); hp.close()


@pytest.mark.gpu
def test_valuef_save_and_load_round_trip(gpu, tmp_path):
    """checkpoint / resume (src/valuefunc.c:226-295): binary and text files reproduce the function bit for
    bit on the same grid; on a finer grid the load re-samples the piecewise-linear cores; a missing file is
    NULL (the examples then start afresh)"""
    L = solver_lib()
    L.valuef_save.argtypes = [vp, C.c_char_p]; L.valuef_savetxt.argtypes = [vp, C.c_char_p]
    L.valuef_load.argtypes = [C.c_char_p, vp, vp]; L.valuef_load.restype = vp
    L.valuef_loadtxt.argtypes = [C.c_char_p, vp, vp]; L.valuef_loadtxt.restype = vp
    cfg = configs.get_config("dubinscar_new", n=12, rank=4)
    hp = HostProblem(L, cfg, arith=1)
    port = make_port(cfg)
    ranks = cfg.ranks()
    cores = synthetic.random_cores(cfg.ngrid, ranks, seed=31)
    vf = hp.valuef(ranks, cores)
    xg = [np.ascontiguousarray(g) for g in port.xgrid]
    garr = (vp * cfg.dx)(*[g.ctypes.data for g in xg])
    L.valuef_set_grid(vf, garr)
    ng = np.ascontiguousarray(cfg.ngrid, np.uintp)
    for save, load, name in ((L.valuef_save, L.valuef_load, "v.bin"), (L.valuef_savetxt, L.valuef_loadtxt, "v.txt")):
        f = str(tmp_path / name).encode()
        assert save(vf, f) == 0
        back = load(f, po._p(ng), garr)
        assert back
        r2, c2 = _host_cores(L, back, cfg.ngrid)
        assert list(r2) == [int(x) for x in ranks]
        for a, b in zip(cores, c2):
            assert np.array_equal(np.asarray(a).reshape(-1), b)
        assert L.valuef_norm2diff(vf, back) == 0.0
        L.valuef_destroy(back)
    assert not L.valuef_load(str(tmp_path / "missing.bin").encode(), po._p(ng), garr)
    # finer grid: every old node is a new node (2n-1 nodes), so the function agrees there and is linear between
    fine_n = np.array([2 * int(n) - 1 for n in cfg.ngrid], dtype=np.uintp)
    fine = [np.ascontiguousarray(np.interp(np.arange(2 * len(g) - 1) / 2.0, np.arange(len(g)), g)) for g in xg]
    farr = (vp * cfg.dx)(*[g.ctypes.data for g in fine])
    back = L.valuef_load(str(tmp_path / "v.bin").encode(), po._p(fine_n), farr)
    assert back
    ft = po.FT(cfg.ngrid, ranks, cores)
    for pt in (np.array([0.3, -1.1, 2.0]), np.array([xg[0][3], xg[1][5], xg[2][7]]), np.array([-2.2, 1.7, 0.4])):
        want = port.ft_eval_linear(ft, np.ascontiguousarray(pt))
        assert abs(L.valuef_eval(back, po._p(np.ascontiguousarray(pt))) - want) <= 1e-12 * max(1.0, abs(want))
    L.valuef_destroy(back); L.valuef_destroy(vf); hp.close()


# ---- reference names added in round 2: process_fibers_neighbor, mca_get_neighbor_node_costs,
# ---- valuef_get_isl, the per-node workspace slabs -------------------------------------------------------
def _geom_config(name, n, rank, dx):
    """a BASELINE config, or "geom7": a 7-dimensional grid with mixed boundary types and an obstacle but NO
    dynamics model (model id 0) -- for the entries of the path that are pure geometry + FT evaluation"""
    if name != "geom7":
        return configs.get_config(name, n=n, rank=rank, dx=dx)
    d = 7
    bc = np.array([configs.ABSORB, configs.REFLECT, configs.PERIODIC, configs.ABSORB, configs.REFLECT, configs.ABSORB, configs.PERIODIC], np.int32)
    return configs.Config("geom7", 0, d, 1, d, n, np.full(d, -1.0), np.linspace(1.0, 2.0, d), bc, 0.1, np.zeros((1, 1)), rank,
                          obs_center=np.full((1, d), 0.1), obs_width=np.full((1, d), 0.9))


def _r2_lib():
    L = solver_lib()
    L.process_fibers_neighbor.argtypes = [sz, vp, sz, vp, vp, vp, vp, vp, vp]
    L.mca_get_neighbor_node_costs.argtypes = [sz, vp, vp, vp, vp, vp, vp, vp]
    L.valuef_get_isl.restype = vp
    L.valuef_get_isl.argtypes = [vp]
    for n in ("workspace_get_drift", "workspace_get_diff", "workspace_get_dt", "workspace_get_prob", "workspace_get_u",
              "workspace_get_grad_drift", "workspace_get_grad_prob"):
        getattr(L, n).restype = vp
        getattr(L, n).argtypes = [vp, sz]
    return L


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,rank,dx", [("double_int", 24, 5, None), ("dubinscar_new", 16, 4, None),
                                             ("skidding5d", 10, 3, None), ("lqgnd_reflect", 8, 3, 6), ("geom7", 6, 2, 7)])
def test_process_fibers_neighbor_reference_signature(gpu, name, n, rank, dx):
    """process_fibers_neighbor (src/nodeutil.c:489-627) with the reference's own argument list (no grids: the
    coordinates come from x), every varying dimension, faces and obstacle fibers: bit-exact flags and indices.
    d = 7 has no device dynamics model and must still work (the entry is geometry only)."""
    L = _r2_lib()
    cfg = _geom_config(name, n, rank, dx)
    port = make_port(cfg)
    lb = np.ascontiguousarray(cfg.lb); ub = np.ascontiguousarray(cfg.ub); ng = np.ascontiguousarray(cfg.ngrid, np.uintp)
    L.c3control_create.restype = vp
    c3c = L.c3control_create(cfg.dx, cfg.du, cfg.dw, po._p(lb), po._p(ub), po._p(ng), cfg.beta)
    for i in range(cfg.dx):
        if cfg.bc[i] != configs.ABSORB:
            L.c3control_set_external_boundary(c3c, i, BC_NAME[int(cfg.bc[i])])
    for o in range(cfg.obs_center.shape[0] if cfg.obs_center.size else 0):
        L.c3control_add_obstacle(c3c, po._p(np.ascontiguousarray(cfg.obs_center[o])), po._p(np.ascontiguousarray(cfg.obs_width[o])))
    bound = L.c3control_get_boundary(c3c)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 3 * cfg.dx, face_frac=0.3)
    if cfg.obs_center.size:                                               # a fiber through the obstacle's centre
        fi[0] = [int(np.argmin(np.abs(port.xgrid[i] - cfg.obs_center[0][i]))) for i in range(cfg.dx)]
    for f in range(len(dv)):
        k = int(dv[f]); N = int(cfg.ngrid[k])
        x = port.fiber_points(k, fi[f])
        fz = np.ascontiguousarray(fi[f], np.uintp); fz[k] = 0
        ab = np.full(N, 9, np.intc); nv = np.zeros(2 * N, np.uintp); nf = np.zeros(max(2 * (cfg.dx - 1), 1), np.uintp)
        assert L.process_fibers_neighbor(cfg.dx, po._p(fz), k, po._p(x), po._p(ab), po._p(nv), po._p(nf), po._p(ng), bound) == 0
        oab, onv, onf = port.fiber_neighbors(k, fi[f])
        assert np.array_equal(ab, oab), (f, k)
        assert np.array_equal(nv.reshape(N, 2), onv), (f, k)
        assert np.array_equal(nf.reshape(-1, 2)[:cfg.dx - 1], onf[:cfg.dx - 1]), (f, k)
    L.c3control_destroy(c3c)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,rank,dx", [("dubinscar_new", 14, 4, None), ("skidding5d", 10, 3, None), ("geom7", 6, 2, 7)])
def test_mca_get_neighbor_costs_and_node_costs_are_model_free(gpu, name, n, rank, dx):
    """mca_get_neighbor_costs (src/nodeutil.c:647) and mca_get_neighbor_node_costs (:718) with the reference's
    signatures on a geometry-only device problem -- including d = 7, which has no instantiated dynamics model."""
    L = _r2_lib()
    cfg = _geom_config(name, n, rank, dx)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    lb = np.ascontiguousarray(cfg.lb); ub = np.ascontiguousarray(cfg.ub); ng = np.ascontiguousarray(cfg.ngrid, np.uintp)
    L.c3control_create.restype = vp
    c3c = L.c3control_create(cfg.dx, cfg.du, cfg.dw, po._p(lb), po._p(ub), po._p(ng), cfg.beta)
    for i in range(cfg.dx):
        if cfg.bc[i] != configs.ABSORB:
            L.c3control_set_external_boundary(c3c, i, BC_NAME[int(cfg.bc[i])])
    for o in range(cfg.obs_center.shape[0] if cfg.obs_center.size else 0):
        L.c3control_add_obstacle(c3c, po._p(np.ascontiguousarray(cfg.obs_center[o])), po._p(np.ascontiguousarray(cfg.obs_width[o])))
    bound, xg = L.c3control_get_boundary(c3c), L.c3control_get_xgrid(c3c)
    nn = np.ascontiguousarray(cfg.ngrid, np.uintp); rr = np.ascontiguousarray(ranks, np.uintp)
    cs = [np.ascontiguousarray(c, np.float64) for c in cores]
    arr = (vp * len(cs))(*[c.ctypes.data for c in cs])
    vf = L.valuef_from_cores(len(cs), po._p(nn), po._p(rr), arr)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 6, face_frac=0.3)
    for f in range(len(dv)):
        k = int(dv[f]); N = int(cfg.ngrid[k])
        x = port.fiber_points(k, fi[f])
        fo = np.zeros(cfg.dx, np.uintp); kk = sz(); ab = np.zeros(N, np.intc); costs = np.zeros((N, 2 * cfg.dx + 1))
        assert L.mca_get_neighbor_costs(cfg.dx, N, po._p(x), bound, vf, po._p(ng), xg, po._p(fo), C.byref(kk), po._p(ab), po._p(costs)) == 0
        oab, oc = port.neighbor_costs(ft, k, fi[f])
        assert kk.value == k and np.array_equal(ab, oab)
        assert rel_err(costs, oc, scale=np.abs(oc).max()) <= 1e-12
    u = synthetic.uniform01(5, 40 * cfg.dx).reshape(40, cfg.dx)
    pts = cfg.lb + u * (cfg.ub - cfg.lb)
    pts[::5] = cfg.lb; pts[1::5] = cfg.ub
    if cfg.obs_center.size:
        pts[2::7] = cfg.obs_center[0]
    for x in pts:
        x = np.ascontiguousarray(x)
        ab = C.c_int(5); out = np.zeros(2 * cfg.dx + 1)
        assert L.mca_get_neighbor_node_costs(cfg.dx, po._p(x), bound, vf, po._p(ng), xg, C.byref(ab), po._p(out)) == 0
        oab, oout = port.neighbor_node_costs(ft, x)
        assert ab.value == oab
        m = 2 * cfg.dx                                                    # slot 2dx is left unset by the reference
        assert rel_err(out[:m], oout[:m], scale=max(np.abs(oout[:m]).max(), 1e-300)) <= 1e-12
    L.valuef_destroy(vf); L.c3control_destroy(c3c)


@pytest.mark.gpu
def test_workspace_slabs_after_bellman_control(gpu):
    """workspace_get_{drift,diff,dt,prob,u} (src/util.c:876-906): after bellman_control(node, u) the node's slab holds
    what the reference leaves there (src/bellman.c:400-449) -- checked against the oracle's transition row"""
    L = _r2_lib()
    cfg = configs.get_config("skidding5d", n=10, rank=3)
    hp = HostProblem(L, cfg, arith=0)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    dx = cfg.dx
    k, fixed = 1, np.array([3, 0, 2, 4, 6], np.int32)
    x = port.fiber_points(k, fixed)
    oab, ocosts = port.neighbor_costs(ft, k, fixed)
    work = L.c3control_get_work(hp.c3c)
    L.control_params_add_time_and_states(hp.cp, 0.0, cfg.n, po._p(x))
    wc = np.ctypeslib.as_array(C.cast(L.workspace_get_costs(work, 0), C.POINTER(dbl)), shape=(cfg.n, 2 * dx + 1))
    wa = np.ctypeslib.as_array(C.cast(L.workspace_get_absorbed(work, 0), C.POINTER(C.c_int)), shape=(cfg.n,))
    wc[:] = ocosts; wa[:] = 0

    class Mem(C.Structure):
        _fields_ = [("shared", vp), ("private_", sz)]
    for j in (1, 4, 8):
        for cand in (0, cfg.nu // 2, cfg.nu - 1):
            u = np.ascontiguousarray(cfg.controls[cand])
            mem = Mem(hp.cp, j)
            L.bellman_control(cfg.du, po._p(u), None, C.byref(mem))
            od, osg, *_ = port.model_eval(x[j:j + 1], u[None, :])
            op, odt, ost = port.transition(od, osg)
            drift = np.ctypeslib.as_array(C.cast(L.workspace_get_drift(work, j), C.POINTER(dbl)), shape=(dx,))
            diff = np.ctypeslib.as_array(C.cast(L.workspace_get_diff(work, j), C.POINTER(dbl)), shape=(dx, cfg.dw))
            dt = C.cast(L.workspace_get_dt(work, j), C.POINTER(dbl))[0]
            prob = np.ctypeslib.as_array(C.cast(L.workspace_get_prob(work, j), C.POINTER(dbl)), shape=(2 * dx + 1,))
            assert rel_err(drift, od[0], scale=1.0) <= 1e-14
            assert np.array_equal(np.diag(diff), osg[0]) and np.count_nonzero(diff - np.diag(np.diag(diff))) == 0
            assert rel_err(prob[:-1], op[0][:-1], scale=1e-3) <= 1e-13 and abs(prob[-1] - op[0][-1]) <= 1e-15
            assert abs(dt - odt[0]) <= 1e-14 * odt[0]
        uo = np.zeros(cfg.du); val = dbl()
        mem = Mem(hp.cp, j)
        assert L.bellman_optimal(cfg.du, po._p(uo), C.byref(val), C.byref(mem)) == 0
        wu = np.ctypeslib.as_array(C.cast(L.workspace_get_u(work, j), C.POINTER(dbl)), shape=(cfg.du,))
        assert np.array_equal(wu, uo)
    # slab layout of src/util.c:738-748: grad_drift follows drift, grad_prob follows prob
    assert L.workspace_get_grad_drift(work, 2) - L.workspace_get_drift(work, 2) == 8 * dx
    assert L.workspace_get_grad_prob(work, 2) - L.workspace_get_prob(work, 2) == 8 * (2 * dx + 1)
    hp.close()


@pytest.mark.gpu
def test_valuef_get_isl_after_a_cross_run(gpu):
    """valuef_get_isl (src/valuefunc.c:218): the left index sets of the cross run that produced the value function,
    r_k multi-indices over dimensions 0..k-1, equal to c3sc_cross_index_sets of the same run"""
    L = _r2_lib()
    cfg = configs.get_config("lqgnd", n=10, rank=3, dx=4)
    hp = HostProblem(L, cfg, arith=1)
    prob = capi.Problem(cfg, arith=1)
    r0, c0 = synthetic.quadratic_cores(prob.xgrid)
    v0 = hp.valuef(r0, c0)
    assert not L.valuef_get_isl(v0)                                         # not produced by a cross run
    a = L.approx_args_init()
    L.approx_args_set_adapt(a, 0); L.approx_args_set_startrank(a, 3); L.approx_args_set_cross_tol(a, 1e-10)
    nev = sz(0)
    v1 = L.c3control_step_vi(hp.c3c, v0, a, hp.opt, 0, C.byref(nev))
    ranks, _ = _host_cores(L, v1, cfg.ngrid)

    class CI(C.Structure):
        _fields_ = [("d", sz), ("n", sz), ("inds", C.POINTER(sz)), ("vals", C.POINTER(dbl))]
    isl = C.cast(L.valuef_get_isl(v1), C.POINTER(C.POINTER(CI)))
    assert isl
    for k in range(cfg.dx):
        ci = isl[k].contents
        assert ci.d == k and ci.n == int(ranks[k])
        for a_ in range(ci.n):
            for i in range(k):
                idx = ci.inds[a_ * k + i]
                assert idx < cfg.ngrid[i]
                if ci.vals:
                    assert ci.vals[a_ * k + i] == prob.xgrid[i][idx]
        if k >= 1:                                                          # multi-indices of a set are distinct
            rows = {tuple(ci.inds[a_ * k + i] for i in range(k)) for a_ in range(ci.n)}
            assert len(rows) == ci.n
    L.valuef_destroy(v0); L.valuef_destroy(v1); L.approx_args_free(a); hp.close(); prob.close()
