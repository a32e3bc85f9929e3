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
                                             ("skidding5d", 10, 3, None), ("lqgnd_reflect", 8, 3, 6)])
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
    L.c3control_destroy(c3c)
