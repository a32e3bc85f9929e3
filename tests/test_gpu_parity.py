"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded
inputs.  Bars (BASELINE.json north_star / SURVEY.md A8):
  bit-exact   : absorbed flags, neighbour indices, transition-stencil structure, argmin index
                (argmin: except ties within tolerance)
  <= 1e-12 rel: neighbour costs, transition probabilities, dt, backed-up values
                (relative to max(|ref|, largest magnitude in the same row/fiber) so entries
                 that cancel to ~0 are judged against the terms that produced them)
"""
import numpy as np
import pytest

from c3sc_b200 import capi, configs, synthetic
from oracle import pyoracle as po
from helpers import (SMALL, make_ft, make_port, rel_err, valid_mask, elem_err_costs, elem_err_values,
                     argmin_mismatches_are_ties)

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _argmin_ok(cfg, port, ft, dv, fi, arg, oarg, oval):
    """identical argmin, or the two candidates' oracle values tie within RTOL"""
    bad = np.argwhere(arg != oarg)
    for f, j in bad:
        if j >= cfg.ngrid[dv[f]]:
            continue
        out, ub, ab, costs = port.vi_fiber_full(ft, dv[f], fi[f])
        x = port.fiber_points(dv[f], fi[f])[j]
        import ctypes as C
        vals = []
        for cand in (arg[f, j], oarg[f, j]):
            prob = np.zeros(2 * cfg.dx + 1); dt = C.c_double(); g = C.c_double(); st = C.c_int()
            port.L.orc_control_value.restype = C.c_double
            u = np.ascontiguousarray(cfg.controls[cand])
            v = port.L.orc_control_value(C.byref(port.p), po._p(np.ascontiguousarray(x)), po._p(u), po._p(np.ascontiguousarray(costs[j])),
                                         po._p(prob), C.byref(dt), C.byref(g), C.byref(st))
            vals.append(v)
        assert abs(vals[0] - vals[1]) <= RTOL * max(abs(vals[1]), np.abs(oval[f]).max()), (f, j, vals)
    return len(bad)


@pytest.mark.parametrize("arith", [0, 1], ids=["exact", "fast"])
@pytest.mark.parametrize("name,n,rank,dx", SMALL)
def test_vi_debug_against_oracle(gpu, name, n, rank, dx, arith):
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=arith)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 96, face_frac=0.25)
    out = prob.vi_batch_debug(vf, dv, fi)
    oval, oarg = port.vi_batch(ft, dv, fi)
    m = valid_mask(cfg, dv)
    for f in range(len(dv)):
        N = int(cfg.ngrid[dv[f]])
        ab, nv, nf = port.fiber_neighbors(dv[f], fi[f])
        assert np.array_equal(out["absorbed"][f, :N], ab), f                      # bit-exact
        assert np.array_equal(out["nbr_vary"][f, :N], nv), f
        if cfg.dx > 1:
            assert np.array_equal(out["nbr_fixed"][f, :cfg.dx - 1], nf), f
        _, costs = port.neighbor_costs(ft, dv[f], fi[f])
        assert rel_err(out["costs"][f, :N], costs, scale=np.abs(costs).max()) <= RTOL, f
    assert rel_err(out["value"][m], oval[m], scale=np.abs(oval[m]).max()) <= RTOL
    nbad = argmin_mismatches_are_ties(cfg, port, ft, dv, fi, out["argmin"], oarg)
    assert nbad <= 0.01 * m.sum()
    # per-element errors (scale = sum|terms| of the element) beside the batch-max figure above
    for f in range(0, len(dv), 7):
        N = int(cfg.ngrid[dv[f]])
        ec, _, scale = elem_err_costs(port, ft, dv[f], fi[f], out["costs"][f, :N])
        assert ec <= RTOL, (f, ec)
        assert elem_err_values(out["value"][f, :N], oval[f, :N], scale) <= RTOL, f
    prob.close(); vf.close()


@pytest.mark.parametrize("name,n,rank,dx", SMALL)
def test_policy_rows_bit_exact_in_exact_mode(gpu, name, n, rank, dx):
    """EXACT arithmetic reproduces transition_assemble's operation order: p and dt at the
    argmin are bit-identical to the oracle's policy rows wherever the drift itself is
    (models without sin/cos); with sin/cos they agree to 1e-14."""
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=0)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 64, face_frac=0.2)
    out = prob.vi_batch_debug(vf, dv, fi)
    _, orows, oarg = port.pi_batch(ft, ft, dv, fi)
    sel = (out["argmin"] == oarg) & (oarg >= 0) & valid_mask(cfg, dv)
    assert sel.sum() > 0
    g, o = out["rows"][sel], orows[sel]
    if cfg.model in (configs.MODEL_LQGND, configs.MODEL_DOUBLE_INT, configs.MODEL_USER):
        assert np.array_equal(g, o)
    else:
        assert rel_err(g[:, :-3], o[:, :-3], scale=1e-3) <= 1e-13      # probabilities
        assert np.abs(g[:, -3] - o[:, -3]).max() <= 1e-15              # p_self: round-off, absolute
        assert rel_err(g[:, -2:], o[:, -2:]) <= 1e-14                  # dt, g
    # stencil structure: which side received the drift term
    assert np.array_equal(g[:, :-3:2] > g[:, 1:-3:2], o[:, :-3:2] > o[:, 1:-3:2])
    prob.close(); vf.close()


@pytest.mark.parametrize("arith", [0, 1], ids=["exact", "fast"])
@pytest.mark.parametrize("name,n,rank,dx", SMALL)
def test_pi_two_subiterations(gpu, name, n, rank, dx, arith):
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=arith)
    port = make_port(cfg)
    ranks, c_pol, ft_pol = make_ft(cfg)
    _, c_it, ft_it = make_ft(cfg, seed=0xABCD00)
    _, c_it2, ft_it2 = make_ft(cfg, seed=0x777700)
    vf_pol = capi.ValueF(cfg.ngrid, ranks, c_pol)
    vf_it = capi.ValueF(cfg.ngrid, ranks, c_it)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 48, face_frac=0.2)
    m = valid_mask(cfg, dv)
    v1, rows, arg = prob.pi_batch(vf_pol, vf_it, dv, fi)
    o1, orows, oarg = port.pi_batch(ft_pol, ft_it, dv, fi)
    assert rel_err(v1[m], o1[m]) <= RTOL
    argmin_mismatches_are_ties(cfg, port, ft_pol, dv, fi, arg, oarg)
    vf_it.update(c_it2)
    v2, _, _ = prob.pi_batch(None, vf_it, dv, fi, rows=rows)
    o2, _, _ = port.pi_batch(ft_pol, ft_it2, dv, fi, rows=orows)
    assert rel_err(v2[m], o2[m]) <= RTOL
    prob.close(); vf_pol.close(); vf_it.close()


@pytest.mark.parametrize("arith", [0, 1], ids=["exact", "fast"])
@pytest.mark.parametrize("dx", [2, 3, 5, 10])
def test_transition_probabilities(gpu, dx, arith):
    """tprob_test.c:327-364: probabilities >= -1e-15 and sum to 1 +- 1e-15; plus parity."""
    name = {2: "lqg2d_new", 3: "dubinscar_new", 5: "skidding5d", 10: "lqgnd"}[dx]
    cfg = configs.get_config(name, n=12, rank=2)
    prob = capi.Problem(cfg, arith=arith)
    port = make_port(cfg)
    n = 4000
    drift = (synthetic.uniform01(11, n * dx).reshape(n, dx) - 0.5) * 8.0
    drift[::7, 0] = 0.0
    drift[::11, dx - 1] = 5e-15                 # inside the 1e-14 dead band
    drift[::13, dx - 1] = -2e-14                # just outside
    sig = synthetic.uniform01(12, n * dx).reshape(n, dx) * 2.0
    sig[::5, 0] = 0.0
    p, dt, st = prob.transition(drift, sig)
    op, odt, ost = port.transition(drift, sig)
    assert np.array_equal(st, ost)
    ok = ost == 0
    assert (p[ok] >= -1e-15).all() and np.abs(p[ok].sum(axis=1) - 1.0).max() <= 1e-15
    if arith == 0:
        assert np.array_equal(p[ok], op[ok]) and np.array_equal(dt[ok], odt[ok])      # bit-exact
    else:
        assert rel_err(p[ok][:, :-1], op[ok][:, :-1], scale=1e-300) <= 1e-14
        assert np.abs(p[ok][:, -1] - op[ok][:, -1]).max() <= 1e-15
        assert rel_err(dt[ok], odt[ok], scale=1e-300) <= 1e-14
    prob.close()


@pytest.mark.parametrize("name,n,rank,dx", SMALL)
def test_device_model_equals_host_callbacks(gpu, name, n, rank, dx):
    """SURVEY §7 hard part 1: the device-resident model must equal the host callback on the grid."""
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=0)
    port = make_port(cfg)
    ne = 3000
    u01 = synthetic.uniform01(5, ne * cfg.dx).reshape(ne, cfg.dx)
    x = cfg.lb + u01 * (cfg.ub - cfg.lb)
    for i in range(min(ne, cfg.n)):                       # include true grid nodes
        x[i] = [prob.xgrid[d][(i * (d + 3)) % cfg.n] for d in range(cfg.dx)]
    u = cfg.controls[np.arange(ne) % cfg.nu]
    g = prob.model_eval(x, u)
    o = port.model_eval(x, u)
    if cfg.model in (configs.MODEL_LQGND, configs.MODEL_DOUBLE_INT, configs.MODEL_USER):
        for a, b in zip(g, o):
            assert np.array_equal(a, b)
    else:
        assert rel_err(g[0], o[0], scale=1.0) <= 4e-16 * 30      # sin/cos: <= 2 ulp of |drift| <= 27+10
        for a, b in zip(g[1:], o[1:]):
            assert np.array_equal(a, b)
    prob.close()


def test_full_size_configs(gpu):
    """BASELINE.json full sizes: a few hundred fibers each against the oracle, plus
    size-independent properties on a larger batch (idempotence, absorbed <-> boundcost,
    PI(policy=V, iter=V) == VI(V))."""
    for name in configs.ALL_CONFIGS:
        cfg = configs.get_config(name)
        prob = capi.Problem(cfg, arith=1)
        port = make_port(cfg)
        ranks, cores, ft = make_ft(cfg)
        vf = capi.ValueF(cfg.ngrid, ranks, cores)
        F = 40 if name == "lqgnd" else 200
        dv, fi = synthetic.random_fibers(cfg.ngrid, F)
        val, arg = prob.vi_batch(vf, dv, fi)
        oval, oarg = port.vi_batch(ft, dv, fi)
        assert rel_err(val, oval) <= RTOL, name
        assert argmin_mismatches_are_ties(cfg, port, ft, dv, fi, arg, oarg) <= 1e-3 * arg.size, name
        # properties on a bigger batch
        dv, fi = synthetic.random_fibers(cfg.ngrid, 2000, seed=99)
        v1, a1 = prob.vi_batch(vf, dv, fi)
        v2, a2 = prob.vi_batch(vf, dv, fi)
        assert np.array_equal(v1, v2) and np.array_equal(a1, a2), "not deterministic"
        p1, rows, pa = prob.pi_batch(vf, vf, dv, fi)
        assert np.array_equal(pa, a1)
        assert rel_err(p1, v1) <= 1e-13
        prob.close(); vf.close()


def test_quadratic_value_function_lqgnd(gpu):
    """structured case of SURVEY §8(d): exact rank-2 FT of sum x_i^2 on the d=10 grid."""
    cfg = configs.get_config("lqgnd_reflect", n=30)
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    ranks, cores = synthetic.quadratic_cores(prob.xgrid)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    ft = po.FT(cfg.ngrid, ranks, cores)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 30)
    out = prob.vi_batch_debug(vf, dv, fi)
    # neighbour costs are sums of squares of the neighbour coordinates
    for f in range(len(dv)):
        _, costs = port.neighbor_costs(ft, dv[f], fi[f])
        assert rel_err(out["costs"][f], costs) <= RTOL
        x = port.fiber_points(dv[f], fi[f])
        assert rel_err(out["costs"][f][:, -1], (x * x).sum(axis=1)) <= 1e-13
    oval, oarg = port.vi_batch(ft, dv, fi)
    assert rel_err(out["value"], oval) <= RTOL
    prob.close(); vf.close()


def test_errors_are_loud(gpu):
    cfg = configs.get_config("lqg2d_new", n=10, rank=2)
    prob = capi.Problem(cfg)
    ranks, cores, ft = make_ft(cfg)
    bad = configs.get_config("lqg2d_new", n=11, rank=2)
    r2, c2, _ = make_ft(bad)
    vf_bad = capi.ValueF(bad.ngrid, r2, c2)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 4)
    with pytest.raises(capi.C3scError):
        prob.vi_batch(vf_bad, dv, fi)
    prob.close(); vf_bad.close()


def test_fiber_descriptors_outside_the_grid_are_refused(gpu):
    """c3sc_fibers_check / the debug entry: the batch analogue of convert_fiber_to_ind's non-zero
    returns (src/nodeutil.c:437-470) -- a descriptor that names no grid node is an error, not a read
    outside the cores."""
    cfg = configs.get_config("dubinscar_new", n=12, rank=3)
    prob = capi.Problem(cfg)
    ranks, cores, ft = make_ft(cfg)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 9)
    prob.fibers_check(dv, fi)
    prob.fibers_check(dv[:0], fi[:0])                                  # empty batch
    for f, (bad_dv, slot, bad_fi) in enumerate([(3, None, None), (-1, None, None), (None, 1, 12), (None, 2, -1)]):
        dv2, fi2 = dv.copy(), fi.copy()
        if bad_dv is not None:
            dv2[4 + f] = bad_dv
        else:
            fi2[4 + f, slot] = bad_fi
        with pytest.raises(capi.C3scError, match=f"fiber {4 + f}"):
            prob.fibers_check(dv2, fi2)
        with pytest.raises(capi.C3scError, match=f"fiber {4 + f}"):
            prob.vi_batch_debug(vf, dv2, fi2)
    out = prob.vi_batch_debug(vf, dv, fi)                              # the refusals left the problem usable
    port = po.Port(cfg, prob.xgrid, prob.h2, prob.t, prob.obs_lb, prob.obs_ub)
    oval, _ = port.vi_batch(ft, dv, fi)
    assert rel_err(out["value"], oval) <= RTOL
    prob.close(); vf.close()


import os as _os
_GOLD = _os.path.join(_os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("arith", [0, 1], ids=["exact", "fast"])
@pytest.mark.parametrize("fname", sorted(f for f in _os.listdir(_GOLD) if f.endswith(".npz")) if _os.path.isdir(_GOLD) else [])
def test_cuda_path_against_golden_fixtures(gpu, fname, arith):
    """the CUDA path directly against the committed outputs of the reference's own object code
    (tests/golden/make_golden.py): flags and neighbour indices bit-exact, neighbour values,
    bellman_vi and both bellman_pi sub-iterations within 1e-12"""
    z = np.load(_os.path.join(_GOLD, fname))
    name, n, rank, dx = str(z["name"]), int(z["n"]), int(z["rank"]), int(z["dx"])
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx if name.startswith("lqgnd") else None)
    prob = capi.Problem(cfg, arith=arith)
    ranks, cores, _ = make_ft(cfg)
    _, cores2, _ = make_ft(cfg, seed=0xABCD00)
    vf, vf2 = capi.ValueF(cfg.ngrid, ranks, cores), capi.ValueF(cfg.ngrid, ranks, cores2)
    dv, fi = z["dim_vary"], z["fixed_ind"]
    m = valid_mask(cfg, dv)
    out = prob.vi_batch_debug(vf, dv, fi)
    assert np.array_equal(out["absorbed"][m], z["absorbed"][m])
    assert np.array_equal(out["nbr_vary"][m], z["nbr_vary"][m])
    if cfg.dx > 1:
        assert np.array_equal(out["nbr_fixed"][:, :cfg.dx - 1].reshape(len(dv), -1), z["nbr_fixed"].reshape(len(dv), -1))
    assert rel_err(out["costs"][m], z["costs"][m], scale=np.abs(z["costs"][m]).max()) <= RTOL
    assert rel_err(out["value"][m], z["vi"][m], scale=np.abs(z["vi"][m]).max()) <= RTOL
    p1, rows, _ = prob.pi_batch(vf, vf2, dv, fi)
    p2, _, _ = prob.pi_batch(None, vf, dv, fi, rows=rows)
    assert rel_err(p1[m], z["pi1"][m], scale=np.abs(z["pi1"][m]).max()) <= RTOL
    assert rel_err(p2[m], z["pi2"][m], scale=np.abs(z["pi2"][m]).max()) <= RTOL
    prob.close(); vf.close(); vf2.close()


def test_timed_workload_full_size_against_oracle(gpu, capsys):
    """The workload bench.py times -- lqgnd_reflect, d = 10, N = 100, r = 20, 243 controls -- at a batch large
    enough for the pipeline the bench runs (two lanes, several chunks: >= 2*148*1024 nodes), all ten varying
    dimensions, faces included: values and argmin of all 4096 fibers against the oracle; flags, neighbour
    indices and neighbour values of a 320-fiber subset through the debug entry.  Errors are judged per
    ELEMENT against sum|terms| of that element, and the batch-max figure is printed beside it."""
    cfg = configs.get_config("lqgnd_reflect")
    assert (cfg.dx, cfg.n, cfg.rank, cfg.nu) == (10, 100, 20, 243)
    prob = capi.Problem(cfg, arith=1)
    assert capi.lib().c3sc_problem_control_path(prob.handle) == 2            # the grid walk, as in the bench
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    F = 4096
    assert F * cfg.n >= 2 * 148 * 1024
    dv, fi = synthetic.random_fibers(cfg.ngrid, F)
    assert set(dv.tolist()) == set(range(10))
    val, arg = prob.vi_batch(vf, dv, fi)                                   # the timed entry point
    oval, oarg = port.vi_batch(ft, dv, fi)
    # debug entry (every intermediate) on a subset that covers every dimension and the face fibers
    sub = np.arange(0, F, F // 320)[:320]
    dbg = prob.vi_batch_debug(vf, dv[sub], fi[sub])
    worst_c = worst_v = 0.0
    for q, f in enumerate(sub):
        ab, nv, nf = port.fiber_neighbors(dv[f], fi[f])
        assert np.array_equal(dbg["absorbed"][q], ab), f
        assert np.array_equal(dbg["nbr_vary"][q], nv), f
        assert np.array_equal(dbg["nbr_fixed"][q, :cfg.dx - 1], nf), f
        ec, ocost, scale = elem_err_costs(port, ft, dv[f], fi[f], dbg["costs"][q])
        worst_c = max(worst_c, ec)
        worst_v = max(worst_v, elem_err_values(val[f], oval[f], scale), elem_err_values(dbg["value"][q], oval[f], scale))
    batch_rel = rel_err(val, oval, scale=np.abs(oval).max())
    # per-element figure over ALL fibers with the node's own magnitude as scale (no batch max involved)
    own = float((np.abs(val - oval) / np.maximum(np.abs(oval), 1e-300)).max())
    nties = argmin_mismatches_are_ties(cfg, port, ft, dv, fi, arg, oarg)
    with capsys.disabled():
        print(f"\n[full size lqgnd_reflect F={F}] costs per-element {worst_c:.2e}; values per-element (sum|terms| scale) "
              f"{worst_v:.2e}, (own magnitude) {own:.2e}, (batch max scale) {batch_rel:.2e}; argmin mismatches {nties} (all ties)")
    assert worst_c <= RTOL and worst_v <= RTOL and batch_rel <= RTOL
    assert own <= 1e-9                               # a value that cancels to ~0 may lose digits; nothing does here
    assert nties <= 1e-4 * F * cfg.n
    # bellman_pi at the same size: improvement against vf, evaluation against a second train, then a sub-iteration
    _, cores2, ft2 = make_ft(cfg, seed=0xABCD00)
    vf2 = capi.ValueF(cfg.ngrid, ranks, cores2)
    p1, rows, parg = prob.pi_batch(vf, vf2, dv, fi)
    o1, orows, opa = port.pi_batch(ft, ft2, dv, fi)
    assert rel_err(p1, o1, scale=np.abs(o1).max()) <= RTOL
    assert argmin_mismatches_are_ties(cfg, port, ft, dv, fi, parg, opa) <= 1e-4 * F * cfg.n
    p2, _, _ = prob.pi_batch(None, vf, dv, fi, rows=rows)
    same = parg == opa
    o2, _, _ = port.pi_batch(ft, ft, dv, fi, rows=orows)
    assert rel_err(p2[same], o2[same], scale=np.abs(o2).max()) <= RTOL
    assert np.array_equal(parg, arg)                                       # same argmin as value iteration on vf
    prob.close(); vf.close(); vf2.close()
