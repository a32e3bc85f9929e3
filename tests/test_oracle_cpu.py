"""CPU tests: the oracle restatement against (a) the reference's own objects in oracle/_ref
(when present), (b) the committed golden fixtures generated from that build, and (c) the
known-answer properties the reference's tests hold for this path."""
import os

import numpy as np
import pytest

from c3sc_b200 import configs, synthetic
from oracle import pyoracle as po
from helpers import SMALL, host_problem, make_ft, make_port, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("name,n,rank,dx", SMALL)
def test_port_equals_reference_objects(built, name, n, rank, dx):
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    ref = po.Ref(cfg)
    xg, h, hmin, h2, t, olb, oub = host_problem(cfg)
    for a, b in zip(xg, ref.grids()):
        assert np.array_equal(a, b)                      # c3_linspace == C3 linspace stand-in
    rh, rhmin, rh2, rt = ref.constants()
    assert np.array_equal(h, rh) and hmin == rhmin and h2 == rh2 and np.array_equal(t, rt)
    rlb, rub = ref.obstacles()
    assert np.array_equal(olb, rlb) and np.array_equal(oub, rub)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    vf = ref.valuef(ft)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 48, face_frac=0.3)
    for f in range(len(dv)):
        for a, b in zip(port.fiber_neighbors(dv[f], fi[f]), ref.fiber_neighbors(dv[f], fi[f])):
            assert np.array_equal(a, b)
        for a, b in zip(port.neighbor_costs(ft, dv[f], fi[f]), ref.neighbor_costs(vf, dv[f], fi[f])):
            assert np.array_equal(a, b)
    o, _ = port.vi_batch(ft, dv, fi, nthreads=2)
    r, _ = ref.vi_fibers(vf, dv, fi)
    assert np.array_equal(o, r)                          # bellman_vi, bit for bit
    _, c2, ft2 = make_ft(cfg, seed=0xABCD00)
    vf2 = ref.valuef(ft2)
    ref.pi_begin(vf)
    r1, _ = ref.pi_fibers(vf2, dv, fi)
    r2, _ = ref.pi_fibers(vf, dv, fi)
    p1, rows, _ = port.pi_batch(ft, ft2, dv, fi)
    p2, _, _ = port.pi_batch(ft, ft, dv, fi, rows=rows)
    assert np.array_equal(p1, r1) and np.array_equal(p2, r2)    # bellman_pi, first + later sub-iteration
    ref.close()


@needs_ref
def test_transition_and_rhs_equal_reference(built):
    cfg = configs.get_config("skidding5d", n=10, rank=2)
    ref = po.Ref(cfg); port = make_port(cfg)
    n, dx = 500, cfg.dx
    drift = (synthetic.uniform01(3, n * dx).reshape(n, dx) - 0.5) * 6.0
    drift[::9, 2] = 3e-15
    sig = synthetic.uniform01(4, n * dx).reshape(n, dx)
    a = port.transition(drift, sig); b = ref.transition(drift, sig)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    cost = synthetic.uniform01(6, 2 * dx + 1)
    for e in range(0, n, 50):
        assert port.L.orc_rhs(dx, 1.3, cfg.beta, po._p(a[0][e]), a[1][e], po._p(cost)) == ref.rhs(1.3, cfg.beta, a[0][e], a[1][e], cost)
    ref.close()


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLD) if f.endswith(".npz")) if os.path.isdir(GOLD) else [])
def test_port_against_golden_fixtures(built, fname):
    """fixtures = outputs of the reference objects (tests/golden/make_golden.py); no /root/reference needed"""
    z = np.load(os.path.join(GOLD, fname))
    name, n, rank, dx = str(z["name"]), int(z["n"]), int(z["rank"]), int(z["dx"])
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx if name.startswith("lqgnd") else None)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    dv, fi = z["dim_vary"], z["fixed_ind"]
    for f in range(len(dv)):
        ab, nv, nf = port.fiber_neighbors(dv[f], fi[f])
        N = int(cfg.ngrid[dv[f]])
        assert np.array_equal(ab, z["absorbed"][f, :N]) and np.array_equal(nv, z["nbr_vary"][f, :N])
        assert np.array_equal(nf, z["nbr_fixed"][f])
        _, costs = port.neighbor_costs(ft, dv[f], fi[f])
        assert np.array_equal(costs, z["costs"][f, :N])
    o, _ = port.vi_batch(ft, dv, fi, nthreads=2)
    assert np.array_equal(o, z["vi"])
    _, c2, ft2 = make_ft(cfg, seed=0xABCD00)
    p1, rows, _ = port.pi_batch(ft, ft2, dv, fi)
    p2, _, _ = port.pi_batch(ft, ft, dv, fi, rows=rows)
    assert np.array_equal(p1, z["pi1"]) and np.array_equal(p2, z["pi2"])


def test_tprob_probsum_known_answer(built):
    """Test_tprob_probsum (tprob_test.c:327-364): 2-D, h={0.1,0.01}, pt=(-2,-0.3),
    drift f1 = (x1, u - x0 ... ) style inputs -> probs >= -1e-15, sum to 1 +- 1e-15."""
    cfg = configs.get_config("lqg2d_new", n=10, rank=2)
    port = make_port(cfg)
    h = np.array([0.1, 0.01]); hmin = h.min(); h2 = hmin * hmin
    port.t[:] = [h2 / h[0], h2 / h[0] / h[0], h2 / h[1], h2 / h[1] / h[1]]
    port.p.h2 = h2
    n = 2000
    drift = (synthetic.uniform01(21, 2 * n).reshape(n, 2) - 0.5) * 10
    sig = synthetic.uniform01(22, 2 * n).reshape(n, 2) * 3
    p, dt, st = port.transition(drift, sig)
    ok = st == 0
    assert ok.all()
    assert (p >= -1e-15).all()
    assert np.abs(p.sum(axis=1) - 1.0).max() <= 1e-15
    assert (dt > 0).all()


def test_absorbed_flags_full_grid_with_obstacle(built):
    """Test_process_fibers_neighbor (tprob_test.c:1068-1169): every fiber of a 3-D grid with a
    centred box obstacle, default ABSORB: flags in {-1,0,1}, faces -> 1, inside box -> -1."""
    cfg = configs.get_config("dubinscar_new", n=13, rank=2)
    cfg.bc[:] = configs.ABSORB
    cfg.obs_width = np.array([[0.8 * 8, 0.8 * 8, 0.8 * 2 * np.pi]]) / 2
    port = make_port(cfg)
    N = cfg.n
    for k in range(3):
        others = [i for i in range(3) if i != k]
        for a in range(N):
            for b in range(N):
                fi = np.zeros(3, np.int32); fi[others[0]] = a; fi[others[1]] = b
                ab, nv, nf = port.fiber_neighbors(k, fi)
                x = port.fiber_points(k, fi)
                on_face = a in (0, N - 1) or b in (0, N - 1)
                inside = np.all((x >= port.obs_lb) & (x <= port.obs_ub), axis=1)
                exp = np.where(inside, -1, 0)
                if on_face:
                    exp[:] = 1
                exp[0] = exp[-1] = 1
                assert np.array_equal(ab, exp)
                inter = np.arange(1, N - 1)
                free = ab[inter] == 0
                assert np.array_equal(nv[inter][free], np.stack([inter - 1, inter + 1], 1)[free])
                assert np.array_equal(nv[inter][~free], np.stack([inter, inter], 1)[~free])


def test_neighbor_eval_matches_pointwise_ft(built):
    """Test_valuef_neighbor_eval (tprob_test.c:535-919): fiber/neighbour values equal the
    pointwise FT evaluation at the neighbour's grid point to 1e-14 (d=3, N={30,43,24}, rank 20
    in the reference; same structure here on a uniform-N grid)."""
    cfg = configs.get_config("dubinscar_new", n=24, rank=20)
    cfg.bc[:] = configs.REFLECT
    cfg.obs_center = np.zeros((0, 0)); cfg.obs_width = np.zeros((0, 0))
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    for k in range(3):
        fi = np.array([3, 5, 9], np.int32)
        ab, nv, nf = port.fiber_neighbors(k, fi)
        _, costs = port.neighbor_costs(ft, k, fi)
        x = port.fiber_points(k, fi)
        fixed_dims = [i for i in range(3) if i != k]
        for j in range(0, cfg.n, 5):
            assert abs(costs[j, 6] - port.ft_eval_linear(ft, x[j])) <= 1e-14
            for s, i in enumerate(fixed_dims):
                for side in range(2):
                    xx = x[j].copy(); xx[i] = port.xgrid[i][nf[s, side]]
                    assert abs(costs[j, 2 * i + side] - port.ft_eval_linear(ft, xx)) <= 1e-14
            for side in range(2):
                xx = x[j].copy(); xx[k] = port.xgrid[k][nv[j, side]]
                assert abs(costs[j, 2 * k + side] - port.ft_eval_linear(ft, xx)) <= 1e-14


def test_fiber_decode(built):
    """Test_valuef_fiber_to_ind (tprob_test.c:921-961): x -> (fixed_ind, dim_vary)."""
    import ctypes as C
    cfg = configs.get_config("skidding5d", n=9, rank=2)
    port = make_port(cfg)
    for k in range(5):
        fi = np.array([4, 0, 8, 3, 7], np.int32)
        x = port.fiber_points(k, fi)
        got = np.zeros(5, np.uintp); kk = C.c_size_t()
        rc = port.L.orc_fiber_to_ind(C.c_size_t(5), C.c_size_t(9), po._p(x), po._p(port.ngrid),
                                     C.cast(port._xg, C.c_void_p), po._p(got), C.byref(kk))
        assert rc == 0 and kk.value == k
        exp = fi.copy(); exp[k] = 0
        assert np.array_equal(got, exp)


def test_c_abi_exports_every_declared_symbol():
    """the C-ABI library loads on a CPU-only box and exports what include/c3sc_b200.h declares"""
    import re
    from c3sc_b200 import capi
    L = capi.lib()
    inc = os.path.join(os.path.dirname(GOLD), "..", "include")
    hdr = "".join(open(os.path.join(inc, h)).read() for h in ("c3sc_b200.h", "c3sc_cross.h", "c3sc_multi.h"))
    declared = set(re.findall(r"\b(c3sc_[a-z_0-9]+)\s*\(", hdr)) - {"c3sc_fiber_batch_fn"}
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.c3sc_version().startswith(b"c3sc_b200")


def test_no_cpu_fallback_without_device():
    from c3sc_b200 import capi
    L = capi.lib()
    if L.c3sc_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.C3scError, match="no CUDA device"):
        capi.Problem(configs.get_config("lqg2d_new", n=8, rank=2))
    assert L.c3sc_cuda_init(0) == 3
