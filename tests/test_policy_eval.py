"""the implicit policy at off-grid states (c3control_policy_eval).
CPU: the oracle restatement of mca_get_neighbor_node_costs against the reference's own object
code; GPU: the C-ABI entries against the oracle."""
import os

import numpy as np
import pytest

from c3sc_b200 import capi, configs, synthetic
from oracle import pyoracle as po
from helpers import make_ft, make_port, rel_err

CASES = [("lqg2d_new", 12, 3, None), ("dubinscar_new", 14, 4, None), ("skidding5d", 10, 3, None), ("lqgnd_reflect", 8, 3, 4)]


def _states(cfg, n=240, seed=3):
    """inside, on / next to every face, outside by a little, and (when there is one) inside the obstacle"""
    u = synthetic.uniform01(seed, n * cfg.dx).reshape(n, cfg.dx)
    pts = cfg.lb + u * (cfg.ub - cfg.lb)
    h = (cfg.ub - cfg.lb) / (cfg.n - 1)
    pts[::7] = cfg.lb + 0.3 * h
    pts[1::7] = cfg.ub - 0.3 * h
    pts[2::11] = cfg.lb
    pts[3::11] = cfg.ub
    if cfg.obs_center.size:
        pts[4::9] = cfg.obs_center[0]
    return pts


@pytest.mark.skipif(not os.path.isdir("/root/reference/src") and not po.have_ref(), reason="reference objects not built")
@pytest.mark.parametrize("name,n,rank,dx", CASES)
def test_oracle_node_costs_equal_reference_objects(name, n, rank, dx):
    """pins orc_neighbor_node_costs bit for bit to nodeutil.c:718-816 compiled from the reference
    (valuef_eval served by the oracle's piecewise-linear FT evaluation on both sides)"""
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    ref = po.Ref(cfg)
    vf = ref.valuef(ft)
    for x in _states(cfg):
        a1, o1 = port.neighbor_node_costs(ft, x)
        a2, o2 = ref.neighbor_node_costs(vf, x)
        assert a1 == a2
        assert np.array_equal(o1[:2 * cfg.dx], o2[:2 * cfg.dx])


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,rank,dx", CASES)
def test_policy_eval_gpu_equals_oracle(gpu, name, n, rank, dx):
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    pts = _states(cfg)
    # valuef_eval
    got = prob.valuef_eval(vf, pts)
    want = np.array([port.ft_eval_linear(ft, np.ascontiguousarray(x)) for x in pts])
    assert rel_err(got, want, scale=np.abs(want).max()) <= 1e-12
    out = cfg.ub + 0.5
    assert prob.valuef_eval(vf, out[None, :])[0] == 0.0                    # outside the grid
    # policy
    u, val, ab, costs = prob.policy_eval(vf, pts)
    for e, x in enumerate(pts):
        ou, oval, oab, ocosts = port.policy_eval(ft, x)
        assert ab[e] == oab                                                 # bit-exact flag
        assert rel_err(costs[e], ocosts, scale=np.abs(ocosts).max()) <= 1e-12
        assert rel_err(val[e], oval, scale=max(abs(oval), np.abs(ocosts).max())) <= 1e-12
    assert (np.abs(u - np.array([port.policy_eval(ft, x)[0] for x in pts])).max(axis=1) == 0).mean() > 0.98
    prob.close(); vf.close()
