"""The host cross driver (include/c3sc_cross.h): CPU tests of the driver itself on analytic
tensors, and the end-to-end parity bar of BASELINE.json -- value iteration through the SAME driver
with the GPU operator and with the CPU oracle must agree to 1e-10 (sup norm, held-out points)."""
import numpy as np
import pytest

from c3sc_b200 import capi, configs, synthetic
from oracle import pyoracle as po
from helpers import make_port


def _tt_full(n, ranks, cores):
    """dense tensor of a train in ValueF::cores layout"""
    d = len(n)
    out = np.ones((1, 1))
    for k in range(d):
        g = np.asarray(cores[k]).reshape(int(n[k]), int(ranks[k + 1]), int(ranks[k])).transpose(2, 0, 1)   # [a][j][b]
        out = np.tensordot(out, g, axes=([-1], [0]))
    return out.reshape([int(x) for x in n])


def test_cross_recovers_low_rank_tensor():
    """a tensor of exact TT rank 3 is reproduced to round-off with ranks >= 3"""
    n = [9, 7, 8, 6]
    grids = [np.linspace(-1, 1, m) for m in n]
    X = np.meshgrid(*grids, indexing="ij")
    full = np.sin(X[0] + X[1] + X[2] + X[3]) + 0.3 * X[0] * X[3]          # rank <= 2 + 1

    def fn(dv, fi):
        F = len(dv)
        out = np.zeros((F, max(n)))
        for f in range(F):
            idx = [int(v) for v in fi[f]]
            k = int(dv[f])
            for j in range(n[k]):
                idx[k] = j
                out[f, j] = full[tuple(idx)]
        return out
    cr = capi.Cross(n, [1, 4, 4, 4, 1])
    cores, nfib, change = cr.run(fn, maxiter=3)
    approx = _tt_full(n, cr.ranks, cores)
    assert np.abs(approx - full).max() <= 1e-11 * np.abs(full).max()
    assert nfib == 3 * 2 * (4 + 16 + 16 + 4)
    assert change < 1e-6          # the norm-difference formula bottoms out at sqrt(eps)
    cr.close()


def test_cross_batches_are_whole_cores():
    """every operator call asks for r_k * r_{k+1} fibers that all vary the same dimension"""
    n = [6, 5, 7]
    seen = []

    def fn(dv, fi):
        seen.append((len(dv), set(int(v) for v in dv)))
        return np.ones((len(dv), max(n)))
    cr = capi.Cross(n, [1, 3, 2, 1])
    cr.run(fn, maxiter=1)
    sizes = {0: 3, 1: 6, 2: 2}
    assert all(len(ks) == 1 and F == sizes[next(iter(ks))] for F, ks in seen)
    cr.close()


def test_rank_is_clipped_to_unfolding():
    cr = capi.Cross([3, 50, 50], [1, 10, 10, 1])
    assert list(cr.ranks) == [1, 3, 10, 1]
    cr.close()


def test_core_norms_against_dense():
    """valuef_norm / valuef_norm2diff stand-ins: discrete l2 of trains with different ranks"""
    n = [5, 6, 4]
    ra, rb = [1, 3, 2, 1], [1, 2, 4, 1]
    ca = synthetic.random_cores(np.array(n, dtype=np.uint64), np.array(ra, dtype=np.uint64), seed=3)
    cb = synthetic.random_cores(np.array(n, dtype=np.uint64), np.array(rb, dtype=np.uint64), seed=4)
    A, B = _tt_full(n, ra, ca), _tt_full(n, rb, cb)
    assert abs(capi.cores_norm(n, ra, ca) - np.linalg.norm(A)) <= 1e-13 * np.linalg.norm(A)
    assert abs(capi.cores_norm2diff(n, ra, ca, rb, cb) - np.linalg.norm(A - B)) <= 1e-12 * np.linalg.norm(A - B)


@pytest.mark.gpu
def test_vi_solve_converges_like_the_python_loop(gpu):
    """c3sc_vi_solve == the step-by-step loop over c3sc_cross_run_vi (same driver state evolution)"""
    cfg = configs.get_config("lqg2d_reflect", n=20, rank=5)
    prob = capi.Problem(cfg, arith=1)
    ranks0, cores0 = synthetic.quadratic_cores(prob.xgrid)
    cr1, cr2 = capi.Cross(cfg.ngrid, cfg.ranks()), capi.Cross(cfg.ngrid, cfg.ranks())
    cores, iters, diff, nfib = cr1.vi_solve(prob, ranks0, cores0, maxiter=5, abs_conv_tol=0.0, sweeps=2)
    assert iters == 5 and nfib > 0 and np.isfinite(diff)
    cg, rg = cores0, np.asarray(ranks0, dtype=np.uint64)
    for _ in range(5):
        vf = capi.ValueF(cfg.ngrid, rg, cg)
        prev, rprev = cg, rg
        cg, _, _ = cr2.run_vi(prob, vf, maxiter=2)
        rg = cr2.ranks
        vf.close()
    for a, b in zip(cores, cg):
        assert np.array_equal(a, b)
    assert abs(diff - capi.cores_norm2diff(cfg.ngrid, rprev, prev, rg, cg)) <= 1e-12 * max(diff, 1e-300)
    prob.close(); cr1.close(); cr2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,rank,dx,iters", [("lqg2d_reflect", 24, 6, None, 6), ("lqgnd", 12, 4, 4, 4),
                                                    ("dubinscar_new", 14, 5, None, 4)])
def test_value_iteration_gpu_equals_oracle(gpu, name, n, rank, dx, iters):
    """north_star: the value function after K value-iteration steps matches the CPU reference
    restatement within 1e-10 in sup norm on a held-out (off-grid) point set."""
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    ranks, cores0 = synthetic.quadratic_cores(prob.xgrid) if cfg.model == configs.MODEL_LQGND else (cfg.ranks(), synthetic.random_cores(cfg.ngrid, cfg.ranks()))
    cr_g = capi.Cross(cfg.ngrid, cfg.ranks())
    cr_o = capi.Cross(cfg.ngrid, cfg.ranks())
    cg, co = cores0, cores0
    rg = ro = np.asarray(ranks, dtype=np.uint64)
    vf = None
    for it in range(iters):
        if vf is not None:
            vf.close()
        vf = capi.ValueF(cfg.ngrid, rg, cg)
        cg, _, _ = cr_g.run_vi(prob, vf, maxiter=2)
        rg = cr_g.ranks
        ft_o = po.FT(cfg.ngrid, ro, co)
        co, _, _ = cr_o.run(lambda dv, fi: port.vi_batch(ft_o, dv, fi)[0], maxiter=2)
        ro = cr_o.ranks
    # held-out grid: random off-grid points, piecewise-linear FT evaluation (valuef_eval, src/valuefunc.c:345-350)
    ft_g, ft_o = po.FT(cfg.ngrid, rg, cg), po.FT(cfg.ngrid, ro, co)
    u = synthetic.uniform01(77, 400 * cfg.dx).reshape(400, cfg.dx)
    pts = cfg.lb + u * (cfg.ub - cfg.lb)
    vg = np.array([port.ft_eval_linear(ft_g, np.ascontiguousarray(x)) for x in pts])
    vo = np.array([port.ft_eval_linear(ft_o, np.ascontiguousarray(x)) for x in pts])
    scale = max(np.abs(vo).max(), 1.0)
    assert np.abs(vg - vo).max() <= 1e-10 * scale, np.abs(vg - vo).max() / scale
    prob.close(); vf.close(); cr_g.close(); cr_o.close()


@pytest.mark.gpu
def test_policy_iteration_step_gpu_equals_oracle(gpu):
    """c3control_step_pi through the driver: policy from one train, evaluation against another"""
    cfg = configs.get_config("lqgnd", n=12, rank=4, dx=4)
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    ranks = cfg.ranks()
    c_pol = synthetic.random_cores(cfg.ngrid, ranks, seed=11)
    c_it = synthetic.random_cores(cfg.ngrid, ranks, seed=12)
    vf_pol, vf_it = capi.ValueF(cfg.ngrid, ranks, c_pol), capi.ValueF(cfg.ngrid, ranks, c_it)
    ft_pol, ft_it = po.FT(cfg.ngrid, ranks, c_pol), po.FT(cfg.ngrid, ranks, c_it)
    cr_g, cr_o = capi.Cross(cfg.ngrid, ranks), capi.Cross(cfg.ngrid, ranks)
    cg, nf_g, _ = cr_g.run_pi(prob, vf_pol, vf_it, maxiter=2)
    co, nf_o, _ = cr_o.run(lambda dv, fi: port.pi_batch(ft_pol, ft_it, dv, fi)[0], maxiter=2)
    assert nf_g == nf_o
    ft_g, ft_o = po.FT(cfg.ngrid, cr_g.ranks, cg), po.FT(cfg.ngrid, cr_o.ranks, co)
    u = synthetic.uniform01(5, 300 * cfg.dx).reshape(300, cfg.dx)
    pts = cfg.lb + u * (cfg.ub - cfg.lb)
    vg = np.array([port.ft_eval_linear(ft_g, np.ascontiguousarray(x)) for x in pts])
    vo = np.array([port.ft_eval_linear(ft_o, np.ascontiguousarray(x)) for x in pts])
    assert np.abs(vg - vo).max() <= 1e-10 * max(np.abs(vo).max(), 1.0)
    prob.close(); vf_pol.close(); vf_it.close(); cr_g.close(); cr_o.close()
