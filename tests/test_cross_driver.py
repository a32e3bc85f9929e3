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


def test_fiber_memo_computes_every_distinct_fiber_once():
    """c3sc_fiber_memo (the reference's memo tables at fiber level, src/bellman.c:1334-1349): behind the memo the operator sees
    every distinct (dim_vary, fixed indices) once -- repeats inside a batch and in later sweeps are served from the store --
    and the driver produces bit-identical cores"""
    n = [9, 7, 8, 6]
    grids = [np.linspace(-1, 1, m) for m in n]
    X = np.meshgrid(*grids, indexing="ij")
    full = np.sin(X[0] + X[1] + X[2] + X[3]) + 0.3 * X[0] * X[3]
    seen = []

    def fn(dv, fi):
        F = len(dv)
        out = np.zeros((F, max(n)))
        for f in range(F):
            idx = [int(v) for v in fi[f]]
            k = int(dv[f])
            idx[k] = 0
            seen.append((k, tuple(idx)))
            for j in range(n[k]):
                idx[k] = j
                out[f, j] = full[tuple(idx)]
        return out
    cr = capi.Cross(n, [1, 4, 4, 4, 1])
    cores0, nfib0, _ = cr.run(fn, maxiter=3)
    plain_calls = len(seen)
    cr.close()
    seen.clear()
    cr = capi.Cross(n, [1, 4, 4, 4, 1])
    cores1, nfib1, _ = cr.run(fn, maxiter=3, memo=True)
    requested, computed = cr.memo_stats
    cr.close()
    assert nfib1 == nfib0 == plain_calls == requested
    assert computed == len(seen) == len(set(seen))            # nothing reached the operator twice
    assert computed < requested                               # ... and the sweeps of a converged cross do repeat fibers
    for a, b in zip(cores0, cores1):
        assert np.array_equal(a, b)


def test_pivoting_team_size_does_not_change_the_numbers(monkeypatch):
    """the pivoting step (twin rows + QR + maxvol) runs on a team of threads over fixed row blocks: cores, index sets and the
    change norm are bit-identical for 1, 3 and 8 threads (C3SC_HOST_THREADS), on unfoldings big enough to use the team"""
    n = [64, 48, 64, 56]
    grids = [np.linspace(-1, 1, m) for m in n]

    def fn(dv, fi):
        k = int(dv[0])
        assert (dv == k).all()
        x = [grids[i][fi[:, i]][:, None] for i in range(4)]
        x[k] = grids[k][None, :]
        out = np.zeros((len(dv), max(n)))
        sm = x[0] + x[1] + x[2] + x[3]
        out[:, :n[k]] = np.sin(2.0 * sm) + 1.0 / (5.0 + sm) + np.exp(-(x[0] ** 2 + x[1] ** 2 + x[2] ** 2 + x[3] ** 2))
        return out
    results = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("C3SC_HOST_THREADS", threads)
        cr = capi.Cross(n, [1, 14, 16, 14, 1])
        cores, nfib, change = cr.run(fn, maxiter=2)
        sets = [cr.index_sets(k) for k in range(1, 4)] if hasattr(cr, "index_sets") else []
        results.append((cores, change, sets))
        cr.close()
    for cores, change, sets in results[1:]:
        assert change == results[0][1]
        for a, b in zip(cores, results[0][0]):
            assert np.array_equal(np.asarray(a), np.asarray(b))
        for a, b in zip(sets, results[0][2]):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # and the answer is a good approximation (the function has modest TT ranks)
    X = np.meshgrid(*grids, indexing="ij")
    sm = X[0] + X[1] + X[2] + X[3]
    full = np.sin(2.0 * sm) + 1.0 / (5.0 + sm) + np.exp(-(X[0] ** 2 + X[1] ** 2 + X[2] ** 2 + X[3] ** 2))
    approx = _tt_full(n, [1, 14, 16, 14, 1], results[0][0])
    assert np.abs(approx - full).max() <= 1e-8 * np.abs(full).max()


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


# ---- rank adaptation: TT rounding + kicked ranks (the adapt == 1 branch of valuef_interp) -------------
def _u64(x):
    return np.array(x, dtype=np.uint64)


def test_round_keeps_a_full_rank_train():
    n = [6, 7, 5, 6]
    r = [1, 4, 5, 3, 1]
    c = synthetic.random_cores(_u64(n), _u64(r), seed=9)
    rr, cc = capi.cores_round(n, r, c, 1e-14)
    assert list(rr) == r
    A = _tt_full(n, r, c)
    assert np.abs(_tt_full(n, rr, cc) - A).max() <= 1e-13 * np.abs(A).max()


def test_round_finds_the_exact_rank():
    """a rank-2 function written with padded rank-5/6 cores (sum of a train with itself and zeros)"""
    n = [8, 9, 7, 8, 6]
    grids = [np.linspace(-1, 1, m) for m in n]
    r2, c2 = synthetic.quadratic_cores(grids)                            # sum x_i^2: exact rank 2
    d = len(n)
    big = [1] + [5, 6, 5, 6] + [1]
    noise = synthetic.random_cores(_u64(n), _u64(big), seed=21)
    cores = []
    for k in range(d):
        g = np.zeros((n[k], big[k + 1], big[k]))                        # [j][b][a] (block j column-major)
        src = np.asarray(c2[k]).reshape(n[k], int(r2[k + 1]), int(r2[k]))
        g[:, :src.shape[1], :src.shape[2]] = src
        if 0 < k < d - 1:                                               # a block the neighbours never reach
            g[:, 2:, 2:] = np.asarray(noise[k]).reshape(n[k], big[k + 1], big[k])[:, 2:, 2:]
        cores.append(g.reshape(-1))
    # make the junk unreachable: the first core feeds only the first two ranks
    A = _tt_full(n, r2, c2)
    assert np.abs(_tt_full(n, big, cores) - A).max() <= 1e-12 * np.abs(A).max()
    rr, cc = capi.cores_round(n, big, cores, 1e-10)
    assert list(rr) == [1, 2, 2, 2, 2, 1]
    assert np.abs(_tt_full(n, rr, cc) - A).max() <= 1e-9 * np.abs(A).max()


@pytest.mark.parametrize("eps", [1e-2, 1e-4, 1e-7])
def test_round_error_is_below_eps(eps):
    """|A - round(A, eps)| <= eps |A| (Frobenius), ranks non-increasing in eps; decaying singular values"""
    n = [10, 9, 11, 8]
    grids = [np.linspace(0, 1, m) for m in n]
    X = np.meshgrid(*grids, indexing="ij")
    full = 1.0 / (1.0 + X[0] + 2 * X[1] + 3 * X[2] + 0.5 * X[3])

    def fn(dv, fi):
        out = np.zeros((len(dv), max(n)))
        for f in range(len(dv)):
            idx = [int(v) for v in fi[f]]; k = int(dv[f])
            for j in range(n[k]):
                idx[k] = j
                out[f, j] = full[tuple(idx)]
        return out
    cr = capi.Cross(n, [1, 8, 8, 8, 1])
    cores, _, _ = cr.run(fn, maxiter=3)
    A = _tt_full(n, cr.ranks, cores)
    rr, cc = capi.cores_round(n, cr.ranks, cores, eps)
    err = np.linalg.norm(_tt_full(n, rr, cc) - A) / np.linalg.norm(A)
    assert err <= eps
    assert all(int(a) <= int(b) for a, b in zip(rr, cr.ranks))
    assert capi.cores_norm2diff(n, cr.ranks, cores, rr, cc) <= 1.01 * eps * capi.cores_norm(n, cr.ranks, cores) + 1e-12
    cr.close()


def test_round_handles_wide_unfoldings_and_bad_ranks():
    n = [2, 3, 9, 3, 2]
    r = [1, 2, 6, 6, 2, 1]
    c = synthetic.random_cores(_u64(n), _u64(r), seed=2)
    A = _tt_full(n, r, c)
    rr, cc = capi.cores_round(n, r, c, 1e-13)
    assert np.abs(_tt_full(n, rr, cc) - A).max() <= 1e-12 * np.abs(A).max()
    with pytest.raises(capi.C3scError):
        capi.cores_round([2, 9, 9], [1, 5, 3, 1], synthetic.random_cores(_u64([2, 9, 9]), _u64([1, 5, 3, 1])), 1e-8)


def test_adaptive_cross_grows_to_the_needed_rank():
    """start at rank 2 on a rank-4 tensor: kicks until the rounding cuts every bond, returns rank 4"""
    n = [9, 8, 10, 7]
    grids = [np.linspace(-1, 1, m) for m in n]
    X = np.meshgrid(*grids, indexing="ij")
    full = np.sin(X[0] + X[1] + X[2] + X[3]) + np.cos(2 * (X[0] - X[1] + X[2] - X[3]))     # TT rank 4
    calls = []

    def fn(dv, fi):
        calls.append(len(dv))
        out = np.zeros((len(dv), max(n)))
        for f in range(len(dv)):
            idx = [int(v) for v in fi[f]]; k = int(dv[f])
            for j in range(n[k]):
                idx[k] = j
                out[f, j] = full[tuple(idx)]
        return out
    cr = capi.Cross(n, [1, 2, 2, 2, 1])
    cores, ranks, nfib, _ = cr.run_adapt(fn, kickrank=2, maxrank=7, round_tol=1e-10, maxiter=3)
    assert list(ranks) == [1, 4, 4, 4, 1]
    assert all(int(x) > 4 for x in cr.ranks[1:-1])                       # the cross ran above the rank it returned
    assert np.abs(_tt_full(n, ranks, cores) - full).max() <= 1e-9
    assert nfib == sum(calls)
    # the next solver step starts from found + 1 (src/valuefunc.c:637-648)
    cr.set_ranks([1] + [int(x) + 1 for x in ranks[1:-1]] + [1])
    assert list(cr.ranks) == [1, 5, 5, 5, 1]
    cores2, ranks2, _, _ = cr.run_adapt(fn, kickrank=2, maxrank=7, round_tol=1e-10, maxiter=2)
    assert list(ranks2) == [1, 4, 4, 4, 1] and list(cr.ranks) == [1, 5, 5, 5, 1]     # nothing to kick this time
    assert np.abs(_tt_full(n, ranks2, cores2) - full).max() <= 1e-9
    cr.close()


def test_adaptive_cross_respects_maxrank_and_kick_zero():
    n = [6, 6, 6]
    rng = np.random.default_rng(4)
    full = rng.standard_normal(n)                                        # full rank: 6, 6

    def fn(dv, fi):
        out = np.zeros((len(dv), 6))
        for f in range(len(dv)):
            idx = [int(v) for v in fi[f]]; k = int(dv[f])
            for j in range(6):
                idx[k] = j
                out[f, j] = full[tuple(idx)]
        return out
    cr = capi.Cross(n, [1, 2, 2, 1])
    _, ranks, _, _ = cr.run_adapt(fn, kickrank=3, maxrank=4, round_tol=1e-12, maxiter=2, maxiter_adapt=6)
    assert list(cr.ranks) == [1, 4, 4, 1] and list(ranks) == [1, 4, 4, 1]
    cr.close()
    cr = capi.Cross(n, [1, 2, 2, 1])
    _, ranks, _, _ = cr.run_adapt(fn, kickrank=0, maxrank=4, round_tol=1e-12, maxiter=2)
    assert list(cr.ranks) == [1, 2, 2, 1] and list(ranks) == [1, 2, 2, 1]
    cr.close()


def test_twin_rows_do_not_degenerate_the_index_sets(built):
    """absorbing LQG with a constant boundary cost: every face fiber is the same constant row.  Grown
    index sets used to pick two of them, which makes the next unfolding rank-deficient and freezes a
    bond below the rank it needs; with the twins withheld from the pivoting the adaptive cross follows
    the dense backup to the rounding tolerance (CPU oracle as the operator)."""
    import itertools
    cfg = configs.get_config("lqgnd", n=12, rank=8, dx=4)
    port = make_port(cfg)
    r0, c0 = synthetic.quadratic_cores(port.xgrid)
    N = 12
    idx = np.array(list(itertools.product(range(N), repeat=3)), dtype=np.int32)
    fi = np.zeros((len(idx), 4), np.int32); fi[:, 1:] = idx
    dv = np.zeros(len(idx), np.int32)
    adapt = capi.Cross(cfg.ngrid, [1, 3, 3, 3, 1])
    ca, ra = c0, np.asarray(r0, dtype=np.uint64)
    for it in range(2):
        ft = po.FT(cfg.ngrid, ra, ca)
        full = port.vi_batch(ft, dv, fi, nthreads=4)[0].reshape(N, N, N, N).transpose(3, 0, 1, 2)
        ca, ra, _, _ = adapt.run_adapt(lambda a, b: port.vi_batch(ft, a, b, nthreads=4)[0], kickrank=2, maxrank=12,
                                       round_tol=1e-5, maxiter=3, maxiter_adapt=8)
        err = np.linalg.norm(_tt_full([N] * 4, ra, ca) - full) / np.linalg.norm(full)
        assert err <= 3e-5, (it, err, ra)
        left, _ = adapt.index_sets(1)
        assert not ({0, N - 1} <= set(int(v) for v in left[:, 0]))      # both faces of dim 0 = the same row twice
        adapt.set_ranks([1] + [min(int(x) + 1, 12) for x in ra[1:-1]] + [1])
    adapt.close()


@pytest.mark.gpu
def test_adaptive_value_iteration_tracks_the_dense_backup(gpu):
    """value iteration with rank adaptation (start rank 3, kick 2, round_tol 1e-5): each step's train
    is within a few round_tol of the dense tensor of the same backup, at ranks below the cap"""
    import itertools
    cfg = configs.get_config("lqgnd", n=12, rank=8, dx=4)
    prob = capi.Problem(cfg, arith=1)
    r0, c0 = synthetic.quadratic_cores(prob.xgrid)
    N = 12
    idx = np.array(list(itertools.product(range(N), repeat=3)), dtype=np.int32)
    fi = np.zeros((len(idx), 4), np.int32); fi[:, 1:] = idx
    dv = np.zeros(len(idx), np.int32)
    adapt = capi.Cross(cfg.ngrid, [1, 3, 3, 3, 1])
    ca, ra = c0, np.asarray(r0, dtype=np.uint64)
    for it in range(3):
        vf = capi.ValueF(cfg.ngrid, ra, ca)
        full = prob.vi_batch(vf, dv, fi)[0].reshape(N, N, N, N).transpose(3, 0, 1, 2)
        ca, ra, nfib, _ = adapt.run_vi_adapt(prob, vf, kickrank=2, maxrank=12, round_tol=1e-5, maxiter=3, maxiter_adapt=8)
        vf.close()
        err = np.linalg.norm(_tt_full([N] * 4, ra, ca) - full) / np.linalg.norm(full)
        assert err <= 5e-5, (it, err, ra)
        assert nfib > 0 and all(int(a) <= int(b) for a, b in zip(ra, adapt.ranks))
        adapt.set_ranks([1] + [min(int(x) + 1, 12) for x in ra[1:-1]] + [1])
    assert int(ra[1]) < 12
    prob.close(); adapt.close()


@pytest.mark.parametrize("name,n,rank,dx,iters", [("lqg2d_new", 30, 5, None, 3), ("lqg2d_reflect", 24, 6, None, 3),
                                                    ("lqgnd", 8, 6, 4, 2)])
def test_pivoting_ignores_round_off_of_the_operator(built, name, n, rank, dx, iters):
    """two operators that agree to round-off (the GPU and the CPU oracle do) must drive the cross to the
    same index sets: symmetric problems tie mirrored rows exactly, and over-specified ranks leave
    numerically dependent columns whose QR directions are pure round-off (tools/pivot_stability.py)"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import pivot_stability as ps
    for seed in (2, 5):
        assert ps.run(name, n, rank, dx, iters, seed) <= 1e-10


def _dense(n, ranks, cores):
    """nodal tensor of a train (blocks column-major r_k x r_{k+1})"""
    d = len(n)
    t = np.ones((1, 1))
    for k in range(d):
        g = np.asarray(cores[k]).reshape(int(n[k]), int(ranks[k + 1]), int(ranks[k])).transpose(2, 0, 1)   # [a, j, b]
        t = np.tensordot(t, g, axes=([t.ndim - 1], [0]))
    return t.reshape([int(x) for x in n])


@pytest.mark.parametrize("d", [1, 2, 3])
def test_continuous_l2_inner_product_against_gauss_quadrature(d):
    """c3sc_cores_dot_l2 (the reference's valuef_norm semantics: integral over the box of the product of the two
    piecewise-linear interpolants) against 3-point Gauss quadrature of the interpolants cell by cell on non-uniform
    grids -- an independent evaluation of the same integral (a product of multilinear functions is quadratic per
    dimension, so the rule is exact)."""
    rng = np.random.default_rng(5 + d)
    n = np.array([7, 5, 6][:d], dtype=np.uint64)
    xg = [np.sort(rng.uniform(-1.0, 2.0, int(m))) for m in n]
    ra = np.array([1] + [3] * (d - 1) + [1], dtype=np.uint64); rb = np.array([1] + [2] * (d - 1) + [1], dtype=np.uint64)
    ca = [rng.standard_normal(int(n[k] * ra[k] * ra[k + 1])) for k in range(d)]
    cb = [rng.standard_normal(int(n[k] * rb[k] * rb[k + 1])) for k in range(d)]
    A, B = _dense(n, ra, ca), _dense(n, rb, cb)
    gp, gw = np.polynomial.legendre.leggauss(3)
    for k in range(d):                                   # interpolate to the Gauss points of every cell, weights along
        x = xg[k]
        P = np.zeros((3 * (len(x) - 1), len(x))); w = np.zeros(3 * (len(x) - 1))
        for c in range(len(x) - 1):
            h = x[c + 1] - x[c]
            for q in range(3):
                t = 0.5 * (gp[q] + 1.0)
                P[3 * c + q, c] = 1.0 - t; P[3 * c + q, c + 1] = t
                w[3 * c + q] = 0.5 * h * gw[q]
        A = np.moveaxis(np.tensordot(P, A, axes=([1], [k])), 0, k)
        B = np.moveaxis(np.tensordot(P, B, axes=([1], [k])), 0, k)
        shape = [1] * d; shape[k] = len(w)
        A = A * w.reshape(shape)
    want = float((A * B).sum())
    got = capi.cores_dot_l2(n, xg, ra, ca, rb, cb)
    assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
    # norm2diff of a train with itself is 0, with zero it is the norm; and the nodal l2 is a different number
    assert capi.cores_norm2diff_l2(n, xg, ra, ca, ra, ca) <= 1e-7 * np.sqrt(abs(capi.cores_dot_l2(n, xg, ra, ca, ra, ca)))
    zero = [np.zeros_like(c) for c in cb]
    assert abs(capi.cores_norm2diff_l2(n, xg, ra, ca, rb, zero) ** 2 - capi.cores_dot_l2(n, xg, ra, ca, ra, ca)) <= 1e-10
    assert abs(capi.cores_norm(n, ra, ca) ** 2 - capi.cores_dot_l2(n, xg, ra, ca, ra, ca)) > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("d,rmax", [(3, 5), (4, 12), (6, 20), (3, 40)])
def test_device_l2_inner_products_equal_the_host_restatement(gpu, d, rmax):
    """c3sc_valuef_dot_l2 / _norm_l2 / _norm2diff_l2 (valuef_norm / valuef_norm2diff next to the device-resident cores, SURVEY
    8(f)-3) against c3sc_cores_*_l2, the host restatement of the same mass-matrix contraction that is pinned against Gauss
    quadrature above: non-uniform grids, two trains of different ranks; 1e-12 of the norms' own size (the summation order over the
    nodes differs: 32 partial sums per dimension)"""
    rng = np.random.default_rng(17 + d)
    n = [int(v) for v in rng.integers(7, 23, size=d)]
    xgrid = [np.sort(rng.uniform(-1.0, 2.0, size=m)) for m in n]
    ra = [1] + [int(v) for v in rng.integers(2, rmax + 1, size=d - 1)] + [1]
    rb = [1] + [int(v) for v in rng.integers(1, rmax + 1, size=d - 1)] + [1]
    ca = [rng.standard_normal(n[k] * ra[k] * ra[k + 1]) for k in range(d)]
    cb = [rng.standard_normal(n[k] * rb[k] * rb[k + 1]) for k in range(d)]
    va = capi.ValueF(n, ra, ca); vb = capi.ValueF(n, rb, cb)
    aa = capi.cores_dot_l2(n, xgrid, ra, ca, ra, ca); ab = capi.cores_dot_l2(n, xgrid, ra, ca, rb, cb); bb = capi.cores_dot_l2(n, xgrid, rb, cb, rb, cb)
    scale = np.sqrt(aa * bb)
    assert abs(va.dot_l2(vb, xgrid) - ab) <= 1e-12 * scale
    assert abs(va.dot_l2(va, xgrid) - aa) <= 1e-12 * aa
    assert abs(va.norm_l2(xgrid) - np.sqrt(aa)) <= 1e-12 * np.sqrt(aa)
    host_diff = capi.cores_norm2diff_l2(n, xgrid, ra, ca, rb, cb)
    assert abs(va.norm2diff_l2(vb, xgrid) - host_diff) <= 1e-10 * (np.sqrt(aa) + np.sqrt(bb))       # a difference of squares
    # a train against a slightly perturbed copy of itself: the case the solvers' Cauchy criterion meets
    cc = [c * (1.0 + 1e-6 * rng.standard_normal(c.size)) for c in ca]
    vc = capi.ValueF(n, ra, cc)
    dd = va.norm2diff_l2(vc, xgrid); hh = capi.cores_norm2diff_l2(n, xgrid, ra, ca, ra, cc)
    assert abs(dd - hh) <= 1e-6 * hh + 1e-9 * np.sqrt(aa)
    assert va.norm2diff_l2(va, xgrid) <= 1e-7 * np.sqrt(aa)      # sqrt of round-off
    va.close(); vb.close(); vc.close()
