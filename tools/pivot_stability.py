"""Pivot stability of the host cross driver: value iteration through the driver with the CPU oracle as
the operator, once as is and once with its values perturbed at round-off level (1e-13 relative, what a
different summation order on the GPU does).  Prints the sup-norm difference of the two value functions.
    PYTHONPATH=.:tests python tools/pivot_stability.py"""
import sys
import numpy as np
from c3sc_b200 import capi, configs, synthetic
from oracle import pyoracle as po
from helpers import make_port


def run(name, n, rank, dx, iters, seed):
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    port = make_port(cfg)
    ranks, cores0 = synthetic.quadratic_cores(port.xgrid) if cfg.model == configs.MODEL_LQGND else (cfg.ranks(), synthetic.random_cores(cfg.ngrid, cfg.ranks()))
    rng = np.random.default_rng(seed)
    crs = [capi.Cross(cfg.ngrid, cfg.ranks()) for _ in range(2)]
    cs = [cores0, cores0]
    rs = [np.asarray(ranks, dtype=np.uint64)] * 2
    for it in range(iters):
        for w in range(2):
            ft = po.FT(cfg.ngrid, rs[w], cs[w])

            def fn(dv, fi, ft=ft, w=w):
                v = port.vi_batch(ft, dv, fi, nthreads=4)[0]
                return v * (1.0 + 1e-13 * rng.standard_normal(v.shape)) if w else v
            cs[w], _, _ = crs[w].run(fn, maxiter=2)
            rs[w] = crs[w].ranks.copy()
    u = synthetic.uniform01(77, 400 * cfg.dx).reshape(400, cfg.dx)
    pts = cfg.lb + u * (cfg.ub - cfg.lb)
    fa, fb = po.FT(cfg.ngrid, rs[0], cs[0]), po.FT(cfg.ngrid, rs[1], cs[1])
    va = np.array([port.ft_eval_linear(fa, np.ascontiguousarray(x)) for x in pts])
    vb = np.array([port.ft_eval_linear(fb, np.ascontiguousarray(x)) for x in pts])
    return np.abs(va - vb).max() / max(np.abs(va).max(), 1.0)


if __name__ == "__main__":
    cases = [("lqg2d_reflect", 24, 6, None, 6), ("lqgnd", 12, 4, 4, 4), ("dubinscar_new", 14, 5, None, 4),
             ("lqg2d_new", 30, 5, None, 5), ("double_int", 24, 6, None, 4), ("lqgnd_reflect", 10, 5, 4, 3)]
    worst = 0.0
    for c in cases:
        for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
            e = run(*c, seed)
            worst = max(worst, e)
            print(c, seed, f"{e:.3e}")
    print("worst", worst)
