"""Shared set-up for the parity tests: one problem described three ways
(GPU library, oracle port, reference objects) from the same Config."""
from __future__ import annotations

import numpy as np

from c3sc_b200 import configs, synthetic
from c3sc_b200.capi import grid_constants
from oracle import pyoracle as po

# (config, n, rank, dx) -- sizes the oracle finishes in well under a second
SMALL = [
    ("lqg2d_new", 20, 4, None),
    ("lqg2d_reflect", 14, 3, None),
    ("double_int", 24, 5, None),
    ("dubinscar_new", 16, 4, None),
    ("skidding5d", 10, 3, None),
    ("lqgnd", 12, 3, 4),
    ("lqgnd_reflect", 8, 3, 6),
    ("user_vdp", 24, 4, None),          # the example USER model (examples/user_model_vdp.cuh), general (non-separable) walk
]


def host_problem(cfg):
    """xgrid, h2, t, obstacle boxes computed on the host exactly as the reference does
    (c3control_create / mca_add_grid_refs / bound_rect_init)."""
    xg = [configs.c3_linspace(cfg.lb[i], cfg.ub[i], int(cfg.ngrid[i])) for i in range(cfg.dx)]
    h, hmin, h2, t = grid_constants(xg, cfg.lb, cfg.ub)
    if cfg.obs_center.size:
        olb = cfg.obs_center - cfg.obs_width / 2.0
        oub = cfg.obs_center + cfg.obs_width / 2.0
    else:
        olb = np.zeros((0, cfg.dx)); oub = np.zeros((0, cfg.dx))
    return xg, h, hmin, h2, t, olb, oub


def make_port(cfg):
    xg, h, hmin, h2, t, olb, oub = host_problem(cfg)
    return po.Port(cfg, xg, h2, t, olb, oub)


def make_ft(cfg, seed=0xC35C0000, rank=None):
    ranks = cfg.ranks(rank)
    cores = synthetic.random_cores(cfg.ngrid, ranks, seed=seed)
    return ranks, cores, po.FT(cfg.ngrid, ranks, cores)


def rel_err(a, b, scale=None):
    """max |a-b| / max(|b|, scale): `scale` guards entries that cancel to ~0."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if scale is None:
        scale = np.abs(b).max() if b.size else 1.0
    den = np.maximum(np.abs(b), scale)
    den = np.where(den == 0, 1.0, den)
    return float((np.abs(a - b) / den).max()) if a.size else 0.0


def valid_mask(cfg, dim_vary):
    """[F, nmax] mask of real nodes (j < ngrid[dim_vary])."""
    n = np.asarray(cfg.ngrid, dtype=np.int64)[np.asarray(dim_vary)]
    return np.arange(int(cfg.ngrid.max()))[None, :] < n[:, None]


def abs_ft(ft):
    """The same train with every core entry replaced by its magnitude.  Evaluated at the same indices it gives
    sum |terms| over all rank paths of an FT value: the condition-aware scale of that value (what a
    re-association of the chain products can legitimately move by a few ulp of)."""
    return po.FT(ft.n, ft.ranks, [np.abs(c) for c in ft.cores])


def elem_err_costs(port, ft, k, fixed, got, ref=None):
    """per-ELEMENT relative error of the neighbour values of one fiber: |got - ref| / sum|terms| of that entry
    (abs_ft), not relative to the largest entry of the batch.  Returns (max error, oracle costs, scales)."""
    if ref is None:
        _, ref = port.neighbor_costs(ft, k, fixed)
    _, scale = port.neighbor_costs(abs_ft(ft), k, fixed)
    scale = np.maximum(scale, np.abs(ref))
    scale = np.where(scale == 0, 1.0, scale)
    return float((np.abs(got - ref) / scale).max()), ref, scale


def elem_err_values(got, ref, cost_scale):
    """per-ELEMENT relative error of backed-up values: a value is dt*g + e^{-beta dt} sum_m p_m cost_m with
    sum p = 1, so sum |p_m cost_m| <= max_m |cost_m|-scale of THAT node's row (cost_scale [N, 2d+1] from
    elem_err_costs); the scale of node j is max(|ref_j|, max_m cost_scale[j, m])."""
    sc = np.maximum(np.abs(ref), cost_scale.max(axis=1))
    sc = np.where(sc == 0, 1.0, sc)
    return float((np.abs(got - ref) / sc).max())


def argmin_mismatches_are_ties(cfg, port, ft, dv, fi, arg, oarg, rtol=1e-12):
    """identical argmin, or the two candidates' ORACLE values tie within rtol (relative to the larger of the two):
    returns the number of (justified) mismatches, asserts on an unjustified one"""
    import ctypes as C
    bad = np.argwhere(arg != oarg)
    port.L.orc_control_value.restype = C.c_double
    cache = {}
    n = 0
    for f, j in bad:
        if j >= cfg.ngrid[dv[f]]:
            continue
        n += 1
        if f not in cache:
            cache[f] = (port.vi_fiber_full(ft, dv[f], fi[f]), port.fiber_points(dv[f], fi[f]))
        (out, ub, ab, costs), xs = cache[f]
        vals = []
        for cand in (arg[f, j], oarg[f, j]):
            assert 0 <= cand < cfg.nu, (f, j, cand)
            prob = np.zeros(2 * cfg.dx + 1); dt = C.c_double(); g = C.c_double(); st = C.c_int()
            u = np.ascontiguousarray(cfg.controls[cand])
            vals.append(port.L.orc_control_value(C.byref(port.p), po._p(np.ascontiguousarray(xs[j])), po._p(u),
                                                 po._p(np.ascontiguousarray(costs[j])), po._p(prob), C.byref(dt), C.byref(g), C.byref(st)))
        assert abs(vals[0] - vals[1]) <= rtol * max(abs(vals[0]), abs(vals[1]), np.abs(costs[j]).max()), (f, j, vals)
    return n
