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
