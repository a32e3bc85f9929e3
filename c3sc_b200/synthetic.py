"""Seeded synthetic inputs of SURVEY.md §8(d): FT cores and fiber lists.

Deterministic (SplitMix64) so the CPU oracle, the reference build and the GPU
path all see bit-identical inputs without shipping data files.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n successive SplitMix64 outputs as uint64 (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = (np.uint64(seed & _M64) + idx * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int) -> np.ndarray:
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def random_cores(ngrid, ranks, seed: int = 0xC35C0000, scale: float = 1.0) -> list[np.ndarray]:
    """G_k[j][a,b] ~ U(-1,1)/sqrt(r_k), seed + k per core; block j column-major r_k x r_{k+1}."""
    cores = []
    for k in range(len(ngrid)):
        cnt = int(ngrid[k]) * int(ranks[k]) * int(ranks[k + 1])
        u = uniform01(seed + k, cnt)
        cores.append(scale * (2.0 * u - 1.0) / np.sqrt(float(ranks[k])))
    return cores


def quadratic_cores(xgrid: list[np.ndarray]) -> tuple[np.ndarray, list[np.ndarray]]:
    """Exact rank-2 FT of sum_i x_i^2 (the `startcost` of examples/lqgnd/lqgnd.c:200-215)."""
    d = len(xgrid)
    ranks = np.full(d + 1, 2, dtype=np.uint64); ranks[0] = ranks[-1] = 1
    cores = []
    for k, g in enumerate(xgrid):
        n = g.size
        q = g * g
        if d == 1:
            blk = q.reshape(n, 1, 1)
        elif k == 0:                      # [x^2, 1]
            blk = np.stack([q, np.ones(n)], axis=1).reshape(n, 1, 2)
        elif k == d - 1:                  # [1; x^2]
            blk = np.stack([np.ones(n), q], axis=1).reshape(n, 2, 1)
        else:                             # [[1, 0], [x^2, 1]]
            blk = np.zeros((n, 2, 2)); blk[:, 0, 0] = 1.0; blk[:, 1, 0] = q; blk[:, 1, 1] = 1.0
        # store each block column-major: index a + b*r_k
        cores.append(np.ascontiguousarray(np.transpose(blk, (0, 2, 1))).reshape(-1))
    return ranks, cores


def random_fibers(ngrid, F: int, seed: int = 0xF1BE, face_frac: float = 0.1):
    """F fibers: dim_vary = f mod d, fixed_ind ~ U{0..N-1}, a fraction forced onto faces."""
    d = len(ngrid)
    dim_vary = (np.arange(F) % d).astype(np.int32)
    u = uniform01(seed, F * d).reshape(F, d)
    fixed = np.minimum((u * np.asarray(ngrid, dtype=np.float64)).astype(np.int64), np.asarray(ngrid, dtype=np.int64) - 1)
    v = uniform01(seed ^ 0x5EED, F * d).reshape(F, d)
    lo = v < face_frac / 2
    hi = (v >= face_frac / 2) & (v < face_frac)
    fixed = np.where(lo, 0, fixed)
    fixed = np.where(hi, np.asarray(ngrid, dtype=np.int64) - 1, fixed)
    fixed[np.arange(F), dim_vary] = 0
    return dim_vary, fixed.astype(np.int32)


def sweep_fibers(ngrid, ranks, seed: int = 0x5EE9):
    """One synthetic TT-cross sweep (left->right then right->left): for core k every
    (left index, right index) pair, i.e. r_k * r_{k+1} fibers per core per direction
    (SURVEY.md §8(d)).  Returns a list of (dim_vary, fixed_ind) batches, one per core
    visit, in visiting order -- batches are sequential in a real cross sweep."""
    d = len(ngrid)
    batches = []
    order = list(range(d)) + list(range(d - 1, -1, -1))
    for visit, k in enumerate(order):
        rl, rr = int(ranks[k]), int(ranks[k + 1])
        F = rl * rr
        u = uniform01(seed + 131 * visit, (rl + rr) * d).reshape(rl + rr, d)
        idx = np.minimum((u * np.asarray(ngrid, dtype=np.float64)).astype(np.int64), np.asarray(ngrid, dtype=np.int64) - 1)
        left, right = idx[:rl], idx[rl:]
        fixed = np.zeros((F, d), dtype=np.int32)
        a, b = np.meshgrid(np.arange(rl), np.arange(rr), indexing="ij")
        a = a.reshape(-1); b = b.reshape(-1)
        fixed[:, :k] = left[a, :k]
        fixed[:, k + 1:] = right[b, k + 1:]
        batches.append((np.full(F, k, dtype=np.int32), fixed))
    return batches
