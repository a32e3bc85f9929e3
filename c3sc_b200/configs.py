"""Problem definitions behind BASELINE.json's five configs.

Each entry restates the *setup* part of a reference example `main()`:
bounds, nodes per dimension, boundary types, obstacles, discount and the
discrete control table.  Dynamics/costs are selected by `model` (ids of
`enum c3sc_model`, include/c3sc_b200.h) and live on the device
(c3sc_b200/csrc/models.cuh); host copies used by the oracle are in
oracle/models.c.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np

MODEL_LQGND, MODEL_DOUBLE_INT, MODEL_DUBINS, MODEL_SKID5D, MODEL_USER = 1, 2, 3, 4, 5
ABSORB, PERIODIC, REFLECT = 1, 2, 3          # enum EBTYPE, reference src/boundary.h:42-47


def c3_linspace(lb: float, ub: float, n: int) -> np.ndarray:
    """C3 `linspace` (array.c): running sum of the interval, not lb + i*step."""
    g = np.empty(n, dtype=np.float64)
    g[0] = lb
    step = (np.float64(ub) - np.float64(lb)) / np.float64(n - 1)
    for i in range(1, n):
        g[i] = g[i - 1] + step
    return g


@dataclass
class Config:
    name: str
    model: int
    dx: int
    du: int
    dw: int
    n: int                       # nodes per dimension
    lb: np.ndarray
    ub: np.ndarray
    bc: np.ndarray               # [dx] EBTYPE
    beta: float
    controls: np.ndarray         # [nu, du] candidate-major
    rank: int                    # FT rank of the synthetic value function
    obs_center: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    obs_width: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    params: np.ndarray = field(default_factory=lambda: np.zeros(0))
    nvec: np.ndarray | None = None   # per-dimension node counts when they differ (tests); n = their maximum

    @property
    def ngrid(self) -> np.ndarray:
        if self.nvec is not None:
            return np.asarray(self.nvec, dtype=np.uint64)
        return np.full(self.dx, self.n, dtype=np.uint64)

    @property
    def nu(self) -> int:
        return int(self.controls.shape[0])

    def ranks(self, rank: int | None = None) -> np.ndarray:
        r = np.full(self.dx + 1, self.rank if rank is None else rank, dtype=np.uint64)
        r[0] = r[-1] = 1
        return r


def _tensor_grid(vals, du):
    return np.array(list(itertools.product(vals, repeat=du)), dtype=np.float64).reshape(-1, du)


def get_config(name: str, n: int | None = None, rank: int | None = None, dx: int | None = None) -> Config:
    if name == "lqg2d_new":          # examples/lqg2d_new/lqg2d.c:241-278 (-n 60); BFGS on [-1,1] -> {-1,0,1}
        c = Config(name, MODEL_LQGND, 2, 1, 2, 60, np.full(2, -2.0), np.full(2, 2.0),
                   np.full(2, ABSORB, np.int32), 0.1, _tensor_grid([-1.0, 0.0, 1.0], 1), 5)
    elif name == "lqg2d_reflect":    # same with -t 1 (lqg2d.c:273-276)
        c = Config(name, MODEL_LQGND, 2, 1, 2, 60, np.full(2, -2.0), np.full(2, 2.0),
                   np.full(2, REFLECT, np.int32), 0.1, _tensor_grid([-1.0, 0.0, 1.0], 1), 5)
    elif name == "double_int":       # examples/double_int/double_int.c:259-314 (-n 100)
        c = Config(name, MODEL_DOUBLE_INT, 2, 1, 2, 100, np.full(2, -2.0), np.full(2, 2.0),
                   np.full(2, ABSORB, np.int32), 0.1, _tensor_grid([-1.0, 0.0, 1.0], 1), 8,
                   obs_center=np.zeros((1, 2)), obs_width=np.full((1, 2), 0.4))
    elif name == "dubinscar_new":    # examples/dubinscar_new/dubinscar.c:282-322 (-n 100)
        c = Config(name, MODEL_DUBINS, 3, 1, 3, 100, np.array([-4.0, -4.0, -math.pi]),
                   np.array([4.0, 4.0, math.pi]), np.array([ABSORB, ABSORB, PERIODIC], np.int32),
                   0.0, _tensor_grid([-1.0, 0.0, 1.0], 1), 10,
                   obs_center=np.zeros((1, 3)), obs_width=np.array([[0.5, 0.5, 2.0 * math.pi]]))
    elif name == "skidding5d":       # examples/skidding5d/scar.c:274-342 (-n 50)
        lbs = np.array([-500.0, -500.0, -math.pi, -0.5, -10.0])
        ubs = np.array([500.0, 500.0, math.pi, 0.5, 10.0])
        ctr = np.array([[0.0, 0.0, 0.0, (ubs[3] + lbs[3]) / 2.0, (ubs[4] + lbs[4]) / 2.0]])
        wid = np.array([[40.0, 40.0, ubs[2] - lbs[2], ubs[3] - lbs[3], ubs[4] - lbs[4]]])
        u = c3_linspace(-5.0 * math.pi / 180.0, 5.0 * math.pi / 180.0, 20).reshape(-1, 1)
        c = Config(name, MODEL_SKID5D, 5, 1, 5, 50, lbs, ubs,
                   np.array([REFLECT, REFLECT, PERIODIC, ABSORB, ABSORB], np.int32), 1.0, u, 15,
                   obs_center=ctr, obs_width=wid)
    elif name in ("lqgnd", "lqgnd_reflect"):   # examples/lqgnd/lqgnd.c:300-353 (-x 10 -n 100); BFGS -> {lbu,0,ubu}^du
        d = 10 if dx is None else dx
        bcv = ABSORB if name == "lqgnd" else REFLECT
        c = Config(name, MODEL_LQGND, d, d // 2, d, 100, np.full(d, -2.0), np.full(d, 2.0),
                   np.full(d, bcv, np.int32), 0.1, _tensor_grid([-1.0, 0.0, 1.0], d // 2), 20)
    elif name == "user_vdp":         # examples/user_model_vdp.cuh: the example USER model (controlled Van der Pol), not a BASELINE config
        c = Config(name, MODEL_USER, 2, 1, 2, 40, np.full(2, -3.0), np.full(2, 3.0),
                   np.array([REFLECT, ABSORB], np.int32), 0.2, c3_linspace(-2.0, 2.0, 9).reshape(-1, 1), 6,
                   obs_center=np.array([[1.5, 1.5]]), obs_width=np.array([[0.6, 0.6]]))
    else:
        raise KeyError(name)
    if n is not None:
        c.n = n
    if rank is not None:
        c.rank = rank
    return c


ALL_CONFIGS = ("lqg2d_new", "double_int", "dubinscar_new", "skidding5d", "lqgnd")
