"""Generate the golden fixtures in this directory from the REFERENCE'S OWN OBJECT CODE
(oracle/_ref/libc3sc_ref.so = /root/reference/src/*.c compiled in place, see oracle/Makefile).

Run here (needs /root/reference):   python tests/golden/make_golden.py
Inputs are seeded (c3sc_b200/synthetic.py), so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from c3sc_b200 import configs, synthetic          # noqa: E402
from oracle import pyoracle as po                  # noqa: E402
from helpers import make_ft                        # noqa: E402

CASES = [("lqg2d_new", 20, 4, 2), ("lqg2d_reflect", 14, 3, 2), ("double_int", 24, 5, 2),
         ("dubinscar_new", 16, 4, 3), ("skidding5d", 10, 3, 5), ("lqgnd", 12, 3, 4), ("lqgnd_reflect", 8, 3, 6)]


def main():
    po.build_ref()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, n, rank, dx in CASES:
        cfg = configs.get_config(name, n=n, rank=rank, dx=dx if name.startswith("lqgnd") else None)
        ref = po.Ref(cfg)
        ranks, cores, ft = make_ft(cfg)
        _, _, ft2 = make_ft(cfg, seed=0xABCD00)
        vf, vf2 = ref.valuef(ft), ref.valuef(ft2)
        F = 24
        dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=0x601D, face_frac=0.3)
        nmax = cfg.n
        absorbed = np.zeros((F, nmax), np.int32); nv = np.zeros((F, nmax, 2), np.int64)
        nf = np.zeros((F, max(cfg.dx - 1, 1), 2), np.int64); costs = np.zeros((F, nmax, 2 * cfg.dx + 1))
        for f in range(F):
            a, v, w = ref.fiber_neighbors(dv[f], fi[f])
            absorbed[f], nv[f], nf[f] = a, v, w
            _, costs[f] = ref.neighbor_costs(vf, dv[f], fi[f])
        vi, _ = ref.vi_fibers(vf, dv, fi)
        ref.pi_begin(vf)
        pi1, _ = ref.pi_fibers(vf2, dv, fi)
        pi2, _ = ref.pi_fibers(vf, dv, fi)
        np.savez_compressed(os.path.join(here, f"{name}_n{n}_r{rank}.npz"), name=name, n=n, rank=rank, dx=cfg.dx,
                            dim_vary=dv, fixed_ind=fi, absorbed=absorbed, nbr_vary=nv, nbr_fixed=nf,
                            costs=costs, vi=vi, pi1=pi1, pi2=pi2)
        ref.close()
        print("wrote", name)


if __name__ == "__main__":
    main()
