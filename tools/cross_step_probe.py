"""One value-iteration step through the host cross driver at the bench shape, timed for several team sizes of the
pivoting step (C3SC_HOST_THREADS); verbose prints the driver's own split (operator / pivoting / norms)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from c3sc_b200 import capi, configs, synthetic
capi.check(capi.lib().c3sc_cuda_init(0))
cfg = configs.get_config("lqgnd_reflect")
prob = capi.Problem(cfg, arith=1); ranks = cfg.ranks()
vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
ref = None
for threads in (sys.argv[1:] or ["1", "2", "4", "8", "16"]):
    os.environ["C3SC_HOST_THREADS"] = threads
    cr = capi.Cross(cfg.ngrid, ranks)
    cr.run_vi(prob, vf, maxiter=1)
    best = 1e9
    for _ in range(5):
        cr2 = capi.Cross(cfg.ngrid, ranks)
        cr2.run_vi(prob, vf, maxiter=1)          # fresh index sets every time: the same work
        t0 = time.perf_counter(); cores, nf, ch = cr2.run_vi(prob, vf, maxiter=1); dt = time.perf_counter() - t0
        best = min(best, dt); cr2.close()
    h = hash(tuple(np.asarray(c).tobytes() for c in cores))
    if ref is None: ref = h
    print("threads %s: cross step %.3f ms (best of 5), %d fibers, same numbers as first: %s" % (threads, best * 1e3, nf, h == ref), flush=True)
    cr.close()
os.environ["C3SC_HOST_THREADS"] = "8"
cr = capi.Cross(cfg.ngrid, ranks)
cr.run_vi(prob, vf, maxiter=1)
cr.run_vi(prob, vf, maxiter=1, verbose=1)
