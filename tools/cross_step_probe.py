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

# where one core batch of the sweep spends its time: enqueue (host, launches) vs completion (device), device buffers
import torch
batches = synthetic.sweep_fibers(cfg.ngrid, ranks)
a, b = batches[len(batches) // 2]
F = len(a)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
out = torch.zeros(F * cfg.n, dtype=torch.float64, device="cuda")
for _ in range(20):
    prob.vi_batch_dev(vf, F, da.data_ptr(), db.data_ptr(), cfg.n, out.data_ptr())
torch.cuda.synchronize()
te = tt = 0.0
R = 200
for _ in range(R):
    t0 = time.perf_counter()
    prob.vi_batch_dev(vf, F, da.data_ptr(), db.data_ptr(), cfg.n, out.data_ptr())
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    te += t1 - t0; tt += t2 - t0
print("core batch of %d fibers, device buffers: enqueue %.1f us, enqueue + completion %.1f us" % (F, te / R * 1e6, tt / R * 1e6))
t0 = time.perf_counter()
for _ in range(R):
    prob.vi_batch_dev(vf, F, da.data_ptr(), db.data_ptr(), cfg.n, out.data_ptr())
torch.cuda.synchronize()
print("back to back: %.1f us per batch" % ((time.perf_counter() - t0) / R * 1e6))
hv = np.zeros(F * cfg.n)
t0 = time.perf_counter()
for _ in range(R):
    prob.vi_batch(vf, a, b, want_argmin=False)
print("host buffers (c3sc_vi_batch, numpy in/out): %.1f us per batch" % ((time.perf_counter() - t0) / R * 1e6))
