"""D2H bandwidth of the value buffer (pinned host memory) on an idle GPU and while the device-resident pipeline runs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3sc_b200 import capi, configs, synthetic
capi.check(capi.lib().c3sc_cuda_init(0))
cfg = configs.get_config("lqgnd_reflect")
prob = capi.Problem(cfg, arith=1); ranks = cfg.ranks()
vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
F = 65536; N = cfg.n
dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=7)
dv_d = torch.from_numpy(dv).cuda(); fi_d = torch.from_numpy(fi).cuda()
out_d = torch.empty(F * N, dtype=torch.float64, device="cuda")
src = torch.randn(F * N, dtype=torch.float64, device="cuda")
out_h = torch.empty(F * N, dtype=torch.float64).pin_memory()
side = [torch.cuda.Stream(), torch.cuda.Stream()]
def dev_step(): prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, out_d.data_ptr())
def copies(pieces, nstreams):
    n = F * N // pieces
    evs = []
    for i in range(pieces):
        s = side[i % nstreams]
        with torch.cuda.stream(s):
            out_h[i * n:(i + 1) * n].copy_(src[i * n:(i + 1) * n], non_blocking=True)
    for s in side[:nstreams]:
        e = torch.cuda.Event(); e.record(s); evs.append(e)
    return evs
def timed(pieces, nstreams, load):
    for _ in range(2):
        if load: dev_step()
        for e in copies(pieces, nstreams): e.synchronize()
        torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        if load:
            dev_step(); dev_step()                   # ~3.5 ms of pipeline in flight under the copies
        t0 = time.perf_counter()
        for e in copies(pieces, nstreams): e.synchronize()
        best = min(best, time.perf_counter() - t0)
        torch.cuda.synchronize()
    return F * N * 8 / best / 1e9, best * 1e3
for load in (False, True):
    for pieces, ns in ((1, 1), (6, 1), (6, 2), (20, 2)):
        bw, ms = timed(pieces, ns, load)
        print("%s %2d pieces on %d stream(s): %.1f GB/s (%.3f ms for 52 MB)" % ("under load," if load else "idle GPU,  ", pieces, ns, bw, ms))
