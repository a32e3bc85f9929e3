"""C3SC_DBG_TIMELINE=1 python tools/e2e_timeline.py: device-side timeline of one host-buffer step of the bench workload."""
import os, sys, time
os.environ.setdefault("C3SC_DBG_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3sc_b200 import capi, configs, synthetic
capi.check(capi.lib().c3sc_cuda_init(0))
cfg = configs.get_config("lqgnd_reflect")
prob = capi.Problem(cfg, arith=1); ranks = cfg.ranks()
vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
F = 65536; N = cfg.n
dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=7)
dv_h = torch.from_numpy(dv).pin_memory(); fi_h = torch.from_numpy(fi).pin_memory()
out_h = torch.empty(F * N, dtype=torch.float64).pin_memory()
L = capi.lib()
for i in range(4):
    sys.stderr.write("step %d\n" % i); sys.stderr.flush()
    t0 = time.perf_counter()
    capi.check(L.c3sc_vi_batch(prob.handle, vf.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, out_h.data_ptr(), None))
    sys.stderr.write("  host wall of the call %.3f ms (includes printing)\n" % ((time.perf_counter() - t0) * 1e3))
