"""Where the host-buffer (e2e) step of the bench workload spends its time against the device-resident one."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3sc_b200 import capi, configs, synthetic
capi.check(capi.lib().c3sc_cuda_init(0))
cfg = configs.get_config("lqgnd_reflect")
prob = capi.Problem(cfg, arith=1); ranks = cfg.ranks()
vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = cfg.n
dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=7)
dv_h = torch.from_numpy(dv).pin_memory(); fi_h = torch.from_numpy(fi).pin_memory()
out_h = torch.empty(F * N, dtype=torch.float64).pin_memory()
dv_d = dv_h.cuda(); fi_d = fi_h.cuda(); out_d = torch.empty(F * N, dtype=torch.float64, device="cuda")
L = capi.lib()
def dev_step(): prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, out_d.data_ptr())
def host_step():
    capi.check(L.c3sc_vi_batch(prob.handle, vf.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, out_h.data_ptr(), None))
def t(fn, sync_each, R=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(R):
        fn()
        if sync_each: torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / R * 1e3
print("device buffers, back to back      : %.3f ms" % t(dev_step, False))
print("device buffers, synchronize / step: %.3f ms" % t(dev_step, True))
t0 = time.perf_counter(); dev_step(); te = time.perf_counter() - t0; torch.cuda.synchronize()
print("  host time to enqueue one step   : %.3f ms" % (te * 1e3))
print("host buffers (c3sc_vi_batch)      : %.3f ms" % t(host_step, False))
