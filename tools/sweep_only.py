"""One synthetic VI sweep (2 x d core batches) a few times: target for `ncu` launch lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from c3sc_b200 import capi, configs, synthetic
capi.check(capi.lib().c3sc_cuda_init(0))
cfg = configs.get_config("lqgnd_reflect")
prob = capi.Problem(cfg, arith=1); ranks = cfg.ranks()
vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
batches = synthetic.sweep_fibers(cfg.ngrid, ranks)
dev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), len(a)) for a, b in batches]
out = torch.zeros(max(f for _, _, f in dev) * cfg.n, dtype=torch.float64, device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for a, b, f in dev:
        prob.vi_batch_dev(vf, f, a.data_ptr(), b.data_ptr(), cfg.n, out.data_ptr())
    torch.cuda.synchronize()
    print("sweep %d: %.3f ms" % (r, 1e3 * (time.perf_counter() - t0)))
