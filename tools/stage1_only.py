import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from c3sc_b200 import capi, configs, synthetic
F = 65536
cfg = configs.get_config("lqgnd_reflect")
capi.check(capi.lib().c3sc_cuda_init(0))
dev = torch.device("cuda", 0)
prob = capi.Problem(cfg, arith=1)
ranks = cfg.ranks(20)
vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
dv, fi = synthetic.random_fibers(cfg.ngrid, F)
dv_d = torch.from_numpy(np.ascontiguousarray(dv)).to(dev); fi_d = torch.from_numpy(np.ascontiguousarray(fi)).to(dev)
st = torch.cuda.current_stream(dev)
for _ in range(3):
    prob.stage1_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), cfg.n, stream=st.cuda_stream)
torch.cuda.synchronize()
