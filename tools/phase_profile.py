import sys, ctypes as C, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c3sc_b200 import capi, configs, synthetic
capi.check(capi.lib().c3sc_cuda_init(0))
L=capi.lib()
for name,F in (("lqgnd_reflect",8192),("skidding5d",8192),("dubinscar_new",8192)):
    cfg=configs.get_config(name)
    prob=capi.Problem(cfg, arith=1); ranks=cfg.ranks(); cores=synthetic.random_cores(cfg.ngrid,ranks); vf=capi.ValueF(cfg.ngrid,ranks,cores)
    dv,fi=synthetic.random_fibers(cfg.ngrid,F)
    dvd=torch.from_numpy(dv).cuda(); fid=torch.from_numpy(fi).cuda(); out=torch.zeros(F*cfg.n,dtype=torch.float64,device='cuda')
    for rep in range(2):
        prob.vi_batch_dev(vf,F,dvd.data_ptr(),fid.data_ptr(),cfg.n,out.data_ptr())
    torch.cuda.synchronize()
    L.c3sc_debug_phase_profile(1,None)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); prob.vi_batch_dev(vf,F,dvd.data_ptr(),fid.data_ptr(),cfg.n,out.data_ptr()); e1.record(); torch.cuda.synchronize()
    buf=(C.c_ulonglong*8)(); L.c3sc_debug_phase_profile(0,buf)
    a=np.array(list(buf),dtype=np.float64)
    print(name, "ms=%.3f"%e0.elapsed_time(e1), "nodes/s=%.3g"%(F*cfg.n/e0.elapsed_time(e1)*1e3), "cycles/fiber:", (a/F).round(0)[:8], "share:", (a/a.sum()).round(3)[:6])
