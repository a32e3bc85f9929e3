#!/usr/bin/env python
"""Sweep of the pipeline's chunking knobs on the headline workload (run under gpurun):
C3SC_CHUNK_MB (cost scratch in flight), C3SC_CHAIN_FIBERS (chain super-chunk), C3SC_LANES, bucketed on / off.
Prints ms per 65 536-fiber step (device-resident fibers) and ms of stage 1 alone."""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from c3sc_b200 import capi, configs, synthetic  # noqa: E402


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    cfg = configs.get_config("lqgnd_reflect")
    capi.check(capi.lib().c3sc_cuda_init(0))
    dev = torch.device("cuda", 0)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks(20)
    vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
    dv, fi = synthetic.random_fibers(cfg.ngrid, F)
    dv_d = torch.from_numpy(np.ascontiguousarray(dv)).to(dev); fi_d = torch.from_numpy(np.ascontiguousarray(fi)).to(dev)
    out = torch.zeros(F * cfg.n, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)

    def timed(fn, n=5, w=2):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(n):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / n

    def full():
        prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), cfg.n, out.data_ptr(), stream=st.cuda_stream)

    def stage1():
        prob.stage1_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), cfg.n, stream=st.cuda_stream)

    ref = None
    # combos from the command line: "K=V,K=V;K=V,..." (argv[2]); default: the pipeline's defaults and the per-fiber chains
    combos = [dict(), dict(C3SC_NO_BUCKETS="1")]
    if len(sys.argv) > 2:
        combos = [dict(kv.split("=") for kv in grp.split(",") if kv) for grp in sys.argv[2].split(";")]
    keys = sorted({k for c in combos for k in c} | {"C3SC_NO_BUCKETS"})
    for c in combos:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(c)
        ms = timed(full)
        ms1 = timed(stage1)
        v = out.cpu().numpy()
        if ref is None:
            ref = v.copy()
        err = float(np.abs(v - ref).max() / np.abs(ref).max())
        print(f"{c}: step {ms:.3f} ms = {F * cfg.n / ms / 1e6:.3f} G node-backups/s; stage 1 alone {ms1:.3f} ms "
              f"({F * cfg.n * 3392 / ms1 / 1e9:.2f} TFLOP/s contract); max dev from per-fiber path {err:.1e}", flush=True)
    prob.check()


if __name__ == "__main__":
    main()
