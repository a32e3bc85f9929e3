#!/usr/bin/env python
"""bench.py -- Bellman node-backups/s of the hot path on synthetic fibers.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                     the reference's own CPU path

One STEP = one pass of the hot path over one batch of synthetic fibers of the headline
config (BASELINE.json configs[4]): d=10 LQG, 100 nodes/dim, FT rank 20,
243 discrete controls, F fibers per GPU (weak scaling: per-GPU work fixed as N grows).
At N>1 a step additionally broadcasts the FT cores from rank 0 and all-gathers the
backed-up fiber values over NCCL (the path's real exchange); the gather of one
quarter of the rank's fibers overlaps the backup of the next quarter on a second stream.

value    node-backups/s, whole job, fiber descriptors + cores already resident in HBM
e2e      same metric through the host-buffer C-ABI call (c3sc_vi_batch): pinned host
         fiber descriptors in, values back out, copies inside the timed region
roofline FP64 pipe.  Top level = the dominant kernels, stage 1 (k_ft_chains + k_ft_nodes, FP64 tensor
         path, 3/4 of a step), timed alone in this run with CUDA events against their algorithmic
         8r^2 + 4dr flops per node, over the DFMA peak measured in this run.  contract_whole_step =
         node-backups/s x W (SURVEY §8(d) contract flops per node-backup; exceeds the peak because
         the grid walk of stage 2 shares partial sums between candidates).  One step = per chunk of
         the batch three kernels of ours (k_ft_chains, k_ft_nodes, k_control_grid) after a 1-CTA
         grouping kernel; their ncu numbers are in profiles/.
vi_sweep seconds per synthetic VI sweep: 2 x d sequential core batches of r_k*r_{k+1} fibers,
         the request pattern of the cross driver (include/c3sc_cross.h)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------------------
def contract_flops_per_node(cfg, rank: int) -> float:
    """W = 8 r^2 + 4 d r + n_u (F_dyn + 12 d + 30)   (contract figure).
    r^2 is the mean of r_k r_{k+1} over the varying core for dim_vary = f mod d."""
    d = cfg.dx
    r = cfg.ranks(rank).astype(np.float64)
    rbar2 = float(np.mean(r[:-1] * r[1:]))
    fdyn = {1: 2 * d + 2 * cfg.du, 2: 2 * d + 2, 3: 50, 4: 80}[cfg.model]
    return 8.0 * rbar2 + 4.0 * d * float(rank) + cfg.nu * (fdyn + 12.0 * d + 30.0)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a thread
    (the timed region of the default run is ~25 ms, shorter than one `nvidia-smi -lms` period); falls back to
    the nvidia-smi query loop of the profiling recipe when the NVML bindings are missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.sm, self.ts, self.mask, self.max_mhz = [], [], 0, None
        self.running = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        while self.running:
            try:
                mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                self.mask |= int(self.reasons_fn(self.handle))
                self.sm.append(mhz)
                self.ts.append(time.perf_counter())         # when the answer arrived
            except Exception:
                pass
            time.sleep(0.002)

    def count(self) -> int:
        return len(self.sm) if self.nvml is not None else len(self.lines)

    def count_between(self, t0: float, t1: float, before: int) -> int:
        """samples answered inside the wall-clock window [t0, t1] (nvidia-smi fallback: lines read since `before`)"""
        if self.nvml is None:
            return max(0, len(self.lines) - before)
        return sum(1 for t in list(self.ts) if t0 <= t <= t1)

    def start(self):
        if self.nvml is not None:
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=1.0)
            reasons = sorted(nm for nm, bit in self.BITS.items() if self.mask & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml, 2 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# --------------------------------------------------------------------------------------
def run_reference(args, cfg, rank_ft):
    """--impl reference: the reference's own CPU implementation (oracle/_ref = its sources
    compiled in place; else the oracle port), all host threads, bounded sample per step."""
    from c3sc_b200 import synthetic
    from oracle import pyoracle as po
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import host_problem

    wrank = int(os.environ.get("RANK", "0"))
    if wrank != 0:
        return
    ranks = cfg.ranks(rank_ft)
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    ft = po.FT(cfg.ngrid, ranks, cores)
    F = args.ref_fibers
    dv, fi = synthetic.random_fibers(cfg.ngrid, F * (args.steps + args.warmup))
    cores_n = os.cpu_count() or 1
    if po.have_ref():
        kind = "reference"
        ref = po.Ref(cfg)
        vf = ref.valuef(ft)
        # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; rank 0 alone runs this arm (the other ranks
        # have exited), so it takes every core of the box
        ref.set_omp_threads(cores_n)
        threads = ref.omp_threads()

        def step(i):
            s = slice(i * F, (i + 1) * F)
            _, secs = ref.vi_fibers(vf, dv[s], fi[s], fresh_per_fiber=False)
            return secs
    else:
        kind = "port"
        if not os.path.exists(po.PORT_PATH):
            po.build_port()
        xg, h, hmin, h2, t, olb, oub = host_problem(cfg)
        port = po.Port(cfg, xg, h2, t, olb, oub)
        threads = cores_n

        def step(i):
            s = slice(i * F, (i + 1) * F)
            t0 = time.perf_counter()
            port.vi_batch(ft, dv[s], fi[s], nthreads=threads)
            return time.perf_counter() - t0
    for i in range(args.warmup):
        step(i)
    total = 0.0
    for i in range(args.steps):
        total += step(args.warmup + i)
    nodes = F * cfg.n * args.steps
    val = nodes / total
    line = {
        "impl": "reference", "metric": "bellman_node_backups_per_s", "value": val, "unit": "node-backups/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, rank_ft, F, args),
        "cpu_baseline": {"value": val, "unit": "node-backups/s", "cores": threads, "kind": kind,
                         "sample": f"{F} fibers x {cfg.n} nodes per step, one bellman_vi call per fiber, "
                                   f"OpenMP over the nodes of a fiber ({threads} threads of {cores_n} host cores)"},
        "e2e": {"value": val, "unit": "node-backups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _ncu_record():
    for name in ("r02e_traffic.json", "r02c_traffic.json", "r02b_traffic.json", "r02_traffic.json", "r01_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
    return None


def roofline_record(stage1, whole, peak_dmma, peak_dfma, W, W1, F, kernel_ms, hbm_bytes, hbm_peak):
    """The `roofline` object of the bench line.  Top level = the dominant kernels: stage 1 (chain stage + node kernel, FP64
    tensor path), timed live by `main` with CUDA events on the launching stream against its ALGORITHMIC 8r^2 + 4dr flops per
    node, over the DMMA peak measured in the same run (both measured FP64 peaks are printed).
    `contract_whole_step` keeps SURVEY 8(d)'s whole-step figure: it exceeds 1 because stage 2 shares partial sums between
    candidates -- algebra, not utilisation.  With more than one rank stage 1 is not re-timed (flagged in `basis`)."""
    rec = _ncu_record() or {}
    traffic = rec.get("dram_bytes_per_fiber")
    roof = {"bound": "tensor", "unit": "TFLOP/s", "peak": peak_dmma,
            "peaks_measured_in_this_run": {"fp64_dmma_tflops": peak_dmma, "fp64_dfma_tflops": peak_dfma,
                                           "note": "one execution resource on B200 (they do not add up); the DMMA figure is the denominator "
                                                   "because stage 1's flops run as DMMAs; MEASURED_PEAKS.json has no FP64 figure"}}
    if stage1:
        roof.update({"achieved": stage1["achieved"], "frac": stage1["achieved"] / peak_dmma,
                     "basis": "stage 1 (chain stage + node kernel) timed alone in this run against its ALGORITHMIC 8r^2 + 4dr flops per node",
                     "kernels": stage1["kernels"], "fibers_per_launch_set": stage1["fibers"], "ms_per_launch_set": stage1["ms"],
                     "flops_per_node": W1, "frac_of_dfma_peak": stage1["achieved"] / peak_dfma, "stage1_live": stage1})
    else:
        roof.update({"achieved": whole, "frac": whole / peak_dmma, "stage1_live": None,
                     "basis": "whole step against the contract flops (stage 1 is timed alone only at --gpus 1)"})
    roof.update({
        "traffic": float(traffic) * F if traffic else None,
        "traffic_source": rec.get("source"),
        "contract_whole_step": {"achieved": whole, "frac": whole / peak_dmma, "flops_per_node_backup": W,
                                "note": "SURVEY 8(d) contract flops per node-backup over the whole step; stage 2 forms every candidate "
                                        "from shared partial sums (~2 FP64 instructions per candidate), so frac > 1 is algebra, not utilisation"},
        "dominant_kernel": rec.get("dominant_kernel"), "other_kernels": rec.get("other_kernels"),
        "hbm": {"algorithmic_bytes_per_step": hbm_bytes, "achieved_gbs": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "note": "not binding: cores stay in L2/SMEM, HBM sees descriptors in and values out"}})
    return roof


def workload_config(cfg, rank_ft, F, args):
    return {"workload": f"{cfg.name}: d={cfg.dx}, {cfg.n} nodes/dim, FT rank {rank_ft}, n_u={cfg.nu} discrete controls, "
                        f"{F} synthetic fibers/GPU/step (dim_vary = f mod d, 10% of fixed indices on faces)",
            "fibers_per_gpu": F, "nodes_per_fiber": cfg.n, "rank": rank_ft, "n_controls": cfg.nu,
            "arith": "fast" if args.arith else "exact",
            "l2": "256 MiB buffer written between timed steps (L2 flush); each step timed by its own CUDA event pair"}


def pin_to_gpu_numa_node(index: int):
    """bind this process (and so its first-touch page-locked allocations) to the CPUs next to its GPU: with eight
    ranks copying results to the host at once, buffers on the far socket halve the D2H rate"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="lqgnd_reflect", help="lqgnd_reflect = examples/lqgnd -t 1 (every node a full backup); lqgnd = absorbing faces")
    ap.add_argument("--fibers", type=int, default=65536, help="fibers per GPU per step")
    ap.add_argument("--rank", type=int, default=20)
    ap.add_argument("--arith", type=int, default=1, help="1 = FAST (default), 0 = EXACT reference order")
    ap.add_argument("--ref-fibers", type=int, default=48, help="fibers per step of the CPU reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the synthetic VI sweep, the cross step, the PI step and the other configs")
    args = ap.parse_args()

    from c3sc_b200 import configs, synthetic
    cfg = configs.get_config(args.config)
    rank_ft = args.rank if args.config.startswith("lqgnd") else cfg.rank

    if args.impl == "reference":
        run_reference(args, cfg, rank_ft)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpus = pin_to_gpu_numa_node(local) if world > 1 else 0

    import ctypes
    import torch
    import torch.distributed as dist
    from c3sc_b200 import capi, sharding

    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the Bellman backup has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = capi.lib()
    capi.check(L.c3sc_cuda_init(local))

    F = args.fibers
    N = cfg.n
    ranks = cfg.ranks(rank_ft)
    prob = capi.Problem(cfg, arith=args.arith)
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    core_ptr, core_cnt = vf.device_buffer()

    class _Arr:  # __cuda_array_interface__ wrapper, zero copy
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    core_view = torch.as_tensor(_Arr(core_ptr, core_cnt), device=dev) if world > 1 else None

    # one global fiber list of F*world fibers; rank g owns the contiguous block g (stable map)
    dv_all, fi_all = synthetic.random_fibers(cfg.ngrid, F * world)
    dv_s, fi_s, nreal = sharding.shard_fibers(dv_all, fi_all, world, rank)
    assert nreal == F
    dv_h = torch.from_numpy(dv_s).pin_memory()
    fi_h = torch.from_numpy(fi_s).pin_memory()
    dv_d = dv_h.to(dev); fi_d = fi_h.to(dev)
    out_h = torch.empty(F * N, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream

    # N > 1: a rank's values reach the others as stores from the control kernel -- to the NVSwitch multicast address of the
    # ranks' symmetric gathered buffers (C3SC_GATHER=mcast, default: one store per value, replicated by the switch; 1.95 against
    # 2.06 ms per step at 8 GPUs) or into every rank's peer-mapped buffer (p2p: CUDA IPC, one store per value and peer; also the
    # fallback without multicast support) --, as one bulk copy per pipeline chunk and peer on the copy engines (copy) or as one
    # small scatter kernel per chunk (scatter), both overlapped with the next chunk; a step ends with a barrier.
    # C3SC_GATHER=nccl: NCCL all-gather.
    gather_mode = os.environ.get("C3SC_GATHER", "mcast") if world > 1 else "none"
    peers = None
    gathered = None
    mcast = None
    if world > 1 and gather_mode == "mcast":
        # NVSwitch multicast (NVLS): the gathered buffers of all ranks are one symmetric allocation with a MULTICAST address;
        # a store to it is replicated by the switch into every rank's copy, so a value leaves the GPU once instead of once
        # per peer.  torch's symmetric memory is the plumbing (allocation + rendezvous); the stores are the control kernel's.
        try:
            import torch.distributed._symmetric_memory as symm_mem
            gname = dist.group.WORLD.group_name
            symt = symm_mem.empty(world * F * N, dtype=torch.float64, device=dev)
            hdl = symm_mem.rendezvous(symt, gname)
            mc_ptr = int(hdl.multicast_ptr or 0)
            if not mc_ptr:
                raise RuntimeError("no multicast address for the symmetric allocation")
            gathered = symt
            mcast = capi.McastPeers(mc_ptr)
            sync_flag = torch.zeros(1, device=dev)
        except Exception as exc:
            if rank == 0:
                print(f"bench: multicast gather unavailable ({exc}); using peer stores", file=sys.stderr)
            gather_mode = "p2p"
            gathered = None
            mcast = None
    if world > 1 and gather_mode in ("p2p", "copy", "scatter"):
        try:
            def _exchange(h):
                got = [None] * world
                dist.all_gather_object(got, h)
                return got
            peers = capi.PeerBuffers(world * F * N * 8, rank, world, _exchange)
            gathered = torch.as_tensor(_Arr(peers.own, world * F * N), device=dev)
            sync_flag = torch.zeros(1, device=dev)
        except Exception as exc:           # no peer access on this box: fall back to NCCL
            if rank == 0:
                print(f"bench: peer-mapped gather unavailable ({exc}); using NCCL", file=sys.stderr)
            peers = None
            gather_mode = "nccl"
    if mcast is not None:
        peers = mcast
    if world > 1 and peers is None:
        gathered = torch.empty(world * F * N, dtype=torch.float64, device=dev)
    # the rank's own output IS its slot of its gathered buffer (no private copy)
    out_d = gathered[rank * F * N:(rank + 1) * F * N] if gathered is not None else torch.zeros(F * N, dtype=torch.float64, device=dev)
    peer_mode = {"copy": 1, "scatter": 2}.get(gather_mode, 0)

    def peers_struct():
        o = capi.BatchOut()
        o.n_peers = len(peers.ptrs)
        for g, ptr in enumerate(peers.ptrs):
            o.value_peers[g] = ptr
        o.peer_offset = rank * F * N
        o.peer_mode = peer_mode
        return o

    # the broadcast of the cores and the rebuild of their derived copies run on a side stream: the batch's grouping and chain plan
    # read the descriptors only and go ahead on the main stream; the library waits for the commit (an event of the value
    # function) right before its first kernel that reads the cores
    side = torch.cuda.Stream(dev) if world > 1 and os.environ.get("C3SC_BCAST_INLINE") is None else None

    def bcast_cores():
        if side is None:
            sharding.broadcast_cores(core_view, src=0)
            vf.commit(stream=sptr)
            return
        side.wait_stream(stream)                      # everything before this step has read the old cores
        with torch.cuda.stream(side):
            sharding.broadcast_cores(core_view, src=0)
            vf.commit(stream=side.cuda_stream)

    def step_resident():
        if world == 1:
            prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, out_d.data_ptr(), stream=sptr)
            return
        bcast_cores()
        if peers is not None:
            prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, out_d.data_ptr(), stream=sptr,
                              peers=peers.ptrs, peer_offset=rank * F * N, peer_mode=peer_mode)
            dist.all_reduce(sync_flag)                # barrier on the stream: every rank's values have landed
            return
        prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, out_d.data_ptr(), stream=sptr)
        dist.all_gather_into_tensor(gathered, out_d)

    ps = peers_struct() if peers is not None else None

    def step_e2e():
        """the sharded step through the host-buffer entry: pinned descriptors in, this rank's values back out, cores
        broadcast and values gathered on the devices (N > 1) -- all inside the call / the timed region"""
        if world == 1:
            capi.check(L.c3sc_vi_batch(prob.handle, vf.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, out_h.data_ptr(), None))
            return
        sharding.broadcast_cores(core_view, src=0)
        vf.commit(stream=sptr)
        stream.synchronize()
        if ps is not None:
            capi.check(L.c3sc_vi_batch_peers(prob.handle, vf.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, out_h.data_ptr(), None,
                                             ctypes.byref(ps)))
            dist.all_reduce(sync_flag)
        else:
            capi.check(L.c3sc_vi_batch(prob.handle, vf.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, out_h.data_ptr(), None))
            out_d.copy_(out_h, non_blocking=True)
            dist.all_gather_into_tensor(gathered, out_d)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if world > 1:                                       # the gather must equal an NCCL all-gather of the ranks' blocks
        step_resident()
        torch.cuda.synchronize(dev)
        dist.barrier()
        ref_g = torch.empty(world * F * N, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(ref_g, out_d.clone())
        torch.cuda.synchronize(dev)
        if not torch.equal(ref_g, gathered):
            raise SystemExit("gathered values differ from the NCCL all-gather")
        del ref_g

    # The sampler runs from the warm-up on (one NVML query takes ~20 ms on these boxes, about the whole timed
    # region at the default K: a query has to be in flight when the region starts to be answered inside it);
    # warm-up, timed steps and the steps below are the identical load.
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize(dev)
    launches0 = L.c3sc_launch_count()
    seen0, wall0 = sampler.count(), time.perf_counter()
    ms_total = timed(step_resident, args.steps, 0)
    wall1 = time.perf_counter()
    launches = L.c3sc_launch_count() - launches0
    in_region, extra = sampler.count_between(wall0, wall1, seen0), 0
    while (extra < 40) if world > 1 else (sampler.count() < 5 and extra < 400):   # fixed count under torchrun: the step has collectives
        step_resident()
        extra += 1
        if extra % 8 == 0:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    timed_out = out_d.clone() if world == 1 else None       # what the timed steps left behind (later sections reuse out_d)
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = in_region
    clocks["untimed_steps_of_the_same_load_while_sampling"] = extra
    prob.check()

    # end-to-end through the host-buffer C-ABI entry (wall clock around the blocking call, max over ranks)
    def timed_host(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize(dev)
        el = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        return el
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed_host(step_e2e, e2e_steps, 2)

    nodes_per_step = F * N * world
    value = nodes_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = nodes_per_step * e2e_steps / (ms_e2e * 1e-3)

    # ---- N > 1 extras: strong scaling (the N = 1 batch cut N ways) and the sharded synthetic VI sweep ------------
    extras = {}
    if world > 1 and peers is not None:
        Fs = F // world                                  # this rank's share of a 65 536-fiber batch

        def step_strong():
            bcast_cores()
            prob.vi_batch_dev(vf, Fs, dv_d.data_ptr(), fi_d.data_ptr(), N, out_d.data_ptr(), stream=sptr,
                              peers=peers.ptrs, peer_offset=rank * Fs * N, peer_mode=peer_mode)
            dist.all_reduce(sync_flag)
        k = max(3, min(args.steps, 10))
        ms_s = timed(step_strong, k, 2) / k
        extras["scaling_strong"] = {"fibers_total": Fs * world, "ms_per_step": ms_s, "node_backups_per_s": Fs * world * N / (ms_s * 1e-3),
                                    "what": "the N = 1 batch split N ways (cores broadcast, values gathered every step)"}
        if not args.no_sweep:
            batches = synthetic.sweep_fibers(cfg.ngrid, ranks)
            shard = []
            for a, b in batches:
                a2, b2, _ = sharding.shard_fibers(a, b, world, rank)
                shard.append((torch.from_numpy(a2).to(dev), torch.from_numpy(b2).to(dev), len(a2)))

            def sweep_sharded():
                for a, b, f in shard:
                    prob.vi_batch_dev(vf, f, a.data_ptr(), b.data_ptr(), N, out_d.data_ptr(), stream=sptr,
                                      peers=peers.ptrs, peer_offset=rank * f * N, peer_mode=peer_mode)
                    dist.all_reduce(sync_flag)            # the driver needs a core batch's values before it asks for the next
            ms_sw = timed(sweep_sharded, 10, 3) / 10
            nodes_sw = sum(len(a) for a, _ in batches) * N
            extras["vi_sweep_sharded"] = {"seconds": ms_sw * 1e-3, "node_backups": nodes_sw, "core_batches": len(shard),
                                          "what": "2 x d sequential core batches, each cut N ways, gathered, barrier per batch"}

    line = None
    if rank == 0:
        W = contract_flops_per_node(cfg, rank_ft)
        peak_dfma = capi.measure_fp64_peak()
        peak_dmma = capi.measure_fp64_tensor_peak()
        kernel_ms = ms_total / args.steps
        rbar2 = float(np.mean([ranks[k] * ranks[k + 1] for k in range(cfg.dx)]))
        W1 = 8.0 * rbar2 + 4.0 * cfg.dx * float(max(ranks))
        stage1 = None
        if world == 1:
            # stage 1 alone (chain stage + node kernel = the neighbour values, 8r^2 + 4dr algorithmic flops per node), launched
            # exactly as the pipeline launches it (same chunks, lanes and scratch), timed live on the whole batch
            def step_stage1():
                prob.stage1_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, stream=sptr)
            k1 = max(3, min(args.steps, 10))
            ms1 = timed(step_stage1, k1, 3) / k1
            stage1 = {"kernels": "k_chain_step x (d-1) + k_ft_nodes (k_ft_chains below 4096 fibers)", "fibers": F, "ms": ms1,
                      "flops_per_node": W1, "achieved": F * N * W1 / (ms1 * 1e-3) / 1e12, "unit": "TFLOP/s"}
        whole = (F * N * W) / (kernel_ms * 1e-3) / 1e12 if world == 1 else (value / world) * W / 1e12
        hbm_bytes = F * N * 8.0 + F * (cfg.dx + 1) * 4.0 + core_cnt * 8.0
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        roof = roofline_record(stage1, whole, peak_dmma, peak_dfma, W, W1, F, kernel_ms, hbm_bytes, hbm_peak)
        gathers = {"copy": "one bulk copy per pipeline chunk and peer on the copy engines into peer-mapped buffers (CUDA IPC over NVLink) + barrier",
                   "p2p": "control kernel stores into every rank's peer-mapped buffer (CUDA IPC over NVLink) + barrier",
                   "scatter": "one small kernel per pipeline chunk stores the chunk into every rank's peer-mapped buffer, on a side stream + barrier",
                   "mcast": "stores from the control kernel to the NVSwitch multicast address of the ranks' symmetric gathered buffers (one store per value, replicated by the switch) + barrier",
                   "nccl": "nccl all_gather_into_tensor"}
        line = {
            "metric": "bellman_node_backups_per_s", "value": value, "unit": "node-backups/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(cfg, rank_ft, F, args),
                           **({"gather": gathers[gather_mode], "cores": "dist.broadcast of the contiguous core buffer every step",
                               "cpus_bound_per_rank": ncpus} if world > 1 else {})),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "node-backups/s",
                    "h2d_bytes_per_step": int(F * (cfg.dx + 1) * 4 * world), "d2h_bytes_per_step": int(F * N * 8 * world),
                    "ms_per_step": ms_e2e / e2e_steps,
                    "api": "c3sc_vi_batch (host buffers, pinned)" if world == 1 else
                           "c3sc_vi_batch_peers per rank (pinned host buffers) + core broadcast + gather + barrier: the sharded step"},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        line.update(extras)

    sweeps = not args.no_sweep and world == 1
    if sweeps:
        batches = synthetic.sweep_fibers(cfg.ngrid, ranks)
        dev_b = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), len(a)) for a, b in batches]

        def sweep():
            for a, b, f in dev_b:
                prob.vi_batch_dev(vf, f, a.data_ptr(), b.data_ptr(), N, out_d.data_ptr(), stream=sptr)
        ms_sw = timed(sweep, 10, 3) / 10
        nodes_sw = sum(f for _, _, f in dev_b) * N
        line["vi_sweep"] = {"seconds": ms_sw * 1e-3, "node_backups": nodes_sw, "launches": len(dev_b),
                            "node_backups_per_s": nodes_sw / (ms_sw * 1e-3),
                            "what": "2 x d sequential core batches of r_k*r_{k+1} fibers (SURVEY §8(d))"}
        # one value-iteration step through the host cross driver (include/c3sc_cross.h)
        cr = capi.Cross(cfg.ngrid, ranks)
        cr.run_vi(prob, vf, maxiter=1)
        t0 = time.perf_counter()
        reps, nf = 3, 0
        for _ in range(reps):
            _, nf, _ = cr.run_vi(prob, vf, maxiter=1)
        dt = (time.perf_counter() - t0) / reps
        # the same with the pivoting step on one host thread (the numbers are identical: fixed row blocks, not thread shares)
        saved = os.environ.get("C3SC_HOST_THREADS")
        os.environ["C3SC_HOST_THREADS"] = "1"
        cr.run_vi(prob, vf, maxiter=1)
        t0 = time.perf_counter()
        for _ in range(reps):
            cr.run_vi(prob, vf, maxiter=1)
        dt1 = (time.perf_counter() - t0) / reps
        if saved is None:
            del os.environ["C3SC_HOST_THREADS"]
        else:
            os.environ["C3SC_HOST_THREADS"] = saved
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        line["vi_cross_step"] = {"seconds": dt, "fibers": nf, "node_backups": nf * N, "node_backups_per_s": nf * N / dt,
                                 "seconds_one_host_thread": dt1,
                                 "host_threads": int(saved) if saved else min(8, ncpu),
                                 "what": "c3sc_cross_run_vi, one left-right + right-left sweep (host TT-cross: twin rows + QR + maxvol "
                                         "on a team of host threads, one batched operator call per core, page-locked host buffers)"}
        cr.close()
        line["pi_step"] = pi_step(prob, cfg, ranks, vf, dv_h, fi_h, dv_d, fi_d, F, N, dev, timed, timed_host, peak_dmma, W1)
        line["other_configs"] = other_configs(args, dev, timed)

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg, rank_ft, args.cpu_seconds)
        line["parity_check"] = parity_check(cfg, rank_ft, prob, vf, dv_s, fi_s, timed_out, N)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def pi_step(prob, cfg, ranks, vf, dv_h, fi_h, dv_d, fi_d, F, N, dev, timed, timed_host, peak_dmma, W1):
    """bellman_pi (src/bellman.c:1702-1886), half of every outer solver iteration: one policy improvement (rows from the argmin
    against the policy value function) + K = 10 sub-iterations against the iterate, rows RESIDENT on the device."""
    import torch
    from c3sc_b200 import capi, synthetic
    L = capi.lib()
    K = 10
    RW = 2 * cfg.dx + 3
    vf_it = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks, seed=0xABCD00))
    rows = torch.empty(F * N * RW, dtype=torch.float64, device=dev)
    val = torch.empty(F * N, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def improve():
        prob.pi_batch_dev(vf, vf_it, F, dv_d.data_ptr(), fi_d.data_ptr(), N, 0, rows.data_ptr(), 0, val.data_ptr(), stream=st)

    def subiter():
        prob.pi_batch_dev(None, vf_it, F, dv_d.data_ptr(), fi_d.data_ptr(), N, 1, rows.data_ptr(), 0, val.data_ptr(), stream=st)

    def whole():
        improve()
        for _ in range(K):
            subiter()
    ms_imp = timed(improve, 3, 1) / 3

    def improve_same():                     # policy function == iterate (the first sub-iteration of a solver step): one stage 1
        prob.pi_batch_dev(vf_it, vf_it, F, dv_d.data_ptr(), fi_d.data_ptr(), N, 0, rows.data_ptr(), 0, val.data_ptr(), stream=st)
    ms_imp_same = timed(improve_same, 3, 1) / 3
    ms_sub = timed(subiter, 5, 1) / 5
    ms_all = timed(whole, 2, 1) / 2
    out_h = torch.empty(F * N, dtype=torch.float64).pin_memory()

    def e2e_sub():
        capi.check(L.c3sc_pi_batch_resident(prob.handle, None, vf_it.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, 1, rows.data_ptr(),
                                            out_h.data_ptr()))
    ms_e2e = timed_host(e2e_sub, 3, 1) / 3
    nodes = F * N
    rec = {"what": f"one policy improvement + {K} sub-iterations on {F} fibers, policy rows resident on the device (c3sc_pi_batch_dev)",
           "ms_improvement": ms_imp, "ms_improvement_policy_is_iterate": ms_imp_same, "ms_sub_iteration": ms_sub, "ms_total": ms_all,
           "node_backups_per_s": nodes * (K + 1) / (ms_all * 1e-3), "sub_iteration_node_evals_per_s": nodes / (ms_sub * 1e-3),
           "e2e_sub_iteration": {"value": nodes / (ms_e2e * 1e-3), "unit": "node-evaluations/s", "ms": ms_e2e,
                                 "api": "c3sc_pi_batch_resident (pinned descriptors in, values out, rows stay on the device)",
                                 "h2d_bytes": int(F * (cfg.dx + 1) * 4), "d2h_bytes": int(F * N * 8)},
           "roofline": {"bound": "tensor", "basis": "a sub-iteration is stage 1 (8r^2 + 4dr flops per node) + a (2d+1) dot per node",
                        "achieved": nodes * (W1 + 2.0 * (2 * cfg.dx + 1)) / (ms_sub * 1e-3) / 1e12, "peak": peak_dmma, "unit": "TFLOP/s",
                        "frac": nodes * (W1 + 2.0 * (2 * cfg.dx + 1)) / (ms_sub * 1e-3) / 1e12 / peak_dmma}}
    vf_it.close()
    del rows, val
    return rec


def other_configs(args, dev, timed):
    """the other BASELINE.json configs at full size, short runs: device-resident rate and the rate through host buffers"""
    import torch
    from c3sc_b200 import capi, configs, synthetic
    L = capi.lib()
    out = {}
    for name, F in (("skidding5d", 16384), ("dubinscar_new", 16384), ("double_int", 8192), ("lqg2d_new", 8192)):
        cfg = configs.get_config(name)
        prob = capi.Problem(cfg, arith=args.arith)
        ranks = cfg.ranks()
        vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
        dv, fi = synthetic.random_fibers(cfg.ngrid, F)
        dv_h = torch.from_numpy(np.ascontiguousarray(dv)).pin_memory(); fi_h = torch.from_numpy(np.ascontiguousarray(fi)).pin_memory()
        dv_d = dv_h.to(dev); fi_d = fi_h.to(dev)
        N = cfg.n
        o = torch.empty(F * N, dtype=torch.float64, device=dev)
        oh = torch.empty(F * N, dtype=torch.float64).pin_memory()
        st = torch.cuda.current_stream(dev).cuda_stream
        ms = timed(lambda: prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, o.data_ptr(), stream=st), 5, 2) / 5
        capi.check(L.c3sc_vi_batch(prob.handle, vf.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, oh.data_ptr(), None))   # warm-up: staging buffers
        t0 = time.perf_counter()
        for _ in range(5):
            capi.check(L.c3sc_vi_batch(prob.handle, vf.handle, F, dv_h.data_ptr(), fi_h.data_ptr(), N, oh.data_ptr(), None))
        ms_h = (time.perf_counter() - t0) / 5 * 1e3
        out[name] = {"d": cfg.dx, "nodes_per_dim": N, "rank": int(max(ranks)), "n_controls": cfg.nu, "fibers": F,
                     "node_backups_per_s": F * N / (ms * 1e-3), "ms_per_step": ms,
                     "e2e_node_backups_per_s": F * N / (ms_h * 1e-3), "contract_flops_per_node": contract_flops_per_node(cfg, int(max(ranks)))}
        prob.close(); vf.close()
    return out


def parity_check(cfg, rank_ft, prob, vf, dv, fi, out_d, N):
    """64 fibers of the TIMED batch's output (as left in the device buffer by the last step) against the oracle port"""
    from c3sc_b200 import synthetic
    from oracle import pyoracle as po
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import host_problem
    if not os.path.exists(po.PORT_PATH):
        po.build_port()
    ranks = cfg.ranks(rank_ft)
    ft = po.FT(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
    xg, h, hmin, h2, t, olb, oub = host_problem(cfg)
    port = po.Port(cfg, xg, h2, t, olb, oub)
    F = len(dv)
    sel = np.unique(np.linspace(0, F - 1, 64).astype(np.int64))
    got = out_d.cpu().numpy().reshape(F, N)[sel]
    val2, arg = prob.vi_batch(vf, dv[sel], fi[sel])          # the same fibers as a small batch: argmin for the check
    oval, oarg = port.vi_batch(ft, dv[sel], fi[sel])
    worst = 0.0
    for q, f in enumerate(sel):
        _, costs = port.neighbor_costs(ft, dv[f], fi[f])
        sc = np.maximum(np.abs(oval[q, :N]), np.abs(costs).max(axis=1))
        worst = max(worst, float((np.abs(got[q] - oval[q, :N]) / np.where(sc == 0, 1.0, sc)).max()))
    return {"fibers_checked": int(sel.size), "against": "oracle port (oracle/c3sc_oracle.c) on the same inputs",
            "max_rel": worst, "max_rel_scale": "per element: max(|oracle value|, largest |neighbour value| of that node)",
            "max_rel_batch_scale": float(np.abs(got - oval[:, :N]).max() / np.abs(oval).max()),
            "argmin_agree": float((arg == oarg).mean()), "timed_output_equals_small_batch": bool(np.array_equal(got, val2[:, :N])),
            "ok": bool(worst <= 1e-12)}



def cpu_baseline(cfg, rank_ft, budget_s):
    """The reference's CPU path (oracle/_ref when present, else the oracle port) on a bounded
    sample of the same workload, on this box's host cores."""
    from c3sc_b200 import synthetic
    from oracle import pyoracle as po
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import host_problem
    ranks = cfg.ranks(rank_ft)
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    ft = po.FT(cfg.ngrid, ranks, cores)
    ncores = os.cpu_count() or 1
    dv, fi = synthetic.random_fibers(cfg.ngrid, 4096)
    if po.have_ref():
        ref = po.Ref(cfg)
        vf = ref.valuef(ft)
        threads = ref.omp_threads()
        _, s0 = ref.vi_fibers(vf, dv[:8], fi[:8], fresh_per_fiber=False)
        nf = int(max(8, min(4000, budget_s / max(s0 / 8, 1e-6))))
        _, secs = ref.vi_fibers(vf, dv[8:8 + nf], fi[8:8 + nf], fresh_per_fiber=False)
        # single-thread figure: a short sample with OpenMP limited to one thread
        ref.set_omp_threads(1)
        n1 = int(max(4, min(400, 3.0 / max(s0 / 8 * threads, 1e-6))))
        _, s1 = ref.vi_fibers(vf, dv[8:8 + n1], fi[8:8 + n1], fresh_per_fiber=False)
        ref.set_omp_threads(threads)
        one_thread = n1 * cfg.n / s1
        kind = "reference"
        how = "oracle/_ref (reference sources compiled in place), one bellman_vi call per fiber, OpenMP over the nodes of a fiber"
    else:
        if not os.path.exists(po.PORT_PATH):
            po.build_port()
        xg, h, hmin, h2, t, olb, oub = host_problem(cfg)
        port = po.Port(cfg, xg, h2, t, olb, oub)
        threads = ncores
        t0 = time.perf_counter(); port.vi_batch(ft, dv[:32], fi[:32], nthreads=threads); s0 = time.perf_counter() - t0
        nf = int(max(32, min(4000, budget_s / max(s0 / 32, 1e-6))))
        t0 = time.perf_counter(); port.vi_batch(ft, dv[32:32 + nf], fi[32:32 + nf], nthreads=threads); secs = time.perf_counter() - t0
        kind = "port"
        how = "oracle port (C restatement), OpenMP over fibers"
    extra = {"value_1_thread": one_thread} if kind == "reference" else {}
    return {"value": nf * cfg.n / secs, "unit": "node-backups/s", "cores": threads, "kind": kind, **extra,
            "sample": f"{nf} fibers x {cfg.n} nodes of the same workload in {secs:.1f} s; {how}; host has {ncores} cores"}


if __name__ == "__main__":
    main()
