"""GPU parity of the stage-1 code paths the standard configs do not reach by themselves:
the general (rank > 32) kernel, odd and mixed ranks on the tensor-core path, group sizes that do
not fill a DMMA tile, chunked batches, and the device-buffer + commit route of the cores."""
import numpy as np
import pytest

from c3sc_b200 import capi, configs, synthetic
from oracle import pyoracle as po
from helpers import make_port, rel_err, valid_mask, argmin_mismatches_are_ties, elem_err_costs, elem_err_values

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _check_costs_and_values(cfg, ranks, F, seed=5, face_frac=0.2):
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    ft = po.FT(cfg.ngrid, ranks, cores)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=seed, face_frac=face_frac)
    out = prob.vi_batch_debug(vf, dv, fi)
    oval, oarg = port.vi_batch(ft, dv, fi)
    m = valid_mask(cfg, dv)
    for f in range(len(dv)):
        N = int(cfg.ngrid[dv[f]])
        ab, nv, nf = port.fiber_neighbors(dv[f], fi[f])
        assert np.array_equal(out["absorbed"][f, :N], ab), f
        assert np.array_equal(out["nbr_vary"][f, :N], nv), f
        ec, costs, scale = elem_err_costs(port, ft, dv[f], fi[f], out["costs"][f, :N])      # per element, scale = sum|terms|
        assert ec <= RTOL, (f, ec)
        assert rel_err(out["costs"][f, :N], costs, scale=np.abs(costs).max()) <= RTOL, f
        assert elem_err_values(out["value"][f, :N], oval[f, :N], scale) <= RTOL, f
    assert rel_err(out["value"][m], oval[m], scale=np.abs(oval[m]).max()) <= RTOL
    assert argmin_mismatches_are_ties(cfg, port, ft, dv, fi, out["argmin"], oarg) <= 0.01 * m.sum()
    prob.close(); vf.close()


@pytest.mark.parametrize("name,n,rank", [("dubinscar_new", 14, 40), ("lqgnd_reflect", 10, 36), ("double_int", 48, 40)])
def test_general_kernel_large_ranks(gpu, name, n, rank):
    """ranks above 32 take k_ft_costs (shared-memory staged DFMA kernel), not the DMMA pair"""
    cfg = configs.get_config(name, n=n, rank=rank, dx=4 if name.startswith("lqgnd") else None)
    _check_costs_and_values(cfg, cfg.ranks(), 40)


@pytest.mark.parametrize("rank", [1, 5, 7, 9, 15, 17, 23, 25, 31, 32])
def test_tensor_core_path_rank_sweep(gpu, rank):
    """every DMMA instantiation (RMAX 8/16/24/32), ranks that are not multiples of the 8x8x4 tile"""
    cfg = configs.get_config("skidding5d", n=12, rank=rank)
    _check_costs_and_values(cfg, cfg.ranks(), 24)


def test_mixed_ranks_and_ragged_groups(gpu):
    """different rank per core; fiber counts that leave partial groups for some dim_vary"""
    cfg = configs.get_config("lqgnd", n=9, dx=6)
    ranks = np.array([1, 3, 8, 5, 12, 2, 1], dtype=np.uint64)
    for F in (1, 3, 13, 67):
        _check_costs_and_values(cfg, ranks, F, seed=F)


def test_ragged_grids_and_twelve_dimensions(gpu):
    """different node counts per dimension (ldo = the largest; shorter fibers leave padding entries
    untouched) and the largest instantiated LQG dimension"""
    cfg = configs.get_config("dubinscar_new", n=17, rank=5)
    cfg.nvec = np.array([11, 17, 13], dtype=np.uint64)
    _check_costs_and_values(cfg, cfg.ranks(), 45)
    cfg = configs.get_config("skidding5d", n=12, rank=4)
    cfg.nvec = np.array([9, 12, 8, 10, 11], dtype=np.uint64)
    _check_costs_and_values(cfg, cfg.ranks(), 45)
    cfg = configs.get_config("lqgnd_reflect", n=5, rank=3, dx=12)
    _check_costs_and_values(cfg, cfg.ranks(), 36)


def test_large_batch_is_chunked_consistently(gpu):
    """a batch larger than one pipeline chunk gives the same numbers as its pieces"""
    cfg = configs.get_config("lqgnd_reflect", n=20, rank=6, dx=8)
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    ranks = cfg.ranks()
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    ft = po.FT(cfg.ngrid, ranks, cores)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    F = 60000                                   # 60000 * 20 * 17 * 8 B = 163 MB of cost scratch -> 2 chunks
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=3)
    val, arg = prob.vi_batch(vf, dv, fi)
    for lo, hi in ((0, 500), (29990, 30400), (F - 300, F)):
        v2, a2 = prob.vi_batch(vf, dv[lo:hi], fi[lo:hi])
        # large and small batches take different stage-2 kernels: same numbers to round-off, same argmin
        assert rel_err(val[lo:hi], v2) <= 1e-13
        assert argmin_mismatches_are_ties(cfg, port, ft, dv[lo:hi], fi[lo:hi], arg[lo:hi], a2) <= 1e-3 * a2.size
    prob.close(); vf.close()


def test_tapered_host_chunks_equal_the_equal_chunks(gpu, monkeypatch):
    """The opt-in chunk layouts of the host-buffer entries (api.cu, ChunkLayout: the last chunk of each lane cut into shrinking
    pieces, C3SC_TAPER=1; the first one into growing pieces as well, C3SC_HEAD_CUTS) at the timed workload's geometry: the
    numbers are those of the default equal chunks bit for bit (so every fiber is covered exactly once), per-lane copy
    streams included; a sample of fibers from every piece against the oracle."""
    cfg = configs.get_config("lqgnd_reflect")
    prob = capi.Problem(cfg, arith=1)
    port = make_port(cfg)
    ranks = cfg.ranks()
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    ft = po.FT(cfg.ngrid, ranks, cores)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    F = 32768 + 40                              # two lanes, chunks of ~5.5 k fibers -> tapered; a ragged last piece
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=21)
    monkeypatch.setenv("C3SC_TAPER", "1")
    monkeypatch.setenv("C3SC_HEAD_CUTS", "85,234,506")
    v1, a1 = prob.vi_batch(vf, dv, fi)
    monkeypatch.delenv("C3SC_HEAD_CUTS")
    v2, a2 = prob.vi_batch(vf, dv, fi)
    monkeypatch.delenv("C3SC_TAPER")
    v0, a0 = prob.vi_batch(vf, dv, fi)                                     # the default: equal chunks
    assert np.array_equal(v2, v0) and np.array_equal(a2, a0)
    assert np.array_equal(v1, v0) and np.array_equal(a1, a0)
    m = valid_mask(cfg, dv)
    sub = np.unique(np.concatenate([np.arange(0, F, 997), np.arange(F - 48, F), np.arange(16380, 16420)]))
    oval, oarg = port.vi_batch(ft, dv[sub], fi[sub])
    assert rel_err(v1[sub], oval, scale=np.abs(oval).max()) <= RTOL
    assert argmin_mismatches_are_ties(cfg, port, ft, dv[sub], fi[sub], a1[sub], oarg) <= 1e-3 * oarg.size
    prob.close(); vf.close()


def test_device_buffer_commit(gpu):
    """cores written straight into the library's device buffer (what the NCCL broadcast does)
    take effect after c3sc_valuef_commit"""
    import torch
    cfg = configs.get_config("dubinscar_new", n=16, rank=6)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks()
    cores_a = synthetic.random_cores(cfg.ngrid, ranks, seed=1)
    cores_b = synthetic.random_cores(cfg.ngrid, ranks, seed=2)
    vf = capi.ValueF(cfg.ngrid, ranks, cores_a)
    ref = capi.ValueF(cfg.ngrid, ranks, cores_b)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 50)
    want, _ = prob.vi_batch(ref, dv, fi)
    ptr, cnt = vf.device_buffer()

    class _Arr:
        def __init__(self, p, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (p, False), "version": 3}
    view = torch.as_tensor(_Arr(ptr, cnt), device="cuda")
    flat = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1) for c in cores_b])
    assert flat.size == cnt
    view.copy_(torch.from_numpy(flat).cuda())
    torch.cuda.synchronize()
    vf.commit()
    got, _ = prob.vi_batch(vf, dv, fi)
    assert np.array_equal(got, want)
    prob.close(); vf.close(); ref.close()


def test_fused_gather_stores(gpu):
    """c3sc_batch_out::value_peers: the control kernels store each value into every listed buffer at
    peer_offset (the fused all-gather; here both 'peers' live on the one device)"""
    import torch
    cfg = configs.get_config("lqgnd_reflect", n=12, rank=5, dx=4)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks()
    vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
    F, N = 300, cfg.n
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=11)
    want, _ = prob.vi_batch(vf, dv, fi)
    dv_d = torch.from_numpy(np.ascontiguousarray(dv)).cuda()
    fi_d = torch.from_numpy(np.ascontiguousarray(fi)).cuda()
    off = 2 * F * N
    bufs = [torch.full((4 * F * N,), -7.0, dtype=torch.float64, device="cuda") for _ in range(2)]
    for own in (True, False):                     # with and without the rank's private value buffer
        out = torch.zeros(F * N, dtype=torch.float64, device="cuda")
        prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, out.data_ptr() if own else 0,
                          peers=[b.data_ptr() for b in bufs], peer_offset=off)
        torch.cuda.synchronize()
        for b in bufs:
            got = b.cpu().numpy()
            assert np.array_equal(got[off:off + F * N].reshape(F, N), want)
            assert (got[:off] == -7.0).all() and (got[off + F * N:] == -7.0).all()
            b.fill_(-7.0)
        if own:
            assert np.array_equal(out.cpu().numpy().reshape(F, N), want)
    prob.close(); vf.close()


@pytest.mark.parametrize("name,n,rank,dx,F", [("lqgnd", 10, 4, 6, 700), ("lqgnd_reflect", 7, 3, 10, 300), ("lqg2d_new", 40, 5, None, 60),
                                               ("double_int", 30, 4, None, 50), ("dubinscar_new", 14, 4, None, 120)])
def test_grid_walk_equals_table_walk(gpu, name, n, rank, dx, F):
    """full {lo, 0, hi}^du control grids take the shared-prefix walk (k_control_grid); it must give the
    brute-force walk's value and argmin -- checked against the oracle and against the library's own
    grouped walk (C3SC_NO_GRID=1), VI and the PI pair (policy rows from the argmin)"""
    import os
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=1)
    assert capi.lib().c3sc_problem_control_path(prob.handle) == 2
    port = make_port(cfg)
    ranks = cfg.ranks()
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    cores2 = synthetic.random_cores(cfg.ngrid, ranks, seed=0xABCD00)
    ft, ft2 = po.FT(cfg.ngrid, ranks, cores), po.FT(cfg.ngrid, ranks, cores2)
    vf, vf2 = capi.ValueF(cfg.ngrid, ranks, cores), capi.ValueF(cfg.ngrid, ranks, cores2)
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=9, face_frac=0.1)
    m = valid_mask(cfg, dv)
    val, arg = prob.vi_batch(vf, dv, fi)
    oval, oarg = port.vi_batch(ft, dv, fi, nthreads=4)
    assert rel_err(val[m], oval[m], scale=np.abs(oval[m]).max()) <= RTOL
    assert argmin_mismatches_are_ties(cfg, port, ft, dv, fi, arg, oarg) <= 0.01 * m.sum()
    p1, rows, _ = prob.pi_batch(vf, vf2, dv, fi)
    o1, orows, _ = port.pi_batch(ft, ft2, dv, fi)
    assert rel_err(p1[m], o1[m], scale=np.abs(o1[m]).max()) <= RTOL
    os.environ["C3SC_NO_GRID"] = "1"
    try:
        val2, arg2 = prob.vi_batch(vf, dv, fi)
    finally:
        del os.environ["C3SC_NO_GRID"]
    assert rel_err(val[m], val2[m], scale=np.abs(val2[m]).max()) <= 1e-13
    assert argmin_mismatches_are_ties(cfg, port, ft, dv, fi, arg, arg2) <= 1e-3 * m.sum()
    prob.close(); vf.close(); vf2.close()


def test_pinned_result_buffer_is_written_in_place(gpu, monkeypatch):
    """c3sc_vi_batch with C3SC_ZEROCOPY=1 and a page-locked (device-mapped) result buffer: the control
    kernels store into it directly; same numbers as the staged copy into pageable memory"""
    import torch
    monkeypatch.setenv("C3SC_ZEROCOPY", "1")
    cfg = configs.get_config("lqgnd_reflect", n=14, rank=5, dx=4)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks()
    vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
    F, N = 9000, cfg.n                                         # > 2 * 148 * 1024 nodes / 14: both lanes
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=21)
    dv = np.ascontiguousarray(dv, dtype=np.int32); fi = np.ascontiguousarray(fi, dtype=np.int32)
    want, warg = prob.vi_batch(vf, dv, fi)                     # numpy (pageable) buffers
    out = torch.full((F * N,), -3.0, dtype=torch.float64).pin_memory()
    arg = torch.full((F * N,), -9, dtype=torch.int32).pin_memory()
    for with_arg in (False, True):
        out.fill_(-3.0)
        capi.check(capi.lib().c3sc_vi_batch(prob.handle, vf.handle, F, dv.ctypes.data, fi.ctypes.data, N,
                                            out.data_ptr(), arg.data_ptr() if with_arg else None))
        assert np.array_equal(out.numpy().reshape(F, N), want)
        if with_arg:
            assert np.array_equal(arg.numpy().reshape(F, N), warg)
    prob.close(); vf.close()


@pytest.mark.parametrize("case", ["skid_r7", "lqg6_mixed", "dubins_r12", "lqg10_r20", "ragged", "double_int"])
def test_bucketed_chain_stage_equals_oracle(gpu, monkeypatch, case):
    """the bucketed chain stage (chain_kernel.cuh: counting sort by core block + streaming DMMA steps) forced on small
    batches (it normally starts at 4096 fibers): same flags, indices, neighbour values, values and argmin as the oracle,
    odd / mixed ranks, ragged grids, every boundary type, one and several super-chunks"""
    monkeypatch.setenv("C3SC_CHAIN_MIN", "1")
    if case == "skid_r7":
        cfg = configs.get_config("skidding5d", n=12, rank=7); _check_costs_and_values(cfg, cfg.ranks(), 150)
    elif case == "lqg6_mixed":
        cfg = configs.get_config("lqgnd", n=9, dx=6)
        for F in (1, 13, 67, 400):
            _check_costs_and_values(cfg, np.array([1, 3, 8, 5, 12, 2, 1], dtype=np.uint64), F, seed=F)
    elif case == "dubins_r12":
        cfg = configs.get_config("dubinscar_new", n=14, rank=12); _check_costs_and_values(cfg, cfg.ranks(), 200)
    elif case == "lqg10_r20":
        monkeypatch.setenv("C3SC_CHAIN_FIBERS", "96")                       # 48-fiber chunks, 2 per super-chunk, 4 super-chunks
        monkeypatch.setenv("C3SC_CHUNK_FIBERS", "48")
        cfg = configs.get_config("lqgnd_reflect", n=16, rank=20, dx=10); _check_costs_and_values(cfg, cfg.ranks(), 300, face_frac=0.3)
    elif case == "ragged":
        cfg = configs.get_config("skidding5d", n=12, rank=4)
        cfg.nvec = np.array([9, 12, 8, 10, 11], dtype=np.uint64)
        _check_costs_and_values(cfg, cfg.ranks(), 90)
    else:
        cfg = configs.get_config("double_int", n=30, rank=9); _check_costs_and_values(cfg, cfg.ranks(), 120)


def test_bucketed_and_per_fiber_chain_stages_agree(gpu, monkeypatch):
    """the two implementations of stage 1a on the same batch: values agree to round-off (different association of the
    same products), flags and argmin identical up to ties"""
    cfg = configs.get_config("lqgnd_reflect", n=20, rank=11, dx=8)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks()
    cores = synthetic.random_cores(cfg.ngrid, ranks)
    vf = capi.ValueF(cfg.ngrid, ranks, cores)
    dv, fi = synthetic.random_fibers(cfg.ngrid, 5000, seed=4)
    monkeypatch.setenv("C3SC_NO_BUCKETS", "1")
    v0, a0 = prob.vi_batch(vf, dv, fi)
    monkeypatch.delenv("C3SC_NO_BUCKETS")
    v1, a1 = prob.vi_batch(vf, dv, fi)
    assert rel_err(v1, v0) <= 1e-13
    assert (a0 != a1).mean() <= 1e-4
    prob.close(); vf.close()


def test_commit_on_a_side_stream_is_waited_for_before_the_cores_are_read(gpu):
    """c3sc_valuef_commit on a stream of its own (bench.py at N > 1: broadcast + commit beside the batch's plan): the value
    function remembers the event of its last commit and a c3sc_vi_batch_dev on ANOTHER stream waits for it before its first kernel
    that reads the cores.  New cores are written into the device buffer on a side stream behind a long dummy kernel queue, then
    committed there; the batch on the main stream must see them -- for the bucketed and the per-fiber chain stage"""
    import torch
    cfg = configs.get_config("lqgnd_reflect", n=16, rank=9, dx=6)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks()
    c0 = synthetic.random_cores(cfg.ngrid, ranks, seed=1)
    c1 = synthetic.random_cores(cfg.ngrid, ranks, seed=2)
    dev = torch.device("cuda", 0)
    for F in (300, 4500):                                # per-fiber chains / bucketed chains
        dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=F)
        ref = capi.ValueF(cfg.ngrid, ranks, c1)
        want, _ = prob.vi_batch(ref, dv, fi)
        ref.close()
        vf = capi.ValueF(cfg.ngrid, ranks, c0)
        ptr, count = vf.device_buffer()
        flat = torch.from_numpy(np.concatenate([np.asarray(c).reshape(-1) for c in c1])).to(dev)
        assert flat.numel() == count
        dv_d = torch.from_numpy(np.ascontiguousarray(dv)).to(dev); fi_d = torch.from_numpy(np.ascontiguousarray(fi)).to(dev)
        out = torch.zeros(F * cfg.n, dtype=torch.float64, device=dev)
        main = torch.cuda.current_stream(dev); side = torch.cuda.Stream(dev)
        dst = torch.as_tensor(_DevArr(ptr, count), device=dev)
        big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            for _ in range(20):
                big.mul_(1.0001)                         # keeps the side stream busy: the commit below completes late
            dst.copy_(flat)
            vf.commit(stream=side.cuda_stream)
        prob.vi_batch_dev(vf, F, dv_d.data_ptr(), fi_d.data_ptr(), cfg.n, out.data_ptr(), stream=main.cuda_stream)
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(F, cfg.n)
        assert np.array_equal(got, np.asarray(want).reshape(F, cfg.n))
        vf.close()
    prob.close()


@pytest.mark.parametrize("name,n,rank,dx,F", [("lqgnd_reflect", 20, 9, 8, 5000), ("lqgnd", 16, 5, 6, 700), ("skidding5d", 12, 4, None, 900)])
def test_pi_first_visit_in_one_pass_equals_two_passes(gpu, monkeypatch, name, n, rank, dx, F):
    """bellman_pi's first visit (improvement against the policy function, then evaluation of the new rows against the iterate,
    src/bellman.c:1831-1871) when both are the SAME value function: the evaluation takes the neighbour values from the chunk's
    scratch of the improvement instead of a second stage 1 -- rows, argmin and values bit-identical to the two-pass form"""
    import torch
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx) if dx else configs.get_config(name, n=n, rank=rank)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks()
    vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=3)
    dev = torch.device("cuda", 0)
    dv_d = torch.from_numpy(np.ascontiguousarray(dv)).to(dev); fi_d = torch.from_numpy(np.ascontiguousarray(fi)).to(dev)
    N = cfg.n; RW = 2 * cfg.dx + 3
    res = []
    for two in (False, True):
        if two:
            monkeypatch.setenv("C3SC_PI_TWO_PASSES", "1")
        rows = torch.zeros(F * N * RW, dtype=torch.float64, device=dev); val = torch.zeros(F * N, dtype=torch.float64, device=dev)
        arg = torch.zeros(F * N, dtype=torch.int32, device=dev)
        prob.pi_batch_dev(vf, vf, F, dv_d.data_ptr(), fi_d.data_ptr(), N, 0, rows.data_ptr(), arg.data_ptr(), val.data_ptr())
        torch.cuda.synchronize()
        res.append((rows.cpu().numpy(), val.cpu().numpy(), arg.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][2], res[1][2])
    assert np.array_equal(res[0][1], res[1][1])
    assert np.abs(res[0][1]).max() > 0
    prob.close(); vf.close()


class _DevArr:
    """a device buffer of the library as a __cuda_array_interface__ object (float64)"""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


@pytest.mark.parametrize("name,n,rank,dx,F", [("lqgnd_reflect", 12, 7, 10, 2600), ("lqgnd", 16, 5, 6, 3000), ("dubinscar_new", 24, 6, None, 2600),
                                               ("double_int", 40, 5, None, 2500), ("skidding5d", 12, 4, None, 2600)])
def test_fused_stage2_equals_the_two_kernel_pipeline(gpu, monkeypatch, name, n, rank, dx, F):
    """C3SC_FUSE=1: the node kernel keeps its fibers' neighbour values in a region of a small ring and runs the control walk
    (or the policy evaluation) itself through a cross-translation-unit device call.  Opt-in (measured slower on B200), kept
    correct: value iteration, the improvement step with rows and argmin, and a sub-iteration give the same numbers."""
    cfg = configs.get_config(name, n=n, rank=rank, dx=dx)
    prob = capi.Problem(cfg, arith=1)
    ranks = cfg.ranks()
    vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
    vf2 = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks, seed=0xABCD00))
    dv, fi = synthetic.random_fibers(cfg.ngrid, F, seed=13, face_frac=0.2)
    v0, a0 = prob.vi_batch(vf, dv, fi)
    p0, rows0, pa0 = prob.pi_batch(vf, vf2, dv, fi)
    q0, _, _ = prob.pi_batch(None, vf, dv, fi, rows=rows0)
    monkeypatch.setenv("C3SC_FUSE", "1")
    n0 = capi.lib().c3sc_launch_count()
    v1, a1 = prob.vi_batch(vf, dv, fi)
    launches = capi.lib().c3sc_launch_count() - n0
    p1, rows1, pa1 = prob.pi_batch(vf, vf2, dv, fi)
    q1, _, _ = prob.pi_batch(None, vf, dv, fi, rows=rows1)
    assert rel_err(v1, v0) <= 1e-14 and np.array_equal(a1, a0)
    assert rel_err(p1, p0) <= 1e-14 and np.array_equal(pa1, pa0) and rel_err(rows1, rows0, scale=1.0) <= 1e-14
    assert rel_err(q1, q0) <= 1e-14
    if name != "skidding5d":                        # a grid-structured control set: no control kernel was launched
        assert launches == 3, launches              # grouping + chains + fused node kernel
    prob.close(); vf.close(); vf2.close()


def test_guard_zones_stay_intact_and_nothing_reads_them(gpu):
    """the library's own memcheck (compute-sanitizer is closed on this GPU pool, profiles/r02_sanitizer_closed.log): in a fresh
    process with C3SC_GUARD=1 every device allocation sits between 0xFF-filled zones; the parity cases of every kernel family
    (both chain stages, tensor-core node kernel at odd / mixed ranks, general kernel, grid / grouped / table walks, PI, fused
    stage 2, off-grid entries, ragged grids) must still match the oracle -- an out-of-bounds read that reached a result would
    be NaN -- and no zone may be damaged afterwards (an out-of-bounds write)."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, %r)
        import numpy as np
        from c3sc_b200 import capi, configs, synthetic
        import test_gpu_paths as T
        L = capi.lib(); capi.check(L.c3sc_cuda_init(0))
        for rank in (1, 7, 20, 31):
            cfg = configs.get_config("skidding5d", n=12, rank=rank); T._check_costs_and_values(cfg, cfg.ranks(), 24)
        cfg = configs.get_config("lqgnd", n=9, dx=6)
        T._check_costs_and_values(cfg, np.array([1, 3, 8, 5, 12, 2, 1], dtype=np.uint64), 67, seed=67)
        cfg = configs.get_config("dubinscar_new", n=14, rank=40); T._check_costs_and_values(cfg, cfg.ranks(), 40)      # general kernel
        cfg = configs.get_config("skidding5d", n=12, rank=4); cfg.nvec = np.array([9, 12, 8, 10, 11], dtype=np.uint64)
        T._check_costs_and_values(cfg, cfg.ranks(), 45)                                                                # ragged
        os.environ["C3SC_CHAIN_MIN"] = "1"                                                                             # bucketed chains
        cfg = configs.get_config("lqgnd_reflect", n=16, rank=20, dx=10); T._check_costs_and_values(cfg, cfg.ranks(), 300, face_frac=0.3)
        cfg = configs.get_config("dubinscar_new", n=14, rank=12); T._check_costs_and_values(cfg, cfg.ranks(), 200)
        del os.environ["C3SC_CHAIN_MIN"]
        os.environ["C3SC_FUSE"] = "1"                                                                                  # fused stage 2
        cfg = configs.get_config("lqgnd_reflect", n=12, rank=7, dx=10)
        prob = capi.Problem(cfg, arith=1); ranks = cfg.ranks()
        vf = capi.ValueF(cfg.ngrid, ranks, synthetic.random_cores(cfg.ngrid, ranks))
        dv, fi = synthetic.random_fibers(cfg.ngrid, 2600, seed=13)
        v1, a1 = prob.vi_batch(vf, dv, fi); p1, rows, _ = prob.pi_batch(vf, vf, dv, fi)
        del os.environ["C3SC_FUSE"]
        v0, a0 = prob.vi_batch(vf, dv, fi)
        assert np.isfinite(v1).all() and np.isfinite(rows).all() and np.array_equal(a0, a1) and np.abs(v1 - v0).max() <= 1e-14 * np.abs(v0).max()
        x = cfg.lb + synthetic.uniform01(3, 50 * cfg.dx).reshape(50, cfg.dx) * (cfg.ub - cfg.lb)
        u, val, ab, costs = prob.policy_eval(vf, x)
        assert np.isfinite(val).all() and np.isfinite(costs).all()
        prob.close(); vf.close()
        print("GUARD_DAMAGED", L.c3sc_guard_check())
    """ % (root, root))
    env = dict(os.environ, C3SC_GUARD="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GUARD_DAMAGED 0" in r.stdout, r.stdout[-500:]
