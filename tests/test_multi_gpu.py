"""include/c3sc_multi.h: one process drives every GPU of the box.  The C program
examples/multi_gpu_b200.c checks the sharded calls bit for bit against one device; with one visible GPU it still
runs the whole multi-device code path with G = 1 (worker thread, shard arithmetic, resident rows)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "multi_gpu_b200")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "c3sc_b200", "lib")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "multi_gpu_b200.c"), "-L" + lib, "-lc3sc_b200", "-Wl,-rpath," + lib,
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64", "-lm", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shard_map_is_contiguous_and_stable():
    """c3sc_multi_shard: contiguous blocks of ceil(F/G), a function of (F, G) only (policy rows never move)"""
    import ctypes as C
    from c3sc_b200 import capi
    L = capi.lib()
    for F in (0, 1, 7, 400, 6480, 65536):
        for G in (1, 2, 3, 4, 8):
            edges = []
            for g in range(G):
                b, e = C.c_size_t(), C.c_size_t()
                L.c3sc_multi_shard(F, G, g, C.byref(b), C.byref(e))
                edges.append((b.value, e.value))
            assert edges[0][0] == 0 and edges[-1][1] == F
            assert all(edges[g][1] == edges[g + 1][0] for g in range(G - 1))
            per = -(-F // G) if F else 0
            assert all(e - b <= per for b, e in edges)
            assert L.c3sc_multi_gathered_count(F, G, 10) == per * G * 10


def test_multi_program_compiles_and_fails_loudly_without_a_gpu(built):
    from c3sc_b200 import capi
    _build()
    if capi.lib().c3sc_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("args", [("0", "4000", "6", "24", "6"), ("0", "9000", "4", "30", "11"), ("0", "300", "10", "12", "5")])
def test_c_program_drives_all_gpus(gpu, args):
    """sharded bellman_vi / bellman_pi (resident rows) / gathered values / a cross step == one device, bit for bit"""
    _build()
    r = subprocess.run([EXE, *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"OK devices (\d+) nccl (\d) fibers (\d+)", r.stdout)
    assert m and int(m.group(1)) == gpu.c3sc_cuda_device_count() and int(m.group(3)) == int(args[1])
    if int(m.group(1)) > 1:
        assert m.group(2) == "1", "NCCL not used on a multi-GPU box"


@pytest.mark.gpu
def test_multi_without_nccl_uses_peer_copies(gpu):
    """C3SC_NO_NCCL=1: cores and gathered values travel by cudaMemcpyPeerAsync; same numbers"""
    _build()
    env = dict(os.environ, C3SC_NO_NCCL="1")
    r = subprocess.run([EXE, "0", "2000", "4", "20", "5"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert re.search(r"OK devices \d+ nccl 0", r.stdout)
