"""examples/lqg2d_b200.c: an examples/lqg2d_new-style main() compiled with gcc against include/c3sc_host.h
and linked to the C-ABI library -- the reference's own call sequence on the GPU path."""
import os
import re
import subprocess
import ctypes as C

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "lqg2d_b200")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "c3sc_b200", "lib")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "lqg2d_b200.c"), "-L" + lib, "-lc3sc_b200", "-Wl,-rpath," + lib, "-lm", "-o", EXE]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_example_compiles_against_the_host_header_and_fails_loudly_without_a_gpu(built):
    from c3sc_b200 import capi
    _build()
    if capi.lib().c3sc_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "16", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_example_program_runs_the_reference_call_sequence(gpu):
    from c3sc_b200 import capi, configs
    from test_host_api import solver_lib, HostProblem, vp, sz, dbl
    _build()
    r = subprocess.run([EXE, "24", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"RESULT norm (\S+) u0 (\S+)", r.stdout)
    assert m and len(re.findall(r"^outer \d+:", r.stdout, flags=re.M)) == 3
    norm, u0 = float(m.group(1)), float(m.group(2))
    assert u0 in (-1.0, 0.0, 1.0) and np.isfinite(norm) and norm > 0
    # the same flow driven through ctypes gives the same number
    L = solver_lib()
    cfg = configs.get_config("lqg2d_reflect", n=24, rank=3)
    hp = HostProblem(L, cfg, arith=1)
    START = C.CFUNCTYPE(C.c_int, sz, C.POINTER(dbl), C.POINTER(dbl), vp)

    def _start(n, x, out, _):
        xs = np.ctypeslib.as_array(x, shape=(n * 2,)).reshape(n, 2)
        np.ctypeslib.as_array(out, shape=(n,))[:] = xs[:, 0] * xs[:, 0] + xs[:, 1] * xs[:, 1]
        return 0
    start = START(_start)
    a = L.approx_args_init()
    L.approx_args_set_cross_tol(a, 1e-8); L.approx_args_set_round_tol(a, 1e-7); L.approx_args_set_kickrank(a, 2)
    L.approx_args_set_startrank(a, 3); L.approx_args_set_maxrank(a, 12); L.approx_args_set_adapt(a, 1)
    cost = L.c3control_init_value(hp.c3c, start, None, a, 0)
    for _ in range(3):
        nxt = L.c3control_pi_solve(hp.c3c, 5, 1e-7, cost, a, hp.opt, 0, None)
        tmp = L.c3control_vi_solve(hp.c3c, 1, 1e-7, nxt, a, hp.opt, 0, None)
        L.valuef_destroy(nxt); L.valuef_destroy(cost)
        cost = tmp
    assert abs(L.valuef_norm(cost) - norm) <= 1e-9 * norm
    L.valuef_destroy(cost); L.approx_args_free(a); hp.close()


@pytest.mark.gpu
def test_example_program_resumes_from_its_checkpoint(gpu, tmp_path):
    """second run with the same checkpoint file continues where the first stopped: its first outer
    iteration moves the function less than the first run's last one did"""
    _build()
    ck = str(tmp_path / "cost.c3sc")
    r1 = subprocess.run([EXE, "20", "3", ck], capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and os.path.exists(ck), r1.stderr
    r2 = subprocess.run([EXE, "20", "1", ck], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0 and "resumed from" in r2.stdout, r2.stderr
    d1 = [float(x) for x in re.findall(r"diff (\S+)", r1.stdout)]
    d2 = [float(x) for x in re.findall(r"diff (\S+)", r2.stdout)]
    assert len(d1) == 3 and len(d2) == 1 and d2[0] < d1[-1]
    n1 = float(re.search(r"RESULT norm (\S+)", r1.stdout).group(1))
    n2 = float(re.search(r"RESULT norm (\S+)", r2.stdout).group(1))
    assert n2 > n1                      # the cost-to-go keeps growing towards its fixed point
