"""The parts of bench.py's JSON contract that can be checked without a GPU: the `roofline` object
(top level = the dominant kernels timed alone, SURVEY 8(d)'s whole-step figure kept beside it) and the
clock sampler's accounting of samples answered inside the timed window."""
import sys
import time
import types

import pytest

import bench


STAGE1 = {"kernels": "k_ft_chains + k_ft_nodes", "fibers": 8192, "ms": 0.283, "flops_per_node": 3392.0,
          "achieved": 8192 * 100 * 3392.0 / 0.283e-3 / 1e12, "unit": "TFLOP/s"}


def test_roofline_top_level_is_the_dominant_kernels():
    r = bench.roofline_record(STAGE1, 135.0, 37.2, 33.7, 47132.0, 3392.0, 65536, 2.29, 57904384.0, 6552.0)
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert r["peak"] == 37.2                                       # the DMMA figure is the denominator of a tensor-bound kernel
    assert r["peaks_measured_in_this_run"] == pytest.approx({"fp64_dmma_tflops": 37.2, "fp64_dfma_tflops": 33.7}) or \
        r["peaks_measured_in_this_run"]["fp64_dfma_tflops"] == 33.7
    assert r["achieved"] == pytest.approx(STAGE1["achieved"]) and r["frac"] == pytest.approx(STAGE1["achieved"] / 37.2)
    assert r["frac_of_dfma_peak"] == pytest.approx(STAGE1["achieved"] / 33.7)
    assert 0.0 < r["frac"] < 1.0                                   # a utilisation, not the contract algebra
    whole = r["contract_whole_step"]
    assert whole["frac"] == pytest.approx(135.0 / 37.2) and whole["flops_per_node_backup"] == 47132.0
    assert r["hbm"]["frac"] == pytest.approx(57904384.0 / 2.29e-3 / 1e9 / 6552.0)
    assert r["stage1_live"]["achieved"] == r["achieved"]


def test_roofline_without_a_stage1_timing_says_so():
    r = bench.roofline_record(None, 135.0, 37.2, 33.7, 47132.0, 3392.0, 65536, 2.29, 57904384.0, 6552.0)
    assert r["stage1_live"] is None and "whole step" in r["basis"]
    assert r["frac"] == pytest.approx(135.0 / 37.2)


def test_contract_flops_match_the_survey_examples():
    from c3sc_b200 import configs
    cfg = configs.get_config("lqgnd_reflect", n=100, rank=20, dx=10)
    w = bench.contract_flops_per_node(cfg, 20)
    assert 45e3 < w < 50e3                                         # 47.7 kflop with r^2 = 400 everywhere


def test_clock_sampler_counts_samples_inside_the_window(monkeypatch):
    m = types.ModuleType("pynvml")
    m.NVML_CLOCK_SM = 1
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda i: object()
    m.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965

    def query(h, c):
        time.sleep(0.01)
        return 1950
    m.nvmlDeviceGetClockInfo = query
    m.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4        # sw_power_cap
    monkeypatch.setitem(sys.modules, "pynvml", m)
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.03)
    seen, t0 = s.count(), time.perf_counter()
    time.sleep(0.08)
    t1 = time.perf_counter()
    inside = s.count_between(t0, t1, seen)
    time.sleep(0.03)
    out = s.stop()
    assert 3 <= inside <= 8 and inside < out["samples"]
    assert out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
