"""Host-side pieces of bench.py that need no GPU: the clock sampler (NVML path with a fake NVML, and the fallback when
NVML is unusable) and the thread count of the CPU reference arm under torchrun's OMP_NUM_THREADS=1."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))

import bench  # noqa: E402


class FakeNvml:
    NVML_CLOCK_SM = 1

    def __init__(self):
        self.calls = 0

    def nvmlInit(self):
        pass

    def nvmlDeviceGetHandleByIndex(self, i):
        return ("h", i)

    def nvmlDeviceGetMaxClockInfo(self, h, kind):
        return 1965

    def nvmlDeviceGetClockInfo(self, h, kind):
        self.calls += 1
        return 1200 if self.calls < 3 else 1800  # idle before the mark, loaded after

    def nvmlDeviceGetCurrentClocksEventReasons(self, h):
        return 0x4 if self.calls >= 3 else 0  # sw_power_cap under load


def test_clock_sampler_nvml_path_counts_only_samples_after_mark():
    s = bench.ClockSampler(0, nvml=FakeNvml())
    time.sleep(0.05)
    s.mark()
    time.sleep(0.08)
    out = s.stop()
    assert out["source"] == "nvml" and out["sm_max_mhz"] == 1965.0 and out["samples"] >= 2
    assert out["sm_mhz"] == 1800.0 and out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_falls_back_without_nvml():
    class Broken:
        def nvmlInit(self):
            raise RuntimeError("no driver")

    out = bench.ClockSampler(0, nvml=Broken()).stop()  # nvidia-smi is absent here too: empty but well-formed
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_thread_count_ignores_torchrun_omp_default(monkeypatch):
    import torch

    calls = {}
    monkeypatch.setattr(torch, "set_num_threads", lambda n: calls.setdefault("n", n))

    class Stop(Exception):
        pass

    def boom():
        raise Stop

    monkeypatch.setattr(torch, "get_num_threads", boom)  # stop right after the thread count was chosen
    try:
        bench.cpu_reference_run(1, 0, budget_s=1.0)
    except Stop:
        pass
    import psutil

    assert 1 <= calls["n"] <= (psutil.cpu_count(logical=False) or os.cpu_count())
    assert calls["n"] == min(len(os.sched_getaffinity(0)), psutil.cpu_count(logical=False) or 10**9)
