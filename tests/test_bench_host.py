"""Host-side pieces of bench.py that need no GPU: the clock-sample window / summary and the workload description."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def _row(sm, mx, reasons):
    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    return [str(sm), str(mx), "700.0"] + ["Active" if n in reasons else "Not Active" for n in names]


def test_clock_sampler_keeps_only_samples_of_the_timed_region():
    import bench
    s = bench.ClockSampler(0)                      # thread never started: rows are injected
    t = time.time()
    s.all_rows = [(t - 5.0, _row(1965, 1965, ())),                       # warm-up: outside the window
                  (t + 0.1, _row(1710, 1965, ("sw_power_cap",))),
                  (t + 0.3, _row(1695, 1965, ("sw_power_cap",))),
                  (t + 0.5, _row(1725, 1965, ("sw_power_cap",)))]
    s.t0, s.t1 = t, t + 0.6
    s.rows = [r for ts, r in s.all_rows if s.t0 <= ts <= s.t1 + 0.25]
    out = s.summary()
    assert out["samples"] == 3 and out["sm_mhz"] == 1710.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_without_nvidia_smi_reports_unavailable():
    import bench
    s = bench.ClockSampler(0)
    s.mark_start()
    s.mark_stop()                                  # no process, no rows
    assert s.summary()["reasons"] == ["unavailable"]


def test_workload_config_names_the_baseline_configuration():
    import bench
    c = bench.workload_config(8)
    assert c["workload"].startswith("configs[1]") and c["global_batch"] == 8 * c["batch_per_gpu"]
    assert c["parallelism"] == "dp8" and "cache" in c and "model" not in c
