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


def test_flop_model_and_edge_counts_follow_survey_8d():
    import bench
    assert [bench.band_edges(L) for L in (64, 100, 256, 512, 1024)] == [3480, 6360, 18840, 39320, 80280]
    assert bench.band_edges(1024, 1023) == 1047552
    assert abs(bench.flops_per_conformer(256, 1, train=False) / 1e9 - 5.116) < 0.01         # per conformer-layer forward
    assert abs(bench.flops_per_conformer(256, 6) / 1e9 - 92.1) < 0.1                        # config 2, fwd + bwd
    assert abs(bench.flops_per_conformer(100, 8, train=False) / 1e9 - 13.9) < 0.1           # config 4, per sample
    for cfg in ("train", "decode", "mixed", "stress"):
        c = bench.workload_config(4, cfg)
        assert c["workload"].startswith("configs[") and c["parallelism"] == "dp4" and "model" not in c


def test_reference_arm_runs_on_a_tiny_sample(capsys):
    """`--impl reference` prints one JSON line with the contract's keys (tiny shapes so that the CPU suite stays short)."""
    import json
    import types
    import bench
    old = dict(bench.CFG)
    bench.CFG.update(L=24, layers=1, z_g=16, z_l=8)
    try:
        bench.run_reference(types.SimpleNamespace(gpus=1, steps=1, warmup=1, config="train"))
    finally:
        bench.CFG.clear()
        bench.CFG.update(old)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == "train_conformers_per_s"
