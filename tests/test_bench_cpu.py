"""Host-side checks of bench.py's bookkeeping (no GPU): both arms describe the workload identically, and the ncu traffic
figure is taken from the newest committed capture only while the kernel sources it was taken from are unchanged."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_both_arms_print_the_same_config():
    b = _bench()
    for name, (W, H, r, sd, n) in b.WORKLOADS.items():
        cfg = b.workload_config(name, W, H, r, sd, n)
        assert cfg["workload"].startswith("synthetic %s statistic buffers (%dx%d" % (name, W, H))
        assert set(cfg) == {"workload", "width", "height", "radius", "sd", "spp"}
    # the float32 range factors of the reference: -.5f / (sd * sd) in float32 arithmetic (estimator.cpp:16)
    assert b.f32_factor(0.1) != -50.0 and abs(b.f32_factor(0.1) + 50.0) < 1e-5
    assert b.f32_factor(10.0) == -0.004999999888241291


def test_ncu_traffic_is_refused_when_the_kernel_changed(tmp_path, monkeypatch):
    b = _bench()
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    csrc = tmp_path / "statmc_b200" / "csrc"
    csrc.mkdir(parents=True)
    for f in b.KERNEL_SOURCES["filter"]:
        (csrc / f).write_text("// " + f)
    assert b.ncu_traffic("filter")["traffic"] is None                      # no capture at all
    sha = b.source_sha("filter")
    body = "dram__bytes_read.sum   Gbyte   1.5\ndram__bytes_write.sum  Mbyte   250\n"
    (prof / "r2_ncu_filter.txt").write_text("# source_sha[filter]: %s\n%s" % (sha, body))
    (prof / "r1e_ncu_filter.txt").write_text(body)                         # older capture without a sha: ignored (r2 sorts later)
    t = b.ncu_traffic("filter")
    assert t["traffic"] == 1.5e9 + 250e6 and t["traffic_source"].endswith("r2_ncu_filter.txt")
    assert b.ncu_traffic("filter", applicable=False)["traffic"] is None    # other workload / N: capture does not apply
    (csrc / b.KERNEL_SOURCES["filter"][0]).write_text("// changed")
    t = b.ncu_traffic("filter")
    assert t["traffic"] is None and "stale" in t["traffic_note"]
    (prof / "r10_ncu_filter.txt").write_text("# source_sha[filter]: %s\n%s" % (b.source_sha("filter"), body))
    assert b.ncu_traffic("filter")["traffic_source"].endswith("r10_ncu_filter.txt")   # natural sort: r10 after r2
