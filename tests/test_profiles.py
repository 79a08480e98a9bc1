"""The committed evidence stays usable: bench.py finds the measured DRAM traffic of every kernel it reports a roofline for, and
the launch-list / summary tools read the committed ncu files."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_roofline_traffic_is_committed_for_the_reported_kernels():
    sys.path.insert(0, ROOT)
    import bench
    for kernel in ("hull_chain_kernel", "raster_tiles_kernel+tile_prims_kernel", "tess_count+scan+emit"):
        traffic = bench.measured_traffic(kernel, 3)
        assert isinstance(traffic, int) and traffic > 1_000_000, kernel
    line = json.load(open(os.path.join(ROOT, "profiles", "bench_r02_c3.json")))
    for r in [line["roofline"]] + line["roofline_other"]:
        assert r["traffic"] == bench.measured_traffic(r["kernel"], 3) or r["traffic"] is not None
        assert 0.0 < r["frac"] < 1.0 and r["peak_kind"] in ("measured", "fallback")
    assert line["e2e"]["h2d_bytes_per_step"] > 20_000_000 and line["gpu_launches"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_launch_summary_reads_the_committed_launch_lists():
    for config in (3, 4, 5):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(ROOT, "profiles", f"launches_r02_c{config}.csv"),
                              "expand" if config != 3 else "tess_count"], capture_output=True, text=True, check=True).stdout
        assert "raster_tiles_kernel" in out and "hull_chain_kernel" in out and "one step:" in out
