"""Prints the headline numbers of bench.py JSON lines: python tools/bench_summary.py gpurun_out/dir/*.json"""
import json
import sys

for path in sys.argv[1:]:
    try:
        lines = [l for l in open(path) if l.startswith("{")]
        d = json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    e2e = d.get("e2e", {})
    print(f"{path}: value {d['value'] / 1e6:.2f} M/s  {d['ms_per_step']:.3f} ms/step (one at a time {d.get('ms_per_step_one_at_a_time')})"
          f"  e2e {e2e.get('value', 0) / 1e6:.2f} M/s {e2e.get('ms_per_step')} ms  launches {d.get('gpu_launches')}")
    print("    kernel_ms", {k: round(v, 4) for k, v in (d.get("kernel_ms_per_step") or {}).items()})
    r = d.get("roofline") or {}
    print("    roofline", r.get("kernel"), "frac", r.get("frac"), "ms", r.get("kernel_ms"), "traffic", r.get("traffic"))
    if d.get("e2e_with_frame"):
        f = d["e2e_with_frame"]
        print(f"    e2e_with_frame {f['value'] / 1e6:.2f} M/s {f['ms_per_step']:.3f} ms")
    if d.get("cpu_baseline"):
        print("    cpu_baseline", d["cpu_baseline"]["value"])
