"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): one line per kernel and the last step's sequence.
python tools/launch_summary.py gpurun_out/x/launches.csv [first_kernel_of_a_step]"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] in ("ns", "nsecond") else (v * 1000 if r[ui] in ("ms", "msecond") else v)
    seq.append((k, v))
first = sys.argv[2] if len(sys.argv) > 2 else "tess_count"
idx = [i for i, (k, _) in enumerate(seq) if k.startswith(first)]
if len(idx) >= 3:
    a, b = idx[-3], idx[-2]
    total = sum(v for _, v in seq[a:b])
    print(f"one step: {b - a} launches, {total:.1f} us summed (cold-cache, serialised)")
    for k, v in seq[a:b]:
        print(f"  {k:46s} {v:8.1f} us  {100 * v / total:5.1f} %")
agg = collections.OrderedDict()
for k, v in seq:
    agg.setdefault(k, []).append(v)
print("all launches:")
for k, v in agg.items():
    print(f"  {k:46s} n={len(v):3d} mean {sum(v) / len(v):8.1f} us")
