"""Instructions executed and stall samples per CUDA source line of one kernel, from
`ncu -i report.ncu-rep --page source --csv --print-source sass,cuda > file.csv` (compile with -lineinfo).
python tools/ncu_lines.py file.csv [top_n]"""
import csv
import sys


def num(x):
    try:
        return int(float(x.replace(",", "")))
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, h, idx, data = "?", None, {}, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        h, idx = r, {}
        for i, n in enumerate(r):
            idx.setdefault(n, i)
        continue
    if h is None or len(r) < len(h) or r[0] == "":
        continue
    key = (cur.split("/")[-1], r[0])
    ie, s = num(r[idx["Instructions Executed"]]), num(r[idx["# Samples"]])
    if key in data:
        data[key][0] += ie
        data[key][1] += s
    else:
        data[key] = [ie, s, r[1].strip()[:110]]
tot = sum(v[0] for v in data.values()) or 1
ts = sum(v[1] for v in data.values()) or 1
print(f"warp instructions {tot}, stall samples {ts}")
for k, v in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]:11d} {100 * v[0] / tot:5.1f} % inst  {100 * v[1] / ts:5.1f} % samples  {k[0]}:{k[1]}  {v[2]}")
