"""Attribute ncu SASS-level samples to CUDA source lines. usage: hot.py <ncu sass csv> <cubin> <kernel substring> [top]"""
import csv, re, subprocess, sys, collections
sass_csv, cubin, kern = sys.argv[1:4]
SORT = 1 if "--inst" in sys.argv else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ai, si, ii = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
insts = [(r[ai], float(r[si] or 0), float(r[ii] or 0), r[1]) for r in rows[2:] if len(r) == len(hdr)]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
# find function section
lines = []; cur_line = None; infunc = False; cur_file=None
for l in dis:
    if l.startswith('.text.') or '.section' in l and '.text.' in l:
        infunc = kern in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur_file, cur_line = m.group(1).split('/')[-1], int(m.group(2)); continue
    if infunc and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append((cur_file, cur_line, l.strip()))
print(len(insts), 'sampled sass instructions;', len(lines), 'disassembled')
n = min(len(insts), len(lines))
agg = collections.defaultdict(lambda: [0.0, 0.0])
for k in range(n):
    f, ln, _ = lines[k]
    agg[(f, ln)][0] += insts[k][1]; agg[(f, ln)][1] += insts[k][2]
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
srcs = {}
def src(f, ln):
    import glob
    if f not in srcs:
        c = glob.glob('/root/repo/contrast_renderer_b200/csrc/**/' + f, recursive=True)
        srcs[f] = open(c[0]).read().splitlines() if c else []
    return srcs[f][ln - 1].strip()[:100] if ln and ln <= len(srcs[f]) else ''
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
    print(f"{f}:{ln:<5} samples {100*v[0]/ts:5.1f}%  inst {100*v[1]/ti:5.1f}%   {src(f, ln)}")
