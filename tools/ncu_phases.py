"""Aggregates the per-line stall samples of an `ncu --page source --csv --print-source cuda,sass` dump by the
device function (of sim_kernel.cu) the line belongs to.  usage: ncu_phases.py dump.csv [top_lines]"""
import csv, re, sys, os
src = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'agarcl_b200', 'csrc', 'sim_kernel.cu')
starts = []
for i, l in enumerate(open(src), 1):
    m = re.match(r'^(?:template.*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(', l)
    if m and not l.startswith('  '): starts.append((i, m.group(1)))
def func_of(line):
    name = 'top'
    for s, n in starts:
        if s <= line: name = n
        else: break
    return name
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] not in ('-', '') and r[2] == '-':
        try: s = int(r[hdr.index("# Samples")])
        except Exception: continue
        agg[(cur, int(r[0]))] = (s, int(r[hdr.index("stall_long_sb")]), int(r[hdr.index("Instructions Executed")]), r[1])
tot = sum(v[0] for v in agg.values())
out = {}
for (f, l), (s, lsb, ie, _) in agg.items():
    name = func_of(l) if f == 'sim_kernel.cu' else 'other:' + f
    o = out.setdefault(name, [0, 0, 0]); o[0] += s; o[1] += lsb; o[2] += ie
for n, (s, lsb, ie) in sorted(out.items(), key=lambda kv: -kv[1][0])[:24]:
    print(f"{n:30s} samples {s:6d} {100*s/tot:5.1f}%  long_sb {lsb:6d}  inst {ie:10d}")
print('total samples', tot)
if len(sys.argv) > 2:
    for (f, l), (s, lsb, ie, srcl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2])]:
        print(f"{f}:{l:5d} {func_of(l) if f=='sim_kernel.cu' else '':22s} samp={s:5d} ({100*s/tot:4.1f}%) lsb={lsb:5d} inst={ie:9d} | {srcl.strip()[:100]}")
