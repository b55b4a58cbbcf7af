"""Static SASS size of k_step by the source function each instruction is attributed to (line info).
usage: python tools/sass_size.py   (needs agarcl_b200/libagarcl_b200.so built with -lineinfo)"""
import collections, os, re, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, 'agarcl_b200', 'csrc', 'sim_kernel.cu')
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
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'agarcl_b200', 'libagarcl_b200.so')], cwd=d, stdout=subprocess.DEVNULL)
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, 'sim_kernel.sm_100a.cubin')], capture_output=True, text=True).stdout
cur = None; cnt = collections.Counter(); infunc = False
for l in dis.splitlines():
    if l.startswith('.text.'): infunc = 'k_step' in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1)); ln = int(m.group(2))
        cur = func_of(ln) if f == 'sim_kernel.cu' else 'other:' + f
        continue
    if infunc and re.match(r'\s+/\*[0-9a-f]+\*/\s', l) and cur: cnt[cur] += 1
tot = sum(cnt.values()); print('k_step SASS instructions', tot, f'= {tot*16/1024:.0f} KB')
for k, v in cnt.most_common(45): print(f'{k:30s} {v:6d} {100*v/tot:5.1f}%  {v*16/1024:6.1f} KB')
