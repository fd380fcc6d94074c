"""Per-source-line instruction counts of one kernel: joins the SASS page of an ncu report (instructions executed,
stall samples per address) with nvdisasm -g line info of the SAME build of libqnb.so.
usage: python tools/sass_lines.py <report.ncu-rep> <kernel-regex> <mangled-substring> [top]"""
import csv, re, subprocess, sys, collections, os, tempfile

rep, kern, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'q6_b200', 'csrc', 'libqnb.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
line_of = {}   # offset -> (file, line)
inside, cur = False, None
for l in dis:
    if l.startswith('.text.'):
        inside = mangled in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
inst, samp, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
base, seen = None, 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        seen += 1
        if seen > 1: break
        continue
    if len(r) < 6 or r[0] == 'Address': continue
    a = int(r[0], 16)
    if base is None: base = a
    src, _ = line_of.get(a - base, (None, None))
    inst[src] += int(r[5]); samp[src] += int(r[2])
    op = [t for t in r[1].split() if not t.startswith('@')][0].split('.')[0]
    ops[src][op] += int(r[5])
ti, ts = sum(inst.values()), sum(samp.values())
print(f'total warp instructions {ti}, samples {ts}')
srcs = {}
for (k, v) in inst.most_common(top):
    text = ''
    if k:
        path = os.path.join(os.path.dirname(lib), k[0])
        if path not in srcs and os.path.exists(path): srcs[path] = open(path).read().splitlines()
        if path in srcs and k[1] - 1 < len(srcs[path]): text = srcs[path][k[1] - 1].strip()[:90]
    o = ' '.join(f'{a}:{b}' for a, b in ops[k].most_common(4))
    print(f'{str(k):28s} {v:9d} {100*v/ti:5.1f}%  smp {100*samp[k]/max(ts,1):5.1f}%  | {text}\n{"":34s}[{o}]')
