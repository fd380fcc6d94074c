"""Static SASS instruction counts per source line of one kernel of libqnb.so (nvdisasm -g line info); a loop body's lines
give its instructions per iteration when the compiler did not unroll it.
usage: python tools/sass_static_lines.py <mangled-substring> [first-line last-line]"""
import collections, os, re, subprocess, sys, tempfile
mangled = sys.argv[1]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10 ** 9)
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'q6_b200', 'csrc', 'libqnb.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
inside, cur = False, None
cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for l in dis:
    if l.startswith('.text.'):
        inside = mangled in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m and cur:
        cnt[cur] += 1; ops[cur][m.group(2).split('.')[0]] += 1
tot = 0
for k in sorted(cnt):
    if k[0].startswith('qnb_lists') and not (lo <= k[1] <= hi): continue
    tot += cnt[k]
    print(f"{k[0]}:{k[1]:5d} {cnt[k]:4d}  " + ' '.join(f"{o}:{c}" for o, c in ops[k].most_common(6)))
print("total", tot, " all lines", sum(cnt.values()))
