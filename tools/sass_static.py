"""Static SASS opcode counts per kernel of libqnb.so (quick check before spending GPU time).
usage: python tools/sass_static.py [kernel-substring ...]"""
import subprocess, sys, re, collections, os
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'q6_b200', 'csrc', 'libqnb.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m: cur = m.group(1); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m and cur:
        op = m.group(1)
        key = 'MOVs' if op.startswith('MOV') or op.startswith('IMAD.MOV') else op.split('.')[0]
        cnt[cur][key] += 1
for k, c in cnt.items():
    if len(sys.argv) > 1 and not any(a in k for a in sys.argv[1:]): continue
    print(k[:70], 'total', sum(c.values()), ' '.join(f'{a}:{b}' for a, b in c.most_common(10)))
