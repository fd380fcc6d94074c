"""Opcode histogram (instructions executed, stall samples) of one kernel from an ncu report's source page.
usage: python tools/sass_hist.py <report.ncu-rep> <kernel-regex>"""
import csv, subprocess, sys, collections

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
first = True
inst, samp = collections.Counter(), collections.Counter()
for r in rows:
    if r and r[0] == 'Kernel Name':
        if not first: break   # first captured launch only
        first = False; print(r[1][:100]); continue
    if len(r) < 6 or r[0] == 'Address': continue
    op = r[1].split()
    op = [t for t in op if not t.startswith('@')][0].rstrip(';')
    key = '.'.join(op.split('.')[:2]) if op.split('.')[0] in ('F2F', 'I2F', 'F2I', 'MUFU', 'LDG', 'LDS', 'STS', 'RED', 'ATOM', 'ATOMG') else op.split('.')[0]
    inst[key] += int(r[5]); samp[key] += int(r[2])
ti, ts = sum(inst.values()), sum(samp.values())
print(f'total warp instructions {ti}, stall samples {ts}')
for k, v in inst.most_common(32):
    print(f'{k:14s} {v:10d} {100*v/ti:6.2f}%   samples {100*samp[k]/max(ts,1):6.2f}%')
