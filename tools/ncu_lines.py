"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line: instructions executed and
stall samples.  usage: ncu_lines.py dump.csv [top]"""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []; f = None; hdr = None; tot_i = 0; tot_s = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = os.path.basename(r[1]); continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] in ('Function Name', 'Kernel Name'): continue
    if hdr is None or r[0] == '': continue
    d = dict(zip(hdr[2:], r[2:]))   # skip the duplicated "Source" header at index 1
    try:
        inst = int(d['Instructions Executed']); smp = int(d['# Samples'])
    except Exception:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '0')}
    out.append((inst, smp, f, r[0], r[1].strip()[:90], stalls)); tot_i += inst; tot_s += smp
print('total inst', tot_i, 'samples', tot_s)
print('--- by instructions')
for o in sorted(out, key=lambda o: -o[0])[:top]:
    print(f'{o[0]:>10} {100*o[0]/tot_i:5.1f}% smp {o[1]:>5} {o[2]}:{o[3]} {o[4]}')
print('--- by samples')
for o in sorted(out, key=lambda o: -o[1])[:top]:
    s = sorted(o[5].items(), key=lambda kv: -kv[1])[:3]
    print(f'{o[1]:>6} {100*o[1]/tot_s:5.1f}% inst {o[0]:>9} {o[2]}:{o[3]} {o[4]}  {s}')
