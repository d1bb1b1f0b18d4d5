"""Per-SASS view of an `ncu --page source --csv --print-source cuda,sass` dump: instructions of the source lines in
[lo, hi] of a file, with samples and top stall reasons.  usage: ncu_sass.py dump.csv file lo hi [min_samples]"""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
fn, lo, hi = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
mins = int(sys.argv[5]) if len(sys.argv) > 5 else 20
f = None; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = os.path.basename(r[1]); continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] in ('Function Name', 'Kernel Name') or hdr is None: continue
    if r[0] != '':
        cur = (f, int(r[0]), r[1].strip()[:70]); continue
    if cur is None or cur[0] != fn or not (lo <= cur[1] <= hi): continue
    d = dict(zip(hdr[4:], r[4:]))
    smp = int(d['# Samples']) if d['# Samples'].isdigit() else 0
    if smp < mins: continue
    stalls = sorted(((k[6:], int(v)) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '0', '-')), key=lambda kv: -kv[1])[:3]
    print(f"{cur[1]:>4} smp {smp:>5} inst {d['Instructions Executed']:>8} {r[3].strip()[:70]:<70} {stalls}")
