"""Where ptxas spilled: LDL / STL per source line of one kernel.  usage: spills.py lib.so kernel-substring"""
import re, subprocess, sys, tempfile, os, glob
lib, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
sass = subprocess.run(['nvdisasm', '-gi'] + glob.glob(d + '/*.cubin'), capture_output=True, text=True).stdout
cur = None; infn = False; res = {}
for line in sass.splitlines():
    if line.startswith('.text.'): infn = pat in line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if infn and re.search(r'\b(STL|LDL)\b', line):
        k = cur; res.setdefault(k, [0, 0])[0 if 'STL' in line else 1] += 1
for k, v in sorted(res.items(), key=lambda kv: kv[0] or ('', 0)): print(k, 'STL', v[0], 'LDL', v[1])
