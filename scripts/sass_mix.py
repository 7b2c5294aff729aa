"""Instruction mix per kernel from `cuobjdump -sass` output (development aid)."""
import collections
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv) > 2 else ""
parts = re.split(r'\n\s*Function : ', txt)
for p in parts[1:]:
    name = p.split('\n', 1)[0].strip()
    dn = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    if pat and not re.search(pat, dn):
        continue
    c = collections.Counter()
    n = 0
    for line in p.split('\n'):
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            n += 1
            c[m.group(1).split('.')[0]] += 1
    keys = ['DFMA', 'DMUL', 'DADD', 'LDL', 'STL', 'LDS', 'STS', 'LDG', 'STG', 'BRA', 'MUFU', 'LDC', 'MOV', 'IMAD']
    print(dn[:90], 'total', n, {k: c[k] for k in keys if c[k]})
