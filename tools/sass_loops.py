"""development aid: per-loop instruction mix of a kernel in the built library"""
import re, subprocess, sys
from collections import Counter
so = sys.argv[1]; fun = sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', '-fun', fun, so], capture_output=True, text=True).stdout
ins = []
for l in txt.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
op = lambda t: re.sub(r'^@!?U?P\w+\s+', '', t).split()[0].split('.')[0]
print('total', len(ins), dict(Counter(op(t) for _, t in ins).most_common(10)))
for a, t in ins:
    if 'BRA' in t:
        mm = re.search(r'0x([0-9a-f]+)', t)
        if mm and int(mm.group(1), 16) < a:
            tgt = int(mm.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= a]
            c = Counter(op(b[1]) for b in body)
            fp = sum(c[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
            print(hex(tgt), '->', hex(a), 'len', len(body), 'fp64', fp, dict(c.most_common(12)))
