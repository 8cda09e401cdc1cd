"""Loops of a SASS listing (cuobjdump -sass -fun <kernel> file.o): size, POPC / local-memory traffic per loop body.
The hot loop of the matching kernel must hold no LDL / STL (spills)."""
import re, sys
from collections import Counter
lines = open(sys.argv[1]).read().splitlines()
addr = re.compile(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);')
ins = []
for l in lines:
    m = addr.search(l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
loops = []
for a, t in ins:
    m = re.search(r'BRA.*0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        loops.append((len(body), tgt, a, body))
for n, tgt, a, body in sorted(loops, key=lambda x: x[0]):
    popc = sum('POPC' in x[1] for x in body)
    if popc == 0:
        continue
    ldl, stl = sum('LDL' in x[1] for x in body), sum('STL' in x[1] for x in body)
    print(f"loop {tgt:#x}-{a:#x}: {n} instructions, POPC {popc}, LDL {ldl}, STL {stl}")
    if popc >= 16 and n < 400:
        c = Counter((x[1].split()[1] if x[1].startswith('@') else x[1].split()[0]).split('.')[0] for x in body)
        print("   mix:", dict(c.most_common(12)))
