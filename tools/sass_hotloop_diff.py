"""Compare the inner loop (the loop with 32 POPC) of two SASS listings instruction by instruction (registers included)."""
import re, sys
def hot(path):
    addr = re.compile(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);')
    ins = [(int(m.group(1), 16), m.group(2)) for m in (addr.search(l) for l in open(path)) if m]
    best = None
    for a, t in ins:
        m = re.search(r'BRA.*0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            body = [x[1] for x in ins if int(m.group(1), 16) <= x[0] <= a]
            if sum('POPC' in x for x in body) == 32 and (best is None or len(body) < len(best)):
                best = body
    return best
a, b = hot(sys.argv[1]), hot(sys.argv[2])
same_op = sum(1 for x, y in zip(a, b) if x.split()[0] == y.split()[0])
same_all = sum(1 for x, y in zip(a, b) if re.sub(r'0x[0-9a-f]+', '', x) == re.sub(r'0x[0-9a-f]+', '', y))
print(f"{len(a)} vs {len(b)} instructions; same opcode at {same_op} positions; identical (registers too) at {same_all}")
