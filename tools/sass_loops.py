"""List the loops (backward branches) of a SASS listing with their instruction counts:
   cuobjdump -sass -fun NAME file.o | python tools/sass_loops.py [--dump START_HEX]"""
import re, sys
lines = [l for l in sys.stdin if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l)]
ins = []
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
for a, t in ins:
    m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        b = int(m.group(1), 16)
        body = [x for _, x in ins[addr[b]:addr[a] + 1]]
        ops = {}
        for x in body:
            op = x.split()[1] if x.startswith("@") else x.split()[0]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        if len(body) <= int(sys.argv[1] if len(sys.argv) > 1 else 400): print("loop %04x..%04x  %d instructions  %s" % (b, a, len(body), " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1]))))
