"""Static census of the loops of one kernel: instructions per backward-branch span, with an opcode histogram.
    cuobjdump -sass file.o | python scripts/sass_loops.py '<substring of the mangled kernel name>'"""
import re, sys, collections
pat = sys.argv[1]
txt = sys.stdin.read()
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+([^;]+);", f):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, len(ins), "instructions")
    for k, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= a]
            if len(body) < 40:
                continue
            hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
            print("  loop %#x..%#x: %d instructions; " % (tgt, a, len(body)) + ", ".join("%s %d" % kv for kv in hist.most_common(14)))
