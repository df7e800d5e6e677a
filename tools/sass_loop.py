"""Dev helper: instruction histogram of the innermost SASS loop that contains given mnemonics.
usage: python tools_sass_loop.py <obj-or-so> <mangled-kernel-substring> [MNEMONIC ...]"""
import re, subprocess, sys
from collections import Counter
obj, fn = sys.argv[1], sys.argv[2]
need = sys.argv[3:] or ["SHFL.UP", "DADD"]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks:
    name = b.split("\n", 1)[0]
    if fn not in name:
        continue
    ops = []
    for l in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ops.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for a, o in ops:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s+)?0x([0-9a-f]+)", o)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in ops if tgt <= x[0] <= a]
            if all(any(n in x[1] for x in body) for n in need):
                if best is None or len(body) < len(best):
                    best = body
    if best is None:
        print(name, "no loop found"); continue
    c = Counter(re.sub(r"^@!?U?P\d\s+", "", x[1]).split()[0].split(".")[0] for x in best)
    print(name.strip(), "loop instrs:", len(best))
    print("  ", c.most_common())
