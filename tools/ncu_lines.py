"""Per-source-line instruction counts of an .ncu-rep captured with --import-source on (top 40 lines): where a kernel's
warp-instructions go.  usage: python tools/ncu_lines.py prof.ncu-rep"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, out, ia, iss = None, [], None, None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        print("==", r[1]); continue
    if r[0] == "Line No":
        ia, iss = r.index("Instructions Executed"), r.index("# Samples"); continue
    if len(r) > 3 and r[2] == "-" and ia is not None:
        try:
            out.append((int(r[ia]), int(r[iss]), cur, r[0], r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
print(f"warp-instructions attributed to source lines: {tot}")
print("  share   stall-samples  file:line  source")
for n, s, fn, ln, src in sorted(out, reverse=True)[:40]:
    print(f"{100.0 * n / tot:6.2f}%  {s:8d}  {fn}:{ln}  {src}")
