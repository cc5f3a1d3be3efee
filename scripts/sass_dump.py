"""Print instructions [lo,hi] (indices as reported by sass_loops.py) of one kernel in a .so."""
import re, subprocess, sys
so, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for l in part.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    for a, t in ins[lo:hi + 1]:
        print(f"{a:06x}  {t}")
