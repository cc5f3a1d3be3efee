"""Per CUDA-source-line sample/instruction shares from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for k, r in enumerate(rows[:6]):
    if "Line No" in r:
        hdr, start = r, k + 1
        break
iL, iS, iSamp, iInst = hdr.index("Line No"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
out = []
for r in rows[start:]:
    if r and r[iL].strip():
        try:
            off = len(r) - len(hdr)
            out.append((float(r[iSamp + off]), float(r[iInst + off]), r[iL], ",".join(r[iS:iS + off + 1]).strip()[:100]))
        except Exception:
            pass
tot, toti = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"samples {tot:.0f} warp-insts {toti:.3e}")
for s, i, l, t in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{100*s/tot:5.1f}%s {100*i/toti:5.1f}%i L{l:>5} {t}")
