"""Aggregate an .ncu-rep's warp-stall samples per CUDA source line (needs -lineinfo and --import-source on)."""
import collections, csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname, hdr, agg = None, None, collections.OrderedDict()
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r) if h not in ("Source",)}; src_i = 1; continue
    if hdr is None or r[2] != "-": continue          # per-line aggregate rows carry "-" in the address column
    try: s = float(r[hdr["# Samples"]]); n = float(r[hdr["Instructions Executed"]])
    except Exception: continue
    st = {k[6:]: float(r[i] or 0) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k}
    agg[(fname, int(r[0]))] = (s, n, r[src_i].strip()[:100], st)
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print(f"samples {tot:.0f}, warp instructions {toti:.3e}")
for (f, l), (s, n, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    main = ", ".join(f"{k} {100*v/max(s,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/tot:5.1f}%  inst {100*n/toti:5.1f}%  {f}:{l:<5d} {src[:70]:70s} [{main}]")
