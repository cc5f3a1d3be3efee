"""Static look at the hot loop of k_engine<NX,NY,DYN> in the built .so: finds backward branches, picks the
loop that contains the Philox multiplier, prints its opcode histogram (LDL/STL = spills inside the loop)."""
import collections, re, subprocess, sys
so = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "k_engineILi4ELi2ELi0"
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
    addr_ix = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_ix:
                loops.append((addr_ix[tgt], i))
    print(name[:50], "instrs", len(ins), "loops", len(loops))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        if not any("-0x2daee0ad" in t.lower() for _, t in body):
            continue
        c = collections.Counter()
        for _, t in body:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", t)
            c[m.group(2)] += 1
        print(f"  loop [{lo},{hi}] size {hi - lo + 1}:", dict(c.most_common(24)))
