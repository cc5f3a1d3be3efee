"""CPU-side performance guard: static checks on the SASS of the headline kernel k_engine<4,2,0,0> in the built library
(no GPU needed).  The steady-state sweeps of a trajectory — the two innermost loops that contain the Philox rounds — were
tuned to 382 (no resampling) and 443-467 (gathering) instructions without a single local-memory access (register spills were
the first thing that cost performance in this kernel).  A change that bloats them or makes them spill
fails here, before any GPU time is spent."""
import collections
import re
import subprocess

import pytest


def _loops(so, pat):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name = part.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in part.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        addr_ix = {a: i for i, (a, _) in enumerate(ins)}
        out = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and tgt in addr_ix:
                    out.append([t for _, t in ins[addr_ix[tgt]:i + 1]])
        return len(ins), out
    return 0, []


def test_steady_state_sweeps_stay_tight_and_spill_free(built):
    total, loops = _loops(built, "k_engineILi4ELi2ELi0ELi0")
    assert total > 0, "k_engine<4,2,0,0> not found in the library"
    philox = [b for b in loops if any("-0x2daee0ad" in t.lower() for t in b)]
    # the sweeps carry the bulk of the FP64 work (the small Philox loops belong to the stratified-threshold search)
    steady = sorted((b for b in philox if 300 <= len(b) < 600 and sum("DFMA" in t for t in b) > 80), key=len)
    assert len(steady) >= 2, [len(b) for b in philox]
    for body, cap in zip(steady[:2], (400, 470)):     # r2: the gathering sweep carries the packed-entry branch (467)
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split(".")[0].split()[0] for t in body)
        assert len(body) <= cap, (len(body), cap)
        assert ops["LDL"] == 0 and ops["STL"] == 0, dict(ops)      # no spills inside the sweep
        assert ops["DFMA"] + ops["DMUL"] + ops["DADD"] <= 160       # FP64 work per particle
        assert ops["CALL"] == 0                                     # everything inlined (no libdevice slow paths)
    assert total < 30000                                            # whole kernel: instruction-cache footprint
