"""sha256 of the SASS instruction stream of one kernel in a built .so (addresses and encodings stripped).
    python scripts/sass_hash.py <lib.so> [name-substring]
Used to check that a change left a tuned kernel byte-identical, and to tie profiles/traffic.json to the kernel it was
measured on."""
import hashlib
import re
import subprocess
import sys


def sass_hashes(so, pats):
    """{pattern: (symbol, instruction count, sha16)} from ONE cuobjdump pass over the library."""
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out = {p: (None, 0, None) for p in pats}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name = part.split("\n", 1)[0]
        for pat in pats:
            if pat not in name or out[pat][0] is not None:
                continue
            ins = []
            for line in part.split("\n"):
                m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
                if m:
                    ins.append(m.group(2).strip())
            out[pat] = (name.strip(), len(ins), hashlib.sha256("\n".join(ins).encode()).hexdigest()[:16])
    return out


def sass_hash(so, pat="k_engineILi4ELi2ELi0ELi0"):
    return sass_hashes(so, [pat])[pat]


if __name__ == "__main__":
    print(*sass_hash(sys.argv[1], *(sys.argv[2:3])))
