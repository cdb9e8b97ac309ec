#!/usr/bin/env python
"""Per-role breakdown of a captured row_update_umma launch from the ncu source page:
instructions, sampled stall reasons and shared-memory wavefronts of the four warp roles
(the role boundaries are the USETMAXREG instructions in program order).

    python scripts/ncu_roles.py <prof.ncu-rep> <rows launch0> [<rows launch1> ...] > out.md
"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
rows_per_launch = [float(x) for x in sys.argv[2:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i > 0 and len(rows[i - 1]) > 1 else "?" for i in hidx]
# the page repeats every launch twice (SASS view per source view); keep distinct kernels in order
seen, launches = set(), []
for li, h0 in enumerate(hidx):
    end = hidx[li + 1] - 1 if li + 1 < len(hidx) else len(rows)
    key = (names[li], sum(float(r[rows[h0].index("Instructions Executed")] or 0) for r in rows[h0 + 1:end]))
    if key in seen:
        continue
    seen.add(key)
    launches.append((names[li], rows[h0], rows[h0 + 1:end]))
print("# Per-role breakdown: `%s`\n" % rep.split("/")[-1])
print("Roles in program order: producers + MMA issuer, drain warpgroup, Cholesky warps. Stall shares are of the"
      " role's own warp samples; wavefronts are LSU shared-memory wavefronts (tensor-core operand reads excluded).\n")
for n, (name, hdr, data) in enumerate(launches):
    nrows = rows_per_launch[n] if n < len(rows_per_launch) else None
    isrc, ins, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    iw = hdr.index("L1 Wavefronts Shared")
    stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    b = [i for i, r in enumerate(data) if "USETMAXREG" in r[isrc]]
    edges = [0] + b + [len(data)]
    # (the 4 + 11 mix keeps the producers at the launch register count: no USETMAXREG there)
    roles = (["prologue", "producers + MMA", "drain", "Cholesky"] if len(b) == 3
             else ["prologue + producers + MMA", "drain", "Cholesky"])
    tot = sum(float(r[ins] or 0) for r in data)
    tots = sum(float(r[isamp] or 0) for r in data)
    print("## launch %d: `%s`%s\n" % (n, name, " (%d rows)" % nrows if nrows else ""))
    print("| role | static instr | executed | share | per row | samples | top stall reasons | smem wavefronts per row |")
    print("|---|---:|---:|---:|---:|---:|---|---:|")
    for role, (s, e) in zip(roles, zip(edges[:-1], edges[1:])):
        seg = data[s:e]
        ii = sum(float(r[ins] or 0) for r in seg)
        ss = sum(float(r[isamp] or 0) for r in seg)
        ww = sum(float(r[iw] or 0) for r in seg)
        st = {hdr[c][6:]: sum(float(r[c] or 0) for r in seg) for c in stall}
        top = sorted(st.items(), key=lambda x: -x[1])[:4]
        print("| %s | %d | %.3e | %.1f%% | %s | %.1f%% | %s | %s |" % (
            role, e - s, ii, 100 * ii / tot, "%.0f" % (ii / nrows) if nrows else "-", 100 * ss / max(tots, 1),
            ", ".join("%s %.0f%%" % (k, 100 * v / max(ss, 1)) for k, v in top),
            "%.0f" % (ww / nrows) if nrows else "-"))
    print()
