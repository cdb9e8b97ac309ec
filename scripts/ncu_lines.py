#!/usr/bin/env python
"""Per-source-line totals of a captured launch (ncu source page, kernels built with -lineinfo):
executed warp instructions, stall samples and shared-memory wavefronts per file:line.

    python scripts/ncu_lines.py <prof.ncu-rep> <launch index> [<rows>] [<top n>]
"""
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0

rep, which = sys.argv[1], int(sys.argv[2])
nrows = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# sections: "File Path", "Function Name", header, lines...
secs = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        path, fn, hdr = rows[i][1], rows[i + 1][1], rows[i + 2]
        j = i + 3
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j])
            j += 1
        secs.append((path, fn, hdr, body))
        i = j
    else:
        i += 1
fns = []
for s in secs:
    if s[1] not in fns:
        fns.append(s[1])
fn = fns[which]
print("# %s\n" % fn[:120])
out = []
for path, f, hdr, body in secs:
    if f != fn:
        continue
    il, isrc = hdr.index("Line No"), hdr.index("Source")
    ii, isamp, iw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
    stall = [c for c, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    for r in body:
        if not r or not r[il]:
            continue
        st = sorted(((num(r[c]), hdr[c][6:]) for c in stall), reverse=True)[:2]
        out.append((num(r[ii]), num(r[isamp]), num(r[iw]), path.split("/")[-1], r[il],
                    r[isrc].strip()[:70], ",".join("%s %.0f" % (n, v) for v, n in st if v > 0)))
ti, ts, tw = (sum(o[k] for o in out) for k in range(3))
print("total: %.0f instr/row, %.0f samples, %.0f smem wavefronts/row\n" % (ti / nrows, ts, tw / nrows))
byfile = {}
for o in out:
    a = byfile.setdefault(o[3], [0, 0, 0])
    a[0] += o[0]; a[1] += o[1]; a[2] += o[2]
for f, a in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print("%-24s instr/row %7.0f  samples %5.1f%%  wavefronts/row %6.0f" % (f, a[0] / nrows, 100 * a[1] / ts, a[2] / nrows))
print()
key = int(sys.argv[5]) if len(sys.argv) > 5 else 1
for o in sorted(out, key=lambda x: -x[key])[:top]:
    print("%7.0f i/row %5.1f%% smp %6.0f wf/row  %s:%s  %s  [%s]" % (o[0] / nrows, 100 * o[1] / ts, o[2] / nrows, o[3], o[4], o[5], o[6]))
