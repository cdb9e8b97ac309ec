#!/usr/bin/env python
"""Aggregate ncu source-page samples of one kernel launch by code region / instruction.
   python scripts/ncu_regions.py <rep> [launch_index] [bucket]"""
import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; B = int(sys.argv[3]) if len(sys.argv) > 3 else 250
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hdr_idx[which]]
end = hdr_idx[which + 1] - 1 if len(hdr_idx) > which + 1 else len(rows)
body = rows[hdr_idx[which] + 1:end]
iS = h.index('# Samples'); iE = h.index('Instructions Executed'); iSrc = h.index('Source')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_')] 
tot = sum(int(r[iS]) for r in body); totE = sum(int(r[iE]) for r in body)
print("instructions:", len(body), "samples", tot, "executed", totE)
TAGS = ('FFMA2', 'F2FP', 'LDTM', 'UTCHMMA', 'SHFL', 'SYNCS', 'LDG', 'LDL', 'STL', 'MUFU', 'WARPSYNC', 'LDS', 'STS')
for b in range(0, len(body), B):
    chunk = body[b:b + B]
    s = sum(int(r[iS]) for r in chunk); e = sum(int(r[iE]) for r in chunk)
    tags = {t: sum(1 for r in chunk if t in r[iSrc]) for t in TAGS}
    print("%5d-%5d samples %5.1f%% exec %5.1f%% %s" % (b, b + B, 100 * s / tot, 100 * e / totE, ' '.join('%s:%d' % (k, v) for k, v in tags.items() if v)))
print("top instructions by samples:")
for r in sorted(body, key=lambda r: -int(r[iS]))[:30]:
    print("%8s %12s %5d  %s" % (r[iS], r[iE], body.index(r), r[iSrc].strip()[:100]))
