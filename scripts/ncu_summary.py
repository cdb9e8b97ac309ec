#!/usr/bin/env python
"""Condense ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python scripts/ncu_summary.py launches <launches.csv> <out.md>     # per-kernel time shares
  python scripts/ncu_summary.py full <prof.ncu-rep> <out.md> [key]   # raw metrics of each captured launch
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_fma.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def launches(path, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            rows.append((int(r["ID"]), short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], v * scale))
    tot = collections.OrderedDict()
    for _, n, _, _, ms in rows:
        a = tot.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(ms for *_, ms in rows)
    with open(out, "w") as f:
        f.write("# ncu launch list: `%s`\n\n" % os.path.basename(path))
        f.write("gpu__time_duration.sum per launch, --clock-control none (cold-cache, serialised: compare shares).\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for n, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (n, c, ms, 100 * ms / total))
        f.write("\n## every launch\n\n| id | kernel | grid | block | ms |\n|---:|---|---|---|---:|\n")
        for i, n, g, b, ms in rows:
            f.write("| %d | `%s` | %s | %s | %.4f |\n" % (i, n, g, b, ms))
    print("wrote", out)


def full(path, out, key=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(out, "w") as f:
        f.write("# ncu --set full capture: `%s`\n\n" % os.path.basename(path))
        for n, r in enumerate(rows[2:]):
            f.write("## launch %d: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % (n, short(r[hdr.index("Kernel Name")])))
            vals = {}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    vals[m] = (r[i], units[i])
                    f.write("| %s | %s | %s |\n" % (m, r[i], units[i]))
            f.write("\n")
            try:
                def to_bytes(v, u):
                    s = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
                    return float(v.replace(",", "")) * s
                t = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
                if t == t:
                    traffic[n] = t
                    f.write("DRAM traffic (read+write) = %.3f GB\n\n" % (t / 1e9))
            except Exception:
                pass
    print("wrote", out)
    if key and traffic:
        p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[key] = traffic[min(traffic)]
        d[key + "__source"] = os.path.basename(out)
        json.dump(d, open(p, "w"), indent=1, sort_keys=True)
        print("updated", p)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
