#!/usr/bin/env python
"""bench.py -- ALS iterations/s on the BASELINE.json headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]

One "step" = one ALS iteration (Gramian + X<-Y, Gramian + Y<-X; AlternatingLeastSquares.java
:227-229) over the synthetic workload of SURVEY.md 8(d), generated on the device.
Prints ONE JSON line (rank 0).  `value` is iterations/s with everything resident in HBM;
`e2e` is the same metric through the C ABI with HOST buffers (CSR + Y0 uploaded from pinned
memory, K iterations, X and Y read back) -- the number to compare with the reference arm.
`--impl reference` times the CPU oracle (the C port of the reference's Java ALS; there is no
JVM in this image) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 1234567890  # RandomManager test seed (common/.../random/RandomManager.java:52)

# BASELINE.json configs: name -> users, items, nnz/user, k
CONFIGS = {
    "c1": dict(users=10_000, items=2_000, nnz_per_user=20, k=16),
    "c2": dict(users=1_000_000, items=100_000, nnz_per_user=50, k=32),
    "c3": dict(users=10_000_000, items=1_000_000, nnz_per_user=100, k=64),
    # c3 at 1/5 scale (same entries per user and per item): profiling-only, ncu replays need
    # the working set small enough to save/restore between passes
    "c3p": dict(users=2_000_000, items=200_000, nnz_per_user=100, k=64),
    # power-law data in the shape of config 5 (50M x 5M, ~200/user, power-law, k=128, 8 GPUs) at 1/25 scale
    # on one GPU and at k=64 (the tensor-core path; k=128 runs on the CUDA-core kernel): rows from 20 to
    # 20 000 entries, a few items that nearly every user has.  Device-resident line only (no CPU arm).
    "c5p": dict(users=2_000_000, items=200_000, nnz_per_user=200, k=64, powerlaw=True, max_nnz=20_000),
    # BASELINE.json configs[4] at full size: needs --gpus 8 (10^10 entries; k = 128 runs on the CUDA-core kernel)
    "c5": dict(users=50_000_000, items=5_000_000, nnz_per_user=200, k=128, powerlaw=True, max_nnz=20_000),
}
WORKLOAD_NAMES = {
    "c1": "10k x 2k, 20 nnz/user, k=16",
    "c2": "1M x 100k, 50 nnz/user, k=32",
    "c3": "10M x 1M, 100 nnz/user, k=64 (headline)",
    "c3p": "2M x 200k, 100 nnz/user, k=64 (1/5-scale headline, profiling only)",
    "c5p": "2M x 200k, ~200 nnz/user power-law (max 20000), Zipf items, k=64 (config 5 at 1/25 scale)",
    "c5": "50M x 5M, ~200 nnz/user power-law (max 20000), Zipf items, k=128 (config 5, 8 GPUs)",
}


def algorithmic_bytes(U, I, nnz, k):
    """B_iter = 2*NNZ*(4k+8) + 2*(U+I)*4k + 2*(U+I+2)*8."""
    return 2 * nnz * (4 * k + 8) + 2 * (U + I) * 4 * k + 2 * (U + I + 2) * 8


def update_launch_bytes(rows, nnz, k):
    """Algorithmic bytes of ONE row-update launch: per entry one 4k-byte factor row + 4 B index
    + 4 B value; per row one 4k-byte write and one 8-byte row pointer; the k x k fp64 Gramian."""
    return nnz * (4 * k + 8) + rows * 4 * k + (rows + 1) * 8 + k * k * 8


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None, t_load=None):
        """Samples received inside the timed region [t0, t1]; when the region is shorter than the
        sampling period, the samples of the warm-up + timed region [t_load, t1] (the same kernels,
        back to back), and `window` says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        window = "timed region"
        rows = [r for t, r in self.rows if t0 is None or t0 <= t <= t1 + 0.02]
        if not rows and t_load is not None:
            rows = [r for t, r in self.rows if t_load + 0.05 <= t <= t1 + 0.05]
            window = "warm-up + timed region (the timed region is shorter than the sampling period)"
        sm, mx, power, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                                  ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons),
                "window": window}


# --------------------------------------------------------------------------------------
# CPU baseline: the oracle (C port of the reference's Java ALS) on a bounded sample.
def cpu_baseline_from_sample(cfg, user_rows, item_rows, Y, X, budget_note):
    """user_rows/item_rows: (ptr, idx, val) CSR slices; Y, X: full opposite factors (host).
    Times one X-half over the user sample and one Y-half over the item sample with all host
    threads (100-row work units, AlternatingLeastSquares.java:77), plus the single-threaded
    transposeTimesSelf (MatrixUtils.java:219-239) on a row sample, and scales each linearly
    to the full workload."""
    from oracle import oracle as O
    U, I, k = cfg["users"], cfg["items"], cfg["k"]
    cores = os.cpu_count() or 1
    nu, ni = user_rows[0].size - 1, item_rows[0].size - 1
    g_rows = min(I, 200_000)
    t0 = time.perf_counter()
    G = O.transpose_times_self(Y[:g_rows])
    t_g = (time.perf_counter() - t0) / g_rows
    out = np.zeros((nu, k), np.float32)
    t0 = time.perf_counter()
    O.als_half(user_rows[0], user_rows[1], user_rows[2], Y, G, out, n_threads=cores)
    t_x = (time.perf_counter() - t0) / max(nu, 1)
    G = O.transpose_times_self(X[:g_rows])
    out = np.zeros((ni, k), np.float32)
    t0 = time.perf_counter()
    O.als_half(item_rows[0], item_rows[1], item_rows[2], X, G, out, n_threads=cores)
    t_y = (time.perf_counter() - t0) / max(ni, 1)
    t_iter = t_x * U + t_y * I + t_g * (U + I)
    return {
        "value": 1.0 / t_iter, "unit": "iterations/s", "cores": cores, "kind": "port",
        "sample": ("C port of the reference Java ALS (oracle/als_oracle.c; no JVM on the box): "
                   "X-half on %d of %d users + Y-half on %d of %d items with %d threads, "
                   "single-threaded Gramian on %d rows, each scaled linearly to the full "
                   "workload; %s" % (nu, U, ni, I, cores, g_rows, budget_note)),
        "seconds_per_iteration_extrapolated": t_iter,
        "split_s": {"x_half": t_x * U, "y_half": t_y * I, "gramians": t_g * (U + I)},
    }


def sample_sizes(cfg, scale=1.0):
    # ~10-30 s of CPU work: cost/row ~ nnz*k^2 (fp64 rank-1) + ~2.7*k^3 (Householder QR)
    k, nnz = cfg["k"], cfg["nnz_per_user"]
    per_user = nnz * k * k + 2.7 * k ** 3
    per_item = nnz * cfg["users"] / cfg["items"] * k * k + 2.7 * k ** 3
    cores = os.cpu_count() or 1
    budget = 3.0e9 * cores * scale  # a few seconds per half at ~1 GFMA/s/core
    nu = int(min(cfg["users"], max(1000, budget / per_user)))
    ni = int(min(cfg["items"], max(200, budget / per_item)))
    return nu, ni


def reference_arm(args, cfg, name):
    """--impl reference: no GPU code on this path. Workload rows from the numpy twin of the
    device generator (oracle/synth.py)."""
    from oracle import synth
    if cfg.get("powerlaw"):
        # (the item-row sampler below walks the uniform generator's strata; the power-law configs are bench
        # lines of this repo only -- their CPU check is the sampled-row parity inside the b200 arm)
        print(json.dumps({"impl": "reference", "unavailable": "no CPU arm for the power-law configs (c5p, c5)"}))
        return
    U, I, nnz, k = cfg["users"], cfg["items"], cfg["nnz_per_user"], cfg["k"]
    nu, ni = sample_sizes(cfg)
    user_rows = synth.synth_rows(0, nu, I, nnz, seed=SEED, neg_fraction=0.0)
    Y = synth.unit_rows(I, k, seed=SEED)
    # Item sample: the first `ni` items. Stratum j of the generator covers items
    # [j*I/nnz, (j+1)*I/nnz); scan every user's draw in the strata that intersect the sample.
    item_ptr, item_idx, item_val = synth_item_rows(cfg, ni)
    # X for the Y-half: timing is data-independent; tile a block of unit rows to full height.
    blk = synth.unit_rows(min(U, 1_000_000), k, seed=SEED + 1)
    X = np.tile(blk, ((U + blk.shape[0] - 1) // blk.shape[0], 1))[:U]
    results, spent = [], []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = cpu_baseline_from_sample(cfg, user_rows, (item_ptr, item_idx, item_val), Y, X,
                                     "per bench step one such sample")
        if step >= args.warmup:
            results.append(r)
            spent.append(time.perf_counter() - t0)
    t_iter = float(np.mean([r["seconds_per_iteration_extrapolated"] for r in results]))
    cb = results[-1]
    cb["value"] = 1.0 / t_iter
    line = {
        "impl": "reference", "metric": "ALS iterations/sec", "value": 1.0 / t_iter,
        "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # one step = one bounded SAMPLE of the workload (base contract, tier section 4): the time
        # actually spent per step; `value` is the sample scaled linearly to the full workload
        "ms_per_step": float(np.mean(spent)) * 1e3,
        "ms_per_full_iteration_extrapolated": t_iter * 1e3,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[name], "users": U, "items": I,
                   "nnz_per_user": nnz, "features": k, "alpha": 1.0, "lambda": 0.1,
                   "seed": SEED, "host_threads": os.cpu_count() or 1,
                   "sample_per_step": "%d users + %d items of the workload" % (nu, ni)},
        "updated_rows_per_s": (U + I) / t_iter,
        "cpu_baseline": cb,
        "e2e": {"value": 1.0 / t_iter, "unit": "iterations/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def synth_item_rows(cfg, ni):
    from oracle import synth
    U, I, nnz = cfg["users"], cfg["items"], cfg["nnz_per_user"]
    strata = sorted({min(nnz - 1, (i * nnz) // I) for i in (0, ni - 1)})
    strata = list(range(strata[0], strata[-1] + 1))
    rows_l, items_l, vals_l = [], [], []
    chunk = 2_000_000
    for j in strata:
        lo, hi = (j * I) // nnz, ((j + 1) * I) // nnz
        for u0 in range(0, U, chunk):
            us = np.arange(u0, min(U, u0 + chunk), dtype=np.uint64)
            h = synth.synth_hash(SEED, us, np.full(us.size, j, dtype=np.uint64))
            it = lo + ((h >> np.uint64(32)) % np.uint64(hi - lo)).astype(np.int64)
            keep = it < ni
            s = (1 + ((h & np.uint64(0xffff)) % np.uint64(5)).astype(np.int64)).astype(np.float32)
            rows_l.append(us[keep].astype(np.int32)); items_l.append(it[keep]); vals_l.append(s[keep])
    rows = np.concatenate(rows_l); items = np.concatenate(items_l); vals = np.concatenate(vals_l)
    order = np.lexsort((rows, items))
    ptr = np.zeros(ni + 1, dtype=np.int64)
    np.cumsum(np.bincount(items, minlength=ni), out=ptr[1:])
    return ptr, np.ascontiguousarray(rows[order]), np.ascontiguousarray(vals[order])


# --------------------------------------------------------------------------------------
def sampled_parity(als, cfg, rank, world, n_user_rows=200, n_item_rows=50):
    """Size-independent parity check at the bench workload's full size, through the C ABI, after
    the timed region: one more X-half and Y-half; sampled rows of this rank's user / item block
    are re-solved by the CPU oracle (oracle/als_oracle.c) from the device's own opposite factor.
    User rows: Gramian from the oracle's transposeTimesSelf over the full Y.  Item rows: the
    device Gramian of X (als_gramian; itself checked against the oracle to 1e-12 in tests/) and
    only the X rows the sampled items reference (als_get_rows) -- the full X is 2.56 GB per rank.
    Returns relative errors (Frobenius, max) per half; bar 1e-4 (BASELINE.json north_star)."""
    from oracle import oracle as O
    from myrrix_recommender_b200.sharding import local_block
    U, I, k = cfg["users"], cfg["items"], cfg["k"]
    rng = np.random.default_rng(1000 + rank)

    def rel(a, b):
        a = a.astype(np.float64); b = b.astype(np.float64)
        return (float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)),
                float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)))

    def rows_csr(rows_local, by_column):
        ptrs, idxs, vals = [0], [], []
        for r in rows_local:
            _, i, v = als.get_interaction_rows(int(r), 1, by_column=by_column, capacity=max(1 << 20, U + 1))
            idxs.append(i); vals.append(v); ptrs.append(ptrs[-1] + i.size)
        return np.array(ptrs, np.int64), np.concatenate(idxs), np.concatenate(vals)

    als.half_x()
    als.sync()
    ub, ue = local_block(U, rank, world)
    loc = np.sort(rng.choice(ue - ub, min(n_user_rows, ue - ub), replace=False))
    up, ui, uv = rows_csr(loc, False)
    Y = als.get_y()
    Xs = als.get_rows(0, loc + ub)
    out = np.zeros((loc.size, k), np.float32)
    O.als_half(up, ui, uv, Y, O.transpose_times_self(Y), out, n_threads=os.cpu_count() or 1)
    ex = rel(Xs, out)
    GX = als.gramian(0)          # collective when sharded; X as the Y-half will read it
    als.half_y()
    als.sync()
    ib, ie = local_block(I, rank, world)
    loc = np.sort(rng.choice(ie - ib, min(n_item_rows, ie - ib), replace=False))
    if cfg.get("powerlaw"):
        # plus three popular items of this block (Zipf ranks >= 40 / 400 / 2500: rows of ~10^6 .. 10^4
        # entries, which the tensor-core kernel walks in chunks) -- a uniform sample rarely meets one
        from oracle import synth
        _, mul, add = synth.powerlaw_params(I, cfg["nnz_per_user"], cfg["max_nnz"], SEED)
        extra = []
        for t in (40, 400, 2500):
            for r in range(t, min(t + 64 * world, I)):
                item = (mul * r + add) % I
                if ib <= item < ie:
                    extra.append(item - ib)
                    break
        loc = np.unique(np.concatenate([loc, np.array(extra, dtype=loc.dtype)]))
    ip, ii, iv = rows_csr(loc, True)
    users, inv = np.unique(ii, return_inverse=True)
    Xc = als.get_rows(0, users)
    Ys = als.get_rows(1, loc + ib)
    out = np.zeros((loc.size, k), np.float32)
    O.als_half(ip, inv.astype(np.int32), iv, Xc, GX, out, n_threads=os.cpu_count() or 1)
    ey = rel(Ys, out)
    return {"x_half": {"rows": int(up.size - 1), "fro": ex[0], "max": ex[1]},
            "y_half": {"rows": int(ip.size - 1), "fro": ey[0], "max": ey[1]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    name = args.config
    cfg = CONFIGS[name]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, cfg, name)
        return 0

    import torch
    import myrrix_recommender_b200 as M

    def dbg(msg):
        if os.environ.get("BENCH_DEBUG"):
            print("[rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    U, I, nnz_pu, k = cfg["users"], cfg["items"], cfg["nnz_per_user"], cfg["k"]
    nnz = U * nnz_pu
    kernel = {"auto": 0, "simt": 1, "tcgen05": 2}[args.kernel]

    als = M.NativeALS(k, device=local_rank, kernel=kernel)
    # a real (non-default) stream: the library launches on it and the CUDA events below are
    # recorded on it, so the events bracket exactly the kernels that are timed
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    als.set_stream(stream.cuda_stream)
    if world > 1:
        import torch.distributed as dist
        uid = [M.factorizer.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        als.comm_init(rank, world, uid[0])
    dbg("comm ready, synthesising")
    if cfg.get("powerlaw"):
        args.no_e2e = args.no_cpu_baseline = True
        als.synth_interactions_powerlaw(U, I, nnz_pu, max_nnz=cfg["max_nnz"], seed=SEED, neg_fraction=0.0)
        nnz = int(als.info().nnz)  # this rank's user block
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([nnz], device="cuda", dtype=torch.int64)
            dist.all_reduce(t)
            nnz = int(t.item())
    else:
        als.synth_interactions(U, I, nnz_pu, seed=SEED, neg_fraction=0.0)
    als.synth_y0(seed=SEED)
    dbg("workload resident")
    h_y0 = None
    if not args.no_e2e:
        import ctypes as C
        h_y0 = torch.empty((I, k), dtype=torch.float32, pin_memory=True)
        als.check(als.lib.als_get_y(als.h, C.cast(h_y0.data_ptr(), C.POINTER(C.c_float))))
    info = als.info()
    kernel_name = {1: "simt", 2: "tcgen05"}[info.kernel]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up + exactly K timed iterations -----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # (nvidia-smi needs a few hundred ms before its first sample)
        time.sleep(0.5)
    t_load = time.perf_counter()
    als.iterate(args.warmup)
    als.sync()
    dbg("warm-up done")
    als.profile(True)
    als.timings(reset=True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    ev0.record(stream)
    als.iterate(args.steps)
    ev1.record(stream)
    barrier()
    t_end = time.perf_counter()
    dbg("timed region done")
    als.sync()
    clocks = sampler.stop(t_begin, t_end, t_load) if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    tm = als.timings(reset=True)
    als.profile(False)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # ---- roofline of the dominant kernel (row update), timed live with CUDA events --------
    peak, peak_src = measured_peak()
    x_ms = tm.update_x_ms / max(tm.n_half_x, 1)
    y_ms = tm.update_y_ms / max(tm.n_half_y, 1)
    bx = update_launch_bytes(U // world, nnz // world, k)
    by = update_launch_bytes(I // world, nnz // world, k)
    dom = "x" if x_ms >= y_ms else "y"
    dom_ms, dom_b = (x_ms, bx) if dom == "x" else (y_ms, by)
    achieved = dom_b / (dom_ms * 1e-3) / 1e9
    traffic = None  # measured at N = 1 only (one ncu pass over the full-size launches)
    try:
        if world > 1:
            raise LookupError
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("%s_%s_%s" % (name, kernel_name, dom))
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic,
        "kernel": "row_update_%s (%s half)" % (kernel_name, "X<-Y" if dom == "x" else "Y<-X"),
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_b,
        "avg_launch_ms": dom_ms,
        "other_half": {"avg_launch_ms": y_ms if dom == "x" else x_ms,
                       "algorithmic_bytes_per_launch": by if dom == "x" else bx},
        "gramian_ms_per_iteration": tm.gramian_ms / args.steps,
        "exchange_ms_per_iteration": tm.exchange_ms / args.steps,
        "fp64_retry_rows_per_iteration": tm.fp64_retry_rows / args.steps,
        "iteration_frac_of_hbm_roof": algorithmic_bytes(U, I, nnz, k) / world /
                                      (ms_per_step * 1e-3) / 1e9 / peak,
    }
    gpu_launches = int(tm.launches)

    # ---- parity at the bench workload's own size (every rank checks rows of its own block) --
    parity = None
    if not args.no_parity:
        par = sampled_parity(als, cfg, rank, world)
        worst = max(par["x_half"]["fro"], par["x_half"]["max"], par["y_half"]["fro"], par["y_half"]["max"])
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([worst], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            worst = float(t.item())
        parity = {"checker": "oracle/als_oracle.c (C port of the reference Java ALS) on sampled rows, "
                             "from the device's own opposite factor, after the timed region",
                  "tolerance": 1e-4, "worst_over_ranks": worst, "ok": bool(worst <= 1e-4),
                  "rank0": par, "ranks_checked": world}

    # ---- e2e through the C ABI with HOST buffers ------------------------------------------
    e2e = None
    cpu_baseline = None
    import ctypes as C
    lib = als.lib
    if world == 1 and not args.no_e2e:
        # untimed: bring the device-generated workload to pinned host memory
        h_ptr = torch.empty(U + 1, dtype=torch.int64, pin_memory=True)
        h_idx = torch.empty(nnz, dtype=torch.int32, pin_memory=True)
        h_val = torch.empty(nnz, dtype=torch.float32, pin_memory=True)
        h_x = torch.empty((U, k), dtype=torch.float32, pin_memory=True)
        h_y = torch.empty((I, k), dtype=torch.float32, pin_memory=True)
        als.check(lib.als_get_interactions(als.h, C.cast(h_ptr.data_ptr(), C.POINTER(C.c_int64)),
                                           C.cast(h_idx.data_ptr(), C.POINTER(C.c_int32)),
                                           C.cast(h_val.data_ptr(), C.POINTER(C.c_float))))
        als.close()

        trace = os.environ.get("BENCH_E2E_TRACE") == "1"  # development: phases of the call, with syncs

        def one_call():
            marks = [("start", time.perf_counter())]

            def mark(name):
                if trace:
                    torch.cuda.synchronize()
                    marks.append((name, time.perf_counter()))
            a = M.NativeALS(k, device=local_rank, kernel=kernel)
            a.set_stream(stream.cuda_stream)
            mark("create")
            a.check(lib.als_set_interactions(a.h, U, I, C.cast(h_ptr.data_ptr(), C.POINTER(C.c_int64)),
                                             C.cast(h_idx.data_ptr(), C.POINTER(C.c_int32)),
                                             C.cast(h_val.data_ptr(), C.POINTER(C.c_float))))
            a.n_users, a.n_items = U, I
            mark("set_interactions")
            a.check(lib.als_set_y(a.h, C.cast(h_y0.data_ptr(), C.POINTER(C.c_float))))
            mark("set_y")
            a.iterate(args.steps)
            mark("iterate")
            a.check(lib.als_get_x(a.h, C.cast(h_x.data_ptr(), C.POINTER(C.c_float))))
            a.check(lib.als_get_y(a.h, C.cast(h_y.data_ptr(), C.POINTER(C.c_float))))
            mark("get_x, get_y")
            a.close()
            mark("destroy")
            if trace:
                dbg("e2e phases: " + ", ".join("%s %.0f ms" % (n, (t1 - t0) * 1e3) for (_, t0), (n, t1)
                                                in zip(marks[:-1], marks[1:])))

        one_call()  # warm-up (allocator, first-touch)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one_call()
        torch.cuda.synchronize()
        t_call = time.perf_counter() - t0
        h2d = (U + 1) * 8 + nnz * 8 + I * k * 4
        d2h = (U + I) * k * 4
        e2e = {"value": args.steps / t_call, "unit": "iterations/s",
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "call": "one factorizer call: upload CSR + Y0 from pinned host memory, build the "
                       "by-item orientation on the device, %d iterations, read X and Y back"
                       % args.steps,
               "seconds_per_call": t_call, "finite": bool(torch.isfinite(h_x[:1000]).all())}
        if not args.no_cpu_baseline:
            nu, ni = sample_sizes(cfg, scale=3.0)  # ~15-25 s of CPU work in total
            ptr_np = h_ptr.numpy()
            e1 = int(ptr_np[nu])
            user_rows = (ptr_np[:nu + 1].copy(), h_idx.numpy()[:e1].copy(), h_val.numpy()[:e1].copy())
            item_rows = synth_item_rows(cfg, ni)
            cpu_baseline = cpu_baseline_from_sample(cfg, user_rows, item_rows, h_y0.numpy(),
                                                    h_x.numpy(), "timed once on rank 0")
    elif world > 1 and not args.no_e2e:
        # Sharded e2e: every rank owns a block of users (and of items).  Host side of one call, per
        # rank: upload the rank's by-user CSR block and the full Y0 from pinned memory; the by-item
        # blocks are built on the devices (all-to-all over NCCL); K iterations (finished rows are
        # pushed to the peers' replicas from the solve epilogue); read back the rank's own block
        # of X and of Y.  The handle and its communicator are created outside the timed region
        # (a serving process keeps them across model builds).
        import torch.distributed as dist
        from myrrix_recommender_b200.sharding import local_block
        ub, ue = local_block(U, rank, world)
        ib, ie = local_block(I, rank, world)
        info = als.info()
        lnnz = int(info.nnz)
        h_ptr = torch.empty(ue - ub + 1, dtype=torch.int64, pin_memory=True)
        h_idx = torch.empty(lnnz, dtype=torch.int32, pin_memory=True)
        h_val = torch.empty(lnnz, dtype=torch.float32, pin_memory=True)
        h_x = torch.empty((ue - ub, k), dtype=torch.float32, pin_memory=True)
        h_y = torch.empty((ie - ib, k), dtype=torch.float32, pin_memory=True)
        als.check(lib.als_get_interactions(als.h, C.cast(h_ptr.data_ptr(), C.POINTER(C.c_int64)),
                                           C.cast(h_idx.data_ptr(), C.POINTER(C.c_int32)),
                                           C.cast(h_val.data_ptr(), C.POINTER(C.c_float))))
        als.close()
        a = M.NativeALS(k, device=local_rank, kernel=kernel)
        a.set_stream(stream.cuda_stream)
        uid = [M.factorizer.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        a.comm_init(rank, world, uid[0])

        def one_call():
            a.check(lib.als_set_interactions(a.h, U, I, C.cast(h_ptr.data_ptr(), C.POINTER(C.c_int64)),
                                             C.cast(h_idx.data_ptr(), C.POINTER(C.c_int32)),
                                             C.cast(h_val.data_ptr(), C.POINTER(C.c_float))))
            a.n_users, a.n_items = U, I
            a.check(lib.als_set_y(a.h, C.cast(h_y0.data_ptr(), C.POINTER(C.c_float))))
            a.iterate(args.steps)
            a.sync()
            a.check(lib.als_get_factor_block(a.h, 0, ub, ue - ub, C.cast(h_x.data_ptr(), C.POINTER(C.c_float))))
            a.check(lib.als_get_factor_block(a.h, 1, ib, ie - ib, C.cast(h_y.data_ptr(), C.POINTER(C.c_float))))

        one_call()  # warm-up
        barrier()
        t0 = time.perf_counter()
        one_call()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_call = float(t.item())
        a.close()
        h2d = (ue - ub + 1) * 8 + lnnz * 8 + I * k * 4       # per rank
        d2h = (ue - ub + ie - ib) * k * 4
        e2e = {"value": args.steps / t_call, "unit": "iterations/s",
               "h2d_bytes_per_step": h2d * world / args.steps, "d2h_bytes_per_step": d2h * world / args.steps,
               "call": "one sharded factorizer call on %d ranks: every rank uploads its by-user CSR block + Y0 "
                       "from pinned host memory, the by-item blocks are built on the devices (all-to-all over "
                       "NCCL), %d iterations, every rank reads its own block of X and Y back; max over ranks; "
                       "handle + communicator created outside the timed region" % (world, args.steps),
               "seconds_per_call": t_call, "finite": bool(torch.isfinite(h_x[:1000]).all())}
    else:
        als.close()  # every rank: destroying the NCCL communicator is collective

    if rank == 0:
        line = {
            "metric": "ALS iterations/sec", "value": value, "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[name], "users": U, "items": I,
                       "nnz_per_user": nnz_pu, "features": k, "alpha": 1.0, "lambda": 0.1,
                       "seed": SEED, "kernel": kernel_name,
                       "warp_roles": ("per launch by row length: 8 solve + 7 gather warps below 384 "
                                      "entries/row, else 4 + 11") if kernel_name == "tcgen05" else None,
                       "l2": "inputs (>= %.1f GB per half) exceed the 126 MB L2; no flush needed"
                             % (nnz * 8 / 1e9),
                       "parallelism": "1 GPU" if world == 1 else
                                      "users/items range-sharded over %d GPUs, factor all-gather" % world},
            "updated_rows_per_s": (U + I) * value,
            "roofline": roofline, "parity": parity, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": gpu_launches, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
