import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import myrrix_recommender_b200 as M
from oracle import oracle as O
rng = np.random.default_rng(11)
n_items, k = 4000, 32
lens = np.concatenate([[1, 2, 3, 31, 32, 33, 63, 64, 65, 3000], np.minimum(3000, (rng.pareto(1.2, 300) * 5 + 1).astype(int))])
ptr, idx, val = [0], [], []
for n in lens:
    idx += list(np.sort(rng.choice(n_items, size=n, replace=False)))
    val += list(rng.integers(1, 6, size=n).astype(np.float32))
    ptr.append(len(idx))
ptr, idx, val = np.array(ptr, np.int64), np.array(idx, np.int32), np.array(val, np.float32)
d = rng.standard_normal((n_items, k))
Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
tp, ti, tv = O.csr_transpose(ptr, idx, val, n_items)
kern = int(sys.argv[1]) if len(sys.argv) > 1 else 1
with M.NativeALS(k, kernel=kern) as als:
    als.set_interactions(len(lens), n_items, ptr, idx, val)
    als.set_y(Y0)
    cp, ci, cv = als.get_interactions(by_column=True)
    print("transpose equal", np.array_equal(cp, tp), np.array_equal(ci, ti), np.array_equal(cv, tv))
    Gd = als.gramian("y"); Go = O.transpose_times_self(Y0)
    print("G_Y err", np.abs(Gd - Go).max())
    als.half_x(); als.sync(); X = als.get_x()
    Xo = np.zeros_like(X); O.als_half(ptr, idx, val, Y0, Go, Xo)
    ex = np.abs(X - Xo).max(axis=1)
    w = np.argsort(-ex)[:8]
    print("X half: max abs err", ex.max(), "max|X|", np.abs(Xo).max())
    for u in w: print("  user", u, "len", lens[u], "err", ex[u], "row max", np.abs(Xo[u]).max())
    Gd = als.gramian("x"); Gx = O.transpose_times_self(X)
    print("G_X err", np.abs(Gd - Gx).max(), np.abs(Gx).max())
    als.half_y(); als.sync(); Y = als.get_y()
    Yo = Y0.copy(); O.als_half(tp, ti, tv, X, Gx, Yo)
    ey = np.abs(Y - Yo).max(axis=1)
    cnt = np.diff(tp)
    w = np.argsort(-ey)[:8]
    print("Y half (from device X): max abs err", ey.max(), "max|Y|", np.abs(Yo).max())
    for i in w: print("  item", i, "cnt", cnt[i], "err", ey[i], "row max", np.abs(Yo[i]).max())
print("---- 2 full iterations vs oracle.als_run")
Xo2, Yo2, its, _ = O.als_run(ptr, idx, val, n_items, Y0, max_iterations=2, convergence_threshold=1e-12, n_threads=8)
print("oracle iterations", its)
with M.NativeALS(k, kernel=kern) as als:
    als.set_interactions(len(lens), n_items, ptr, idx, val)
    als.set_y(Y0)
    als.iterate(2); als.sync()
    X2, Y2 = als.get_x(), als.get_y()
for name, a, b, cnts in (("X", X2, Xo2, lens), ("Y", Y2, Yo2, np.diff(tp))):
    e = np.abs(a - b).max(axis=1)
    print(name, "max abs err", e.max(), "max|ref|", np.abs(b).max(), "fro rel", np.linalg.norm(a-b)/np.linalg.norm(b))
    for r in np.argsort(-e)[:6]: print("   row", r, "cnt", cnts[r], "err", e[r], "row max", np.abs(b[r]).max())
# manual oracle chain with explicit halves
Y = Y0.copy(); X = np.zeros((len(lens), k), np.float32)
for it in range(2):
    G = O.transpose_times_self(Y); O.als_half(ptr, idx, val, Y, G, X)
    G = O.transpose_times_self(X); O.als_half(tp, ti, tv, X, G, Y)
print("manual chain vs als_run:", np.abs(X - Xo2).max(), np.abs(Y - Yo2).max())
print("manual chain vs gpu    :", np.abs(X - X2).max(), np.abs(Y - Y2).max())
