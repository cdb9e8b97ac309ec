"""Where does each warp role of the tcgen05 row-update kernel spend its time?
Builds a -DALS_PROFILE_WAITS copy of the library (scripts/_prof/libmyrrix_als.so), runs a few
iterations of a BASELINE config and prints, per barrier kind, the average cycles per CTA a
role's warps were blocked on it, next to the kernel's total cycles."""
import ctypes as C, os, subprocess, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF = os.path.join(ROOT, "scripts", "_prof")
os.makedirs(PROF, exist_ok=True)
lib = os.path.join(PROF, "libmyrrix_als.so")
if "--build" in sys.argv:
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                           "-DALS_PROFILE_WAITS", *[a for a in sys.argv if a.startswith("-D")], "-Xcompiler", "-fPIC", "-shared", "-o", lib,
                           os.path.join(ROOT, "myrrix-recommender_b200", "csrc", "als_abi.cu"), "-ldl"])
    sys.exit(0)
import myrrix_recommender_b200 as M
M._native.LIB_PATH = lib
cfgs = {"c2": (1000000, 100000, 50, 32), "c3p": (2000000, 200000, 100, 64), "c3": (10000000, 1000000, 100, 64)}
name = sys.argv[1] if len(sys.argv) > 1 else "c3p"
U, I, nnz, k = cfgs[name]
L = M._native.load()
L.als_debug_wait_cycles.argtypes = [C.POINTER(C.c_uint64)]
names = ["prod:b_empty", "prod:empty", "mma:acc_empty", "mma:full", "drain:w_empty", "drain:acc_full",
         "chol:w_full", "chol:b_full", "kernel total (per CTA sum)", "chol:lockstep", "chol:factor_solve",
         "prod:raw_full"]
warps = [7, 7, 1, 1, 4, 4, 8, 8, 1, 8, 8, 7]
if name == "c2":
    warps = [7, 7, 1, 1, 4, 4, 12, 12, 1, 12, 12, 7]
with M.NativeALS(k) as als:
    als.synth_interactions(U, I, nnz, seed=1234567890)
    als.synth_y0(seed=1234567890)
    als.iterate(2); als.sync()
    buf = (C.c_uint64 * 16)()
    L.als_debug_wait_cycles(buf)
    for half, fn in (("X", als.half_x), ("Y", als.half_y)):
        fn(); als.sync()
        L.als_debug_wait_cycles(buf)
        tot = buf[8] / 148.0
        if half == "Y":  # 4 Cholesky + 11 producer warps
            warps = [11, 11, 1, 1, 4, 4, 4, 4, 1, 4, 4, 11]
        print("%s-half: kernel cycles per CTA %.3e" % (half, tot))
        for i in (0, 1, 11, 2, 3, 4, 5, 6, 7, 9, 10):
            per_warp = buf[i] / 148.0 / warps[i]
            print("   %-16s blocked %5.1f%% of the kernel (avg per warp)" % (names[i], 100.0 * per_warp / tot))
