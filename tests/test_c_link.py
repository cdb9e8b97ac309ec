"""A C program LINKED against libmyrrix_als.so (tests/c/golden_link.c): the boundary is a real C
ABI, not something only ctypes can call.  CPU run: builds, links, runs and must report "no CUDA
device" through the library's own status code (exit 77: there is no CPU fallback).  GPU run: the
reference's golden AlternatingLeastSquaresTest.testALS through that C program."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "myrrix-recommender_b200")


def _build(tmp_path):
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        g = json.load(f)["als"]
    R, Y0, P = g["R"], g["Y0"], g["product"]

    def arr(name, a, ty):
        rows = ",\n  ".join("{" + ", ".join(repr(float(v)) + ("f" if ty == "float" else "") for v in r) + "}" for r in a)
        return "static const %s %s[%d][%d] = {\n  %s};\n" % (ty, name, len(a), len(a[0]), rows)
    hdr = ("#define N_USERS %d\n#define N_ITEMS %d\n#define FEATURES %d\n#define THRESHOLD %r\n#define MAX_ITER %d\n"
           % (len(R), len(R[0]), g["features"], g["threshold"], g["max_iterations"]))
    hdr += arr("R", R, "float") + arr("Y0", Y0, "float") + arr("PRODUCT", P, "double")
    (tmp_path / "golden_data.h").write_text(hdr)
    exe = str(tmp_path / "golden_link")
    lib = os.path.join(PKG, "libmyrrix_als.so")
    if not os.path.exists(lib):
        import sys
        sys.path.insert(0, PKG)
        import build
        build.build()
    subprocess.check_call(["gcc", "-O1", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", str(tmp_path),
                           os.path.join(ROOT, "tests", "c", "golden_link.c"), "-o", exe,
                           "-L", PKG, "-lmyrrix_als", "-lm", "-Wl,-rpath," + PKG])
    return exe


def test_c_program_links_and_reports_no_device_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
    assert "ALS_E_CUDA" in r.stdout


@pytest.mark.gpu
def test_c_program_reproduces_the_reference_golden(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
