"""The C++ host facade (include/gms_b200/*.hpp) over the C ABI: compile on any host, run on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gms_b200", "lib")


def build(src, out):
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), src, "-L" + LIBDIR, "-lgmsb",
                           "-Wl,-rpath," + LIBDIR, "-o", out])
    return out


def test_facade_compiles_and_fails_loudly_without_gpu(gms):
    exe = build(os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), os.path.join(ROOT, "build", "facade_test"))
    build(os.path.join(ROOT, "examples", "triangle_counting.cpp"), os.path.join(ROOT, "build", "triangle_counting"))
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 3 and "CUDA" in r.stderr        # GMSB_ERR_CUDA, never a CPU fallback


@pytest.mark.gpu
def test_facade_on_gpu(gms):
    exe = build(os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), os.path.join(ROOT, "build", "facade_test"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


def test_sets_facade_compiles_and_fails_loudly_without_gpu(gms):
    exe = build(os.path.join(ROOT, "tests", "cpp", "sets_test.cpp"), os.path.join(ROOT, "build", "sets_test"))
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 3 and "CUDA" in r.stderr


@pytest.mark.gpu
def test_reference_set_cases_on_gpu(gms):
    """testing/sets.cpp (54 typed cases) with CudaSortedSet as the set type, plus Tomita's pivot rule and one
    Bron-Kerbosch expansion written against the Set concept."""
    exe = build(os.path.join(ROOT, "tests", "cpp", "sets_test.cpp"), os.path.join(ROOT, "build", "sets_test"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_example_on_gpu(gms):
    exe = build(os.path.join(ROOT, "examples", "triangle_counting.cpp"), os.path.join(ROOT, "build", "triangle_counting"))
    r = subprocess.run([exe, "-g", "16", "-n", "2", "-v", "-k", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("PASS") == 2 and "triangles=15656287" in r.stdout and "count=291383976" in r.stdout


@pytest.mark.gpu
def test_reference_harness_drop_in(gms):
    """The reference's own benchmark main + verifier with the graph type swapped (oracle/dropin_tc.cpp); the binary
    embeds reference code, so it is built only where /root/reference exists and travels under oracle/_ref/."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_tc")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_tc not built (needs /root/reference)")
    r = subprocess.run([exe, "-g", "kronecker", "14", "--deg", "16", "-n", "2", "-v"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("@@@")]
    cuda = [ln for ln in lines if "CudaSetGraph" in ln]
    assert len(cuda) == 4 and all("PASS" in ln for ln in cuda), r.stdout[-3000:]     # 2 kernels x 2 trials
    assert all("PASS" in ln for ln in lines)
