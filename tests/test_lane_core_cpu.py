"""The per-lane clique search (gms_b200/csrc/kclique_lane_core.cuh) is __host__ __device__: the exact code the GPU
lanes run — task split by residue class, path-only depth-first search, last-two-levels pair count, set compaction —
is compiled with g++ and checked here against a brute-force count on random DAG bit matrices."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lane_core_against_brute_force():
    exe = os.path.join(ROOT, "build", "lane_core_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wno-unknown-pragmas",
                           os.path.join(ROOT, "tests", "cpp", "lane_core_test.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all checks passed" in r.stdout


def test_schedule_owner_deal_fast_form():
    """tc.cu's owner-of-a-vertex arithmetic (gms_b200/csrc/owner.cuh, __host__ __device__) compiled for the host."""
    exe = os.path.join(ROOT, "build", "owner_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "cpp", "owner_test.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all checks passed" in r.stdout
