"""Host logic of the run kernel's planner (panda3d_b200/csrc/p3cs_runplan.h) on the CPU: every row in exactly one item, items
inside their tile and run, the half-warp residue rule that makes the row walk bank-conflict free, and the two-palette-copy
assignment (an augmenting-path matching) against the exhaustive optimum.  tests/run_plan_check.cpp does the work."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_run_plan_invariants_and_two_copy_matching(tmp_path):
    exe = str(tmp_path / "run_plan_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "run_plan_check.cpp")], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(res.stdout[-3000:])
    assert res.returncode == 0 and "FAIL" not in res.stdout and res.stdout.rstrip().endswith("ok all"), res.stdout[-3000:] + res.stderr
