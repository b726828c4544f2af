"""Host logic of bench.py that can be checked without a GPU: the roofline object and the reference arm's fallbacks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _measurement(kernel):
    # 10 profiled steps of 1000 four point constraints, 10 velocity iterations, 2 ms in the velocity kernel
    return {"prof": {kernel: {"ms": 2.0, "launches": 100}}, "prof_steps": 10, "prof_gpu_ms": 10.0,
            "pagg": {"num_constraints": 10000, "num_contact_points": 40000, "velocity_iterations": 100}}


def test_roofline_object_follows_the_survey_formula():
    r = bench.roofline_of(_measurement("KSolveVelocity"))
    # C(c) + 4 S_v + 4 (3 + c) with C(c) = 220 + 64 c -> 600 B per four point constraint and iteration
    total = 10 * 1000 * 10 * 600.0
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "KSolveVelocity"
    assert abs(r["achieved"] - total / 2.0e-3 / 1e9) < 1e-6
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["bytes_per_launch"] - total / 100) < 1e-6
    # traffic comes from the committed ncu capture of the SAME kernel (profiles/*_solve_traffic.json) or is null
    t = bench.ncu_traffic_ratio()
    if t is not None and t[0].get("kernel") == "KSolveVelocity":
        assert abs(r["traffic"] - t[0]["dram_bytes_per_algorithmic_byte"] * total / 100) < 1e-6
    else:
        assert r["traffic"] is None
    assert abs(r["share_of_step"] - 0.2) < 1e-12


def test_roofline_of_the_one_launch_solver_counts_the_warm_start_pass():
    r = bench.roofline_of(_measurement("KSolveVelocityAll"))
    assert r["kernel"] == "KSolveVelocityAll" and r["passes_per_step"] == 11
    assert abs(r["achieved"] - 10 * 1000 * 11 * 600.0 / 2.0e-3 / 1e9) < 1e-6


def test_both_arms_name_the_same_config():
    import argparse
    a = argparse.Namespace(workload="batch", worlds=4096, bodies=1000000, steps=20, warmup=5, gpus=1)
    c = bench.job_config(a)
    assert c["bodies"] == 4096 * 1240 and c["worlds"] == 4096 and "[5, 25)" in c["window"]
    assert bench.job_config(argparse.Namespace(workload="pile", worlds=4096, bodies=1000000, steps=20, warmup=5, gpus=2))["bodies"] == 1000000


def test_roofline_falls_back_to_the_small_world_solver():
    assert bench.roofline_of(_measurement("KSolveSmallVelocity"))["kernel"] == "KSolveSmallVelocity"
    assert bench.roofline_of({"prof": {}, "prof_steps": 1, "prof_gpu_ms": 1.0, "pagg": {"num_constraints": 0, "num_contact_points": 0, "velocity_iterations": 0}}) is None


def test_reference_arm_prints_one_json_line():
    # a tiny CPU run of the reference arm (2 steps of 16 concurrent small worlds would still be the Pyramid: keep it to the contract check)
    env = dict(os.environ, RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "", "ranks other than 0 exit without work"
