"""bench.py's contract on the side that runs without a GPU: the CPU arm (`--impl reference`) prints exactly one JSON line with
the keys the driver reads, on the same workload string as the GPU arm; under torchrun only rank 0 works and prints; and the GPU
arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)


def test_reference_arm_line():
    p = run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--pairs", "20000"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GCUPS" and d["unit"] == "GCUPS" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("BASELINE.json configs[1]") and d["config"]["pairs_per_gpu"] == 20000
    # the GPU arm names its workload with the same function: the driver compares the two strings
    sys.path.insert(0, ROOT)
    import bench
    class A:
        dist, ref_bases = 0, 16_000_000
    assert bench.workload_config(A, 8, 20000)["workload"] == d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    p = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--pairs", "1000"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    p = run(["--steps", "1", "--warmup", "1", "--pairs", "1000", "--no-aux"])
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]          # no number without a GPU


def test_config2_checksum_annotation():
    """The strong-scaling leg puts the CPU checker's full-size checksum beside the GPU's for the exact workload it was computed on
    (and only for that one); the constant is the one on record under profiles/."""
    sys.path.insert(0, ROOT)
    import bench
    rec = json.load(open(os.path.join(ROOT, "profiles", "config2_full_checksum_cpu_r02.json")))
    assert rec["pairs"] == 100_000_000 and rec["equal"] is True
    assert bench.CONFIG2_CPU_CHECKSUM[(100_000_000, 0, 150, 500)] == rec["checksum64"] == rec["gpu_checksum64"]
    r = bench.annotate_strong({"checksum64": rec["checksum64"]}, 100_000_000, 0, 150, 500)
    assert r["equals_cpu_checker_on_all_pairs"] is True and r["cpu_checker_checksum64"] == rec["checksum64"]
    r = bench.annotate_strong({"checksum64": "0" * 16}, 100_000_000, 0, 150, 500)
    assert r["equals_cpu_checker_on_all_pairs"] is False
    r = bench.annotate_strong({"checksum64": rec["checksum64"]}, 50_000_000, 0, 150, 500)       # another workload: no claim
    assert "equals_cpu_checker_on_all_pairs" not in r
