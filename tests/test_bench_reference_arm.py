"""bench.py plumbing that runs without a GPU: the reference arm / cpu_baseline sample and the clock sampler parser."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_sample_small():
    import bench
    r = bench.reference_sample("cfg3", cells=60, n_species=2, threads=2)
    assert r["kind"] in ("reference", "port")
    assert r["value"] > 1e5 and r["cores"] == 2
    assert r["evaluations"] == 60 * 2 * 32 * 24 * 21
    r2 = bench.reference_sample("cfg2", cells=200, n_species=1, threads=2)        # 2+1D sample: 10 cells x 241 eta points
    assert r2["evaluations"] == 50 * 1 * 32 * 24 * 241


def test_reference_arm_prints_one_json_line(monkeypatch):
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", OMP_NUM_THREADS="2")
    code = ("import bench, sys; bench.reference_sample.__defaults__ = (40, 1, 2, 1, 1); "
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1']; bench.main()")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out_lines = r.stdout.strip().splitlines()
    assert len(out_lines) == 1, out_lines          # stdout carries the JSON record only (the reference's progress goes to stderr)
    line = json.loads(out_lines[0])
    assert line["impl"] == "reference" and line["unit"] == "evaluations/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    # non-zero ranks of a torchrun launch do no work and exit 0
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_sample_operation_0():
    import bench
    r = bench.reference_sample("cfg3", cells=160, n_species=2, threads=2, operation=0)      # 40 cells after the operation-0 cut
    assert r["evaluations"] == 40 * 2 * 32 * 24 * 21 and r["value"] > 1e5
    if r["kind"] == "reference":
        assert r["cores"] == 1 and "no OpenMP pragma" in r["sample"]


def test_bench_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1", "--cells", "100"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
