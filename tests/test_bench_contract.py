"""bench.py's contract on a box without a GPU: the reference arm (`--impl reference`) is the CPU oracle and nothing else — one
JSON line with the driver's keys, the product library never mapped into the process — and the product arm
refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WRAP = r"""
import sys, runpy
sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1']
try:
    runpy.run_path('bench.py', run_name='__main__')
except SystemExit as e:
    assert not e.code, e.code
maps = open('/proc/self/maps').read()
print('MAPS', int('libqspush' in maps), int('libqs_oracle' in maps))
"""


def test_reference_arm_is_the_oracle_and_never_maps_the_product():
    r = subprocess.run([sys.executable, "-c", _WRAP], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    js = [l for l in lines if l.startswith("{")]
    assert len(js) == 1                                           # ONE JSON line
    d = json.loads(js[0])
    assert d["impl"] == "reference" and d["metric"] == "sqp_rti_iterations_per_sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("config3") and d["config"]["horizon"] == 40
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    maps = [l for l in lines if l.startswith("MAPS")][0].split()
    assert maps[1] == "0" and maps[2] == "1"                      # libqspush.so not mapped, the oracle is


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU path" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]      # no bench line is printed


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """N > 1: the driver launches the reference arm like the product arm (torchrun, one rank per GPU); rank 0 alone runs the oracle
    on the N-GPU workload's config (config 4, strong scaling) and prints, the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29513", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    js = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(js) == 1
    d = json.loads(js[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["workload"].startswith("config4")
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
