"""bench.py --impl reference (the CPU arm: oracle port timed on the host cores) prints exactly one JSON line with the
keys the driver reads, for a UCI config and for the MNIST config.  No GPU involved."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config", ["power", "mnist"])
def test_reference_arm_prints_one_json_line(config):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config,
                          "--steps", "1", "--warmup", "1", "--batch", "4096"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "PM-VAE train samples/s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_does_not_load_the_product():
    """The CPU arm times the oracle port only: neither the package nor libpmvae.so may be mapped into its process."""
    code = ("import runpy, sys; sys.argv = ['bench.py', '--impl', 'reference', '--config', 'gas', '--batch', '256', '--steps', '1', "
            "'--warmup', '1']; runpy.run_path(%r, run_name='__main__'); "
            "assert not any(m.startswith('posterior_matching_b200') for m in sys.modules), 'package imported'; "
            "assert 'libpmvae' not in open('/proc/self/maps').read(), 'libpmvae.so mapped'" % os.path.join(ROOT, "bench.py"))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["config"]["rows_per_gpu_per_step"] == 256


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
