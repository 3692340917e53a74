"""CPU-only checks of bench.py's contract: the reference arm prints exactly one JSON line with the required keys, and every
kernel named in a workload's roofline accounting exists in the library sources (kernel renames must not silently turn the
reported dominant kernel into the wrong one)."""
import glob
import json
import os
import re
import subprocess
import sys

import pytest

import cabi

ROOT = cabi.ROOT
sys.path.insert(0, ROOT)


def launched_kernel_names():
    names = set()
    for path in glob.glob(os.path.join(ROOT, "ppo.c_b200", "csrc", "*.cu")):
        src = open(path).read()
        for m in re.finditer(r"B200_LAUNCH(?:_PDL|_COOP)?\(\s*\(?\s*([A-Za-z0-9_]+(?:<[^()]*?>)?)", src):
            names.add(m.group(1).replace(" ", ""))
        for m in re.finditer(r"profile_mark\(\"\(?([A-Za-z0-9_]+)", src):
            names.add(m.group(1))
    return names


def test_roofline_keys_name_real_kernels():
    import bench
    names = launched_kernel_names()
    assert len(names) > 25
    for wname, cls in bench.WORKLOADS.items():
        wl = cls(None, 0, 1)
        for probe in ({}, {"fused_phase_kernel": {}}):          # c2 / c1 account differently when the persistent kernel ran
            for key in wl.roofline_work(probe):
                assert any(key in n for n in names), "%s: roofline key %r matches no launched kernel" % (wname, key)


@pytest.mark.parametrize("workload", ["c5", "adam"])
def test_reference_arm_prints_one_json_line(workload):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]:
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
