"""bench.py's reference arm (the CPU side of the contract) on the smallest workloads: one JSON line with the keys the
driver reads.  The GPU arm needs a device and is exercised by the driver."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # exactly one JSON line on stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("args,env", [
    (["--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0"], None),
    (["--impl", "reference", "--workload", "regular_400", "--steps", "1", "--warmup", "0"], {"VRT_REG_SHAPE": "20,14,14"}),
    (["--impl", "reference", "--workload", "searchlight", "--steps", "1", "--warmup", "0"], None),
])
def test_reference_arm_prints_one_contract_line(args, env):
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "output_sites")) or os.path.exists("/root/reference/rt_preprocessing/output_sites")
    if args[3] in ("small", "searchlight") and not staged:
        pytest.skip("voro++ driver not staged")
    d = run(args, env)
    assert KEYS <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "updates/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    if args[3] != "regular_400":
        assert d["libvrt_mapped"] is False               # the CPU arm never loads the product library
