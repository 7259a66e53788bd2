"""bench.py's reference arm (the CPU oracle port on the host cores) and the JSON contract shared by both arms.
The GPU arm itself needs a B200; its line is checked by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--ref-seconds", "0.5", "--workload", "kirk_default_100x100x1000"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "bellman_state_control_updates_per_s"
    assert line["unit"] == "updates/s" and line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["config"]["workload"] == "kirk_default_100x100x1000" and line["config"]["grid"] == [100, 100]
    assert line["config"]["controls"] == 1000 and "l2" in line["config"] and "step" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_both_arms_share_one_config_object():
    sys.path.insert(0, ROOT)
    import bench
    import bellman_b200 as bb
    d = bench.make_desc(bb, "kirk_scaled_8192x8192x512")
    c = bench.workload_config("kirk_scaled_8192x8192x512", d)
    assert c == {"workload": "kirk_scaled_8192x8192x512", "grid": [8192, 8192], "controls": 512, "problems": 1,
                 "step": "one backward stage over the whole grid",
                 "l2": "J_{k+1} (537 MB) exceeds the 126 MB L2; no flush needed"}
