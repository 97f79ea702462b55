"""bench.py --impl reference on the CPU (the reference arm of the measurement contract): it runs EXACTLY the --steps /
--warmup it is given — a step is one frame of the oracle, or a centred band of rows when the run would not fit its time
budget — and prints the contract's JSON line.  Rank > 0 under torchrun exits without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env, *args):
    env = dict(os.environ, **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cornell_512"] + list(args),
                       capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_runs_the_steps_it_is_given():
    lines = run({}, "--steps", "3", "--warmup", "2")
    line = json.loads(lines[-1])
    assert line["impl"] == "reference" and line["steps"] == 3 and line["warmup"] == 2 and line["n_gpus"] == 1
    assert line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert "3 steps of 1 full frame" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "cornell" in line["config"]["workload"]
    full = line

    # a budget the full frames cannot meet: the same number of steps, each a band of rows, and a lower spp/s to match
    lines = run({"PTC_REFERENCE_BUDGET_S": "0.05"}, "--steps", "4", "--warmup", "1")
    band = json.loads(lines[-1])
    assert band["steps"] == 4 and band["warmup"] == 1
    assert "4 steps of rows" in band["cpu_baseline"]["sample"] and "of 512 rows" in band["cpu_baseline"]["sample"]
    assert band["ms_per_step"] < full["ms_per_step"]
    assert 0.3 < band["value"] / full["value"] < 3.0          # a band of the same image traces at a comparable rate


def test_reference_arm_other_ranks_do_nothing():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "2", "--warmup", "1") == []
