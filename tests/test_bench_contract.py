"""CPU suite, part 4: bench.py's reference arm runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "chain-steps/s" and line["dtype"] == "f64" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"] == "cfg2" and line["vs_baseline"] is None
    # the config object is built by one function for both arms, so the driver's `same_config` comparison holds
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config_block("cfg2", bench.WORKLOADS["cfg2"], 1)
    for key in ("chains_per_gpu", "seed", "eps", "len", "parallelism", "N", "d", "sampler"):
        assert key in bench.config_block("cfg4", bench.WORKLOADS["cfg4"], 8)
