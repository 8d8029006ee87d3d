"""CPU: the JSON line of `bench.py --impl reference` (the arm the driver runs beside the GPU arm) carries every key of the bench
contract, on a bounded sample of the headline workload (the CPU restatement of the whole hot path, oracle/prover.py)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-log-n", "15"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ntt_field_ops_per_s" and line["unit"] == "field-ops/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert "starknet layout, 2^22 Cairo steps" in line["config"]["workload"]            # the GPU arm's config, bounded sample stated
    assert "2^11-step" in line["config"]["sample"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and e2e["value"] > 0 and e2e["unit"] == line["unit"]
    # every stage of the GPU step has its CPU counterpart
    for stage in ("lde_base", "merkle_base", "lde_ext", "merkle_ext", "constraint_eval", "ntt_comp_inv", "ntt_comp_fwd", "merkle_comp", "ood", "deep", "fri"):
        assert line["stages_s"][stage] > 0, stage
