"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port of the reference's CPU path on the host
cores) prints ONE JSON line with the keys the driver reads, and the algorithmic-byte figures bench.py's roofline uses
are the ones DESIGN.md / SURVEY.md 8(d) state."""
import json
import os
import subprocess
import sys

from tests.util import ROOT


def test_reference_arm_prints_the_contract_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "msda_fwd_bwd_queries_per_sec"
    assert line["unit"] == "queries/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1 and line["value"] > 0
    assert abs(line["value"] - 22223 / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]   # one frame per step
    assert line["config"]["workload"].startswith("msda_core_op_fwd_bwd") and "model" not in line["config"]
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_algorithmic_bytes_match_the_stated_per_query_figures():
    sys.path.insert(0, ROOT)
    import bench
    s = sum(h * w for h, w in bench.COCO_SHAPES)
    assert s == 22223
    fwd32, bwd32 = bench.algorithmic_bytes(1, 4)
    fwd16, bwd16 = bench.algorithmic_bytes(1, 2)
    assert (fwd32, bwd32) == (3584 * s, 6144 * s)              # fp32 3584 / 6144 B per query
    # 16-bit values / outputs / gradients, fp32 locations and weights: 512 + 1536 + 512 and 3 x 512 + 2 x 1536 + ... by
    # SURVEY 8(d)'s own formula = 2560 / 4608 B per query (its table prints 4096 for the backward: an addition slip)
    assert (fwd16, bwd16) == (2560 * s, 4608 * s)
