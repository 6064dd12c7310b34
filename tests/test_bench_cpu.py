"""bench.py on the CPU: the reference arm's JSON line (driven through the UNMODIFIED reference under baseline/_ref when it is
installed; the oracle port otherwise) carries the keys of the benchmark contract and the same `config` the B200 arm prints,
and the FLOP accounting of the estimator is self-consistent."""

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_sweep_flop_accounting():
    g1, g5 = bench.sweep_gflop("large", 1), bench.sweep_gflop("large", 5)
    # one magnitude: the shared work is not amortised; five magnitudes: a fifth of it per pair
    assert g1["executed_per_pair"] > g5["executed_per_pair"] > 0
    assert abs(g5["executed_per_image"] - 5 * g5["executed_per_pair"]) < 1e-6
    # reference-faithful 70.25 GFLOP/pair (ViT-B), 246.2 (ViT-L)
    assert abs(bench.sweep_gflop("base", 1)["reference_faithful_per_pair"] - 70.25) < 0.05
    assert abs(g1["reference_faithful_per_pair"] - 246.2) < 0.1
    assert bench.sweep_gflop("base", 1)["executed_per_pair"] < 0.6 * 70.25


def test_strong_scaling_split_and_config():
    class A:
        scaling, batch, global_batch = "strong", 512, 512

    assert bench.per_rank_batch(A, 1) == 512 and bench.per_rank_batch(A, 8) == 64
    A.scaling = "weak"
    assert bench.per_rank_batch(A, 8) == 512
    c1 = bench.finetune_config("base", 64, 8, [], 85806346, "strong")
    assert c1["global_batch"] == 512 and c1["parallelism"] == "dp8" and c1["scaling"] == "strong"
    with pytest.raises(SystemExit):
        A.scaling, A.global_batch = "strong", 510
        bench.per_rank_batch(A, 8)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "img/s" and line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 1
    # "reference" when baseline/_ref is installed (the build container, the GPU box), "port" (the oracle) otherwise
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    if (ROOT / "baseline" / "_ref" / "vitef").is_dir():
        assert line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"] == {"value": line["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same `config` object the B200 arm prints for this workload
    assert line["config"] == bench.finetune_config("base", 512, 1, [], 85806346, "strong")


def test_kernel_regression_guard_on_kept_logs():
    """tools/kernel_regression.py on logs kept under profiles/: round 1's drift of the attention forward (236 -> 263 us while the
    torch controls of the same runs moved by 5 %) is flagged, the step from the end of round 1 to the start of round 2 is
    clean, and the final library against the start of round 2 shows the attention backward more than 25 % faster."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    tool, prof = str(root / "tools" / "kernel_regression.py"), root / "profiles"

    def run(new, base):
        return subprocess.run([sys.executable, tool, str(prof / new), str(prof / base)], capture_output=True, text=True)

    drift = run("r02_z_kernel_bench.log", "r01_m_kernel_bench.log")
    assert drift.returncode == 1 and "attention fwd" in drift.stdout.split("FAIL")[-1]
    clean = run("r03_a_kernel_bench.log", "r02_z_kernel_bench.log")
    assert clean.returncode == 0 and "OK: no kernel regressed" in clean.stdout
    final = run("r04_g_kernel_bench.log", "r03_a_kernel_bench.log")
    line = next(l for l in final.stdout.splitlines() if l.startswith("attention bwd"))
    assert float(line.rsplit("normalised x", 1)[1].split()[0]) < 0.75, line
