"""The JSON lines bench.py prints are what the round driver parses: keep their keys and types pinned.
CPU: the reference arm (oracle port on the host cores).  GPU: the B200 arm with a short run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": float,
             "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict, "cpu_baseline": dict,
             "e2e": dict, "gpu_launches": int}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]           # exactly ONE JSON line
    return json.loads(lines[0])


def check_base(d):
    for k, t in BASE_KEYS.items():
        assert k in d, k
        assert isinstance(d[k], t) or (t is float and isinstance(d[k], int)), (k, type(d[k]))
    assert d["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-step-seconds", "0.4")
    check_base(d)
    # the same `config` object as the B200 arm prints (the driver compares them)
    assert {"workload", "family", "envs_per_gpu", "nodes", "elements", "free_dofs", "l2", "actions", "parallelism"} <= set(d["config"])
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"] > 0


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--steps", "6", "--warmup", "3", "--cpu-seconds", "0.5", "--batch", "512")
    check_base(d)
    assert "impl" not in d or d["impl"] == "b200"
    assert d["gpu_launches"] == 6 * 3                     # Pareto branch + fused actor network + env-step per step
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert d["status_nonzero_envs"] == 0
    # round-2 keys: per-rank link rates, the compact-state end-to-end number, the legs at BASELINE configs 3 / 4 shapes
    assert d["e2e"]["h2d_gbs_per_rank"] > 0 and d["e2e"]["d2h_gbs_per_rank"] > 0
    assert d["e2e_compact_state"]["d2h_bytes_per_step"] < d["e2e"]["d2h_bytes_per_step"]
    ex = d["extra"]
    assert ex["config3"]["family"] == "small_roof" and ex["config3"]["envs_per_gpu"] == 16384 and ex["config3"]["value"] > 0
    assert ex["config4"]["family"] == "large_bridge" and ex["config4"]["envs_per_gpu"] == 1024
    assert 0 < ex["config3"]["roofline_fem"]["frac"] < 1 and ex["actor_pareto_P50"]["P"] == 50
    assert d["config"]["nodes"] == 16 and d["config"]["free_dofs"] == 28


@pytest.mark.gpu
def test_training_leg_line():
    d = run_bench("--train", "--steps", "3")
    assert d["family"] == "large_roof" and d["envs_per_gpu"] == 1024 and d["value"] > 0 and d["n_gpus"] == 1
    assert d["models_identical_across_ranks"] is True and d["learner_update"].startswith("one captured CUDA graph")
    assert set(d["stages_ms"]) >= {"rollout_3x_act_3x_env_step", "replay_push", "learner_train_update"}


def test_e2e_piece_schedule_argument():
    """--e2e-pieces: 'auto' (HostRollout decides), a count of equal pieces, explicit sizes, or fractions of the batch"""
    import bench
    assert bench.e2e_piece_schedule("auto", 4096) == "auto"
    assert bench.e2e_piece_schedule("3", 4096) == 3
    assert bench.e2e_piece_schedule("512,1536,2048", 4096) == [512, 1536, 2048]
    frac = bench.e2e_piece_schedule("0.125,0.375,0.5", 4096)
    assert frac == [512, 1536, 2048] and all(v % 32 == 0 for v in frac[:-1])
    with pytest.raises(SystemExit):
        bench.e2e_piece_schedule("512,512", 4096)
