"""bench.py's CPU-only parts: the reference arm (the CPU oracle at full size of whatever workload is named), its JSON
contract, and the workload each GPU count resolves to.  No GPU is touched."""
import json
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_line_is_self_consistent():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--log-n", "10", "--kind", "1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "proofs_per_sec" and d["unit"] == "proofs/s"
    assert d["steps"] == 2 and d["steps_requested"] == 2 and d["gpu_launches"] == 0
    assert abs(d["ms_per_step"] - 1000.0 / d["value"]) < 1e-6 * d["ms_per_step"]  # one number, two units
    assert d["e2e"] == {"value": d["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1
    assert "scaled" not in cb["sample"] and "2^10" in cb["sample"]  # the full trace is proved, nothing is extrapolated
    assert d["config"]["trace_rows"] == 1 << 10


def test_reference_arm_is_silent_on_other_ranks():
    env = {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "PATH": "/usr/bin:/bin"}
    import os
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--log-n", "10"],
                       capture_output=True, text=True, timeout=120, env={**os.environ, **env})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_names_the_workload_of_the_gpu_count_without_torchrun():
    """`bench.py --impl reference --gpus 4` started as a plain process (no WORLD_SIZE) still resolves to configs[3];
    checked through --help-free argument resolution only (a 2^22 CPU proof takes minutes)."""
    sys.path.insert(0, str(ROOT))
    import bench
    src = (ROOT / "bench.py").read_text()
    assert "run_reference(args, rank, max(world, args.gpus), out)" in src
    a = SimpleNamespace(log_n=None, kind=None, gpus=4)
    bench.resolve_workload(a, max(1, a.gpus))
    assert (a.log_n, a.kind) == (22, 3)


def test_workload_per_gpu_count():
    """configs[2] (2^20, ciphertext program) on 1 and 2 GPUs, configs[3] (2^22, mixed) on 4 and 8; explicit flags win."""
    sys.path.insert(0, str(ROOT))
    import bench
    for world, want in [(1, (20, 2)), (2, (20, 2)), (4, (22, 3)), (8, (22, 3))]:
        a = SimpleNamespace(log_n=None, kind=None, gpus=world)
        bench.resolve_workload(a, world)
        assert (a.log_n, a.kind) == want
        cfg = bench.workload_config(a, world)
        assert cfg["trace_rows"] == 1 << want[0] and ("sharded" in cfg["parallelism"]) == (world > 1)
    a = SimpleNamespace(log_n=16, kind=1, gpus=8)
    bench.resolve_workload(a, 8)
    assert (a.log_n, a.kind) == (16, 1)
    assert sum(bench.algo_bytes_per_proof(1 << 20).values()) == (4032 + 4352 + 3712 + 2928 + 1280 + 274) << 20
