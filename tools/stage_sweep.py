"""BASELINE.json configs[4]: LDE + Merkle + FRI stage sweep, 2^18..2^24 rows x trace width, device-resident.

    python tools/stage_sweep.py [--min 18] [--max 24] [--widths 7,14,28,56] [--iters 3] [--cpu-log 14] [--out FILE]

One JSON line per (log_n, width): interpolation + coset LDE (blowup 8) of `width` random columns, BLAKE3 row hash +
Merkle tree over the 8n rows, and (once per log_n) the FRI layer loop over one column of 8n evaluations.  Times are
CUDA-event times inside the library (`ezk_bench_lde_merkle`, `ezk_bench_fri`); GB/s figures are the ALGORITHMIC
bytes of SURVEY 8(d) (LDE 9 n W 16; Merkle 8n (16 W + 3*32); FRI 274 n) over the measured time, next to the measured
HBM copy peak of MEASURED_PEAKS.json.  Shapes whose working set (W (2n + 16n) + 32n elements) exceeds --mem-gb are
skipped and say so.  `--cpu-log k` also times the CPU oracle (test infrastructure; one column LDE + Merkle of 2^k rows,
1 thread) so that the line carries a CPU figure scaled by n log n for the same shape.
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import encrypt_zkvm_b200 as ezk  # noqa: E402


def hbm_peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured"
    except Exception:
        return 7700.0, "nominal"


def cpu_reference(log_k: int):
    """seconds per (column x row) unit of the CPU oracle at 2^log_k rows: (lde_s_per_column, merkle_s_per_row_of_28)"""
    import numpy as np
    sys.path.insert(0, str(ROOT))
    from tests import _oracle
    orc = _oracle.load()
    orc.lib.orc_set_num_threads(1)
    n = 1 << log_k
    rng = np.random.default_rng(7)
    col = rng.integers(0, 1 << 62, size=(n, 2), dtype=np.uint64)
    col[:, 1] >>= 2
    t0 = time.perf_counter()
    orc.lde_column(col)
    lde_s = time.perf_counter() - t0
    rows = rng.integers(0, 1 << 62, size=(8 * n, 28, 2), dtype=np.uint64)
    t0 = time.perf_counter()
    orc.merkle_rows(rows)
    mk_s = time.perf_counter() - t0
    return lde_s, mk_s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=18)
    ap.add_argument("--max", type=int, default=24)
    ap.add_argument("--widths", default="7,14,28,56")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--mem-gb", type=float, default=150.0)
    ap.add_argument("--cpu-log", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    widths = [int(w) for w in args.widths.split(",")]
    peak, peak_src = hbm_peak()
    cpu = None
    if args.cpu_log:
        lde_s, mk_s = cpu_reference(args.cpu_log)
        cpu = {"log_n": args.cpu_log, "lde_s_per_column": lde_s, "merkle_s_width28": mk_s, "threads": 1}
    out = open(args.out, "w") if args.out else None

    def emit(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    with ezk.ExecutionProver(ezk.ProofOptions(), [0, 0], [0] * 16, ezk.ServerKey()) as p:
        for log_n in range(args.min, args.max + 1):
            n = 1 << log_n
            for w in widths:
                need_gb = (w * 18 * n + 32 * n + 1024) * 16 / 1e9
                rec = {"stage": "lde_merkle", "log_n": log_n, "width": w, "working_set_gb": round(need_gb, 2)}
                if need_gb > args.mem_gb:
                    rec["skipped"] = f"working set above {args.mem_gb} GB"
                    emit(rec)
                    continue
                lde_ms, mk_ms = p.bench_lde_merkle(w, n, args.iters)
                lde_bytes = 9 * n * w * 16
                mk_bytes = 8 * n * (16 * w + 3 * 32)
                rec.update({"lde_ms": round(lde_ms, 4), "lde_algo_gbs": round(lde_bytes / lde_ms / 1e6, 1),
                            "lde_frac_hbm": round(lde_bytes / lde_ms / 1e6 / peak, 4),
                            "merkle_ms": round(mk_ms, 4), "merkle_algo_gbs": round(mk_bytes / mk_ms / 1e6, 1),
                            "merkle_frac_hbm": round(mk_bytes / mk_ms / 1e6 / peak, 4),
                            "hbm_peak_gbs": peak, "peak_source": peak_src})
                if cpu:
                    scale = (n * log_n) / ((1 << cpu["log_n"]) * cpu["log_n"])
                    rec["cpu_lde_ms_scaled"] = round(cpu["lde_s_per_column"] * w * scale * 1e3, 1)
                    rec["cpu_merkle_ms_scaled"] = round(cpu["merkle_s_width28"] * (n >> cpu["log_n"]) * (16 * w + 96) / (16 * 28 + 96) * 1e3, 1)
                    rec["cpu"] = f"oracle, 1 thread, measured at 2^{cpu['log_n']} rows and scaled"
                emit(rec)
            fri_ms = p.bench_fri(n, args.iters)
            fri_bytes = 274 * n
            emit({"stage": "fri", "log_n": log_n, "fri_ms": round(fri_ms, 4), "fri_algo_gbs": round(fri_bytes / fri_ms / 1e6, 1),
                  "fri_frac_hbm": round(fri_bytes / fri_ms / 1e6 / peak, 4), "hbm_peak_gbs": peak})
    if out:
        out.close()


if __name__ == "__main__":
    main()
