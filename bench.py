#!/usr/bin/env python
"""Benchmark of the trace -> proof path (BASELINE.json metric: proofs/sec at a 2^20-row trace).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n L] [--kind k]

One "step" = one full proof (trace in, serialized proof out).  Prints ONE JSON line (rank 0).

N = 1   the synthetic ciphertext program of BASELINE.json configs[2] (READ2/READ/SMUL/ADD2/SADD, 2^20 rows).
N = 2   the same proof sharded over both GPUs (one prover, coset-sharded)          -> strong scaling
N = 4/8 BASELINE.json configs[3]: the mixed program at 2^22 rows, one proof sharded over all GPUs -> strong scaling
        (`single_gpu_same_config` carries the one-GPU time of the same proof, measured in the same run)

* value    : proofs/s with the trace (N > 1: this rank's column shard) already resident in HBM, device-timed with
             CUDA events on the prover's stream, max over ranks.
* e2e      : the same metric through the host-buffer C-ABI call (ezk_prover_prove) from PAGEABLE host memory - what
             the reference's `TraceTable` columns are (vm/src/lib.rs:18) - proof bytes out, copies inside the timed
             region; `e2e.pinned` is the same from page-locked memory.
* roofline : dominant kernel of the timed region (per-kernel CUDA events inside the library).
* cpu_baseline / --impl reference : the CPU oracle (restated reference path) proving the SAME trace at full size.
Other measurements (pageable-upload sweeps, micro-benchmarks, stage sweeps) live in tools/ and write profiles/*.json.
"""
from __future__ import annotations

import argparse
import datetime
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "proofs_per_sec"
UNIT = "proofs/s"
DTYPE = "u128 (f128 field) + u32 (BLAKE3)"
KINDS = {1: "scalar PUSH/READ/ADD/MUL", 2: "ciphertext READ2/READ/SMUL/ADD2/SADD", 3: "mixed"}


def algo_bytes_per_proof(n: int) -> dict:
    """Minimum compulsory HBM traffic per stage with unfused stages (BASELINE.md / SURVEY 8d), bytes."""
    return {"trace_lde": 4032 * n, "trace_commit": 4352 * n, "constraints": 3712 * n, "composition": 2928 * n,
            "deep": 1280 * n, "fri": 274 * n}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run (B200_PROFILING.md recipe); samples
    are attributed to the timed regions by their timestamps."""
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.windows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def window(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows, smax = [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 11:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, util = float(parts[2]), float(parts[10])
                smax.append(float(parts[3]))
            except ValueError:
                continue
            rows.append((ts, clk, util, {name for name, val in zip(names, parts[6:10]) if val.lower().startswith("active")}))
        timed = [r for r in rows if any(a - 0.01 <= r[0] <= b + 0.01 for a, b in self.windows)]
        # fall back to the samples taken under load when the timestamps cannot be matched (clock skew, short regions)
        used = timed or [r for r in rows if r[2] >= 50.0] or rows
        reasons = set().union(*[r[3] for r in used]) if used else set()
        sm = [r[1] for r in used]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "samples_in_timed_regions": len(timed), "samples_total": len(rows), "reasons": sorted(reasons)}


def workload_config(args, world: int) -> dict:
    ref = {20: "configs[2]", 22: "configs[3]", 16: "configs[1]"}.get(args.log_n, "sweep size")
    return {"workload": f"synthetic {KINDS[args.kind]} program padded to 2^{args.log_n} trace rows, 28 columns, blowup 8, "
                        f"32 queries, FRI folding 8 (BASELINE.json {ref})",
            "trace_rows": 1 << args.log_n, "lde_rows": 8 << args.log_n, "field": "f128", "hash": "blake3-256",
            "parallelism": "1 GPU" if world == 1 else f"one proof sharded over {world} GPUs by LDE coset (NCCL)",
            "l2": "inputs larger than L2 (trace 28*n*16 B, LDE 8x that)"}


def resolve_workload(args, world: int):
    """configs[2] on 1 and 2 GPUs, configs[3] (2^22 rows, mixed program) on 4 and 8 unless the caller names a size."""
    if args.log_n is None:
        args.log_n = 22 if world >= 4 else 20
    if args.kind is None:
        args.kind = 3 if args.log_n >= 22 else 2


# ----------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle proves the SAME trace at full size (no scaling model)

def cpu_oracle_prove(trace, pub, threads: int):
    from tests import _oracle
    o = _oracle.load()
    o.lib.orc_set_num_threads(threads)
    art = o.prove(trace, pub)
    return art.seconds, art.proof


def run_reference(args, rank: int, world: int, out):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust prover cannot be built here (no
    cargo/rustc, winterfell 0.9.0 not vendored), so this arm times the CPU oracle - the restated reference algorithm -
    with all host threads (NOT the reference's configuration, which is single-threaded: Cargo.toml:13 enables no
    `concurrent` feature) on the full-size workload.  Every step is one full proof; the number of timed steps is capped
    by wall time, not by shrinking the trace.  Loads oracle/liborc.so only (its own trace generator)."""
    if rank != 0:
        return
    resolve_workload(args, world)
    from tests import _oracle
    o = _oracle.load()
    threads = os.cpu_count() or 1
    o.lib.orc_set_num_threads(threads)
    t_gen = time.perf_counter()
    trace, pub = o.synthetic_trace(args.kind, args.log_n)
    t_gen = time.perf_counter() - t_gen
    budget = args.cpu_budget_s
    t_start = time.perf_counter()
    warm = 0
    secs = []
    # one warm-up proof only when a proof is cheap enough to leave room for a timed one
    first, _ = cpu_oracle_prove(trace, pub, threads)
    if args.warmup >= 1 and first * 3 < budget:
        warm = 1
    else:
        secs.append(first)
    while len(secs) < args.steps and (not secs or time.perf_counter() - t_start + statistics.mean(secs) < budget):
        s, _ = cpu_oracle_prove(trace, pub, threads)
        secs.append(s)
    per = statistics.mean(secs)
    value = 1.0 / per
    sample = (f"oracle (C++ restatement of winterfell 0.9.0 + ProcessorAir) proving the full 2^{args.log_n}-row trace: "
              f"{len(secs)} proofs of {per:.2f} s on {threads} OpenMP threads - not the reference's configuration "
              f"(single thread); timed steps capped at {budget:.0f} s of wall time")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(secs),
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "trace_generation_s": t_gen, "wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """Everything the libraries print (NCCL's version banner goes to stdout) is sent to stderr; the returned file is
    the real stdout, used only for the ONE JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=None)
    ap.add_argument("--kind", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall-time cap of the CPU legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # the CPU arm needs no process group: when it is started without torchrun, --gpus N still names the workload of N GPUs
        run_reference(args, rank, max(world, args.gpus), out)
        return
    resolve_workload(args, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import encrypt_zkvm_b200 as ezk
    from encrypt_zkvm_b200 import parallel

    if ezk.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device visible; this backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return parallel.max_over_ranks(x, device="cuda")

    def sum_over_ranks(x: float) -> float:
        return parallel.sum_over_ranks(x, device="cuda")

    # ---- workload: the host VM builds the trace (north star: trace generation stays on the host); every rank builds
    # the same one (N > 1: one proof, each rank uploads / keeps only the columns it owns) ----
    n = 1 << args.log_n
    warm_req = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
    prog, ex = ezk.synthetic_case(args.kind, args.log_n)
    trace_np = ex.trace()                                              # pageable numpy memory, (28, n, 2) uint64
    program_hash, outputs = prog.hash(), ex.outputs()
    pub = program_hash + outputs
    pinned = torch.from_numpy(trace_np.view(np.int64)).pin_memory()    # page-locked copy for e2e.pinned
    pinned_np = pinned.numpy().view(np.uint64)
    dev = pinned.to(f"cuda:{local_rank}", non_blocking=False)          # resident copy for the `value` arm
    del ex
    h2d_bytes = 28 * n * 16

    prover = ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey(), device=local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- N > 1: one GPU alone on the same proof first (strong-scaling denominator), then join the group ----
    single = None
    if world > 1:
        for _ in range(2):
            ref_proof = prover.prove_device(dev.data_ptr(), n).to_bytes()
        barrier()
        prover.timer_start()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            prover.prove_device(dev.data_ptr(), n)
        ms1 = max_over_ranks(prover.timer_stop()) / reps
        single = {"ms_per_proof": ms1, "proofs_per_s": 1e3 / ms1, "steps": reps}
        prover.join_group()

    # ---- device-resident arm ----
    for _ in range(warm_req):
        proof = prover.prove_device(dev.data_ptr(), n)
    ezk.profile_enable(True)
    ezk.profile_reset()
    barrier()
    launches0 = ezk.kernel_launch_count()
    w0 = time.time()
    prover.timer_start()
    for _ in range(args.steps):
        proof = prover.prove_device(dev.data_ptr(), n)
    ms = prover.timer_stop()
    sampler.window(w0, time.time())
    barrier()
    launches = ezk.kernel_launch_count() - launches0
    profile = ezk.profile_read()
    ezk.profile_enable(False)
    stage_ms = prover.stage_times_ms()
    t_dev = max_over_ranks(ms / 1e3)
    value = args.steps / t_dev if world > 1 else world * args.steps / t_dev
    proof_bytes_dev = proof.to_bytes()

    # ---- end-to-end arms: host trace in, proof bytes out (pageable = the drop-in caller; pinned as a sub-key) ----
    def e2e_arm(host_array):
        for _ in range(2):
            p = prover.prove(host_array)
        barrier()
        w0 = time.time()
        prover.timer_start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p = prover.prove(host_array)
        dev_ms = prover.timer_stop()
        wall = time.perf_counter() - t0
        sampler.window(w0, time.time())
        barrier()
        return max_over_ranks(max(dev_ms / 1e3, wall)), p.to_bytes()

    t_e2e, proof_e2e = e2e_arm(trace_np)
    t_pin, proof_pin = e2e_arm(pinned_np)
    # N = 1, sub-key: the eight bookkeeping columns (clk, op bits, flag, depth) generated on the device from the operation
    # list instead of uploaded (ezk_prover_prove_ops): 20 columns + one byte per operation cross PCIe
    e2e_ops = None
    if world == 1:
        codes = prog.op_codes()
        for _ in range(2):
            p_ops = prover.prove_with_ops(trace_np, codes)
        w0 = time.time()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            p_ops = prover.prove_with_ops(trace_np, codes)
        torch.cuda.synchronize()
        t_ops = time.perf_counter() - t0
        sampler.window(w0, time.time())
        e2e_ops = {"value": args.steps / t_ops, "ms_per_step": t_ops * 1e3 / args.steps, "h2d_bytes_per_step": 20 * n * 16 + len(codes),
                   "identical_bytes": p_ops.to_bytes() == proof_e2e}
    clocks = sampler.stop()
    launches_total = int(sum_over_ranks(float(launches)))
    h2d_step = h2d_bytes if world == 1 else sum(n * 16 for c in range(28) if c % world == rank)
    h2d_step = int(sum_over_ranks(float(h2d_step))) if world > 1 else h2d_step

    # ---- the timed proofs are checked: verifier of the product, identical bytes across the three arms (and, N > 1,
    # with the single-GPU proof); the oracle comparison follows with the CPU baseline ----
    checks = {"verified": False, "arms_identical": proof_bytes_dev == proof_e2e == proof_pin,
              "proof_sha256": hashlib.sha256(proof_bytes_dev).hexdigest()[:16]}
    try:
        prover.verify(proof_bytes_dev)
        checks["verified"] = True
    except Exception as e:
        checks["verify_error"] = f"{type(e).__name__}: {e}"[:200]
    if world > 1:
        same = sum_over_ranks(1.0 if proof_bytes_dev == ref_proof else 0.0) == world
        checks["bytes_identical_to_single_gpu"] = bool(same)
        prover.leave_group()

    # ---- N = 1, informational: two provers on this GPU, one host thread each ----
    pipelined = None
    if world == 1:
        try:
            import threading
            extra = ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey(), device=local_rank)
            pair, got = [prover, extra], [None, None]

            def work(k, reps):
                for _ in range(reps):
                    got[k] = pair[k].prove_device(dev.data_ptr(), n).to_bytes()

            for reps in (2, args.steps):
                th = [threading.Thread(target=work, args=(k, reps)) for k in range(2)]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                torch.cuda.synchronize()
                pipe_wall = time.perf_counter() - t0
            extra.close()
            pipelined = {"provers_per_gpu": 2, "proofs_per_s": 2 * args.steps / pipe_wall,
                         "identical_bytes": got[0] == got[1] == proof_bytes_dev}
        except Exception as e:  # optional mode: the contract lines above must survive its failure
            pipelined = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- N = 1, informational: BASELINE.json configs[1] (scalar program, 2^16 rows) on the same prover ----
    config1 = None
    if world == 1 and args.log_n == 20:
        try:
            p1, e1 = ezk.synthetic_case(1, 16)
            t1 = torch.from_numpy(e1.trace().view(np.int64)).to(f"cuda:{local_rank}")
            with ezk.ExecutionProver(ezk.ProofOptions(), p1.hash(), e1.outputs(), ezk.ServerKey(), device=local_rank) as q:
                for _ in range(3):
                    q.prove_device(t1.data_ptr(), 1 << 16)
                q.timer_start()
                for _ in range(20):
                    pr = q.prove_device(t1.data_ptr(), 1 << 16)
                ms16 = q.timer_stop() / 20
                q.verify(pr)
            config1 = {"workload": "BASELINE.json configs[1]: scalar PUSH/READ/ADD/MUL program, 2^16 rows", "ms_per_proof": ms16,
                       "proofs_per_s": 1e3 / ms16, "verified": True}
            del t1
        except Exception as e:
            config1 = {"error": f"{type(e).__name__}: {e}"[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (device time from CUDA events around every launch) ----
    peak, peak_kind = hbm_peak()
    top = max(profile.items(), key=lambda kv: kv[1]["ms"]) if profile else None
    roofline = None
    if top:
        name, st = top
        achieved = st["algo_bytes"] / (st["ms"] * 1e-3) / 1e9 if st["ms"] > 0 else 0.0
        traffic = None
        tf = ROOT / "profiles" / "roofline_traffic.json"
        if tf.exists() and world == 1 and args.log_n == 20:
            try:
                traffic = json.loads(tf.read_text()).get(name)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "peak_source": peak_kind,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "launches_per_step": st["launches"] / args.steps,
                    "avg_launch_ms": st["ms"] / st["launches"],
                    "algo_bytes_per_launch": st["algo_bytes"] / st["launches"],
                    "share_of_step": st["ms"] / (ms if ms else 1.0)}
    kernels = {k: {"ms": round(v["ms"] / args.steps, 4), "n": round(v["launches"] / args.steps, 1),
                   "GBps": round(v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
               for k, v in sorted(profile.items(), key=lambda kv: -kv[1]["ms"])[:8]}
    ab = algo_bytes_per_proof(n)
    stages = {k: {"ms": round(stage_ms[k], 4),
                  "frac_of_hbm_peak": round(ab[k] / world / (stage_ms[k] * 1e-3) / 1e9 / peak, 4) if k in ab and stage_ms[k] > 0 else None}
              for k in stage_ms}

    # the ceiling that actually binds: the integer pipes; peak = the modmul micro-benchmark
    lde_modmuls = 28 * (n * (args.log_n / 2 + 2) + 8 * n * (args.log_n / 2 + 3)) / world
    int_pipe = {"stage": "trace_lde", "achieved_gmodmul_per_s": lde_modmuls / (stage_ms["trace_lde"] * 1e-3) / 1e9,
                "peak_gmodmul_per_s": 327.0}
    int_pipe["frac"] = int_pipe["achieved_gmodmul_per_s"] / int_pipe["peak_gmodmul_per_s"]

    # ---- CPU baseline (rank 0, N = 1 only): the oracle proves the SAME full-size trace; its proof bytes double as the
    # parity check of the timed GPU proofs ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        secs, want = cpu_oracle_prove(trace_np, pub, threads)
        checks["bytes_identical_to_cpu_oracle"] = want == proof_bytes_dev
        cpu = {"value": 1.0 / secs, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"one full proof of the same 2^{args.log_n}-row trace by the CPU oracle: {secs:.2f} s on {threads} OpenMP "
                         f"threads (not the reference's configuration: Winterfell runs single-threaded there)"}
        rec = ROOT / "profiles" / "r02_cpu_single_thread.json"  # measured once by tools/cpu_single_thread.py on the GPU box
        if rec.exists():
            try:
                cpu["single_thread_recorded"] = json.loads(rec.read_text()).get(f"{args.kind}_{args.log_n}")
            except Exception:
                pass

    ok = checks["verified"] and checks["arms_identical"] and checks.get("bytes_identical_to_cpu_oracle", True) and \
        checks.get("bytes_identical_to_single_gpu", True)
    line = {
        "metric": METRIC, "value": value if ok else None, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm_req,
        "ms_per_step": t_dev * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": args.steps / t_e2e if ok else None, "unit": UNIT, "h2d_bytes_per_step": h2d_step,
                "d2h_bytes_per_step": len(proof_e2e), "ms_per_step": t_e2e * 1e3 / args.steps,
                "host_memory": "pageable (numpy array; staged through the prover's page-locked ring)",
                "pinned": {"value": args.steps / t_pin, "ms_per_step": t_pin * 1e3 / args.steps},
                "bookkeeping_columns_on_device": e2e_ops},
        "gpu_launches": launches_total, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "checks": checks,
        "int_pipe_roofline": int_pipe, "stages": stages, "kernels": kernels, "proof_bytes": len(proof_bytes_dev),
        "single_gpu_same_config": single, "pipelined": pipelined, "config1_2p16": config1,
        "hbm_roofline_proofs_per_s": peak * 1e9 / sum(ab.values()),
    }
    if single:
        line["speedup_vs_one_gpu"] = single["ms_per_proof"] / (t_dev * 1e3 / args.steps)
    if not ok:
        line["error"] = "the timed proofs failed their checks (see `checks`): no value is reported"
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
