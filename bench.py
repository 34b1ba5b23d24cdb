#!/usr/bin/env python
"""Benchmark of the trace -> proof path (BASELINE.json metric: proofs/sec at a 2^20-row trace).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 20] [--kind 2]

One "step" = one full proof of the synthetic ciphertext program (READ2/READ/SMUL/ADD2/SADD, BASELINE.md
config 2) padded to 2^log_n trace rows.  Prints ONE JSON line (rank 0).

* value     : proofs/s with the trace already resident in HBM (ezk_prover_prove_device), device-timed with CUDA
              events on the prover's stream, max over ranks.  N > 1: every rank proves its own trace on its own
              GPU (independent proofs, no data-path collective) -> weak scaling.
* e2e       : the same metric through the host-buffer C-ABI call (ezk_prover_prove): pinned host trace in,
              proof bytes out, copies inside the timed region.
* roofline  : dominant kernel of the timed region (per-kernel CUDA events inside the library).
* cpu_baseline / --impl reference : the CPU oracle (restated reference path) on a bounded sample.
"""
from __future__ import annotations

import argparse
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
CPU_SAMPLE_LOG_N = 16


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
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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
            if len(parts) < 10:
                continue
            try:
                clk, util = float(parts[1]), float(parts[9])
                smax.append(float(parts[2]))
            except ValueError:
                continue
            rows.append((clk, util, {name for name, val in zip(names, parts[5:9]) if val.lower().startswith("active")}))
        # the sampler runs from before the warm-up proofs: keep the samples taken under load
        loaded = [r for r in rows if r[1] >= 50.0] or rows
        reasons = set().union(*[r[2] for r in loaded]) if loaded else set()
        sm = [r[0] for r in loaded]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "samples_total": len(rows), "reasons": sorted(reasons)}


def cpu_oracle_run(log_n: int, kind: int, threads: int, repeats: int = 1):
    """Times the CPU oracle's prove() on a bounded sample; returns (seconds per proof, threads used)."""
    import encrypt_zkvm_b200 as ezk
    from tests import _oracle
    o = _oracle.load()
    o.lib.orc_set_num_threads(threads)
    prog, ex = ezk.synthetic_case(kind, log_n)
    trace, pub = ex.trace(), prog.hash() + ex.outputs()
    best = None
    for _ in range(repeats):
        art = o.prove(trace, pub)
        best = art.seconds if best is None else min(best, art.seconds)
    return best, threads


def scale_to_full(seconds_sample: float, sample_log_n: int, log_n: int) -> float:
    """proofs/s at 2^log_n extrapolated from a 2^sample_log_n sample with the n*log2(n) work model."""
    factor = (2 ** (log_n - sample_log_n)) * (log_n / sample_log_n)
    return 1.0 / (seconds_sample * factor)


def run_reference(args, rank: int, out):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust prover cannot be built here
    (no cargo/rustc, winterfell 0.9.0 not vendored), so this arm times the CPU oracle - the restated reference
    algorithm - with all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = min(CPU_SAMPLE_LOG_N, args.log_n)
    for _ in range(args.warmup and 1):  # one warm-up proof is enough for a CPU run that takes seconds
        cpu_oracle_run(sample, args.kind, threads)
    t0 = time.perf_counter()
    secs = []
    for _ in range(args.steps):
        s, _ = cpu_oracle_run(sample, args.kind, threads)
        secs.append(s)
    wall = time.perf_counter() - t0
    per = sum(secs) / len(secs)
    value = scale_to_full(per, sample, args.log_n)
    sample_txt = (f"oracle prove of the same synthetic program at 2^{sample} rows: {per:.2f} s/proof on {threads} threads "
                  f"(OpenMP); scaled to 2^{args.log_n} rows by n*log2(n) (x{2 ** (args.log_n - sample) * args.log_n / sample:.1f})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u128 (f128 field) + u32 (BLAKE3)", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), file=out, flush=True)


def workload_config(args) -> dict:
    kinds = {1: "scalar PUSH/READ/ADD/MUL", 2: "ciphertext READ2/READ/SMUL/ADD2/SADD", 3: "mixed"}
    return {"workload": f"synthetic {kinds[args.kind]} program padded to 2^{args.log_n} trace rows, 28 columns, "
                        f"blowup 8, 32 queries, FRI folding 8 (BASELINE.json configs[2])",
            "trace_rows": 1 << args.log_n, "lde_rows": 8 << args.log_n, "field": "f128", "hash": "blake3-256",
            "parallelism": f"{args.gpus} independent provers (one per GPU)" if args.gpus > 1 else "1 GPU",
            "l2": "inputs larger than L2 (trace 28*n*16 B, LDE 8x that)"}


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
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--kind", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, out)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import encrypt_zkvm_b200 as ezk

    if ezk.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device visible; this backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from encrypt_zkvm_b200 import parallel

    def max_over_ranks(x: float) -> float:
        return parallel.max_over_ranks(x, device="cuda")

    def sum_over_ranks(x: float) -> float:
        return parallel.sum_over_ranks(x, device="cuda")

    # ---- workload: host VM builds the trace (north star: trace generation stays on the host) ----
    n = 1 << args.log_n
    # EZK_TRACE_CACHE=dir (profiling sessions only) keeps the generated case between invocations: the host VM needs
    # 2-3 s for 2^20 rows, all of it outside the timed regions
    seed = parallel.unit_seed(0xE2C0DE00, args.log_n, rank)
    cache = os.environ.get("EZK_TRACE_CACHE")
    cache_file = Path(cache) / f"bench_{args.kind}_{args.log_n}_{seed}.pkl" if cache else None
    if cache_file and cache_file.exists():
        import pickle
        trace_np, program_hash, outputs = pickle.loads(cache_file.read_bytes())
        ex = None
    else:
        prog, ex = ezk.synthetic_case(args.kind, args.log_n, seed=seed)
        trace_np = ex.trace()
        program_hash, outputs = prog.hash(), ex.outputs()
        if cache_file:
            import pickle
            cache_file.parent.mkdir(parents=True, exist_ok=True)
            cache_file.write_bytes(pickle.dumps((trace_np, program_hash, outputs), protocol=4))
    host = torch.from_numpy(trace_np.view(np.int64)).pin_memory()      # (28, n, 2) pinned
    host_np = host.numpy().view(np.uint64)
    dev = host.to(f"cuda:{local_rank}", non_blocking=False)            # resident copy for the `value` arm
    del trace_np, ex
    h2d_bytes = 28 * n * 16

    prover = ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey(), device=local_rank)

    # ---- device-resident arm ----
    # the clock sampler starts before the warm-up proofs (nvidia-smi needs ~0.2 s to deliver its first sample and the
    # timed region of a few proofs is shorter than that); warm-up and timed steps run the same kernels back to back
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # at least W warm-up proofs, and at least 0.6 s of them: nvidia-smi's utilisation figure (which marks the samples
    # taken under load) is averaged over a window longer than a handful of proofs
    warmups, t_w = 0, time.perf_counter()
    while warmups < max(args.warmup, 3) or time.perf_counter() - t_w < 0.6:
        proof = prover.prove_device(dev.data_ptr(), n)
        warmups += 1
    ezk.profile_enable(True)
    ezk.profile_reset()
    barrier()
    launches0 = ezk.kernel_launch_count()
    prover.timer_start()
    for _ in range(args.steps):
        proof = prover.prove_device(dev.data_ptr(), n)
    ms = prover.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = ezk.kernel_launch_count() - launches0
    profile = ezk.profile_read()
    ezk.profile_enable(False)
    stage_ms = prover.stage_times_ms()
    t_dev = max_over_ranks(ms / 1e3)
    value = world * args.steps / t_dev

    # ---- end-to-end arm: pinned host trace in, proof bytes out ----
    for _ in range(2):
        proof = prover.prove(host_np)
    barrier()
    prover.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proof = prover.prove(host_np)
    ms_e2e_dev = prover.timer_stop()
    wall_e2e = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(max(ms_e2e_dev / 1e3, wall_e2e))
    e2e_value = world * args.steps / t_e2e
    proof_bytes = len(proof)
    launches_total = int(sum_over_ranks(float(launches)))

    # ---- proof-level pipelining, informational: two provers on this GPU, one host thread each, the
    # same resident trace.  The latency-bound stretches of one proof (Merkle tops, FRI tail, transcript round trips)
    # are filled with the kernels of the other.  Wall clock between device synchronisations (several streams).
    pipelined, pipe_wall, pipe_same, pipe_err = None, 0.0, 0.0, None
    try:  # no collective inside: a rank that fails here must not leave the others waiting
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
        pipe_same = 1.0 if got[0] == got[1] == proof.to_bytes() else 0.0
        extra.close()
    except Exception as e:  # optional mode: the contract lines above must survive its failure
        pipe_err = f"{type(e).__name__}: {e}"[:300]
    ok_ranks = sum_over_ranks(0.0 if pipe_err else 1.0)
    t_pipe = max_over_ranks(pipe_wall)
    same_ranks = sum_over_ranks(pipe_same)
    if ok_ranks == world and t_pipe > 0:
        pipelined = {"provers_per_gpu": 2, "proofs_per_s": world * 2 * args.steps / t_pipe,
                     "ms_per_proof": t_pipe * 1e3 / (2 * args.steps), "identical_bytes": bool(same_ranks == world),
                     "timing": "wall clock between device synchronisations, max over ranks"}
    else:
        pipelined = {"error": pipe_err or "failed on another rank"}

    # ---- N > 1: the same GPUs as ONE prover (coset-sharded single proof, NCCL all-gathers) ----
    sharded = None
    if world > 1:
        try:
            prog0, ex0 = ezk.synthetic_case(args.kind, args.log_n, seed=parallel.unit_seed(0xE2C0DE00, args.log_n, 0))
            dev0 = torch.from_numpy(ex0.trace().view(np.int64)).to(f"cuda:{local_rank}")
            with ezk.ExecutionProver(ezk.ProofOptions(), prog0.hash(), ex0.outputs(), ezk.ServerKey(), device=local_rank) as sp:
                ref_bytes = sp.prove_device(dev0.data_ptr(), n).to_bytes()
                sp.join_group()
                for _ in range(2):
                    got = sp.prove_device(dev0.data_ptr(), n).to_bytes()
                barrier()
                sp.timer_start()
                for _ in range(args.steps):
                    got = sp.prove_device(dev0.data_ptr(), n).to_bytes()
                ms_sh = sp.timer_stop()
                barrier()
                sp.leave_group()
            t_sh = max_over_ranks(ms_sh / 1e3)
            same = sum_over_ranks(1.0 if got == ref_bytes else 0.0) == world
            sharded = {"ms_per_proof": t_sh * 1e3 / args.steps, "proofs_per_s": args.steps / t_sh, "gpus_per_proof": world,
                       "speedup_vs_one_gpu": (t_dev / args.steps) / (t_sh / args.steps),
                       "bytes_identical_to_single_gpu": bool(same),
                       "collectives": "ncclAllGather of leaf digests (2x), constraint evaluations, DEEP evaluations, opened rows"}
            del dev0, ex0
        except Exception as e:  # the throughput line above must survive a failure of the optional mode
            sharded = {"error": f"{type(e).__name__}: {e}"[:300]}

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
        if tf.exists():
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
    kernels = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                   "GBps": (v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None}
               for k, v in sorted(profile.items(), key=lambda kv: -kv[1]["ms"])}
    ab = algo_bytes_per_proof(n)
    stages = {k: {"ms": stage_ms[k], "GBps": (ab[k] / (stage_ms[k] * 1e-3) / 1e9) if k in ab and stage_ms[k] > 0 else None,
                  "frac_of_hbm_peak": (ab[k] / (stage_ms[k] * 1e-3) / 1e9 / peak) if k in ab and stage_ms[k] > 0 else None}
              for k in stage_ms}

    # ---- the ceiling that actually binds: the integer pipes.  Peak = the modmul micro-benchmark
    # (tools/ubench/field_ubench.cu, 327 G modmul/s per B200); the LDE stage does W * (n (log n / 2 + 2) +
    # 8 n (log n / 2 + 3)) modular products (butterfly twiddles, coset / inter-pass factors, interpolation scaling).
    lde_modmuls = 28 * (n * (args.log_n / 2 + 2) + 8 * n * (args.log_n / 2 + 3))
    int_pipe = {"stage": "trace_lde", "modmuls": lde_modmuls, "achieved_gmodmul_per_s": lde_modmuls / (stage_ms["trace_lde"] * 1e-3) / 1e9,
                "peak_gmodmul_per_s": 327.0, "peak_source": "tools/ubench/field_ubench.cu on this pool's B200 (profiles/README.md)"}
    int_pipe["frac"] = int_pipe["achieved_gmodmul_per_s"] / int_pipe["peak_gmodmul_per_s"]

    # ---- CPU baseline (rank 0, N = 1 only): the oracle, single thread like the reference's configuration ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = min(CPU_SAMPLE_LOG_N, args.log_n)
        secs, threads = cpu_oracle_run(sample, args.kind, 1)
        cpu = {"value": scale_to_full(secs, sample, args.log_n), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle prove of the same synthetic program at 2^{sample} rows: {secs:.2f} s on 1 thread (the "
                         f"reference runs Winterfell single-threaded); scaled to 2^{args.log_n} rows by n*log2(n)"}

    # ---- informational: the same proofs from PAGEABLE host memory (what a Rust Vec column is), plain and staged
    # upload, in a subprocess so that nothing it does can touch the contract's numbers above ----
    pageable = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, str(ROOT / "tools" / "pageable_e2e.py"), str(args.log_n), str(args.kind),
                                str(args.steps), str(local_rank)], capture_output=True, text=True, timeout=300)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            pageable = json.loads(last[-1]) if r.returncode == 0 and last else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as e:
            pageable = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- informational: the stand-alone micro-benchmarks of tools/ubench (integer-pipe issue costs, the f128 product,
    # and whether DFMA issues beside the integer pipes - DESIGN.md sections 4 and 9); a few milliseconds of GPU time
    # each, after everything above is measured ----
    ubench = None
    if world == 1 and not args.no_cpu_baseline:
        ubench = {}
        env = {**os.environ, "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank))}
        for src in sorted((ROOT / "tools" / "ubench").glob("*.cu")):
            exe = src.with_suffix("")
            if not exe.exists():
                continue
            try:
                r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, env=env)
                ubench[exe.name] = ([ln.strip() for ln in r.stdout.splitlines() if ln.strip()] if r.returncode == 0
                                    else {"error": (r.stderr or r.stdout)[-300:]})
            except Exception as e:
                ubench[exe.name] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- informational: experimental build variants (encrypt_zkvm_b200/build.py VARIANTS) against the default build, each
    # in its own subprocess: same proof bytes?  stage times? ----
    variants = None
    if world == 1 and not args.no_cpu_baseline:
        variants = {}
        libs = [ROOT / "encrypt_zkvm_b200" / "libezkvm.so"] + sorted((ROOT / "encrypt_zkvm_b200").glob("libezkvm_*.so"))
        if len(libs) > 1:
            # each library with the default NTT launch shape (256 threads x 3 CTAs/SM, <= 80 registers) and with the
            # 256 x 2 shape (<= 128 registers: no spills for the precomputed-form twiddles)
            for lib, shape in [(lib, shape) for lib in libs for shape in ("0", "2") if not (lib is libs[0] and shape == "2")]:
                key = lib.name if shape == "0" else f"{lib.name}@ntt_shape{shape}"
                try:
                    r = subprocess.run([sys.executable, str(ROOT / "tools" / "variant_probe.py"), str(args.log_n), str(args.kind),
                                        "5", str(local_rank)], capture_output=True, text=True, timeout=300,
                                       env={**os.environ, "EZKVM_LIB": str(lib), "EZK_NTT_VARIANT": shape})
                    last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                    variants[key] = json.loads(last[-1]) if r.returncode == 0 and last else {"error": (r.stderr or r.stdout)[-300:]}
                except Exception as e:
                    variants[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            want = variants.get("libezkvm.so", {}).get("proof_sha256")
            for name, v in variants.items():
                if "proof_sha256" in v:
                    v["same_bytes_as_default"] = bool(want) and v["proof_sha256"] == want

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmups,
        "ms_per_step": t_dev * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u128 (f128 field) + u32 (BLAKE3)", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": proof_bytes,
                "ms_per_step": t_e2e * 1e3 / args.steps},
        "gpu_launches": launches_total, "clocks": clocks, "roofline": roofline, "int_pipe_roofline": int_pipe, "cpu_baseline": cpu,
        "stages": stages, "kernels": kernels, "proof_bytes": proof_bytes, "sharded_single_proof": sharded,
        "pipelined": pipelined, "pageable_e2e": pageable, "ubench": ubench, "variants": variants,
        "hbm_roofline_proofs_per_s": peak * 1e9 / sum(ab.values()),
    }
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
