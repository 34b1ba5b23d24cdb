"""The reference's configuration of the CPU path, measured at full size: the CPU oracle (C++ restatement of
winterfell 0.9.0 + ProcessorAir) proving the benchmark trace on ONE thread (the reference enables no `concurrent`
feature: /root/reference/Cargo.toml:13).  Minutes of CPU time per 2^20-row proof, so it is run once per round on the
GPU box's host (gpurun) and the result is committed; bench.py attaches it as cpu_baseline.single_thread_recorded.

    python tools/cpu_single_thread.py [kind log_n]...      # default: 2 20  -> profiles/r02_cpu_single_thread.json
"""
import json
import os
import platform
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests import _oracle

o = _oracle.load()
args = [int(x) for x in sys.argv[1:]] or [2, 20]
out_path = ROOT / "profiles" / "r02_cpu_single_thread.json"
rec = json.loads(out_path.read_text()) if out_path.exists() else {}
cpu = ""
try:
    cpu = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
except Exception:
    cpu = platform.processor()
for kind, log_n in zip(args[0::2], args[1::2]):
    trace, pub = o.synthetic_trace(kind, log_n)
    o.lib.orc_set_num_threads(1)
    art = o.prove(trace, pub)
    rec[f"{kind}_{log_n}"] = {"seconds_per_proof": art.seconds, "proofs_per_s": 1.0 / art.seconds, "threads": 1, "cpu": cpu,
                             "host_cores": os.cpu_count(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                             "what": f"oracle.prove of the synthetic kind-{kind} program at 2^{log_n} rows, full size, one thread"}
    print(json.dumps(rec[f"{kind}_{log_n}"]), flush=True)
out_path.write_text(json.dumps(rec, indent=1) + "\n")
