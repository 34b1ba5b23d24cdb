"""Runs the stand-alone micro-benchmarks of tools/ubench on the visible GPU and writes their output to
profiles/r02_ubench.json (integer-pipe issue costs, the f128 product, DFMA beside the integer pipes).

    python tools/run_ubench.py
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from encrypt_zkvm_b200 import build

res = {}
for exe in build.build_ubench():
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    res[exe.name] = [ln.strip() for ln in r.stdout.splitlines() if ln.strip()] if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-400:]}
(ROOT / "profiles" / "r02_ubench.json").write_text(json.dumps(res, indent=1) + "\n")
print(json.dumps(res, indent=1))
