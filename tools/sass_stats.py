"""Static SASS statistics of the built kernels (cuobjdump -sass on the object files of encrypt_zkvm_b200/_build):
instructions per kernel and the counts of the mnemonics that matter for the integer-pipe analysis of DESIGN.md.

    python tools/sass_stats.py [object-name-substring] [kernel-name-substring]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
objs = sorted((ROOT / "encrypt_zkvm_b200" / "_build").glob("*.o"))
want_obj = sys.argv[1] if len(sys.argv) > 1 else ""
want_k = sys.argv[2] if len(sys.argv) > 2 else ""
KEYS = ["IMAD.WIDE", "IMAD", "IADD3", "LOP3", "SHF", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "ISETP", "SEL", "PRMT", "MOV"]
for obj in objs:
    if want_obj not in obj.name:
        continue
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    name, counts = None, None
    def flush():
        if name and want_k in name:
            total = sum(counts["__all__"].values())
            parts = []
            for k in KEYS:
                c = sum(v for m, v in counts["__all__"].items() if (m.startswith(k) if k != "IMAD" else (m.startswith("IMAD") and not m.startswith("IMAD.WIDE"))))
                if c:
                    parts.append(f"{k}={c}")
            print(f"{obj.name}: {name[:110]}\n    instructions={total} " + " ".join(parts))
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, counts = m.group(1), {"__all__": collections.Counter()}
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and counts is not None:
            counts["__all__"][m.group(1)] += 1
    flush()
