"""SASS of the f128 field primitives in isolation (tools/sass/field_primitives.cu, compiled for sm_100a): per primitive
the instruction count between the operand loads and the result store, the mix by mnemonic and the full listing.
Writes profiles/r02_sass_field_primitives.md - the evidence behind the instruction counts quoted in DESIGN.md section 4.

    python tools/sass_primitives.py
"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tools" / "sass" / "field_primitives.cu"
OUT = ROOT / "profiles" / "r02_sass_field_primitives.md"

with tempfile.TemporaryDirectory() as tmp:
    cubin = Path(tmp) / "p.cubin"
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", "-o", str(cubin), str(SRC)]
    subprocess.run(cmd, check=True)
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True, check=True).stdout
    ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2:]

funcs = re.split(r"\n\s*Function : ", sass)[1:]
doc = ["# SASS of the field primitives (sm_100a)", "",
       "`python tools/sass_primitives.py` — `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin tools/sass/field_primitives.cu`, " + "; ".join(ver) + ".",
       "Counts are the instructions strictly between the last operand load (`LDG`) and the first result store (`STG`)",
       "of each probe kernel: the primitive itself, the flag update (`VIMNMX`) of the branch-free forms and the address",
       "of the store (one `LDC` + one `IMAD.WIDE`, which ptxas schedules into this window) - subtract 2 for the primitive",
       "alone: 58 for the flagged product, 38 in precomputed form, 11-12 for an addition, 9 for a subtraction. In a",
       "kernel the primitives are inlined and interleaved; the executed counts per element are in",
       "`r02_ncu_source_strided_lde_summary_final2.json`.", "",
       "| primitive | instructions | IMAD.WIDE(.X) | other IMAD | IADD3(.X) | SEL / MOV / HFMA2 | other |", "|---|---|---|---|---|---|---|"]
listings = []
for fn in sorted(funcs, key=lambda f: f.split("\n")[0]):
    name = fn.split("\n")[0].strip()
    lines = []
    for l in fn.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines.append(m.group(2).strip())
    last_ld = max(i for i, l in enumerate(lines) if re.search(r"\bLDG", l))
    first_st = min(i for i, l in enumerate(lines) if re.search(r"\bSTG", l))
    body = [l for l in lines[last_ld + 1:first_st]]
    mn = collections.Counter()
    for l in body:
        t = l.split()
        op = t[1] if t[0].startswith("@") else t[0]
        mn[op] += 1
    wide = sum(v for k, v in mn.items() if k.startswith("IMAD.WIDE"))
    imad = sum(v for k, v in mn.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
    iadd = sum(v for k, v in mn.items() if k.startswith("IADD3"))
    mov = sum(v for k, v in mn.items() if k.startswith(("SEL", "MOV", "HFMA2")))
    other = len(body) - wide - imad - iadd - mov
    doc.append(f"| `{name}` | {len(body)} | {wide} | {imad} | {iadd} | {mov} | {other} |")
    listings += ["", f"## {name}", "", "```", *body, "```"]
OUT.write_text("\n".join(doc + listings) + "\n")
print("\n".join(doc[8:]))
