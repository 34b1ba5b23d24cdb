"""Summarises ncu output into profiles/ (tracked):

    python tools/ncu_summary.py full   gpurun_out/X.ncu-rep  profiles/NAME      # -> NAME.md + roofline_traffic.json
    python tools/ncu_summary.py list   gpurun_out/X_launches.csv profiles/NAME  # -> NAME.md (per-kernel share of a step)
    python tools/ncu_summary.py traffic gpurun_out/X_traffic.csv profiles/NAME  # -> NAME.csv copy + roofline_traffic.json
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_cycles_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]


def short(name: str) -> str:
    m = re.search(r"(\w+)(<[^>]*>)?\(", name.replace("ezk::<unnamed>::", "").replace("unnamed>::", ""))
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def to_bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1)


def to_ms(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6, "second": 1e3}.get(unit, 1)


def full(rep: str, out: str):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["| kernel | grid | ms | DRAM read GB | DRAM write GB | achieved GB/s | DRAM % | SM % | issue % | ALU % | FMA cyc % | FMA-heavy cyc % | occ % | regs | L2 hit % | top stalls (warps per issue) |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                  and "selected" not in h]
    traffic = collections.defaultdict(list)
    for r in rows[2:]:
        g = lambda k: r[idx[k]] if k in idx else ""
        name = short(g("Kernel Name"))
        ms = to_ms(g("gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]])
        rd = to_bytes(g("dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(g("dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]])
        traffic[name.split("<")[0]].append(rd + wr)
        f = lambda k: f"{float(g(k).replace(',', '')):.1f}" if g(k) not in ("", "n/a") else "-"
        stalls = sorted(((float(g(k).replace(",", "")) if g(k) not in ("", "n/a") else 0.0,
                          k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in stall_keys), reverse=True)[:3]
        stall_txt = ", ".join(f"{n} {v:.2f}" for v, n in stalls)
        lines.append(f"| {name} | {g('Grid Size')} | {ms:.3f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | {(rd + wr) / ms / 1e6:.0f} | "
                     f"{f('dram__throughput.avg.pct_of_peak_sustained_elapsed')} | {f('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                     f"{f('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {f('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')} | "
                     f"{f('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')} | "
                     f"{f('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed')} | {f('sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                     f"{g('launch__registers_per_thread')} | {f('lts__t_sector_hit_rate.pct')} | {stall_txt} |")
    Path(out + ".md").write_text(f"# ncu --set full --clock-control none: {Path(rep).name}\n\n"
                                 "Per-launch values (cold cache, serialised under the profiler; never a bench number).\n\n" + "\n".join(lines) + "\n")
    kmap = {"ntt_strided_pass": "ntt_strided_pass", "ntt_final_pass": "ntt_final_pass", "constraint_kernel": "constraints",
            "hash_rows_kernel": "hash_rows"}
    tf = ROOT / "profiles" / "roofline_traffic.json"
    cur = json.loads(tf.read_text()) if tf.exists() else {}
    for k, v in traffic.items():
        cur[kmap.get(k, k)] = sum(v) / len(v)
    cur["_source"] = f"{Path(rep).name}: mean dram__bytes_read.sum + dram__bytes_write.sum per captured launch"
    tf.write_text(json.dumps(cur, indent=1) + "\n")
    print("\n".join(lines))


def launch_list(path: str, out: str):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = short(r[idx["Kernel Name"]])
        ms = to_ms(r[idx["Metric Value"]], r[idx["Metric Unit"]])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list: {Path(path).name}", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` over `bench.py --steps 2 --warmup 1`: the first 400 "
             "launches (about four proofs).  Times are cold-cache and serialised: compare the SHARES with bench.py's `kernels`, not the absolutes.",
             "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {name} | {cnt} | {ms:.3f} | {100 * ms / total:.1f}% |")
    Path(out + ".md").write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


KMAP = {"ntt_strided_pass": "ntt_strided_pass", "ntt_final_pass": "ntt_final_pass", "constraint_kernel": "constraints",
        "hash_rows_kernel": "hash_rows", "merkle_level_kernel": "merkle_level", "merkle_subtree_kernel": "merkle_level", "pair_inverse_kernel": "pair_inverse",
        "eval_partial_kernel": "eval_polys", "fri_fold_kernel": "fri_fold", "deep_combine_kernel": "deep_combine",
        "deep_pointwise_kernel": "deep_pointwise"}


def traffic(path: str, out: str):
    """`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` over every launch of one proof: mean DRAM bytes
    per launch per kernel -> profiles/roofline_traffic.json (what bench.py prints as roofline.traffic)."""
    text = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(text))
    idx = {h: i for i, h in enumerate(rows[0])}
    per_launch = collections.OrderedDict()
    for r in rows[1:]:
        if r[idx["Metric Name"]] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        key = (r[idx["ID"]], short(r[idx["Kernel Name"]]).split("<")[0])
        per_launch[key] = per_launch.get(key, 0.0) + to_bytes(r[idx["Metric Value"]], r[idx["Metric Unit"]])
    agg = collections.defaultdict(list)
    for (_, name), b in per_launch.items():
        agg[name].append(b)
    Path(out + ".csv").write_text("".join(text))
    res = {KMAP.get(k, k): sum(v) / len(v) for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))}
    res["_source"] = (f"{Path(out).name}.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one "
                      "2^20 proof (tools/profile_prove.py 20); mean per launch")
    (ROOT / "profiles" / "roofline_traffic.json").write_text(json.dumps(res, indent=1) + "\n")
    for k, v in res.items():
        print(k, v)


if __name__ == "__main__":
    {"full": full, "list": launch_list, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
