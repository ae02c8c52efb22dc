#!/usr/bin/env python
"""Summarise an `ncu --set full` report (here, no GPU needed):
  ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/raw.csv
  python tools/ncu_summary.py /tmp/raw.csv profiles/r02_ncu_full_cfg4_5M.md [profiles/r02_traffic.json WORKLOAD]
Writes a per-kernel table (time, DRAM bytes, DRAM %, issue-active %, warp instructions, registers,
occupancy limiters) and, optionally, the per-launch traffic file bench.py reads for roofline.traffic."""
import csv
import json
import sys

raw, out_md = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
M = {"ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "inst": "smsp__inst_executed.sum",
     "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
     "lim_regs": "launch__occupancy_limit_registers", "lim_smem": "launch__occupancy_limit_shared_mem",
     "l2_hit": "lts__t_sector_hit_rate.pct", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed"}


def val(r, key):
    i = col[M[key]]
    v = float(r[i].replace(",", ""))
    u = units[i]
    if key in ("rd", "wr"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if key == "ms":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
    return v


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "")
    return n.split("(")[0]


lines = ["| # | kernel | grid | ms | DRAM read MB | DRAM write MB | DRAM % of peak | issue active % | warp instr (M) | warps active % | regs | CTAs/SM limit (regs / smem) | L2 hit % |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
agg = {}
for k, r in enumerate(data):
    name = short(r[col["Kernel Name"]])
    d = {key: val(r, key) for key in M}
    lines.append(f"| {k} | `{name}` | {r[col['Grid Size']]} | {d['ms']:.3f} | {d['rd'] / 1e6:.0f} | {d['wr'] / 1e6:.0f} | "
                 f"{d['dram_pct']:.1f} | {d['issue_pct']:.1f} | {d['inst'] / 1e6:.0f} | {d['warps_pct']:.0f} | {d['regs']:.0f} | "
                 f"{d['lim_regs']:.0f} / {d['lim_smem']:.0f} | {d['l2_hit']:.0f} |")
    a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "warp_instructions": 0.0})
    a["launches"] += 1
    a["ms"] += d["ms"]
    a["dram_bytes"] += d["rd"] + d["wr"]
    a["warp_instructions"] += d["inst"]
open(out_md, "a").write("\n".join(lines) + "\n")
print("\n".join(lines))
if len(sys.argv) > 4:
    kern = {}
    for name, a in agg.items():
        key = {"blend_fwd_kernel": "blend_fwd", "blend_bwd_kernel": "blend_bwd"}.get(name, name)
        n = a["launches"]
        kern[key] = {"dram_bytes": a["dram_bytes"] / n, "warp_instructions": a["warp_instructions"] / n,
                     "ms_under_ncu": a["ms"] / n, "launches_captured": n}
    json.dump({"workload": sys.argv[4], "source": raw, "note": "per-launch averages from one ncu --set full capture "
               "(cold-cache, serialised: use shares, not absolutes)", "kernels": kern}, open(sys.argv[3], "w"), indent=1)
