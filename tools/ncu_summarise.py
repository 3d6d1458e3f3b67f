#!/usr/bin/env python3
"""profiles/<tag>_final_launches.csv (ncu --csv --metrics ..., one bench step) -> profiles/<tag>_final_launches.md and
profiles/dram_traffic.json (per-kernel DRAM bytes per step + the fingerprint of the kernel sources they were measured on; read by
bench.py for roofline.traffic, which prints them only while the sources are unchanged)."""
import csv, json, os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_final_launches.csv")
tag = os.path.basename(src).split("_")[0]
sys.path.insert(0, ROOT)
from bench import kernel_source_sha
rows = list(csv.reader(open(src)))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ix = {h: i for i, h in enumerate(hdr)}
d = OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].replace("srsb200::", "").replace("void ", "").split("(")[0]
    d.setdefault((int(r[ix["ID"]]), name), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
lines = ["# " + tag + " final: per-launch ncu metrics of one bench step (16384 CB x K=6144, max 8 half-iterations, early stop)", "",
         "Command (B200, after the same command exited 0 without ncu): `tools/ncu_final_launches.sh` (`ncu --metrics gpu__time_duration.sum,"
         "dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active...,sm__warps_active...,sm__pipe_alu_cycles_active...,"
         "sm__pipe_fma_cycles_active... -k regex:\"scan_kernel|job_kernel|extract_kernel|emit_kernel\" --clock-control none -s 54 -c 18 --csv "
         "python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu`). Raw CSV: `profiles/" + tag + "_final_launches.csv`. Launches are serialised and "
         "cold-cache under ncu: compare shares.", "",
         "| # | kernel | us | DRAM read MB | DRAM write MB | achieved DRAM TB/s | warp-instr M | issue active % | ALU pipe % | FMA pipe % |",
         "|---|---|---|---|---|---|---|---|---|---|"]
per_b, per_us = OrderedDict(), OrderedDict()
tot_us = 0.0
for n, ((_, name), v) in enumerate(d.items()):
    us = v["gpu__time_duration.sum"] / 1e3
    rd, wr = v["dram__bytes_read.sum"], v["dram__bytes_write.sum"]
    base = name.split("<")[0]
    per_b[base] = per_b.get(base, 0.0) + rd + wr
    per_us[base] = per_us.get(base, 0.0) + us
    tot_us += us
    lines.append("| %d | %s | %.1f | %.0f | %.0f | %.2f | %.1f | %.1f | %.1f | %.1f |" % (
        n, name, us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e6, v["smsp__inst_executed.sum"] / 1e6,
        v["smsp__issue_active.avg.pct_of_peak_sustained_active"], v["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"],
        v["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]))
lines += ["", "Per kernel per step: " + ", ".join("%s %.0f us / %.2f GB" % (k, per_us[k], per_b[k] / 1e9) for k in per_b) +
          "; total %.0f us, %.2f GB of DRAM traffic (= %.2f ms at the measured 6.54 TB/s copy peak)." % (tot_us, sum(per_b.values()) / 1e9, sum(per_b.values()) / 6542.1e9 * 1e3)]
open(os.path.join(ROOT, "profiles", tag + "_final_launches.md"), "w").write("\n".join(lines) + "\n")
json.dump({"source": "profiles/" + tag + "_final_launches.csv (ncu, one bench step, 16384 CB x K=6144, serialised launches)", "kernel_source_sha": kernel_source_sha(),
           "job_kernel_bytes_per_step": per_b.get("job_kernel"), "per_kernel_bytes_per_step": per_b, "per_kernel_us_per_step_under_ncu": per_us,
           "total_bytes_per_step": sum(per_b.values())}, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
print("\n".join(lines[-1:]))
