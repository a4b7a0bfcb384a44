#!/usr/bin/env python
"""Condense gpurun_out/prof_<tag>_*_raw.csv (ncu --set full, one kernel each) and launches_<tag>.csv
(gpu__time_duration per launch) into small tracked files under profiles/."""
import collections, csv, glob, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_tag = sys.argv[2] if len(sys.argv) > 2 else tag
KEYS = [
    ("duration_us", "gpu__time_duration.sum"), ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("warp_inst", "smsp__inst_executed.sum"), ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"),
]
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
rows = []
paths = [p for p in glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{tag}_*_raw.csv")) if re.findall(r"_(\d+)_raw", p)]
for path in sorted(paths, key=lambda p: int(re.findall(r"_(\d+)_raw", p)[0])):
    r = list(csv.reader(open(path)))
    if len(r) < 3:
        continue
    hdr, units, d = r[0], r[1], r[2]
    idx = {h: i for i, h in enumerate(hdr)}
    row = {"kernel": re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void <unnamed>::", "").replace("<unnamed>::", "")}
    for k, m in KEYS:
        if m in idx and d[idx[m]] not in ("", "n/a"):
            v = float(d[idx[m]].replace(",", ""))
            v *= UNIT.get(units[idx[m]], 1.0) if k in ("duration_us", "dram_read_MB", "dram_write_MB") else 1.0
            row[k] = round(v, 3)
    stalls = [(h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(d[i].replace(",", "")))
              for h, i in idx.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[i] not in ("", "n/a")]
    row["top_stalls"] = " ".join(f"{a}={b:.1f}" for a, b in sorted(stalls, key=lambda x: -x[1])[:4])
    rows.append(row)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
if rows:
    cols = ["kernel"] + [k for k, _ in KEYS] + ["top_stalls"]
    with open(os.path.join(ROOT, "profiles", f"{out_tag}_ncu_full_kernels.csv"), "w") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for row in rows:
            w.writerow(row)
            print({k: row.get(k) for k in ("kernel", "duration_us", "dram_pct", "issue_pct", "warps_active_pct", "regs", "top_stalls")})
lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void <unnamed>::", "").replace("<unnamed>::", "")[:90]
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0) / 1e3
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{out_tag}_launches_by_kernel.csv"), "w") as f:
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{a[0]},{a[1]:.4f},{100 * a[1] / tot:.2f}\n")
    print("launch list total ms", round(tot, 2))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"  {a[1]:8.3f} ms {a[0]:5d}x {100*a[1]/tot:5.1f}%  {k}")
