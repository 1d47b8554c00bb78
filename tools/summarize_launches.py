"""Turns an `ncu --csv --metrics ...` launch list into the per-kernel JSON summary kept under profiles/.
usage: python tools/summarize_launches.py <launches.csv> <out.json> [config-name]
Keeps the launches of the LAST frame in the file (the list usually holds a warm-up frame first)."""
import csv, json, sys, collections

src, dst = sys.argv[1], sys.argv[2]
cfg = sys.argv[3] if len(sys.argv) > 3 else "c3"
rows = []
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
by_id = collections.OrderedDict()
for r in rows:
    k = by_id.setdefault(r["ID"], {"kernel": r["Kernel Name"].split("(")[0].replace("void ", "").replace("b200sgm::", ""), "grid": r["Grid Size"], "block": r["Block Size"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    name, unit = r["Metric Name"], r["Metric Unit"]
    if name == "gpu__time_duration.sum":
        k["time_us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    elif name == "dram__bytes_read.sum":
        k["dram_read_bytes"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif name == "dram__bytes_write.sum":
        k["dram_write_bytes"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif name == "smsp__inst_executed.sum":
        k["warp_inst"] = v
    elif name == "smsp__issue_active.avg.pct_of_peak_sustained_active":
        k["issue_active_pct"] = v
    elif name == "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active":
        k["pipe_alu_pct"] = v
ks = list(by_id.values())
# last frame = from the last k_prefilter launch onward (one launch filters both images)
idx = [i for i, k in enumerate(ks) if k["kernel"].startswith("k_prefilter")]
start = idx[-1] if idx else 0
ks = ks[start:]
tot_t = sum(k.get("time_us", 0) for k in ks)
for k in ks:
    k["share_of_frame_pct"] = round(100 * k.get("time_us", 0) / tot_t, 2) if tot_t else None
    by = k.get("dram_read_bytes", 0) + k.get("dram_write_bytes", 0)
    if k.get("time_us"):
        k["dram_gbs"] = round(by / (k["time_us"] * 1e-6) / 1e9, 1)
out = {"config": cfg, "note": "ncu --metrics pass (cold-cache, serialised launches); last frame of the run",
       "frame_time_us": tot_t, "frame_dram_bytes": sum(k.get("dram_read_bytes", 0) + k.get("dram_write_bytes", 0) for k in ks),
       "kernels": ks}
json.dump(out, open(dst, "w"), indent=1)
print("frame %.1f us, %.2f GB DRAM, %d launches" % (tot_t, out["frame_dram_bytes"] / 1e9, len(ks)))
for k in ks:
    print("  %-28s %9.1f us %5.1f%%  %7.1f GB/s  inst %.0fM" % (k["kernel"][:28], k.get("time_us", 0), k["share_of_frame_pct"] or 0, k.get("dram_gbs", 0), k.get("warp_inst", 0) / 1e6))
