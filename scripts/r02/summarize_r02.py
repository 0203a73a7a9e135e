"""gpurun_out/ncu_<tag>_*.csv (ncu --page raw --csv exports) + launches_<tag>.csv -> committed summaries under profiles/.

    python scripts/r02/summarize_r02.py            # tag r02  (scripts/r02/ncu_r02.sh)
    TAG=r02b python scripts/r02/summarize_r02.py   # tag r02b (scripts/r02/ncu_r02b.sh: after the wide-stage sweep / resident knn_data filter)
"""
import collections
import csv
import json
import os
import re

TAG = os.environ.get("TAG", "r02")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
KEEP = ("gpu__time_duration", "sm__cycles_elapsed.avg.per_second", "dram__bytes", "dram__throughput", "lts__t_sector_hit_rate", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes", "pipe_tensor", "sm__issue_active", "pipe_fma_cycles", "pipe_fp64", "launch__", "sm__throughput",
        "smsp__inst_executed.sum", "sm__warps_active", "smsp__average_warp", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.sum",
        "gpu__compute_memory_throughput", "l1tex__t_bytes", "smsp__cycles_active")

lp = os.path.join(G, f"launches_{TAG}.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    open(os.path.join(OUT, f"launches_{TAG}.csv"), "w").writelines(lines)
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("mdsctk::", "")
        a = agg.setdefault(name, {"launches": 0, "total_ms": 0.0})
        a["launches"] += 1
        a["total_ms"] += float(r["Metric Value"]) / 1e6
    tot = sum(a["total_ms"] for a in agg.values())
    for a in agg.values():
        a["share"] = a["total_ms"] / tot
    json.dump(agg, open(os.path.join(OUT, f"launches_{TAG}_summary.json"), "w"), indent=1)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["total_ms"]):
        print(f"{k:40s} {a['launches']:5d} {a['total_ms']:10.2f} ms {100 * a['share']:6.2f} %")

traffic = {}
for f in sorted(os.listdir(G)):
    m = re.match(r"ncu_" + TAG + r"_(\w+)\.csv$", f)
    if not m:
        continue
    rows = list(csv.reader(open(os.path.join(G, f))))
    if len(rows) < 3:
        print("empty", f)
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    met = {h: {"value": v, "unit": u} for h, u, v in zip(hdr, units, vals) if any(k in h for k in KEEP) and v != ""}
    name = next((v for h, v in zip(hdr, vals) if h == "Kernel Name"), "")

    def val(key):
        e = met.get(key)
        return None if e is None else float(e["value"].replace(",", "")) * UNIT.get(e["unit"], 1.0)
    dur, rd, wr = val("gpu__time_duration.sum"), val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    summ = {"kernel": name, "capture": m.group(1), "duration_ms": dur * 1e3 if dur else None,
            "dram_bytes_per_launch": (rd or 0) + (wr or 0), "dram_GBps": ((rd or 0) + (wr or 0)) / dur / 1e9 if dur else None,
            "sm_clock_GHz": float(met["sm__cycles_elapsed.avg.per_second"]["value"]) if "sm__cycles_elapsed.avg.per_second" in met else None,
            "tensor_pipe_active_pct_of_elapsed": float(met["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]["value"])
            if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in met else None,
            "issue_active_pct": float(met["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]["value"])
            if "sm__issue_active.avg.pct_of_peak_sustained_elapsed" in met else None,
            "l2_to_sm_bytes": val("l1tex__m_xbar2l1tex_read_bytes.sum")}
    json.dump({"summary": summ, "metrics": met}, open(os.path.join(OUT, f"ncu_{m.group(1)}_{TAG}.json"), "w"), indent=1)
    traffic[m.group(1)] = summ["dram_bytes_per_launch"]
    print(json.dumps(summ))
if traffic:
    t = {"c4": traffic.get("sweep_c4"), "c3": traffic.get("sweep_c3"), "c3_single_basin": traffic.get("sweep_single_basin"), "c5": traffic.get("data_sweep_c5"),
         "unit": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (ncu --set full): a 131072-row block for c4 / c5, the whole "
                 "100k-row query for c3"}
    old = os.path.join(OUT, "traffic_r02.json")
    if TAG != "r02" and os.path.exists(old):           # captures not repeated keep the earlier value
        o = json.load(open(old))
        for k in ("c4", "c3", "c3_single_basin", "c5"):
            if t.get(k) is None:
                t[k] = o.get(k)
    json.dump(t, open(os.path.join(OUT, f"traffic_{TAG}.json"), "w"), indent=1)
