"""Turns the scratch ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [kernel_id]

  gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum ... --csv python bench.py ...)
  gpurun_out/prof_tc_<tag>.ncu-rep (ncu --set full ... -k regex:rms_sweep_tc)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "3"
out = os.path.join(ROOT, "profiles")

lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    open(os.path.join(out, f"launches_{tag}.csv"), "w").writelines(lines)
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("mdsctk::", "")
        a = agg.setdefault(name, {"launches": 0, "total_ms": 0.0})
        a["launches"] += 1
        a["total_ms"] += float(r["Metric Value"]) / 1e6
    tot = sum(a["total_ms"] for a in agg.values())
    for a in agg.values():
        a["share"] = a["total_ms"] / tot
    json.dump(agg, open(os.path.join(out, f"launches_{tag}_summary.json"), "w"), indent=1)
    print(json.dumps(agg, indent=1))

rp = os.path.join(ROOT, "gpurun_out", f"prof_tc_{tag}.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = ("gpu__time_duration", "sm__cycles_elapsed.avg.per_second", "dram__bytes", "dram__throughput", "lts__t_sector_hit_rate",
            "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes", "pipe_tensor", "mem_tensor", "sm__issue_active", "pipe_fma_cycles",
            "pipe_fmaheavy", "launch__", "sm__throughput", "smsp__inst_executed.sum", "sm__warps_active", "smsp__warp_issue_stalled",
            "smsp__average_warp", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.sum")
    m = {h: {"value": v, "unit": u} for h, u, v in zip(hdr, units, vals) if any(k in h for k in keep) and v != ""}
    name = next((v for h, v in zip(hdr, vals) if h == "Kernel Name"), "")
    json.dump({"kernel": name, "report": os.path.basename(rp), "metrics": m},
              open(os.path.join(out, f"ncu_rms_sweep_tc_{tag}.json"), "w"), indent=1)
    rd = float(m["dram__bytes_read.sum"]["value"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12}[m["dram__bytes_read.sum"]["unit"]]
    wr = float(m["dram__bytes_write.sum"]["value"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12}[m["dram__bytes_write.sum"]["unit"]]
    json.dump({kid: rd + wr, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, C3 workload)"},
              open(os.path.join(out, f"traffic_{tag}.json"), "w"), indent=1)
    print("traffic", rd + wr, "tensor", {k: v for k, v in m.items() if "tensor_cycles_active" in k})
