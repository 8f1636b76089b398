"""ncu CSV of every launch of one training step (tools/one_step.py under `ncu --profile-from-start off --metrics
gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,gpu__dram_throughput...
--csv`) -> profiles/traffic_rN.json (what bench.py fills roofline.traffic from) + a markdown share table.
usage: python tools/summarize_ncu_step.py gpurun_out/r3_launches.csv profiles/traffic_r3.json profiles/r3/launches_r3_summary.md \
       ["round 3" [step_ms]]"""
import collections
import csv
import json
import re
import sys

src, out_json, out_md = sys.argv[1:4]
label = sys.argv[4] if len(sys.argv) > 4 else "round 2"          # e.g. "round 3"
step_ms = sys.argv[5] if len(sys.argv) > 5 else "40.8"           # the overlapped multi-stream step of that round
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
per = {}
for row in csv.DictReader(lines):
    d = per.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})
    try:
        d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        pass


def short(n):
    n = re.sub(r"\(.*", "", re.sub(r"^void ", "", n)).replace("bvae::", "")
    return re.sub(r"\((int|bool)\)", "", n).replace(", ", ",")


T, TP, DP = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
agg = collections.OrderedDict()
for k in sorted(per):
    d = per[k]
    a = agg.setdefault(short(d["name"]), {"launches": 0, "ms": 0.0, "dram": 0.0, "tw": 0.0, "dw": 0.0})
    t = d.get(T, 0.0)
    a["launches"] += 1
    a["ms"] += t / 1e6
    a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["tw"] += d.get(TP, 0.0) * t
    a["dw"] += d.get(DP, 0.0) * t
tot = sum(a["ms"] for a in agg.values())
how = ("ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
       "dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg."
       "pct_of_peak_sustained_elapsed --csv python tools/one_step.py  (one eager 512-bar training step after 3 warm-up steps, "
       "every launch; %s)" % src)
table = {}
for n, a in agg.items():
    table[n] = {"launches_per_step": a["launches"], "ms_per_step_under_ncu": round(a["ms"], 4),
                "share_of_step": round(a["ms"] / tot, 4), "dram_bytes_per_launch": round(a["dram"] / a["launches"]),
                "dram_bytes_per_step": round(a["dram"]),
                "tensor_pipe_pct_time_weighted": round(a["tw"] / (a["ms"] * 1e6 + 1e-9), 1),
                "dram_throughput_pct_time_weighted": round(a["dw"] / (a["ms"] * 1e6 + 1e-9), 1),
                "own_kernel": not n.startswith("at::") and "nccl" not in n.lower() and "cub::" not in n}
nb = [n for n in agg if n.startswith(("nb_", "nbf_", "nbs_"))]
co = [n for n in agg if n.startswith(("conv_tc", "wgrad_tc", "wgrad_halo", "stem_", "wgrad_unpack"))]
table["norm_blocks"] = {"kernels": nb, "launches_per_step": sum(agg[n]["launches"] for n in nb),
                        "ms_per_step_under_ncu": round(sum(agg[n]["ms"] for n in nb), 3),
                        "dram_bytes_per_step": round(sum(agg[n]["dram"] for n in nb)),
                        "algorithmic_bytes_per_step": int(4.26e6 * 10 * 512)}
table["contractions"] = {"kernels": co, "launches_per_step": sum(agg[n]["launches"] for n in co),
                         "ms_per_step_under_ncu": round(sum(agg[n]["ms"] for n in co), 3),
                         "dram_bytes_per_step": round(sum(agg[n]["dram"] for n in co))}
table["_meta"] = {"how": how, "total_ms_serialised": round(tot, 3), "launches": len(per),
                  "note": "per-launch times are cold-cache and serialised: compare shares, not absolutes"}
json.dump(table, open(out_json, "w"), indent=1)
with open(out_md, "w") as f:
    f.write("# ncu launch list of ONE 512-bar training step (%s, final kernels, eager launches)\n\n`%s`\n\n" % (label, how))
    f.write("%d launches, %.2f ms serialised (the overlapped multi-stream step takes %s ms).  Library glue (ATen fills / copies / "
            "cat on [B,1152]-sized tensors) is the `at::` rows.\n\n" % (len(per), tot, step_ms))
    f.write("| kernel instance | launches | ms | share | DRAM MB / launch | tensor pipe % | DRAM % |\n|---|---:|---:|---:|---:|---:|---:|\n")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        t = table[n]
        f.write("| `%s` | %d | %.3f | %.1f %% | %.1f | %.1f | %.1f |\n" % (n[:80], a["launches"], a["ms"], 100 * a["ms"] / tot,
                t["dram_bytes_per_launch"] / 1e6, t["tensor_pipe_pct_time_weighted"], t["dram_throughput_pct_time_weighted"]))
    x = table["norm_blocks"]
    f.write("\nNorm blocks: %d launches, %.2f ms, %.1f GB DRAM traffic per step against %.1f GB algorithmic (SURVEY.md 8d) = %.2fx.\n"
            % (x["launches_per_step"], x["ms_per_step_under_ncu"], x["dram_bytes_per_step"] / 1e9,
               x["algorithmic_bytes_per_step"] / 1e9, x["dram_bytes_per_step"] / x["algorithmic_bytes_per_step"]))
    x = table["contractions"]
    f.write("Contractions: %d launches, %.2f ms, %.1f GB DRAM traffic per step.\n" % (x["launches_per_step"],
            x["ms_per_step_under_ncu"], x["dram_bytes_per_step"] / 1e9))
print(json.dumps({k: table[k] for k in ("norm_blocks", "contractions", "_meta")}, indent=1)[:1500])
