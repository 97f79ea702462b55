"""Folds the ncu CSVs of tools/trace_counters.sh into trace_counters.json: warp instructions per ray and lanes per
instruction of the traversal kernels (all k_trace* launches of 3 frames / the rays those frames traced)."""
import csv, json, os, sys

out = sys.argv[1]
res = {}
for w in sys.argv[2:]:
    log = [l for l in open(os.path.join(out, f"tc_{w}.log")) if l.startswith("{")]
    if not log:
        continue
    info = json.loads(log[-1])
    rows = list(csv.reader(l for l in open(os.path.join(out, f"tc_{w}.csv")) if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        kind = "trace" if "k_trace" in r[ik] else "shade"
        agg.setdefault(kind, {}).setdefault(r[im], 0.0)
        agg[kind][r[im]] += float(r[iv].replace(",", ""))
    t, s = agg["trace"], agg.get("shade", {})
    res[w] = {"warp_inst_per_ray": t["smsp__inst_executed.sum"] / info["rays"],
              "thread_inst_per_warp_inst": t["smsp__thread_inst_executed.sum"] / t["smsp__inst_executed.sum"],
              "thread_inst_per_ray": t["smsp__thread_inst_executed.sum"] / info["rays"],
              "trace_us_per_frame_under_ncu": t["gpu__time_duration.sum"] / 1e3 / info["frames"],
              "shade_warp_inst_per_ray": s.get("smsp__inst_executed.sum", 0.0) / info["rays"],
              "rays": info["rays"], "frames": info["frames"],
              "source": f"ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none, every k_trace* launch of {info['frames']} frames (PTC_PARTS=1), tools/trace_counters.sh"}
json.dump(res, open(os.path.join(out, "trace_counters.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
