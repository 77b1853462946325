"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_list_summary.py in.csv out.json "command" """
import collections, csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[d["Metric Unit"]]
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += ms
total = sum(v[1] for v in agg.values())
out = {"command": sys.argv[3] if len(sys.argv) > 3 else "", "note": "per-launch times under ncu are serialised and cold-cache: the SHARES are what compares with bench.py's roofline.share_of_step",
       "total_ms": total, "launches": sum(v[0] for v in agg.values()),
       "kernels": [{"kernel": k, "launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / total, 5)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["kernels"][:6]))
