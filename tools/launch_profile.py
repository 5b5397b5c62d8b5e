"""profiles/ summary of an ncu launch list with time + DRAM bytes (see profiles/r2_launches_bf16_b64.txt).

    python tools/launch_profile.py gpurun_out/launches.csv profiles/rN_launches_bf16_b64 "title" [precision batch traffic.json]
writes <out>.txt, copies the csv to <out>.csv; with the optional arguments it also records the per-launch DRAM traffic of the
conv and Activation1d kernels under traffic.json[precision] (what bench.py reports as roofline.traffic)."""
import collections
import csv
import json
import shutil
import sys

src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
prec, batch, tjson = (sys.argv[4], sys.argv[5], sys.argv[6]) if len(sys.argv) > 6 else ("bf16", "1", None)
rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", ""), "grid": r["Grid Size"]})
    v, u, m = float(r["Metric Value"]), r["Metric Unit"], r["Metric Name"]
    if m.startswith("dram"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    elif m.startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
    d[m] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in byid.values():
    a = agg[d["name"]]
    a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d["dram__bytes_read.sum"]; a[3] += d["dram__bytes_write.sum"]
tot = sum(a[1] for a in agg.values())
with open(out + ".txt", "w") as f:
    f.write(title + "\n")
    f.write("command: ncu --nvtx --nvtx-include alcm_decode/ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
            f"--clock-control none python tools/one_decode.py {prec} {batch}\n")
    f.write("(cold-cache, serialised launches: compare SHARES with bench.py's class_ms, not absolutes; raw list: "
            + out.split("/")[-1] + ".csv)\n\n")
    f.write(f"{len(byid)} launches, {tot:.1f} us total\n")
    f.write(f"{'us':>10} {'launches':>8} {'share':>6} {'dram rd MB':>10} {'dram wr MB':>10}  kernel\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{a[1]:10.1f} {a[0]:8d} {100 * a[1] / tot:5.1f}% {a[2] / 1e6:10.1f} {a[3] / 1e6:10.1f}  {k}\n")
    for pat, ttl in (("conv_umma", "conv_umma_kernel by grid (1-D: tiles, or resident slots for persistent launches)"),
                     ("act1d", "act1d kernels by grid"), ("gn_fused", "gn_fused_kernel by grid")):
        g = collections.defaultdict(lambda: [0, 0.0, 0.0])
        for d in byid.values():
            if pat in d["name"]:
                a = g[d["grid"]]
                a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        f.write(f"\n{ttl}:\n")
        for k, a in sorted(g.items(), key=lambda kv: -kv[1][1]):
            f.write(f"   {k:>16} x{a[0]:3d} {a[1]:8.1f} us ({a[1] / a[0]:6.1f} us each, dram {a[2] / a[0] / 1e6:6.1f} MB each)\n")
if len(byid) <= 400:
    shutil.copy(src, out + ".csv")
cv = [d for d in byid.values() if "conv_umma" in d["name"]]
print(json.dumps({"kernel": "conv_umma_kernel", "launches": len(cv),
                  "dram_bytes_per_launch": round(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in cv) / len(cv)),
                  "source": out + ".txt (ncu dram__bytes_read.sum + dram__bytes_write.sum, all conv_umma launches of one batch-1 decode, cold cache)"},
                 indent=1))

if tjson:
    import os
    ac = [d for d in byid.values() if "act1d" in d["name"]]
    rec = {"workload": f"one decode of {batch} x 10 s clips, {prec} (tools/one_decode.py {prec} {batch})",
           "conv_launches": len(cv),
           "conv_dram_bytes_per_launch": round(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in cv) / max(len(cv), 1)),
           "conv_dram_bytes_total": round(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in cv)),
           "act_launches": len(ac),
           "act_dram_bytes_per_launch": round(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in ac) / max(len(ac), 1)),
           "act_dram_bytes_total": round(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in ac)),
           "source": out + ".txt (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, --clock-control none)"}
    allj = json.load(open(tjson)) if os.path.exists(tjson) else {}
    allj[prec] = rec
    json.dump(allj, open(tjson, "w"), indent=1)
