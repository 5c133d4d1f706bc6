#!/usr/bin/env python
"""Turn an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,...`
launch list into the per-kernel summary kept under profiles/ (shares of a step, DRAM bytes) and the
profiles/dominant_kernel_traffic.json that bench.py reads for `roofline.traffic`.

    python tools/summarize_launches.py gpurun_out/launches.csv --first K --count N --out profiles/NAME.csv \
        [--traffic-json profiles/dominant_kernel_traffic.json --dominant gemm_h_kernel,resblock]
        [--manifest gpurun_out/manifest.json --workload music256 --class-json profiles/r2_traffic_by_class.json]

`--manifest` is the launch-by-launch category list bench.py writes when HILCODEC_DUMP_LAUNCHES=<path> is set (the
library's own record of the step, `hil_profile_launches`); the ncu launch list of the same step has the same launches in
the same order, so zipping the two attributes the measured DRAM bytes to the layer classes bench.py reports rooflines
for (`roofline.traffic`).
"""
import argparse
import collections
import csv
import json
import re


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    byid = collections.OrderedDict()
    for x in csv.DictReader(lines):
        d = byid.setdefault(int(x["ID"]), {"name": re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").strip(),
                                           "grid": x["Grid Size"]})
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        if x["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        elif u.lower().startswith("kbyte"):
            v *= 1e3
        elif u.lower().startswith("mbyte"):
            v *= 1e6
        elif u.lower().startswith("gbyte"):
            v *= 1e9
        d[x["Metric Name"]] = v
    return list(byid.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--first", type=int, default=0, help="index (in the capture) of the first launch of the step")
    ap.add_argument("--count", type=int, default=0, help="launches per step (0 = all)")
    ap.add_argument("--out")
    ap.add_argument("--title", default="")
    ap.add_argument("--traffic-json")
    ap.add_argument("--dominant", default="gemm_h_kernel,resblock_kernel")
    ap.add_argument("--manifest")
    ap.add_argument("--workload", default="music256")
    ap.add_argument("--class-json")
    ap.add_argument("--last-step", action="store_true",
                    help="take the LAST --count (or manifest length) launches of this library: torch's own kernels "
                         "(at::, cub::, ...) are dropped first, so the window is the profiled step bench.py ran last")
    a = ap.parse_args()
    rows = load(a.csv)
    if a.last_step:
        rows = [r for r in rows if not re.search(r"\bat::|at_cuda_detail|cub::|elementwise_kernel|vectorized_|distribution_|Memset|memset", r["name"])]
        n = a.count or (len(json.load(open(a.manifest))["launches"]) if a.manifest else 0)
        rows = rows[-n:] if n else rows
    else:
        rows = rows[a.first:a.first + a.count] if a.count else rows[a.first:]
    agg = collections.OrderedDict()
    for r in rows:
        g = agg.setdefault(r["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tensor": 0.0})
        g["n"] += 1
        g["us"] += r.get("gpu__time_duration.sum", 0.0)
        g["rd"] += r.get("dram__bytes_read.sum", 0.0)
        g["wr"] += r.get("dram__bytes_write.sum", 0.0)
        g["tensor"] += r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * \
            r.get("gpu__time_duration.sum", 0.0)
    total = sum(g["us"] for g in agg.values()) or 1.0
    out = []
    if a.title:
        out.append("# " + a.title)
    out.append("# per-launch times are cold-cache and serialised (ncu replay): compare SHARES, not absolutes")
    out.append("kernel,launches,total_us,share,dram_read_MB,dram_write_MB,dram_GBps,tensor_pipe_active_pct")
    for name, g in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.append(f"{name},{g['n']},{g['us']:.1f},{g['us'] / total:.4f},{g['rd'] / 1e6:.0f},{g['wr'] / 1e6:.0f},"
                   f"{(g['rd'] + g['wr']) / g['us'] / 1e3 if g['us'] else 0:.0f},{g['tensor'] / g['us'] if g['us'] else 0:.1f}")
    out.append(f"# total {total:.1f} us over {len(rows)} launches")
    text = "\n".join(out) + "\n"
    if a.out:
        open(a.out, "w").write(text)
    print(text)
    if a.traffic_json:
        keys = [k for k in a.dominant.split(",") if k]
        dom = [r for r in rows if any(k in r["name"] for k in keys)]
        by = sum(r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0) for r in dom)
        json.dump({"kernel": "pointwise / fused-DWS / fused-ResBlock tensor-core GEMM launches of one step (" + a.dominant + ")",
                   "launches": len(dom), "source": (a.out or a.csv) + " (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                   "dram_bytes_per_launch_avg": by / max(len(dom), 1), "dram_bytes_per_step": by},
                  open(a.traffic_json, "w"), indent=1)
    if a.manifest and a.class_json:
        by_class(rows, a.manifest, a.workload, a.class_json, (a.out or a.csv) + " (ncu dram__bytes_read.sum + dram__bytes_write.sum)")


def by_class(rows, manifest_path, workload, out_path, source):
    man = json.load(open(manifest_path))["launches"]
    # ncu sees every kernel (also cudaMemset2D etc. are not kernels: same count expected); align by count
    if len(man) != len(rows):
        raise SystemExit(f"manifest has {len(man)} launches, the ncu window {len(rows)}: adjust --first / --count")
    agg = collections.OrderedDict()
    for m, r in zip(man, rows):
        g = agg.setdefault(m["cat"], {"launches": 0, "dram_bytes_per_step": 0.0, "us": 0.0, "algorithmic_bytes_per_step": 0.0,
                                      "kernels": set()})
        g["launches"] += 1
        g["dram_bytes_per_step"] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
        g["us"] += r.get("gpu__time_duration.sum", 0.0)
        g["algorithmic_bytes_per_step"] += m["bytes"]
        g["kernels"].add(r["name"])
    try:
        doc = json.load(open(out_path))
    except Exception:
        doc = {}
    doc[workload] = {k: {"launches": g["launches"], "dram_bytes_per_step": g["dram_bytes_per_step"],
                         "dram_bytes_per_launch_avg": g["dram_bytes_per_step"] / g["launches"],
                         "algorithmic_bytes_per_step": g["algorithmic_bytes_per_step"],
                         "ncu_us_per_step": g["us"], "kernels": sorted(g["kernels"]), "source": source}
                     for k, g in agg.items()}
    json.dump(doc, open(out_path, "w"), indent=1)
    for k, v in doc[workload].items():
        print(f"{k:24s} {v['launches']:3d} launches  {v['dram_bytes_per_step'] / 1e9:8.2f} GB DRAM  "
              f"{v['algorithmic_bytes_per_step'] / 1e9:8.2f} GB algorithmic  {v['ncu_us_per_step'] / 1e3:7.2f} ms")


if __name__ == "__main__":
    main()
