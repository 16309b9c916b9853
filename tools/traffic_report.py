"""Aggregate an ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`) of
tools/profile_forward.py: per-kernel DRAM bytes and time of the LAST forward, and the totals of the tensor-core
convolution kernels (what bench.py reports as roofline.traffic).  With a 4th argument also writes the per-kernel table of
every NON-conv kernel (post-processing, refine_mask, thin layers): launches, time, DRAM bytes and achieved GB/s per batch.
Usage: traffic_report.py CSV [n_forwards] [conv_out.json] [postproc_out.json]"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    path = sys.argv[1]
    nfwd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per_id = collections.OrderedDict()
    for r in rd:
        if len(r) < len(hdr):
            continue
        kid = int(r[ix["ID"]])
        d = per_id.setdefault(kid, {"name": re.sub(r"\(.*", "", r[ix["Kernel Name"]])})
        val = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
        d[r[ix["Metric Name"]]] = val
    launches = list(per_id.values())
    n = len(launches) // nfwd
    last = launches[-n:]
    agg = collections.OrderedDict()
    for d in last:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    print("%-48s %4s %10s %10s %10s" % ("kernel", "n", "us", "rd MB", "wr MB"))
    for k, (c, t, rdb, wrb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %4d %10.1f %10.1f %10.1f" % (k[:48], c, t, rdb / 1e6, wrb / 1e6))
    conv = [d for d in last if any(k in d["name"] for k in ("conv_tc", "conv_halo", "conv_hs", "conv_sw", "conv_bneck", "conv_segtail"))]
    tot = {"launches": len(conv), "us": sum(d.get("gpu__time_duration.sum", 0.0) for d in conv),
           "dram_read_bytes": sum(d.get("dram__bytes_read.sum", 0.0) for d in conv),
           "dram_write_bytes": sum(d.get("dram__bytes_write.sum", 0.0) for d in conv)}
    tot["dram_bytes"] = tot["dram_read_bytes"] + tot["dram_write_bytes"]
    print("tensor-core conv kernels: %d launches, %.1f us, DRAM %.1f MB read + %.1f MB written per forward" % (
        tot["launches"], tot["us"], tot["dram_read_bytes"] / 1e6, tot["dram_write_bytes"] / 1e6))
    if len(sys.argv) > 3 and sys.argv[3] != "-":
        json.dump(tot, open(sys.argv[3], "w"), indent=1)
    if len(sys.argv) > 4:
        rows = []
        for k, (c, t, rdb, wrb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if any(x in k for x in ("conv_tc", "conv_halo", "conv_hs", "conv_sw", "conv_bneck", "conv_segtail")):
                continue
            rows.append({"kernel": k, "launches": c, "us": round(t, 1), "dram_read_MB": round(rdb / 1e6, 2),
                         "dram_write_MB": round(wrb / 1e6, 2),
                         "achieved_GBps": round((rdb + wrb) / (t * 1e-6) / 1e9, 1) if t > 0 else None})
        json.dump({"what": "non-conv kernels of one batch (ncu dram__bytes_* and gpu__time_duration per launch, summed per kernel); "
                           "achieved_GBps = DRAM bytes / kernel time; HBM peak 6572 GB/s (MEASURED_PEAKS.json)",
                   "kernels": rows}, open(sys.argv[4], "w"), indent=1)


if __name__ == "__main__":
    main()
