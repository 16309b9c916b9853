"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the
LAST forward in the log and, given the op table printed by tools/profile_forward.py, per-layer TFLOP/s."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    idx = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rd:
        if len(r) < len(hdr):
            continue
        if "Metric Name" in idx and r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue   # multi-metric lists (traffic captures): one row per metric and launch
        u = r[idx["Metric Unit"]]
        t = float(r[idx["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
        data.append((r[idx["Kernel Name"]], t, r[idx["Grid Size"]]))
    return data


def main():
    data = load(sys.argv[1])
    nfwd = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    n = len(data) // nfwd
    second = data[-n:]
    tot = sum(t for _, t, _ in second)
    agg = collections.OrderedDict()
    for k, t, g in second:
        name = re.sub(r"\(.*", "", k)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    print("launches per forward: %d, total %.1f us" % (n, tot))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-52s n=%3d %10.1f us %5.1f%%" % (k[:52], c, t, 100 * t / tot))
    if len(sys.argv) > 2:
        ops = [l.split() for l in open(sys.argv[2]) if l.startswith("op ")]
        gemm = [o for o in ops if o[3] in ("0", "1", "2", "6", "7", "10")]
        tc = [(k, t, g) for (k, t, g) in second if "conv_tc" in k or "conv_halo" in k or "conv_hs" in k or "conv_sw" in k or "conv_bneck" in k or "conv_segtail" in k]
        bs = 16
        print()
        tflop = 0.0
        for o, (k, t, g) in zip(gemm, tc):
            kind, kk, s, cin, cout, down = int(o[3]), int(o[5]), int(o[7]), int(o[9]), int(o[11]), int(o[13])
            hw = (1024 // down) ** 2 * bs
            if kind == 10:
                fl = 2 * hw * 10 * cin * cout
            elif kind == 7:
                fl = 2 * hw * 16 * cin
            elif kind == 0:
                fl = 2 * (1024 // 2) ** 2 * bs * 108 * cout
            else:
                fl = 2 * hw * 16 * cin * cout if kind == 2 else 2 * (hw // (s * s)) * kk * kk * cin * cout
            tflop += fl
            bn = re.search(r"conv_(?:tc|halo|hs)_kernel<\(int\)(\d+)>|conv_(?:tc|halo|hs)_kernel<(\d+)>", k)
            print("op%-3s kind %d k%d s%d cin %4d cout %4d /%-2d BN=%-3s grid=%-16s %8.1f us %7.1f TF/s" % (
                o[1], kind, kk, s, cin, cout, down, ((bn.group(1) or bn.group(2)) + ("h" if "halo" in k else ("s" if "conv_hs" in k else ""))) if bn else ("128w" if "conv_sw" in k else ("fuse" if "conv_bneck" in k else ("tail" if "conv_segtail" in k else "?"))), g, t, fl / t / 1e6))
        ttc = sum(t for _, t, _ in tc)
        print("conv_tc total %.1f us, %.1f GFLOP -> %.1f TF/s" % (ttc, tflop / 1e9, tflop / ttc / 1e6))


if __name__ == "__main__":
    main()
