"""Summarise an .ncu-rep (ncu --set full --import-source on): per kernel the headline metrics of the raw page and the
SASS instructions that hold the most warp-stall samples (source page).  Usage: ncu_summary.py REP [REP...] > summary.txt"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    for rep in sys.argv[1:]:
        print("=" * 100)
        print(rep)
        raw = page(rep, "raw")
        hdr, units = raw[0], raw[1]
        src = page(rep, "source", ("--print-source", "sass"))
        starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
        for ki, r in enumerate(raw[2:]):
            print("-" * 100)
            print("kernel:", r[hdr.index("Kernel Name")], " grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
            for m in RAW:
                if m in hdr:
                    print("  %-88s %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
            if ki < len(starts):
                a = starts[ki]
                b = starts[ki + 1] if ki + 1 < len(starts) else len(src)
                h, seg = src[a + 1], src[a + 2:b]
                ie, ss = h.index("Instructions Executed"), h.index("# Samples")
                sc = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
                tot = sum(int(x[ss] or 0) for x in seg) or 1
                agg = sorted(((sum(int(x[i] or 0) for x in seg), h[i]) for i in sc), reverse=True)[:6]
                print("  warp-stall samples: %d; by reason: %s" % (tot, ", ".join("%s %.0f%%" % (n, 100.0 * v / tot) for v, n in agg)))
                print("  SASS instructions: %d, executed %d warp-instructions" % (len(seg), sum(int(x[ie] or 0) for x in seg)))
                top = sorted(range(len(seg)), key=lambda i: -int(seg[i][ss] or 0))[:8]
                for i in top:
                    x = seg[i]
                    st = max(((int(x[c] or 0), h[c]) for c in sc))
                    print("    %5.1f%% of samples  #%-5d %-58s (%s)" % (100.0 * int(x[ss] or 0) / tot, i, x[1].strip()[:58], st[1]))


if __name__ == "__main__":
    main()
