#!/usr/bin/env python
"""Text summary of one kernel of an .ncu-rep (run here, no GPU needed): headline metrics, stall reasons per issue,
warp-state samples by reason and the most sampled instructions.

    python tools/ncu_summary.py report.ncu-rep > profiles/rNN_<kernel>_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    ix = {h: i for i, h in enumerate(hdr)}
    print("report: %s" % rep.split("/")[-1])
    print("kernel: %s" % vals[ix["Kernel Name"]][:160])
    for w in WANT:
        if w in ix:
            print("%s %s %s" % (w, vals[ix[w]], units[ix[w]]))
    print("--- stalls (warps per issue-active cycle)")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = float(vals[ix[h]] or 0)
            if v >= 0.005:
                print("   %-22s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    src = page(rep, "source")
    sh, data = src[1], src[2:]
    sx = {h: i for i, h in enumerate(sh)}
    tot = sum(int(r[sx["# Samples"]] or 0) for r in data)
    print("--- warp-state samples by reason (source page), total %d" % tot)
    for k in [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]:
        n = sum(int(r[sx[k]] or 0) for r in data)
        if n * 200 >= tot:
            print("  %-22s %9d  %.1f%%" % (k, n, 100.0 * n / max(tot, 1)))
    print("--- twelve most sampled instructions")
    for r in sorted(data, key=lambda r: -int(r[sx["# Samples"]] or 0))[:12]:
        st = {k: int(r[sx[k]] or 0) for k in sh if k.startswith("stall_") and "Not Issued" not in k}
        top = max(st.items(), key=lambda x: x[1])
        print("  %7s  %-58s %s" % (r[sx["# Samples"]], r[sx["Source"]][:58], top[0]))


if __name__ == "__main__":
    main()
