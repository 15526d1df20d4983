"""Summarise .ncu-rep files into a small text file for profiles/ (run here, no GPU needed):
    python scripts/ncu_summary.py out.txt rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def main():
    out = open(sys.argv[1], "w")
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        out.write("# %s\n" % rep)
        for r in rows[2:]:
            out.write("kernel: %s\n" % r[idx["Kernel Name"]])
            for w in WANT:
                if w in idx:
                    out.write("  %-70s %s %s\n" % (w, r[idx[w]], units[idx[w]]))
            tr = float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])
            out.write("  traffic (dram read+write, units as above)                              %.6f\n" % tr)
            st = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_")
                         and not h.endswith("_not_issued")), reverse=True)[:6]
            tot = sum(float(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_")
                      and not h.endswith("_not_issued"))
            out.write("  top stall reasons (pc samples, share): " + ", ".join(
                "%s %.0f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / max(tot, 1)) for v, h in st) + "\n")
    out.close()
    print(open(sys.argv[1]).read())


if __name__ == "__main__":
    main()
