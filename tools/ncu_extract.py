#!/usr/bin/env python
"""ncu-rep -> the `kernel,metric,value,unit` rows committed under profiles/ (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.sum", "lts__t_requests_srcunit_tex.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps")


def main():
    rep, name, out = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    with open(out, "w") as f:
        f.write("kernel,metric,value,unit\n")
        for i, h in enumerate(hdr):
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio"):
                f.write("%s,%s,%s,%s\n" % (name, h, vals[i], units[i]))
    print("wrote", out)


if __name__ == "__main__":
    main()
