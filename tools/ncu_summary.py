"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line per launch."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"),
        ("sm__inst_executed.sum", "winst"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "stsect"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "redsect"),
        ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long%"),
        ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "st_wait%"),
        ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg%"),
        ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_bar%")]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    parts = [name[:14].ljust(14)]
    for key, short in WANT:
        if key in hdr:
            v = r[hdr.index(key)]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            u = units[hdr.index(key)]
            parts.append("%s=%s%s" % (short, v, u if u in ("Mbyte", "Kbyte", "Gbyte", "byte") else ""))
    print(" ".join(parts))
