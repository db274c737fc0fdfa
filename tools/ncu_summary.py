"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs."""
import csv, io, re, subprocess, sys
KEYS = [
 "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
 "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
 "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
 "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
 "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed",
 "lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
 "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])[-60:] for r in data]
    print(f"# {path}")
    print(f"{'metric':88s} {'unit':10s} " + "  ".join(f"{n[-28:]:>28s}" for n in names))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:88s} {units[i]:10s} " + "  ".join(f"{r[i]:>28s}" for r in data))
if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
