"""Per-kernel summary of an .ncu-rep (--set full): one block per distinct kernel (first launch; `xN` = launches in the report) with the
metrics the roofline discussion uses.  usage: ncu_kernels_summary.py report.ncu-rep [raw.csv] > summary.txt"""
import csv, io, re, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_red.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
seen = {}
for r in data:
    seen.setdefault(r[ki], []).append(r)
for name, rs in seen.items():
    print(f"\n## {name}   x{len(rs)}  (durations, ms: " + ", ".join(f"{float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')):.3f}" for r in rs) + ")")
    r = rs[0]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:86s} {units[i]:16s} {r[i]}")
