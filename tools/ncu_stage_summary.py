#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep of the stand-alone stage kernels (tools/stage_step.py): duration, DRAM and L2
bytes and rates, hit rates, sectors per request, issue / occupancy - the counters north_star asks for on the gather stage."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "smsp__inst_executed.sum"]
def val(r, k):
    return r[hdr.index(k)] if k in hdr else None
for r in rows[2:]:
    name = val(r, "Kernel Name")
    print(f"KERNEL {name[:80]}  grid {val(r, 'Grid Size')} block {val(r, 'Block Size')}")
    for k in KEYS:
        v = val(r, k)
        if v is not None:
            print(f"   {k:72s} {v:>18s} {units[hdr.index(k)]}")
    try:
        f = lambda k: float(val(r, k).replace(",", ""))
        unit_t = units[hdr.index("gpu__time_duration.sum")]
        t = f("gpu__time_duration.sum") * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit_t, 1e-9)
        ub = lambda k: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index(k)], 1.0)
        dram = f("dram__bytes_read.sum") * ub("dram__bytes_read.sum") + f("dram__bytes_write.sum") * ub("dram__bytes_write.sum")
        l2 = f("lts__t_sectors.sum") * 32.0
        print(f"   => DRAM {dram / 1e6:9.1f} MB at {dram / t / 1e9:7.1f} GB/s   L2 {l2 / 1e6:9.1f} MB at {l2 / t / 1e9:7.1f} GB/s   ({t * 1e6:.1f} us)")
        ld_s, ld_r = f("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), f("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
        if ld_r:
            print(f"   => global loads: {ld_s / ld_r:.2f} sectors / request")
        st_s, st_r = f("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"), f("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum")
        if st_r:
            print(f"   => global stores: {st_s / st_r:.2f} sectors / request")
    except Exception as e:
        print("   (derived figures unavailable:", e, ")")
