#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into text: key counters of every captured kernel
and the stall-sample distribution per code region (split at barrier waits) of the first kernel."""
import csv, subprocess, sys, io, collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("KERNEL", r[hdr.index("Kernel Name")][:70], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for k in KEYS:
        if k in hdr:
            print(f"   {k:90s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seg = rows[ks[0] + 1:ks[1]]
hdr, data = seg[0], seg[1:]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print(f"\nSOURCE PAGE of {rows[ks[0]][1][:60]}: {len(data)} SASS instructions, {tot} samples")
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stall}
print("  " + ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
print("  top instructions by samples:")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:40]):
    r = data[i]
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall), reverse=True)[:2]
    print(f"  {i:6d} {int(r[iS]):8d} {100*int(r[iS])/tot:5.1f}% exec {r[iEx]:>10s}  {r[iSrc].strip()[:70]:70s} {st}")

# ---- barrier-wait attribution: wait_bar(bar, parity, tag) leaves `tag` as an immediate right after its spin loop
import re
TAGS = {100: "producer: empty[slot]", 200: "mma: a_ready[0]", 201: "mma: a_ready[1]", 202: "mma: acc_free[0]", 203: "mma: acc_free[1]",
        220: "mma: full[slot]", 230: "mma: full[slot] (skip)", 300: "epi GATE acc_full", 310: "epi L acc_full[0]", 311: "epi L acc_full[1]",
        320: "epi FEAT acc_full[0]", 321: "epi FEAT acc_full[1]", 340: "epi VIEWS acc_full", 350: "epi RGB acc_full"}
waits = collections.Counter()
i = 0
while i < len(data):
    if "TRYWAIT" in data[i][iSrc]:
        n, j = 0, i
        while j < len(data):   # the wait shows up on the branch that consumes the try_wait predicate
            n += int(data[j][iS])
            if "BRA" in data[j][iSrc]:
                break
            j += 1
        tag = None
        for j in range(i + 1, min(i + 30, len(data))):
            m = re.search(r"(?:MOV|IMAD\.MOV\.U32) R\d+, (?:RZ, RZ, )?0x([0-9a-f]+) ;?$", data[j][iSrc].strip().rstrip(";").strip() + " ;")
            if m and 100 <= int(m.group(1), 16) <= 400:
                tag = int(m.group(1), 16); break
        key = tag - (tag % 10 if tag and 100 <= tag < 110 or tag and 220 <= tag < 240 else 0) if tag else None
        waits[key] += n
    i += 1
print("\n  barrier-wait samples by site (share of all samples; one fully-waiting warp = %.1f%%):" % (100.0 / 10))
for k, v in sorted(waits.items(), key=lambda x: -x[1]):
    print(f"    {TAGS.get(k, k)!s:32s} {v:9d} {100*v/tot:5.1f}%")
