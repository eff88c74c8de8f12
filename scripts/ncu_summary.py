"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + stall samples segmented by marker instructions.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "derived__memory_l1_wavefronts_shared_excessive", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_lookup_hit.sum", "l1tex__t_sectors_lookup_miss.sum"]
for k in keys:
    for h in hdr:
        if h == k or h.endswith("." + k):
            print("%-90s %s %s" % (k, m[h], units[hdr.index(h)]))
            break
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(m[h] or 0) > 0.05:
        print("%-90s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall "), m[h]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2, data = rows[1], rows[2:]
ia, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "SASS instructions", len(data))
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"BAR\.SYNC|SYNCS\.PHASECHK|UTCBAR|SYNCS\.ARRIVE|USETMAXREG")
cur = [0, 0, 0]
for i, r in enumerate(data):
    cur[0] += int(r[isamp]); cur[1] += int(r[iex]); cur[2] += 1
    if pat.search(r[ia]):
        if cur[0] > tot * 0.003:
            print("%5d %-58s samples %6d (%4.1f%%) warp-instr %10d sass %d" % (i, r[ia].strip()[:58], cur[0], 100 * cur[0] / tot, cur[1], cur[2]))
        cur = [0, 0, 0]
print("tail", cur)
