"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per kernel + hottest SASS lines.
usage: python scripts/ncu_summary.py file.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct"]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    print("=" * 100)
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d: print("%-70s %s" % (k, d[k][:90]))
    st = sorted(((float(v), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items() if k.startswith(STALL) and v not in ("", "n/a")), reverse=True)
    print("stalls/issue: " + ", ".join("%s=%.2f" % (k, v) for v, k in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:]:
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    name = lines[0][1][:80]; h = lines[1]; data = [l for l in lines[2:] if len(l) == len(h)]
    ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    tot = sum(int(x[isamp]) for x in data); ti = sum(int(x[ia]) for x in data)
    print("-" * 100); print(name, "samples", tot, "warp-instrs", ti)
    for i, x in sorted(sorted(enumerate(data), key=lambda t: -int(t[1][isamp]))[:topn]):
        print("%6d %-72s %10s %7s %5.1f%%" % (i, x[isrc][:72], x[ia], x[isamp], 100.0 * int(x[isamp]) / max(tot, 1)))
