"""Per-SASS-line stall breakdown of the hottest region of a kernel in an .ncu-rep (read here, no GPU needed).
usage: python scripts/ncu_stalls.py file.ncu-rep [kernel-substring] [first-line last-line]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
for blk in src.split('"Kernel Name"')[1:]:
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    name = lines[0][1]
    if pat not in name: continue
    h = lines[1]; data = [l for l in lines[2:] if len(l) == len(h)]
    isrc, ia, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    st = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(x[isamp]) for x in data)
    agg = {c: sum(int(x[i] or 0) for x in data) for i, c in st}
    print(name[:90], "samples", tot)
    print("  totals:", ", ".join("%s=%.1f%%" % (c[6:], 100.0 * v / max(tot, 1)) for c, v in sorted(agg.items(), key=lambda t: -t[1])[:10]))
    lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0; hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(data)
    for k in range(lo, min(hi, len(data))):
        x = data[k]
        if int(x[isamp]) == 0 and len(sys.argv) <= 3: continue
        top = sorted(((int(x[i] or 0), c[6:]) for i, c in st), reverse=True)[:3]
        print("%6d %-64s %9s %6s  %s" % (k, x[isrc][:64], x[ia], x[isamp], " ".join("%s:%d" % (c, v) for v, c in top if v)))
