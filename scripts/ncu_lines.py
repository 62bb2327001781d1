"""Instructions / stall samples per CUDA source line (needs -lineinfo + --import-source).  usage: ncu_lines.py rep kernel-regex [topN]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = {}
hdr = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ia = r.index("Instructions Executed"); isamp = r.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[2] == "-":          # a CUDA source line (its SASS rows follow with an address)
        try: a = int(r[ia]); sm = int(r[isamp])
        except ValueError: continue
        if a: agg[(cur, int(r[0]), r[1].strip()[:100])] = (a, sm)
tot = sum(a for a, _ in agg.values()); ts = sum(s for _, s in agg.values())
print("total warp-instrs", tot, "samples", ts)
for k, (a, sm) in sorted(agg.items(), key=lambda t: -t[1][0])[:topn]:
    print("%5.1f%% instr %5.1f%% smp  %s:%d  %s" % (100.0 * a / tot, 100.0 * sm / max(ts, 1), k[0], k[1], k[2]))
