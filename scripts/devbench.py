"""Developer micro-benchmark (not the judged bench.py): times one batched kernel on HBM-resident data.
usage: python scripts/devbench.py huf_dec [MiB] [p]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import probagen
from gpu_common import cpu_compress, BLOCK, SLOT
import finitestateentropy_b200 as fb

what = sys.argv[1] if len(sys.argv) > 1 else "huf_dec"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.14
n = mib << 20
t0 = time.time(); data = probagen(n, p); t1 = time.time()
codec = "huf" if what.startswith("huf") else "fse"
cbuf, cs, slot = cpu_compress(codec, data, slot=SLOT); t2 = time.time()
print("gen %.2fs cpu-compress %.2fs (%d threads) ratio %.4f" % (t1 - t0, t2 - t1, os.cpu_count(), cs.sum() / n))
d_c = torch.from_numpy(cbuf).cuda(); d_s = torch.from_numpy(cs.view(np.int64)).cuda(); d_d = torch.from_numpy(data).cuda()
out = torch.empty(n, dtype=torch.uint8, device="cuda"); res = torch.empty(len(cs), dtype=torch.int64, device="cuda")
fn = fb.huf_decompress_batch if codec == "huf" else fb.fse_decompress_batch
for _ in range(3):
    fn(d_c, d_s, n, BLOCK, slot, out=out, results=res)
torch.cuda.synchronize()
ok = bool((out == d_d).all()) and bool((res == torch.tensor([min(BLOCK, n - b * BLOCK) for b in range(len(cs))], device="cuda")).all())
ts = []
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(d_c, d_s, n, BLOCK, slot, out=out, results=res); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
best = min(ts); med = sorted(ts)[len(ts) // 2]
alg = n + int(cs.sum())
print("%s %d MiB p=%.2f ok=%s best %.3f ms med %.3f ms -> %.1f GB/s uncompressed, %.1f GB/s algorithmic (%.1f%% of 6566)" %
      (what, mib, p, ok, best, med, n / best / 1e6, alg / best / 1e6, alg / best / 1e6 / 65.66))
