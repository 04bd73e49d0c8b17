"""Dev aid: run the tensor-core path on a small synthetic corpus and compare with the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
SEED = 0xDA5EA2C4
idx = D.new_index(D.IndexOptions(capacity=n))
idx.add_synthetic(SEED, 0, n)
idx.set_option("force_path", 2)
qs = O.make_queries(SEED, 11, batch, n)
t0 = time.time()
gl, gd, cnt = idx.search_batch(qs, k)
print("gemm search done in %.3f s" % (time.time() - t0), idx.profile())
stored = O.synth_rows_f16(SEED, 0, n)
wl, wd, wc, _ = O.cpu_scan_f16(stored, None, qs, k)
bad = 0
for i in range(batch):
    same = (gl[i] == wl[i]).all() and (gd[i].view(np.uint32) == wd[i].view(np.uint32)).all()
    if not same:
        bad += 1
        if bad <= 3:
            print("MISMATCH q", i, "\n got ", gl[i][:10], gd[i][:5], "\n want", wl[i][:10], wd[i][:5])
print("mismatching queries:", bad, "of", batch)
