"""upper bound on what cross-shard threshold seeding could save.

Every query gets, from the very first tile, a threshold just below its final k-th score (the distance_limit push-down
does exactly that) -- no seeding scheme can do better.  Compared with the ordinary search on the same box/process."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
from dawnsearch_b200 import synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
batch, k = 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 10
idx = D.new_index(D.IndexOptions(capacity=rows))
idx.add_synthetic(0xDA5EA2C4, 0, rows)
q = synth.make_queries(0xDA5EA2C4, 3, batch, rows)
gl, gd, cnt = idx.search_batch(q, k)
kth = float(gd[:, k - 1].max())            # the loosest k-th distance of the batch
limit = kth + 1e-3                         # one limit for all queries, just beyond every query's k-th hit
ll, ld, lc = idx.search_batch_limit(q, k, limit)
assert (lc == k).all() and (ll == gl).all() and (ld.view(np.uint32) == gd.view(np.uint32)).all()
res = {"plain": [], "seeded": []}
for rep in range(4):
    for name in ("plain", "seeded"):
        idx.set_profiling(True); idx.profile(reset=True)
        for _ in range(6):
            if name == "plain": idx.search_batch(q, k)
            else: idx.search_batch_limit(q, k, limit)
        p = idx.profile(reset=True); idx.set_profiling(False)
        res[name].append(round(p["gemm_ms"] / p["gemm_batches"], 4))
med = {n: sorted(v)[len(v) // 2] for n, v in res.items()}
print(json.dumps({"rows": rows, "batch": batch, "k": k, "kth_distance_max": kth, "limit": limit, "gemm_ms": res,
                  "median_ms": med, "best_case_saving_pct": round(100 * (1 - med["seeded"] / med["plain"]), 2)}))
