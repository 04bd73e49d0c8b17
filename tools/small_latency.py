"""Latency breakdown of single-query searches on small corpora (the reference's real scale, <= 1M pages)."""
import os, sys, time, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
from dawnsearch_b200 import synth
out = {}
for n in (100_000, 1_000_000):
    with D.new_index(D.IndexOptions(capacity=n)) as idx:
        idx.add_synthetic(0xDA5EA2C4, 0, n)
        qs = synth.make_queries(0xDA5EA2C4, 4, 64, n)
        for q in qs[:10]: idx.search(q, 20)
        lat = []
        for i in range(300):
            t = time.perf_counter(); idx.search(qs[i % 64], 20); lat.append((time.perf_counter() - t) * 1e6)
        idx.set_profiling(True); idx.profile(reset=True)
        for i in range(100): idx.search(qs[i % 64], 20)
        p = idx.profile(reset=True)
        out[n] = {"e2e_p50_us": round(statistics.median(lat), 1), "e2e_p99_us": round(sorted(lat)[296], 1),
                  "scan_us": round(p["scan_ms"] * 10, 1), "finalize_us": round(p["finalize_ms"] * 10, 1)}
print(json.dumps(out))
