"""Single-query scan time at a given corpus size for the library named by DAWN_B200_LIB (A/B of builds on one box)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dawnsearch_b200 as D
from dawnsearch_b200 import synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
with D.new_index(D.IndexOptions(capacity=rows)) as idx:
    idx.add_synthetic(0xDA5EA2C4, 0, rows)
    qs = synth.make_queries(0xDA5EA2C4, 4, 16, rows)
    for q in qs[:5]: idx.search(q, k)
    out = []
    for rep in range(3):
        idx.set_profiling(True); idx.profile(reset=True)
        for i in range(30): idx.search(qs[i % 16], k)
        p = idx.profile(reset=True); idx.set_profiling(False)
        ms = p["scan_ms"] / p["scan_launches"]
        out.append(round(rows * 768 / ms / 1e6, 1))
    print(os.environ.get("DAWN_B200_LIB", "current"), rows, "GB/s per rep:", out)
