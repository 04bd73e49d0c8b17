"""One small workload for an ncu capture: python tools/ncu_target.py <kind> [rows] [batch] [k]
kind: i8gemm | f16gemm | f16scan | i8scan.  Runs one warm-up search and one measured search through the host API."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dawnsearch_b200 as D
from dawnsearch_b200 import synth
kind = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else (1024 if "gemm" in kind else 1)
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
quant = D.ScalarKind.I8 if kind.startswith("i8") else D.ScalarKind.F16
with D.new_index(D.IndexOptions(capacity=rows, quantization=quant)) as idx:
    idx.add_synthetic(0xDA5EA2C4, 0, rows)
    for kv in filter(None, os.environ.get("DAWN_OPTS", "").split(",")):  # e.g. DAWN_OPTS=gemm_chunk_tiles=8,gemm_cta_group=1
        key, val = kv.split("=")
        idx.set_option(key, int(val))
    qs = synth.make_queries(0xDA5EA2C4, 4, batch, rows, planted_fraction=0.25)
    for _ in range(2):
        r = idx.search_batch(qs, k)
    print(kind, rows, batch, k, r[0][0][:3], idx.profile())
