"""Batch-1024 search over an int8 shard through the opt-in tensor-core path (i8_tensor.cu). Prints JSON."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
from dawnsearch_b200 import synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 62_500_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
out = {"rows": rows, "batch": batch}
with D.new_index(D.IndexOptions(capacity=rows, quantization=D.ScalarKind.I8)) as idx:
    idx.add_synthetic(0xDA5EA2C4, 0, rows)
    qs = synth.make_queries(0xDA5EA2C4, 4, batch, rows)
    idx.set_option("i8_tensor_min_batch", 8)
    for k in (10, 100):
        idx.search_batch(qs, k)
        p0 = idx.profile()
        t = time.perf_counter()
        for _ in range(3): idx.search_batch(qs, k)
        dt = (time.perf_counter() - t) / 3
        p1 = idx.profile()
        out[f"k{k}"] = {"ms_per_batch": round(dt * 1e3, 2), "qps": round(batch / dt), "chunks": int(p1["gemm_batches"] - p0["gemm_batches"]) // 3,
                        "escalations": int(p1["escalations"] - p0["escalations"]), "uncertified": int(p1["uncertified"] - p0["uncertified"])}
print(json.dumps(out))
