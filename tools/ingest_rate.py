"""Measures K1 (host f32 -> device fp16/int8 append) through dawn_index_add_batch. Prints JSON."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(0)
rows = rng.standard_normal((n, 384)).astype(np.float32)
rows /= np.linalg.norm(rows, axis=1, keepdims=True)
labels = np.arange(1, n + 1, dtype=np.uint64)
out = {"rows": n, "host_bytes": n * 1536}
for name, kind in (("f16", D.ScalarKind.F16), ("i8", D.ScalarKind.I8)):
    with D.new_index(D.IndexOptions(capacity=n, quantization=kind)) as idx:
        idx.add_batch(labels[:10000], rows[:10000])  # warm
    with D.new_index(D.IndexOptions(capacity=n, quantization=kind)) as idx:
        t0 = time.perf_counter()
        idx.add_batch(labels, rows)
        idx.search(rows[0], 1)  # forces the last flush
        dt = time.perf_counter() - t0
        out[name] = {"seconds": round(dt, 4), "rows_per_s": round(n / dt), "host_GBps": round(n * 1536 / dt / 1e9, 2)}
print(json.dumps(out))
