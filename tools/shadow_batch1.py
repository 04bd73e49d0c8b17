"""Single queries and small batches over an fp16 corpus: the TMA scan over the fp16 rows (768 B per row) against the int8-shadow
tensor path (388 B per row, option shadow_i8 with gemm_small_batch = 1).  Device-timed by the library's own events + wall clock."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
from dawnsearch_b200 import synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
idx = D.new_index(D.IndexOptions(capacity=rows))
idx.add_synthetic(0xDA5EA2C4, 0, rows)
out = {"rows": rows}
for batch in (1, 4, 64):
    qs = synth.make_queries(0xDA5EA2C4, 5, 64 * batch, rows).reshape(64, batch, 384)
    for name, opts in (("fp16_scan_or_tiles", {"shadow_i8": 0, "gemm_small_batch": 2}), ("int8_shadow", {"shadow_i8": 1, "gemm_small_batch": 1})):
        for kk, v in opts.items():
            idx.set_option(kk, v)
        ref = None
        for i in range(5):
            idx.search_batch(qs[i], 10)
        lat = []
        for i in range(64):
            t0 = time.perf_counter()
            r = idx.search_batch(qs[i], 10)
            lat.append((time.perf_counter() - t0) * 1e3)
        out[f"batch{batch}_{name}_e2e_ms_p50"] = round(float(np.median(lat)), 4)
        out[f"batch{batch}_{name}_shadow_batches"] = idx.profile(reset=True)["shadow_batches"]
        res = idx.search_batch(qs[0], 10)
        if name == "fp16_scan_or_tiles":
            keep = res
        else:
            out[f"batch{batch}_bit_identical"] = bool((res[0] == keep[0]).all() and (res[1].view(np.uint32) == keep[1].view(np.uint32)).all())
print(json.dumps(out))
