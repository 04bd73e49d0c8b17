"""Measures K1 (host f32 -> device fp16/int8 append) through dawn_index_add_batch (the bulk pipeline: 4 copier
threads, pinned double buffers, 4 streams), and save/load of a 10M-row index.  Prints JSON.

    python tools/ingest_rate.py [rows=4000000] [save_rows=10000000] [dir=/dev/shm]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
save_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
where = sys.argv[3] if len(sys.argv) > 3 else "/dev/shm"
rng = np.random.default_rng(0)
base = rng.standard_normal((200_000, 384)).astype(np.float32)
base /= np.linalg.norm(base, axis=1, keepdims=True)
rows = np.concatenate([base] * (n // len(base) + 1))[:n]  # content does not matter for the copy rate; pages are touched
labels = np.arange(1, n + 1, dtype=np.uint64)
out = {"rows": n, "host_bytes": n * 1536, "host_threads": os.cpu_count(),
       "note": "pageable host memory in, every row through the f32->fp16 / int8 conversion kernel"}
for name, kind in (("f16", D.ScalarKind.F16), ("i8", D.ScalarKind.I8)):
    with D.new_index(D.IndexOptions(capacity=n, quantization=kind)) as idx:
        idx.add_batch(labels[:100000], rows[:100000])  # warm: creates the pipeline's pinned buffers and streams
        best = None
        for rep in range(3):
            idx2 = D.new_index(D.IndexOptions(capacity=n, quantization=kind))
            idx2.add_batch(labels[:40000], rows[:40000])
            t0 = time.perf_counter()
            idx2.add_batch(labels[40000:], rows[40000:])   # returns when every row is on the device
            dt = time.perf_counter() - t0
            m = n - 40000
            if best is None or dt < best[0]:
                best = (dt, m)
            assert idx2.size() == n
            idx2.close()
        dt, m = best
        out[name] = {"seconds": round(dt, 4), "rows_per_s": round(m / dt), "host_GBps": round(m * 1536 / dt / 1e9, 2)}
# save / load of a synthetic index (device -> file -> device), double-buffered through pinned memory
try:
    path = os.path.join(where, "dawn_ingest_rate.idx")
    with D.new_index(D.IndexOptions(capacity=save_rows)) as idx:
        idx.add_synthetic(0xDA5EA2C4, 0, save_rows)
        t0 = time.perf_counter()
        idx.save(path)
        ts = time.perf_counter() - t0
        nbytes = os.path.getsize(path)
        ref = idx.search(base[0], 10)
    with D.new_index(D.IndexOptions()) as idx:
        t0 = time.perf_counter()
        idx.load(path)
        tl = time.perf_counter() - t0
        got = idx.search(base[0], 10)
        assert (got.labels == ref.labels).all() and idx.size() == save_rows
        v = idx.verify()
    os.remove(path)
    out["save_load"] = {"rows": save_rows, "file_bytes": nbytes, "dir": where, "save_s": round(ts, 3), "save_GBps": round(nbytes / ts / 1e9, 2),
                        "load_s": round(tl, 3), "load_GBps": round(nbytes / tl / 1e9, 2), "verify_after_load": v}
except Exception as e:  # noqa: BLE001
    out["save_load"] = f"failed: {e!r}"
print(json.dumps(out))
