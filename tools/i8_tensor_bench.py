"""Large batches over an int8 shard on the tensor cores: native kind::i8 (gemm_i8.cu) against the fp16-tile variant
(i8_tensor.cu) in one process.  Times whole host-API calls and the library's own CUDA events around the rounds. Prints JSON.

    python tools/i8_tensor_bench.py [rows=62500000] [batch=1024] [variants=1,0]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
from dawnsearch_b200 import synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 62_500_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
variants = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,0").split(",")]
out = {"rows": rows, "batch": batch, "algorithmic_int8_ops_per_batch": 2.0 * rows * batch * 384,
       "algorithmic_bytes_per_pass": rows * 388}
with D.new_index(D.IndexOptions(capacity=rows, quantization=D.ScalarKind.I8)) as idx:
    idx.add_synthetic(0xDA5EA2C4, 0, rows)
    qs = synth.make_queries(0xDA5EA2C4, 4, batch, rows, planted_fraction=0.25)
    idx.set_option("i8_tensor_min_batch", 8)
    ref = {}
    for native in variants:
        idx.set_option("i8_native", native)
        name = "native_kind_i8" if native else "fp16_tiles_via_dequant_scratch"
        out[name] = {}
        for k in (10, 100):
            r = idx.search_batch(qs, k)
            if k in ref:  # both variants are exact: they must agree bit for bit
                assert (r[0] == ref[k][0]).all() and (r[1].view(np.uint32) == ref[k][1].view(np.uint32)).all()
            ref[k] = r
            idx.set_profiling(True)
            p0 = idx.profile(reset=True)
            t = time.perf_counter()
            reps = 4
            for _ in range(reps): idx.search_batch(qs, k)
            dt = (time.perf_counter() - t) / reps
            p1 = idx.profile(reset=True)
            idx.set_profiling(False)
            gemm_ms = p1["gemm_ms"] / reps
            out[name][f"k{k}"] = {"ms_per_batch_host_api": round(dt * 1e3, 2), "qps": round(batch / dt),
                                  "rounds_ms_cuda_events": round(gemm_ms, 2), "finalize_ms": round(p1["finalize_ms"] / reps, 3),
                                  "tera_int8_ops_per_s": round(2.0 * rows * batch * 384 / (gemm_ms / 1e3) / 1e12, 1),
                                  "corpus_stream_GBps": round(rows * 388 / (gemm_ms / 1e3) / 1e9, 1),
                                  "gemm_runs": int(p1["gemm_batches"]) // reps, "kernel_launches": int(p1["kernel_launches"]) // reps,
                                  "escalations": int(p1["escalations"]), "uncertified": int(p1["uncertified"])}
print(json.dumps(out))
