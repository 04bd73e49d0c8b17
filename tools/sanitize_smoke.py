"""Small shapes of every kernel path, for compute-sanitizer (memcheck / racecheck) runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dawnsearch_b200 as D
from oracle import oracle as O

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 3000
rows = O.np_synth_rows_f32(5, 0, n)
labels = np.arange(1, n + 1, dtype=np.uint64)
qs = O.make_queries(5, 6, 300, n)
stored = O.store_f16(rows)
if which in ("all", "scan"):
    with D.new_index(D.IndexOptions(capacity=n)) as idx:
        idx.add_batch(labels, rows)
        for b in (1, 2, 4):
            gl, gd, cnt = idx.search_batch(qs[:b], 10)
            assert (gl[0] == O.search_f16(stored, None, qs[0], 10)[0]).all()
        print("scan ok", flush=True)
if which in ("all", "gemm"):
    with D.new_index(D.IndexOptions(capacity=n)) as idx:
        idx.add_batch(labels, rows)
        idx.set_option("force_path", 2)
        for b in (8, 300):
            gl, gd, cnt = idx.search_batch(qs[:b], 10)
            assert (gl[0] == O.search_f16(stored, None, qs[0], 10)[0]).all()
        print("gemm ok", idx.profile()["gemm_batches"], flush=True)
if which in ("all", "i8"):
    q8, sc = O.store_i8(rows)
    with D.new_index(D.IndexOptions(capacity=n, quantization=D.ScalarKind.I8)) as idx:
        idx.add_batch(labels, rows)
        for b in (1, 2):
            gl, gd, cnt = idx.search_batch(qs[:b], 10)
            assert (gl[0] == O.search_i8(q8, sc, None, qs[0], 10)[0]).all()
        print("i8 ok", flush=True)
if which in ("all", "i8gemm"):
    # native int8 tensor-core kernel (gemm_i8.cu) + exact re-score selects; needs >= 65,536 rows
    n8 = 66_000
    with D.new_index(D.IndexOptions(capacity=n8, quantization=D.ScalarKind.I8)) as idx:
        idx.add_synthetic(5, 0, n8)
        idx.set_option("i8_tensor_min_batch", 4)
        rows8 = np.concatenate([O.np_synth_rows_f32(5, i, min(11000, n8 - i)) for i in range(0, n8, 11000)])
        q8, sc = O.store_i8(rows8)
        q2 = O.make_queries(5, 7, 140, n8)
        for b in (8, 140):  # one CTA per tile / CTA pairs
            gl, gd, cnt = idx.search_batch(q2[:b], 10)
            assert (gl[0] == O.search_i8(q8, sc, None, q2[0], 10)[0]).all()
        print("i8gemm ok", idx.profile()["gemm_batches"], flush=True)
if which in ("all", "shadow"):
    # fp16 corpus filtered through its int8 shadow (gemm_i8.cu in shadow mode + shadow_quantize_kernel); needs >= 65,536 rows
    n8 = 66_000
    with D.new_index(D.IndexOptions(capacity=n8)) as idx:
        idx.set_option("shadow_i8", 1)
        idx.add_synthetic(5, 0, n8)
        st16 = O.synth_rows_f16(5, 0, n8)
        q2 = O.make_queries(5, 7, 140, n8)
        for b in (8, 140):
            gl, gd, cnt = idx.search_batch(q2[:b], 10)
            assert (gl[0] == O.search_f16(st16, None, q2[0], 10)[0]).all()
        print("shadow ok", idx.profile()["shadow_batches"], flush=True)
if which in ("all", "f32"):
    with D.new_index(D.IndexOptions(capacity=n, quantization=D.ScalarKind.F32)) as idx:
        idx.add_batch(labels, rows)
        for force in (1, 2):
            idx.set_option("force_path", force)
            gl, gd, cnt = idx.search_batch(qs[:9], 10)
            assert (gl[0] == O.search_f32(rows, None, qs[0], 10)[0]).all()
        assert idx.verify()["bad_rows"] == 0
        print("f32 ok", flush=True)
if which in ("all", "multi"):
    with D.MultiIndex([0, 0]) as m:  # two shards on one GPU: peer-copy exchange + merge kernel (with its duplicate-label pass)
        m.reserve(n)
        m.add_batch(labels, rows)
        gl, gd, cnt = m.search_batch(qs[:5], 10)
        assert (gl[0] == O.search_f16(stored, None, qs[0], 10)[0]).all()
        print("multi ok", flush=True)
