"""Generates tests/golden/golden_v1.npz.

The reference holds no golden vectors for this path ("parity unpinned"),
so these are produced by the *numpy* restatement in oracle/oracle.py (np_synth_rows_f32 +
np_search), which shares no code with the C oracle or the CUDA kernels.  Both are then
checked against this file.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SEED = 0xDA5EA2C4
N = 6000
NQ = 12
KS = (1, 10, 20, 100)


def main():
    rows_f32 = O.np_synth_rows_f32(SEED, 0, N)
    stored = rows_f32.astype(np.float16)
    # non-monotone labels: a fixed permutation of 1..N offset by 1000
    perm = np.argsort(O.np_mix64(np.arange(N, dtype=np.uint64) + np.uint64(99)), kind="stable")
    labels = (perm.astype(np.uint64) + np.uint64(1001))
    queries = O.make_queries(SEED, SEED + 1, NQ, N)
    out = {
        "seed": np.uint64(SEED), "n": np.int64(N), "labels": labels, "queries": queries,
        # spot rows pin the generator and the fp16 rounding
        "rows_f32_head": rows_f32[:4], "stored_f16_head": stored[:4].view(np.uint16),
        "rows_f32_tail": rows_f32[-2:], "stored_f16_tail": stored[-2:].view(np.uint16),
    }
    sf = stored.astype(np.float32)
    for k in KS:
        labs = np.zeros((NQ, k), dtype=np.uint64)
        dist = np.zeros((NQ, k), dtype=np.float32)
        for i in range(NQ):
            l, d = O.np_search(sf, labels, queries[i], k)
            labs[i], dist[i] = l, d
        out[f"labels_k{k}"] = labs
        out[f"dist_k{k}"] = dist.view(np.uint32)  # bit patterns
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
