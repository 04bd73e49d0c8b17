"""Generates tests/golden/golden_v2.npz: int8 storage + search, and the i24 wire codec.

Like golden_v1 these come from the *numpy* restatements in oracle/oracle.py (np_store_i8, np_search with
row scales, np_to24 / np_from24), which share no code with the C oracle or the CUDA kernels; both are checked
against this file.  Re-run:  python tests/golden/make_golden_v2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SEED = 0xDA5EA2C4
N = 4000
NQ = 8
KS = (1, 10, 20, 100)


def main():
    rows = O.np_synth_rows_f32(SEED + 7, 0, N)
    q8, scales = O.np_store_i8(rows)
    perm = np.argsort(O.np_mix64(np.arange(N, dtype=np.uint64) + np.uint64(5)), kind="stable")
    labels = perm.astype(np.uint64) + np.uint64(77)
    queries = O.make_queries(SEED + 7, SEED + 8, NQ, N)
    out = {"seed": np.uint64(SEED + 7), "n": np.int64(N), "labels": labels, "queries": queries,
           "i8_head": q8[:3], "scales_head": scales[:3].view(np.uint32),
           "i8_checksum": np.int64(q8.astype(np.int64).sum()), "scales_checksum": np.uint64(scales.view(np.uint32).astype(np.uint64).sum())}
    deq = q8.astype(np.float32)
    for k in KS:
        labs = np.zeros((NQ, k), dtype=np.uint64)
        dist = np.zeros((NQ, k), dtype=np.float32)
        for i in range(NQ):
            l, d = O.np_search(deq, labels, queries[i], k, row_scale=scales)
            labs[i], dist[i] = l, d
        out[f"labels_k{k}"] = labs
        out[f"dist_k{k}"] = dist.view(np.uint32)
    wire = np.stack([np.frombuffer(O.np_to24(queries[i]), dtype=np.uint8) for i in range(3)])
    out["i24_wire"] = wire
    out["i24_decoded"] = np.stack([O.np_from24(wire[i].tobytes()) for i in range(3)]).view(np.uint32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
