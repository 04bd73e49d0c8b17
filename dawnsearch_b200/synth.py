"""Host-side (numpy) twin of the device corpus generator in csrc/ingest.cu (synth_f16_kernel).

Used by bench.py and examples to make queries for a corpus that was generated on the GPU with
`Index.add_synthetic` (100M vectors cannot cross PCIe in a test budget).  Every step is integer
arithmetic or a correctly rounded IEEE f64 operation, so host and device agree bit for bit.
This is a workload generator, not a search path.
"""
from __future__ import annotations

import numpy as np

EM_LEN = 384
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def mix64(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def rows_f32(seed: int, first_row: int, n: int) -> np.ndarray:
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(first_row))[:, None]
    cols = np.arange(EM_LEN, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = mix64(np.uint64(seed) + _GOLD * (rows * np.uint64(EM_LEN) + cols + np.uint64(1)))
    m = np.uint64(0xFFFF)
    s = (h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))
    raw = s.astype(np.int64) - 131070
    sumsq = (raw * raw).sum(axis=1)
    zero = sumsq == 0
    if zero.any():
        raw[zero, 0] = 1
        sumsq[zero] = 1
    inv = 1.0 / np.sqrt(sumsq.astype(np.float64))
    return (raw.astype(np.float64) * inv[:, None]).astype(np.float32)


def planted_rows(query_seed: int, nq: int, n_rows: int, planted_fraction: float = 0.5) -> np.ndarray:
    n_pl = int(nq * planted_fraction) if n_rows > 0 else 0
    return (mix64(np.arange(n_pl, dtype=np.uint64) + np.uint64(query_seed + 77)) % np.uint64(max(n_rows, 1))).astype(np.int64)


def make_queries(corpus_seed: int, query_seed: int, nq: int, n_rows: int,
                 planted_fraction: float = 0.5) -> np.ndarray:
    """f32 unit queries: the first nq*planted_fraction are noisy copies of stored rows (so a real
    nearest neighbour exists: label planted_rows()[i] + 1), the rest are unrelated unit vectors."""
    qs = rows_f32(query_seed, 0, nq).astype(np.float64)
    for i, r in enumerate(planted_rows(query_seed, nq, n_rows, planted_fraction)):
        qs[i] = rows_f32(corpus_seed, int(r), 1)[0].astype(np.float64) + 0.35 * qs[i]
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    return qs.astype(np.float32)
