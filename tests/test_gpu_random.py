"""Randomised differential test: random corpus sizes, duplicate rows, arbitrary (even repeated)
u64 labels, random k / batch / storage / path -- the C-ABI result must equal the oracle bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make_case(oracle, rng, case):
    n = int(rng.choice([1, 2, 7, 8, 9, 15, 16, 17, 100, 255, 256, 257, 1023, 1024, 1025, 2500, 6000]))
    rows = oracle.np_synth_rows_f32(1000 + case, 0, n)
    if n > 4 and rng.random() < 0.5:  # duplicate blocks -> exact ties
        src = rng.integers(0, n, size=max(1, n // 3))
        dst = rng.integers(0, n, size=len(src))
        rows[dst] = rows[src]
    mode = rng.integers(0, 3)
    if mode == 0:
        labels = np.arange(1, n + 1, dtype=np.uint64)                       # SQLite rowids
    elif mode == 1:
        labels = rng.integers(1, 2 ** 63, size=n, dtype=np.int64).astype(np.uint64) * np.uint64(2) + np.uint64(1)  # huge ids
    else:
        labels = rng.integers(1, max(2, n // 2), size=n, dtype=np.int64).astype(np.uint64)  # repeated labels
    return n, rows, labels


@pytest.mark.parametrize("case", range(24))
def test_random_case(dawn, oracle, case):
    rng = np.random.default_rng(case)
    n, rows, labels = make_case(oracle, rng, case)
    storage = "i8" if case % 3 == 2 else "f16"
    k = int(rng.choice([1, 2, 10, 20, 33, 100, 120]))
    batch = int(rng.choice([1, 2, 3, 4, 5, 9, 40]))
    qs = oracle.make_queries(1000 + case, 77 + case, batch, n)
    if n > 2 and rng.random() < 0.5:
        qs[0] = rows[rng.integers(0, n)]  # a stored vector as query (self match, maybe tied)
    opts = dawn.IndexOptions(quantization=dawn.ScalarKind.I8 if storage == "i8" else dawn.ScalarKind.F16)
    with dawn.new_index(opts) as idx:
        idx.reserve(n)
        cut = int(rng.integers(0, n + 1))
        idx.add_batch(labels[:cut], rows[:cut])
        for i in range(cut, min(n, cut + 3)):
            idx.add(int(labels[i]), rows[i])
        if cut + 3 < n:
            idx.add_batch(labels[cut + 3:], rows[cut + 3:])
        assert idx.size() == n
        if storage == "f16" and n >= 1024 and case % 2 == 0:
            idx.set_option("force_path", 2)  # tensor-core path (escalates to the scan where it cannot certify)
        gl, gd, cnt = idx.search_batch(qs, k)
        if storage == "f16":
            stored = oracle.store_f16(rows)
            want = [oracle.search_f16(stored, labels, q, k) for q in qs]
        else:
            q8, sc = oracle.store_i8(rows)
            want = [oracle.search_i8(q8, sc, labels, q, k) for q in qs]
        for i, (wl, wd) in enumerate(want):
            assert cnt[i] == len(wl) == min(k, n)
            assert (gl[i, : cnt[i]] == wl).all(), (case, n, k, batch, storage, i, gl[i, : cnt[i]], wl)
            assert (bits(gd[i, : cnt[i]]) == bits(wd)).all(), (case, n, k, batch, storage, i)
