"""GPU parity tests: the CUDA path, called through the C ABI (via dawnsearch_b200.index, a thin
ctypes layer), against the CPU oracle on the same inputs.  Labels and distances must be
BIT-IDENTICAL: the scan only selects candidates, the final scores are recomputed on the device
in the reference's order of summation (src/search/vector.rs:128-134).

Written to read like uses of the reference's index (src/search/search_provider.rs):
new_index(&INDEX_OPTIONS), reserve, add(id, &v), search(&q, 20) -> Matches{labels, distances}.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
SEED = 0xDA5EA2C4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_same(got: "tuple", want: "tuple", ctx=""):
    gl, gd = got
    wl, wd = want
    assert len(gl) == len(wl), (ctx, len(gl), len(wl))
    assert (np.asarray(gl) == np.asarray(wl)).all(), (ctx, gl, wl)
    assert (bits(gd) == bits(wd)).all(), (ctx, gd, wd)


def perm_labels(oracle, n, salt=3, base=1):
    return (np.argsort(oracle.np_mix64(np.arange(n, dtype=np.uint64) + np.uint64(salt)), kind="stable")
            + base).astype(np.uint64)


@pytest.fixture(scope="module")
def corpus5k(oracle):
    n = 5000
    rows = oracle.np_synth_rows_f32(SEED, 0, n)
    labels = perm_labels(oracle, n, base=1001)
    return rows, labels, oracle.store_f16(rows)


@pytest.fixture(scope="module")
def index5k(dawn, corpus5k):
    rows, labels, _ = corpus5k
    idx = dawn.new_index(dawn.IndexOptions())
    idx.reserve(len(rows))
    idx.add_batch(labels, rows)
    yield idx
    idx.close()


def test_golden_vectors_through_the_c_abi(dawn, oracle):
    g = np.load(GOLDEN)
    n = int(g["n"])
    rows = oracle.np_synth_rows_f32(int(g["seed"]), 0, n)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        for i in range(0, n, 1000):  # several add_batch calls, like a rebuild from SQLite
            idx.add_batch(g["labels"][i:i + 1000], rows[i:i + 1000])
        assert idx.size() == n
        for k in (1, 10, 20, 100):
            labels, dist, counts = idx.search_batch(g["queries"], k)
            assert (counts == k).all()
            assert (labels == g[f"labels_k{k}"]).all()
            assert (bits(dist) == g[f"dist_k{k}"]).all()


@pytest.mark.parametrize("k", [1, 10, 20, 100, 120])
def test_single_query_matches_oracle(index5k, corpus5k, oracle, k):
    rows, labels, stored = corpus5k
    for q in oracle.make_queries(SEED, 5, 6, len(rows)):
        m = index5k.search(q, k)
        assert_same((m.labels, m.distances), oracle.search_f16(stored, labels, q, k), f"k={k}")


@pytest.mark.parametrize("batch", [1, 2, 3, 4, 5, 7, 8, 33])
def test_batches_match_oracle(index5k, corpus5k, oracle, batch):
    rows, labels, stored = corpus5k
    qs = oracle.make_queries(SEED, 100 + batch, batch, len(rows))
    gl, gd, cnt = index5k.search_batch(qs, 10)
    for i, q in enumerate(qs):
        assert cnt[i] == 10
        assert_same((gl[i], gd[i]), oracle.search_f16(stored, labels, q, 10), f"batch={batch} i={i}")


@pytest.mark.parametrize("n", [1, 7, 8, 9, 19, 255, 256, 257, 1000, 4097])
def test_ragged_sizes_and_k_larger_than_n(dawn, oracle, n):
    rows = oracle.np_synth_rows_f32(77, 0, n)
    stored = oracle.store_f16(rows)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(np.arange(1, n + 1, dtype=np.uint64), rows)
        for k in (1, 20, 100):
            for q in oracle.make_queries(77, 78, 3, n):
                m = idx.search(q, k)
                assert len(m.labels) == min(k, n)  # fewer than k when size < k
                assert_same((m.labels, m.distances), oracle.search_f16(stored, None, q, k), f"n={n} k={k}")


def test_empty_index_and_k_zero(dawn, oracle):
    q = oracle.np_synth_rows_f32(1, 0, 1)[0]
    with dawn.new_index(dawn.IndexOptions()) as idx:
        assert idx.size() == 0 and idx.capacity() == 0 and idx.dimensions() == 384
        m = idx.search(q, 20)
        assert len(m.labels) == 0
        idx.reserve(10)
        idx.add(5, q)
        assert len(idx.search(q, 0).labels) == 0
        with pytest.raises(dawn.DawnError):
            idx.search(q, 121)


def test_self_match_and_ordering_invariants(index5k, corpus5k):
    rows, labels, _ = corpus5k
    for r in (0, 1234, 4999):
        m = index5k.search(rows[r], 20)  # the reference's k (search_provider.rs:214)
        assert m.labels[0] == labels[r]
        assert m.distances[0] < 1e-3     # src/net/web.rs:339 "exploring" test
        assert (np.diff(m.distances) >= 0).all()


def test_ties_break_on_lower_label(dawn, oracle):
    base = oracle.np_synth_rows_f32(5, 0, 3)
    rows = np.concatenate([np.repeat(base[:1], 700, axis=0), base[1:], np.repeat(base[:1], 300, axis=0)])
    n = len(rows)
    labels = perm_labels(oracle, n, salt=9)
    stored = oracle.store_f16(rows)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(labels, rows)
        for k in (10, 100):
            m = idx.search(base[0], k)
            assert_same((m.labels, m.distances), oracle.search_f16(stored, labels, base[0], k), f"k={k}")
            assert list(m.labels) == sorted(m.labels.tolist())  # all tied: ascending labels


def test_long_runs_of_equal_scores_across_all_cta_lists(dawn, oracle):
    """Tens of thousands of identical rows: every per-CTA list is full of tied candidates, so the
    merge cannot prune by score and has to order thousands of ties by label."""
    base = oracle.np_synth_rows_f32(6, 0, 2)
    n_dup = 40_000
    rows = np.concatenate([np.repeat(base[:1], n_dup, axis=0), oracle.np_synth_rows_f32(7, 0, 20_000)])
    n = len(rows)
    labels = perm_labels(oracle, n, salt=11)
    stored = oracle.store_f16(rows)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(labels, rows)
        for k in (1, 20, 100):
            m = idx.search(base[0], k)
            assert_same((m.labels, m.distances), oracle.search_f16(stored, labels, base[0], k), f"k={k}")
            assert list(m.labels) == sorted(labels[:n_dup].tolist())[:k]
        qs = np.stack([base[0], base[1], base[0]])
        gl, gd, cnt = idx.search_batch(qs, 10)
        for i, q in enumerate(qs):
            assert cnt[i] == 10
            assert_same((gl[i], gd[i]), oracle.search_f16(stored, labels, q, 10), f"batch row {i}")


def test_add_then_search_and_reserve_growth(dawn, oracle):
    """The reference's insert pattern: reserve(size+1024) when full, then add one vector
    (search_provider.rs:280-284); every add is visible to the next search."""
    n = 2600
    rows = oracle.np_synth_rows_f32(91, 0, n)
    stored = oracle.store_f16(rows)
    q = oracle.make_queries(91, 92, 1, n)[0]
    with dawn.new_index(dawn.IndexOptions()) as idx:
        for i in range(n):
            if idx.size() == idx.capacity():
                idx.reserve(idx.size() + 1024)
            idx.add(i + 1, rows[i])
            if i in (0, 1, 1023, 1024, 2047, 2599):
                m = idx.search(q, 20)
                assert_same((m.labels, m.distances), oracle.search_f16(stored[: i + 1], None, q, 20), f"i={i}")
        assert idx.size() == n and idx.capacity() >= n
        with pytest.raises(dawn.DawnError) as ei:  # beyond capacity: error, not a silent grow
            idx.add_batch(np.arange(5000, dtype=np.uint64), np.zeros((5000, 384), dtype=np.float32))
        assert ei.value.code == -3


def test_save_load_roundtrip(dawn, oracle, index5k, corpus5k, tmp_path):
    rows, labels, stored = corpus5k
    path = str(tmp_path / "index.dawn")
    index5k.save(path)
    assert os.path.getsize(path) == 32 + len(rows) * (8 + 768)
    q = oracle.make_queries(SEED, 55, 1, len(rows))[0]
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.load(path)
        assert idx.size() == len(rows)
        m = idx.search(q, 20)
        assert_same((m.labels, m.distances), oracle.search_f16(stored, labels, q, 20))
        # a failed load leaves the index unchanged (the reference then rebuilds, :115-116)
        with open(path, "r+b") as f:
            f.truncate(1000)
        with pytest.raises(dawn.DawnError) as ei:
            idx.load(path)
        assert ei.value.code == -4
        assert idx.size() == len(rows)
        with pytest.raises(dawn.DawnError):
            idx.load(str(tmp_path / "missing"))


def test_get_returns_the_stored_vector(index5k, corpus5k, dawn):
    rows, labels, stored = corpus5k
    for r in (0, 77, 4999):
        v = index5k.get(int(labels[r]))
        assert (bits(v) == bits(stored[r].astype(np.float32))).all()
    with pytest.raises(dawn.DawnError):
        index5k.get(999_999_999)


def test_device_generator_is_bit_identical_to_oracle(dawn, oracle):
    n = 3000
    first = 123_456_789
    with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
        idx.add_synthetic(SEED, first, n)
        assert idx.size() == n
        want = oracle.synth_rows_f16(SEED, first, n)
        for r in (0, 1, 255, 256, 1500, n - 1):
            v = idx.get(first + r + 1)  # labels are first_row + i + 1
            assert (bits(v) == bits(want[r].astype(np.float32))).all()
        labels = np.arange(first + 1, first + n + 1, dtype=np.uint64)
        for q in oracle.make_queries(SEED, 66, 3, 1000):
            m = idx.search(q, 10)
            assert_same((m.labels, m.distances), oracle.search_f16(want, labels, q, 10))


@pytest.mark.timeout(300)
def test_two_million_rows_against_threaded_cpu_scan(dawn, oracle):
    n = 2_000_000
    with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
        idx.add_synthetic(SEED, 0, n)
        stored = oracle.synth_rows_f16(SEED, 0, n)  # same rows on the host (bit-identical generator)
        qs = oracle.make_queries(SEED, SEED + 1, 16, n)
        for k in (10, 100):
            gl, gd, cnt = idx.search_batch(qs, k)
            wl, wd, wc, _ = oracle.cpu_scan_f16(stored, None, qs, k)
            assert (cnt == wc).all()
            assert (gl == wl).all()
            assert (bits(gd) == bits(wd)).all()
        prof = idx.profile()
        assert prof["uncertified"] == 0
        # The same corpus one query at a time (the streaming-scan path, ring slots reused ~50 times per CTA),
        # first for fp16-path k' = 16, 32 and 128.  Regression: the scan used to hang here when a consumer
        # warp waited at the prune barrier with landed stages of its own still in the ring.
        for k in (10, 20, 100):
            wl, wd, wc, _ = oracle.cpu_scan_f16(stored, None, qs[:6], k)
            for i in range(6):
                m = idx.search(qs[i], k)
                assert_same((m.labels, m.distances), (wl[i][: wc[i]], wd[i][: wc[i]]), f"single query {i} k={k}")
        # and right after a bulk add (the ingest path), as a rebuild from SQLite would do it
    rows = oracle.np_synth_rows_f32(SEED + 9, 0, 200_000)
    labels = np.arange(1, len(rows) + 1, dtype=np.uint64)
    stored = oracle.store_f16(rows)
    with dawn.new_index(dawn.IndexOptions(capacity=len(rows))) as idx:
        idx.add_batch(labels, rows)
        for k in (1, 10):
            m = idx.search(rows[0], k)
            assert_same((m.labels, m.distances), oracle.search_f16(stored, labels, rows[0], k), f"after add_batch k={k}")


def test_full_size_properties_10m(dawn, oracle):
    """BASELINE config C2 (10M x 384 fp16, batch 1, k=10) through size-independent properties:
    planted neighbours are found, every returned distance equals the oracle's exact score of
    that (regenerated) row, results are sorted, top-10 is a prefix of top-100, and two
    half-corpus shards merged on the device give the same answer as the whole corpus."""
    n = 10_000_000
    nq = 8
    qs = oracle.make_queries(SEED, SEED + 1, nq, n)
    planted = oracle.planted_rows(SEED + 1, nq, n)
    with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
        idx.add_synthetic(SEED, 0, n)
        l10, d10, c10 = idx.search_batch(qs, 10)
        l100, d100, c100 = idx.search_batch(qs, 100)
        assert (c10 == 10).all() and (c100 == 100).all()
        assert (l100[:, :10] == l10).all() and (bits(d100[:, :10]) == bits(d10)).all()
        assert (np.diff(d100, axis=1) >= 0).all()
        for i, r in enumerate(planted):
            assert l10[i, 0] == r + 1
        for i in range(nq):
            for j in range(10):
                row = oracle.synth_rows_f16(SEED, int(l10[i, j]) - 1, 1)[0]
                want = np.float32(1.0) - np.float32(oracle.score_f16(row, qs[i]))
                assert bits(d10[i, j]) == bits(want)
        assert idx.profile()["uncertified"] == 0
        whole = (l10.copy(), d10.copy())

    # the same corpus as two id-range shards + device merge (the multi-GPU path on one GPU)
    import torch

    half = n // 2
    a = dawn.new_index(dawn.IndexOptions(capacity=half))
    b = dawn.new_index(dawn.IndexOptions(capacity=half))
    a.add_synthetic(SEED, 0, half)
    b.add_synthetic(SEED, half, half)
    la, da, ca = a.search_batch(qs, 10)
    lb, db, cb = b.search_batch(qs, 10)
    a.close()
    b.close()
    dev = torch.device("cuda:0")
    L = torch.from_numpy(np.stack([la, lb]).astype(np.int64)).to(dev)
    Dd = torch.from_numpy(np.stack([da, db])).to(dev)
    Cn = torch.from_numpy(np.stack([ca, cb]).astype(np.int32)).to(dev)
    Lo = torch.zeros((nq, 10), dtype=torch.int64, device=dev)
    Do = torch.zeros((nq, 10), dtype=torch.float32, device=dev)
    Co = torch.zeros(nq, dtype=torch.int32, device=dev)
    dawn.merge_results_device(0, L.data_ptr(), Dd.data_ptr(), Cn.data_ptr(), 2, nq, 10, Lo.data_ptr(),
                              Do.data_ptr(), Co.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (Co.cpu().numpy() == 10).all()
    assert (Lo.cpu().numpy().astype(np.uint64) == whole[0]).all()
    assert (bits(Do.cpu().numpy()) == bits(whole[1])).all()


# ------------------------------------------------------------------ K3: tensor-core path


@pytest.fixture(scope="module")
def index300k(dawn):
    n = 300_000
    idx = dawn.new_index(dawn.IndexOptions(capacity=n))
    idx.add_synthetic(SEED, 0, n)
    yield idx, n
    idx.close()


@pytest.fixture(scope="module")
def stored300k(oracle):
    return oracle.synth_rows_f16(SEED, 0, 300_000)


@pytest.mark.parametrize("batch,k", [(1, 10), (7, 20), (16, 10), (128, 10), (129, 100), (300, 100), (1024, 10)])
def test_gemm_path_matches_oracle(index300k, stored300k, oracle, batch, k):
    """Large batches run as tcgen05 tiles with the top-k fused in the epilogue; the fp16-rounded
    queries only SELECT candidates, so labels and distances are still bit-identical."""
    idx, n = index300k
    idx.set_option("force_path", 2)
    try:
        idx.profile(reset=True)
        qs = oracle.make_queries(SEED, 1000 + batch, batch, n)
        gl, gd, cnt = idx.search_batch(qs, k)
        prof = idx.profile(reset=True)
        assert prof["gemm_batches"] == 1 and prof["scan_launches"] == 0  # really the tensor-core path
        assert prof["escalations"] == 0 and prof["uncertified"] == 0
        wl, wd, wc, _ = oracle.cpu_scan_f16(stored300k, None, qs, k)
        assert (cnt == wc).all()
        assert (gl == wl).all()
        assert (bits(gd) == bits(wd)).all()
    finally:
        idx.set_option("force_path", 0)


@pytest.mark.parametrize("n", [1024, 1025, 1279, 1280, 4097, 70_001])
def test_gemm_path_ragged_row_counts(dawn, oracle, n):
    rows = oracle.np_synth_rows_f32(4242, 0, n)
    stored = oracle.store_f16(rows)
    labels = perm_labels(oracle, n, salt=11, base=7)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(labels, rows)
        idx.set_option("force_path", 2)
        qs = oracle.make_queries(4242, 4243, 20, n)
        for k in (10, 100):
            gl, gd, cnt = idx.search_batch(qs, k)
            for i, q in enumerate(qs):
                assert_same((gl[i, : cnt[i]], gd[i, : cnt[i]]), oracle.search_f16(stored, labels, q, k), f"n={n} k={k} i={i}")
        assert idx.profile()["gemm_batches"] == 2


def test_gemm_path_duplicates_escalate_to_exact_scan(dawn, oracle):
    """Thousands of identical pages: the tensor-core path cannot certify the label tie-break,
    the host API re-runs those queries through the exact scan, results stay exact."""
    base = oracle.np_synth_rows_f32(5, 0, 2)
    rows = np.concatenate([np.repeat(base[:1], 3000, axis=0), oracle.np_synth_rows_f32(6, 0, 2000)])
    n = len(rows)
    labels = perm_labels(oracle, n, salt=13)
    stored = oracle.store_f16(rows)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(labels, rows)
        idx.set_option("force_path", 2)
        qs = np.stack([base[0], base[1]])
        gl, gd, cnt = idx.search_batch(qs, 10)
        for i, q in enumerate(qs):
            assert_same((gl[i], gd[i]), oracle.search_f16(stored, labels, q, 10), f"i={i}")
        assert idx.profile()["escalations"] >= 1


def test_auto_path_selection(index300k, oracle):
    idx, n = index300k
    idx.profile(reset=True)
    idx.search_batch(oracle.make_queries(SEED, 1, 4, n), 10)
    p = idx.profile(reset=True)
    assert p["gemm_batches"] == 0 and p["scan_launches"] == 1
    idx.search_batch(oracle.make_queries(SEED, 2, 64, n), 10)
    p = idx.profile(reset=True)
    assert p["gemm_batches"] == 1 and p["scan_launches"] == 0


def test_gemm_path_adversarial_row_order_stays_exact(dawn, oracle):
    """Rows stored so that thousands of near neighbours of the query come LAST.  With sequential
    rounds the thresholds learnt early would be useless and the candidate log would overflow (the
    overflow -> exact-scan fallback was verified that way); the strided tile permutation makes every
    round a uniform sample, so the tensor-core path copes.  Either way the answer is bit-identical."""
    n_bg, n_hot = 70_000, 6_000
    bg = oracle.np_synth_rows_f32(8, 0, n_bg)
    q = oracle.np_synth_rows_f32(9, 0, 1)[0]
    noise = oracle.np_synth_rows_f32(10, 0, n_hot)
    hot = q[None, :] + 0.3 * noise
    hot = (hot / np.linalg.norm(hot, axis=1, keepdims=True)).astype(np.float32)
    rows = np.concatenate([bg, hot])
    n = len(rows)
    stored = oracle.store_f16(rows)
    with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
        idx.add_batch(np.arange(1, n + 1, dtype=np.uint64), rows)
        idx.set_option("force_path", 2)
        qs = np.stack([q] + list(oracle.make_queries(8, 11, 3, n_bg)))
        gl, gd, cnt = idx.search_batch(qs, 10)
        prof = idx.profile()
        assert prof["gemm_batches"] == 1 and prof["uncertified"] == 0
        wl, wd, wc, _ = oracle.cpu_scan_f16(stored, None, qs, 10)
        assert (gl == wl).all() and (bits(gd) == bits(wd)).all()
