"""DAWN_SCALAR_F32 -- the reference's own index setting (ScalarKind::F32, src/search/search_provider.rs:38): vectors are kept
as given, candidates are SELECTED on fp16 copies by the same kernels, and the survivors are re-scored on the f32 vectors in
the reference's order of summation (src/search/vector.rs:128-134).  Labels and distances must equal an exact f32 brute force
over the vectors as added (oracle/dawn_oracle.c:dawn_oracle_search_f32), bit for bit -- this is BASELINE config C1's data
type (100k f32 page vectors, single query, k = 10)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 0xDA5EA2C4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def c1(dawn, oracle):
    n = 100_000
    rows = np.concatenate([oracle.np_synth_rows_f32(SEED, i, 20000) for i in range(0, n, 20000)])
    labels = np.arange(1, n + 1, dtype=np.uint64)
    idx = dawn.new_index(dawn.IndexOptions(quantization=dawn.ScalarKind.F32))
    idx.reserve(n)
    idx.add_batch(labels, rows)
    yield idx, rows, labels
    idx.close()


def test_c1_single_queries_match_exact_f32_brute_force(c1, oracle):
    idx, rows, labels = c1
    qs = oracle.make_queries(SEED, SEED + 1, 40, len(rows))
    for q in qs:
        for k in (10, 20):  # k = 20 is what SearchProvider::search_embedding asks for (search_provider.rs:214)
            m = idx.search(q, k)
            wl, wd = oracle.search_f32(rows, labels, q, k)
            assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all(), k
    # a stored page queried with its own embedding comes back first with distance < 0.001 (src/net/web.rs:339)
    m = idx.search(rows[777], 20)
    assert m.labels[0] == 778 and m.distances[0] < 1e-3
    assert idx.profile()["uncertified"] == 0


@pytest.mark.parametrize("force", [1, 2])
def test_f32_batches_on_both_selection_paths(c1, oracle, force):
    idx, rows, labels = c1
    idx.set_option("force_path", force)
    try:
        for batch, k in ((7, 10), (130, 100)):
            qs = oracle.make_queries(SEED, 50 + batch, batch, len(rows))
            idx.profile(reset=True)
            gl, gd, cnt = idx.search_batch(qs, k)
            p = idx.profile()
            assert (p["gemm_batches"] > 0) == (force == 2) and p["uncertified"] == 0
            for i, q in enumerate(qs):
                wl, wd = oracle.search_f32(rows, labels, q, k)
                assert cnt[i] == k and (gl[i] == wl).all() and (bits(gd[i]) == bits(wd)).all(), (force, batch, i)
    finally:
        idx.set_option("force_path", 0)


def test_f32_get_save_load_and_growth(dawn, oracle, tmp_path):
    n = 9000
    rows = oracle.np_synth_rows_f32(3, 0, n)
    labels = (np.arange(n, dtype=np.uint64) * np.uint64(5) + np.uint64(11))
    q = oracle.make_queries(3, 4, 1, n)[0]
    path = str(tmp_path / "f32.idx")
    with dawn.new_index(dawn.IndexOptions(quantization=dawn.ScalarKind.F32)) as idx:
        for lo in range(0, n, 1000):            # the reference's reserve(size + 1024) pattern (search_provider.rs:280-283)
            idx.reserve(idx.size() + 1024)
            idx.add_batch(labels[lo:lo + 1000], rows[lo:lo + 1000])
        assert (idx.get(int(labels[4321])) == rows[4321]).all()   # the vector exactly as added, not an fp16 rounding of it
        idx.save(path)
        want = oracle.search_f32(rows, labels, q, 20)
        m = idx.search(q, 20)
        assert (m.labels == want[0]).all() and (bits(m.distances) == bits(want[1])).all()
    with dawn.new_index(dawn.IndexOptions(quantization=dawn.ScalarKind.F32)) as idx:
        idx.load(path)
        assert idx.size() == n
        m = idx.search(q, 20)
        assert (m.labels == want[0]).all() and (bits(m.distances) == bits(want[1])).all()
        assert idx.verify()["bad_rows"] == 0
    with dawn.new_index(dawn.IndexOptions()) as idx16:    # an fp16 index refuses the f32 file, and stays usable
        with pytest.raises(dawn.DawnError):
            idx16.load(path)


def test_f32_synthetic_rows_equal_the_oracle_generator(dawn, oracle):
    n = 30_000
    with dawn.new_index(dawn.IndexOptions(quantization=dawn.ScalarKind.F32, capacity=n)) as idx:
        idx.add_synthetic(SEED, 5, n)
        rows = oracle.np_synth_rows_f32(SEED, 5, n)
        for r in (0, 1, 12345, n - 1):
            assert (idx.get(5 + r + 1) == rows[r]).all()
        q = oracle.make_queries(SEED, 9, 1, n)[0]
        m = idx.search(q, 10)
        wl, wd = oracle.search_f32(rows, np.arange(6, n + 6, dtype=np.uint64), q, 10)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
