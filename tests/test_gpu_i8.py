"""GPU parity for the int8-stored corpus (K4, BASELINE config C5): per-row absmax/127 scales,
dp4a scan with a two-level int8 query split, exact f32 re-score.  Labels and distances must be
bit-identical to oracle/dawn_oracle.c:dawn_oracle_search_i8 over the same stored bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 0xDA5EA2C4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def perm_labels(oracle, n, salt=3, base=1):
    return (np.argsort(oracle.np_mix64(np.arange(n, dtype=np.uint64) + np.uint64(salt)), kind="stable") + base).astype(np.uint64)


def i8_options(dawn, **kw):
    return dawn.IndexOptions(quantization=dawn.ScalarKind.I8, **kw)


def test_i8_golden_vectors_through_the_c_abi(dawn, oracle):
    """tests/golden/golden_v2.npz: int8 storage + search pinned by the numpy restatement (make_golden_v2.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz"))
    n = int(g["n"])
    rows = oracle.np_synth_rows_f32(int(g["seed"]), 0, n)
    with dawn.new_index(i8_options(dawn)) as idx:
        idx.reserve(n)
        idx.add_batch(g["labels"], rows)
        for k in (1, 10, 20, 100):
            labels, dist, counts = idx.search_batch(g["queries"], k)
            assert (counts == k).all()
            assert (labels == g[f"labels_k{k}"]).all()
            assert (bits(dist) == g[f"dist_k{k}"]).all()
        for i in range(3):  # the wire form of the same queries gives the same hits as the decoded query
            wire = g["i24_wire"][i].tobytes()
            assert dawn.encode_i24(g["queries"][i]) == wire
            dec = dawn.decode_i24(wire)
            assert (bits(dec) == g["i24_decoded"][i]).all()
            m = idx.search_i24(wire, 10)
            wl, wd = oracle.search_i8(*oracle.store_i8(rows), g["labels"], dec, 10)
            assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()


@pytest.mark.parametrize("n", [1, 7, 8, 9, 255, 257, 1000, 5000])
def test_i8_search_matches_oracle(dawn, oracle, n):
    rows = oracle.np_synth_rows_f32(SEED, 0, n)
    q8, sc = oracle.store_i8(rows)
    labels = perm_labels(oracle, n, base=100)
    with dawn.new_index(i8_options(dawn)) as idx:
        idx.reserve(n)
        idx.add_batch(labels, rows)
        assert idx.size() == n
        for k in (1, 10, 100):
            for q in oracle.make_queries(SEED, 9, 4, n):
                m = idx.search(q, k)
                wl, wd = oracle.search_i8(q8, sc, labels, q, k)
                assert len(m.labels) == min(k, n)
                assert (m.labels == wl).all(), (n, k, m.labels, wl)
                assert (bits(m.distances) == bits(wd)).all(), (n, k, m.distances, wd)
        assert idx.profile()["uncertified"] == 0


@pytest.mark.parametrize("batch", [1, 2, 3, 5, 8])
def test_i8_batches(dawn, oracle, batch):
    n = 3000
    rows = oracle.np_synth_rows_f32(12, 0, n)
    q8, sc = oracle.store_i8(rows)
    with dawn.new_index(i8_options(dawn)) as idx:
        idx.reserve(n)
        idx.add_batch(np.arange(1, n + 1, dtype=np.uint64), rows)
        qs = oracle.make_queries(12, 13 + batch, batch, n)
        gl, gd, cnt = idx.search_batch(qs, 20)
        for i, q in enumerate(qs):
            wl, wd = oracle.search_i8(q8, sc, None, q, 20)
            assert cnt[i] == 20 and (gl[i] == wl).all() and (bits(gd[i]) == bits(wd)).all()


def test_i8_stored_vector_generator_and_persistence(dawn, oracle, tmp_path):
    n, first = 3000, 7_000_000
    rows = oracle.synth_rows_f32(SEED, first, n)
    q8, sc = oracle.store_i8(rows)
    labels = np.arange(first + 1, first + n + 1, dtype=np.uint64)
    with dawn.new_index(i8_options(dawn, capacity=n)) as idx:
        idx.add_synthetic(SEED, first, n)  # device generator + device quantiser
        for r in (0, 7, 8, 1500, n - 1):
            v = idx.get(first + r + 1)
            want = q8[r].astype(np.float32) * sc[r]
            assert (bits(v) == bits(want)).all()
        qs = oracle.make_queries(SEED, 21, 3, 1000)
        for q in qs:
            m = idx.search(q, 10)
            wl, wd = oracle.search_i8(q8, sc, labels, q, 10)
            assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
        path = str(tmp_path / "i8.dawn")
        idx.save(path)
        with dawn.new_index(i8_options(dawn)) as idx2:
            idx2.load(path)
            assert idx2.size() == n
            m = idx2.search(qs[0], 10)
            wl, wd = oracle.search_i8(q8, sc, labels, qs[0], 10)
            assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
        with dawn.new_index(dawn.IndexOptions()) as f16idx:  # storage kinds do not mix
            with pytest.raises(dawn.DawnError):
                f16idx.load(path)


def test_i8_ties_and_growth(dawn, oracle):
    base = oracle.np_synth_rows_f32(5, 0, 2)
    rows = np.concatenate([np.repeat(base[:1], 500, axis=0), oracle.np_synth_rows_f32(6, 0, 700)])
    n = len(rows)
    labels = perm_labels(oracle, n, salt=5)
    q8, sc = oracle.store_i8(rows)
    with dawn.new_index(i8_options(dawn)) as idx:
        for i in range(n):  # the reference's insert pattern (search_provider.rs:280-284)
            if idx.size() == idx.capacity():
                idx.reserve(idx.size() + 1024)
            idx.add(int(labels[i]), rows[i])
        m = idx.search(base[0], 10)
        wl, wd = oracle.search_i8(q8, sc, labels, base[0], 10)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
        assert list(m.labels) == sorted(m.labels.tolist())


def test_i8_200k_rows(dawn, oracle):
    n = 200_000
    rows = oracle.synth_rows_f32(SEED, 0, n) if False else None
    with dawn.new_index(i8_options(dawn, capacity=n)) as idx:
        idx.add_synthetic(SEED, 0, n)
        # regenerate the same rows on the host in slices (bit-identical generator + quantiser)
        f32 = np.concatenate([oracle.np_synth_rows_f32(SEED, i, 20000) for i in range(0, n, 20000)])
        q8, sc = oracle.store_i8(f32)
        qs = oracle.make_queries(SEED, 31, 6, n)
        for k in (10, 100):
            gl, gd, cnt = idx.search_batch(qs, k)
            for i, q in enumerate(qs):
                wl, wd = oracle.search_i8(q8, sc, None, q, k)
                assert (gl[i] == wl).all() and (bits(gd[i]) == bits(wd)).all()
        assert idx.profile()["uncertified"] == 0


def test_i8_tensor_core_path_matches_oracle(dawn, oracle):
    """Opt-in path for large batches over an int8 corpus (i8_tensor.cu): chunks dequantised to an fp16 scratch, the
    tcgen05 rounds over each chunk, per-chunk candidate lists gathered, exact int8 re-score.  Labels and distances
    must still be bit-identical to the int8 oracle, over several chunks and with ragged chunk sizes."""
    n = 200_000
    with dawn.new_index(i8_options(dawn, capacity=n)) as idx:
        idx.add_synthetic(SEED, 0, n)
        f32 = np.concatenate([oracle.np_synth_rows_f32(SEED, i, 20000) for i in range(0, n, 20000)])
        q8, sc = oracle.store_i8(f32)
        qs = oracle.make_queries(SEED, 33, 40, n)
        idx.set_option("i8_tensor_min_batch", 8)
        idx.set_option("i8_tensor_chunk_rows", 65536)  # 4 chunks of 50,176 / 50,176 / 50,176 / 49,472 rows
        for k in (10, 100):
            before = idx.profile()
            gl, gd, cnt = idx.search_batch(qs, k)
            after = idx.profile()
            assert after["gemm_batches"] - before["gemm_batches"] == 4  # one run of the rounds per chunk
            for i, q in enumerate(qs):
                wl, wd = oracle.search_i8(q8, sc, None, q, k)
                assert cnt[i] == k
                assert (gl[i] == wl).all() and (bits(gd[i]) == bits(wd)).all(), (k, i)
        assert idx.profile()["uncertified"] == 0
        # below the batch threshold the scan path answers, with the same result
        m = idx.search(qs[0], 10)
        wl, wd = oracle.search_i8(q8, sc, None, qs[0], 10)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
