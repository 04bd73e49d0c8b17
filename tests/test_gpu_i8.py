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


@pytest.mark.parametrize("native", [1, 0])
def test_i8_tensor_core_path_matches_oracle(dawn, oracle, native):
    """Large batches over an int8 corpus on the tensor cores.  native = 1 (gemm_i8.cu, the default): tcgen05 kind::i8
    straight from the blocked int8 arena with a one-level int8 query, every logged candidate re-scored exactly between
    rounds.  native = 0 (i8_tensor.cu, the A/B partner): chunks dequantised to an fp16 scratch, the fp16 rounds over each
    chunk, per-chunk lists gathered.  Either way labels and distances must be bit-identical to the int8 oracle, with
    ragged sizes, a query batch that is not a multiple of the tile, and both CTA-pair and one-CTA tiles."""
    n = 200_003
    with dawn.new_index(i8_options(dawn, capacity=n)) as idx:
        idx.add_synthetic(SEED, 0, n)
        f32 = np.concatenate([oracle.np_synth_rows_f32(SEED, i, min(20000, n - i)) for i in range(0, n, 20000)])
        q8, sc = oracle.store_i8(f32)
        idx.set_option("i8_native", native)
        idx.set_option("i8_tensor_min_batch", 8)
        idx.set_option("i8_tensor_chunk_rows", 65536)  # native = 0: 4 chunks
        for nq in (40, 130):                           # one 128-query tile (one CTA per tile) / CTA pairs
            qs = oracle.make_queries(SEED, 33 + nq, nq, n)
            for k in (10, 100):
                before = idx.profile()
                gl, gd, cnt = idx.search_batch(qs, k)
                after = idx.profile()
                assert after["gemm_batches"] - before["gemm_batches"] == (1 if native else 4)
                assert after["escalations"] == before["escalations"]
                for i, q in enumerate(qs):
                    wl, wd = oracle.search_i8(q8, sc, None, q, k)
                    assert cnt[i] == k
                    assert (gl[i] == wl).all() and (bits(gd[i]) == bits(wd)).all(), (native, nq, k, i)
        assert idx.profile()["uncertified"] == 0
        # below the batch threshold the scan path answers, with the same result
        m = idx.search(qs[0], 10)
        wl, wd = oracle.search_i8(q8, sc, None, qs[0], 10)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()


def test_i8_native_tensor_path_ties_limit_and_small_corpus(dawn, oracle):
    """kind::i8 path corner cases: exact duplicates (ties break on the label; the overflowing candidate log escalates to the
    exact scan), a pushed-down distance limit, and a corpus smaller than one round."""
    n = 70_000
    f32 = oracle.np_synth_rows_f32(SEED + 9, 0, n).copy()
    f32[1000:1300] = f32[999]                                # 301 identical pages
    labels = np.arange(1, n + 1, dtype=np.uint64)
    q8, sc = oracle.store_i8(f32)
    with dawn.new_index(i8_options(dawn, capacity=n)) as idx:
        idx.add_batch(labels, f32)
        idx.set_option("i8_tensor_min_batch", 4)
        qs = np.concatenate([f32[999:1000], oracle.make_queries(SEED + 9, 5, 19, n)])
        for k in (10, 100):
            gl, gd, cnt = idx.search_batch(qs, k)
            for i, q in enumerate(qs):
                wl, wd = oracle.search_i8(q8, sc, labels, q, k)
                assert (gl[i] == wl).all() and (bits(gd[i]) == bits(wd)).all(), (k, i)
        assert idx.profile()["gemm_batches"] >= 2
        wl, wd = oracle.search_i8(q8, sc, labels, qs[3], 20)
        for lim in (float(wd[5]), float(wd[-1]) + 0.3, 0.0):
            gl, gd, cnt = idx.search_batch_limit(qs, 20, lim)
            for i, q in enumerate(qs):
                wl, wd = oracle.search_i8(q8, sc, labels, q, 20)
                keep = int((wd < np.float32(lim)).sum())
                assert cnt[i] == keep and (gl[i][:keep] == wl[:keep]).all() and (bits(gd[i][:keep]) == bits(wd[:keep])).all()
