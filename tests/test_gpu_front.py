"""GPU tests of the callers / formats either side of the hot path:
micro-batching front, distance_limit, i24 wire queries, legacy .emb bulk load."""
import struct
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 0xDA5EA2C4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def small(dawn, oracle):
    n = 20000
    rows = oracle.np_synth_rows_f32(SEED, 0, n)
    idx = dawn.new_index(dawn.IndexOptions(capacity=n))
    idx.add_batch(np.arange(1, n + 1, dtype=np.uint64), rows)
    yield idx, rows, oracle.store_f16(rows)
    idx.close()


def test_batcher_coalesces_concurrent_callers(dawn, oracle, small):
    """32 threads x 8 single-query searches (the reference's calling pattern, one query per
    message: search_service.rs:55-104) are answered in batches and equal the oracle."""
    idx, rows, stored = small
    qs = oracle.make_queries(SEED, 77, 256, len(rows))
    want = [oracle.search_f16(stored, None, q, 20) for q in qs]
    b = dawn.Batcher(idx, max_batch=64, max_wait_us=2000)
    got = [None] * len(qs)

    def worker(t):
        for i in range(t, len(qs), 32):
            got[i] = b.search(qs[i], 20)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(32)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for m, (wl, wd) in zip(got, want):
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
    st = b.stats()
    assert st["queries"] == 256 and st["batches"] < 256 and st["largest_batch"] > 1
    b.close()


def test_batcher_mixed_k_and_full_batches(dawn, small):
    """More callers than max_batch, two different k in flight (a batch holds one k; the others wait for the next batch to
    open), a lone caller afterwards: every answer equals the unbatched dawn_index_search, bit for bit."""
    idx, rows, stored = small
    import oracle.oracle as O

    qs = O.make_queries(SEED, 78, 480, len(rows))
    b = dawn.Batcher(idx, max_batch=16, max_wait_us=500)
    got = [None] * len(qs)

    def worker(t):
        for i in range(t, len(qs), 48):
            got[i] = b.search(qs[i], 20 if (i % 5) else 7)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(48)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i, m in enumerate(got):
        want = idx.search(qs[i], 20 if (i % 5) else 7)
        assert len(m.labels) == len(want.labels)
        assert (m.labels == want.labels).all() and (bits(m.distances) == bits(want.distances)).all()
    st = b.stats()
    assert st["queries"] == len(qs) and st["largest_batch"] <= 16
    lone = b.search(qs[3], 20)  # one caller: answered without waiting for a window to fill
    assert (lone.labels == idx.search(qs[3], 20).labels).all()
    with pytest.raises(dawn.DawnError) as e:  # the index's error reaches the caller of the batch that failed ...
        b.search(qs[3], 100_000)
    assert "DAWN_MAX_K" in str(e.value)
    again = b.search(qs[4], 20)  # ... and the front keeps working
    assert (again.labels == idx.search(qs[4], 20).labels).all()
    b.close()


def test_batcher_over_the_multi_handle(dawn, oracle, small):
    """dawn_batcher_create_multi: single-query callers -> batches -> two shards (on one GPU here) -> device merge."""
    idx, rows, stored = small
    qs = oracle.make_queries(SEED, 79, 96, len(rows))
    with dawn.MultiIndex([0, 0]) as m:
        m.reserve(len(rows))
        m.add_batch(np.arange(1, len(rows) + 1, dtype=np.uint64), rows)
        b = dawn.Batcher(m, max_batch=32, max_wait_us=500)
        got = [None] * len(qs)

        def worker(t):
            for i in range(t, len(qs), 24):
                got[i] = b.search(qs[i], 10)

        threads = [threading.Thread(target=worker, args=(t,)) for t in range(24)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for i, g in enumerate(got):
            wl, wd = oracle.search_f16(stored, None, qs[i], 10)
            assert (g.labels == wl).all() and (bits(g.distances) == bits(wd)).all()
        assert b.stats()["queries"] == len(qs)
        b.close()


def test_distance_limit_across_shards(dawn, oracle, small):
    """dawn_multi_search_batch_limit / dawn_index_search_device_limit: every shard applies the limit before the exchange;
    the merged hits are exactly the oracle's hits with distance < limit, on the scan path and on the tensor-core path."""
    idx, rows, stored = small
    qs = oracle.make_queries(SEED, 80, 40, len(rows))
    with dawn.MultiIndex([0, 0]) as m:
        m.reserve(len(rows))
        m.add_batch(np.arange(1, len(rows) + 1, dtype=np.uint64), rows)
        for force in (1, 2):
            m.set_option("force_path", force)
            full_l, full_d, _ = m.search_batch(qs, 20)
            for limit in (float(np.median(full_d[:, 6])), 0.0, 2.5, float("inf")):
                gl, gd, gc = m.search_batch_limit(qs, 20, limit)
                for i, q in enumerate(qs):
                    wl, wd = oracle.search_f16(stored, None, q, 20)
                    keep = int((wd < limit).sum())
                    assert gc[i] == keep
                    assert (gl[i, :keep] == wl[:keep]).all() and (bits(gd[i, :keep]) == bits(wd[:keep])).all()
        nan_l, nan_d, nan_c = m.search_batch_limit(qs, 20, float("nan"))  # NaN = no limit
        assert (nan_c == 20).all() and (nan_l == full_l).all()


def test_distance_limit_drops_far_hits(small, oracle):
    idx, rows, stored = small
    q = oracle.make_queries(SEED, 5, 1, len(rows))[0]
    full = idx.search(q, 20)
    limit = float(full.distances[7])  # udp_service.rs:196-199: distance >= limit is dropped
    m = idx.search_limit(q, 20, limit)
    assert len(m.labels) == int((full.distances < limit).sum()) <= 7
    assert (m.labels == full.labels[: len(m.labels)]).all()
    assert len(idx.search_limit(q, 20, float("inf")).labels) == 20
    assert len(idx.search_limit(q, 20, -1.0).labels) == 0


@pytest.mark.parametrize("scalar", ["f16", "i8"])
def test_distance_limit_pushed_into_the_scan_is_exact(dawn, oracle, scalar):
    """The limit is a score floor inside the scan kernels: whatever its value, the hits must be exactly the
    oracle's hits with distance < limit (bit-identical, same order), for fp16 and int8 storage."""
    n = 30_000
    rows = oracle.np_synth_rows_f32(SEED + 3, 0, n)
    labels = np.arange(1, n + 1, dtype=np.uint64)
    if scalar == "i8":
        opts = dawn.IndexOptions(quantization=dawn.ScalarKind.I8)
        stored = oracle.store_i8(rows)
        ref = lambda q, k: oracle.search_i8(stored[0], stored[1], labels, q, k)
    else:
        opts = dawn.IndexOptions()
        stored = oracle.store_f16(rows)
        ref = lambda q, k: oracle.search_f16(stored, labels, q, k)
    with dawn.new_index(opts) as idx:
        idx.reserve(n)
        idx.add_batch(labels, rows)
        for q in oracle.make_queries(SEED + 3, 9, 4, n):
            for k in (20, 100):
                wl, wd = ref(q, k)
                limits = [float(wd[0]), float(wd[1]), float(wd[k // 2]), float(wd[-1]),
                          float(np.nextafter(wd[3], np.float32(2))), float(wd[-1]) + 0.5, 0.0]
                for lim in limits:
                    keep = int((wd < np.float32(lim)).sum())
                    m = idx.search_limit(q, k, lim)
                    assert len(m.labels) == keep, (scalar, k, lim)
                    assert (m.labels == wl[:keep]).all() and (bits(m.distances) == bits(wd[:keep])).all()
        # the next unlimited search is not affected by the previous limit
        q = oracle.make_queries(SEED + 3, 10, 1, n)[0]
        idx.search_limit(q, 20, 0.1)
        wl, wd = ref(q, 20)
        m = idx.search(q, 20)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()


def test_i24_wire_query_and_stored_vector(dawn, small, oracle):
    idx, rows, stored = small
    q = oracle.make_queries(SEED, 6, 1, len(rows))[0]
    wire = dawn.encode_i24(q)
    assert len(wire) == 1152
    decoded = dawn.decode_i24(wire)
    m = idx.search_i24(wire, 20)
    wl, wd = oracle.search_f16(stored, None, decoded, 20)  # the peer searches with the DECODED query
    assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
    lim = float(wd[5])
    assert len(idx.search_i24(wire, 20, lim).labels) == int((wd < lim).sum())
    sv = idx.get_i24(17)
    assert sv == oracle.to24(stored[16].astype(np.float32))
    with pytest.raises(dawn.DawnError):
        idx.search_i24(bytes(1152), 20)  # decodes to all -1: not normalised


def test_legacy_emb_bulk_load(dawn, oracle):
    """`.emb` files are arrays of repr(C) PageEntry (src/index/warc.rs:35-43)."""
    n = 3000
    rows = oracle.np_synth_rows_f32(31, 0, n)
    rows[10] *= 3.0  # a record that fails the normalisation gate is skipped
    blob = b"".join(struct.pack("<QQ", i * 7, i * 11) + rows[i].tobytes() + struct.pack("<QQ", 20, 30) for i in range(n))
    assert len(blob) == n * 1568
    with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
        skipped = idx.add_page_entries(blob, first_label=1)
        assert skipped == 1 and idx.size() == n - 1
        keep = np.ones(n, dtype=bool)
        keep[10] = False
        stored = oracle.store_f16(rows[keep])
        labels = (np.arange(n, dtype=np.uint64) + 1)[keep]
        q = oracle.make_queries(31, 32, 1, n)[0]
        m = idx.search(q, 20)
        wl, wd = oracle.search_f16(stored, labels, q, 20)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
