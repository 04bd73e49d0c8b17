"""CPU tests of the oracle (oracle/ is test infrastructure; PARITY UNPINNED by the reference).

The reference has no tests (/root/reference/.github/workflows/build.yml:32), so what is
checked here is: (1) the C oracle against an independent numpy restatement and against the
committed golden vectors, (2) the runtime invariants the reference enforces in
src/search/vector.rs, src/search/best_results.rs and src/net/web.rs.
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------ formats


def test_f32_to_f16_matches_ieee_rne(oracle):
    rng = np.random.default_rng(1)
    xs = [rng.standard_normal(20000).astype(np.float32) * s for s in (1, 1e-2, 1e-4, 6e-5, 1e-7, 3e4, 1e5)]
    xs.append(np.array([0.0, -0.0, 65504, 65519.996, 65520, 1e9, -1e9, np.inf, -np.inf, 2.0 ** -24,
                        2.0 ** -25, 2.0 ** -25 * 1.0001, 2.0 ** -14, 6.1e-5, 1 + 2.0 ** -11,
                        1 + 2.0 ** -11 + 2.0 ** -20, 1 + 3 * 2.0 ** -11], dtype=np.float32))
    x = np.concatenate(xs)
    pad = (-len(x)) % 384
    x = np.concatenate([x, np.zeros(pad, dtype=np.float32)]).reshape(-1, 384)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    got = oracle.store_f16(x).view(np.uint16)
    assert (got == want).all()


def test_f16_to_f32_all_patterns(oracle):
    L = oracle.lib()
    h = np.arange(65536, dtype=np.uint16)
    want = h.view(np.float16).astype(np.float32)
    got = np.array([L.dawn_oracle_f16_to_f32(int(v)) for v in h], dtype=np.float32)
    ok = (bits(got) == bits(want)) | np.isnan(want)
    assert ok.all()


def test_synth_c_equals_numpy(oracle):
    for seed, first in ((0xDA5EA2C4, 0), (7, 123456789), (2 ** 63 + 5, 2 ** 33)):
        a = oracle.synth_rows_f32(seed, first, 40)
        b = oracle.np_synth_rows_f32(seed, first, 40)
        assert (bits(a) == bits(b)).all()
        h = oracle.synth_rows_f16(seed, first, 40).view(np.uint16)
        assert (h == b.astype(np.float16).view(np.uint16)).all()
    norms = np.linalg.norm(oracle.np_synth_rows_f32(1, 0, 200).astype(np.float64), axis=1)
    assert np.abs(norms - 1).max() < 1e-6


def test_i8_store(oracle):
    rows = oracle.np_synth_rows_f32(3, 0, 16)
    q, s = oracle.store_i8(rows)
    amax = np.abs(rows).max(axis=1)
    assert (s == (amax / np.float32(127)).astype(np.float32)).all()
    assert np.abs(q).max() == 127
    deq = q.astype(np.float32) * s[:, None]
    assert np.abs(deq - rows).max() <= s.max() / 2 * 1.0001


# ------------------------------------------------------ vector.rs invariants


def test_distance_functions_follow_vector_rs(oracle):
    rng = np.random.default_rng(2)
    a = oracle.normalize(rng.standard_normal(384))
    b = oracle.normalize(rng.standard_normal(384))
    acc = np.float32(0)
    for i in range(384):  # vector.rs:128-134, one rounding per op
        acc = np.float32(acc + np.float32(a[i] * b[i]))
    assert bits(oracle.distance_ip(a, b)) == bits(acc)
    assert bits(oracle.distance_cosine(a, b)) == bits(np.float32(1) - acc)
    # L2^2 = 2 - 2 dot on unit vectors (vector.rs:95-97 vs 99-101): same ranking
    assert abs(oracle.distance_l2sq(a, b) - (2 - 2 * float(acc))) < 1e-5
    assert oracle.is_normalized(a) and oracle.is_normalized(b)


def test_is_normalized_gate(oracle):
    v = np.zeros(384, dtype=np.float32)
    v[0] = 1.0
    assert oracle.is_normalized(v)
    assert not oracle.is_normalized(v * 0.98)      # vector.rs:185-192: (0.99, 1.01) exclusive
    assert not oracle.is_normalized(v * 1.02)
    assert oracle.is_normalized(v * 0.995) and oracle.is_normalized(v * 1.005)
    w = v.copy()
    w[1] = np.nan
    assert not oracle.is_normalized(w)
    w[1] = np.inf
    assert not oracle.is_normalized(w)
    assert not oracle.is_normalized(np.zeros(384, dtype=np.float32))


def test_i16_quantisation(oracle):
    L = oracle.lib()
    # vector.rs:30-32: round-half-away-from-zero of x*32767, saturating
    for x, want in ((0.0, 0), (1.0, 32767), (-1.0, -32767), (0.5, 16384), (-0.5, -16384),
                    (1.5 / 32767, 2), (-1.5 / 32767, -2), (2.0, 32767), (-2.0, -32768)):
        assert L.dawn_oracle_f32_to_i16(x) == want, x


def test_i24_wire_codec_roundtrip(oracle):
    # vector.rs:48-87: a normalised vector must survive to24 -> from24 and stay normalised
    for seed in range(5):
        v = oracle.np_synth_rows_f32(seed, 0, 1)[0]
        data = oracle.to24(v)
        assert len(data) == 1152  # udp_packets.rs:35-38
        back, ok = oracle.from24(data)
        assert ok
        assert np.abs(back - v).max() < 3e-7
    v = np.zeros(384, dtype=np.float32)
    v[0] = -1.0  # encodes to 0
    assert oracle.to24(v)[:3] == b"\x00\x00\x00"
    v[0] = 1.0
    assert oracle.to24(v)[:3] == b"\xff\xff\x7f"


def test_best_results_semantics(oracle):
    # best_results.rs:44-65
    b = oracle.BestResults(3)
    assert b.worst_distance() == 0.0  # quirk: stays 0 until full
    assert b.insert(1, 0.5) and b.insert(2, 0.3)
    assert not b.insert(1, 0.1)  # dedupe by id while filling
    assert b.worst_distance() == 0.0
    assert b.insert(3, 0.7)
    assert b.worst_distance() == pytest.approx(0.7)
    assert not b.insert(4, 0.7)  # strict <
    assert b.insert(4, 0.6)      # replaces worst (id 3)
    assert not b.insert(2, 0.0)  # already present
    b.sort()
    assert [i for i, _ in b.results()] == [2, 1, 4]
    d = [x for _, x in b.results()]
    assert d == sorted(d)


# ------------------------------------------------------------- the hot path


@pytest.mark.parametrize("k", [1, 10, 20, 100])
def test_c_oracle_equals_numpy_restatement(oracle, k):
    n = 3000
    stored = oracle.np_synth_rows_f16(11, 0, n)
    labels = (np.argsort(oracle.np_mix64(np.arange(n, dtype=np.uint64))) + 5).astype(np.uint64)
    qs = oracle.make_queries(11, 12, 6, n)
    for q in qs:
        l1, d1 = oracle.search_f16(stored, labels, q, k)
        l2, d2 = oracle.np_search(stored.astype(np.float32), labels, q, k)
        assert (l1 == l2).all() and (bits(d1) == bits(d2)).all()


def test_golden_vectors(oracle):
    g = np.load(GOLDEN)
    seed, n = int(g["seed"]), int(g["n"])
    rows = oracle.synth_rows_f32(seed, 0, n)
    assert (bits(rows[:4]) == bits(g["rows_f32_head"])).all()
    assert (bits(rows[-2:]) == bits(g["rows_f32_tail"])).all()
    stored = oracle.store_f16(rows)
    assert (stored[:4].view(np.uint16) == g["stored_f16_head"]).all()
    assert (stored[-2:].view(np.uint16) == g["stored_f16_tail"]).all()
    for k in (1, 10, 20, 100):
        for i, q in enumerate(g["queries"]):
            l, d = oracle.search_f16(stored, g["labels"], q, k)
            assert (l == g[f"labels_k{k}"][i]).all()
            assert (bits(d) == g[f"dist_k{k}"][i]).all()


def test_golden_v2_int8_storage_and_i24_codec(oracle):
    """tests/golden/golden_v2.npz (numpy restatements, make_golden_v2.py) pins the C oracle's int8 storage,
    int8 search and i24 wire codec."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz"))
    n = int(g["n"])
    rows = oracle.np_synth_rows_f32(int(g["seed"]), 0, n)
    q8, sc = oracle.store_i8(rows)
    assert (q8[:3] == g["i8_head"]).all() and (sc[:3].view(np.uint32) == g["scales_head"]).all()
    assert int(q8.astype(np.int64).sum()) == int(g["i8_checksum"])
    assert int(sc.view(np.uint32).astype(np.uint64).sum()) == int(g["scales_checksum"])
    q8n, scn = oracle.np_store_i8(rows)
    assert (q8 == q8n).all() and (sc.view(np.uint32) == scn.view(np.uint32)).all()
    for k in (1, 10, 20, 100):
        for i, q in enumerate(g["queries"]):
            l, d = oracle.search_i8(q8, sc, g["labels"], q, k)
            assert (l == g[f"labels_k{k}"][i]).all() and (bits(d) == g[f"dist_k{k}"][i]).all()
    for i in range(3):
        wire = oracle.to24(g["queries"][i])
        assert wire == g["i24_wire"][i].tobytes() == oracle.np_to24(g["queries"][i])
        dec, ok = oracle.from24(wire)
        assert ok and (bits(dec) == g["i24_decoded"][i]).all()
        assert (bits(oracle.np_from24(wire)) == g["i24_decoded"][i]).all()


def test_reference_invariants_on_search(oracle):
    n = 2000
    rows = oracle.np_synth_rows_f32(21, 0, n)
    stored = rows.astype(np.float16)
    # self query: rank 0, distance < 0.001 (src/net/web.rs:330-343)
    for r in (0, 17, n - 1):
        l, d = oracle.search_f16(stored, None, rows[r], 20)
        assert l[0] == r + 1 and d[0] < 1e-3
        assert (np.diff(d) >= 0).all()          # ascending (best_results.rs:71-79)
        assert len(l) == 20                      # count <= k (search_provider.rs:214)
    # k > N: count == N
    l, d = oracle.search_f16(stored[:7], None, rows[0], 20)
    assert len(l) == 7
    # empty corpus
    l, d = oracle.search_f16(stored[:0], None, rows[0], 20)
    assert len(l) == 0


def test_ties_break_on_lower_label(oracle):
    row = oracle.np_synth_rows_f16(5, 0, 1)
    stored = np.repeat(row, 50, axis=0)
    labels = np.arange(50, 0, -1).astype(np.uint64) * 3  # descending labels
    l, d = oracle.search_f16(stored, labels, row[0].astype(np.float32), 10)
    assert list(l) == sorted(labels.tolist())[:10]
    assert len(set(bits(d).tolist())) == 1


def test_f32_and_i8_variants(oracle):
    n = 1500
    rows = oracle.np_synth_rows_f32(31, 0, n)
    q = oracle.make_queries(31, 32, 1, n)[0]
    l1, d1 = oracle.search_f32(rows, None, q, 10)
    l2, d2 = oracle.np_search(rows, None, q, 10)
    assert (l1 == l2).all() and (bits(d1) == bits(d2)).all()
    qi8, sc = oracle.store_i8(rows)
    l3, d3 = oracle.search_i8(qi8, sc, None, q, 10)
    l4, d4 = oracle.np_search(qi8.astype(np.float32), None, q, 10, row_scale=sc)
    assert (l3 == l4).all() and (bits(d3) == bits(d4)).all()
    # the three storage precisions agree on the planted neighbour
    lf16, _ = oracle.search_f16(rows.astype(np.float16), None, q, 1)
    assert l1[0] == l3[0] == lf16[0]


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_cpu_scan_equals_scalar_oracle(oracle, threads):
    n = 40000
    stored = oracle.synth_rows_f16(41, 0, n)
    labels = (np.argsort(oracle.np_mix64(np.arange(n, dtype=np.uint64) + np.uint64(3))) + 1).astype(np.uint64)
    qs = oracle.make_queries(41, 42, 5, n)
    for k in (1, 10, 100):
        lo, do, cnt, cert = oracle.cpu_scan_f16(stored, labels, qs, k, threads=threads)
        for i, q in enumerate(qs):
            l, d = oracle.search_f16(stored, labels, q, k)
            assert cnt[i] == len(l)
            assert (lo[i, : cnt[i]] == l).all() and (bits(do[i, : cnt[i]]) == bits(d)).all()


def test_cpu_scan_duplicates_fall_back_exactly(oracle):
    row = oracle.np_synth_rows_f16(5, 0, 1)
    stored = np.repeat(row, 3000, axis=0)
    labels = np.arange(3000, 0, -1).astype(np.uint64)
    lo, do, cnt, cert = oracle.cpu_scan_f16(stored, labels, row.astype(np.float32), 10, threads=4)
    assert list(lo[0]) == list(range(1, 11))


def test_hnsw_restatement_behaves_like_an_ann_index(oracle):
    """The stand-in for the reference's USearch index (hnsw_restatement.c, NOT USearch 0.22.3):
    approximate, ascending distances = 1 - dot, recall rising to 1 with efSearch."""
    n = 3000
    rows = oracle.np_synth_rows_f32(3, 0, n)
    h = oracle.Hnsw(16, 128, 64, 1)
    h.add_batch(np.arange(1, n + 1), rows)
    assert h.size() == n
    qs = oracle.make_queries(3, 4, 40, n)
    truth = [oracle.search_f32(rows, None, q, 10) for q in qs]
    recalls = []
    for ef in (16, 64, 2000):
        h.set_ef_search(ef)
        hit = 0
        for q, (tl, td) in zip(qs, truth):
            l, d = h.search(q, 10)
            assert len(l) == 10 and (np.diff(d) >= 0).all()
            hit += len(set(l.tolist()) & set(tl.tolist()))
        recalls.append(hit / 400)
    assert recalls[0] <= recalls[1] <= recalls[2] and recalls[2] > 0.99
    l, d = h.search(rows[7], 20)  # self query (src/net/web.rs:339)
    assert l[0] == 8 and abs(d[0]) < 1e-3


# ---- property tests (hypothesis): C oracle == numpy restatement == threaded scan on arbitrary small inputs
from hypothesis import HealthCheck, given, settings, strategies as st


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 300), k=st.integers(1, 120), seed=st.integers(0, 2 ** 32 - 1), dup=st.booleans(),
       label_mode=st.sampled_from(["rowid", "huge", "repeated"]))
def test_property_oracles_agree(oracle, n, k, seed, dup, label_mode):
    rng = np.random.default_rng(seed)
    rows = oracle.np_synth_rows_f32(seed, 0, n)
    if dup and n > 2:
        rows[rng.integers(0, n, size=n // 2)] = rows[rng.integers(0, n)]
    if label_mode == "rowid":
        labels = np.arange(1, n + 1, dtype=np.uint64)
    elif label_mode == "huge":
        labels = rng.integers(1, 2 ** 63, size=n, dtype=np.int64).astype(np.uint64)
    else:
        labels = rng.integers(1, max(2, n // 2), size=n, dtype=np.int64).astype(np.uint64)
    stored = rows.astype(np.float16)
    q = oracle.make_queries(seed, seed ^ 0xABCD, 1, n)[0]
    l1, d1 = oracle.search_f16(stored, labels, q, k)
    l2, d2 = oracle.np_search(stored.astype(np.float32), labels, q, k)
    lo, do, cnt, _ = oracle.cpu_scan_f16(stored, labels, q[None, :], k, threads=3)
    assert len(l1) == min(k, n)
    assert (l1 == l2).all() and (bits(d1) == bits(d2)).all()
    assert cnt[0] == len(l1) and (lo[0, : cnt[0]] == l1).all() and (bits(do[0, : cnt[0]]) == bits(d1)).all()
    assert (np.diff(d1) >= 0).all()
    # ties are ordered by ascending label
    for i in range(len(l1) - 1):
        if bits(d1[i:i + 1])[0] == bits(d1[i + 1:i + 2])[0]:
            assert l1[i] <= l1[i + 1]
