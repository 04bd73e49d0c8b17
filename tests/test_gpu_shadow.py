"""Option "shadow_i8": an fp16 corpus keeps an int8 copy of itself; batches that would take the fp16 tensor-core path are
FILTERED on the copy by the int8 tensor cores (gemm_i8.cu in shadow mode) and every candidate is re-scored on the fp16
rows.  The filter's band is a rigorous bound (query quantisation + the rows' own quantisation error, measured when the copy is
made), so labels and distances must stay bit-identical to the oracle's exact scan of the fp16 rows -- whatever the data."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 0xDA5EA2C4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def check(oracle, stored, labels, qs, k, got):
    gl, gd, gc = got
    wl, wd, wc, _ = oracle.cpu_scan_f16(stored, labels, qs, k)
    assert (gc == wc).all()
    assert (gl == wl).all()
    assert (bits(gd) == bits(wd)).all()


@pytest.fixture(scope="module")
def shadow300k(dawn, oracle):
    n = 300_000
    idx = dawn.new_index(dawn.IndexOptions(capacity=n))
    idx.add_synthetic(SEED, 0, n)
    idx.set_option("shadow_i8", 1)
    idx.set_option("shadow_big_k_rows", 0)  # take the shadow path whatever k (by default k > 32 waits for 40M rows)
    yield idx, n, oracle.synth_rows_f16(SEED, 0, n)
    idx.close()


@pytest.mark.parametrize("batch,k", [(16, 10), (129, 100), (300, 20), (1024, 10), (520, 1)])
def test_shadow_path_matches_oracle(dawn, oracle, shadow300k, batch, k):
    idx, n, stored = shadow300k
    qs = oracle.make_queries(SEED, 40 + batch, batch, n)
    idx.profile(reset=True)
    got = idx.search_batch(qs, k)
    p = idx.profile(reset=True)
    assert p["shadow_batches"] == 1 and p["gemm_batches"] == 1 and p["uncertified"] == 0
    assert p["max_selection_error"] == 0.0  # the rounds rank exact scores
    check(oracle, stored, None, qs, k, got)
    # and equals what the fp16 tensor-core path answers
    idx.set_option("shadow_i8", 0)
    plain = idx.search_batch(qs, k)
    idx.set_option("shadow_i8", 1)
    assert idx.profile(reset=True)["shadow_batches"] == 0
    assert (plain[0] == got[0]).all() and (bits(plain[1]) == bits(got[1])).all()


def test_shadow_with_distance_limit(dawn, oracle, shadow300k):
    idx, n, stored = shadow300k
    qs = oracle.make_queries(SEED, 41, 200, n)
    wl, wd, wc, _ = oracle.cpu_scan_f16(stored, None, qs, 20)
    for limit in (float(np.median(wd[:, 5])), 0.3, 1.5):
        gl, gd, gc = idx.search_batch_limit(qs, 20, limit)
        for i in range(len(qs)):
            keep = int((wd[i] < limit).sum())
            assert gc[i] == keep
            assert (gl[i, :keep] == wl[i, :keep]).all() and (bits(gd[i, :keep]) == bits(wd[i, :keep])).all()


def test_large_k_on_a_small_shard_stays_on_the_fp16_tiles(dawn, oracle, shadow300k):
    idx, n, stored = shadow300k
    qs = oracle.make_queries(SEED, 43, 200, n)
    idx.set_option("shadow_big_k_rows", 40_000_000)
    idx.profile(reset=True)
    got100 = idx.search_batch(qs, 100)       # compute-bound batch, k > 32, small shard: fp16 tiles
    got100_small = idx.search_batch(qs[:64], 100)  # one query tile: HBM-bound, the shadow's 388 B per row win
    got20 = idx.search_batch(qs, 20)
    p = idx.profile(reset=True)
    idx.set_option("shadow_big_k_rows", 0)
    assert p["gemm_batches"] == 3 and p["shadow_batches"] == 2
    check(oracle, stored, None, qs, 100, got100)
    check(oracle, stored, None, qs[:64], 100, got100_small)
    check(oracle, stored, None, qs, 20, got20)


def test_shadow_follows_appends_growth_and_load(dawn, oracle, tmp_path):
    """The copy is extended lazily after appends, rebuilt after the arena is reallocated, and rebuilt after a load."""
    n = 140_000
    rows = oracle.np_synth_rows_f32(SEED + 5, 0, n)
    labels = (np.arange(n, dtype=np.uint64) * 3 + 7)
    stored = oracle.store_f16(rows)
    qs = oracle.make_queries(SEED + 5, 9, 150, n)
    with dawn.new_index(dawn.IndexOptions(capacity=80_000)) as idx:
        idx.set_option("shadow_i8", 1)
        idx.add_batch(labels[:70_000], rows[:70_000])
        idx.profile(reset=True)
        check(oracle, stored[:70_000], labels[:70_000], qs, 10, idx.search_batch(qs, 10))
        idx.add_batch(labels[70_000:80_000], rows[70_000:80_000])  # same arena: the copy grows by 10,000 rows
        check(oracle, stored[:80_000], labels[:80_000], qs, 10, idx.search_batch(qs, 10))
        assert idx.profile(reset=True)["shadow_batches"] == 2
        idx.reserve(n)                                             # new arena: the copy is dropped and rebuilt
        idx.add_batch(labels[80_000:], rows[80_000:])
        idx.profile(reset=True)
        check(oracle, stored, labels, qs, 20, idx.search_batch(qs, 20))
        assert idx.profile(reset=True)["shadow_batches"] == 1
        path = str(tmp_path / "shadow.idx")
        idx.save(path)
    with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
        idx.set_option("shadow_i8", 1)
        idx.load(path)
        idx.profile(reset=True)
        check(oracle, stored, labels, qs, 10, idx.search_batch(qs, 10))
        assert idx.profile(reset=True)["shadow_batches"] == 1
    os.remove(path)


def test_shadow_adversarial_rows(dawn, oracle):
    """Rows the int8 copy represents badly (a few large components and many small ones, queries living in the small ones),
    exact duplicates, and near-duplicates that differ below the int8 resolution: the band must still contain every true hit."""
    rng = np.random.default_rng(11)
    n = 90_000
    rows = oracle.np_synth_rows_f32(SEED + 6, 0, n)
    spiky = rng.normal(size=(3000, 384)).astype(np.float32) * 0.02
    spiky[:, :2] = np.array([0.9, 0.4], dtype=np.float32)
    spiky /= np.linalg.norm(spiky, axis=1, keepdims=True)
    rows[1000:4000] = spiky
    rows[5000:9000] = rows[4999]                      # 4001 exact duplicates
    near = np.repeat(rows[9500:9501], 2000, axis=0) + rng.normal(size=(2000, 384)).astype(np.float32) * 2e-4
    rows[10_000:12_000] = near / np.linalg.norm(near, axis=1, keepdims=True)
    labels = np.arange(1, n + 1, dtype=np.uint64)
    stored = oracle.store_f16(rows)
    tail = rows[1000:1300].copy()
    tail[:, :2] = 0
    tail /= np.linalg.norm(tail, axis=1, keepdims=True)
    qs = np.concatenate([tail, rows[4999:5000], rows[9500:9501], rows[10_000:10_040],
                         oracle.make_queries(SEED + 6, 3, 58, n)]).astype(np.float32)
    with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
        idx.set_option("shadow_i8", 1)
        idx.set_option("shadow_big_k_rows", 0)
        idx.add_batch(labels, rows)
        for k in (10, 100):
            idx.profile(reset=True)
            got = idx.search_batch(qs, k)
            assert idx.profile(reset=True)["shadow_batches"] >= 1
            check(oracle, stored, labels, qs, k, got)
