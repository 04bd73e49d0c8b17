"""CPU model of the int8-shadow filter (gemm_i8.cu in shadow mode, DESIGN.md "K4d") -- numpy restatements of the arithmetic,
not of the kernels:

1. The band.  For a query q (f32), its one-level int8 form (s1, hi), an fp16 row x16 and its int8 copy (s_row, x8) as
   shadow_quantize_kernel makes it:  |q.x16 - s1 s_row HI| <= e1 + |q| kappa s_row  with  HI = hi.x8 (exact integer),
   e1 = 1.10 |q - s1 hi| + 3e-5  and  kappa = max_rows |x16 / s_row - x8|.  Hence the epilogue's test
   s_row (HI + c_q) >= (T - e1) / s1,  c_q = |q| kappa / s1,  never drops a row whose exact score reaches T.
2. The two-stage select.  Entries get a fast score within eta of the exact one; with tau the k'-th best fast score, an entry
   below tau - 2 eta is not among the k' best by exact score, so re-scoring only the others exactly loses nothing.

Random rows, spiky rows (a few large components), near-duplicates and queries aligned with the small components."""
import numpy as np
import pytest

from oracle import oracle as O

DIM = 384


def quantize_rows(x16):
    """shadow_quantize_kernel: scale = absmax / 127 (f32 division), x8 = rint(x / scale) clamped, kappa measured on t = x / scale."""
    x = x16.astype(np.float32)
    amax = np.abs(x).max(axis=1)
    s = np.where(amax > 0, (amax / np.float32(127.0)).astype(np.float32), np.float32(1.0)).astype(np.float32)
    t = (x / s[:, None]).astype(np.float32)
    x8 = np.clip(np.rint(t), -127, 127).astype(np.int32)
    kappa = float(np.sqrt(((t - x8.astype(np.float32)) ** 2).sum(axis=1).astype(np.float32)).max()) * 1.00001
    return s, x8, kappa * 1.0001 + 1e-4  # what ensure_shadow stores


def quantize_query(q):
    """prep_queries_i8_gemm_kernel in shadow mode."""
    amax = np.abs(q).max()
    s1 = np.float32(amax / np.float32(127.0)) if amax > 0 else np.float32(1.0)
    hi = np.clip(np.rint(q / s1), -127, 127).astype(np.int32)
    delta = float(np.sqrt(((q.astype(np.float64) - float(s1) * hi) ** 2).sum()))
    e1 = delta * 1.10 + 3.0e-5
    qn = float(np.sqrt((q.astype(np.float64) ** 2).sum())) * 1.00001
    return s1, hi, e1, qn


def corpus(kind, n, rng):
    rows = O.np_synth_rows_f32(77, 0, n)
    if kind == "spiky":
        rows = rng.normal(size=(n, DIM)).astype(np.float32) * 0.02
        rows[:, :3] = rng.normal(size=(n, 3)).astype(np.float32) + 0.5
    elif kind == "near_duplicates":
        rows = np.repeat(rows[:1], n, axis=0) + rng.normal(size=(n, DIM)).astype(np.float32) * 3e-4
    elif kind == "one_hot_ish":
        rows = rng.normal(size=(n, DIM)).astype(np.float32) * 1e-3
        rows[np.arange(n), rng.integers(0, DIM, n)] = 1.0
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    return O.store_f16(rows)  # fp16 values as the index stores them


@pytest.mark.parametrize("kind", ["gaussian", "spiky", "near_duplicates", "one_hot_ish"])
def test_band_contains_every_exact_score(kind):
    rng = np.random.default_rng(5)
    x16 = corpus(kind, 4000, rng)
    s_row, x8, kappa = quantize_rows(x16)
    assert kappa <= np.sqrt(DIM) / 2 * 1.001 + 2e-4  # the theoretical ceiling of the measured constant
    x64 = x16.astype(np.float64)
    queries = [O.np_synth_rows_f32(78, i, 1)[0] for i in range(6)]
    small = x16[0].astype(np.float32).copy()
    small[np.argsort(-np.abs(small))[:4]] = 0  # a query living in the row's small components
    queries.append((small / np.linalg.norm(small)).astype(np.float32))
    queries.append(x16[1].astype(np.float32))
    worst = 0.0
    for q in queries:
        s1, hi, e1, qn = quantize_query(q)
        HI = x8 @ hi                                            # exact integers, as the int8 tensor cores give them
        a = float(s1) * s_row.astype(np.float64) * HI           # the approximate score
        exact = x64 @ q.astype(np.float64)                      # real dot product; the f32 sequential sum is within 2.4e-5 of it
        band = e1 - 2.4e-5 + qn * kappa * s_row.astype(np.float64)
        assert (np.abs(exact - a) <= band).all(), (kind, float(np.abs(exact - a).max()))
        worst = max(worst, float((np.abs(exact - a) / band).max()))
        # the epilogue's form of the same test, in f32 like the kernel: a row with exact >= T is never dropped
        T = np.sort(exact)[-10]
        thr = np.float32((np.float32(T) - np.float32(e1)) / s1)
        thr = np.float32(thr - abs(thr) * np.float32(4.0e-7))
        cq = np.float32(qn * kappa / float(s1))
        f = (HI.astype(np.float32) + cq) * s_row
        assert (f[exact >= T] >= thr).all(), kind
    assert worst < 1.0


@pytest.mark.parametrize("kp,n", [(16, 300), (16, 2048), (104, 1500), (32, 40)])
def test_two_stage_select_keeps_the_exact_top_kp(kp, n):
    rng = np.random.default_rng(kp * 7 + n)
    eta = 6.0e-5
    for trial in range(20):
        exact = rng.normal(size=n).astype(np.float64) * 0.05
        if trial % 3 == 0:  # a crowd of near-ties around the k'-th score
            m = min(3 * kp, n)
            exact[:m] = np.sort(exact)[-min(kp, n)] + rng.uniform(-1.5 * eta, 1.5 * eta, size=m)
        fast = exact + rng.uniform(-eta, eta, size=n)
        kept = rng.random(n) < 0.05
        fast[kept] = exact[kept]  # entries kept from the previous select already carry exact scores
        if n >= kp:
            tau = np.sort(fast)[-kp]
            marked = fast >= tau - 2 * eta
        else:
            marked = np.ones(n, dtype=bool)
        want = set(np.argsort(-exact, kind="stable")[: min(kp, n)].tolist())
        assert want <= set(np.nonzero(marked)[0].tolist())
        # and ranking the marked entries by exact score gives exactly the top-k'
        idx = np.nonzero(marked)[0]
        got = set(idx[np.argsort(-exact[idx], kind="stable")[: min(kp, n)]].tolist())
        assert got == want
