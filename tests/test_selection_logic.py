"""CPU checks of the two pieces of selection logic the CUDA path relies on (numpy restatements of the
algorithms, not of the kernels):

1. finalize.cu:select_lists -- the k'-th best of the lists' first ceil(k'/n_lists) entries is a lower bound on the
   k'-th best overall, so filtering by it and ranking the survivors yields exactly the top-k' of the union,
   ties included.
2. The exactness certificate: candidates are selected by an approximate score that is within
   eps of the exact one; if 1 - (weakest kept approximate score + eps) is strictly above the k-th exact distance,
   the k results ARE the exact top-k.  Whenever the certificate holds the answer must be exact; it may fail to
   hold (then the library escalates), but it must never hold for a wrong answer.
"""
import numpy as np
import pytest


def cand_order(scores, labels, rows):
    """indices sorted by (score desc, label asc, row asc) -- cand_better in dawn_common.cuh"""
    return np.lexsort((rows, labels, -scores.astype(np.float64)))


def select_lists(lists, kp):
    """lists: [n_lists][kp] of (score, label, row) sorted by cand_order, padded with row = -1."""
    n_lists = len(lists)
    m = -(-kp // n_lists)
    probe = np.array([lists[l][j][0] if lists[l][j][2] >= 0 else -np.inf for l in range(n_lists) for j in range(m)])
    bound = np.sort(probe)[::-1][kp - 1] if len(probe) >= kp else -np.inf
    surv = [c for lst in lists for c in lst if c[2] >= 0 and c[0] >= bound]
    s = np.array([c[0] for c in surv]); lab = np.array([c[1] for c in surv]); row = np.array([c[2] for c in surv])
    order = cand_order(s, lab, row)[:kp]
    return [surv[i] for i in order], len(surv)


@pytest.mark.parametrize("n_lists,kp,ties", [(148, 16, False), (148, 32, True), (148, 128, False), (3, 128, True),
                                             (7, 32, False), (2, 16, True), (148, 128, True)])
def test_select_lists_equals_top_kp_of_the_union(n_lists, kp, ties):
    rng = np.random.default_rng(n_lists * 1000 + kp)
    for trial in range(6):
        rows_per_list = rng.integers(0, 3 * kp, size=n_lists)  # some lists short or empty
        lists, union = [], []
        next_row = 0
        for l in range(n_lists):
            n = int(rows_per_list[l])
            sc = rng.standard_normal(n).astype(np.float32)
            if ties:
                sc = np.round(sc * 4) / 4  # many equal scores
            lab = rng.integers(0, 2 ** 40, size=n)  # duplicates are as good as impossible, and rows break them anyway
            row = np.arange(next_row, next_row + n)
            next_row += n
            order = cand_order(sc, lab, row)[:kp]
            lst = [(float(sc[i]), int(lab[i]), int(row[i])) for i in order]
            union += lst
            lst += [(-np.inf, 2 ** 63, -1)] * (kp - len(lst))
            lists.append(lst)
        got, n_surv = select_lists(lists, kp)
        s = np.array([c[0] for c in union]); lab = np.array([c[1] for c in union]); row = np.array([c[2] for c in union])
        want = [union[i] for i in cand_order(s, lab, row)[:kp]] if union else []
        assert got == want
        assert n_surv >= min(kp, len(union))


def certificate_run(rng, n, k, kp, eps, gap_scale):
    exact = (rng.standard_normal(n) * gap_scale).astype(np.float32)
    approx = (exact + rng.uniform(-eps, eps, n).astype(np.float32) * np.float32(0.999)).astype(np.float32)
    labels = rng.permutation(n).astype(np.int64)
    rows = np.arange(n)
    keep = cand_order(approx, labels, rows)[:kp]                     # what the scan / GEMM epilogue delivers
    scan_min = approx[keep].min() if len(keep) == kp else None
    dist = (np.float32(1.0) - exact[keep]).astype(np.float32)         # re-scored exactly
    order = np.lexsort((rows[keep], labels[keep], dist))[:k]
    result = keep[order]
    kth = dist[order][-1]
    certified = True if scan_min is None else bool(np.float32(1.0) - (scan_min + np.float32(eps)) > kth)
    all_dist = (np.float32(1.0) - exact).astype(np.float32)
    truth = np.lexsort((rows, labels, all_dist))[:k]
    return certified, bool((result == truth).all())


@pytest.mark.parametrize("k,kp", [(1, 16), (10, 16), (20, 32), (100, 128)])
def test_certificate_never_holds_for_a_wrong_answer(k, kp):
    rng = np.random.default_rng(k * 7 + kp)
    eps = 3e-5
    seen = {"certified": 0, "uncertified": 0, "uncertified_and_wrong": 0}
    for trial in range(300):
        # gap_scale sweeps from "scores far apart" to "scores closer together than eps"
        gap_scale = 10.0 ** rng.uniform(-6.5, -1.0)
        certified, exact = certificate_run(rng, 3000, k, kp, eps, gap_scale)
        if certified:
            assert exact, (trial, gap_scale)
            seen["certified"] += 1
        else:
            seen["uncertified"] += 1
            seen["uncertified_and_wrong"] += 0 if exact else 1
    # the sweep must actually exercise both outcomes, including selections the certificate rightly refused
    assert seen["certified"] > 20 and seen["uncertified"] > 20 and seen["uncertified_and_wrong"] > 0, seen


# 3. gemm_topk.cu: rounds of tiles visited in a strided permutation; a round appends every score >= the k'-th best
#    of all earlier rounds to a per-query log; a select keeps the best k' and publishes the new threshold.  The
#    survivors of the last select must be the global top-k' by (score desc, row asc) whatever the order of the rows,
#    unless the log overflowed (which the kernel flags and the host answers with the exact scan).
def rounds_select(scores, kp, tile=256, growth=8, log_cap=2048, sequential=False):
    n = len(scores)
    total_tiles = -(-n // tile)
    mult = int(total_tiles * 0.6180339887498949) | 1
    from math import gcd
    while mult > 1 and gcd(mult, total_tiles) != 1:
        mult += 2
    if total_tiles <= 2 or sequential:
        mult = 1
    mult %= total_tiles
    mult = mult or 1
    kept = np.zeros(0, dtype=np.int64)
    thr = -np.inf
    overflow = False
    begin, end = 0, max(1, 1024 // tile)
    while begin < total_tiles:
        if end > total_tiles or end + end // 4 > total_tiles:
            end = total_tiles
        rows = []
        for t in range(begin, end):
            phys = (t * mult) % total_tiles
            rows.append(np.arange(phys * tile, min(n, (phys + 1) * tile)))
        rows = np.concatenate(rows)
        passed = rows[scores[rows] >= thr]
        if len(kept) + len(passed) > log_cap:
            overflow = True
            passed = passed[: log_cap - len(kept)]
        log = np.concatenate([kept, passed])
        order = np.lexsort((log, -scores[log].astype(np.float64)))[:kp]
        kept = log[order]
        thr = scores[kept[-1]] if len(kept) == kp else -np.inf
        begin, end = end, end * growth
    return kept, overflow


@pytest.mark.parametrize("kp,growth", [(16, 8), (32, 8), (128, 4), (16, 32)])
def test_rounds_with_thresholds_keep_the_global_top_kp(kp, growth):
    rng = np.random.default_rng(kp + growth)
    for n in (5, 1000, 1024, 70_000, 300_001):
        scores = rng.standard_normal(n).astype(np.float32)
        kept, overflow = rounds_select(scores, kp, growth=growth)
        assert not overflow
        want = np.lexsort((np.arange(n), -scores.astype(np.float64)))[:kp]
        assert (kept == want).all(), (n, kp, growth)


def test_adversarial_row_order_overflows_only_without_the_permutation():
    """The near neighbours of the query all sit at the end of the corpus (pages crawled last): in stored order the
    last round meets thousands of rows above a threshold learnt from unrelated pages and the log overflows (the
    kernel flags the query and the host re-runs it through the exact scan); visiting the tiles in the strided
    permutation makes every round a sample of the whole corpus and the log stays small.
    (A corpus sorted by similarity to the query from start to end can still overflow -- whole tiles beat the
    threshold at once -- which is what the overflow flag and the exact-scan fallback are for.)"""
    rng = np.random.default_rng(3)
    n, kp = 400_000, 128
    scores = rng.standard_normal(n).astype(np.float32)
    scores[-4000:] += np.float32(10.0)
    _, overflow_seq = rounds_select(scores, kp, growth=4, sequential=True)
    kept, overflow_perm = rounds_select(scores, kp, growth=4)
    assert overflow_seq and not overflow_perm
    want = np.lexsort((np.arange(n), -scores.astype(np.float64)))[:kp]
    assert (kept == want).all()
