"""The exactness certificate rests on bounds eps on |selection score - exact score|.  This test DERIVES them instead of
trusting them: it measures the worst error of the tensor-core kernel over more than 1e9 (query,row)
pairs -- isotropic, same-sign (worst case for accumulation: every product positive), near-duplicate (scores ~ 1) and
norm-edge inputs -- and of the scan kernels over every candidate they hand to the re-score, and asserts that every
constant compiled into the library exceeds the observed maximum at least twofold.  The histogram goes to
gpurun_out/r02_slack_histogram.json (committed under profiles/)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def test_tensor_core_accumulation_slack_is_derived_from_a_billion_pairs(dawn, oracle):
    rng = np.random.default_rng(20261017)
    n = 1024
    sets = {
        "isotropic": (unit(rng.standard_normal((n, 384))), lambda b: unit(rng.standard_normal((b, 384))), 130),
        "same_sign": (unit(np.abs(rng.standard_normal((n, 384)))), lambda b: unit(np.abs(rng.standard_normal((b, 384)))), 90),
        "norm_edge_same_sign": (unit(np.abs(rng.standard_normal((n, 384)))) * np.float32(1.0095),
                                lambda b: unit(np.abs(rng.standard_normal((b, 384)))) * np.float32(1.0095), 30),
    }
    report = {}
    total = {"pairs": 0, "max_mma_vs_f64": 0.0, "max_seq_vs_f64": 0.0, "max_mma_vs_seq": 0.0, "max_err_over_eps_q": 0.0,
             "hist": [0] * 40}
    for name, (rows, make_q, batches) in sets.items():
        acc = dawn.ScoreError()
        with dawn.new_index(dawn.IndexOptions(capacity=n)) as idx:
            idx.add_batch(np.arange(1, n + 1, dtype=np.uint64), rows)
            for b in range(batches):
                qs = make_q(4096)
                if name == "isotropic" and b % 4 == 0:  # near-duplicates of stored rows: scores close to 1
                    qs[:1024] = unit(rows + 0.02 * rng.standard_normal((n, 384)).astype(np.float32))
                idx.debug_gemm_score_error(qs, acc)
        d = acc.as_dict()
        report[name] = d
        assert d["pairs"] == batches * 4096 * n
        assert d["max_err_over_eps_q"] < 0.5, (name, d)
        total["pairs"] += d["pairs"]
        total["hist"] = [a + b for a, b in zip(total["hist"], d["hist"])]
        for key in ("max_mma_vs_f64", "max_seq_vs_f64", "max_mma_vs_seq", "max_err_over_eps_q"):
            total[key] = max(total[key], d[key])
        total["gemm_accum_slack"] = d["gemm_accum_slack"]
    t = total
    report["all"] = t
    assert t["pairs"] > 1_000_000_000
    # the slack has to cover the MMA's accumulation error AND the rounding of the sequential re-score, with margin 2
    need = t["max_mma_vs_f64"] + t["max_seq_vs_f64"]
    assert t["gemm_accum_slack"] >= 2.0 * need, (t["gemm_accum_slack"], need)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        report["bins"] = "hist[0]: |diff| == 0; hist[b]: |tcgen05 score - sequential f32 score| in [2^(b-40), 2^(b-39))"
        with open(os.path.join(out_dir, "r02_slack_histogram.json"), "w") as f:
            json.dump(report, f, indent=1)


@pytest.mark.parametrize("scalar", ["f16", "i8"])
def test_scan_eps_exceeds_twice_the_observed_selection_error(dawn, oracle, scalar):
    """K2 / K4 hand k' candidates with their scan scores to the re-score; the finalize kernel records the largest
    |scan score - exact score| it ever sees.  Same-sign rows and queries make every product positive (largest partial sums)."""
    rng = np.random.default_rng(7)
    n = 100_000
    worst = 0.0
    for kind in ("isotropic", "same_sign"):
        rows = rng.standard_normal((n, 384))
        rows = unit(np.abs(rows) if kind == "same_sign" else rows)
        quant = dawn.ScalarKind.I8 if scalar == "i8" else dawn.ScalarKind.F16
        with dawn.new_index(dawn.IndexOptions(capacity=n, quantization=quant)) as idx:
            idx.add_batch(np.arange(1, n + 1, dtype=np.uint64), rows)
            idx.set_option("force_path", 1)
            idx.set_option("i8_tensor_min_batch", 0)
            idx.profile(reset=True)
            for b in range(6):
                qs = rng.standard_normal((256, 384))
                qs = unit(np.abs(qs) if kind == "same_sign" else qs)
                idx.search_batch(qs, 100)  # k' = 128 candidates per query
            p = idx.profile()
            assert p["scan_launches"] > 0 and p["gemm_batches"] == 0
            worst = max(worst, p["max_selection_error"])
    acc = dawn.ScoreError()
    with dawn.new_index(dawn.IndexOptions(capacity=1024)) as idx:
        idx.add_batch(np.arange(1, 1025, dtype=np.uint64), unit(rng.standard_normal((1024, 384))))
        idx.debug_gemm_score_error(unit(rng.standard_normal((8, 384))), acc)  # only to read the compiled-in constants
    if scalar == "f16":
        assert 0 < worst and acc.scan_eps >= 2.0 * worst, (acc.scan_eps, worst)
    else:
        # the int8 scan's eps is per query: 1.03 * |q - s1 hi - s2 lo| + 3e-5; the observed error must leave the constant part twofold margin
        assert 0 < worst and 3.0e-5 >= 2.0 * worst - 2.0 * 1.2e-5, worst
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"r02_scan_selection_error_{scalar}.json"), "w") as f:
            json.dump({"scalar": scalar, "max_abs_scan_minus_exact": worst, "scan_eps_f16": acc.scan_eps,
                       "candidates_checked": 2 * 6 * 256 * 128}, f)
