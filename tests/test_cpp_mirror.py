"""The compiled-language host mirror (include/dawn_index.hpp): dawn::ffi::Index with the method
names of usearch::ffi::Index, dawn::SearchProvider with search_provider.rs's flow, BestResults."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "dawnsearch_b200", "lib")


@pytest.fixture(scope="module")
def demo(tmp_path_factory, dawn):
    out = str(tmp_path_factory.mktemp("cpp") / "provider_demo")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "provider_demo.cpp"), "-o", out, "-L", LIBDIR,
                    "-ldawn_b200", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return out


def test_cpp_mirror_builds_and_refuses_without_gpu(demo):
    import torch

    if torch.cuda.is_available():
        pytest.skip("refusal path is for hosts without a GPU")
    r = subprocess.run([demo, "--no-gpu-check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "refused code=-2" in r.stdout and "no CPU fallback" in r.stdout


def test_cpp_host_mirrors_match_oracle(demo, oracle, tmp_path):
    """to24 / from24 / is_normalized / BestResults in the C++ mirror against the oracle's restatement of
    src/search/vector.rs:48-87,185-192 and best_results.rs:44-79 (no GPU involved)."""
    vs = oracle.make_queries(9, 10, 3, 1000)
    vs.tofile(tmp_path / "v.f32")
    r = subprocess.run([demo, "--host-only", str(tmp_path / "v.f32")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    lines = r.stdout.splitlines()
    for i in range(3):
        wire = bytes.fromhex([l for l in lines if l.startswith(f"wire {i} ")][0].split()[2])
        assert wire == oracle.to24(vs[i])
        back = np.array([int(x) for x in [l for l in lines if l.startswith(f"back {i}")][0].split()[2:]], dtype=np.uint32)
        ob, ok = oracle.from24(wire)
        assert ok and (back == ob.view(np.uint32)).all()
        assert f"norm {i} 1" in lines
    # BestResults(3): ids 0..5 with id 4 re-submitted as id 1 (duplicate id is rejected), sorted ascending
    b = oracle.BestResults(3)
    d = [0.5, 0.25, 0.75, 0.125, 0.25, 0.9]
    for i in range(6):
        b.insert(1 if i == 4 else i, d[i])
    b.sort()
    want = "best " + " ".join(f"{i}:{dist:g}" for i, dist in b.results()) + f" worst={b.worst_distance():g}"
    assert want in lines, (want, lines[-1])


@pytest.mark.gpu
def test_cpp_search_provider_matches_oracle(demo, oracle, tmp_path):
    n, nq = 3000, 6
    rows = oracle.np_synth_rows_f32(61, 0, n)
    qs = oracle.make_queries(61, 62, nq, n)
    rows.tofile(tmp_path / "rows.f32")
    qs.tofile(tmp_path / "q.f32")
    r = subprocess.run([demo, str(tmp_path / "rows.f32"), str(n), str(tmp_path / "q.f32"), str(nq)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    stored = oracle.store_f16(rows)
    got = {}
    for line in r.stdout.splitlines():
        f = line.split()
        if f[0] == "like":
            assert f[1] == "5" and f[2] == "1"
            continue
        got.setdefault(int(f[0]), []).append((int(f[1]), int(f[2]), f[3]))
    for q in range(nq):
        wl, wd = oracle.search_f16(stored, None, qs[q], 20)  # the reference's k (search_provider.rs:214)
        assert [g[0] for g in got[q]] == wl.tolist()
        assert [g[1] for g in got[q]] == wd.view(np.uint32).tolist()
        assert got[q][0][2] == f"https://example.org/{wl[0] - 1}"  # rowids start at 1
