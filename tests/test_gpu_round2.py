"""Round-2 GPU tests: the exactness hole at N > 1 (certificate flags now travel with the result blocks),
cross-shard duplicate labels in the merge, the norm gate over the device corpus (SearchProvider::verify,
search_provider.rs:289-327), distance_limit on the tensor-core path, the parallel bulk-load pipeline,
load() leaving the index untouched on failure, and re-entrant searches on one handle."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0xDA5EA2C4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# Two ranks, id-range shards, packed all-gather, device merge.  DAWN_TEST_BACKEND=gloo puts both ranks on GPU 0
# (the exchange is staged through the host), so the whole sharded logic -- including the certificate-driven exact
# re-runs -- is exercised on a one-GPU box; with 2+ GPUs the same worker runs over NCCL.
SHARD_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["DAWN_ROOT"])
import numpy as np, torch, torch.distributed as dist
from dawnsearch_b200.sharded import ShardedIndex, shard_range
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
backend = os.environ["DAWN_TEST_BACKEND"]
local = int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo")
torch.cuda.set_device(local)
SEED, rows = 0xDA5EA2C4, 300_007
first, n = shard_range(rank, world, rows)
ok = True
sh = ShardedIndex(local, n)
sh.index.add_synthetic(SEED, first, n)
stored = O.synth_rows_f16(SEED, 0, rows) if rank == 0 else None
for batch, k in ((1, 10), (5, 20), (64, 10), (130, 100)):
    qs = O.make_queries(SEED, 50 + batch, batch, rows)
    gl, gd, cnt = sh.search(qs, k)
    if rank == 0:
        wl, wd, wc, _ = O.cpu_scan_f16(stored, None, qs, k)
        same = (cnt == wc).all() and (gl == wl).all() and (gd.view(np.uint32) == wd.view(np.uint32)).all()
        print("synthetic batch", batch, "k", k, "ok" if same else "MISMATCH", flush=True)
        ok &= bool(same)
    if batch in (5, 64):  # distance_limit across the shards: every shard cuts before the exchange
        limit = 0.78
        ll, ld, lc = sh.search(qs, k, distance_limit=limit)
        if rank == 0:
            same = True
            for i in range(batch):
                keep = int((wd[i, :wc[i]] < limit).sum())
                same &= lc[i] == keep and (ll[i, :keep] == wl[i, :keep]).all() and \
                    (ld[i, :keep].view(np.uint32) == wd[i, :keep].view(np.uint32)).all()
            print("limit batch", batch, "k", k, "ok" if same else "MISMATCH", flush=True)
            ok &= bool(same)
sh.close()
# duplicates: every shard is full of identical pages, so no shard can certify its top-k from the candidate slack
# alone -> each re-runs the query exactly before the merge (the hole round 1 left open at N > 1)
base = O.np_synth_rows_f32(5, 0, 3)
n_dup = 6000
rows_dup = np.concatenate([np.repeat(base[:1], n_dup, axis=0), O.np_synth_rows_f32(6, 0, 4000)])
labels = np.concatenate([np.arange(n_dup, 0, -1), np.arange(10_000, 14_000)]).astype(np.uint64)
perm = np.argsort(O.np_mix64(np.arange(len(labels), dtype=np.uint64) + np.uint64(9)), kind="stable")
rows_dup, labels = rows_dup[perm], labels[perm]
first, n = shard_range(rank, world, len(labels))
sh = ShardedIndex(local, n)
sh.index.add_batch(labels[first:first + n], rows_dup[first:first + n])
stored = O.store_f16(rows_dup)
escalated = 0
for batch, k in ((1, 10), (24, 20), (24, 100)):
    qs = np.concatenate([base[:1], O.make_queries(6, 7, batch - 1, 4000)]) if batch > 1 else base[:1]
    if batch > 1:
        sh.index.set_option("force_path", 2)   # tensor-core path on every shard
    gl, gd, cnt = sh.search(qs, k)
    escalated += sh.last_search_escalated
    if rank == 0:
        wl, wd, wc, _ = O.cpu_scan_f16(stored, labels, qs, k)
        same = (cnt == wc).all() and (gl == wl).all() and (gd.view(np.uint32) == wd.view(np.uint32)).all()
        print("duplicates batch", batch, "k", k, "ok" if same else "MISMATCH", flush=True)
        ok &= bool(same)
        ok &= list(gl[0][:10]) == list(range(1, 11))
    if batch == 24 and k == 20:  # a limit on top of the exact re-runs
        limit = 0.5
        ll, ld, lc = sh.search(qs, k, distance_limit=limit)
        if rank == 0:
            same = True
            for i in range(batch):
                keep = int((wd[i, :wc[i]] < limit).sum())
                same &= lc[i] == keep and (ll[i, :keep] == wl[i, :keep]).all()
            print("limit duplicates ok" if same else "limit duplicates MISMATCH", flush=True)
            ok &= bool(same)
t = torch.tensor([escalated], dtype=torch.int64)
if backend == "nccl":
    t = t.cuda()
dist.all_reduce(t)
if rank == 0:
    print("escalated", int(t.item()), flush=True)
    ok &= int(t.item()) > 0
flag = torch.tensor([1 if ok else 0])
if backend == "nccl":
    flag = flag.cuda()
dist.broadcast(flag, 0)
sh.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
'''


def _run_shard_worker(tmp_path, backend, port):
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER)
    env = dict(os.environ, DAWN_ROOT=ROOT, DAWN_TEST_BACKEND=backend)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == 10, r.stdout
    assert "MISMATCH" not in r.stdout


def test_two_rank_sharded_search_incl_duplicates_on_one_gpu(tmp_path):
    """Runs on ANY box: both ranks on GPU 0, exchange over gloo (host-staged)."""
    _run_shard_worker(tmp_path, "gloo", 29621)


def test_two_gpu_sharded_search_incl_duplicates_over_nccl(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    _run_shard_worker(tmp_path, "nccl", 29622)


def test_merge_drops_a_label_two_shards_both_hold(dawn):
    """BestResults::insert rejects an id it already contains (best_results.rs:46,57); the device merge drops the copy
    from the later shard instead of returning the page twice."""
    import torch

    k, batch = 4, 3
    lab = torch.tensor([[[1, 5, 9, 12], [1, 2, 3, 4], [7, 8, 0, 0]],           # shard 0
                        [[5, 6, 9, 13], [1, 2, 3, 4], [7, 9, 10, 11]]],        # shard 1
                       dtype=torch.int64, device="cuda")
    dst = torch.tensor([[[.1, .2, .3, .4], [.1, .2, .3, .4], [.5, .6, 0, 0]],
                        [[.2, .25, .3, .35], [.1, .2, .3, .4], [.5, .55, .7, .8]]],
                       dtype=torch.float32, device="cuda")
    cnt = torch.tensor([[4, 4, 2], [4, 4, 4]], dtype=torch.int32, device="cuda")
    ol = torch.zeros((batch, k), dtype=torch.int64, device="cuda")
    od = torch.zeros((batch, k), dtype=torch.float32, device="cuda")
    oc = torch.zeros(batch, dtype=torch.int32, device="cuda")
    dawn.merge_results_device(0, lab.data_ptr(), dst.data_ptr(), cnt.data_ptr(), 2, batch, k, ol.data_ptr(),
                              od.data_ptr(), oc.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
    torch.cuda.synchronize()
    assert ol.cpu().tolist() == [[1, 5, 6, 9], [1, 2, 3, 4], [7, 9, 8, 10]]
    assert oc.cpu().tolist() == [4, 4, 4]
    assert np.allclose(od.cpu().numpy()[0], [.1, .2, .25, .3])


def test_verify_applies_the_norm_gate_to_the_device_corpus(dawn, oracle):
    n = 20_000
    rows = oracle.np_synth_rows_f32(11, 0, n)
    labels = np.arange(1, n + 1, dtype=np.uint64)
    for quant in (dawn.ScalarKind.F16, dawn.ScalarKind.I8):
        with dawn.new_index(dawn.IndexOptions(quantization=quant)) as idx:
            idx.reserve(n)
            idx.add_batch(labels, rows)
            v = idx.verify()
            assert v["bad_rows"] == 0 and 0.99 < v["min_norm"] <= v["max_norm"] < 1.01, v


def test_rows_outside_the_norm_gate_are_reported_and_results_stay_exact(dawn, oracle):
    """The raw ABI (like usearch) accepts any vector.  verify() counts stored rows outside (0.99, 1.01), and the
    certificate's bounds are scaled by the longest stored row, so answers are still the oracle's, bit for bit."""
    n = 40_000
    rows = oracle.np_synth_rows_f32(12, 0, n).copy()
    rows[100:160] *= np.float32(1.7)
    rows[5000] *= np.float32(0.5)
    labels = np.arange(1, n + 1, dtype=np.uint64)
    stored = oracle.store_f16(rows)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(labels, rows)
        v = idx.verify()
        assert v["bad_rows"] == 61 and 1.69 < v["max_norm"] < 1.71 and 0.49 < v["min_norm"] < 0.51, v
        qs = np.concatenate([oracle.make_queries(12, 13, 20, n), rows[100:104] / np.float32(1.7)])
        for force in (1, 2):
            idx.set_option("force_path", force)
            gl, gd, cnt = idx.search_batch(qs, 10)
            wl, wd, wc, _ = oracle.cpu_scan_f16(stored, labels, qs, 10)
            assert (gl == wl).all() and (bits(gd) == bits(wd)).all(), force
        assert idx.profile()["uncertified"] == 0


def test_distance_limit_on_the_tensor_core_path(dawn, oracle):
    """f3 on K3: the limit seeds every query's threshold; hits must be exactly the oracle's hits below the limit."""
    n = 300_000
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_synthetic(SEED, 0, n)
        stored = oracle.synth_rows_f16(SEED, 0, n)
        qs = oracle.make_queries(SEED, 21, 40, n)
        idx.set_option("force_path", 2)
        for k in (10, 100):
            wl, wd, wc, _ = oracle.cpu_scan_f16(stored, None, qs, k)
            for lim in (float(np.median(wd[:, k // 2])), float(wd[:, 0].max()) + 1e-4, 0.0, 2.5):
                idx.profile(reset=True)
                gl, gd, cnt = idx.search_batch_limit(qs, k, lim)
                p = idx.profile()
                assert p["gemm_batches"] >= 1 and p["uncertified"] == 0
                for i in range(len(qs)):
                    keep = int((wd[i] < np.float32(lim)).sum())
                    assert cnt[i] == keep, (k, lim, i, cnt[i], keep)
                    assert (gl[i][:keep] == wl[i][:keep]).all() and (bits(gd[i][:keep]) == bits(wd[i][:keep])).all()


@pytest.mark.parametrize("scalar", ["f16", "i8"])
def test_bulk_add_pipeline_keeps_order_and_bits(dawn, oracle, scalar):
    """add_batch of >= 32768 rows takes the parallel pipeline (4 copier threads, 8 pinned buffers, 4 streams);
    rows must land in the caller's order with exactly the bytes the oracle stores."""
    n = 100_003
    rows = np.concatenate([oracle.np_synth_rows_f32(31, i, min(20000, n - i)) for i in range(0, n, 20000)])
    labels = (np.arange(n, dtype=np.uint64) * np.uint64(3) + np.uint64(7))
    quant = dawn.ScalarKind.I8 if scalar == "i8" else dawn.ScalarKind.F16
    with dawn.new_index(dawn.IndexOptions(quantization=quant)) as idx:
        idx.reserve(n + 5)
        idx.add(1, rows[0])                       # a staged single add first: the bulk path must flush it
        idx.add_batch(labels[1:], rows[1:])
        assert idx.size() == n
        qs = oracle.make_queries(31, 32, 6, n)
        labels2 = labels.copy()
        labels2[0] = 1
        if scalar == "i8":
            q8, sc = oracle.store_i8(rows)
            for q in qs:
                m = idx.search(q, 20)
                wl, wd = oracle.search_i8(q8, sc, labels2, q, 20)
                assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
        else:
            stored = oracle.store_f16(rows)
            gl, gd, cnt = idx.search_batch(qs, 20)
            wl, wd, wc, _ = oracle.cpu_scan_f16(stored, labels2, qs, 20)
            assert (gl == wl).all() and (bits(gd) == bits(wd)).all()
            for r in (0, 1, 8191, 8192, 65536, n - 1):  # slice boundaries of the pipeline
                assert (idx.get(int(labels2[r])) == stored[r].astype(np.float32)).all(), r


def test_failed_load_leaves_the_index_unchanged(dawn, oracle, tmp_path):
    n = 12_000
    rows = oracle.np_synth_rows_f32(41, 0, n)
    labels = np.arange(1, n + 1, dtype=np.uint64)
    stored = oracle.store_f16(rows)
    q = oracle.make_queries(41, 42, 1, n)[0]
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(labels[:10_000], rows[:10_000])
        good = str(tmp_path / "good.idx")
        idx.save(good)
        idx.add_batch(labels[10_000:], rows[10_000:])       # staged, not yet flushed
        blob = open(good, "rb").read()
        for name, data in (("short", blob[: len(blob) // 2]), ("magic", b"XXXXXXXX" + blob[8:]), ("empty", b"")):
            bad = str(tmp_path / name)
            open(bad, "wb").write(data)
            with pytest.raises(dawn.DawnError):
                idx.load(bad)
            assert idx.size() == n                           # staged adds survived, nothing was dropped
            m = idx.search(q, 10)
            wl, wd = oracle.search_f16(stored, labels, q, 10)
            assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
        idx.load(good)                                       # a good file replaces everything
        assert idx.size() == 10_000 and idx.capacity() >= n
    # an EMPTY, pre-reserved index (the start-up sequence reserve -> load) is filled in place: no second arena
    with dawn.new_index(dawn.IndexOptions(capacity=50_000)) as idx:
        idx.load(good)
        assert idx.size() == 10_000 and idx.capacity() == 50_000
        m = idx.search(q, 10)
        wl, wd = oracle.search_f16(stored[:10_000], labels[:10_000], q, 10)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
        idx.add_batch(labels[10_000:], rows[10_000:])       # and keeps growing from there
        m = idx.search(q, 10)
        wl, wd = oracle.search_f16(stored, labels, q, 10)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n)
        idx.add_batch(labels[:10_000], rows[:10_000])
        idx.load(good)
        m = idx.search(q, 10)
        wl, wd = oracle.search_f16(stored[:10_000], labels[:10_000], q, 10)
        assert (m.labels == wl).all() and (bits(m.distances) == bits(wd)).all()


def test_concurrent_searches_on_one_handle(dawn, oracle):
    """SURVEY 8(b) threading: search* is re-entrant (per-call stream + workspace).  Four threads hammer one handle with
    different batch shapes (scan path, tensor-core path, single queries) while a fifth appends rows; every answer
    must equal the oracle's over the rows that were visible when the index was first filled."""
    n = 200_000
    extra = 3_000
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(n + extra)
        idx.add_synthetic(SEED, 0, n)
        stored = oracle.synth_rows_f16(SEED, 0, n)
        far = -oracle.np_synth_rows_f32(77, 0, extra)       # appended rows: irrelevant to the queries below? not guaranteed,
        shapes = [(1, 10), (2, 20), (300, 10), (520, 100)]  # so appended rows use labels the check filters out
        # (300 and 520 queries = 2 and 3 query tiles: two tensor-core launches share the SMs here, so the chunk rendezvous of
        #  gemm_pipe.cuh runs into its timeout instead of its partners -- it must only cost time, never block or change results)
        want = {}
        qsets = {}
        for t, (b, k) in enumerate(shapes):
            qsets[t] = oracle.make_queries(SEED, 300 + t, b, n)
            want[t] = oracle.cpu_scan_f16(stored, None, qsets[t], k + 5)
        errors = []

        def worker(t):
            b, k = shapes[t]
            try:
                for _ in range(12):
                    gl, gd, cnt = idx.search_batch(qsets[t], k)
                    wl, wd = want[t][0], want[t][1]
                    for i in range(b):
                        keep = gl[i] <= n                    # drop rows appended meanwhile
                        g_l, g_d = gl[i][keep], gd[i][keep]
                        m = len(g_l)
                        if m < k - 5 or not ((g_l == wl[i][:m]).all() and (bits(g_d) == bits(wd[i][:m])).all()):
                            errors.append((t, i, g_l, wl[i][:m]))
                            return
            except Exception as e:  # noqa: BLE001
                errors.append((t, repr(e)))

        def adder():
            try:
                for i in range(0, extra, 100):
                    idx.add_batch(np.arange(n + 1 + i, n + 101 + i, dtype=np.uint64), far[i:i + 100])
            except Exception as e:  # noqa: BLE001
                errors.append(("adder", repr(e)))

        threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)] + [threading.Thread(target=adder)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        assert not errors, errors[:2]
        assert idx.size() == n + extra
        assert idx.profile()["device_status"] == 0


def test_chunk_rendezvous_on_and_off_give_the_same_bits(dawn, oracle):
    """gemm_unit_sync only changes WHEN the CTA pairs that share a corpus chunk start it (L2 sharing); results are exact either way."""
    n = 400_000
    for quant in (dawn.ScalarKind.F16, dawn.ScalarKind.I8):
        with dawn.new_index(dawn.IndexOptions(capacity=n, quantization=quant)) as idx:
            idx.add_synthetic(SEED, 0, n)
            qs = oracle.make_queries(SEED, 77, 700, n)   # 3 query tiles of 256
            idx.set_option("force_path", 2)
            idx.set_option("i8_tensor_min_batch", 8)
            out = []
            for sync in (1, 0):
                idx.set_option("gemm_unit_sync", sync)
                out.append(idx.search_batch(qs, 10))
            assert (out[0][0] == out[1][0]).all() and (bits(out[0][1]) == bits(out[1][1])).all()
            if quant == dawn.ScalarKind.F16:
                stored = oracle.synth_rows_f16(SEED, 0, n)
                wl, wd, wc, _ = oracle.cpu_scan_f16(stored, None, qs[:64], 10)
                assert (out[0][0][:64] == wl).all() and (bits(out[0][1][:64]) == bits(wd)).all()
            p = idx.profile()
            assert p["gemm_batches"] == 2 and p["uncertified"] == 0


def test_device_api_counts_uncertified_results(dawn, oracle):
    """dawn_index_search_device only enqueues and cannot escalate; the finalize kernel counts what it could not
    certify on the device, so a device-resident caller (bench.py's `value` leg) can report it."""
    import torch

    base = oracle.np_synth_rows_f32(5, 0, 1)
    rows = np.repeat(base, 5000, axis=0)
    labels = np.arange(1, 5001, dtype=np.uint64)
    with dawn.new_index(dawn.IndexOptions()) as idx:
        idx.reserve(5000)
        idx.add_batch(labels, rows)
        q = torch.from_numpy(np.repeat(base, 3, axis=0)).cuda()
        k = 10
        ol = torch.zeros((3, k), dtype=torch.int64, device="cuda")
        od = torch.zeros((3, k), dtype=torch.float32, device="cuda")
        oc = torch.zeros(3, dtype=torch.int32, device="cuda")
        of = torch.zeros(3, dtype=torch.int32, device="cuda")
        idx.profile(reset=True)
        idx.search_device(q.data_ptr(), 3, k, ol.data_ptr(), od.data_ptr(), oc.data_ptr(), of.data_ptr(),
                          torch.cuda.current_stream().cuda_stream or 1)
        torch.cuda.synchronize()
        assert of.cpu().tolist() == [0, 0, 0]                # 5000 exact ties: the slack cannot certify
        assert idx.profile()["device_uncertified"] == 3
        assert ol.cpu().numpy()[0].tolist() == list(range(1, 11))   # ...yet ties still break on the label
