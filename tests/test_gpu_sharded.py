"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU under
torchrun, id-range shards, one packed NCCL all-gather, device merge -- the merged answer must be
bit-identical to the oracle over the whole corpus."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["DAWN_ROOT"])
import numpy as np, torch, torch.distributed as dist
from dawnsearch_b200.sharded import ShardedIndex, shard_range
from oracle import oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
SEED, rows = 0xDA5EA2C4, 1_000_003
first, n = shard_range(rank, world, rows)
sh = ShardedIndex(local, n)
sh.index.add_synthetic(SEED, first, n)
stored = O.synth_rows_f16(SEED, 0, rows) if rank == 0 else None
ok = True
for batch, k in ((1, 10), (4, 20), (64, 10), (300, 100)):
    qs = O.make_queries(SEED, 50 + batch, batch, rows)
    gl, gd, cnt = sh.search(qs, k)
    if rank == 0:
        wl, wd, wc, _ = O.cpu_scan_f16(stored, None, qs, k)
        same = (cnt == wc).all() and (gl == wl).all() and (gd.view(np.uint32) == wd.view(np.uint32)).all()
        print("batch", batch, "k", k, "ok" if same else "MISMATCH", flush=True)
        ok &= bool(same)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
sh.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
'''


def test_two_gpu_sharded_search_matches_oracle(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DAWN_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == 4, r.stdout


def _multi_devices():
    import torch

    n = torch.cuda.device_count()
    return [0, 1] if n >= 2 else [0, 0]  # two shards on one GPU still exercise the whole exchange


def test_single_process_multi_index_matches_oracle(dawn, oracle):
    """dawn_multi_*: one process, one shard per device, peer-copy gather + device merge."""
    import numpy as np

    n = 30_001
    rows = oracle.np_synth_rows_f32(71, 0, n)
    labels = (np.argsort(oracle.np_mix64(np.arange(n, dtype=np.uint64) + np.uint64(3)), kind="stable") + 10).astype(np.uint64)
    stored = oracle.store_f16(rows)
    with dawn.MultiIndex(_multi_devices()) as m:
        assert m.shards() == 2
        m.reserve(n + 10)
        m.add_batch(labels[:20000], rows[:20000])
        for i in range(20000, 20010):  # single adds, like SearchProvider::insert
            m.add(int(labels[i]), rows[i])
        m.add_batch(labels[20010:], rows[20010:])
        assert m.size() == n
        for batch, k in ((1, 20), (5, 10), (40, 100)):
            qs = oracle.make_queries(71, 72 + batch, batch, n)
            gl, gd, cnt = m.search_batch(qs, k)
            wl, wd, wc, _ = oracle.cpu_scan_f16(stored, labels, qs, k)
            assert (cnt == wc).all() and (gl == wl).all()
            assert (gd.view(np.uint32) == wd.view(np.uint32)).all()
        with pytest.raises(dawn.DawnError):
            m.add_batch(np.arange(100, dtype=np.uint64), np.zeros((100, 384), dtype=np.float32))  # beyond capacity


def test_single_process_multi_index_synthetic_and_duplicates(dawn, oracle):
    import numpy as np

    n = 200_000
    with dawn.MultiIndex(_multi_devices()) as m:
        m.reserve(n)
        m.add_synthetic(0xDA5EA2C4, 0, n)
        stored = oracle.synth_rows_f16(0xDA5EA2C4, 0, n)
        qs = oracle.make_queries(0xDA5EA2C4, 5, 24, n)
        gl, gd, cnt = m.search_batch(qs, 10)
        wl, wd, wc, _ = oracle.cpu_scan_f16(stored, None, qs, 10)
        assert (gl == wl).all() and (gd.view(np.uint32) == wd.view(np.uint32)).all()
    base = oracle.np_synth_rows_f32(5, 0, 1)
    rows = np.repeat(base, 2000, axis=0)  # every shard full of ties: uncertified -> exact re-run per shard
    labels = np.arange(2000, 0, -1).astype(np.uint64)
    with dawn.MultiIndex(_multi_devices()) as m:
        m.reserve(2000)
        m.add_batch(labels, rows)
        r = m.search(base[0], 10)
        assert list(r.labels) == list(range(1, 11))


PIPE_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["DAWN_ROOT"])
import numpy as np, torch, torch.distributed as dist
from dawnsearch_b200.sharded import ShardedIndex, shard_range
from oracle import oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
SEED, rows = 0xDA5EA2C4, 400_000
first, n = shard_range(rank, world, rows)
sh = ShardedIndex(local, n)
sh.index.add_synthetic(SEED, first, n)
stored = O.synth_rows_f16(SEED, 0, rows) if rank == 0 else None
ok = True
k, batch, steps = 10, 32, 5
qs = [O.make_queries(SEED, 90 + i, batch, rows) for i in range(steps)]
dq = [torch.from_numpy(q).cuda() for q in qs]
outs = []
for i in range(steps):  # back-to-back pipelined batches; copy each result out as soon as it is complete
    blk = sh.search_device(dq[i], k, pipelined=True)
    sh.wait_results()
    outs.append(blk.clone())
torch.cuda.synchronize()
if rank == 0:
    from dawnsearch_b200.sharded import ResultBlock
    rb = ResultBlock(batch, k)
    for i in range(steps):
        L, D, Cn = rb.views(outs[i].cpu())
        wl, wd, wc, _ = O.cpu_scan_f16(stored, None, qs[i], k)
        same = (L.numpy().astype(np.uint64) == wl).all() and (D.numpy().view(np.uint32) == wd.view(np.uint32)).all()
        print("step", i, "ok" if same else "MISMATCH", flush=True)
        ok &= bool(same)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
sh.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
'''


def test_two_gpu_pipelined_exchange_matches_oracle(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "pipe_worker.py"
    script.write_text(PIPE_WORKER)
    env = dict(os.environ, DAWN_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == 5, r.stdout
