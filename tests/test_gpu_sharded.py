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
