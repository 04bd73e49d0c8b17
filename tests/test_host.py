"""CPU tests of the host-side logic: the numpy twin of the device generator, shard ranges, the
packed result block, and the N>1 exchange step on world_size-2 gloo (the GPU kernels are
replaced by the oracle here, which is exactly what the parity tests prove they equal)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dawnsearch_b200 import synth
from dawnsearch_b200.sharded import ResultBlock, all_gather_blocks, shard_range


def test_synth_twin_matches_oracle(oracle):
    for seed, first in ((0xDA5EA2C4, 0), (9, 10 ** 9)):
        a = synth.rows_f32(seed, first, 33)
        b = oracle.synth_rows_f32(seed, first, 33)
        assert (a.view(np.uint32) == b.view(np.uint32)).all()
    qa = synth.make_queries(5, 6, 10, 1000)
    qb = oracle.make_queries(5, 6, 10, 1000)
    assert (qa.view(np.uint32) == qb.view(np.uint32)).all()
    assert (synth.planted_rows(6, 10, 1000) == oracle.planted_rows(6, 10, 1000)).all()
    for q in qa:
        assert oracle.is_normalized(q)  # what the reference's gate demands (search_provider.rs:206)


@pytest.mark.parametrize("rows,world", [(100, 1), (100, 8), (7, 8), (0, 4), (100_000_000, 8), (13, 2)])
def test_shard_ranges_partition_the_corpus(rows, world):
    seen = 0
    for r in range(world):
        first, n = shard_range(r, world, rows)
        assert first == min(seen, rows) and n >= 0
        seen += n
    assert seen == rows


def test_result_block_layout():
    blk = ResultBlock(batch=3, k=10)
    assert blk.nbytes % 16 == 0 and blk.nbytes >= 3 * 10 * 12 + 12
    buf = torch.zeros(blk.nbytes, dtype=torch.uint8)
    labels, dists, counts = blk.views(buf)
    labels[:] = torch.arange(30).view(3, 10)
    dists[:] = 0.5
    counts[:] = torch.tensor([10, 9, 0], dtype=torch.int32)
    l2, d2, c2 = blk.views(buf.clone())
    assert (l2 == labels).all() and (d2 == 0.5).all() and c2.tolist() == [10, 9, 0]


def _merge_on_host(labels, dists, counts, k):
    """(distance asc, label asc) merge of per-shard lists: the contract of dawn_merge_results_device."""
    out_l, out_d = [], []
    for b in range(labels.shape[1]):
        lab = np.concatenate([labels[s, b, : counts[s, b]] for s in range(labels.shape[0])])
        dst = np.concatenate([dists[s, b, : counts[s, b]] for s in range(labels.shape[0])])
        order = np.lexsort((lab, dst))[:k]
        out_l.append(lab[order])
        out_d.append(dst[order])
    return out_l, out_d


def _worker(rank, world, port, rows, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O

        seed = 77
        first, n = shard_range(rank, world, rows)
        stored = O.np_synth_rows_f16(seed, first, n)
        labels = np.arange(first + 1, first + n + 1, dtype=np.uint64)
        qs = synth.make_queries(seed, seed + 1, 4, rows)
        blk = ResultBlock(len(qs), k)
        local = torch.zeros(blk.nbytes, dtype=torch.uint8)
        L, D, Cn = blk.views(local)
        for i, q in enumerate(qs):  # the shard-local exact top-k (the GPU kernels' job)
            l, d = O.search_f16(stored, labels, q, k)
            L[i, : len(l)] = torch.from_numpy(l.astype(np.int64))
            D[i, : len(d)] = torch.from_numpy(d)
            Cn[i] = len(l)
        gathered = torch.zeros((world, blk.nbytes), dtype=torch.uint8)
        all_gather_blocks(local, gathered)
        gl = np.stack([blk.views(gathered[s])[0].numpy() for s in range(world)]).astype(np.uint64)
        gd = np.stack([blk.views(gathered[s])[1].numpy() for s in range(world)])
        gc = np.stack([blk.views(gathered[s])[2].numpy() for s in range(world)])
        ml, md = _merge_on_host(gl, gd, gc, k)
        whole = O.np_synth_rows_f16(seed, 0, rows)
        ok = True
        for i, q in enumerate(qs):
            wl, wd = O.search_f16(whole, None, q, k)
            ok &= bool((ml[i] == wl).all() and (md[i].view(np.uint32) == wd.view(np.uint32)).all())
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rows,k", [(3001, 10), (15, 10)])
def test_sharded_exchange_world2_gloo(rows, k):
    """Two ranks, id-range shards, one all-gather of packed blocks, merge == whole-corpus oracle
    (including the ragged case where a shard holds fewer than k rows)."""
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, k, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


# ---- host mirrors of src/search/vector.rs inside libdawn_b200 (no GPU needed) -------------------


def test_library_vector_helpers_match_oracle(dawn, oracle):
    rng = np.random.default_rng(5)
    for _ in range(20):
        v = rng.standard_normal(384).astype(np.float32)
        assert (dawn.normalize(v).view(np.uint32) == oracle.normalize(v).view(np.uint32)).all()
        u = dawn.normalize(v)
        assert dawn.is_normalized(u) and oracle.is_normalized(u)
        for scale in (0.9899, 0.99, 0.9901, 1.0099, 1.01, 1.0101):
            w = (u * np.float32(scale)).astype(np.float32)
            assert dawn.is_normalized(w) == oracle.is_normalized(w)
        data = dawn.encode_i24(u)
        assert data == oracle.to24(u)  # vector.rs:74-86
        back = dawn.decode_i24(data)
        ob, ok = oracle.from24(data)
        assert ok and (back.view(np.uint32) == ob.view(np.uint32)).all()  # vector.rs:52-72
    bad = np.zeros(384, dtype=np.float32)
    assert not dawn.is_normalized(bad)
    with pytest.raises(dawn.DawnError):
        dawn.decode_i24(dawn.encode_i24(bad))  # "Embedding is not normalized"
    nan = np.full(384, np.nan, dtype=np.float32)
    assert not dawn.is_normalized(nan)


def test_library_i24_codec_matches_golden_v2(dawn):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz"))
    for i in range(3):
        wire = g["i24_wire"][i].tobytes()
        assert dawn.encode_i24(g["queries"][i]) == wire
        assert (dawn.decode_i24(wire).view(np.uint32) == g["i24_decoded"][i]).all()
