"""Multi-GPU layer: the corpus is sharded by contiguous id range, one process per GPU.

The reference already answers a query this way across WAN peers: scatter the query, every
instance returns its local top-k, the caller merges by distance
(/root/reference/src/net/udp_service.rs:314-330, src/search/search_service.rs:201-277).  On
one NVSwitch box the same shape becomes: local exact top-k on every GPU (libdawn_b200) ->
ONE all-gather of k (label, distance) pairs per query over NCCL/NVLink -> device-side merge
(dawn_merge_results_device).  Because every shard returns bit-exact distances, the merged
result is bit-identical to a single index holding the whole corpus.

torch is plumbing here (device buffers, streams, torch.distributed); all compute is in
libdawn_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .index import EM_LEN, Index, IndexOptions, merge_results_device, new_index


class ShardedIndex:
    """One rank's shard plus the collective search.  All ranks must call `search*` together."""

    def __init__(self, device: int, capacity: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device
        self.tdev = torch.device("cuda", device)
        self.index: Index = new_index(IndexOptions(device=device, capacity=capacity))
        self._ws = {}

    def close(self):
        self.index.close()

    def _workspace(self, batch: int, k: int):
        key = (batch, k)
        ws = self._ws.get(key)
        if ws is None:
            d = self.tdev
            ws = {
                "q": torch.empty((batch, EM_LEN), dtype=torch.float32, device=d),
                "labels": torch.empty((batch, k), dtype=torch.int64, device=d),
                "dist": torch.empty((batch, k), dtype=torch.float32, device=d),
                "counts": torch.empty(batch, dtype=torch.int32, device=d),
                "flags": torch.empty(batch, dtype=torch.int32, device=d),
                "g_labels": torch.empty((self.world, batch, k), dtype=torch.int64, device=d),
                "g_dist": torch.empty((self.world, batch, k), dtype=torch.float32, device=d),
                "g_counts": torch.empty((self.world, batch), dtype=torch.int32, device=d),
                "o_labels": torch.empty((batch, k), dtype=torch.int64, device=d),
                "o_dist": torch.empty((batch, k), dtype=torch.float32, device=d),
                "o_counts": torch.empty(batch, dtype=torch.int32, device=d),
                "h_q": torch.empty((batch, EM_LEN), dtype=torch.float32).pin_memory(),
                "h_labels": torch.empty((batch, k), dtype=torch.int64).pin_memory(),
                "h_dist": torch.empty((batch, k), dtype=torch.float32).pin_memory(),
                "h_counts": torch.empty(batch, dtype=torch.int32).pin_memory(),
            }
            self._ws[key] = ws
        return ws

    def search_device(self, d_queries: torch.Tensor, k: int):
        """Queries already on this rank's GPU ([B,384] f32).  Enqueues local search, all-gather and
        merge on the current stream; returns device tensors (labels int64, distances, counts)."""
        batch = d_queries.shape[0]
        ws = self._workspace(batch, k)
        stream = torch.cuda.current_stream(self.tdev).cuda_stream
        self.index.search_device(d_queries.data_ptr(), batch, k, ws["labels"].data_ptr(), ws["dist"].data_ptr(),
                                 ws["counts"].data_ptr(), ws["flags"].data_ptr(), stream)
        if self.world == 1:
            return ws["labels"], ws["dist"], ws["counts"]
        dist.all_gather_into_tensor(ws["g_labels"], ws["labels"], group=self.group)
        dist.all_gather_into_tensor(ws["g_dist"], ws["dist"], group=self.group)
        dist.all_gather_into_tensor(ws["g_counts"], ws["counts"], group=self.group)
        merge_results_device(self.device, ws["g_labels"].data_ptr(), ws["g_dist"].data_ptr(),
                             ws["g_counts"].data_ptr(), self.world, batch, k, ws["o_labels"].data_ptr(),
                             ws["o_dist"].data_ptr(), ws["o_counts"].data_ptr(), stream)
        return ws["o_labels"], ws["o_dist"], ws["o_counts"]

    def search(self, queries: np.ndarray, k: int):
        """Host queries in, host results out (every rank gets the full merged answer)."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, EM_LEN)
        batch = q.shape[0]
        ws = self._workspace(batch, k)
        with torch.cuda.device(self.tdev):
            ws["h_q"].copy_(torch.from_numpy(q))
            ws["q"].copy_(ws["h_q"], non_blocking=True)
            labels, dists, counts = self.search_device(ws["q"], k)
            ws["h_labels"].copy_(labels, non_blocking=True)
            ws["h_dist"].copy_(dists, non_blocking=True)
            ws["h_counts"].copy_(counts, non_blocking=True)
            torch.cuda.current_stream(self.tdev).synchronize()
        return (ws["h_labels"].numpy().astype(np.uint64), ws["h_dist"].numpy().copy(),
                ws["h_counts"].numpy().astype(np.int64))
