"""dawnsearch_b200 -- B200-native exact top-k search behind DawnSearch's index add/search interface.

Only the hot path lives here: `csrc/` (hand-written sm_100a kernels + the C ABI in
include/dawn_index.h) and `index.py`, a ctypes mirror of the `usearch::ffi::Index` surface the
reference calls (src/search/search_provider.rs).  There is no CPU fallback.
"""
from .index import (EM_LEN, MAX_K, Batcher, DawnError, Index, IndexOptions, Matches, MetricKind, MultiIndex, ScalarKind, ScoreError,
                    decode_i24, encode_i24, is_normalized, load_library, merge_results_device, new_index, normalize)

__all__ = ["EM_LEN", "MAX_K", "Batcher", "decode_i24", "encode_i24", "is_normalized", "normalize", "DawnError", "Index", "IndexOptions", "Matches", "MetricKind", "MultiIndex",
           "ScalarKind", "ScoreError", "load_library", "merge_results_device", "new_index"]
