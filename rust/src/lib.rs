//! dawn-index-b200: device-resident exact top-k index for DawnSearch on NVIDIA B200.
//! `index::gpu_index` mirrors `usearch::ffi` method for method (see that file's header).
pub mod index {
    pub mod gpu_index;
}
pub use index::gpu_index::{new_index, Batcher, Index, IndexOptions, Matches, MetricKind, MultiIndex, ScalarKind};
