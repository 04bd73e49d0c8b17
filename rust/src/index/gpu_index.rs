//! gpu_index.rs -- drop-in for `usearch::ffi::{new_index, Index, IndexOptions, ...}` backed by
//! libdawn_b200 (include/dawn_index.h).  UNCOMPILED IN THIS REPO (no Rust toolchain in the image).
//!
//! `src/search/search_provider.rs` changes exactly one line:
//!     -use usearch::ffi::{new_index, Index, IndexOptions, MetricKind, ScalarKind};
//!     +use crate::index::gpu_index::{new_index, Index, IndexOptions, MetricKind, ScalarKind};
//! Every method keeps the name, argument meaning and `Result` behaviour of the cxx bridge the
//! reference calls at search_provider.rs:102,115,117,133,149,178,214,221,246,280-284.
use std::ffi::{c_char, c_int, c_void, CStr, CString};

#[repr(C)]
struct DawnOptions { dimensions: u32, metric: u32, scalar: u32, device: i32, capacity: u64, flags: u32, reserved: u32 }

extern "C" {
    fn dawn_last_error() -> *const c_char;
    fn dawn_index_create(opts: *const DawnOptions, out: *mut *mut c_void) -> c_int;
    fn dawn_index_free(idx: *mut c_void);
    fn dawn_index_reserve(idx: *mut c_void, n: usize) -> c_int;
    fn dawn_index_add(idx: *mut c_void, label: u64, v: *const f32) -> c_int;
    fn dawn_index_add_batch(idx: *mut c_void, labels: *const u64, v: *const f32, n: usize) -> c_int;
    fn dawn_index_search(idx: *mut c_void, q: *const f32, k: usize, labels: *mut u64, dist: *mut f32, count: *mut usize) -> c_int;
    fn dawn_index_search_batch(idx: *mut c_void, q: *const f32, batch: usize, k: usize, labels: *mut u64, dist: *mut f32, counts: *mut usize) -> c_int;
    fn dawn_index_size(idx: *const c_void) -> usize;
    fn dawn_index_capacity(idx: *const c_void) -> usize;
    fn dawn_index_dimensions(idx: *const c_void) -> usize;
    fn dawn_index_save(idx: *mut c_void, path: *const c_char) -> c_int;
    fn dawn_index_load(idx: *mut c_void, path: *const c_char) -> c_int;
    // callers and formats either side of the path
    fn dawn_index_get(idx: *mut c_void, label: u64, out384: *mut f32) -> c_int;
    fn dawn_index_get_i24(idx: *mut c_void, label: u64, out1152: *mut u8) -> c_int;
    fn dawn_index_search_limit(idx: *mut c_void, q: *const f32, k: usize, distance_limit: f32, labels: *mut u64, dist: *mut f32, count: *mut usize) -> c_int;
    fn dawn_index_search_i24(idx: *mut c_void, q1152: *const u8, k: usize, has_limit: c_int, distance_limit: f32, labels: *mut u64, dist: *mut f32, count: *mut usize) -> c_int;
    fn dawn_batcher_create(idx: *mut c_void, max_batch: usize, max_wait_us: u32, out: *mut *mut c_void) -> c_int;
    fn dawn_batcher_search(b: *mut c_void, q: *const f32, k: usize, labels: *mut u64, dist: *mut f32, count: *mut usize) -> c_int;
    fn dawn_batcher_free(b: *mut c_void);
}

#[derive(Clone, Copy)] pub enum MetricKind { IP }
#[derive(Clone, Copy)] pub enum ScalarKind { F32, F16, F8 }   // F32 is accepted and stored as F16 on the device

pub struct IndexOptions {          // same fields as usearch::ffi::IndexOptions (search_provider.rs:35-42)
    pub dimensions: usize, pub metric: MetricKind, pub quantization: ScalarKind,
    pub connectivity: usize, pub expansion_add: usize, pub expansion_search: usize,
}

pub struct Matches { pub labels: Vec<u64>, pub distances: Vec<f32> }   // search_provider.rs:221

pub struct Index { h: *mut c_void }
unsafe impl Send for Index {}      // the library serialises calls on a handle internally

fn err() -> anyhow::Error {
    anyhow::anyhow!(unsafe { CStr::from_ptr(dawn_last_error()) }.to_string_lossy().into_owned())
}
fn ck(rc: c_int) -> anyhow::Result<()> { if rc == 0 { Ok(()) } else { Err(err()) } }

pub fn new_index(o: &IndexOptions) -> anyhow::Result<Box<Index>> {       // search_provider.rs:102
    let scalar = match o.quantization { ScalarKind::F8 => 1, _ => 0 };   // F8 -> int8 storage (388 B / page), F32/F16 -> fp16
    let opts = DawnOptions { dimensions: o.dimensions as u32, metric: 0, scalar, device: 0, capacity: 0, flags: 0, reserved: 0 };
    let mut h = std::ptr::null_mut();
    ck(unsafe { dawn_index_create(&opts, &mut h) })?;
    Ok(Box::new(Index { h }))
}

impl Index {
    pub fn reserve(&self, capacity: usize) -> anyhow::Result<()> { ck(unsafe { dawn_index_reserve(self.h, capacity) }) }      // :133,282
    pub fn add(&self, label: u64, vector: &[f32]) -> anyhow::Result<()> {                                                     // :149,284
        anyhow::ensure!(vector.len() == 384, "vector must have 384 dimensions");
        ck(unsafe { dawn_index_add(self.h, label, vector.as_ptr()) })
    }
    pub fn add_batch(&self, labels: &[u64], vectors: &[f32]) -> anyhow::Result<()> {      // bulk rebuild (fill_index_from_db, :127-153)
        anyhow::ensure!(vectors.len() == labels.len() * 384, "vectors must be labels.len() x 384");
        ck(unsafe { dawn_index_add_batch(self.h, labels.as_ptr(), vectors.as_ptr(), labels.len()) })
    }
    pub fn search(&self, query: &[f32], count: usize) -> anyhow::Result<Matches> {                                            // :214
        anyhow::ensure!(query.len() == 384, "query must have 384 dimensions");
        let (mut labels, mut distances, mut n) = (vec![0u64; count], vec![0f32; count], 0usize);
        ck(unsafe { dawn_index_search(self.h, query.as_ptr(), count, labels.as_mut_ptr(), distances.as_mut_ptr(), &mut n) })?;
        labels.truncate(n); distances.truncate(n);
        Ok(Matches { labels, distances })
    }
    pub fn search_batch(&self, queries: &[f32], count: usize) -> anyhow::Result<Vec<Matches>> {   // for a batching front-end
        let b = queries.len() / 384;
        let (mut labels, mut distances, mut counts) = (vec![0u64; b * count], vec![0f32; b * count], vec![0usize; b]);
        ck(unsafe { dawn_index_search_batch(self.h, queries.as_ptr(), b, count, labels.as_mut_ptr(), distances.as_mut_ptr(), counts.as_mut_ptr()) })?;
        Ok((0..b).map(|i| Matches { labels: labels[i * count..i * count + counts[i]].to_vec(),
                                    distances: distances[i * count..i * count + counts[i]].to_vec() }).collect())
    }
    /// `UdpPacket::Search { distance_limit, .. }` (src/net/udp_packets.rs:29-39): hits with `distance >= limit` are not
    /// returned (src/net/udp_service.rs:196-199); the limit is pushed down into the scan kernels.
    pub fn search_limit(&self, query: &[f32], count: usize, distance_limit: Option<f32>) -> anyhow::Result<Matches> {
        anyhow::ensure!(query.len() == 384, "query must have 384 dimensions");
        let (mut labels, mut distances, mut n) = (vec![0u64; count], vec![0f32; count], 0usize);
        ck(unsafe { dawn_index_search_limit(self.h, query.as_ptr(), count, distance_limit.unwrap_or(f32::NAN),
                                            labels.as_mut_ptr(), distances.as_mut_ptr(), &mut n) })?;
        labels.truncate(n); distances.truncate(n);
        Ok(Matches { labels, distances })
    }
    /// The peer side of a remote search: the 1152-byte i24 embedding straight off the wire (src/search/vector.rs:48-87).
    pub fn search_i24(&self, wire: &[u8], count: usize, distance_limit: Option<f32>) -> anyhow::Result<Matches> {
        anyhow::ensure!(wire.len() == 1152, "i24 embedding must be 1152 bytes");
        let (mut labels, mut distances, mut n) = (vec![0u64; count], vec![0f32; count], 0usize);
        ck(unsafe { dawn_index_search_i24(self.h, wire.as_ptr(), count, distance_limit.is_some() as c_int, distance_limit.unwrap_or(0.0),
                                          labels.as_mut_ptr(), distances.as_mut_ptr(), &mut n) })?;
        labels.truncate(n); distances.truncate(n);
        Ok(Matches { labels, distances })
    }
    /// `embedding_for_page` (search_provider.rs:183-195) served from the device corpus (the stored, fp16-rounded vector).
    pub fn get(&self, label: u64) -> anyhow::Result<Vec<f32>> {
        let mut v = vec![0f32; 384];
        ck(unsafe { dawn_index_get(self.h, label, v.as_mut_ptr()) })?;
        Ok(v)
    }
    /// `GetEmbedding` over UDP (udp_service.rs:254-276): the stored vector in wire format.
    pub fn get_i24(&self, label: u64) -> anyhow::Result<Vec<u8>> {
        let mut v = vec![0u8; 1152];
        ck(unsafe { dawn_index_get_i24(self.h, label, v.as_mut_ptr()) })?;
        Ok(v)
    }
    pub fn size(&self) -> usize { unsafe { dawn_index_size(self.h) } }                 // :246,280
    pub fn capacity(&self) -> usize { unsafe { dawn_index_capacity(self.h) } }         // :280
    pub fn dimensions(&self) -> usize { unsafe { dawn_index_dimensions(self.h) } }
    pub fn save(&self, path: &str) -> anyhow::Result<()> { let p = CString::new(path)?; ck(unsafe { dawn_index_save(self.h, p.as_ptr()) }) }   // :117,178
    pub fn load(&self, path: &str) -> anyhow::Result<()> { let p = CString::new(path)?; ck(unsafe { dawn_index_load(self.h, p.as_ptr()) }) }   // :115
    pub fn view(&self, path: &str) -> anyhow::Result<()> { self.load(path) }          // examples_old/search_usearch.rs:47
}

impl Drop for Index { fn drop(&mut self) { unsafe { dawn_index_free(self.h) } } }

/// Micro-batching front for `SearchService` (src/search/search_service.rs:55-104): any number of threads call
/// `search` with one query each; a worker thread inside the library answers them in batches, which is what feeds the
/// tensor-core path.  The batcher borrows the index: drop it before the index.
pub struct Batcher { b: *mut c_void }
unsafe impl Send for Batcher {}
unsafe impl Sync for Batcher {}
impl Batcher {
    pub fn new(index: &Index, max_batch: usize, max_wait_us: u32) -> anyhow::Result<Batcher> {
        let mut b = std::ptr::null_mut();
        ck(unsafe { dawn_batcher_create(index.h, max_batch, max_wait_us, &mut b) })?;
        Ok(Batcher { b })
    }
    pub fn search(&self, query: &[f32], count: usize) -> anyhow::Result<Matches> {
        anyhow::ensure!(query.len() == 384, "query must have 384 dimensions");
        let (mut labels, mut distances, mut n) = (vec![0u64; count], vec![0f32; count], 0usize);
        ck(unsafe { dawn_batcher_search(self.b, query.as_ptr(), count, labels.as_mut_ptr(), distances.as_mut_ptr(), &mut n) })?;
        labels.truncate(n); distances.truncate(n);
        Ok(Matches { labels, distances })
    }
}
impl Drop for Batcher { fn drop(&mut self) { unsafe { dawn_batcher_free(self.b) } } }
