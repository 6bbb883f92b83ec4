//! Rust shim over `libsst_b200.so` (C ABI in `include/sst_b200.h`).
//!
//! Drop-in for the hot path of the reference workspace: implement the reference's own traits
//! (`static_search_tree::{SearchIndex, SearchScheme}`, static-search-tree/src/lib.rs:30-61) for
//! GPU-resident indices, so `bench.rs` / `test.rs` can use them unchanged.  Reference errors are
//! panics (`assert!`, `unwrap`, `panic = 'abort'`), so every non-zero status panics here too;
//! `try_new` maps `SST_ERR_CAPACITY` to `None` (partitioned_s_tree.rs:271-274,463-466,594-597).
//!
//! NOT COMPILED in the authoring image (no rustc/cargo there); kept in lock-step with the tested
//! C++ mirror `host/sst.hpp`.
use std::ffi::{c_char, c_int, c_void, CStr};
use std::marker::PhantomData;

#[repr(C)] pub struct SstIndex { _p: [u8; 0] }
#[repr(C)] pub struct SstSa { _p: [u8; 0] }
#[repr(C)] pub struct SstMultiSa { _p: [u8; 0] }

pub const SST_LEFT_MAX: u32 = 1;
pub const SST_REVERSE_STORAGE: u32 = 2;
pub const SST_FULL_ARRAY: u32 = 4;
pub const SST_ERR_CAPACITY: c_int = 3;

extern "C" {
    fn sst_last_error() -> *const c_char;
    fn sst_last_status() -> c_int;
    fn sst_stree_build(sorted: *const u32, n: usize, node_b: u32, flags: u32, device: c_int) -> *mut SstIndex;
    fn sst_pstree_build(sorted: *const u32, n: usize, b: u32, variant: c_int, device: c_int) -> *mut SstIndex;
    fn sst_index_free(idx: *mut SstIndex);
    fn sst_index_size_bytes(idx: *const SstIndex) -> usize;
    fn sst_index_layers(idx: *const SstIndex) -> usize;
    fn sst_query(idx: *const SstIndex, qs: *const u32, nq: usize, out_vals: *mut u32, out_idx: *mut u64, scheme: c_int) -> c_int;
    fn sst_eytzinger_build(sorted: *const u32, n: usize, device: c_int) -> *mut SstIndex;
    fn sst_bind_thread_to_device(device: c_int) -> c_int;
    fn sst_host_alloc(bytes: usize) -> *mut c_void;
    fn sst_host_free(p: *mut c_void);
    fn sst_fasta_encode(fasta: *const c_char, len: usize, out_codes: *mut u8, out_len: *mut usize, device: c_int) -> c_int;
    fn sst_kmer_keys(codes: *const u8, n: usize, k: u32, max_keys: usize, out_keys: *mut u32, out_count: *mut usize,
                     sort: c_int, device: c_int) -> c_int;
    fn sst_sa_build(text: *const u8, n: usize, device: c_int) -> *mut SstSa;
    fn sst_sa_free(sa: *mut SstSa);
    fn sst_sa_check(sa: *const SstSa, out_violations: *mut u64) -> c_int;
    fn sst_sa_search(sa: *const SstSa, pats: *const u8, pat_off: *const u64, npat: usize, mode: c_int,
                     out_lo: *mut u32, out_hi: *mut u32, out_pos: *mut u32) -> c_int;
    fn sst_sa_search_probes(sa: *const SstSa, pats: *const u8, pat_off: *const u64, npat: usize,
                            out_pos: *mut u32, out_probes: *mut u32) -> c_int;
    fn sst_multi_sa_build(text: *const u8, n: usize, devices: *const c_int, n_devices: c_int) -> *mut SstMultiSa;
    fn sst_multi_sa_search(m: *const SstMultiSa, pats: *const u8, pat_off: *const u64, npat: usize, mode: c_int,
                           out_lo: *mut u32, out_hi: *mut u32, out_pos: *mut u32) -> c_int;
    fn sst_multi_sa_devices(m: *const SstMultiSa) -> c_int;
    fn sst_multi_sa_free(m: *mut SstMultiSa);
    fn sst_query_reserve(idx: *const SstIndex, nq: usize, want_index: c_int) -> c_int;
    fn sst_query_calibrate(idx: *mut SstIndex, max_nq: usize, out_min_nq: *mut usize) -> c_int;
    fn sst_set_option(name: *const c_char, value: i64) -> c_int;
}

fn last_error() -> String { unsafe { CStr::from_ptr(sst_last_error()).to_string_lossy().into_owned() } }
fn check(rc: c_int) { if rc != 0 { panic!("sst_b200: {}", last_error()); } }

/// Owns one GPU index.  `Sync`: the C ABI allows concurrent queries on one handle.
pub struct GpuIndex { h: *mut SstIndex }
unsafe impl Send for GpuIndex {}
unsafe impl Sync for GpuIndex {}
impl Drop for GpuIndex { fn drop(&mut self) { unsafe { sst_index_free(self.h) } } }

impl GpuIndex {
    fn from_raw(h: *mut SstIndex) -> Self { if h.is_null() { panic!("sst_b200: {}", last_error()); } Self { h } }
    pub fn size(&self) -> usize { unsafe { sst_index_size_bytes(self.h) } }
    pub fn layers(&self) -> usize { unsafe { sst_index_layers(self.h) } }
    /// Pre-sizes the calling thread's pipeline scratch: later device-side queries of up to `nq` allocate nothing.
    pub fn reserve(&self, nq: usize, want_index: bool) { check(unsafe { sst_query_reserve(self.h, nq, want_index as c_int) }) }
    /// Measures on this device from which batch size on the reordered-batch pipeline wins; `SST_SCHEME_AUTO` then uses it.
    pub fn calibrate(&mut self, max_nq: usize) -> usize {
        let mut v = 0usize;
        check(unsafe { sst_query_calibrate(self.h, max_nq, &mut v) });
        v
    }
    /// `SearchScheme::query`: values of the first key >= q, same length and order as `qs`.
    pub fn query(&self, qs: &[u32]) -> Vec<u32> {
        let mut out = vec![0u32; qs.len()];
        check(unsafe { sst_query(self.h, qs.as_ptr(), qs.len(), out.as_mut_ptr(), std::ptr::null_mut(), 0) });
        out
    }
    /// values and the sorted-array indices (`l` of binary_search.rs:36-49)
    pub fn query_with_index(&self, qs: &[u32]) -> (Vec<u32>, Vec<u64>) {
        let (mut v, mut i) = (vec![0u32; qs.len()], vec![0u64; qs.len()]);
        check(unsafe { sst_query(self.h, qs.as_ptr(), qs.len(), v.as_mut_ptr(), i.as_mut_ptr(), 0) });
        (v, i)
    }
}

/// `STree<B, 16>` on the GPU (s_tree.rs:14-20).
pub struct GpuSTree<const B: usize>(pub GpuIndex);
pub type GpuSTree16 = GpuSTree<16>;
pub type GpuSTree15 = GpuSTree<15>;
impl<const B: usize> GpuSTree<B> {
    /// `STree::new_params(vals, left_max, reverse_storage, full_array)` (s_tree.rs:72-77).
    pub fn new_params(vals: &[u32], left_max: bool, reverse_storage: bool, full_array: bool) -> Self {
        let flags = (left_max as u32) * SST_LEFT_MAX | (reverse_storage as u32) * SST_REVERSE_STORAGE | (full_array as u32) * SST_FULL_ARRAY;
        Self(GpuIndex::from_raw(unsafe { sst_stree_build(vals.as_ptr(), vals.len(), B as u32, flags, 0) }))
    }
    pub fn search(&self, q: u32) -> u32 { self.0.query(&[q])[0] }
    pub fn batch<const P: usize>(&self, qb: &[u32; P]) -> [u32; P] { self.0.query(qb).try_into().unwrap() }
    pub fn batch_interleave_all_128(&self, qs: &[u32]) -> Vec<u32> { self.0.query(qs) }
}

/// `Eytzinger` on the GPU (eytzinger.rs:9-89): baseline layout, unsigned compares.
pub struct GpuEytzinger(pub GpuIndex);
impl GpuEytzinger {
    pub fn new(vals: &[u32]) -> Self { Self(GpuIndex::from_raw(unsafe { sst_eytzinger_build(vals.as_ptr(), vals.len(), 0) })) }
    pub fn search(&self, q: u32) -> u32 { self.0.query(&[q])[0] }
}

/// Binds the calling thread to the CPUs local to `device` (NUMA); returns the size of the CPU set, 0 if unknown.
pub fn bind_thread_to_device(device: i32) -> i32 {
    unsafe { sst_bind_thread_to_device(device as c_int) as i32 }
}

/// Page-locked `u32` buffer (`sst_host_alloc`): lets `GpuIndex::query_into` overlap H2D, kernel and D2H.
pub struct PinnedU32 { p: *mut u32, len: usize }
unsafe impl Send for PinnedU32 {}
impl PinnedU32 {
    pub fn new(len: usize) -> Self {
        let p = unsafe { sst_host_alloc(len * 4) } as *mut u32;
        if p.is_null() { panic!("sst_b200: {}", last_error()); }
        Self { p, len }
    }
    pub fn as_slice(&self) -> &[u32] { unsafe { std::slice::from_raw_parts(self.p, self.len) } }
    pub fn as_mut_slice(&mut self) -> &mut [u32] { unsafe { std::slice::from_raw_parts_mut(self.p, self.len) } }
}
impl Drop for PinnedU32 { fn drop(&mut self) { unsafe { sst_host_free(self.p as *mut c_void) } } }
impl GpuIndex {
    /// Like `query`, writing into a caller-provided (ideally pinned) buffer.
    pub fn query_into(&self, qs: &[u32], out: &mut [u32]) {
        assert_eq!(qs.len(), out.len());
        check(unsafe { sst_query(self.h, qs.as_ptr(), qs.len(), out.as_mut_ptr(), std::ptr::null_mut(), 0) });
    }
}

/// `read_fasta_file`'s decoding (suffix-array-searching/src/util.rs:144-169) on the GPU.
pub fn read_fasta(fasta: &[u8]) -> Vec<u8> {
    let mut out = vec![0u8; fasta.len().max(1)];
    let mut n = 0usize;
    check(unsafe { sst_fasta_encode(fasta.as_ptr() as *const c_char, fasta.len(), out.as_mut_ptr(), &mut n, 0) });
    out.truncate(n);
    out
}
/// The `--human` key mode of bench.rs:60-76 (+ the sort of :89).
pub fn kmer_keys(codes: &[u8], k: u32, sort: bool) -> Vec<u32> {
    let cap = codes.len().saturating_sub(k as usize - 1);
    let mut out = vec![0u32; cap.max(1)];
    let mut n = 0usize;
    check(unsafe { sst_kmer_keys(codes.as_ptr(), codes.len(), k, cap, out.as_mut_ptr(), &mut n, sort as c_int, 0) });
    out.truncate(n);
    out
}

/// Marker layouts of partitioned_s_tree.rs:34-81.
pub trait Layout { const VARIANT: c_int; }
pub struct Simple; pub struct Compact; pub struct L1; pub struct Overlapping; pub struct Map;
impl Layout for Simple { const VARIANT: c_int = 1; }
impl Layout for Compact { const VARIANT: c_int = 2; }
impl Layout for L1 { const VARIANT: c_int = 3; }
impl Layout for Overlapping { const VARIANT: c_int = 4; }
impl Layout for Map { const VARIANT: c_int = 5; }

/// `PartitionedSTree<16, 16, Tp>` on the GPU.
pub struct GpuPartitionedSTree<Tp: Layout>(pub GpuIndex, PhantomData<Tp>);
impl<Tp: Layout> GpuPartitionedSTree<Tp> {
    pub fn new(vals: &[u32], b: usize) -> Self { Self::try_new(vals, b).unwrap() }
    pub fn try_new(vals: &[u32], b: usize) -> Option<Self> {
        let h = unsafe { sst_pstree_build(vals.as_ptr(), vals.len(), b as u32, Tp::VARIANT, 0) };
        if h.is_null() && unsafe { sst_last_status() } == SST_ERR_CAPACITY { return None; }
        Some(Self(GpuIndex::from_raw(h), PhantomData))
    }
    pub fn search(&self, q: u32) -> u32 { self.0.query(&[q])[0] }
}

// In the reference workspace, add (static-search-tree/src/lib.rs):
//
//   impl<const B: usize> SearchIndex for sst_b200::GpuSTree<B> {
//       fn new(vals: &[u32]) -> Self { Self::new_params(vals, false, false, false) }
//       fn size(&self) -> usize { self.0.size() }
//       fn layers(&self) -> usize { self.0.layers() }
//   }
//   // one scheme serves every GPU index: the whole slice goes down in one call
//   pub const GPU: Full<..> = full(|idx: &sst_b200::GpuSTree16, qs: &[u32]| idx.0.query(qs));

/// `SaNaive` on the GPU (sa_search.rs:11-57): text and suffix array live in HBM.
pub struct GpuSa<'t> { h: *mut SstSa, _t: PhantomData<&'t [u8]> }
unsafe impl Send for GpuSa<'_> {}
unsafe impl Sync for GpuSa<'_> {}
impl Drop for GpuSa<'_> { fn drop(&mut self) { unsafe { sst_sa_free(self.h) } } }
impl<'t> GpuSa<'t> {
    /// `SaNaive::build(t)`: builds the suffix array on the GPU and asserts strict suffix order
    /// exactly like sa_search.rs:36-38.
    pub fn build(t: &'t [u8]) -> Self {
        let h = unsafe { sst_sa_build(t.as_ptr(), t.len(), 0) };
        if h.is_null() { panic!("sst_b200: {}", last_error()); }
        let mut bad = 0u64;
        check(unsafe { sst_sa_check(h, &mut bad) });
        assert!(bad == 0);
        Self { h, _t: PhantomData }
    }
    /// Batched `binary_search` (sa_search.rs:98-112): returns `sa[l]` per pattern.
    pub fn binary_search_batch(&self, qs: &[&[u8]], mlr: bool) -> Vec<usize> {
        let mut flat = Vec::new();
        let mut off = vec![0u64];
        for q in qs { flat.extend_from_slice(q); off.push(flat.len() as u64); }
        let (mut lo, mut pos) = (vec![0u32; qs.len()], vec![0u32; qs.len()]);
        check(unsafe { sst_sa_search(self.h, flat.as_ptr(), off.as_ptr(), qs.len(), mlr as c_int, lo.as_mut_ptr(),
                                     std::ptr::null_mut(), pos.as_mut_ptr()) });
        pos.into_iter().map(|p| p as usize).collect()
    }
}
/// `fn(&SaNaive, &[u8], &mut usize) -> usize` (type F1, sa_search.rs:453).  `cnt` advances by the number of probes the
/// reference's loop makes for this pattern (sa_search.rs:98-112: one per iteration), counted on the device by the plain
/// binary-search kernel (`sst_sa_search_probes`); use `GpuSa::binary_search_batch` when the counter is not needed.
pub fn binary_search(sa: &GpuSa, q: &[u8], cnt: &mut usize) -> usize {
    let off = [0u64, q.len() as u64];
    let (mut pos, mut probes) = (0u32, 0u32);
    check(unsafe { sst_sa_search_probes(sa.h, q.as_ptr(), off.as_ptr(), 1, &mut pos, &mut probes) });
    *cnt += probes as usize;
    pos as usize
}

/// Text + suffix array replicated on several GPUs, the pattern batch sharded contiguously (chunk = ceil(npat / G)): the serial
/// callers of sa_search.rs:423-451 in the harness shape of static-search-tree/src/bin/bench.rs:558-573.
pub struct GpuMultiSa { h: *mut SstMultiSa }
unsafe impl Send for GpuMultiSa {}
unsafe impl Sync for GpuMultiSa {}
impl Drop for GpuMultiSa { fn drop(&mut self) { unsafe { sst_multi_sa_free(self.h) } } }
impl GpuMultiSa {
    pub fn build(t: &[u8], devices: &[i32]) -> Self {
        let h = unsafe { sst_multi_sa_build(t.as_ptr(), t.len(), devices.as_ptr(), devices.len() as c_int) };
        if h.is_null() { panic!("sst_b200: {}", last_error()); }
        Self { h }
    }
    pub fn devices(&self) -> usize { unsafe { sst_multi_sa_devices(self.h) as usize } }
    /// Batched `binary_search`: `sa[l]` per pattern, in pattern order.
    pub fn binary_search_batch(&self, qs: &[&[u8]], mlr: bool) -> Vec<usize> {
        let mut flat = Vec::new();
        let mut off = vec![0u64];
        for q in qs { flat.extend_from_slice(q); off.push(flat.len() as u64); }
        let (mut lo, mut pos) = (vec![0u32; qs.len()], vec![0u32; qs.len()]);
        check(unsafe { sst_multi_sa_search(self.h, flat.as_ptr(), off.as_ptr(), qs.len(), mlr as c_int, lo.as_mut_ptr(),
                                           std::ptr::null_mut(), pos.as_mut_ptr()) });
        pos.into_iter().map(|p| p as usize).collect()
    }
}

/// Library option by name (the table of csrc/common.cuh), e.g. `set_option("SA_CHUNK", 1 << 22)`.
pub fn set_option(name: &str, value: i64) {
    let c = std::ffi::CString::new(name).unwrap();
    check(unsafe { sst_set_option(c.as_ptr(), value) });
}
