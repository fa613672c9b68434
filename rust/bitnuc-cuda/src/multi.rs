//! `Multi`: the hot path sharded over the GPUs of one box from one process (`bn_multi_*` of `include/bitnuc_cuda.h`).
//!
//! The reference is single-threaded and has no analogue (bitnuc `src/lib.rs:214-220` is the whole surface); the
//! methods keep the reference's argument order and `Vec` semantics so that `multi.encode(seq, &mut ebuf)` reads like
//! `bitnuc::encode(seq, &mut ebuf)`.  Shards are contiguous (base ranges on 64-base boundaries, records / pairs /
//! reads by index, variable-length reads by byte volume); the four base counters are summed over the shards by
//! `ncclAllReduce` (`Reduce::Nccl`) or by the library's own all-reduce kernel over NVLink peer memory (`Reduce::P2p`).
//!
//! NOT COMPILED IN THE AUTHORING ENVIRONMENT (no Rust toolchain there); see INTEGRATION.md.
use crate::error::{check, NucleotideError};
use crate::ffi::*;
use std::ptr;

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Reduce {
    Nccl,
    P2p,
}

pub struct Multi(*mut bn_multi);

// the library serialises calls on one bn_multi internally
unsafe impl Send for Multi {}
unsafe impl Sync for Multi {}

impl Drop for Multi {
    fn drop(&mut self) {
        unsafe { bn_multi_destroy(self.0) }
    }
}

/// A packed batch of variable-length reads: read `r` is `words[word_offsets[r]..word_offsets[r + 1]]`.
#[derive(Debug, Default, PartialEq, Eq, Clone)]
pub struct PackedBatch {
    pub words: Vec<u64>,
    pub word_offsets: Vec<u64>,
}

/// `InvalidBase` of a batch call, with the read that holds it and the position inside the read.
#[derive(Debug, PartialEq, Eq)]
pub struct BatchError {
    pub error: NucleotideError,
    pub record: u64,
    pub position: u64,
}

impl Multi {
    /// `devices`: device ordinals (empty = every visible device).  Panics when no sm_100 device is usable or the
    /// collective cannot be set up (no libnccl.so.2 / no NVLink peer access): there is no CPU fallback.
    pub fn new(devices: &[i32], reduce: Reduce) -> Self {
        let mut h = ptr::null_mut();
        let mode = if reduce == Reduce::Nccl { BN_REDUCE_NCCL } else { BN_REDUCE_P2P };
        let rc = unsafe {
            if devices.is_empty() { bn_multi_create(ptr::null(), 0, mode, &mut h) } else { bn_multi_create(devices.as_ptr(), devices.len() as i32, mode, &mut h) }
        };
        assert_eq!(rc, BN_OK, "bitnuc-cuda: bn_multi_create failed with {rc}");
        Multi(h)
    }

    pub fn size(&self) -> usize {
        unsafe { bn_multi_size(self.0) as usize }
    }

    /// ncclGetVersion() of the library in use (0 with `Reduce::P2p`).
    pub fn nccl_version(&self) -> i32 {
        unsafe { bn_multi_nccl_version(self.0) }
    }

    /// `bitnuc::encode` over all devices: `ebuf` is cleared, then filled; on `InvalidBase` it keeps the words of the
    /// chunks before the failing chunk, like the reference (`packing/avx.rs:142-143`).
    pub fn encode(&self, sequence: &[u8], ebuf: &mut Vec<u64>) -> Result<(), NucleotideError> {
        let e0 = bn_error_t::default();
        if sequence.is_empty() {
            check(BN_ERR_EMPTY_ENCODE, &e0)?; // panics, like the reference
        }
        ebuf.clear();
        ebuf.resize(sequence.len().div_ceil(32), 0);
        let (mut n_words, mut e) = (0usize, bn_error_t::default());
        let rc = unsafe { bn_multi_encode(self.0, sequence.as_ptr(), sequence.len(), ebuf.as_mut_ptr(), &mut n_words, &mut e) };
        ebuf.truncate(n_words);
        check(rc, &e)
    }

    /// `bitnuc::decode` over all devices: appends `n_bases` bytes to `dbuf`.
    pub fn decode(&self, ebuf: &[u64], n_bases: usize, dbuf: &mut Vec<u8>) -> Result<(), NucleotideError> {
        let old = dbuf.len();
        dbuf.resize(old + n_bases, 0);
        let mut e = bn_error_t::default();
        let rc = unsafe { bn_multi_decode(self.0, ebuf.as_ptr(), ebuf.len(), n_bases, dbuf.as_mut_ptr().add(old), &mut e) };
        if rc != BN_OK {
            dbuf.truncate(old);
        }
        check(rc, &e)
    }

    /// Batched `as_2bit`: `n` records of `k` bases every `stride` bytes.
    pub fn as_2bit_batch(&self, recs: &[u8], n: usize, k: usize, stride: usize) -> Result<Vec<u64>, NucleotideError> {
        assert!(n == 0 || k > 32 || (stride >= k && recs.len() >= (n - 1) * stride + k));
        let mut out = vec![0u64; n];
        let mut e = bn_error_t::default();
        let rc = unsafe { bn_multi_as_2bit_batch(self.0, recs.as_ptr(), n, k.min(u32::MAX as usize) as u32, stride, out.as_mut_ptr(), &mut e) };
        check(rc, &e).map(|_| out)
    }

    /// Batched `from_2bit`: the low `k` bases of every word, records `stride` bytes apart.
    pub fn from_2bit_batch(&self, packed: &[u64], k: usize, stride: usize) -> Result<Vec<u8>, NucleotideError> {
        let n = packed.len();
        let mut out = vec![0u8; if n > 0 && k <= 32 && stride >= k { (n - 1) * stride + k } else { 0 }];
        let mut e = bn_error_t::default();
        let rc = unsafe { bn_multi_from_2bit_batch(self.0, packed.as_ptr(), n, k.min(u32::MAX as usize) as u32, out.as_mut_ptr(), stride.max(1), &mut e) };
        check(rc, &e).map(|_| out)
    }

    /// Exact mismatch count (`bitnuc::hdist` wraps its `u32` accumulator above 2^32 - 1).
    pub fn hdist_total(&self, ebuf1: &[u64], ebuf2: &[u64], n_bases: usize) -> Result<u64, NucleotideError> {
        let (mut total, mut e) = (0u64, bn_error_t::default());
        let rc = unsafe { bn_multi_hdist(self.0, ebuf1.as_ptr(), ebuf1.len(), ebuf2.as_ptr(), ebuf2.len(), n_bases, &mut total, &mut e) };
        check(rc, &e).map(|_| total)
    }

    pub fn hdist(&self, ebuf1: &[u64], ebuf2: &[u64], n_bases: usize) -> Result<u32, NucleotideError> {
        self.hdist_total(ebuf1, ebuf2, n_bases).map(|t| t as u32)
    }

    /// `hdist_scalar(u[i], v[i], len)` for every pair.
    pub fn hdist_pairs(&self, u: &[u64], v: &[u64], len: usize) -> Result<Vec<u32>, NucleotideError> {
        assert_eq!(u.len(), v.len());
        let mut out = vec![0u32; u.len()];
        let mut e = bn_error_t::default();
        let rc = unsafe { bn_multi_hdist_pairs(self.0, u.as_ptr(), v.as_ptr(), u.len(), len.min(u32::MAX as usize) as u32, out.as_mut_ptr(), &mut e) };
        check(rc, &e).map(|_| out)
    }

    /// `BaseCount::base_counts` + `GCContent::gc_content` of one packed sequence; the four counters are all-reduced.
    pub fn base_counts_gc(&self, data: &[u64], length: usize) -> Result<([usize; 4], f64), NucleotideError> {
        let (mut counts, mut gc, mut e) = ([0u64; 4], 0f64, bn_error_t::default());
        let rc = unsafe { bn_multi_base_counts(self.0, data.as_ptr(), data.len(), length, counts.as_mut_ptr(), &mut gc, &mut e) };
        check(rc, &e).map(|_| (counts.map(|x| x as usize), gc))
    }

    /// Per-read `[A, C, G, T]` and gc of a packed batch plus the all-reduced totals.
    pub fn base_counts_batch(&self, batch: &PackedBatch, lens: &[u64]) -> Result<(Vec<[u64; 4]>, Vec<f64>, [u64; 4]), NucleotideError> {
        let n = lens.len();
        assert!(batch.word_offsets.len() >= n);
        let (mut counts4, mut gc, mut totals, mut e) = (vec![[0u64; 4]; n], vec![0f64; n], [0u64; 4], bn_error_t::default());
        let rc = unsafe {
            bn_multi_base_counts_batch(self.0, batch.words.as_ptr(), batch.words.len(), batch.word_offsets.as_ptr(), lens.as_ptr(), n,
                                       counts4.as_mut_ptr() as *mut u64, gc.as_mut_ptr(), totals.as_mut_ptr(), &mut e)
        };
        check(rc, &e).map(|_| (counts4, gc, totals))
    }

    /// `PackedSequence::new` over a batch of reads `bytes[offsets[r]..offsets[r + 1]]`, sharded by byte volume.
    pub fn encode_batch(&self, bytes: &[u8], offsets: &[u64]) -> Result<PackedBatch, BatchError> {
        assert!(!offsets.is_empty() && offsets.windows(2).all(|w| w[0] <= w[1]) && *offsets.last().unwrap() as usize <= bytes.len());
        let n = offsets.len() - 1;
        let cap = ((offsets[n] - offsets[0]) / 32) as usize + n;
        let mut b = PackedBatch { words: vec![0; cap], word_offsets: vec![0; n + 1] };
        let mut e = bn_error_t::default();
        let rc = unsafe {
            bn_multi_encode_batch(self.0, bytes.as_ptr(), offsets.as_ptr(), n, b.words.as_mut_ptr(), b.word_offsets.as_mut_ptr(), ptr::null_mut(), &mut e)
        };
        check(rc, &e).map_err(|error| BatchError { error, record: e.record, position: e.b })?;
        b.words.truncate(b.word_offsets[n] as usize);
        Ok(b)
    }
}
