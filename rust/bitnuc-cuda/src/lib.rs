//! bitnuc-cuda: bitnuc's hot path on B200 GPUs behind bitnuc's own signatures.
//!
//! `use bitnuc_cuda as bitnuc;` is the intended switch: `as_2bit`, `from_2bit`, `from_2bit_alloc`,
//! `encode`, `encode_alloc`, `decode`, `hdist`, `hdist_scalar`, `PackedSequence`, `BaseCount`,
//! `GCContent` and `NucleotideError` keep the reference's signatures and error semantics
//! (bitnuc `src/lib.rs:214-220`).  The batch and device-resident variants are additions.
//!
//! NOT COMPILED IN THE AUTHORING ENVIRONMENT (no Rust toolchain there); see INTEGRATION.md.
mod error;
pub mod ffi;
pub mod multi;

pub use error::NucleotideError;
pub use multi::{Multi, Reduce};
use error::check;
use ffi::*;
use std::cell::RefCell;
use std::ops::Range;
use std::ptr;

struct Ctx(*mut bn_ctx);
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { bn_ctx_destroy(self.0) }
    }
}
thread_local! {
    // one context per host thread: the reference's functions are re-entrant
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}
fn with_ctx<T>(f: impl FnOnce(*mut bn_ctx) -> T) -> T {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            let device = std::env::var("BITNUC_DEVICE").ok().and_then(|d| d.parse().ok()).unwrap_or(0);
            let mut h = ptr::null_mut();
            let rc = unsafe { bn_ctx_create(device, &mut h) };
            assert_eq!(rc, BN_OK, "bitnuc-cuda: no usable sm_100 CUDA device (there is no CPU fallback)");
            *c = Some(Ctx(h));
        }
        f(c.as_ref().unwrap().0)
    })
}

pub fn as_2bit(seq: &[u8]) -> Result<u64, NucleotideError> {
    let (mut out, mut e) = (0u64, bn_error_t::default());
    let k = seq.len().min(u32::MAX as usize) as u32;
    let rc = with_ctx(|c| unsafe { bn_as_2bit_batch(c, seq.as_ptr(), 1, k, seq.len().max(1), &mut out, &mut e) });
    check(rc, &e).map(|_| out)
}

pub fn encode(sequence: &[u8], ebuf: &mut Vec<u64>) -> Result<(), NucleotideError> {
    let e0 = bn_error_t::default();
    if sequence.is_empty() {
        check(BN_ERR_EMPTY_ENCODE, &e0)?; // panics, like the reference
    }
    ebuf.clear();
    ebuf.resize(sequence.len().div_ceil(32), 0);
    let (mut n_words, mut e) = (0usize, bn_error_t::default());
    let rc = with_ctx(|c| unsafe { bn_encode(c, sequence.as_ptr(), sequence.len(), ebuf.as_mut_ptr(), &mut n_words, &mut e) });
    ebuf.truncate(n_words); // on InvalidBase: the words of the chunks before the failing chunk
    check(rc, &e)
}

pub fn encode_alloc(sequence: &[u8]) -> Result<Vec<u64>, NucleotideError> {
    let mut ebuf = Vec::new();
    encode(sequence, &mut ebuf)?;
    Ok(ebuf)
}

pub fn decode(ebuf: &[u64], n_bases: usize, dbuf: &mut Vec<u8>) -> Result<(), NucleotideError> {
    let old = dbuf.len();
    dbuf.resize(old + n_bases, 0); // appends, never clears
    let mut e = bn_error_t::default();
    let rc = with_ctx(|c| unsafe { bn_decode(c, ebuf.as_ptr(), ebuf.len(), n_bases, dbuf.as_mut_ptr().add(old), &mut e) });
    if rc != BN_OK {
        dbuf.truncate(old);
    }
    check(rc, &e)
}

pub fn from_2bit(packed: u64, expected_size: usize, sequence: &mut Vec<u8>) -> Result<(), NucleotideError> {
    let old = sequence.len();
    let k = expected_size.min(u32::MAX as usize) as u32;
    sequence.resize(old + if k <= 32 { k as usize } else { 0 }, 0);
    let mut e = bn_error_t::default();
    let rc = with_ctx(|c| unsafe { bn_from_2bit_batch(c, &packed, 1, k, sequence.as_mut_ptr().add(old), (k as usize).max(1), &mut e) });
    if rc != BN_OK {
        sequence.truncate(old);
    }
    check(rc, &e)
}

pub fn from_2bit_alloc(packed: u64, expected_size: usize) -> Result<Vec<u8>, NucleotideError> {
    let mut sequence = Vec::with_capacity(expected_size.min(32));
    from_2bit(packed, expected_size, &mut sequence)?;
    Ok(sequence)
}

/// Exact mismatch count (the reference's `u32` accumulator wraps above 2^32 - 1 mismatches).
pub fn hdist_total(ebuf1: &[u64], ebuf2: &[u64], n_bases: usize) -> Result<u64, NucleotideError> {
    let (mut total, mut e) = (0u64, bn_error_t::default());
    let rc = with_ctx(|c| unsafe { bn_hdist(c, ebuf1.as_ptr(), ebuf1.len(), ebuf2.as_ptr(), ebuf2.len(), n_bases, &mut total, &mut e) });
    check(rc, &e).map(|_| total)
}

pub fn hdist(ebuf1: &[u64], ebuf2: &[u64], n_bases: usize) -> Result<u32, NucleotideError> {
    hdist_total(ebuf1, ebuf2, n_bases).map(|t| t as u32) // release-build wrap of bitnuc's u32 accumulator
}

pub fn hdist_scalar(u: u64, v: u64, len: usize) -> Result<u32, NucleotideError> {
    let (mut out, mut e) = (0u32, bn_error_t::default());
    let rc = with_ctx(|c| unsafe { bn_hdist_pairs(c, &u, &v, 1, len.min(u32::MAX as usize) as u32, &mut out, &mut e) });
    check(rc, &e).map(|_| out)
}

/// `seq.windows(k).map(as_2bit).collect()` in one kernel launch (README.md:160-180): one packed word per window.
pub fn kmers(seq: &[u8], k: usize) -> Result<Vec<u64>, NucleotideError> {
    assert!(k != 0, "window size must be non-zero"); // slice::windows panics
    let mut out = vec![0u64; if seq.len() >= k { seq.len() - k + 1 } else { 0 }];
    let (mut n_out, mut e) = (0usize, bn_error_t::default());
    let rc = with_ctx(|c| unsafe { bn_kmers(c, seq.as_ptr(), seq.len(), k.min(u32::MAX as usize) as u32, out.as_mut_ptr(), &mut n_out, &mut e) });
    check(rc, &e)?;
    out.truncate(n_out);
    Ok(out)
}

/// What a FASTQ text holds once it is parsed and packed: read `r` is `words[word_offsets[r]..word_offsets[r+1]]`
/// (`seq_lens[r]` bases, each read on fresh words like `PackedSequence::new`), its sequence line sits at
/// `text[seq_offsets[r]..][..seq_lens[r]]`.
#[derive(Debug, Default, PartialEq, Eq, Clone)]
pub struct FastqBatch {
    pub words: Vec<u64>,
    pub word_offsets: Vec<u64>,
    pub seq_offsets: Vec<u64>,
    pub seq_lens: Vec<u64>,
}

/// Malformed FASTQ or an invalid base: the reference has no parser, so the format faults get their own variant.
#[derive(Debug, PartialEq, Eq)]
pub enum FastqError {
    /// `fault`: 1 header without '@', 2 separator without '+', 3 quality/sequence lengths differ, 4 truncated
    Malformed { record: u64, fault: u8 },
    Nucleotide { error: NucleotideError, record: u64, position: u64 },
}

/// The caller's loop `for record in reader { PackedSequence::new(record.seq())? }` (bitnuc README.md:160-180) in two
/// calls on the raw text: records are found and encoded on the device.
pub fn fastq_encode(text: &[u8]) -> Result<FastqBatch, FastqError> {
    fastx_encode(text, false)
}

/// The same for FASTA text with one sequence line per record ('>' header line, sequence line).
pub fn fasta_encode(text: &[u8]) -> Result<FastqBatch, FastqError> {
    fastx_encode(text, true)
}

/// Wrapped (multi-line) FASTA, the genome-file form: a '>' header line, then any number of sequence lines per record,
/// joined into the record's sequence (what a FASTA reader hands to `PackedSequence::new(record.seq())`).
/// `seq_offsets` holds the byte offset of every HEADER line: a record's bases are not contiguous in the text.
pub fn fasta_wrapped_encode(text: &[u8]) -> Result<FastqBatch, FastqError> {
    let (mut n_records, mut n_bases, mut n_words, mut e) = (0usize, 0usize, 0usize, bn_error_t::default());
    let rc = with_ctx(|c| unsafe { bn_fasta_wrapped_scan(c, text.as_ptr(), text.len(), &mut n_records, &mut n_bases, &mut n_words, &mut e) });
    if rc == -5 {
        return Err(FastqError::Malformed { record: e.record, fault: e.a as u8 });
    }
    check(rc, &e).map_err(|error| FastqError::Nucleotide { error, record: e.record, position: e.b })?;
    let mut b = FastqBatch { words: vec![0; n_words], word_offsets: vec![0; n_records + 1], seq_offsets: vec![0; n_records], seq_lens: vec![0; n_records] };
    let rc = with_ctx(|c| unsafe {
        bn_fasta_wrapped_encode(c, text.as_ptr(), text.len(), n_records, n_words, b.words.as_mut_ptr(), b.word_offsets.as_mut_ptr(),
                                b.seq_offsets.as_mut_ptr(), b.seq_lens.as_mut_ptr(), &mut e)
    });
    check(rc, &e).map_err(|error| FastqError::Nucleotide { error, record: e.record, position: e.b })?;
    Ok(b)
}

fn fastx_encode(text: &[u8], fasta: bool) -> Result<FastqBatch, FastqError> {
    let (mut n_reads, mut n_words, mut e) = (0usize, 0usize, bn_error_t::default());
    let rc = with_ctx(|c| unsafe {
        if fasta { bn_fasta_scan(c, text.as_ptr(), text.len(), &mut n_reads, &mut n_words, &mut e) }
        else { bn_fastq_scan(c, text.as_ptr(), text.len(), &mut n_reads, &mut n_words, &mut e) }
    });
    if rc == -5 {
        return Err(FastqError::Malformed { record: e.record, fault: e.a as u8 });
    }
    check(rc, &e).map_err(|error| FastqError::Nucleotide { error, record: e.record, position: e.b })?;
    let mut b = FastqBatch { words: vec![0; n_words], word_offsets: vec![0; n_reads + 1], seq_offsets: vec![0; n_reads], seq_lens: vec![0; n_reads] };
    let rc = with_ctx(|c| unsafe {
        let f = if fasta { bn_fasta_encode } else { bn_fastq_encode };
        f(c, text.as_ptr(), text.len(), n_reads, n_words, b.words.as_mut_ptr(), b.word_offsets.as_mut_ptr(), b.seq_offsets.as_mut_ptr(),
          b.seq_lens.as_mut_ptr(), &mut e)
    });
    check(rc, &e).map_err(|error| FastqError::Nucleotide { error, record: e.record, position: e.b })?;
    Ok(b)
}

/// `bitnuc::split_packed` (src/utils/functions/split.rs:14-20): validates, then clears and fills both buffers.
pub fn split_packed(ebuf: &[u64], slen: usize, idx: usize, lbuf: &mut Vec<u64>, rbuf: &mut Vec<u64>) -> Result<(), NucleotideError> {
    let word_offsets = [0u64, ebuf.len() as u64];
    let (len64, idx64) = (slen as u64, idx as u64);
    let (mut lo, mut ro) = ([0u64; 2], [0u64; 2]);
    let (mut left, mut right) = (vec![0u64; ebuf.len() + 1], vec![0u64; ebuf.len() + 1]);
    let mut e = bn_error_t::default();
    let rc = with_ctx(|c| unsafe {
        bn_split_packed_batch(c, ebuf.as_ptr(), ebuf.len(), word_offsets.as_ptr(), &len64, &idx64, 1, left.as_mut_ptr(), lo.as_mut_ptr(),
                              right.as_mut_ptr(), ro.as_mut_ptr(), &mut e)
    });
    check(rc, &e)?;
    left.truncate(lo[1] as usize);
    right.truncate(ro[1] as usize);
    *lbuf = left;
    *rbuf = right;
    Ok(())
}

/// Batched `as_2bit`: `n` records of `k` bases every `stride` bytes (an addition to the reference API).
pub fn as_2bit_batch(recs: &[u8], n: usize, k: usize, stride: usize) -> Result<Vec<u64>, NucleotideError> {
    assert!(n == 0 || k > 32 || recs.len() >= (n - 1) * stride + k);
    let mut out = vec![0u64; n];
    let mut e = bn_error_t::default();
    let rc = with_ctx(|c| unsafe { bn_as_2bit_batch(c, recs.as_ptr(), n, k.min(u32::MAX as usize) as u32, stride, out.as_mut_ptr(), &mut e) });
    check(rc, &e).map(|_| out)
}

#[derive(Debug, PartialEq, Eq, Clone, Hash)]
pub struct PackedSequence {
    data: Vec<u64>,
    length: usize,
}

impl PackedSequence {
    pub fn new(seq: &[u8]) -> Result<Self, NucleotideError> {
        let mut data = Vec::new();
        if !seq.is_empty() {
            encode(seq, &mut data)?;
        }
        Ok(Self { data, length: seq.len() })
    }
    pub fn len(&self) -> usize {
        self.length
    }
    pub fn is_empty(&self) -> bool {
        self.length == 0
    }
    pub fn get(&self, index: usize) -> Result<u8, NucleotideError> {
        if index >= self.length {
            return Err(NucleotideError::IndexOutOfBounds { index, length: self.length });
        }
        Ok(b"ACGT"[((self.data[index / 32] >> ((index % 32) * 2)) & 0b11) as usize])
    }
    pub fn slice(&self, range: Range<usize>) -> Result<Vec<u8>, NucleotideError> {
        if range.start > range.end || range.end > self.length {
            return Err(NucleotideError::InvalidRange { start: range.start, end: range.end, length: self.length });
        }
        range.map(|i| self.get(i)).collect()
    }
    pub fn to_vec(&self) -> Result<Vec<u8>, NucleotideError> {
        let mut out = Vec::with_capacity(self.length);
        if self.length > 0 {
            decode(&self.data, self.length, &mut out)?;
        }
        Ok(out)
    }
    fn counts_gc(&self) -> ([usize; 4], f64) {
        let (mut counts, mut gc, mut e) = ([0u64; 4], 0f64, bn_error_t::default());
        let rc = with_ctx(|c| unsafe { bn_base_counts(c, self.data.as_ptr(), self.data.len(), self.length, counts.as_mut_ptr(), &mut gc, &mut e) });
        check(rc, &e).expect("PackedSequence invariant: data.len() == ceil(length / 32)");
        (counts.map(|x| x as usize), gc)
    }
}

pub trait GCContent {
    fn gc_content(&self) -> f64;
}
impl GCContent for PackedSequence {
    fn gc_content(&self) -> f64 {
        self.counts_gc().1
    }
}
pub trait BaseCount {
    fn base_counts(&self) -> [usize; 4];
}
impl BaseCount for PackedSequence {
    fn base_counts(&self) -> [usize; 4] {
        self.counts_gc().0
    }
}
