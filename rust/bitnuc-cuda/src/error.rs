//! `NucleotideError`: the reference's enum, variant for variant (bitnuc `src/error.rs`).
use crate::ffi::bn_error_t;
use std::fmt;

#[derive(Debug, PartialEq, Eq)]
pub enum NucleotideError {
    InvalidBase(u8),
    SequenceTooLong(usize),
    InvalidLength(usize),
    IndexOutOfBounds { index: usize, length: usize },
    InvalidRange { start: usize, end: usize, length: usize },
    Unsupported,
}

impl fmt::Display for NucleotideError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            Self::InvalidBase(b) => write!(f, "Invalid nucleotide base: {}", b),
            Self::SequenceTooLong(len) => write!(f, "Sequence length {} exceeds maximum", len),
            Self::InvalidLength(len) => write!(f, "Invalid length: {}", len),
            Self::IndexOutOfBounds { index, length } => write!(f, "Index {} out of bounds for sequence of length {}", index, length),
            Self::InvalidRange { start, end, length } => write!(f, "Invalid range {}..{} for sequence of length {}", start, end, length),
            Self::Unsupported => write!(f, "Unsupported architecture"),
        }
    }
}

impl std::error::Error for NucleotideError {}

/// Maps a `bn_status` + payload onto the reference's error vocabulary.  Negative codes (CUDA failure,
/// bad arguments) have no reference variant: like an allocation failure in `Vec`, they panic.
pub(crate) fn check(rc: i32, e: &bn_error_t) -> Result<(), NucleotideError> {
    match rc {
        0 => Ok(()),
        1 => Err(NucleotideError::InvalidBase(e.base)),
        2 => Err(NucleotideError::SequenceTooLong(e.a as usize)),
        3 => Err(NucleotideError::InvalidLength(e.a as usize)),
        4 => Err(NucleotideError::IndexOutOfBounds { index: e.a as usize, length: e.b as usize }),
        5 => Err(NucleotideError::InvalidRange { start: e.a as usize, end: e.b as usize, length: e.c as usize }),
        6 => Err(NucleotideError::Unsupported),
        -3 => panic!("attempt to subtract with overflow"), // encode(b""): what the reference does (packing/avx.rs:138)
        -6 => panic!("bitnuc-cuda: collective failed (NCCL result {})", e.cuda_error),
        other => panic!("bitnuc-cuda: CUDA/argument failure {} (cuda error {})", other, e.cuda_error),
    }
}
