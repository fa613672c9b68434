"""bitnuc_b200 -- B200-native (sm_100a CUDA) implementation of bitnuc's data-parallel hot path.

The public names are the reference's (/root/reference/src/lib.rs:214-220): ``as_2bit``, ``from_2bit``,
``from_2bit_alloc``, ``encode``, ``encode_alloc``, ``decode``, ``hdist``, ``hdist_scalar``, ``split_packed``,
``PackedSequence`` (with ``base_counts`` / ``gc_content``) and ``NucleotideError``; plus array/batch
forms and, in ``bitnuc_b200.device``, device-resident variants on torch CUDA tensors.

All computation happens in hand-written CUDA kernels behind the C ABI of ``libbitnuc_cuda.so``
(include/bitnuc_cuda.h).  There is no CPU fallback: importing works anywhere, calling needs a B200.
"""
from .errors import BitnucCudaError, FastqError, NucleotideError, ReferencePanic
from .api import (Context, as_2bit, as_2bit_batch, base_counts_batch, base_counts_gc, decode, decode_np,
                  default_context, encode, encode_alloc, encode_batch, encode_np, fasta_encode, fasta_wrapped_encode, fastq_encode, from_2bit, from_2bit_alloc,
                  from_2bit_batch, hdist, hdist_pairs, hdist_scalar, hdist_total, get_batch, kmers, kmers_batch, slice_batch, split_packed, split_packed_batch)
from .multi import MultiContext
from .sequence import PackedSequence

__all__ = [
    "NucleotideError", "ReferencePanic", "BitnucCudaError", "PackedSequence", "Context", "default_context",
    "as_2bit", "from_2bit", "from_2bit_alloc", "encode", "encode_alloc", "decode", "hdist", "hdist_scalar",
    "encode_np", "decode_np", "as_2bit_batch", "from_2bit_batch", "hdist_pairs", "hdist_total",
    "base_counts_gc", "base_counts_batch", "encode_batch", "split_packed", "split_packed_batch", "slice_batch", "get_batch", "kmers", "kmers_batch",
    "fastq_encode", "fasta_encode", "fasta_wrapped_encode", "FastqError", "MultiContext",
]
