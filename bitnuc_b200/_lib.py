"""ctypes binding of libbitnuc_cuda.so -- one prototype per symbol declared in include/bitnuc_cuda.h."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .errors import BitnucCudaError, FastqError, NucleotideError, ReferencePanic

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libbitnuc_cuda.so"

BN_OK = 0
BN_ERR_CUDA, BN_ERR_ARGUMENT, BN_ERR_EMPTY_ENCODE, BN_ERR_NOMEM, BN_ERR_FASTQ, BN_ERR_COLLECTIVE = -1, -2, -3, -4, -5, -6


class BnError(C.Structure):
    _fields_ = [("code", C.c_int32), ("base", C.c_uint8), ("pad_", C.c_uint8 * 3), ("a", C.c_uint64),
                ("b", C.c_uint64), ("c", C.c_uint64), ("offset", C.c_uint64), ("record", C.c_uint64),
                ("cuda_error", C.c_int32), ("pad2_", C.c_int32)]


_vp, _sz, _u64, _u32, _int = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int
_errp = C.POINTER(BnError)

# name -> (restype, argtypes); must list every symbol of include/bitnuc_cuda.h
PROTOTYPES = {
    "bn_abi_version": (_int, []),
    "bn_device_count": (_int, []),
    "bn_error_string": (_int, [_errp, C.c_char_p, _sz]),
    "bn_ctx_create": (_int, [_int, C.POINTER(_vp)]),
    "bn_ctx_destroy": (None, [_vp]),
    "bn_ctx_device": (_int, [_vp]),
    "bn_ctx_stream": (_vp, [_vp]),
    "bn_ctx_synchronize": (_int, [_vp]),
    "bn_ctx_set_chunk_bytes": (_int, [_vp, _sz]),
    "bn_ctx_set_compat": (_int, [_vp, _int]),
    "bn_ctx_compat": (_int, [_vp]),
    "bn_ctx_set_timing": (_int, [_vp, _int]),
    "bn_last_kernel_ms": (_int, [_vp, C.POINTER(C.c_float)]),
    "bn_dev_alloc": (_int, [_vp, _sz, C.POINTER(_vp)]),
    "bn_dev_free": (_int, [_vp, _vp]),
    "bn_host_alloc": (_int, [_vp, _sz, C.POINTER(_vp)]),
    "bn_host_free": (_int, [_vp, _vp]),
    "bn_copy_h2d": (_int, [_vp, _vp, _vp, _sz]),
    "bn_copy_d2h": (_int, [_vp, _vp, _vp, _sz]),
    "bn_encode": (_int, [_vp, _vp, _sz, _vp, C.POINTER(_sz), _errp]),
    "bn_decode": (_int, [_vp, _vp, _sz, _sz, _vp, _errp]),
    "bn_as_2bit_batch": (_int, [_vp, _vp, _sz, _u32, _sz, _vp, _errp]),
    "bn_from_2bit_batch": (_int, [_vp, _vp, _sz, _u32, _vp, _sz, _errp]),
    "bn_hdist": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, C.POINTER(_u64), _errp]),
    "bn_hdist_pairs": (_int, [_vp, _vp, _vp, _sz, _u32, _vp, _errp]),
    "bn_base_counts": (_int, [_vp, _vp, _sz, _sz, C.POINTER(_u64), C.POINTER(C.c_double), _errp]),
    "bn_base_counts_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _vp, _errp]),
    "bn_encode_batch": (_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _errp]),
    "bn_encode_dev": (_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "bn_decode_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "bn_as_2bit_batch_dev": (_int, [_vp, _vp, _vp, _sz, _u32, _sz, _vp, _vp]),
    "bn_from_2bit_batch_dev": (_int, [_vp, _vp, _vp, _sz, _u32, _vp, _sz]),
    "bn_hdist_dev": (_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "bn_hdist_pairs_dev": (_int, [_vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "bn_base_counts_dev": (_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "bn_base_counts_batch_dev": (_int, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _vp]),
    "bn_base_counts_fixed_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp]),
    "bn_encode_batch_scratch_bytes": (_sz, [_sz, _sz]),
    "bn_encode_batch_dev": (_int, [_vp, _vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "bn_split_packed_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _errp]),
    "bn_split_packed_scratch_bytes": (_sz, [_sz]),
    "bn_split_packed_batch_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bn_kmers": (_int, [_vp, _vp, _sz, _u32, _vp, C.POINTER(_sz), _errp]),
    "bn_kmers_dev": (_int, [_vp, _vp, _vp, _sz, _u32, _vp, _vp]),
    "bn_kmers_batch": (_int, [_vp, _vp, _vp, _sz, _u32, _vp, _sz, _vp, _errp]),
    "bn_kmers_batch_scratch_bytes": (_sz, [_sz, _sz]),
    "bn_kmers_batch_dev": (_int, [_vp, _vp, _vp, _vp, _sz, _sz, _u32, _vp, _vp, _vp, _vp]),
    "bn_slice_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _errp]),
    "bn_get_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _errp]),
    "bn_slice_batch_scratch_bytes": (_sz, [_sz]),
    "bn_slice_batch_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "bn_get_batch_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "bn_fastq_scan": (_int, [_vp, _vp, _sz, C.POINTER(_sz), C.POINTER(_sz), _errp]),
    "bn_fastq_encode": (_int, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _errp]),
    "bn_fasta_scan": (_int, [_vp, _vp, _sz, C.POINTER(_sz), C.POINTER(_sz), _errp]),
    "bn_fasta_encode": (_int, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _errp]),
    "bn_fasta_wrapped_scan": (_int, [_vp, _vp, _sz, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), _errp]),
    "bn_fasta_wrapped_encode": (_int, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _errp]),
    "bn_fasta_count_dev": (_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "bn_fasta_index_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bn_fasta_encode_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bn_fasta_status_fetch": (_int, [_vp, _vp, _vp, _u64, _vp, _sz, _errp]),
    "bn_fastq_scratch_bytes": (_sz, [_sz]),
    "bn_fastq_index_scratch_bytes": (_sz, [_sz]),
    "bn_fastq_count_dev": (_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "bn_fastq_index_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bn_fastq_encode_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bn_fastq_status_fetch": (_int, [_vp, _vp, _vp, _u64, _vp, _sz, _errp]),
    "bn_status_fetch": (_int, [_vp, _vp, _vp, _errp]),
    "bn_synth_words_dev": (_int, [_vp, _vp, _u64, _u64, _u64, _sz, _vp]),
    "bn_synth_ascii_dev": (_int, [_vp, _vp, _u64, _u64, _u64, _sz, _vp]),
    # multi-GPU (one process, N devices)
    "bn_multi_create": (_int, [C.POINTER(_int), _int, _int, C.POINTER(_vp)]),
    "bn_multi_destroy": (None, [_vp]),
    "bn_multi_size": (_int, [_vp]),
    "bn_multi_ctx": (_vp, [_vp, _int]),
    "bn_multi_reduce": (_int, [_vp]),
    "bn_multi_nccl_version": (_int, [_vp]),
    "bn_multi_set_chunk_bytes": (_int, [_vp, _sz]),
    "bn_multi_synchronize": (_int, [_vp]),
    "bn_multi_shard_units": (_int, [_vp, _sz, _sz, C.POINTER(_sz)]),
    "bn_multi_shard_reads": (_int, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "bn_multi_encode": (_int, [_vp, _vp, _sz, _vp, C.POINTER(_sz), _errp]),
    "bn_multi_decode": (_int, [_vp, _vp, _sz, _sz, _vp, _errp]),
    "bn_multi_as_2bit_batch": (_int, [_vp, _vp, _sz, _u32, _sz, _vp, _errp]),
    "bn_multi_from_2bit_batch": (_int, [_vp, _vp, _sz, _u32, _vp, _sz, _errp]),
    "bn_multi_hdist": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, C.POINTER(_u64), _errp]),
    "bn_multi_hdist_pairs": (_int, [_vp, _vp, _vp, _sz, _u32, _vp, _errp]),
    "bn_multi_base_counts": (_int, [_vp, _vp, _sz, _sz, C.POINTER(_u64), C.POINTER(C.c_double), _errp]),
    "bn_multi_base_counts_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _vp, _errp]),
    "bn_multi_encode_batch": (_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _errp]),
    "bn_multi_base_counts_dev": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "bn_multi_base_counts_fixed_dev": (_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "bn_multi_hdist_dev": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "bn_multi_allreduce_u64_dev": (_int, [_vp, _vp, _int]),
    "bn_multi_last_ms": (_int, [_vp, C.POINTER(C.c_float)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise BitnucCudaError(
                f"{LIB_PATH} is missing: build it with `python -m bitnuc_b200.build` (needs nvcc). "
                "bitnuc_b200 has no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library disagree
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def error_string(err: BnError) -> str:
    buf = C.create_string_buffer(256)
    load().bn_error_string(C.byref(err), buf, 256)
    return buf.value.decode()


def raise_for(rc: int, err: BnError | None = None):
    """Translate a bn_status into the reference's error vocabulary."""
    if rc == BN_OK:
        return
    if rc == 1:
        raise NucleotideError.InvalidBase(err.base if err is not None else 0)
    if rc == 2:
        raise NucleotideError.SequenceTooLong(err.a)
    if rc == 3:
        raise NucleotideError.InvalidLength(err.a)
    if rc == 4:
        raise NucleotideError.IndexOutOfBounds(err.a, err.b)
    if rc == 5:
        raise NucleotideError.InvalidRange(err.a, err.b, err.c)
    if rc == 6:
        raise NucleotideError.Unsupported()
    if rc == BN_ERR_FASTQ and err is not None:
        raise FastqError(err.record, err.a)
    if rc == BN_ERR_EMPTY_ENCODE:
        raise ReferencePanic("encode of an empty sequence: the reference panics (packing/avx.rs:138)")
    if err is not None and err.code == rc:
        raise BitnucCudaError(error_string(err))
    raise BitnucCudaError({BN_ERR_CUDA: "CUDA failure (no usable sm_100 device?)", BN_ERR_ARGUMENT: "invalid argument",
                           BN_ERR_NOMEM: "out of memory",
                           BN_ERR_COLLECTIVE: "collective failed (libnccl.so.2 missing or failing, no NVLink peer access, or a peer "
                                              "never arrived)"}.get(rc, f"bn_status {rc}"))
