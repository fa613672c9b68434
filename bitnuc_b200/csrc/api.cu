// api.cu -- the C ABI of libbitnuc_cuda.so (see include/bitnuc_cuda.h): contexts, the device-pointer
// entry points (enqueue only) and the host-pointer entry points (H2D -> kernel -> D2H; the stream-, record- and
// read-parallel ones are chunked over a 3-stage multi-stream pipeline so PCIe copies overlap the kernels, and
// pageable caller memory is bounced through pinned stage buffers by a multi-threaded memcpy).
#include "ctx.h"

extern "C" {

// ------------------------------------------------------------------ library / context -------

int bn_abi_version(void) { return BN_ABI_VERSION; }

int bn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int bn_error_string(const bn_error_t* e, char* buf, size_t cap) {
    if (!e || !buf || cap == 0) return 0;
    switch (e->code) {
    case BN_OK: return snprintf(buf, cap, "Ok");
    case BN_INVALID_BASE: return snprintf(buf, cap, "Invalid nucleotide base: %u", (unsigned)e->base);
    case BN_SEQUENCE_TOO_LONG: return snprintf(buf, cap, "Sequence length %llu exceeds maximum", (unsigned long long)e->a);
    case BN_INVALID_LENGTH: return snprintf(buf, cap, "Invalid length: %llu", (unsigned long long)e->a);
    case BN_INDEX_OUT_OF_BOUNDS:
        return snprintf(buf, cap, "Index %llu out of bounds for sequence of length %llu", (unsigned long long)e->a,
                        (unsigned long long)e->b);
    case BN_INVALID_RANGE:
        return snprintf(buf, cap, "Invalid range %llu..%llu for sequence of length %llu", (unsigned long long)e->a,
                        (unsigned long long)e->b, (unsigned long long)e->c);
    case BN_UNSUPPORTED: return snprintf(buf, cap, "Unsupported architecture");
    case BN_ERR_CUDA: return snprintf(buf, cap, "CUDA error %d: %s", e->cuda_error, cudaGetErrorString((cudaError_t)e->cuda_error));
    case BN_ERR_ARGUMENT: return snprintf(buf, cap, "invalid argument");
    case BN_ERR_EMPTY_ENCODE: return snprintf(buf, cap, "encode of an empty sequence (the reference panics)");
    case BN_ERR_NOMEM: return snprintf(buf, cap, "out of memory");
    case BN_ERR_FASTQ: {
        static const char* const what[] = {"malformed record", "header line does not start with '@' (FASTA: '>')", "separator line does not start with '+'",
                                           "quality and sequence lengths differ", "text ends inside the record"};
        return snprintf(buf, cap, "record %llu: %s", (unsigned long long)e->record, what[e->a <= 4 ? e->a : 0]);
    }
    case BN_ERR_COLLECTIVE: return snprintf(buf, cap, "collective failed (NCCL result %d)", e->cuda_error);
    default: return snprintf(buf, cap, "unknown error %d", e->code);
    }
}

int bn_ctx_create(int device, bn_ctx** out) {
    if (!out) return BN_ERR_ARGUMENT;
    *out = nullptr;
    int n = bn_device_count();
    if (device < 0 || device >= n) return n == 0 ? BN_ERR_CUDA : BN_ERR_ARGUMENT;
    bn_ctx* ctx = new (std::nothrow) bn_ctx();
    if (!ctx) return BN_ERR_NOMEM;
    DeviceGuard g(device);
    ctx->di.device = device;
    cudaDeviceProp prop;
    bool ok = cudaGetDeviceProperties(&prop, device) == cudaSuccess;
    if (ok) ctx->di.sm_count = prop.multiProcessorCount;
    ok = ok && prop.major >= 10;  // kernels are built for sm_100a only; no fallback
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int s = 0; ok && s < kStages; ++s) {
        ok = cudaStreamCreateWithFlags(&ctx->stage_stream[s], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->stage_done[s], cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&ctx->d_words, (5 * kStages + 4) * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaHostAlloc(&ctx->h_words, (5 * kStages + 4) * sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        bn_ctx_destroy(ctx);
        return BN_ERR_CUDA;
    }
    *out = ctx;
    return BN_OK;
}

void bn_ctx_destroy(bn_ctx* ctx) {
    if (!ctx) return;
    {
        DeviceGuard g(ctx->di.device);
        cudaDeviceSynchronize();
        for (int s = 0; s < kStages; ++s) {
            if (ctx->stage_stream[s]) cudaStreamDestroy(ctx->stage_stream[s]);
            if (ctx->stage_done[s]) cudaEventDestroy(ctx->stage_done[s]);
            if (ctx->stage_in[s].p) cudaFree(ctx->stage_in[s].p);
            if (ctx->stage_out[s].p) cudaFree(ctx->stage_out[s].p);
            for (auto& b : ctx->stage_aux[s])
                if (b.p) cudaFree(b.p);
            for (auto& b : ctx->hstage_in[s])
                if (b.p) cudaFreeHost(b.p);
            if (ctx->hstage_out[s].p) cudaFreeHost(ctx->hstage_out[s].p);
        }
        for (auto& b : ctx->slot)
            if (b.p) cudaFree(b.p);
        for (auto& b : ctx->fq)
            if (b.p) cudaFree(b.p);
        if (ctx->d_words) cudaFree(ctx->d_words);
        if (ctx->h_words) cudaFreeHost(ctx->h_words);
        if (ctx->t0) cudaEventDestroy(ctx->t0);
        if (ctx->t1) cudaEventDestroy(ctx->t1);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        cudaGetLastError();
    }
    delete ctx;
}

int bn_ctx_device(const bn_ctx* ctx) { return ctx ? ctx->di.device : -1; }
void* bn_ctx_stream(const bn_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int bn_ctx_synchronize(bn_ctx* ctx) {
    if (!ctx) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    bn_error_t* err = nullptr;
    BN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < kStages; ++s) BN_CUDA(cudaStreamSynchronize(ctx->stage_stream[s]));
    return BN_OK;
}

int bn_ctx_set_chunk_bytes(bn_ctx* ctx, size_t bytes) {
    if (!ctx) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (bytes == 0) bytes = kDefaultChunk;
    ctx->chunk = std::max<size_t>(4096, (bytes + 4095) & ~(size_t)4095);  // whole words, 16-byte aligned shards
    return BN_OK;
}

int bn_ctx_set_compat(bn_ctx* ctx, int mode) {
    if (!ctx || (mode != BN_COMPAT_X86_64 && mode != BN_COMPAT_AARCH64)) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->compat = mode;
    return BN_OK;
}

int bn_ctx_compat(const bn_ctx* ctx) { return ctx ? ctx->compat : BN_ERR_ARGUMENT; }

int bn_ctx_set_timing(bn_ctx* ctx, int on) {
    if (!ctx) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (on && !ctx->t0) {
        if (cudaEventCreate(&ctx->t0) != cudaSuccess || cudaEventCreate(&ctx->t1) != cudaSuccess) {
            cudaGetLastError();
            return BN_ERR_CUDA;
        }
    }
    ctx->timing = on != 0;
    ctx->timed = false;
    return BN_OK;
}

int bn_last_kernel_ms(bn_ctx* ctx, float* ms) {
    if (!ctx || !ms) return BN_ERR_ARGUMENT;
    *ms = 0.0f;
    if (!ctx->timing || !ctx->timed) return BN_ERR_ARGUMENT;   // timing is off, or no device-pointer call since it was turned on
    DeviceGuard g(ctx->di.device);
    if (cudaEventSynchronize(ctx->t1) != cudaSuccess || cudaEventElapsedTime(ms, ctx->t0, ctx->t1) != cudaSuccess) {
        cudaGetLastError();
        return BN_ERR_CUDA;
    }
    return BN_OK;
}

int bn_dev_alloc(bn_ctx* ctx, size_t bytes, void** out) {
    bn_error_t* err = nullptr;
    if (!ctx || !out) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    BN_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return BN_OK;
}
int bn_dev_free(bn_ctx* ctx, void* ptr) {
    bn_error_t* err = nullptr;
    if (!ctx) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    BN_CUDA(cudaFree(ptr));
    return BN_OK;
}
int bn_host_alloc(bn_ctx* ctx, size_t bytes, void** out) {
    bn_error_t* err = nullptr;
    if (!ctx || !out) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    BN_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return BN_OK;
}
int bn_host_free(bn_ctx* ctx, void* ptr) {
    bn_error_t* err = nullptr;
    if (!ctx) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    BN_CUDA(cudaFreeHost(ptr));
    return BN_OK;
}
int bn_copy_h2d(bn_ctx* ctx, void* dst, const void* src, size_t bytes) {
    bn_error_t* err = nullptr;
    if (!ctx) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    BN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    BN_CUDA(cudaStreamSynchronize(ctx->stream));
    return BN_OK;
}
int bn_copy_d2h(bn_ctx* ctx, void* dst, const void* src, size_t bytes) {
    bn_error_t* err = nullptr;
    if (!ctx) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    BN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BN_CUDA(cudaStreamSynchronize(ctx->stream));
    return BN_OK;
}

// ------------------------------------------------------------------ device-pointer calls ----

#define BN_LAUNCH(expr)                                  \
    do {                                                 \
        cudaError_t e__ = (expr);                        \
        if (e__ != cudaSuccess) {                        \
            cudaGetLastError();                          \
            return BN_ERR_CUDA;                          \
        }                                                \
    } while (0)

int bn_encode_dev(bn_ctx* ctx, void* stream, const uint8_t* d_seq, size_t n, uint64_t* d_out, uint64_t* d_status) {
    if (!ctx || !d_status || (n && (!d_seq || !d_out))) return BN_ERR_ARGUMENT;
    if (n == 0) return BN_ERR_EMPTY_ENCODE;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_encode(ctx->di, d_seq, n, d_out, reinterpret_cast<unsigned long long*>(d_status), pick(ctx, stream)));
    return BN_OK;
}

int bn_decode_dev(bn_ctx* ctx, void* stream, const uint64_t* d_words, size_t n_words, size_t n_bases, uint8_t* d_out) {
    if (!ctx) return BN_ERR_ARGUMENT;
    if (n_words < (n_bases + 31) / 32) return BN_INVALID_LENGTH;
    if (n_bases && (!d_words || !d_out)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_decode(ctx->di, d_words, n_bases, d_out, pick(ctx, stream)));
    return BN_OK;
}

int bn_as_2bit_batch_dev(bn_ctx* ctx, void* stream, const uint8_t* d_recs, size_t n, uint32_t k, size_t stride,
                         uint64_t* d_out, uint64_t* d_status) {
    if (!ctx) return BN_ERR_ARGUMENT;
    if (k > 32) return BN_SEQUENCE_TOO_LONG;
    if (stride < k || !d_status || (n && (!d_out || (k && !d_recs)))) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_as_2bit_batch(ctx->di, d_recs, n, k, stride, d_out, reinterpret_cast<unsigned long long*>(d_status),
                                       pick(ctx, stream)));
    return BN_OK;
}

int bn_from_2bit_batch_dev(bn_ctx* ctx, void* stream, const uint64_t* d_packed, size_t n, uint32_t k, uint8_t* d_out,
                           size_t stride) {
    if (!ctx) return BN_ERR_ARGUMENT;
    if (k > 32) return BN_INVALID_LENGTH;
    if (stride < k || (n && k && (!d_packed || !d_out))) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_from_2bit_batch(ctx->di, d_packed, n, k, d_out, stride, pick(ctx, stream)));
    return BN_OK;
}

int bn_hdist_dev(bn_ctx* ctx, void* stream, const uint64_t* d_a, const uint64_t* d_b, size_t n_bases, uint64_t* d_total) {
    if (!ctx || !d_total || (n_bases && (!d_a || !d_b))) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_hdist(ctx->di, d_a, d_b, n_bases, reinterpret_cast<unsigned long long*>(d_total), pick(ctx, stream)));
    return BN_OK;
}

int bn_hdist_pairs_dev(bn_ctx* ctx, void* stream, const uint64_t* d_u, const uint64_t* d_v, size_t n_pairs, uint32_t len,
                       uint32_t* d_out) {
    if (!ctx) return BN_ERR_ARGUMENT;
    if (len > 32) return BN_INVALID_LENGTH;
    if (n_pairs && (!d_u || !d_v || !d_out)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_hdist_pairs(ctx->di, d_u, d_v, n_pairs, len, d_out, pick(ctx, stream)));
    return BN_OK;
}

int bn_base_counts_dev(bn_ctx* ctx, void* stream, const uint64_t* d_words, size_t n_bases, uint64_t* d_counts, double* d_gc) {
    if (!ctx || !d_counts || (n_bases && !d_words)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_base_counts(ctx->di, d_words, n_bases, reinterpret_cast<unsigned long long*>(d_counts), d_gc,
                                     pick(ctx, stream)));
    return BN_OK;
}

int bn_base_counts_batch_dev(bn_ctx* ctx, void* stream, const uint64_t* d_words, size_t n_words,
                             const uint64_t* d_word_offsets, const uint64_t* d_lens, size_t n_reads, uint64_t* d_counts4,
                             double* d_gc, uint64_t* d_totals) {
    if (!ctx || (n_reads && (!d_word_offsets || !d_lens))) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_base_counts_batch(ctx->di, d_words, d_word_offsets, d_lens, n_reads, 0, n_words,
                                           reinterpret_cast<unsigned long long*>(d_counts4), d_gc,
                                           reinterpret_cast<unsigned long long*>(d_totals), pick(ctx, stream)));
    return BN_OK;
}

int bn_base_counts_fixed_dev(bn_ctx* ctx, void* stream, const uint64_t* d_words, size_t n_reads, size_t read_len,
                             uint64_t* d_counts4, double* d_gc, uint64_t* d_totals) {
    if (!ctx || (n_reads && read_len && !d_words)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_base_counts_batch(ctx->di, d_words, nullptr, nullptr, n_reads, read_len, n_reads * ((read_len + 31) / 32),
                                           reinterpret_cast<unsigned long long*>(d_counts4), d_gc,
                                           reinterpret_cast<unsigned long long*>(d_totals), pick(ctx, stream)));
    return BN_OK;
}

size_t bn_encode_batch_scratch_bytes(size_t n_reads, size_t n_bytes) { return bn::encode_batch_scratch_bytes(n_reads, n_bytes); }

int bn_encode_batch_dev(bn_ctx* ctx, void* stream, const uint8_t* d_bytes, const uint64_t* d_offsets, size_t n_reads, size_t n_bytes,
                        uint64_t* d_out_words, uint64_t* d_out_word_offsets, uint32_t* d_read_status, uint64_t* d_status,
                        void* d_scratch) {
    if (!ctx || !d_status || !d_out_word_offsets || (n_reads && (!d_offsets || !d_scratch))) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_encode_batch(ctx->di, d_bytes, d_offsets, n_reads, n_bytes, d_out_words, d_out_word_offsets, d_read_status,
                                      reinterpret_cast<unsigned long long*>(d_status), d_scratch, pick(ctx, stream)));
    return BN_OK;
}

size_t bn_split_packed_scratch_bytes(size_t n_reads) { return bn::split_packed_scratch_bytes(n_reads); }

int bn_split_packed_batch_dev(bn_ctx* ctx, void* stream, const uint64_t* d_words, const uint64_t* d_word_offsets,
                              const uint64_t* d_lens, const uint64_t* d_idx, size_t n_reads, uint64_t* d_left,
                              uint64_t* d_left_offsets, uint64_t* d_right, uint64_t* d_right_offsets, uint64_t* d_status,
                              void* d_scratch) {
    if (!ctx || !d_status || !d_left_offsets || !d_right_offsets || (n_reads && (!d_word_offsets || !d_lens || !d_idx || !d_scratch)))
        return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_split_packed_batch(ctx->di, d_words, d_word_offsets, d_lens, d_idx, n_reads, d_left, d_left_offsets, d_right,
                                            d_right_offsets, reinterpret_cast<unsigned long long*>(d_status), d_scratch,
                                            pick(ctx, stream)));
    return BN_OK;
}

int bn_kmers_dev(bn_ctx* ctx, void* stream, const uint8_t* d_seq, size_t n, uint32_t k, uint64_t* d_out, uint64_t* d_status) {
    if (!ctx || !d_status || k == 0) return BN_ERR_ARGUMENT;
    if (n >= k && k > 32) return BN_SEQUENCE_TOO_LONG;
    if (n >= k && (!d_seq || !d_out)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_kmer_windows(ctx->di, d_seq, n, k, d_out, reinterpret_cast<unsigned long long*>(d_status), pick(ctx, stream)));
    return BN_OK;
}

size_t bn_kmers_batch_scratch_bytes(size_t n_reads, size_t n_bytes) { return bn::kmer_windows_batch_scratch_bytes(n_reads, n_bytes); }

int bn_kmers_batch_dev(bn_ctx* ctx, void* stream, const uint8_t* d_bytes, const uint64_t* d_offsets, size_t n_reads, size_t n_bytes, uint32_t k,
                       uint64_t* d_out, uint64_t* d_out_offsets, uint64_t* d_status, void* d_scratch) {
    if (!ctx || !d_status || !d_out_offsets || k == 0 || (n_reads && (!d_offsets || !d_scratch))) return BN_ERR_ARGUMENT;
    if (k > 32) return BN_SEQUENCE_TOO_LONG;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_kmer_windows_batch(ctx->di, d_bytes, d_offsets, n_reads, n_bytes, k, d_out, d_out_offsets,
                                            reinterpret_cast<unsigned long long*>(d_status), d_scratch, pick(ctx, stream)));
    return BN_OK;
}

size_t bn_slice_batch_scratch_bytes(size_t nq) { return bn::slice_batch_scratch_bytes(nq); }

int bn_slice_batch_dev(bn_ctx* ctx, void* stream, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens,
                       size_t n_reads, const uint64_t* d_q_read, const uint64_t* d_q_start, const uint64_t* d_q_end, size_t nq, uint8_t* d_out,
                       uint64_t* d_out_offsets, uint64_t* d_status, void* d_scratch) {
    if (!ctx || !d_status || !d_out_offsets || (nq && (!d_q_read || !d_q_start || !d_q_end || !d_scratch || !d_word_offsets || !d_lens)))
        return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_slice_batch(ctx->di, d_words, d_word_offsets, d_lens, n_reads, d_q_read, d_q_start, d_q_end, nq, d_out, d_out_offsets,
                                     reinterpret_cast<unsigned long long*>(d_status), d_scratch, pick(ctx, stream)));
    return BN_OK;
}

int bn_get_batch_dev(bn_ctx* ctx, void* stream, const uint64_t* d_words, const uint64_t* d_word_offsets, const uint64_t* d_lens, size_t n_reads,
                     const uint64_t* d_q_read, const uint64_t* d_q_index, size_t nq, uint8_t* d_out, uint64_t* d_status) {
    if (!ctx || !d_status || (nq && (!d_q_read || !d_q_index || !d_out || !d_word_offsets || !d_lens))) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_get_batch(ctx->di, d_words, d_word_offsets, d_lens, n_reads, d_q_read, d_q_index, nq, d_out,
                                   reinterpret_cast<unsigned long long*>(d_status), pick(ctx, stream)));
    return BN_OK;
}

size_t bn_fastq_scratch_bytes(size_t n_bytes) { return bn::fastq_scratch_bytes(n_bytes); }
size_t bn_fastq_index_scratch_bytes(size_t n_reads) { return bn::fastq_index_scratch_bytes(n_reads); }

static int fastx_count_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, void* d_scratch, uint64_t* d_n_lines, int fasta) {
    if (!ctx || !d_n_lines || (n_bytes && (!d_text || !d_scratch)) || (reinterpret_cast<uintptr_t>(d_text) & 15u)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_fastq_count(ctx->di, d_text, n_bytes, d_scratch, d_n_lines, fasta, pick(ctx, stream)));
    return BN_OK;
}

static int fastx_index_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, size_t n_reads, void* d_scratch,
                           void* d_index_scratch, uint64_t* d_seq_offsets, uint64_t* d_seq_lens, uint64_t* d_word_offsets, uint64_t* d_status,
                           int fasta) {
    if (!ctx || !d_status || !d_word_offsets || (reinterpret_cast<uintptr_t>(d_text) & 15u) ||
        (reinterpret_cast<uintptr_t>(d_index_scratch) & 15u) ||
        (n_bytes && (!d_text || !d_scratch)) || (n_reads && (!d_index_scratch || !d_seq_offsets || !d_seq_lens)))
        return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_fastq_index(ctx->di, d_text, n_bytes, n_reads, d_scratch, d_index_scratch, d_seq_offsets, d_seq_lens, d_word_offsets,
                                     reinterpret_cast<unsigned long long*>(d_status), fasta, pick(ctx, stream)));
    return BN_OK;
}

static int fastx_encode_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, size_t n_reads, void* d_scratch,
                            const uint64_t* d_seq_offsets, const uint64_t* d_seq_lens, const uint64_t* d_word_offsets, uint64_t* d_out_words,
                            uint64_t* d_status, int fasta) {
    if (!ctx || !d_status || (reinterpret_cast<uintptr_t>(d_text) & 15u) ||
        (n_reads && (!d_text || !d_scratch || !d_seq_offsets || !d_seq_lens || !d_word_offsets)))
        return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_fastq_encode(ctx->di, d_text, n_bytes, n_reads, d_scratch, d_seq_offsets, d_seq_lens, d_word_offsets, d_out_words,
                                      reinterpret_cast<unsigned long long*>(d_status), fasta, pick(ctx, stream)));
    return BN_OK;
}

int bn_fastq_count_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, void* d_scratch, uint64_t* d_n_lines) {
    return fastx_count_dev(ctx, stream, d_text, n_bytes, d_scratch, d_n_lines, 0);
}
int bn_fasta_count_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, void* d_scratch, uint64_t* d_n_lines) {
    return fastx_count_dev(ctx, stream, d_text, n_bytes, d_scratch, d_n_lines, 1);
}
int bn_fastq_index_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, size_t n_reads, void* d_scratch,
                       void* d_index_scratch, uint64_t* d_seq_offsets, uint64_t* d_seq_lens, uint64_t* d_word_offsets, uint64_t* d_status) {
    return fastx_index_dev(ctx, stream, d_text, n_bytes, n_reads, d_scratch, d_index_scratch, d_seq_offsets, d_seq_lens, d_word_offsets, d_status, 0);
}
int bn_fasta_index_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, size_t n_reads, void* d_scratch,
                       void* d_index_scratch, uint64_t* d_seq_offsets, uint64_t* d_seq_lens, uint64_t* d_word_offsets, uint64_t* d_status) {
    return fastx_index_dev(ctx, stream, d_text, n_bytes, n_reads, d_scratch, d_index_scratch, d_seq_offsets, d_seq_lens, d_word_offsets, d_status, 1);
}
int bn_fastq_encode_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, size_t n_reads, void* d_scratch,
                        const uint64_t* d_seq_offsets, const uint64_t* d_seq_lens, const uint64_t* d_word_offsets, uint64_t* d_out_words,
                        uint64_t* d_status) {
    return fastx_encode_dev(ctx, stream, d_text, n_bytes, n_reads, d_scratch, d_seq_offsets, d_seq_lens, d_word_offsets, d_out_words, d_status, 0);
}
int bn_fasta_encode_dev(bn_ctx* ctx, void* stream, const uint8_t* d_text, size_t n_bytes, size_t n_reads, void* d_scratch,
                        const uint64_t* d_seq_offsets, const uint64_t* d_seq_lens, const uint64_t* d_word_offsets, uint64_t* d_out_words,
                        uint64_t* d_status) {
    return fastx_encode_dev(ctx, stream, d_text, n_bytes, n_reads, d_scratch, d_seq_offsets, d_seq_lens, d_word_offsets, d_out_words, d_status, 1);
}

static int fastq_fault(bn_error_t* err, unsigned long long key) {
    set_err(err, BN_ERR_FASTQ, key & 0xFFu);
    if (err) err->record = key >> 8;
    return BN_ERR_FASTQ;
}

static int fastx_status_fetch(bn_ctx* ctx, void* stream, const uint64_t* d_status, uint64_t n_lines, const uint64_t* d_seq_offsets,
                              size_t n_reads, bn_error_t* err, unsigned lines_per_record);

int bn_fastq_status_fetch(bn_ctx* ctx, void* stream, const uint64_t* d_status, uint64_t n_lines, const uint64_t* d_seq_offsets,
                          size_t n_reads, bn_error_t* err) {
    return fastx_status_fetch(ctx, stream, d_status, n_lines, d_seq_offsets, n_reads, err, 4);
}
int bn_fasta_status_fetch(bn_ctx* ctx, void* stream, const uint64_t* d_status, uint64_t n_lines, const uint64_t* d_seq_offsets,
                          size_t n_reads, bn_error_t* err) {
    return fastx_status_fetch(ctx, stream, d_status, n_lines, d_seq_offsets, n_reads, err, 2);
}

static int fastx_status_fetch(bn_ctx* ctx, void* stream, const uint64_t* d_status, uint64_t n_lines, const uint64_t* d_seq_offsets,
                              size_t n_reads, bn_error_t* err, unsigned lines_per_record) {
    if (!ctx || !d_status) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t s = pick(ctx, stream);
    BN_CUDA(cudaMemcpyAsync(ctx->h_words, d_status, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    BN_CUDA(cudaStreamSynchronize(s));
    unsigned long long fault = ctx->h_words[1];
    const unsigned long long key = ctx->h_words[0];
    if (n_lines % lines_per_record) fault = std::min<unsigned long long>(fault, ((n_lines / lines_per_record) << 8) | BN_FASTQ_TRUNCATED);
    if (fault != kNoError) return fastq_fault(err, fault);
    if (key == kNoError) return set_err(err, BN_OK);
    invalid_base(err, key, 0);
    if (err && d_seq_offsets && n_reads) {  // the read holding the byte: the last r with seq_offsets[r] <= offset (rare path: a few 8-byte copies)
        const uint64_t off = key >> 8;
        size_t lo = 0, hi = n_reads - 1;
        uint64_t v = 0;
        while (lo < hi) {
            const size_t mid = lo + (hi - lo + 1) / 2;
            BN_CUDA(cudaMemcpyAsync(&v, d_seq_offsets + mid, 8, cudaMemcpyDeviceToHost, s));
            BN_CUDA(cudaStreamSynchronize(s));
            if (v <= off) lo = mid; else hi = mid - 1;
        }
        BN_CUDA(cudaMemcpyAsync(&v, d_seq_offsets + lo, 8, cudaMemcpyDeviceToHost, s));
        BN_CUDA(cudaStreamSynchronize(s));
        err->record = lo;
        err->b = off - v;
    }
    return BN_INVALID_BASE;
}

int bn_status_fetch(bn_ctx* ctx, void* stream, const uint64_t* d_status, bn_error_t* err) {
    if (!ctx || !d_status) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t s = pick(ctx, stream);
    BN_CUDA(cudaMemcpyAsync(ctx->h_words, d_status, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    BN_CUDA(cudaStreamSynchronize(s));
    const unsigned long long key = ctx->h_words[0];
    if (key == kNoError) return set_err(err, BN_OK);
    return invalid_base(err, key, 0);
}

int bn_synth_words_dev(bn_ctx* ctx, void* stream, uint64_t seed, uint64_t stream_id, uint64_t first_word, size_t n_words,
                       uint64_t* d_out) {
    if (!ctx || (n_words && !d_out)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_synth_words(ctx->di, seed, stream_id, first_word, n_words, d_out, pick(ctx, stream)));
    return BN_OK;
}

int bn_synth_ascii_dev(bn_ctx* ctx, void* stream, uint64_t seed, uint64_t stream_id, uint64_t first_base, size_t n,
                       uint8_t* d_out) {
    if (!ctx || (n && !d_out) || (first_base % 32)) return BN_ERR_ARGUMENT;
    DeviceGuard g(ctx->di.device);
    LaunchTimer lt(ctx, pick(ctx, stream));
    BN_LAUNCH(bn::launch_synth_ascii(ctx->di, seed, stream_id, first_base, n, d_out, pick(ctx, stream)));
    return BN_OK;
}

// ------------------------------------------------------------------ host-pointer calls ------
// encode / decode: the sequence is cut into chunks of ctx->chunk bases (a multiple of 4096, so every
// chunk starts on a word boundary and all device buffers stay 16-byte aligned).  Chunk c runs on
// stage c % kStages: H2D -> kernel -> D2H on that stage's stream, so with pinned host buffers the
// upload of chunk c+1 overlaps the kernel of chunk c and the download of chunk c-1.

int bn_encode(bn_ctx* ctx, const uint8_t* seq, size_t n, uint64_t* out, size_t* n_words, bn_error_t* err) {
    if (n_words) *n_words = 0;
    if (!ctx || (n && (!seq || !out))) return set_err(err, BN_ERR_ARGUMENT);
    if (n == 0) {
        if (ctx->compat != BN_COMPAT_AARCH64) return set_err(err, BN_ERR_EMPTY_ENCODE);
        if (!out) return set_err(err, BN_ERR_ARGUMENT);   // aarch64: as_2bit(b"") = 0 is pushed (packing/aarch64.rs:223-227)
        out[0] = 0;
        if (n_words) *n_words = 1;
        return set_err(err, BN_OK);
    }
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t chunk = ctx->chunk;
    const size_t n_chunks = (n + chunk - 1) / chunk;
    const bool in_pageable = is_pageable(seq), out_pageable = is_pageable(out);
    unsigned long long best = kNoError;  // smallest global (offset << 8 | byte)
    for (int s = 0; s < kStages && (size_t)s < n_chunks; ++s) {
        BN_CUDA(ensure(ctx->stage_in[s], std::min(chunk, n)));
        BN_CUDA(ensure(ctx->stage_out[s], (std::min(chunk, n) + 31) / 32 * 8));
        if (in_pageable) BN_CUDA(ensure_host(ctx->hstage_in[s][0], std::min(chunk, n)));
        if (out_pageable) BN_CUDA(ensure_host(ctx->hstage_out[s], (std::min(chunk, n) + 31) / 32 * 8));
    }
    auto retire = [&](size_t c) -> cudaError_t {
        const int s = (int)(c % kStages);
        cudaError_t e = cudaEventSynchronize(ctx->stage_done[s]);
        if (e != cudaSuccess) return e;
        if (out_pageable) {
            const size_t off = c * chunk, len = std::min(chunk, n - off);
            parallel_memcpy(out + off / 32, ctx->hstage_out[s].p, (len + 31) / 32 * 8);
        }
        const unsigned long long key = ctx->h_words[s];
        if (key != kNoError) {
            const unsigned long long global = (((key >> 8) + (unsigned long long)c * chunk) << 8) | (key & 0xFFu);
            best = std::min(best, global);
        }
        return cudaSuccess;
    };
    size_t issued = 0, retired = 0;
    while (retired < n_chunks && (issued < n_chunks || retired < issued)) {
        if (issued < n_chunks && issued - retired < (size_t)kStages && best == kNoError) {
            const size_t c = issued, off = c * chunk, len = std::min(chunk, n - off);
            const int s = (int)(c % kStages);
            cudaStream_t st = ctx->stage_stream[s];
            unsigned long long* d_status = ctx->d_words + s;
            const void* src = seq + off;
            if (in_pageable) {
                parallel_memcpy(ctx->hstage_in[s][0].p, seq + off, len);
                src = ctx->hstage_in[s][0].p;
            }
            BN_CUDA(cudaMemcpyAsync(ctx->stage_in[s].p, src, len, cudaMemcpyHostToDevice, st));
            BN_CUDA(bn::launch_encode(ctx->di, static_cast<const uint8_t*>(ctx->stage_in[s].p), len,
                                      static_cast<uint64_t*>(ctx->stage_out[s].p), d_status, st));
            BN_CUDA(cudaMemcpyAsync(out_pageable ? ctx->hstage_out[s].p : static_cast<void*>(out + off / 32), ctx->stage_out[s].p,
                                    (len + 31) / 32 * 8, cudaMemcpyDeviceToHost, st));
            BN_CUDA(cudaMemcpyAsync(ctx->h_words + s, d_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            BN_CUDA(cudaEventRecord(ctx->stage_done[s], st));
            ++issued;
        } else if (retired < issued) {
            BN_CUDA(retire(retired));
            ++retired;
        } else {
            break;  // an error stopped the issue of further chunks and everything issued has retired
        }
    }
    if (best != kNoError) {
        if (n_words) *n_words = (size_t)((best >> 8) / 32);
        invalid_base(err, best, 0);
        const size_t off = (size_t)(best >> 8);
        if (ctx->compat == BN_COMPAT_AARCH64 && err && n >= 32 && off < n / 32 * 32) {
            err->base = seq[off & ~(size_t)31];   // the block's first byte, src/utils/packing/aarch64.rs:194-196
            err->a = err->base;
        }
        return BN_INVALID_BASE;
    }
    if (n_words) *n_words = (n + 31) / 32;
    return set_err(err, BN_OK);
}

int bn_decode(bn_ctx* ctx, const uint64_t* words, size_t n_words, size_t n_bases, uint8_t* out, bn_error_t* err) {
    if (!ctx) return set_err(err, BN_ERR_ARGUMENT);
    if (n_words < (n_bases + 31) / 32) return set_err(err, BN_INVALID_LENGTH, n_bases);  // before any pointer is looked at
    if (n_bases == 0) return set_err(err, BN_OK);
    if (!words || !out) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t chunk = ctx->chunk;
    const size_t n_chunks = (n_bases + chunk - 1) / chunk;
    const bool in_pageable = is_pageable(words), out_pageable = is_pageable(out);
    for (int s = 0; s < kStages && (size_t)s < n_chunks; ++s) {
        BN_CUDA(ensure(ctx->stage_out[s], std::min(chunk, n_bases)));
        BN_CUDA(ensure(ctx->stage_in[s], (std::min(chunk, n_bases) + 31) / 32 * 8));
        if (in_pageable) BN_CUDA(ensure_host(ctx->hstage_in[s][0], (std::min(chunk, n_bases) + 31) / 32 * 8));
        if (out_pageable) BN_CUDA(ensure_host(ctx->hstage_out[s], std::min(chunk, n_bases)));
    }
    auto retire = [&](size_t c) -> cudaError_t {  // yields the core while waiting (blocking-sync event)
        const int s = (int)(c % kStages);
        cudaError_t e = cudaEventSynchronize(ctx->stage_done[s]);
        if (e == cudaSuccess && out_pageable) parallel_memcpy(out + c * chunk, ctx->hstage_out[s].p, std::min(chunk, n_bases - c * chunk));
        return e;
    };
    for (size_t c = 0; c < n_chunks; ++c) {
        const size_t off = c * chunk, len = std::min(chunk, n_bases - off);
        const int s = (int)(c % kStages);
        cudaStream_t st = ctx->stage_stream[s];
        if (c >= (size_t)kStages) BN_CUDA(retire(c - kStages));
        const void* src = words + off / 32;
        if (in_pageable) {
            parallel_memcpy(ctx->hstage_in[s][0].p, src, (len + 31) / 32 * 8);
            src = ctx->hstage_in[s][0].p;
        }
        BN_CUDA(cudaMemcpyAsync(ctx->stage_in[s].p, src, (len + 31) / 32 * 8, cudaMemcpyHostToDevice, st));
        BN_CUDA(bn::launch_decode(ctx->di, static_cast<const uint64_t*>(ctx->stage_in[s].p), len,
                                  static_cast<uint8_t*>(ctx->stage_out[s].p), st));
        BN_CUDA(cudaMemcpyAsync(out_pageable ? ctx->hstage_out[s].p : static_cast<void*>(out + off), ctx->stage_out[s].p, len,
                                cudaMemcpyDeviceToHost, st));
        BN_CUDA(cudaEventRecord(ctx->stage_done[s], st));
    }
    for (size_t c = n_chunks > (size_t)kStages ? n_chunks - kStages : 0; c < n_chunks; ++c) BN_CUDA(retire(c));
    return set_err(err, BN_OK);
}

// The record- and stream-parallel calls below use the same 3-stage pipeline through one helper: chunk c is issued on
// stage c % kStages (its own stream: H2D -> kernel -> D2H) once the chunk that used the stage before has retired, so
// with pinned host buffers uploads, kernels and downloads of neighbouring chunks overlap.  Per-stage status /
// accumulator words live at d_words[s] (status) and d_words[kStages + 4 s ..] (up to four accumulators).

}  // extern "C"

namespace {

// check(r, t) over r in [0, n) on up to eight host threads (thread t takes a contiguous range, ranges in index order); a
// non-zero result stops that thread.  Returns the smallest failing r (SIZE_MAX: none) and its result in *kind.  The read and
// query tables of the batch calls are hundreds of megabytes: single-threaded, these passes cost as much as the PCIe transfers.
constexpr unsigned kHostCheckThreads = 8;
template <class F>
static size_t parallel_first_failing(size_t n, F check, int* kind) {
    const unsigned n_thr = (unsigned)std::max<size_t>(1, std::min<size_t>({kHostCheckThreads, std::thread::hardware_concurrency(), n / 262144}));
    std::vector<size_t> first_bad(n_thr, SIZE_MAX);
    std::vector<int> kinds(n_thr, 0);
    auto run = [&](unsigned t) {
        for (size_t r = n * t / n_thr; r < n * (t + 1) / n_thr; ++r)
            if (const int k = check(r, t)) {
                first_bad[t] = r;
                kinds[t] = k;
                return;
            }
    };
    std::vector<std::thread> workers;
    for (unsigned t = 1; t < n_thr; ++t) workers.emplace_back(run, t);
    run(0);
    for (auto& w : workers) w.join();
    for (unsigned t = 0; t < n_thr; ++t)
        if (first_bad[t] != SIZE_MAX) {
            *kind = kinds[t];
            return first_bad[t];
        }
    return SIZE_MAX;
}


// issue(c, s, stream) -> cudaError_t enqueues chunk c; retire(c, s) runs on the host once chunk c has completed.
template <class Issue, class Retire>
int run_pipeline(bn_ctx* ctx, size_t n_chunks, bn_error_t* err, Issue issue, Retire retire) {
    for (size_t c = 0; c < n_chunks; ++c) {
        const int s = (int)(c % kStages);
        if (c >= (size_t)kStages) {
            BN_CUDA(cudaEventSynchronize(ctx->stage_done[s]));
            retire(c - kStages, s);
        }
        BN_CUDA(issue(c, s, ctx->stage_stream[s]));
        BN_CUDA(cudaEventRecord(ctx->stage_done[s], ctx->stage_stream[s]));
    }
    for (size_t c = n_chunks > (size_t)kStages ? n_chunks - kStages : 0; c < n_chunks; ++c) {
        const int s = (int)(c % kStages);
        BN_CUDA(cudaEventSynchronize(ctx->stage_done[s]));
        retire(c, s);
    }
    return BN_OK;
}

#define BN_TRY(expr)                        \
    do {                                    \
        cudaError_t e__ = (expr);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

inline size_t units_per_chunk(const bn_ctx* ctx, size_t bytes_per_unit, size_t multiple) {
    size_t u = ctx->chunk / (bytes_per_unit ? bytes_per_unit : 1);
    u = u / multiple * multiple;
    return u ? u : multiple;
}

// Pageable caller memory goes through the stage's pinned bounce buffers (see is_pageable): the source of an upload,
// and the destination of a download (copied out to the caller by `unbounce` once the chunk has retired).
inline cudaError_t bounce_in(bn_ctx* ctx, int s, int which, const void*& src, size_t bytes, bool pageable) {
    if (!pageable || bytes == 0) return cudaSuccess;
    cudaError_t e = ensure_host(ctx->hstage_in[s][which], bytes);
    if (e != cudaSuccess) return e;
    parallel_memcpy(ctx->hstage_in[s][which].p, src, bytes);
    src = ctx->hstage_in[s][which].p;
    return cudaSuccess;
}
inline cudaError_t bounce_out(bn_ctx* ctx, int s, void*& dst, size_t bytes, bool pageable) {
    if (!pageable || bytes == 0) return cudaSuccess;
    cudaError_t e = ensure_host(ctx->hstage_out[s], bytes);
    if (e == cudaSuccess) dst = ctx->hstage_out[s].p;
    return e;
}
inline void unbounce(bn_ctx* ctx, int s, void* user_dst, size_t bytes, bool pageable) {
    if (pageable && bytes) parallel_memcpy(user_dst, ctx->hstage_out[s].p, bytes);
}

}  // namespace

extern "C" {

int bn_as_2bit_batch(bn_ctx* ctx, const uint8_t* recs, size_t n, uint32_t k, size_t stride, uint64_t* out, bn_error_t* err) {
    if (!ctx) return set_err(err, BN_ERR_ARGUMENT);
    if (k > 32) return set_err(err, BN_SEQUENCE_TOO_LONG, k);  // checked before any content (naive.rs:5-7)
    if (stride < k || (n && (!out || (k && !recs)))) return set_err(err, BN_ERR_ARGUMENT);
    if (n == 0) return set_err(err, BN_OK);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t per = units_per_chunk(ctx, stride, 64), n_chunks = (n + per - 1) / per;
    const bool in_pg = is_pageable(recs), out_pg = is_pageable(out);
    unsigned long long best = kNoError;  // smallest global (offset << 8 | byte)
    const int rc = run_pipeline(
        ctx, n_chunks, err,
        [&](size_t c, int s, cudaStream_t st) -> cudaError_t {
            const size_t r0 = c * per, cnt = std::min(per, n - r0), in_bytes = k ? (cnt - 1) * stride + k : 0;
            BN_TRY(ensure(ctx->stage_in[s], in_bytes ? in_bytes : 1));
            BN_TRY(ensure(ctx->stage_out[s], cnt * 8));
            const void* src = recs + r0 * stride;
            void* dst = out + r0;
            BN_TRY(bounce_in(ctx, s, 0, src, in_bytes, in_pg));
            BN_TRY(bounce_out(ctx, s, dst, cnt * 8, out_pg));
            if (in_bytes) BN_TRY(cudaMemcpyAsync(ctx->stage_in[s].p, src, in_bytes, cudaMemcpyHostToDevice, st));
            BN_TRY(bn::launch_as_2bit_batch(ctx->di, static_cast<const uint8_t*>(ctx->stage_in[s].p), cnt, k, stride,
                                            static_cast<uint64_t*>(ctx->stage_out[s].p), ctx->d_words + s, st));
            BN_TRY(cudaMemcpyAsync(dst, ctx->stage_out[s].p, cnt * 8, cudaMemcpyDeviceToHost, st));
            return cudaMemcpyAsync(ctx->h_words + s, ctx->d_words + s, 8, cudaMemcpyDeviceToHost, st);
        },
        [&](size_t c, int s) {
            unbounce(ctx, s, out + c * per, std::min(per, n - c * per) * 8, out_pg);
            const unsigned long long key = ctx->h_words[s];
            if (key != kNoError) best = std::min(best, (((key >> 8) + (unsigned long long)(c * per * stride)) << 8) | (key & 0xFFu));
        });
    if (rc != BN_OK) return rc;
    if (best != kNoError) {
        invalid_base(err, best, 0);
        if (err) err->record = (best >> 8) / stride;
        return BN_INVALID_BASE;
    }
    return set_err(err, BN_OK);
}

int bn_from_2bit_batch(bn_ctx* ctx, const uint64_t* packed, size_t n, uint32_t k, uint8_t* out, size_t stride, bn_error_t* err) {
    if (!ctx) return set_err(err, BN_ERR_ARGUMENT);
    if (k > 32) return set_err(err, BN_INVALID_LENGTH, k);  // unpacking/naive.rs:8-10
    if (stride < k || (n && k && (!packed || !out))) return set_err(err, BN_ERR_ARGUMENT);
    if (n == 0 || k == 0) return set_err(err, BN_OK);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t per = units_per_chunk(ctx, stride, 64), n_chunks = (n + per - 1) / per;
    const bool in_pg = is_pageable(packed), out_pg = is_pageable(out);
    const int rc = run_pipeline(
        ctx, n_chunks, err,
        [&](size_t c, int s, cudaStream_t st) -> cudaError_t {
            const size_t r0 = c * per, cnt = std::min(per, n - r0), out_bytes = (cnt - 1) * stride + k;
            BN_TRY(ensure(ctx->stage_in[s], cnt * 8));
            BN_TRY(ensure(ctx->stage_out[s], out_bytes));
            const void* src = packed + r0;
            void* dst = out + r0 * stride;
            BN_TRY(bounce_in(ctx, s, 0, src, cnt * 8, in_pg));
            BN_TRY(cudaMemcpyAsync(ctx->stage_in[s].p, src, cnt * 8, cudaMemcpyHostToDevice, st));
            if (stride != k) {  // bytes between records must come back untouched
                const void* gaps = out + r0 * stride;
                BN_TRY(bounce_in(ctx, s, 1, gaps, out_bytes, out_pg));
                BN_TRY(cudaMemcpyAsync(ctx->stage_out[s].p, gaps, out_bytes, cudaMemcpyHostToDevice, st));
            }
            BN_TRY(bounce_out(ctx, s, dst, out_bytes, out_pg));
            BN_TRY(bn::launch_from_2bit_batch(ctx->di, static_cast<const uint64_t*>(ctx->stage_in[s].p), cnt, k,
                                              static_cast<uint8_t*>(ctx->stage_out[s].p), stride, st));
            return cudaMemcpyAsync(dst, ctx->stage_out[s].p, out_bytes, cudaMemcpyDeviceToHost, st);
        },
        [&](size_t c, int s) { unbounce(ctx, s, out + c * per * stride, (std::min(per, n - c * per) - 1) * stride + k, out_pg); });
    return rc != BN_OK ? rc : set_err(err, BN_OK);
}

int bn_hdist(bn_ctx* ctx, const uint64_t* a, size_t n_words_a, const uint64_t* b, size_t n_words_b, size_t n_bases,
             uint64_t* total, bn_error_t* err) {
    if (!ctx || !total) return set_err(err, BN_ERR_ARGUMENT);
    const size_t need = (n_bases + 31) / 32;
    if (n_words_a < need || n_words_b < need) return set_err(err, BN_INVALID_LENGTH, n_bases);  // multi.rs:124-127
    *total = 0;
    if (n_bases == 0) return set_err(err, BN_OK);
    if (!a || !b) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t per = units_per_chunk(ctx, 16, 2), n_chunks = (need + per - 1) / per;  // words per chunk, 16-byte aligned shards
    const bool a_pg = is_pageable(a), b_pg = is_pageable(b);
    unsigned long long sum = 0;
    const int rc = run_pipeline(
        ctx, n_chunks, err,
        [&](size_t c, int s, cudaStream_t st) -> cudaError_t {
            const size_t w0 = c * per, cnt = std::min(per, need - w0), bases = std::min(n_bases - w0 * 32, cnt * 32);
            BN_TRY(ensure(ctx->stage_in[s], cnt * 8));
            BN_TRY(ensure(ctx->stage_out[s], cnt * 8));
            const void *sa = a + w0, *sb = b + w0;
            BN_TRY(bounce_in(ctx, s, 0, sa, cnt * 8, a_pg));
            BN_TRY(bounce_in(ctx, s, 1, sb, cnt * 8, b_pg));
            BN_TRY(cudaMemcpyAsync(ctx->stage_in[s].p, sa, cnt * 8, cudaMemcpyHostToDevice, st));
            BN_TRY(cudaMemcpyAsync(ctx->stage_out[s].p, sb, cnt * 8, cudaMemcpyHostToDevice, st));
            BN_TRY(bn::launch_hdist(ctx->di, static_cast<const uint64_t*>(ctx->stage_in[s].p), static_cast<const uint64_t*>(ctx->stage_out[s].p),
                                    bases, ctx->d_words + s, st));
            return cudaMemcpyAsync(ctx->h_words + s, ctx->d_words + s, 8, cudaMemcpyDeviceToHost, st);
        },
        [&](size_t, int s) { sum += ctx->h_words[s]; });
    if (rc != BN_OK) return rc;
    *total = sum;
    return set_err(err, BN_OK);
}

int bn_hdist_pairs(bn_ctx* ctx, const uint64_t* u, const uint64_t* v, size_t n_pairs, uint32_t len, uint32_t* out, bn_error_t* err) {
    if (!ctx) return set_err(err, BN_ERR_ARGUMENT);
    if (len > 32) return set_err(err, BN_INVALID_LENGTH, len);  // scalar.rs:13-15
    if (n_pairs == 0) return set_err(err, BN_OK);
    if (!u || !v || !out) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t per = units_per_chunk(ctx, 16, 4), n_chunks = (n_pairs + per - 1) / per;
    const bool u_pg = is_pageable(u), v_pg = is_pageable(v), out_pg = is_pageable(out);
    const int rc = run_pipeline(
        ctx, n_chunks, err,
        [&](size_t c, int s, cudaStream_t st) -> cudaError_t {
            const size_t p0 = c * per, cnt = std::min(per, n_pairs - p0);
            BN_TRY(ensure(ctx->stage_in[s], cnt * 8));
            BN_TRY(ensure(ctx->stage_aux[s][0], cnt * 8));
            BN_TRY(ensure(ctx->stage_out[s], cnt * 4));
            const void *su = u + p0, *sv = v + p0;
            void* dst = out + p0;
            BN_TRY(bounce_in(ctx, s, 0, su, cnt * 8, u_pg));
            BN_TRY(bounce_in(ctx, s, 1, sv, cnt * 8, v_pg));
            BN_TRY(bounce_out(ctx, s, dst, cnt * 4, out_pg));
            BN_TRY(cudaMemcpyAsync(ctx->stage_in[s].p, su, cnt * 8, cudaMemcpyHostToDevice, st));
            BN_TRY(cudaMemcpyAsync(ctx->stage_aux[s][0].p, sv, cnt * 8, cudaMemcpyHostToDevice, st));
            BN_TRY(bn::launch_hdist_pairs(ctx->di, static_cast<const uint64_t*>(ctx->stage_in[s].p),
                                          static_cast<const uint64_t*>(ctx->stage_aux[s][0].p), cnt, len,
                                          static_cast<uint32_t*>(ctx->stage_out[s].p), st));
            return cudaMemcpyAsync(dst, ctx->stage_out[s].p, cnt * 4, cudaMemcpyDeviceToHost, st);
        },
        [&](size_t c, int s) { unbounce(ctx, s, out + c * per, std::min(per, n_pairs - c * per) * 4, out_pg); });
    return rc != BN_OK ? rc : set_err(err, BN_OK);
}

int bn_base_counts(bn_ctx* ctx, const uint64_t* words, size_t n_words, size_t n_bases, uint64_t counts[4], double* gc, bn_error_t* err) {
    if (!ctx || !counts) return set_err(err, BN_ERR_ARGUMENT);
    const size_t need = (n_bases + 31) / 32;
    if (n_words < need) return set_err(err, BN_INVALID_LENGTH, n_bases);
    if (n_bases && !words) return set_err(err, BN_ERR_ARGUMENT);
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    if (gc) *gc = 0.0;
    if (n_bases == 0) return set_err(err, BN_OK);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t per = units_per_chunk(ctx, 8, 2), n_chunks = (need + per - 1) / per;
    const bool in_pg = is_pageable(words);
    const int rc = run_pipeline(
        ctx, n_chunks, err,
        [&](size_t c, int s, cudaStream_t st) -> cudaError_t {
            const size_t w0 = c * per, cnt = std::min(per, need - w0), bases = std::min(n_bases - w0 * 32, cnt * 32);
            BN_TRY(ensure(ctx->stage_in[s], cnt * 8));
            const void* src = words + w0;
            BN_TRY(bounce_in(ctx, s, 0, src, cnt * 8, in_pg));
            BN_TRY(cudaMemcpyAsync(ctx->stage_in[s].p, src, cnt * 8, cudaMemcpyHostToDevice, st));
            BN_TRY(bn::launch_base_counts(ctx->di, static_cast<const uint64_t*>(ctx->stage_in[s].p), bases, ctx->d_words + kStages + 4 * s, nullptr, st));
            return cudaMemcpyAsync(ctx->h_words + kStages + 4 * s, ctx->d_words + kStages + 4 * s, 4 * 8, cudaMemcpyDeviceToHost, st);
        },
        [&](size_t, int s) {
            for (int i = 0; i < 4; ++i) counts[i] += ctx->h_words[kStages + 4 * s + i];
        });
    if (rc != BN_OK) return rc;
    // analysis.rs:14, exactly this operation order on exact integer counts: (gc as f64 / len as f64) * 100.0
    if (gc) {
        volatile double q = (double)(counts[1] + counts[2]) / (double)n_bases;
        *gc = q * 100.0;
    }
    return set_err(err, BN_OK);
}

int bn_base_counts_batch(bn_ctx* ctx, const uint64_t* words, size_t n_words, const uint64_t* word_offsets, const uint64_t* lens,
                         size_t n_reads, uint64_t* counts4, double* gc, uint64_t totals[4], bn_error_t* err) {
    if (!ctx || (n_reads && (!word_offsets || !lens))) return set_err(err, BN_ERR_ARGUMENT);
    // One pass over the read table, in blocks of 65536 reads spread over a few host threads (10 M reads are 160 MB of
    // offsets and lengths: single-threaded, this pass cost more than the PCIe transfers): every read must lie inside
    // `words` (InvalidLength as in decode/hdist); are the reads laid out in order (word offsets never decrease: what
    // bn_encode_batch / bn_fastq_encode produce)?  Where does each block's furthest-reaching read end?
    constexpr size_t kBlockReads = 65536;
    const size_t n_blocks = (n_reads + kBlockReads - 1) / kBlockReads;
    std::vector<uint64_t> block_end(n_blocks, 0);
    const unsigned n_thr = (unsigned)std::max<size_t>(1, std::min<size_t>({8, std::thread::hardware_concurrency(), n_blocks / 4}));
    std::vector<size_t> first_bad(n_thr, SIZE_MAX);
    std::vector<char> ordered(n_thr, 1);
    auto scan_blocks = [&](unsigned t) {
        for (size_t b = n_blocks * t / n_thr; b < n_blocks * (t + 1) / n_thr; ++b) {
            uint64_t end = 0;
            const size_t r1 = std::min(n_reads, (b + 1) * kBlockReads);
            for (size_t r = b * kBlockReads; r < r1; ++r) {
                const uint64_t off = word_offsets[r], need = (lens[r] + 31) / 32;
                if (off > n_words || need > n_words - off) {
                    first_bad[t] = std::min(first_bad[t], r);
                    return;
                }
                if (r && off < word_offsets[r - 1]) ordered[t] = 0;
                if (need) end = std::max(end, off + need);
            }
            block_end[b] = end;
        }
    };
    {
        std::vector<std::thread> workers;
        for (unsigned t = 1; t < n_thr; ++t) workers.emplace_back(scan_blocks, t);
        scan_blocks(0);
        for (auto& w : workers) w.join();
    }
    const size_t bad = *std::min_element(first_bad.begin(), first_bad.end());
    if (bad != SIZE_MAX) {   // the first failing read in index order
        set_err(err, BN_INVALID_LENGTH, lens[bad]);
        if (err) err->record = bad;
        return BN_INVALID_LENGTH;
    }
    const bool in_order = std::all_of(ordered.begin(), ordered.end(), [](char c) { return c != 0; });
    // chunks of whole blocks of about ctx->chunk bytes of words: chunk c = reads [cut[c], cut[c+1]), words up to cut_end[c]
    const size_t chunk_words = ctx->chunk / 8;
    std::vector<size_t> cut{0};
    std::vector<uint64_t> cut_end{0};
    for (size_t b = 0; in_order && b < n_blocks; ++b) {
        const size_t r = b * kBlockReads;
        if (r > cut.back() && std::max(cut_end.back(), block_end[b]) > word_offsets[cut.back()] + chunk_words) {
            cut.push_back(r);
            cut_end.push_back(0);
        }
        cut_end.back() = std::max(cut_end.back(), block_end[b]);
    }
    cut.push_back(n_reads);
    if (totals) totals[0] = totals[1] = totals[2] = totals[3] = 0;
    if (n_reads == 0) return set_err(err, BN_OK);
    if (n_words && !words) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    // Reads in order are cut into chunks of whole reads and run through the 3-stage pipeline: the upload of chunk c+1
    // (40 B of words + 16 B of offsets / lengths per 150 bp read) overlaps the kernel of chunk c and the download of
    // chunk c-1 (32 B + 8 B).
    if (in_order && n_words * 8 > ctx->chunk) {
        uint64_t acc[4] = {0, 0, 0, 0};
        const int rc = run_pipeline(
            ctx, cut.size() - 1, err,
            [&](size_t c, int s, cudaStream_t st) -> cudaError_t {
                const size_t r0 = cut[c], r1 = cut[c + 1], cnt = r1 - r0;
                const uint64_t w0 = word_offsets[r0], w1 = std::max(w0, cut_end[c]);
                const size_t nw = (size_t)(w1 - w0);
                BN_TRY(ensure(ctx->stage_in[s], nw ? nw * 8 : 8));
                BN_TRY(ensure(ctx->stage_aux[s][0], cnt * 8));
                BN_TRY(ensure(ctx->stage_aux[s][1], cnt * 8));
                if (counts4) BN_TRY(ensure(ctx->stage_out[s], cnt * 32));
                if (gc) BN_TRY(ensure(ctx->stage_aux[s][2], cnt * 8));
                if (nw) BN_TRY(cudaMemcpyAsync(ctx->stage_in[s].p, words + w0, nw * 8, cudaMemcpyHostToDevice, st));
                BN_TRY(cudaMemcpyAsync(ctx->stage_aux[s][0].p, word_offsets + r0, cnt * 8, cudaMemcpyHostToDevice, st));
                BN_TRY(cudaMemcpyAsync(ctx->stage_aux[s][1].p, lens + r0, cnt * 8, cudaMemcpyHostToDevice, st));
                unsigned long long* d_tot = ctx->d_words + kStages + 4 * s;
                // the kernel indexes with the caller's absolute word offsets: hand it the chunk's base moved back by w0
                BN_TRY(bn::launch_base_counts_batch(ctx->di, static_cast<const uint64_t*>(ctx->stage_in[s].p) - w0,
                                                    static_cast<const uint64_t*>(ctx->stage_aux[s][0].p),
                                                    static_cast<const uint64_t*>(ctx->stage_aux[s][1].p), cnt, 0, nw,
                                                    counts4 ? static_cast<unsigned long long*>(ctx->stage_out[s].p) : nullptr,
                                                    gc ? static_cast<double*>(ctx->stage_aux[s][2].p) : nullptr, d_tot, st));
                if (counts4) BN_TRY(cudaMemcpyAsync(counts4 + 4 * r0, ctx->stage_out[s].p, cnt * 32, cudaMemcpyDeviceToHost, st));
                if (gc) BN_TRY(cudaMemcpyAsync(gc + r0, ctx->stage_aux[s][2].p, cnt * 8, cudaMemcpyDeviceToHost, st));
                return cudaMemcpyAsync(ctx->h_words + kStages + 4 * s, d_tot, 4 * 8, cudaMemcpyDeviceToHost, st);
            },
            [&](size_t, int s) {
                for (int i = 0; i < 4; ++i) acc[i] += ctx->h_words[kStages + 4 * s + i];
            });
        if (rc != BN_OK) return rc;
        if (totals)
            for (int i = 0; i < 4; ++i) totals[i] = acc[i];
        return set_err(err, BN_OK);
    }
    cudaStream_t st = ctx->stream;
    BN_CUDA(ensure(ctx->slot[0], n_words ? n_words * 8 : 8));
    BN_CUDA(ensure(ctx->slot[1], n_reads * 8));
    BN_CUDA(ensure(ctx->slot[2], n_reads * 8));
    if (counts4) BN_CUDA(ensure(ctx->slot[3], n_reads * 32));
    if (gc) BN_CUDA(ensure(ctx->slot[4], n_reads * 8));
    if (n_words) BN_CUDA(cudaMemcpyAsync(ctx->slot[0].p, words, n_words * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(ctx->slot[1].p, word_offsets, n_reads * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(ctx->slot[2].p, lens, n_reads * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(bn::launch_base_counts_batch(ctx->di, static_cast<const uint64_t*>(ctx->slot[0].p),
                                         static_cast<const uint64_t*>(ctx->slot[1].p), static_cast<const uint64_t*>(ctx->slot[2].p),
                                         n_reads, 0, n_words, counts4 ? static_cast<unsigned long long*>(ctx->slot[3].p) : nullptr,
                                         gc ? static_cast<double*>(ctx->slot[4].p) : nullptr, ctx->d_words + 8, st));
    if (counts4) BN_CUDA(cudaMemcpyAsync(counts4, ctx->slot[3].p, n_reads * 32, cudaMemcpyDeviceToHost, st));
    if (gc) BN_CUDA(cudaMemcpyAsync(gc, ctx->slot[4].p, n_reads * 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 8, ctx->d_words + 8, 4 * 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    if (totals)
        for (int i = 0; i < 4; ++i) totals[i] = ctx->h_words[8 + i];
    return set_err(err, BN_OK);
}

// Variable-length batches are cut into chunks of whole reads (>= ctx->chunk bytes each, so a read longer than the
// chunk is a chunk of its own) and run through the same 3-stage pipeline as bn_encode: the upload of chunk c+1
// overlaps the scan + encode of chunk c and the download of chunk c-1.  Word offsets come back relative to their
// chunk and are rebased on the host; nothing short-circuits (the per-read variant needs every read anyway, and the
// plain variant reports the minimum offset over all chunks = the first invalid base in input order).
int bn_encode_batch(bn_ctx* ctx, const uint8_t* bytes, const uint64_t* offsets, size_t n_reads, uint64_t* out_words,
                    uint64_t* out_word_offsets, uint32_t* read_status, bn_error_t* err) {
    if (!ctx || !out_word_offsets || (n_reads && !offsets)) return set_err(err, BN_ERR_ARGUMENT);
    out_word_offsets[0] = 0;
    if (n_reads == 0) return set_err(err, BN_OK);
    const uint64_t lo = offsets[0], hi = offsets[n_reads];
    struct Chunk {
        size_t r0, r1;      // reads [r0, r1)
        uint64_t w0, nw;    // first output word, number of output words
    };
    std::vector<Chunk> chunks;
    {   // one pass over the offsets: validate, count words, cut
        const size_t want = ctx->chunk;
        Chunk c{0, 0, 0, 0};
        uint64_t w = 0;
        for (size_t r = 0; r < n_reads; ++r) {
            if (offsets[r + 1] < offsets[r]) return set_err(err, BN_ERR_ARGUMENT);
            w += (offsets[r + 1] - offsets[r] + 31) / 32;
            if (offsets[r + 1] - offsets[c.r0] >= want || r + 1 == n_reads) {
                c.r1 = r + 1;
                c.nw = w - c.w0;
                chunks.push_back(c);
                c = Chunk{r + 1, r + 1, w, 0};
            }
        }
    }
    if (hi > lo && (!bytes || !out_words)) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    unsigned long long best = kNoError;  // smallest global (offset << 8 | byte)
    const bool in_pg = is_pageable(bytes), out_pg = is_pageable(out_words);
    auto retire = [&](size_t c) -> cudaError_t {
        const int s = (int)(c % kStages);
        cudaError_t e = cudaEventSynchronize(ctx->stage_done[s]);
        if (e != cudaSuccess) return e;
        unbounce(ctx, s, out_words + chunks[c].w0, chunks[c].nw * 8, out_pg);
        best = std::min<unsigned long long>(best, ctx->h_words[s]);  // device offsets are already global (base pointer trick below)
        if (c) {  // rebase this chunk's word offsets (entry r0 was written by the previous chunk's total, same value)
            const Chunk& ch = chunks[c];
            for (size_t r = ch.r0 + 1; r <= ch.r1; ++r) out_word_offsets[r] += ch.w0;
        }
        return cudaSuccess;
    };
    for (size_t c = 0; c < chunks.size(); ++c) {
        const Chunk& ch = chunks[c];
        const int s = (int)(c % kStages);
        cudaStream_t st = ctx->stage_stream[s];
        if (c >= (size_t)kStages) BN_CUDA(retire(c - kStages));
        const size_t nr = ch.r1 - ch.r0;
        const uint64_t b0 = offsets[ch.r0], b1 = offsets[ch.r1];
        const size_t phase = b0 & 15u;  // staged at the same 16-byte phase as bytes + b0, so the offsets are used unchanged
        BN_CUDA(ensure(ctx->stage_in[s], (b1 - b0) + phase + 16));
        BN_CUDA(ensure(ctx->stage_out[s], ch.nw * 8 + 8));
        BN_CUDA(ensure(ctx->stage_aux[s][0], (nr + 1) * 8));
        BN_CUDA(ensure(ctx->stage_aux[s][1], (nr + 1) * 8));
        if (read_status) BN_CUDA(ensure(ctx->stage_aux[s][2], nr * 4));
        BN_CUDA(ensure(ctx->stage_aux[s][3], bn::encode_batch_scratch_bytes(nr, b1 - b0)));
        uint8_t* d_bytes = static_cast<uint8_t*>(ctx->stage_in[s].p) + phase;
        unsigned long long* d_status = ctx->d_words + s;
        const void* src = bytes + b0;
        void* dst = out_words + ch.w0;
        BN_CUDA(bounce_in(ctx, s, 0, src, b1 - b0, in_pg));
        BN_CUDA(bounce_out(ctx, s, dst, ch.nw * 8, out_pg));
        if (b1 > b0) BN_CUDA(cudaMemcpyAsync(d_bytes, src, b1 - b0, cudaMemcpyHostToDevice, st));
        BN_CUDA(cudaMemcpyAsync(ctx->stage_aux[s][0].p, offsets + ch.r0, (nr + 1) * 8, cudaMemcpyHostToDevice, st));
        BN_CUDA(bn::launch_encode_batch(ctx->di, d_bytes - b0, static_cast<const uint64_t*>(ctx->stage_aux[s][0].p), nr, b1 - b0,
                                        static_cast<uint64_t*>(ctx->stage_out[s].p), static_cast<uint64_t*>(ctx->stage_aux[s][1].p),
                                        read_status ? static_cast<uint32_t*>(ctx->stage_aux[s][2].p) : nullptr, d_status,
                                        ctx->stage_aux[s][3].p, st));
        if (ch.nw) BN_CUDA(cudaMemcpyAsync(dst, ctx->stage_out[s].p, ch.nw * 8, cudaMemcpyDeviceToHost, st));
        // entries r0+1 .. r1 (entry r0 is the previous chunk's last entry; chunk 0 writes its own zero too)
        BN_CUDA(cudaMemcpyAsync(out_word_offsets + ch.r0 + 1, static_cast<uint64_t*>(ctx->stage_aux[s][1].p) + 1, nr * 8,
                                cudaMemcpyDeviceToHost, st));
        if (read_status) BN_CUDA(cudaMemcpyAsync(read_status + ch.r0, ctx->stage_aux[s][2].p, nr * 4, cudaMemcpyDeviceToHost, st));
        BN_CUDA(cudaMemcpyAsync(ctx->h_words + s, d_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        BN_CUDA(cudaEventRecord(ctx->stage_done[s], st));
    }
    for (size_t c = chunks.size() > (size_t)kStages ? chunks.size() - kStages : 0; c < chunks.size(); ++c) BN_CUDA(retire(c));
    if (best != kNoError) {
        invalid_base(err, best, 0);
        if (err) {
            const uint64_t off = best >> 8;
            const size_t r = (size_t)(std::upper_bound(offsets, offsets + n_reads + 1, off) - offsets) - 1;
            err->record = r;
            err->b = off - offsets[r];  // position inside the read
        }
        return BN_INVALID_BASE;
    }
    return set_err(err, BN_OK);
}

int bn_split_packed_batch(bn_ctx* ctx, const uint64_t* words, size_t n_words, const uint64_t* word_offsets, const uint64_t* lens,
                          const uint64_t* idx, size_t n_reads, uint64_t* left, uint64_t* left_offsets, uint64_t* right,
                          uint64_t* right_offsets, bn_error_t* err) {
    if (!ctx || !left_offsets || !right_offsets || (n_reads && (!word_offsets || !lens || !idx))) return set_err(err, BN_ERR_ARGUMENT);
    // The first failing read in index order, as the caller's loop with `?` would report.  kind 1: idx > len (split.rs:22-27);
    // 2: a malformed read table; 3: an ebuf too short for slen (the reference panics at split.rs:77 or truncates the right half).
    {
        int kind = 0;
        const size_t r = parallel_first_failing(n_reads, [&](size_t q, unsigned) {
            if (idx[q] > lens[q]) return 1;
            if (word_offsets[q + 1] < word_offsets[q] || word_offsets[q + 1] > n_words) return 2;
            const uint64_t have = word_offsets[q + 1] - word_offsets[q];
            return idx[q] && idx[q] < lens[q] && have && have < (lens[q] + 31) / 32 ? 3 : 0;
        }, &kind);
        if (r != SIZE_MAX) {
            if (kind == 2) return set_err(err, BN_ERR_ARGUMENT);
            if (kind == 1) set_err(err, BN_INDEX_OUT_OF_BOUNDS, idx[r], lens[r]);
            else set_err(err, BN_INVALID_LENGTH, lens[r]);
            if (err) err->record = r;
            return kind == 1 ? BN_INDEX_OUT_OF_BOUNDS : BN_INVALID_LENGTH;
        }
    }
    left_offsets[0] = right_offsets[0] = 0;
    if (n_reads == 0) return set_err(err, BN_OK);
    if (n_words && (!words || !left || !right)) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    cudaStream_t st = ctx->stream;
    BN_CUDA(ensure(ctx->slot[0], n_words ? n_words * 8 : 8));
    BN_CUDA(ensure(ctx->slot[1], (n_reads + 1) * 8));
    BN_CUDA(ensure(ctx->slot[2], n_reads * 8));
    BN_CUDA(ensure(ctx->slot[3], n_reads * 8));
    BN_CUDA(ensure(ctx->slot[4], (n_words + n_reads) * 8 + 8));
    BN_CUDA(ensure(ctx->slot[5], n_words * 8 + 8));
    BN_CUDA(ensure(ctx->slot[6], 2 * (n_reads + 1) * 8));
    BN_CUDA(ensure(ctx->slot[7], bn::split_packed_scratch_bytes(n_reads)));
    uint64_t* d_lo = static_cast<uint64_t*>(ctx->slot[6].p);
    uint64_t* d_ro = d_lo + n_reads + 1;
    if (n_words) BN_CUDA(cudaMemcpyAsync(ctx->slot[0].p, words, n_words * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(ctx->slot[1].p, word_offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(ctx->slot[2].p, lens, n_reads * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(ctx->slot[3].p, idx, n_reads * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(bn::launch_split_packed_batch(ctx->di, static_cast<const uint64_t*>(ctx->slot[0].p), static_cast<const uint64_t*>(ctx->slot[1].p),
                                          static_cast<const uint64_t*>(ctx->slot[2].p), static_cast<const uint64_t*>(ctx->slot[3].p), n_reads,
                                          static_cast<uint64_t*>(ctx->slot[4].p), d_lo, static_cast<uint64_t*>(ctx->slot[5].p), d_ro,
                                          ctx->d_words + 8, ctx->slot[7].p, st));
    BN_CUDA(cudaMemcpyAsync(left_offsets, d_lo, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaMemcpyAsync(right_offsets, d_ro, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    if (left_offsets[n_reads]) BN_CUDA(cudaMemcpyAsync(left, ctx->slot[4].p, left_offsets[n_reads] * 8, cudaMemcpyDeviceToHost, st));
    if (right_offsets[n_reads]) BN_CUDA(cudaMemcpyAsync(right, ctx->slot[5].p, right_offsets[n_reads] * 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    return set_err(err, BN_OK);
}

// get / slice: the packed batch and the query arrays are staged whole; validation happens on the host in query order
// (the caller's loop with `?` stops at the first failing query), so the device never sees a bad query here.
static int stage_packed_batch(bn_ctx* ctx, const uint64_t* words, size_t n_words, const uint64_t* word_offsets, const uint64_t* lens,
                              size_t n_reads, cudaStream_t st, bn_error_t* err) {
    {
        int kind = 0;
        const size_t r = parallel_first_failing(n_reads, [&](size_t q, unsigned) {
            return word_offsets[q] > n_words || (lens[q] + 31) / 32 > n_words - word_offsets[q] ? 1 : 0;
        }, &kind);
        if (r != SIZE_MAX) {
            set_err(err, BN_INVALID_LENGTH, lens[r]);
            if (err) err->record = r;
            return BN_INVALID_LENGTH;
        }
    }
    BN_CUDA(ensure(ctx->slot[0], n_words ? n_words * 8 : 8));
    BN_CUDA(ensure(ctx->slot[1], n_reads ? n_reads * 8 : 8));
    BN_CUDA(ensure(ctx->slot[2], n_reads ? n_reads * 8 : 8));
    if (n_words) BN_CUDA(cudaMemcpyAsync(ctx->slot[0].p, words, n_words * 8, cudaMemcpyHostToDevice, st));
    if (n_reads) {
        BN_CUDA(cudaMemcpyAsync(ctx->slot[1].p, word_offsets, n_reads * 8, cudaMemcpyHostToDevice, st));
        BN_CUDA(cudaMemcpyAsync(ctx->slot[2].p, lens, n_reads * 8, cudaMemcpyHostToDevice, st));
    }
    return BN_OK;
}

int bn_slice_batch(bn_ctx* ctx, const uint64_t* words, size_t n_words, const uint64_t* word_offsets, const uint64_t* lens, size_t n_reads,
                   const uint64_t* q_read, const uint64_t* q_start, const uint64_t* q_end, size_t nq, uint8_t* out, size_t out_cap,
                   uint64_t* out_offsets, bn_error_t* err) {
    if (!ctx || !out_offsets || (n_reads && (!word_offsets || !lens)) || (nq && (!q_read || !q_start || !q_end)))
        return set_err(err, BN_ERR_ARGUMENT);
    size_t total = 0;
    {
        size_t sums[kHostCheckThreads] = {};
        int kind = 0;
        const size_t bad = parallel_first_failing(nq, [&](size_t q, unsigned t) {
            if (q_read[q] >= n_reads) return 2;
            if (q_start[q] > q_end[q] || q_end[q] > lens[q_read[q]]) return 1;  // sequence.rs:199-205
            sums[t] += q_end[q] - q_start[q];
            return 0;
        }, &kind);
        if (bad != SIZE_MAX) {
            if (kind == 2) return set_err(err, BN_ERR_ARGUMENT);
            set_err(err, BN_INVALID_RANGE, q_start[bad], q_end[bad], lens[q_read[bad]]);
            if (err) err->record = bad;
            return BN_INVALID_RANGE;
        }
        for (size_t v : sums) total += v;
    }
    out_offsets[0] = 0;
    if (nq == 0) return set_err(err, BN_OK);
    if (total > out_cap || (total && !out)) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    cudaStream_t st = ctx->stream;
    const int rc = stage_packed_batch(ctx, words, n_words, word_offsets, lens, n_reads, st, err);
    if (rc != BN_OK) return rc;
    BN_CUDA(ensure(ctx->slot[3], 3 * nq * 8));
    BN_CUDA(ensure(ctx->slot[4], total + 16));
    BN_CUDA(ensure(ctx->slot[5], (nq + 1) * 8));
    BN_CUDA(ensure(ctx->slot[6], bn::slice_batch_scratch_bytes(nq)));
    uint64_t* dq = static_cast<uint64_t*>(ctx->slot[3].p);
    BN_CUDA(cudaMemcpyAsync(dq, q_read, nq * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(dq + nq, q_start, nq * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(dq + 2 * nq, q_end, nq * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(bn::launch_slice_batch(ctx->di, static_cast<const uint64_t*>(ctx->slot[0].p), static_cast<const uint64_t*>(ctx->slot[1].p),
                                   static_cast<const uint64_t*>(ctx->slot[2].p), n_reads, dq, dq + nq, dq + 2 * nq, nq,
                                   static_cast<uint8_t*>(ctx->slot[4].p), static_cast<uint64_t*>(ctx->slot[5].p), ctx->d_words + 8,
                                   ctx->slot[6].p, st));
    BN_CUDA(cudaMemcpyAsync(out_offsets, ctx->slot[5].p, (nq + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (total) BN_CUDA(cudaMemcpyAsync(out, ctx->slot[4].p, total, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    return set_err(err, BN_OK);
}

int bn_get_batch(bn_ctx* ctx, const uint64_t* words, size_t n_words, const uint64_t* word_offsets, const uint64_t* lens, size_t n_reads,
                 const uint64_t* q_read, const uint64_t* q_index, size_t nq, uint8_t* out, bn_error_t* err) {
    if (!ctx || (n_reads && (!word_offsets || !lens)) || (nq && (!q_read || !q_index || !out))) return set_err(err, BN_ERR_ARGUMENT);
    {
        int kind = 0;
        const size_t bad = parallel_first_failing(nq, [&](size_t q, unsigned) {
            if (q_read[q] >= n_reads) return 2;
            return q_index[q] >= lens[q_read[q]] ? 1 : 0;  // sequence.rs:117-122
        }, &kind);
        if (bad != SIZE_MAX) {
            if (kind == 2) return set_err(err, BN_ERR_ARGUMENT);
            set_err(err, BN_INDEX_OUT_OF_BOUNDS, q_index[bad], lens[q_read[bad]]);
            if (err) err->record = bad;
            return BN_INDEX_OUT_OF_BOUNDS;
        }
    }
    if (nq == 0) return set_err(err, BN_OK);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    cudaStream_t st = ctx->stream;
    const int rc = stage_packed_batch(ctx, words, n_words, word_offsets, lens, n_reads, st, err);
    if (rc != BN_OK) return rc;
    BN_CUDA(ensure(ctx->slot[3], 2 * nq * 8));
    BN_CUDA(ensure(ctx->slot[4], nq));
    uint64_t* dq = static_cast<uint64_t*>(ctx->slot[3].p);
    BN_CUDA(cudaMemcpyAsync(dq, q_read, nq * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(dq + nq, q_index, nq * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(bn::launch_get_batch(ctx->di, static_cast<const uint64_t*>(ctx->slot[0].p), static_cast<const uint64_t*>(ctx->slot[1].p),
                                 static_cast<const uint64_t*>(ctx->slot[2].p), n_reads, dq, dq + nq, nq, static_cast<uint8_t*>(ctx->slot[4].p),
                                 ctx->d_words + 8, st));
    BN_CUDA(cudaMemcpyAsync(out, ctx->slot[4].p, nq, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    return set_err(err, BN_OK);
}

int bn_kmers(bn_ctx* ctx, const uint8_t* seq, size_t n, uint32_t k, uint64_t* out, size_t* n_out, bn_error_t* err) {
    if (n_out) *n_out = 0;
    if (!ctx || k == 0) return set_err(err, BN_ERR_ARGUMENT);
    if (n < k) return set_err(err, BN_OK);                         // windows(k) yields nothing
    if (k > 32) return set_err(err, BN_SEQUENCE_TOO_LONG, k);      // the first window already fails (naive.rs:5-7)
    if (!seq || !out) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    const size_t n_win = n - k + 1;
    const size_t per = units_per_chunk(ctx, 8, 2048), n_chunks = (n_win + per - 1) / per;  // the output (8 B per window) dominates
    const bool in_pg = is_pageable(seq), out_pg = is_pageable(out);
    unsigned long long best = kNoError;
    const int rc = run_pipeline(
        ctx, n_chunks, err,
        [&](size_t c, int s, cudaStream_t st) -> cudaError_t {
            const size_t i0 = c * per, cnt = std::min(per, n_win - i0), in_bytes = cnt + k - 1;  // chunks share k-1 bytes of input
            BN_TRY(ensure(ctx->stage_in[s], in_bytes));
            BN_TRY(ensure(ctx->stage_out[s], cnt * 8));
            const void* src = seq + i0;
            void* dst = out + i0;
            BN_TRY(bounce_in(ctx, s, 0, src, in_bytes, in_pg));
            BN_TRY(bounce_out(ctx, s, dst, cnt * 8, out_pg));
            BN_TRY(cudaMemcpyAsync(ctx->stage_in[s].p, src, in_bytes, cudaMemcpyHostToDevice, st));
            BN_TRY(bn::launch_kmer_windows(ctx->di, static_cast<const uint8_t*>(ctx->stage_in[s].p), in_bytes, k,
                                           static_cast<uint64_t*>(ctx->stage_out[s].p), ctx->d_words + s, st));
            BN_TRY(cudaMemcpyAsync(dst, ctx->stage_out[s].p, cnt * 8, cudaMemcpyDeviceToHost, st));
            return cudaMemcpyAsync(ctx->h_words + s, ctx->d_words + s, 8, cudaMemcpyDeviceToHost, st);
        },
        [&](size_t c, int s) {
            unbounce(ctx, s, out + c * per, std::min(per, n_win - c * per) * 8, out_pg);
            const unsigned long long key = ctx->h_words[s];
            if (key != kNoError) best = std::min(best, (((key >> 8) + (unsigned long long)(c * per)) << 8) | (key & 0xFFu));
        });
    if (rc != BN_OK) return rc;
    if (best != kNoError) {  // the windows before the failing one are what the caller's loop had produced: they are in `out`
        const size_t off = (size_t)(best >> 8), first_bad = off >= k - 1 ? off - (k - 1) : 0;
        if (n_out) *n_out = first_bad;
        invalid_base(err, best, 0);
        if (err) err->record = first_bad;
        return BN_INVALID_BASE;
    }
    if (n_out) *n_out = n_win;
    return set_err(err, BN_OK);
}

int bn_kmers_batch(bn_ctx* ctx, const uint8_t* bytes, const uint64_t* offsets, size_t n_reads, uint32_t k, uint64_t* out, size_t out_cap,
                   uint64_t* out_offsets, bn_error_t* err) {
    if (!ctx || !out_offsets || k == 0 || (n_reads && !offsets)) return set_err(err, BN_ERR_ARGUMENT);
    out_offsets[0] = 0;
    if (n_reads == 0) return set_err(err, BN_OK);
    size_t total = 0;
    for (size_t r = 0; r < n_reads; ++r) {
        if (offsets[r + 1] < offsets[r]) return set_err(err, BN_ERR_ARGUMENT);
        const uint64_t len = offsets[r + 1] - offsets[r];
        if (len >= k) {
            if (k > 32) {  // the first read with a window fails on its first window (naive.rs:5-7)
                set_err(err, BN_SEQUENCE_TOO_LONG, k);
                if (err) err->record = r;
                return BN_SEQUENCE_TOO_LONG;
            }
            total += len - k + 1;
        }
    }
    const uint64_t lo = offsets[0], hi = offsets[n_reads];
    if (total > out_cap || (total && (!bytes || !out))) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    cudaStream_t st = ctx->stream;
    const size_t phase = lo & 15u;  // staged at the same 16-byte phase as bytes + lo, so the offsets are used unchanged
    BN_CUDA(ensure(ctx->slot[0], (hi - lo) + phase + 16));
    BN_CUDA(ensure(ctx->slot[1], (n_reads + 1) * 8));
    BN_CUDA(ensure(ctx->slot[2], total * 8 + 8));
    BN_CUDA(ensure(ctx->slot[3], (n_reads + 1) * 8));
    BN_CUDA(ensure(ctx->slot[4], bn::kmer_windows_batch_scratch_bytes(n_reads, hi - lo)));
    uint8_t* d_bytes = static_cast<uint8_t*>(ctx->slot[0].p) + phase;
    if (hi > lo) BN_CUDA(cudaMemcpyAsync(d_bytes, bytes + lo, hi - lo, cudaMemcpyHostToDevice, st));
    BN_CUDA(cudaMemcpyAsync(ctx->slot[1].p, offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    BN_CUDA(bn::launch_kmer_windows_batch(ctx->di, d_bytes - lo, static_cast<const uint64_t*>(ctx->slot[1].p), n_reads, hi - lo, k,
                                          static_cast<uint64_t*>(ctx->slot[2].p), static_cast<uint64_t*>(ctx->slot[3].p), ctx->d_words + 8,
                                          ctx->slot[4].p, st));
    BN_CUDA(cudaMemcpyAsync(out_offsets, ctx->slot[3].p, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (total) BN_CUDA(cudaMemcpyAsync(out, ctx->slot[2].p, total * 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 8, ctx->d_words + 8, 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    const unsigned long long key = ctx->h_words[8];
    if (key != kNoError) {
        invalid_base(err, key, 0);
        if (err) {
            const uint64_t off = key >> 8;
            const size_t r = (size_t)(std::upper_bound(offsets, offsets + n_reads + 1, off) - offsets) - 1;
            err->record = r;
            err->b = off - offsets[r];
        }
        return BN_INVALID_BASE;
    }
    return set_err(err, BN_OK);
}

}  // extern "C"

namespace {

// Upload of a whole host buffer on the context stream.  Pageable caller memory goes through two pinned stage buffers in
// chunks (multi-threaded memcpy of chunk c+1 while chunk c is on the link) instead of the driver's single-threaded staging.
int upload_whole(bn_ctx* ctx, void* d_dst, const uint8_t* src, size_t bytes, bn_error_t* err) {
    cudaStream_t st = ctx->stream;
    if (!is_pageable(src)) {
        BN_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, st));
        return BN_OK;
    }
    const size_t chunk = ctx->chunk;
    for (int s = 0; s < 2; ++s) BN_CUDA(ensure_host(ctx->hstage_in[s][0], std::min(chunk, bytes)));
    size_t c = 0;
    for (size_t off = 0; off < bytes; off += chunk, ++c) {
        const int s = (int)(c & 1);
        const size_t len = std::min(chunk, bytes - off);
        if (c >= 2) BN_CUDA(cudaEventSynchronize(ctx->stage_done[s]));   // the copy that last used this stage buffer
        parallel_memcpy(ctx->hstage_in[s][0].p, src + off, len);
        BN_CUDA(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, ctx->hstage_in[s][0].p, len, cudaMemcpyHostToDevice, st));
        BN_CUDA(cudaEventRecord(ctx->stage_done[s], st));
    }
    return BN_OK;
}

}  // namespace

extern "C" {

// FASTQ text -> records -> packed words.  The text is staged whole (pageable text in chunks through pinned buffers);
// scan and encode are two calls because the caller has to allocate the outputs in between.
static int fastx_scan(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t* n_reads, size_t* n_words, bn_error_t* err, int fasta) {
    if (!ctx || !n_reads || !n_words || (n_bytes && !text)) return set_err(err, BN_ERR_ARGUMENT);
    const unsigned lpr = fasta ? 2u : 4u;   // lines per record
    *n_reads = *n_words = 0;
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    ctx->fq_valid = false;
    ctx->fq_text = text;
    ctx->fq_bytes = n_bytes;
    ctx->fq_reads = ctx->fq_words = 0;
    ctx->fq_fasta = fasta;
    if (n_bytes == 0) {
        ctx->fq_valid = true;
        return set_err(err, BN_OK);
    }
    cudaStream_t st = ctx->stream;
    BN_CUDA(ensure(ctx->fq[0], n_bytes + 16));
    BN_CUDA(ensure(ctx->fq[1], bn::fastq_scratch_bytes(n_bytes)));
    const uint8_t* d_text = static_cast<const uint8_t*>(ctx->fq[0].p);
    if (const int rc = upload_whole(ctx, ctx->fq[0].p, text, n_bytes, err)) return rc;
    BN_CUDA(bn::launch_fastq_count(ctx->di, d_text, n_bytes, ctx->fq[1].p, reinterpret_cast<uint64_t*>(ctx->d_words + 10), fasta, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 10, ctx->d_words + 10, 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    const unsigned long long n_lines = ctx->h_words[10];
    const size_t nr = (size_t)(n_lines / lpr);
    BN_CUDA(ensure(ctx->fq[2], bn::fastq_index_scratch_bytes(nr)));
    BN_CUDA(ensure(ctx->fq[3], nr * 8 + 8));
    BN_CUDA(ensure(ctx->fq[4], nr * 8 + 8));
    BN_CUDA(ensure(ctx->fq[5], (nr + 1) * 8));
    uint64_t* d_wo = static_cast<uint64_t*>(ctx->fq[5].p);
    BN_CUDA(bn::launch_fastq_index(ctx->di, d_text, n_bytes, nr, ctx->fq[1].p, ctx->fq[2].p, static_cast<uint64_t*>(ctx->fq[3].p),
                                   static_cast<uint64_t*>(ctx->fq[4].p), d_wo, ctx->d_words + 12, fasta, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 11, d_wo + nr, 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 12, ctx->d_words + 12, 16, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    unsigned long long fault = ctx->h_words[13];
    if (n_lines % lpr) fault = std::min<unsigned long long>(fault, ((unsigned long long)nr << 8) | BN_FASTQ_TRUNCATED);
    if (fault != kNoError) return fastq_fault(err, fault);
    ctx->fq_reads = nr;
    ctx->fq_words = (size_t)ctx->h_words[11];
    ctx->fq_valid = true;
    *n_reads = nr;
    *n_words = ctx->fq_words;
    return set_err(err, BN_OK);
}

static int fastx_encode(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t n_reads, size_t n_words, uint64_t* out_words,
                        uint64_t* out_word_offsets, uint64_t* seq_offsets, uint64_t* seq_lens, bn_error_t* err, int fasta) {
    if (!ctx) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    // only valid right after bn_fastq_scan of the same text on this context
    if (!ctx->fq_valid || ctx->fq_fasta != fasta || ctx->fq_text != text || ctx->fq_bytes != n_bytes || ctx->fq_reads != n_reads ||
        ctx->fq_words != n_words ||
        (n_words && !out_words))
        return set_err(err, BN_ERR_ARGUMENT);
    if (n_reads == 0) {
        if (out_word_offsets) out_word_offsets[0] = 0;
        return set_err(err, BN_OK);
    }
    cudaStream_t st = ctx->stream;
    BN_CUDA(ensure(ctx->fq[6], n_words * 8 + 8));
    const uint64_t* d_so = static_cast<const uint64_t*>(ctx->fq[3].p);
    BN_CUDA(cudaMemsetAsync(ctx->d_words + 12, 0xFF, 8, st));
    BN_CUDA(bn::launch_fastq_encode(ctx->di, static_cast<const uint8_t*>(ctx->fq[0].p), n_bytes, n_reads, ctx->fq[1].p, d_so,
                                    static_cast<const uint64_t*>(ctx->fq[4].p), static_cast<const uint64_t*>(ctx->fq[5].p),
                                    static_cast<uint64_t*>(ctx->fq[6].p), ctx->d_words + 12, fasta, st));
    if (n_words) BN_CUDA(cudaMemcpyAsync(out_words, ctx->fq[6].p, n_words * 8, cudaMemcpyDeviceToHost, st));
    if (out_word_offsets) BN_CUDA(cudaMemcpyAsync(out_word_offsets, ctx->fq[5].p, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (seq_offsets) BN_CUDA(cudaMemcpyAsync(seq_offsets, ctx->fq[3].p, n_reads * 8, cudaMemcpyDeviceToHost, st));
    if (seq_lens) BN_CUDA(cudaMemcpyAsync(seq_lens, ctx->fq[4].p, n_reads * 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 12, ctx->d_words + 12, 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    const unsigned long long key = ctx->h_words[12];
    if (key != kNoError) {
        invalid_base(err, key, 0);
        if (err) {
            std::vector<uint64_t> so(n_reads);
            BN_CUDA(cudaMemcpyAsync(so.data(), d_so, n_reads * 8, cudaMemcpyDeviceToHost, st));
            BN_CUDA(cudaStreamSynchronize(st));
            const uint64_t off = key >> 8;
            const size_t r = (size_t)(std::upper_bound(so.begin(), so.end(), off) - so.begin()) - 1;
            err->record = r;
            err->b = off - so[r];
        }
        return BN_INVALID_BASE;
    }
    return set_err(err, BN_OK);
}

// Wrapped (multi-line) FASTA: '>' header line, then any number of sequence lines per record.  bn_fasta_wrapped_scan uploads
// the text and does all the device work (line table, compaction of the sequence bytes, batch encode of the records); the
// results stay resident on the context and bn_fasta_wrapped_encode copies them out into buffers the caller sized from the scan.
int bn_fasta_wrapped_scan(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t* n_records, size_t* n_bases, size_t* n_words, bn_error_t* err) {
    if (!ctx || !n_records || !n_bases || !n_words || (n_bytes && !text)) return set_err(err, BN_ERR_ARGUMENT);
    *n_records = *n_bases = *n_words = 0;
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    ctx->fq_valid = false;
    ctx->fq_text = text;
    ctx->fq_bytes = n_bytes;
    ctx->fq_reads = ctx->fq_words = 0;
    ctx->fq_fasta = 2;   // wrapped
    if (n_bytes == 0) {
        ctx->fq_valid = true;
        return set_err(err, BN_OK);
    }
    cudaStream_t st = ctx->stream;
    BN_CUDA(ensure(ctx->fq[0], n_bytes + 16));
    BN_CUDA(ensure(ctx->fq[1], bn::fastq_scratch_bytes(n_bytes)));
    const uint8_t* d_text = static_cast<const uint8_t*>(ctx->fq[0].p);
    if (const int rc = upload_whole(ctx, ctx->fq[0].p, text, n_bytes, err)) return rc;
    BN_CUDA(bn::launch_fastq_count(ctx->di, d_text, n_bytes, ctx->fq[1].p, reinterpret_cast<uint64_t*>(ctx->d_words + 10), 1, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 10, ctx->d_words + 10, 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    const size_t n_lines = (size_t)ctx->h_words[10];
    BN_CUDA(ensure(ctx->fq[2], bn::fasta_wrapped_scratch_bytes(n_lines)));
    BN_CUDA(bn::launch_fasta_wrapped_index(ctx->di, d_text, n_bytes, n_lines, ctx->fq[1].p, ctx->fq[2].p,
                                           reinterpret_cast<uint64_t*>(ctx->d_words + 14), ctx->d_words + 12, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 12, ctx->d_words + 12, 32, cudaMemcpyDeviceToHost, st));   // status pair, records, bases
    BN_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_words[13] != kNoError) return fastq_fault(err, ctx->h_words[13]);
    const size_t nr = (size_t)ctx->h_words[14], nb = (size_t)ctx->h_words[15];
    // fq[3] header offsets | fq[4] record offsets into the compacted sequence | fq[5] word offsets | fq[6] words; slot[0] the compacted bytes
    BN_CUDA(ensure(ctx->fq[3], nr * 8 + 8));
    BN_CUDA(ensure(ctx->fq[4], (nr + 1) * 8));
    BN_CUDA(ensure(ctx->fq[5], (nr + 1) * 8));
    BN_CUDA(ensure(ctx->fq[6], (nb / 32 + nr) * 8 + 8));
    BN_CUDA(ensure(ctx->slot[0], nb + 16));
    BN_CUDA(ensure(ctx->slot[7], bn::encode_batch_scratch_bytes(nr, nb)));
    BN_CUDA(bn::launch_fasta_wrapped_compact(ctx->di, d_text, n_lines, ctx->fq[2].p, nr, static_cast<uint8_t*>(ctx->slot[0].p),
                                             static_cast<uint64_t*>(ctx->fq[4].p), static_cast<uint64_t*>(ctx->fq[3].p), st));
    BN_CUDA(bn::launch_encode_batch(ctx->di, static_cast<const uint8_t*>(ctx->slot[0].p), static_cast<const uint64_t*>(ctx->fq[4].p), nr, nb,
                                    static_cast<uint64_t*>(ctx->fq[6].p), static_cast<uint64_t*>(ctx->fq[5].p), nullptr, ctx->d_words + 12,
                                    ctx->slot[7].p, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 11, static_cast<uint64_t*>(ctx->fq[5].p) + nr, 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaMemcpyAsync(ctx->h_words + 12, ctx->d_words + 12, 8, cudaMemcpyDeviceToHost, st));
    BN_CUDA(cudaStreamSynchronize(st));
    const unsigned long long key = ctx->h_words[12];
    if (key != kNoError) {   // the first invalid base in file order: record and position inside the record's sequence
        invalid_base(err, key, 0);
        if (err && nr) {
            std::vector<uint64_t> ro(nr + 1);
            BN_CUDA(cudaMemcpyAsync(ro.data(), ctx->fq[4].p, (nr + 1) * 8, cudaMemcpyDeviceToHost, st));
            BN_CUDA(cudaStreamSynchronize(st));
            const uint64_t off = key >> 8;   // offset in the concatenated sequence bytes
            const size_t r = (size_t)(std::upper_bound(ro.begin(), ro.end(), off) - ro.begin()) - 1;
            err->record = r;
            err->b = off - ro[r];
        }
        return BN_INVALID_BASE;
    }
    ctx->fq_reads = nr;
    ctx->fq_words = nr ? (size_t)ctx->h_words[11] : 0;
    ctx->fq_valid = true;
    *n_records = nr;
    *n_bases = nb;
    *n_words = ctx->fq_words;
    return set_err(err, BN_OK);
}

int bn_fasta_wrapped_encode(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t n_records, size_t n_words, uint64_t* out_words,
                            uint64_t* out_word_offsets, uint64_t* header_offsets, uint64_t* seq_lens, bn_error_t* err) {
    if (!ctx) return set_err(err, BN_ERR_ARGUMENT);
    DeviceGuard g(ctx->di.device);
    std::lock_guard<std::mutex> lk(ctx->mu);
    StageDrain drain{ctx};
    // only valid right after bn_fasta_wrapped_scan of the same text on this context
    if (!ctx->fq_valid || ctx->fq_fasta != 2 || ctx->fq_text != text || ctx->fq_bytes != n_bytes || ctx->fq_reads != n_records ||
        ctx->fq_words != n_words || (n_words && !out_words))
        return set_err(err, BN_ERR_ARGUMENT);
    if (n_records == 0) {
        if (out_word_offsets) out_word_offsets[0] = 0;
        return set_err(err, BN_OK);
    }
    cudaStream_t st = ctx->stream;
    if (n_words) BN_CUDA(cudaMemcpyAsync(out_words, ctx->fq[6].p, n_words * 8, cudaMemcpyDeviceToHost, st));
    if (out_word_offsets) BN_CUDA(cudaMemcpyAsync(out_word_offsets, ctx->fq[5].p, (n_records + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (header_offsets) BN_CUDA(cudaMemcpyAsync(header_offsets, ctx->fq[3].p, n_records * 8, cudaMemcpyDeviceToHost, st));
    std::vector<uint64_t> ro;
    if (seq_lens) {
        ro.resize(n_records + 1);
        BN_CUDA(cudaMemcpyAsync(ro.data(), ctx->fq[4].p, (n_records + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    BN_CUDA(cudaStreamSynchronize(st));
    if (seq_lens)
        for (size_t r = 0; r < n_records; ++r) seq_lens[r] = ro[r + 1] - ro[r];
    return set_err(err, BN_OK);
}

int bn_fastq_scan(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t* n_reads, size_t* n_words, bn_error_t* err) {
    return fastx_scan(ctx, text, n_bytes, n_reads, n_words, err, 0);
}
int bn_fasta_scan(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t* n_reads, size_t* n_words, bn_error_t* err) {
    return fastx_scan(ctx, text, n_bytes, n_reads, n_words, err, 1);
}
int bn_fastq_encode(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t n_reads, size_t n_words, uint64_t* out_words,
                    uint64_t* out_word_offsets, uint64_t* seq_offsets, uint64_t* seq_lens, bn_error_t* err) {
    return fastx_encode(ctx, text, n_bytes, n_reads, n_words, out_words, out_word_offsets, seq_offsets, seq_lens, err, 0);
}
int bn_fasta_encode(bn_ctx* ctx, const uint8_t* text, size_t n_bytes, size_t n_reads, size_t n_words, uint64_t* out_words,
                    uint64_t* out_word_offsets, uint64_t* seq_offsets, uint64_t* seq_lens, bn_error_t* err) {
    return fastx_encode(ctx, text, n_bytes, n_reads, n_words, out_words, out_word_offsets, seq_offsets, seq_lens, err, 1);
}

}  // extern "C"
