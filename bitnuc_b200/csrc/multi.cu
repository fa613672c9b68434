// multi.cu -- bn_multi: one process, N devices (see include/bitnuc_cuda.h, "multi-GPU").
//
// The path has no exchange step (/root/reference/src/utils/packing/avx.rs:138-145 carries nothing between words), so
// the host-pointer calls are the single-device calls of api.cu run on contiguous shards, one host worker thread and
// one bn_ctx (streams + pinned staging) per device, all devices busy before anything waits.  The one collective of
// the path -- the sum of the four base counters (and of the hdist total) -- is either ncclAllReduce over
// ncclCommInitAll communicators (grouped; libnccl is resolved with dlopen) or our own mailbox all-reduce over NVLink
// peer memory (p2p_allreduce_kernel below): every device stores its <= 4 words plus an epoch flag into a slot of every
// peer's mailbox and sums the n slots of its own -- one 32-thread launch per device for a message that is pure latency.
#include "ctx.h"

#include <condition_variable>
#include <functional>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- NCCL through dlopen
// Only these seven entry points are used; their prototypes and the two enum values are NCCL 2.x ABI.
typedef struct ncclComm* ncclComm_t;
constexpr int kNcclSum = 0, kNcclUint64 = 5;
struct Nccl {
    int (*GetVersion)(int*) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

const Nccl& nccl() {
    static const Nccl lib = [] {
        Nccl n;
        // a libnccl.so.2 already in the process (e.g. torch's bundled one) is found first by its soname
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return n;
        auto sym = [&](const char* name) { return dlsym(h, name); };
        n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(sym("ncclGetVersion"));
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
        n.ok = n.GetVersion && n.CommInitAll && n.CommDestroy && n.GroupStart && n.GroupEnd && n.AllReduce;
        return n;
    }();
    return lib;
}

// ---------------------------------------------------------------------------------------------- mailbox all-reduce
constexpr int kMaxDev = 16;
constexpr int kSlotWords = 8;                                   // 4 values, the epoch flag, 3 pad words: one 64-byte slot
constexpr size_t kMailWords = 2 * kMaxDev * kSlotWords;         // [epoch parity][sender][slot]
constexpr long long kSpinLimit = 6000000000ll;                  // ~3 s of SM clocks: a peer that never arrives is reported, not waited for

struct PeerBox {
    unsigned long long* mail[kMaxDev];   // mail[p]: the mailbox that lives on shard p's device
    int n, self;
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// data[0..count) <- sum over the n shards of their data[0..count), on every device.  Lane p talks to peer p: it stores
// this device's words and then (release) the epoch into slot `self` of p's mailbox -- over NVLink when p is another
// device -- and waits (acquire) for slot p of its own mailbox to show the epoch.  Slots alternate with the epoch's
// parity: a sender can only be two epochs ahead of a receiver that has not yet read, never one (it needs the
// receiver's flag of the epoch in between, which the receiver's stream issues after its own read).
__global__ void __launch_bounds__(32)
p2p_allreduce_kernel(PeerBox box, unsigned long long* __restrict__ data, int count, unsigned long long epoch,
                     double* __restrict__ gc, unsigned int* __restrict__ fault) {
    const int lane = threadIdx.x;
    const size_t par = (size_t)(epoch & 1ull) * kMaxDev;
    unsigned long long v[4] = {0, 0, 0, 0};
    bool ok = true;
    if (lane < box.n) {
        unsigned long long* dst = box.mail[lane] + (par + box.self) * kSlotWords;
#pragma unroll
        for (int j = 0; j < 4; ++j) st_relaxed_sys(dst + j, j < count ? data[j] : 0ull);
        st_release_sys(dst + 4, epoch);
        const unsigned long long* src = box.mail[box.self] + (par + lane) * kSlotWords;
        const long long t0 = clock64();
        while (ld_acquire_sys(src + 4) != epoch) {
            if (clock64() - t0 > kSpinLimit) {
                ok = false;
                break;
            }
        }
        if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = ld_relaxed_sys(src + j);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = bn::warp_sum_u64(v[j]);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!all_ok) {
            atomicExch(fault, 1u);
        } else {
            for (int j = 0; j < count; ++j) data[j] = v[j];
            if (gc) {   // analysis.rs:14 on the reduced integer counts: (gc as f64 / len as f64) * 100.0
                const unsigned long long len = v[0] + v[1] + v[2] + v[3];
                *gc = len ? __dmul_rn(__ddiv_rn(__ull2double_rn(v[1] + v[2]), __ull2double_rn(len)), 100.0) : 0.0;
            }
        }
    }
}

__global__ void gc_from_counts_kernel(const unsigned long long* __restrict__ counts, double* __restrict__ gc) {
    const unsigned long long len = counts[0] + counts[1] + counts[2] + counts[3];
    *gc = len ? __dmul_rn(__ddiv_rn(__ull2double_rn(counts[1] + counts[2]), __ull2double_rn(len)), 100.0) : 0.0;
}

// ---------------------------------------------------------------------------------------------- host workers
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool pending = false, stop = false;

    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return pending || stop; });
            if (stop) return;
            lk.unlock();
            job();
            lk.lock();
            pending = false;
            cv.notify_all();
        }
    }
    void post(std::function<void()> f) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(f);
        pending = true;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !pending; });
    }
};

}  // namespace

struct bn_multi {
    int n = 0;
    std::vector<int> dev;
    std::vector<bn_ctx*> ctx;
    std::vector<Worker*> worker;
    bool distinct = true;                       // no device named twice (NCCL needs that)
    int reduce = BN_REDUCE_NCCL;
    std::vector<ncclComm_t> comm;               // NCCL mode
    int nccl_version = 0;
    std::vector<unsigned long long*> mail;      // P2P mode: mail[i] on device i
    std::vector<unsigned int*> fault;           // P2P mode: per-device "a peer never arrived" flag
    std::vector<unsigned long long*> coll;      // per-device scratch for host-pointer reductions (4 words)
    std::vector<cudaEvent_t> t0, t1;            // bn_multi_last_ms
    bool timed = false;
    unsigned long long epoch = 0;
    std::mutex mu;
};

namespace {

template <class F>
void run_all(bn_multi* m, F f) {
    for (int i = 1; i < m->n; ++i) m->worker[i]->post([=] { f(i); });
    f(0);   // the calling thread takes shard 0 itself
    for (int i = 1; i < m->n; ++i) m->worker[i]->wait();
}

void shard_units(int n, size_t n_units, size_t align, size_t* starts) {
    if (align == 0) align = 1;
    const size_t blocks = (n_units + align - 1) / align, per = blocks / n, extra = blocks % n;
    size_t b = 0;
    for (int i = 0; i <= n; ++i) {
        starts[i] = std::min(b * align, n_units);
        b += per + ((size_t)i < extra ? 1 : 0);
    }
    starts[n] = n_units;
}

int coll_fail(bn_error_t* err, int nccl_rc) {
    set_err(err, BN_ERR_COLLECTIVE);
    if (err) err->cuda_error = nccl_rc;
    return BN_ERR_COLLECTIVE;
}

int init_nccl(bn_multi* m) {
    if (!m->comm.empty()) return BN_OK;
    if (!m->distinct || m->n < 2) return BN_ERR_ARGUMENT;
    const Nccl& nc = nccl();
    if (!nc.ok) return BN_ERR_COLLECTIVE;
    m->comm.assign(m->n, nullptr);
    if (nc.CommInitAll(m->comm.data(), m->n, m->dev.data()) != 0) {
        m->comm.clear();
        return BN_ERR_COLLECTIVE;
    }
    nc.GetVersion(&m->nccl_version);
    return BN_OK;
}

int init_p2p(bn_multi* m) {
    if (!m->mail.empty()) return BN_OK;
    if (m->n > kMaxDev) return BN_ERR_ARGUMENT;
    for (int i = 0; i < m->n && m->distinct; ++i) {
        DeviceGuard g(m->dev[i]);
        for (int j = 0; j < m->n; ++j) {
            if (j == i) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[j]) != cudaSuccess || !can) {
                cudaGetLastError();
                return BN_ERR_COLLECTIVE;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                return BN_ERR_COLLECTIVE;
            }
            cudaGetLastError();
        }
    }
    m->mail.assign(m->n, nullptr);
    m->fault.assign(m->n, nullptr);
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        if (cudaMalloc(&m->mail[i], kMailWords * 8) != cudaSuccess || cudaMemset(m->mail[i], 0, kMailWords * 8) != cudaSuccess ||
            cudaMalloc(&m->fault[i], 4) != cudaSuccess || cudaMemset(m->fault[i], 0, 4) != cudaSuccess) {
            cudaGetLastError();
            return BN_ERR_CUDA;
        }
    }
    for (int i = 0; i < m->n; ++i) {   // the zeroed mailboxes must be in place before any peer stores into them
        DeviceGuard g(m->dev[i]);
        cudaDeviceSynchronize();
    }
    return BN_OK;
}

// In place, on every context's own stream; d_gc (array or entries may be null) receives the gc of the reduced counts.
int allreduce_enqueue(bn_multi* m, unsigned long long* const* d_buf, int count, double* const* d_gc) {
    if (m->n == 1) {
        if (d_gc && d_gc[0]) {
            DeviceGuard g(m->dev[0]);
            gc_from_counts_kernel<<<1, 1, 0, m->ctx[0]->stream>>>(d_buf[0], d_gc[0]);
            if (cudaGetLastError() != cudaSuccess) return BN_ERR_CUDA;
        }
        return BN_OK;
    }
    if (m->reduce == BN_REDUCE_NCCL) {
        const Nccl& nc = nccl();
        int rc = nc.GroupStart();
        for (int i = 0; i < m->n && rc == 0; ++i)
            rc = nc.AllReduce(d_buf[i], d_buf[i], (size_t)count, kNcclUint64, kNcclSum, m->comm[i], m->ctx[i]->stream);
        const int rc_end = nc.GroupEnd();
        if (rc != 0 || rc_end != 0) return BN_ERR_COLLECTIVE;
        for (int i = 0; i < m->n; ++i) {
            if (!d_gc || !d_gc[i]) continue;
            DeviceGuard g(m->dev[i]);
            gc_from_counts_kernel<<<1, 1, 0, m->ctx[i]->stream>>>(d_buf[i], d_gc[i]);
            if (cudaGetLastError() != cudaSuccess) return BN_ERR_CUDA;
        }
        return BN_OK;
    }
    const unsigned long long epoch = ++m->epoch;
    PeerBox box{};
    box.n = m->n;
    for (int i = 0; i < m->n; ++i) box.mail[i] = m->mail[i];
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        box.self = i;
        p2p_allreduce_kernel<<<1, 32, 0, m->ctx[i]->stream>>>(box, d_buf[i], count, epoch, d_gc ? d_gc[i] : nullptr, m->fault[i]);
        if (cudaGetLastError() != cudaSuccess) return BN_ERR_CUDA;
    }
    return BN_OK;
}

void time_begin(bn_multi* m) {
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        cudaEventRecord(m->t0[i], m->ctx[i]->stream);
    }
}
void time_end(bn_multi* m) {
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        cudaEventRecord(m->t1[i], m->ctx[i]->stream);
    }
    m->timed = true;
}

int sync_all(bn_multi* m, bn_error_t* err) {
    int rc = BN_OK;
    for (int i = 0; i < m->n; ++i) {
        const int r = bn_ctx_synchronize(m->ctx[i]);
        if (r != BN_OK && rc == BN_OK) rc = r;
    }
    if (rc != BN_OK) return set_err(err, rc);
    for (int i = 0; i < m->n && !m->fault.empty(); ++i) {   // a mailbox wait that gave up
        DeviceGuard g(m->dev[i]);
        unsigned int f = 0;
        if (cudaMemcpy(&f, m->fault[i], 4, cudaMemcpyDeviceToHost) != cudaSuccess) return cuda_fail(err, cudaGetLastError());
        if (f) {
            cudaMemset(m->fault[i], 0, 4);
            return coll_fail(err, 0);
        }
    }
    return BN_OK;
}

// Host-pointer reductions: every shard's four words are already on the host; they go through the same collective as
// the device-resident calls (up, all-reduce, down from shard 0) so that one reduction path serves -- and is tested by -- both.
int reduce_host4(bn_multi* m, const uint64_t (*part)[4], uint64_t out[4], bn_error_t* err) {
    if (m->n == 1) {
        for (int j = 0; j < 4; ++j) out[j] = part[0][j];
        return BN_OK;
    }
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        BN_CUDA(cudaMemcpyAsync(m->coll[i], part[i], 32, cudaMemcpyHostToDevice, m->ctx[i]->stream));
    }
    const int rc = allreduce_enqueue(m, m->coll.data(), 4, nullptr);
    if (rc != BN_OK) return set_err(err, rc);
    {
        DeviceGuard g(m->dev[0]);
        BN_CUDA(cudaMemcpyAsync(m->ctx[0]->h_words + 12, m->coll[0], 32, cudaMemcpyDeviceToHost, m->ctx[0]->stream));
    }
    const int rs = sync_all(m, err);
    if (rs != BN_OK) return rs;
    for (int j = 0; j < 4; ++j) out[j] = m->ctx[0]->h_words[12 + j];
    return BN_OK;
}

// The first shard in input order that did not succeed decides the call's result.
int first_failure(const std::vector<int>& rc) {
    for (size_t i = 0; i < rc.size(); ++i)
        if (rc[i] != BN_OK) return (int)i;
    return -1;
}

}  // namespace

extern "C" {

int bn_multi_create(const int* devs, int n, int reduce, bn_multi** out) {
    if (!out || (reduce != BN_REDUCE_NCCL && reduce != BN_REDUCE_P2P)) return BN_ERR_ARGUMENT;
    *out = nullptr;
    const int visible = bn_device_count();
    if (visible == 0) return BN_ERR_CUDA;
    if (n <= 0) {
        n = visible;
        devs = nullptr;
    }
    if (n > kMaxDev) return BN_ERR_ARGUMENT;
    bn_multi* m = new (std::nothrow) bn_multi();
    if (!m) return BN_ERR_NOMEM;
    m->n = n;
    for (int i = 0; i < n; ++i) {
        const int d = devs ? devs[i] : i;
        if (d < 0 || d >= visible) {
            delete m;
            return BN_ERR_ARGUMENT;
        }
        if (std::find(m->dev.begin(), m->dev.end(), d) != m->dev.end()) m->distinct = false;
        m->dev.push_back(d);
    }
    m->reduce = reduce;
    int rc = BN_OK;
    for (int i = 0; i < n && rc == BN_OK; ++i) {
        bn_ctx* c = nullptr;
        rc = bn_ctx_create(m->dev[i], &c);
        if (rc == BN_OK) m->ctx.push_back(c);
    }
    for (int i = 0; i < n && rc == BN_OK; ++i) {
        DeviceGuard g(m->dev[i]);
        unsigned long long* p = nullptr;
        cudaEvent_t a = nullptr, b = nullptr;
        if (cudaMalloc(&p, 64) != cudaSuccess || cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
            cudaGetLastError();
            rc = BN_ERR_CUDA;
        }
        m->coll.push_back(p);
        m->t0.push_back(a);
        m->t1.push_back(b);
    }
    if (rc == BN_OK && n > 1) {
        if (reduce == BN_REDUCE_NCCL && !m->distinct) rc = BN_ERR_ARGUMENT;   // NCCL refuses one device twice: ask for BN_REDUCE_P2P
        else rc = reduce == BN_REDUCE_NCCL ? init_nccl(m) : init_p2p(m);
    }
    if (rc == BN_OK) {
        m->worker.assign(n, nullptr);
        for (int i = 1; i < n; ++i) {
            m->worker[i] = new Worker();
            m->worker[i]->th = std::thread([w = m->worker[i]] { w->loop(); });
        }
    }
    if (rc != BN_OK) {
        bn_multi_destroy(m);
        return rc;
    }
    *out = m;
    return BN_OK;
}

void bn_multi_destroy(bn_multi* m) {
    if (!m) return;
    for (Worker* w : m->worker) {
        if (!w) continue;
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->stop = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (size_t i = 0; i < m->ctx.size(); ++i) bn_ctx_synchronize(m->ctx[i]);
    for (ncclComm_t c : m->comm)
        if (c) nccl().CommDestroy(c);
    for (size_t i = 0; i < m->dev.size(); ++i) {
        DeviceGuard g(m->dev[i]);
        if (i < m->mail.size() && m->mail[i]) cudaFree(m->mail[i]);
        if (i < m->fault.size() && m->fault[i]) cudaFree(m->fault[i]);
        if (i < m->coll.size() && m->coll[i]) cudaFree(m->coll[i]);
        if (i < m->t0.size() && m->t0[i]) cudaEventDestroy(m->t0[i]);
        if (i < m->t1.size() && m->t1[i]) cudaEventDestroy(m->t1[i]);
        cudaGetLastError();
    }
    for (bn_ctx* c : m->ctx) bn_ctx_destroy(c);
    delete m;
}

int bn_multi_size(const bn_multi* m) { return m ? m->n : 0; }
bn_ctx* bn_multi_ctx(bn_multi* m, int i) { return m && i >= 0 && i < m->n ? m->ctx[i] : nullptr; }
int bn_multi_reduce(const bn_multi* m) { return m ? m->reduce : BN_ERR_ARGUMENT; }
int bn_multi_nccl_version(const bn_multi* m) { return m ? m->nccl_version : 0; }

int bn_multi_set_chunk_bytes(bn_multi* m, size_t bytes) {
    if (!m) return BN_ERR_ARGUMENT;
    for (bn_ctx* c : m->ctx) bn_ctx_set_chunk_bytes(c, bytes);
    return BN_OK;
}

int bn_multi_synchronize(bn_multi* m) {
    if (!m) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(m->mu);
    return sync_all(m, nullptr);
}

int bn_multi_shard_units(const bn_multi* m, size_t n_units, size_t align, size_t* starts) {
    if (!m || !starts) return BN_ERR_ARGUMENT;
    shard_units(m->n, n_units, align, starts);
    return BN_OK;
}

int bn_multi_shard_reads(const bn_multi* m, const uint64_t* offsets, size_t n_reads, size_t* starts) {
    if (!m || !starts || (n_reads && !offsets)) return BN_ERR_ARGUMENT;
    starts[0] = 0;
    const uint64_t lo = n_reads ? offsets[0] : 0, hi = n_reads ? offsets[n_reads] : 0;
    for (int g = 1; g < m->n; ++g) {
        // the first read whose start reaches the ideal byte cut (128-bit product: hi - lo may be tens of gigabytes)
        const uint64_t target = lo + (uint64_t)(((unsigned __int128)(hi - lo) * (unsigned)g) / (unsigned)m->n);
        const size_t r = n_reads ? (size_t)(std::lower_bound(offsets, offsets + n_reads, target) - offsets) : 0;
        starts[g] = std::max(starts[g - 1], std::min(n_reads, r));
    }
    starts[m->n] = n_reads;
    return BN_OK;
}

// ------------------------------------------------------------------ host-pointer calls --------------------------

int bn_multi_encode(bn_multi* m, const uint8_t* seq, size_t n, uint64_t* out, size_t* n_words, bn_error_t* err) {
    if (n_words) *n_words = 0;
    if (!m) return set_err(err, BN_ERR_ARGUMENT);
    std::lock_guard<std::mutex> lk(m->mu);
    if (n == 0 || m->n == 1) return bn_encode(m->ctx[0], seq, n, out, n_words, err);   // empty: the reference's panic / aarch64 word
    if (!seq || !out) return set_err(err, BN_ERR_ARGUMENT);
    std::vector<size_t> st(m->n + 1), nw(m->n, 0);
    shard_units(m->n, n, 64, st.data());
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    run_all(m, [&](int i) {
        const size_t len = st[i + 1] - st[i];
        if (len) rc[i] = bn_encode(m->ctx[i], seq + st[i], len, out + st[i] / 32, &nw[i], &e[i]);
    });
    const int f = first_failure(rc);
    if (f < 0) {
        if (n_words) *n_words = (n + 31) / 32;
        return set_err(err, BN_OK);
    }
    // the words before the failing chunk: every earlier shard is complete, the failing one holds nw[f] of its own
    if (rc[f] == BN_INVALID_BASE && n_words) *n_words = st[f] / 32 + nw[f];
    if (err) {
        *err = e[f];
        if (rc[f] == BN_INVALID_BASE) err->offset += st[f];
    }
    return rc[f];
}

int bn_multi_decode(bn_multi* m, const uint64_t* words, size_t n_words, size_t n_bases, uint8_t* out, bn_error_t* err) {
    if (!m) return set_err(err, BN_ERR_ARGUMENT);
    if (n_words < (n_bases + 31) / 32) return set_err(err, BN_INVALID_LENGTH, n_bases);
    if (n_bases == 0) return set_err(err, BN_OK);
    if (!words || !out) return set_err(err, BN_ERR_ARGUMENT);
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<size_t> st(m->n + 1);
    shard_units(m->n, n_bases, 64, st.data());
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    run_all(m, [&](int i) {
        const size_t len = st[i + 1] - st[i];
        if (len) rc[i] = bn_decode(m->ctx[i], words + st[i] / 32, (len + 31) / 32, len, out + st[i], &e[i]);
    });
    const int f = first_failure(rc);
    if (f < 0) return set_err(err, BN_OK);
    if (err) *err = e[f];
    return rc[f];
}

int bn_multi_as_2bit_batch(bn_multi* m, const uint8_t* recs, size_t n, uint32_t k, size_t stride, uint64_t* out, bn_error_t* err) {
    if (!m) return set_err(err, BN_ERR_ARGUMENT);
    if (k > 32) return set_err(err, BN_SEQUENCE_TOO_LONG, k);
    if (stride < k || (n && (!out || (k && !recs)))) return set_err(err, BN_ERR_ARGUMENT);
    if (n == 0) return set_err(err, BN_OK);
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<size_t> st(m->n + 1);
    shard_units(m->n, n, 64, st.data());   // 64 records: every shard's first record starts 16-byte aligned whatever the stride
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    run_all(m, [&](int i) {
        const size_t cnt = st[i + 1] - st[i];
        if (cnt) rc[i] = bn_as_2bit_batch(m->ctx[i], recs + st[i] * stride, cnt, k, stride, out + st[i], &e[i]);
    });
    const int f = first_failure(rc);
    if (f < 0) return set_err(err, BN_OK);
    if (err) {
        *err = e[f];
        if (rc[f] == BN_INVALID_BASE) {
            err->offset += st[f] * stride;
            err->record += st[f];
        }
    }
    return rc[f];
}

int bn_multi_from_2bit_batch(bn_multi* m, const uint64_t* packed, size_t n, uint32_t k, uint8_t* out, size_t stride, bn_error_t* err) {
    if (!m) return set_err(err, BN_ERR_ARGUMENT);
    if (k > 32) return set_err(err, BN_INVALID_LENGTH, k);
    if (stride < k || (n && k && (!packed || !out))) return set_err(err, BN_ERR_ARGUMENT);
    if (n == 0 || k == 0) return set_err(err, BN_OK);
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<size_t> st(m->n + 1);
    shard_units(m->n, n, 64, st.data());
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    run_all(m, [&](int i) {
        const size_t cnt = st[i + 1] - st[i];
        if (cnt) rc[i] = bn_from_2bit_batch(m->ctx[i], packed + st[i], cnt, k, out + st[i] * stride, stride, &e[i]);
    });
    const int f = first_failure(rc);
    if (f < 0) return set_err(err, BN_OK);
    if (err) *err = e[f];
    return rc[f];
}

int bn_multi_hdist(bn_multi* m, const uint64_t* a, size_t n_words_a, const uint64_t* b, size_t n_words_b, size_t n_bases,
                   uint64_t* total, bn_error_t* err) {
    if (!m || !total) return set_err(err, BN_ERR_ARGUMENT);
    const size_t need = (n_bases + 31) / 32;
    if (n_words_a < need || n_words_b < need) return set_err(err, BN_INVALID_LENGTH, n_bases);   // multi.rs:124-127
    *total = 0;
    if (n_bases == 0) return set_err(err, BN_OK);
    if (!a || !b) return set_err(err, BN_ERR_ARGUMENT);
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<size_t> st(m->n + 1);
    shard_units(m->n, n_bases, 64, st.data());
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    std::vector<uint64_t> part(m->n, 0);
    run_all(m, [&](int i) {
        const size_t len = st[i + 1] - st[i], nw = (len + 31) / 32;
        if (len) rc[i] = bn_hdist(m->ctx[i], a + st[i] / 32, nw, b + st[i] / 32, nw, len, &part[i], &e[i]);
    });
    const int f = first_failure(rc);
    if (f >= 0) {
        if (err) *err = e[f];
        return rc[f];
    }
    for (int i = 0; i < m->n; ++i) *total += part[i];
    return set_err(err, BN_OK);
}

int bn_multi_hdist_pairs(bn_multi* m, const uint64_t* u, const uint64_t* v, size_t n_pairs, uint32_t len, uint32_t* out, bn_error_t* err) {
    if (!m) return set_err(err, BN_ERR_ARGUMENT);
    if (len > 32) return set_err(err, BN_INVALID_LENGTH, len);   // scalar.rs:13-15
    if (n_pairs == 0) return set_err(err, BN_OK);
    if (!u || !v || !out) return set_err(err, BN_ERR_ARGUMENT);
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<size_t> st(m->n + 1);
    shard_units(m->n, n_pairs, 4, st.data());
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    run_all(m, [&](int i) {
        const size_t cnt = st[i + 1] - st[i];
        if (cnt) rc[i] = bn_hdist_pairs(m->ctx[i], u + st[i], v + st[i], cnt, len, out + st[i], &e[i]);
    });
    const int f = first_failure(rc);
    if (f < 0) return set_err(err, BN_OK);
    if (err) *err = e[f];
    return rc[f];
}

int bn_multi_base_counts(bn_multi* m, const uint64_t* words, size_t n_words, size_t n_bases, uint64_t counts[4], double* gc, bn_error_t* err) {
    if (!m || !counts) return set_err(err, BN_ERR_ARGUMENT);
    const size_t need = (n_bases + 31) / 32;
    if (n_words < need) return set_err(err, BN_INVALID_LENGTH, n_bases);
    if (n_bases && !words) return set_err(err, BN_ERR_ARGUMENT);
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    if (gc) *gc = 0.0;
    if (n_bases == 0) return set_err(err, BN_OK);
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<size_t> st(m->n + 1);
    shard_units(m->n, n_bases, 64, st.data());
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    std::vector<uint64_t> part(4 * (size_t)m->n, 0);
    run_all(m, [&](int i) {
        const size_t len = st[i + 1] - st[i];
        if (len) rc[i] = bn_base_counts(m->ctx[i], words + st[i] / 32, (len + 31) / 32, len, &part[4 * i], nullptr, &e[i]);
    });
    const int f = first_failure(rc);
    if (f >= 0) {
        if (err) *err = e[f];
        return rc[f];
    }
    const int rr = reduce_host4(m, reinterpret_cast<const uint64_t(*)[4]>(part.data()), counts, err);
    if (rr != BN_OK) return rr;
    if (gc) {   // analysis.rs:14, exactly this operation order on exact integer counts
        volatile double q = (double)(counts[1] + counts[2]) / (double)n_bases;
        *gc = q * 100.0;
    }
    return set_err(err, BN_OK);
}

int bn_multi_base_counts_batch(bn_multi* m, const uint64_t* words, size_t n_words, const uint64_t* word_offsets, const uint64_t* lens,
                               size_t n_reads, uint64_t* counts4, double* gc, uint64_t totals[4], bn_error_t* err) {
    if (!m || (n_reads && (!word_offsets || !lens))) return set_err(err, BN_ERR_ARGUMENT);
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->n == 1 || n_reads == 0) return bn_base_counts_batch(m->ctx[0], words, n_words, word_offsets, lens, n_reads, counts4, gc, totals, err);
    std::vector<size_t> st(m->n + 1);
    shard_units(m->n, n_reads, 1, st.data());
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    std::vector<uint64_t> part(4 * (size_t)m->n, 0);
    run_all(m, [&](int i) {
        const size_t r0 = st[i], cnt = st[i + 1] - r0;
        if (cnt)   // the shard indexes the caller's `words` with the caller's absolute word offsets
            rc[i] = bn_base_counts_batch(m->ctx[i], words, n_words, word_offsets + r0, lens + r0, cnt, counts4 ? counts4 + 4 * r0 : nullptr,
                                         gc ? gc + r0 : nullptr, &part[4 * i], &e[i]);
    });
    const int f = first_failure(rc);
    if (f >= 0) {
        if (err) {
            *err = e[f];
            if (rc[f] == BN_INVALID_LENGTH) err->record += st[f];
        }
        return rc[f];
    }
    uint64_t sum[4];
    const int rr = reduce_host4(m, reinterpret_cast<const uint64_t(*)[4]>(part.data()), sum, err);
    if (rr != BN_OK) return rr;
    if (totals)
        for (int j = 0; j < 4; ++j) totals[j] = sum[j];
    return set_err(err, BN_OK);
}

int bn_multi_encode_batch(bn_multi* m, const uint8_t* bytes, const uint64_t* offsets, size_t n_reads, uint64_t* out_words,
                          uint64_t* out_word_offsets, uint32_t* read_status, bn_error_t* err) {
    if (!m || !out_word_offsets || (n_reads && !offsets)) return set_err(err, BN_ERR_ARGUMENT);
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->n == 1 || n_reads == 0) return bn_encode_batch(m->ctx[0], bytes, offsets, n_reads, out_words, out_word_offsets, read_status, err);
    std::vector<size_t> st(m->n + 1);
    bn_multi_shard_reads(m, offsets, n_reads, st.data());
    // pass 1 (all shards at once): validate the shard's offsets and count its output words, so that every shard knows
    // where its words start before anything is encoded
    std::vector<uint64_t> nw(m->n, 0);
    std::vector<char> bad(m->n, 0);
    run_all(m, [&](int i) {
        uint64_t w = 0;
        for (size_t r = st[i]; r < st[i + 1]; ++r) {
            if (offsets[r + 1] < offsets[r]) {
                bad[i] = 1;
                return;
            }
            w += (offsets[r + 1] - offsets[r] + 31) / 32;
        }
        nw[i] = w;
    });
    if (std::any_of(bad.begin(), bad.end(), [](char c) { return c != 0; })) return set_err(err, BN_ERR_ARGUMENT);
    std::vector<uint64_t> w0(m->n + 1, 0);
    for (int i = 0; i < m->n; ++i) w0[i + 1] = w0[i] + nw[i];
    if (offsets[n_reads] > offsets[0] && (!bytes || !out_words)) return set_err(err, BN_ERR_ARGUMENT);
    std::vector<int> rc(m->n, BN_OK);
    std::vector<bn_error_t> e(m->n);
    run_all(m, [&](int i) {
        const size_t r0 = st[i], cnt = st[i + 1] - r0;
        if (!cnt) return;
        // absolute byte offsets into the caller's `bytes`; word offsets come back relative to the shard
        rc[i] = bn_encode_batch(m->ctx[i], bytes, offsets + r0, cnt, out_words + w0[i], out_word_offsets + r0,
                                read_status ? read_status + r0 : nullptr, &e[i]);
        if (rc[i] == BN_OK || rc[i] == BN_INVALID_BASE)
            for (size_t r = r0 + 1; r <= r0 + cnt; ++r) out_word_offsets[r] += w0[i];
    });
    // entry st[i] is written by two shards (the last entry of one, the zero first entry of the next): settle it here
    for (int i = 0; i < m->n; ++i) out_word_offsets[st[i]] = w0[i];
    out_word_offsets[n_reads] = w0[m->n];
    const int f = first_failure(rc);
    if (f < 0) return set_err(err, BN_OK);
    if (err) {
        *err = e[f];
        if (rc[f] == BN_INVALID_BASE) err->record += st[f];
    }
    return rc[f];
}

// ------------------------------------------------------------------ device-resident sharded reductions ----------

int bn_multi_allreduce_u64_dev(bn_multi* m, uint64_t* const* d_buf, int count) {
    if (!m || !d_buf || count < 1 || count > 4) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(m->mu);
    time_begin(m);
    const int rc = allreduce_enqueue(m, reinterpret_cast<unsigned long long* const*>(d_buf), count, nullptr);
    time_end(m);
    return rc;
}

int bn_multi_base_counts_dev(bn_multi* m, const uint64_t* const* d_words, const size_t* n_bases, uint64_t* const* d_counts, double* const* d_gc) {
    if (!m || !d_words || !n_bases || !d_counts) return BN_ERR_ARGUMENT;
    for (int i = 0; i < m->n; ++i)
        if (!d_counts[i] || (n_bases[i] && !d_words[i])) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(m->mu);
    time_begin(m);
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        if (bn::launch_base_counts(m->ctx[i]->di, d_words[i], n_bases[i], reinterpret_cast<unsigned long long*>(d_counts[i]), nullptr,
                                   m->ctx[i]->stream) != cudaSuccess) {
            cudaGetLastError();
            return BN_ERR_CUDA;
        }
    }
    const int rc = allreduce_enqueue(m, reinterpret_cast<unsigned long long* const*>(d_counts), 4, d_gc);
    time_end(m);
    return rc;
}

int bn_multi_base_counts_fixed_dev(bn_multi* m, const uint64_t* const* d_words, const size_t* n_reads, size_t read_len,
                                   uint64_t* const* d_counts4, double* const* d_gc_reads, uint64_t* const* d_totals, double* const* d_gc) {
    if (!m || !d_words || !n_reads || !d_totals) return BN_ERR_ARGUMENT;
    for (int i = 0; i < m->n; ++i)
        if (!d_totals[i] || (n_reads[i] && read_len && !d_words[i])) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(m->mu);
    time_begin(m);
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        if (bn::launch_base_counts_batch(m->ctx[i]->di, d_words[i], nullptr, nullptr, n_reads[i], read_len, n_reads[i] * ((read_len + 31) / 32),
                                         d_counts4 ? reinterpret_cast<unsigned long long*>(d_counts4[i]) : nullptr,
                                         d_gc_reads ? d_gc_reads[i] : nullptr, reinterpret_cast<unsigned long long*>(d_totals[i]),
                                         m->ctx[i]->stream) != cudaSuccess) {
            cudaGetLastError();
            return BN_ERR_CUDA;
        }
    }
    const int rc = allreduce_enqueue(m, reinterpret_cast<unsigned long long* const*>(d_totals), 4, d_gc);
    time_end(m);
    return rc;
}

int bn_multi_hdist_dev(bn_multi* m, const uint64_t* const* d_a, const uint64_t* const* d_b, const size_t* n_bases, uint64_t* const* d_total) {
    if (!m || !d_a || !d_b || !n_bases || !d_total) return BN_ERR_ARGUMENT;
    for (int i = 0; i < m->n; ++i)
        if (!d_total[i] || (n_bases[i] && (!d_a[i] || !d_b[i]))) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(m->mu);
    time_begin(m);
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        if (bn::launch_hdist(m->ctx[i]->di, d_a[i], d_b[i], n_bases[i], reinterpret_cast<unsigned long long*>(d_total[i]),
                             m->ctx[i]->stream) != cudaSuccess) {
            cudaGetLastError();
            return BN_ERR_CUDA;
        }
    }
    const int rc = allreduce_enqueue(m, reinterpret_cast<unsigned long long* const*>(d_total), 1, nullptr);
    time_end(m);
    return rc;
}

int bn_multi_last_ms(bn_multi* m, float* ms) {
    if (!m || !ms) return BN_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(m->mu);
    if (!m->timed) return BN_ERR_ARGUMENT;
    for (int i = 0; i < m->n; ++i) {
        DeviceGuard g(m->dev[i]);
        if (cudaEventSynchronize(m->t1[i]) != cudaSuccess || cudaEventElapsedTime(&ms[i], m->t0[i], m->t1[i]) != cudaSuccess) {
            cudaGetLastError();
            return BN_ERR_CUDA;
        }
    }
    return BN_OK;
}

}  // extern "C"
