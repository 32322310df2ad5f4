// mrag.cu -- the C ABI of include/mrag.h: index lifecycle, write side, and the search driver
// that strings the kernels together on one CUDA stream:
//
//   query_prep_kernel -> [pool_bitmap_kernel] -> [filter_mask_kernel]      (prepare)
//   -> scan_mma_kernel x ceil(nq / 64)   (bf16 corpus, dim <= 768: TMA + tcgen05 + TMEM)
//      or scan_gemv_kernel x ceil(nq / 4) (CUDA cores: fp32 corpus, wide rows, single queries,
//                                          selective filters -- it never loads a masked row)
//   -> merge_kernel -> nan_tail_kernel                                     (ORDER BY .. LIMIT k)
//
// Stands in for the SQL statement of app/services/vector_store.py:274-287 and
// app/services/corpus_search.py:1525-1536 (see include/mrag.h for the clause-by-clause map).
// There is no CPU path in this file: without a device every compute entry point fails.
#include "../../include/mrag.h"
#include "common.cuh"
#include "prep.cuh"
#include "scan_gemv.cuh"
#include "scan_mma.cuh"
#include "scan_mma128.cuh"
#include "scan_mma256w.cuh"
#include "select.cuh"
#include "rescore.cuh"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <string>
#include <vector>

using namespace mrag;

// ------------------------------------------------------------------------------------------
// errors / counters
// ------------------------------------------------------------------------------------------
static thread_local std::string t_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
    return code;
}

#define CU(expr)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            int c_ = (e_ == cudaErrorMemoryAllocation) ? MRAG_ERR_OOM : MRAG_ERR_CUDA;        \
            return fail(c_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                            \
        }                                                                                     \
    } while (0)

#define LAUNCHED()                                                                         \
    do {                                                                                   \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess)                                                             \
            return fail(MRAG_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                 \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                       \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int host_next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ------------------------------------------------------------------------------------------
// per-call workspace (one per concurrent search; pooled on the index)
// ------------------------------------------------------------------------------------------
struct EventSet {
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};   // start, prepared, scanned, end
    bool recorded = false;
    int create() {
        for (auto& x : e) CU(cudaEventCreate(&x));
        return 0;
    }
    void destroy() {
        for (auto& x : e) if (x) { cudaEventDestroy(x); x = nullptr; }
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = n + n / 4 + 64;
        CU(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Workspace {
    cudaStream_t own_stream = nullptr;
    cudaStream_t last_stream = nullptr;   // stream of the last (possibly still running) use
    bool in_use = false;
    EventSet ev;
    DevBuf<float> qraw, qpad, qinv, scores;
    DevBuf<__nv_bfloat16> qbf;
    DevBuf<uint32_t> mask, pool, pool_bits, gthr, qhl, gmax;
    DevBuf<uint64_t> part, part2, part3, ub, gcand;
    DevBuf<int64_t> rows, crows;
    DevBuf<int32_t> counts, ccounts;
    DevBuf<float> cscores;
    DevBuf<uint64_t> ckeys;
    DevBuf<int> fb;                 // [0] count, [1..] query list of the exact fallback
    DevBuf<DevHyb> hyb;             // hybrid search: per-query parameters
    DevBuf<HybChunk> hchunks;       // hybrid search: the same, transposed per chunk of 32 queries (hybrid_mask_kernel phase 1)
    DevBuf<uint32_t> hmask;         // hybrid search: per-query row bitmaps
    DevBuf<uint32_t> pair_rows;     // hybrid search, pair path: rows of the surviving (query, row) pairs, one segment per query
    DevBuf<int64_t> segoff;         // [nq + 1] segment offsets
    DevBuf<uint16_t> codes;         // d-tag arm: the codes asked for
    DevBuf<int> flags;              // [0] need_tail
    DevBuf<unsigned long long> npass, stats;
    void release() {
        qraw.release(); qpad.release(); qinv.release(); scores.release(); qbf.release(); qhl.release();
        mask.release(); pool.release(); pool_bits.release(); gthr.release(); gmax.release(); part.release(); part2.release(); part3.release(); ub.release(); gcand.release();
        rows.release(); counts.release(); crows.release(); ccounts.release(); cscores.release(); ckeys.release(); fb.release(); hyb.release(); hchunks.release(); hmask.release(); pair_rows.release(); segoff.release(); codes.release(); flags.release(); npass.release(); stats.release();
        ev.destroy();
        if (own_stream) cudaStreamDestroy(own_stream);
        own_stream = nullptr;
    }
};

// ------------------------------------------------------------------------------------------
// the index
// ------------------------------------------------------------------------------------------
struct mrag_index {
    int dim = 0, ld = 0, dtype = 0, device = 0, num_sms = 0;
    int64_t capacity = 0, size = 0, row_base = 0;
    int64_t n_docs = 0;                 // max doc_idx seen + 1
    void* rows = nullptr;               // [capacity][ld] storage dtype
    __nv_bfloat16* shadow = nullptr;    // fp32 indexes, ld <= 768: bf16 copy scanned by the candidate-generating kernel
    float* inv_norm = nullptr;          // [capacity] 1/|x| of the stored row (+inf: zero norm)
    MetaCols cols{};
    uint64_t* doc_tags = nullptr;       // [tag_docs_cap][MRAG_TAG_WORDS]
    int64_t n_tag_docs = 0, tag_docs_cap = 0;
    mrag_chunkfeat* feat = nullptr;     // [capacity] text features of the hybrid rerank (allocated on first use, zero = none)
    uint32_t* over_rows = nullptr;      // chunk d-tag keys beyond the 4 inline slots: (row, code) pairs sorted by row
    uint16_t* over_codes = nullptr;
    int64_t n_over = 0;
    int64_t* row_ids = nullptr;         // [capacity] caller-assigned id of every row (mrag_set_row_ids), nullptr = row + row_base
    uint64_t* doc_jtags = nullptr;      // [jtag_docs_cap][MRAG_JTAG_WORDS]
    int64_t n_jtag_docs = 0, jtag_docs_cap = 0;
    CUtensorMap tmap;                   // bf16 rows (or the shadow) as a 2-D tensor, 64x64 boxes, SWIZZLE_128B
    CUtensorMap tmap32;                 // the same tensor with 64x32 boxes (half tiles of the CTA-pair scan)
    bool has_tmap = false;
    cudaStream_t wstream = nullptr;     // write-side stream
    std::atomic<cudaEvent_t> prepared_event{nullptr};   // mrag_set_prepared_event
    std::shared_mutex lock;             // searches share, writers exclude
    std::mutex ws_lock;
    std::vector<Workspace*> pool;
    // request coalescing (MRAG_OPT_COALESCE): unfiltered host-buffer searches that arrive while a scan is in flight
    // queue up and run as ONE batch (one pass over the corpus serves up to 1024 queries) when it ends
    struct Pending {
        const float* q; int nq; int k; uint32_t options;
        float* scores; int64_t* rows; int32_t* counts;
        int rc = 0; std::string err; bool done = false;
    };
    std::mutex co_lock;
    std::condition_variable co_cv;
    std::vector<Pending*> co_queue;
    bool co_busy = false;
    std::atomic<int64_t> co_batches{0}, co_requests{0};
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static size_t elem_size(int dtype) { return dtype == MRAG_BF16 ? 2 : 4; }

// a candidate pool resident on the device: the four cascade levels as document bitmaps + the level in use
struct mrag_pool {
    mrag_index* idx = nullptr;
    uint32_t* bits = nullptr;           // [MRAG_POOL_LEVELS][words]
    int64_t words = 0, n_docs = 0;
    int level = -1;
    int64_t counts[MRAG_POOL_LEVELS] = {0, 0, 0, 0};
};

// thread-local record of the last search (for mrag_last_kernel_ms / mrag_last_scan_kind)
static thread_local EventSet t_last_ev;          // borrowed handles (owned by a workspace / ring)
static thread_local bool t_last_valid = false;
static thread_local const char* t_last_kind = "none";
static thread_local unsigned long long* t_stats_ptr = nullptr;
static thread_local int* t_fb_ptr = nullptr;
static thread_local bool t_hybrid_pairs = false;      // fallback counter of the last approx+rescore search
static thread_local std::vector<EventSet> t_ring;
static thread_local int t_ring_used = 0;
static thread_local int t_ring_device = -1;

// ------------------------------------------------------------------------------------------
// TMA descriptor of the corpus (driver entry point fetched through the runtime: no -lcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_corpus_tmap(mrag_index* x, CUtensorMap* out, void* base, int64_t alloc_rows, int box_rows) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(MRAG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[2] = {cuuint64_t(x->ld), cuuint64_t(alloc_rows)};
    cuuint64_t gstride[1] = {cuuint64_t(x->ld) * 2};
    cuuint32_t box[2] = {cuuint32_t(kMmaKBlock), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<PFN_tmapEncodeTiled>(fn)(
        out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MRAG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
    x->has_tmap = true;
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------
static int set_search_chain_carveout(int device);

extern "C" int mrag_create(mrag_index** out, int dim, int dtype, int device, int64_t capacity) {
    if (!out) return fail(MRAG_ERR_ARG, "mrag_create: out is null");
    *out = nullptr;
    if (dim < 1 || dim > 16384) return fail(MRAG_ERR_ARG, "mrag_create: dim %d out of range [1,16384]", dim);
    if (dtype != MRAG_F32 && dtype != MRAG_BF16) return fail(MRAG_ERR_ARG, "mrag_create: unknown dtype %d", dtype);
    if (capacity < 1 || capacity > (int64_t(1) << 31) - 64)
        return fail(MRAG_ERR_ARG, "mrag_create: capacity %lld out of range [1, 2^31-64]", (long long)capacity);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(MRAG_ERR_CUDA, "mrag_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(MRAG_ERR_ARG, "mrag_create: device %d not in [0,%d)", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_create: cudaSetDevice(%d) failed", device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(MRAG_ERR_CUDA, "mrag_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    if (int rc = set_search_chain_carveout(device)) return rc;
    mrag_index* x = new (std::nothrow) mrag_index();
    if (!x) return fail(MRAG_ERR_OOM, "mrag_create: host allocation failed");
    x->dim = dim;
    x->ld = int(ceil_div(dim, 64) * 64);
    x->dtype = dtype;
    x->device = device;
    x->num_sms = prop.multiProcessorCount;
    x->capacity = capacity;
    const int64_t cap32 = ceil_div(capacity, 32) * 32;
    const size_t row_bytes = size_t(x->ld) * elem_size(dtype);
    cudaError_t e = cudaSuccess;
    // 2 MB of slack after the last row: the TMA / vector paths may touch a whole tile past `size`
    if (e == cudaSuccess) e = cudaMalloc(&x->rows, size_t(cap32) * row_bytes + (2u << 20));
    if (e == cudaSuccess) e = cudaMalloc(&x->inv_norm, size_t(cap32 + 128) * 4);  // bulk-copied per 64- or 128-row tile
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.doc_idx, size_t(cap32) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.payer, size_t(cap32) * 2);
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.state, size_t(cap32));
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.program, size_t(cap32));
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.authority, size_t(cap32));
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.source_type, size_t(cap32));
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.valid, size_t(cap32 / 32 + 1) * 4);
    if (e == cudaSuccess) e = cudaMemset(x->cols.valid, 0, size_t(cap32 / 32 + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&x->cols.live, size_t(cap32 / 32 + 1) * 4);
    if (e == cudaSuccess) e = cudaMemset(x->cols.live, 0, size_t(cap32 / 32 + 1) * 4);
    if (e == cudaSuccess) e = cudaMemset(x->inv_norm, 0, size_t(cap32 + 128) * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&x->wstream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        int code = (e == cudaErrorMemoryAllocation) ? MRAG_ERR_OOM : MRAG_ERR_CUDA;
        fail(code, "mrag_create: %s (capacity %lld x ld %d x %zu B)", cudaGetErrorString(e),
             (long long)capacity, x->ld, elem_size(dtype));
        cudaGetLastError();
        mrag_destroy(x);
        t_err = std::string("mrag_create: ") + cudaGetErrorString(e);
        return code;
    }
    const char* shadow_env = getenv("MRAG_F32_SHADOW");
    if (dtype == MRAG_F32 && !(shadow_env && shadow_env[0] == '0')) {
        // +50% memory buys candidate scans at 2 B per element (tensor-core, 128 queries per pass, for dim <= 768;
        // CUDA-core over the shadow otherwise) with exact rescoring
        cudaError_t es = cudaMalloc(&x->shadow, size_t(cap32) * x->ld * 2 + (2u << 20));
        if (es != cudaSuccess) { x->shadow = nullptr; cudaGetLastError(); }   // not fatal: exact scan only
    }
    // bf16 rows of up to 1536 elements feed the tensor-core scans (one CTA up to 768, k-split CTA pairs beyond); the bf16
    // shadow of an fp32 shard only the candidate scans (<= 768)
    if ((dtype == MRAG_BF16 && x->ld <= kMmaKsMaxLd) || (x->shadow && x->ld <= kMmaMaxLd)) {
        // rows past `capacity` inside the allocation are never selected (mask bits are zero)
        void* tbase = dtype == MRAG_BF16 ? x->rows : static_cast<void*>(x->shadow);
        if (make_corpus_tmap(x, &x->tmap, tbase, cap32 + 64, kMmaTileRows) != MRAG_OK ||
            make_corpus_tmap(x, &x->tmap32, tbase, cap32 + 64, kMma256HalfRows) != MRAG_OK) {
            std::string keep = t_err;
            mrag_destroy(x);
            t_err = keep;
            return MRAG_ERR_CUDA;
        }
    }
    *out = x;
    return MRAG_OK;
}

extern "C" int mrag_destroy(mrag_index* x) {
    if (!x) return MRAG_OK;
    DeviceGuard g(x->device);
    cudaDeviceSynchronize();
    for (Workspace* w : x->pool) { w->release(); delete w; }
    x->pool.clear();
    if (x->rows) cudaFree(x->rows);
    if (x->shadow) cudaFree(x->shadow);
    if (x->inv_norm) cudaFree(x->inv_norm);
    if (x->cols.doc_idx) cudaFree(x->cols.doc_idx);
    if (x->cols.payer) cudaFree(x->cols.payer);
    if (x->cols.state) cudaFree(x->cols.state);
    if (x->cols.program) cudaFree(x->cols.program);
    if (x->cols.authority) cudaFree(x->cols.authority);
    if (x->cols.source_type) cudaFree(x->cols.source_type);
    if (x->cols.valid) cudaFree(x->cols.valid);
    if (x->cols.live) cudaFree(x->cols.live);
    if (x->doc_tags) cudaFree(x->doc_tags);
    if (x->feat) cudaFree(x->feat);
    if (x->row_ids) cudaFree(x->row_ids);
    if (x->over_rows) cudaFree(x->over_rows);
    if (x->over_codes) cudaFree(x->over_codes);
    if (x->doc_jtags) cudaFree(x->doc_jtags);
    if (x->wstream) cudaStreamDestroy(x->wstream);
    t_last_valid = false;
    delete x;
    return MRAG_OK;
}

extern "C" int64_t mrag_size(const mrag_index* x) { return x ? x->size : -1; }
extern "C" int64_t mrag_capacity(const mrag_index* x) { return x ? x->capacity : -1; }
extern "C" int mrag_dim(const mrag_index* x) { return x ? x->dim : -1; }
extern "C" int mrag_index_dtype(const mrag_index* x) { return x ? x->dtype : -1; }
extern "C" int mrag_device(const mrag_index* x) { return x ? x->device : -1; }

extern "C" int mrag_set_prepared_event(mrag_index* x, void* cuda_event) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_set_prepared_event: null index");
    x->prepared_event.store(static_cast<cudaEvent_t>(cuda_event), std::memory_order_release);
    return MRAG_OK;
}

extern "C" int mrag_set_row_base(mrag_index* x, int64_t row_base) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_set_row_base: null index");
    std::unique_lock<std::shared_mutex> wl(x->lock);
    x->row_base = row_base;
    return MRAG_OK;
}

// returned row = ids[row] instead of row + row_base (several shards of ONE table in one process: the id is the row's
// position in the host table, so ties break exactly as in an unsharded table)
__global__ void remap_rows_kernel(int64_t* __restrict__ rows, size_t count, const int64_t* __restrict__ ids, int64_t row_base) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) {
        const int64_t r = rows[i];
        if (r >= 0) rows[i] = ids[r - row_base];
    }
}

extern "C" int mrag_set_row_ids(mrag_index* x, int64_t first_row, const int64_t* ids, int64_t n) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_set_row_ids: null index");
    if (first_row < 0 || n < 0 || first_row + n > x->capacity)
        return fail(MRAG_ERR_ARG, "mrag_set_row_ids: rows [%lld, %lld) outside the capacity %lld", (long long)first_row,
                    (long long)(first_row + n), (long long)x->capacity);
    if (n == 0) return MRAG_OK;
    if (!ids) return fail(MRAG_ERR_ARG, "mrag_set_row_ids: ids is null");
    std::unique_lock<std::shared_mutex> wl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_set_row_ids: cudaSetDevice failed");
    if (!x->row_ids) {
        // rows without an explicit id keep the default (row + row_base): fill the identity first
        std::vector<int64_t> ident(size_t(x->capacity));
        for (int64_t i = 0; i < x->capacity; ++i) ident[size_t(i)] = i + x->row_base;
        CU(cudaMalloc(&x->row_ids, size_t(x->capacity) * 8));
        CU(cudaMemcpyAsync(x->row_ids, ident.data(), size_t(x->capacity) * 8, cudaMemcpyHostToDevice, x->wstream));
        CU(cudaStreamSynchronize(x->wstream));
    }
    CU(cudaMemcpyAsync(x->row_ids + first_row, ids, size_t(n) * 8, cudaMemcpyHostToDevice, x->wstream));
    CU(cudaStreamSynchronize(x->wstream));
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// write side
// ------------------------------------------------------------------------------------------
static int append_device_locked(mrag_index* x, const float* d_rows, int64_t n, const mrag_rowmeta* meta,
                                int64_t first, cudaStream_t s, mrag_rowmeta* d_meta_scratch) {
    const int wpb = 256 / 32;
    if (x->dtype == MRAG_BF16)
        store_rows_kernel<1><<<unsigned(ceil_div(n, wpb)), 256, 0, s>>>(d_rows, n, x->dim, x->rows, x->ld, first, x->inv_norm, nullptr);
    else
        store_rows_kernel<0><<<unsigned(ceil_div(n, wpb)), 256, 0, s>>>(d_rows, n, x->dim, x->rows, x->ld, first, x->inv_norm, x->shadow);
    LAUNCHED();
    CU(cudaMemcpyAsync(d_meta_scratch, meta, size_t(n) * sizeof(mrag_rowmeta), cudaMemcpyHostToDevice, s));
    scatter_meta_kernel<<<unsigned(ceil_div(n, 256)), 256, 0, s>>>(d_meta_scratch, n, first, x->cols);
    LAUNCHED();
    return MRAG_OK;
}

static int append_common(mrag_index* x, const float* rows, bool rows_on_device, int64_t n,
                         const mrag_rowmeta* meta, int64_t* first_row, cudaStream_t user_stream) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_append: null index");
    if (n < 0) return fail(MRAG_ERR_ARG, "mrag_append: n < 0");
    if (n == 0) { if (first_row) *first_row = x->size; return MRAG_OK; }
    if (!rows) return fail(MRAG_ERR_ARG, "mrag_append: rows is null");
    std::unique_lock<std::shared_mutex> wl(x->lock);
    if (x->size + n > x->capacity)
        return fail(MRAG_ERR_OOM, "mrag_append: %lld + %lld rows exceed capacity %lld", (long long)x->size,
                    (long long)n, (long long)x->capacity);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_append: cudaSetDevice(%d) failed", x->device);
    cudaStream_t s = x->wstream;
    if (rows_on_device && user_stream) {
        // order our stream after the producer of d_rows
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e1 = cudaEventRecord(ev, user_stream);
        if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(s, ev, 0);
        cudaEventDestroy(ev);                            // (released on the error path too)
        if (e1 != cudaSuccess) return fail(MRAG_ERR_CUDA, "mrag_append_device: cannot order the write stream after the caller's: %s", cudaGetErrorString(e1));
    }
    // default metadata: every row its own document, no codes, vector present
    std::vector<mrag_rowmeta> defmeta;
    const int64_t chunk = rows_on_device
        ? std::min<int64_t>(n, int64_t(1) << 20)
        : std::max<int64_t>(1, std::min<int64_t>(n, (int64_t(64) << 20) / (int64_t(x->dim) * 4)));
    mrag_rowmeta* d_meta = nullptr;
    float* d_stage = nullptr;
    CU(cudaMalloc(&d_meta, size_t(chunk) * sizeof(mrag_rowmeta)));
    if (!rows_on_device) {
        cudaError_t e = cudaMalloc(&d_stage, size_t(chunk) * x->dim * 4);
        if (e != cudaSuccess) { cudaFree(d_meta); return fail(MRAG_ERR_OOM, "mrag_append: staging alloc failed"); }
    }
    int rc = MRAG_OK;
    int64_t max_doc = x->n_docs - 1;
    for (int64_t off = 0; off < n && rc == MRAG_OK; off += chunk) {
        const int64_t m = std::min(chunk, n - off);
        const mrag_rowmeta* mp;
        if (meta) {
            mp = meta + off;
        } else {
            defmeta.resize(size_t(m));
            for (int64_t i = 0; i < m; ++i) {
                mrag_rowmeta r{};
                r.doc_idx = uint32_t(x->size + off + i);
                r.payer = MRAG_CODE_NONE; r.state = 0xFF; r.program = 0xFF; r.authority = 0xFF; r.source_type = 0xFF;
                r.valid = 1;
                defmeta[size_t(i)] = r;
            }
            mp = defmeta.data();
        }
        for (int64_t i = 0; i < m; ++i) max_doc = std::max<int64_t>(max_doc, mp[i].doc_idx);
        const float* src = rows + off * x->dim;
        if (!rows_on_device) {
            cudaError_t e = cudaMemcpyAsync(d_stage, src, size_t(m) * x->dim * 4, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) { rc = fail(MRAG_ERR_CUDA, "mrag_append: H2D failed: %s", cudaGetErrorString(e)); break; }
            src = d_stage;
        }
        rc = append_device_locked(x, src, m, mp, x->size + off, s, d_meta);
        // the host meta chunk / staging buffer are reused by the next iteration
        if (rc == MRAG_OK) {
            cudaError_t e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_append: %s", cudaGetErrorString(e));
        }
    }
    cudaFree(d_meta);
    if (d_stage) cudaFree(d_stage);
    if (rc != MRAG_OK) return rc;
    if (first_row) *first_row = x->size;
    x->size += n;
    x->n_docs = max_doc + 1;
    return MRAG_OK;
}

extern "C" int mrag_append(mrag_index* x, const float* rows, int64_t n, const mrag_rowmeta* meta, int64_t* first_row) {
    return append_common(x, rows, false, n, meta, first_row, nullptr);
}

extern "C" int mrag_append_device(mrag_index* x, const void* d_rows_f32, int64_t n, const mrag_rowmeta* meta,
                                  int64_t* first_row, void* stream) {
    return append_common(x, static_cast<const float*>(d_rows_f32), true, n, meta, first_row,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int mrag_set_doc_tags(mrag_index* x, int64_t first_doc, const uint64_t* bits, int64_t n_docs) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_set_doc_tags: null index");
    if (first_doc < 0 || n_docs < 0) return fail(MRAG_ERR_ARG, "mrag_set_doc_tags: negative range");
    if (n_docs == 0) return MRAG_OK;
    if (!bits) return fail(MRAG_ERR_ARG, "mrag_set_doc_tags: bits is null");
    if (first_doc + n_docs > (int64_t(1) << 32)) return fail(MRAG_ERR_ARG, "mrag_set_doc_tags: doc index overflow");
    std::unique_lock<std::shared_mutex> wl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_set_doc_tags: cudaSetDevice failed");
    const int64_t need = first_doc + n_docs;
    if (need > x->tag_docs_cap) {
        int64_t cap = std::max<int64_t>(need + need / 2, 1024);
        uint64_t* p = nullptr;
        CU(cudaMalloc(&p, size_t(cap) * MRAG_TAG_WORDS * 8));
        CU(cudaMemsetAsync(p, 0, size_t(cap) * MRAG_TAG_WORDS * 8, x->wstream));
        if (x->doc_tags) {
            CU(cudaMemcpyAsync(p, x->doc_tags, size_t(x->n_tag_docs) * MRAG_TAG_WORDS * 8, cudaMemcpyDeviceToDevice, x->wstream));
            CU(cudaStreamSynchronize(x->wstream));
            cudaFree(x->doc_tags);
        }
        x->doc_tags = p;
        x->tag_docs_cap = cap;
    }
    CU(cudaMemcpyAsync(x->doc_tags + size_t(first_doc) * MRAG_TAG_WORDS, bits, size_t(n_docs) * MRAG_TAG_WORDS * 8,
                       cudaMemcpyHostToDevice, x->wstream));
    CU(cudaStreamSynchronize(x->wstream));
    x->n_tag_docs = std::max(x->n_tag_docs, need);
    return MRAG_OK;
}

extern "C" int mrag_tombstone_doc(mrag_index* x, uint32_t doc_idx, int64_t* n_rows) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_tombstone_doc: null index");
    if (n_rows) *n_rows = 0;
    std::unique_lock<std::shared_mutex> wl(x->lock);
    if (x->size == 0) return MRAG_OK;
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_tombstone_doc: cudaSetDevice failed");
    unsigned long long* d_hit = nullptr;
    CU(cudaMalloc(&d_hit, 8));
    CU(cudaMemsetAsync(d_hit, 0, 8, x->wstream));
    tombstone_kernel<<<unsigned(ceil_div(x->size, 256)), 256, 0, x->wstream>>>(x->cols.doc_idx, x->size, doc_idx,
                                                                              x->cols.valid, x->cols.live, d_hit);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    unsigned long long hit = 0;
    cudaError_t e = cudaMemcpyAsync(&hit, d_hit, 8, cudaMemcpyDeviceToHost, x->wstream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(x->wstream);
    cudaFree(d_hit);
    if (e != cudaSuccess) return fail(MRAG_ERR_CUDA, "mrag_tombstone_doc: %s", cudaGetErrorString(e));
    if (n_rows) *n_rows = int64_t(hit);
    return MRAG_OK;
}

// Rows that still exist / that still have a vector, of the mrag_size() slots in use: the shard is append-only (a
// re-published document gets new slots, publish.py:310-362), so size - live is what a rebuild would reclaim.
extern "C" int mrag_live_rows(mrag_index* x, int64_t* live_rows, int64_t* rows_with_vector) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_live_rows: null index");
    if (live_rows) *live_rows = 0;
    if (rows_with_vector) *rows_with_vector = 0;
    std::shared_lock<std::shared_mutex> rl(x->lock);
    if (x->size == 0) return MRAG_OK;
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_live_rows: cudaSetDevice failed (no CPU path)");
    unsigned long long* d_out = nullptr;
    CU(cudaMalloc(&d_out, 16));
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemsetAsync(d_out, 0, 16, x->wstream);
    if (e == cudaSuccess) {
        const int64_t nwords = ceil_div(x->size, 32);
        count_bits_kernel<<<unsigned(std::min<int64_t>(1024, ceil_div(nwords, 256))), 256, 0, x->wstream>>>(x->cols.live, x->cols.valid, x->size, d_out);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaMemcpyAsync(h, d_out, 16, cudaMemcpyDeviceToHost, x->wstream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(x->wstream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(MRAG_ERR_CUDA, "mrag_live_rows: %s", cudaGetErrorString(e));
    if (live_rows) *live_rows = int64_t(h[0]);
    if (rows_with_vector) *rows_with_vector = int64_t(h[1]);
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// workspaces
// ------------------------------------------------------------------------------------------
// The kernels of a search alternate with the tensor-core scans, which need the SM's largest shared-memory carve-out; an SM
// only changes its carve-out while it is empty, so every small kernel of the chain asks for the same (maximal) carve-out
// instead of its own default (MRAG_CARVEOUT=0: leave the defaults, for A/B).
static int set_search_chain_carveout(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return MRAG_OK;
    const char* e = getenv("MRAG_CARVEOUT");
    if (!(e && e[0] == '0')) {
        const int c = cudaSharedmemCarveoutMaxShared;
        CU(cudaFuncSetAttribute(query_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(filter_mask_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(pool_bitmap_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(nan_tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(remap_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(xmerge_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(rescore_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(rescore_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CU(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c));
    }
    done[device] = true;
    return MRAG_OK;
}

static Workspace* acquire_ws(mrag_index* x, cudaStream_t want_stream) {
    std::lock_guard<std::mutex> l(x->ws_lock);
    Workspace* pick = nullptr;
    for (Workspace* w : x->pool) {
        if (w->in_use) continue;
        // a workspace left running by a NO_SYNC call may be reused in stream order on the same
        // stream, or by anyone once its end event has completed
        if (w->last_stream == nullptr || (want_stream && w->last_stream == want_stream)) { pick = w; break; }
        if (w->ev.recorded && cudaEventQuery(w->ev.e[3]) == cudaSuccess) { w->last_stream = nullptr; pick = w; break; }
    }
    if (!pick) {
        pick = new (std::nothrow) Workspace();
        if (!pick) return nullptr;
        if (pick->ev.create() != 0 ||
            cudaStreamCreateWithFlags(&pick->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
            pick->release();
            delete pick;
            return nullptr;
        }
        x->pool.push_back(pick);
    }
    pick->in_use = true;
    return pick;
}

static void release_ws(mrag_index* x, Workspace* w, cudaStream_t running_on) {
    std::lock_guard<std::mutex> l(x->ws_lock);
    w->last_stream = running_on;   // nullptr when the call synchronised
    w->in_use = false;
}

// ------------------------------------------------------------------------------------------
// filter -> device form
// ------------------------------------------------------------------------------------------
static void to_dev_filter(const mrag_filter* f, DevFilter* d) {
    memset(d, 0, sizeof *d);
    d->flags = f->flags;
    memcpy(d->payer_any, f->payer_any, sizeof d->payer_any);
    memcpy(d->payer_alt_any, f->payer_alt_any, sizeof d->payer_alt_any);
    d->alt_state = f->alt_state; d->state_eq = f->state_eq; d->program_eq = f->program_eq;
    d->authority_eq = f->authority_eq; d->source_type_eq = f->source_type_eq;
    d->doc_eq = f->doc_eq;
    memcpy(d->tag_state_any, f->tag_state_any, sizeof d->tag_state_any);
    memcpy(d->tag_program_any, f->tag_program_any, sizeof d->tag_program_any);
    memcpy(d->tag_payer_any, f->tag_payer_any, sizeof d->tag_payer_any);
    memcpy(d->tag_any, f->tag_any, sizeof d->tag_any);
}

// builds the row bitmap for `f` into `mask_out` (ceil(n/32) words) on stream s
static int build_mask(mrag_index* x, Workspace* w, const mrag_filter* f, int64_t n, uint32_t* mask_out,
                      unsigned long long* d_npass, cudaStream_t s, bool include_null_vec = false) {
    DevFilter df;
    to_dev_filter(f, &df);
    const uint32_t* pool_bits = nullptr;
    if (df.flags & MRAG_F_DOC_POOL_HANDLE) {
        // the pool is already a document bitmap on this device (mrag_pool_build): nothing to marshal
        const mrag_pool* p = f->pool;
        if (df.flags & MRAG_F_DOC_POOL) return fail(MRAG_ERR_ARG, "mrag_filter: MRAG_F_DOC_POOL and MRAG_F_DOC_POOL_HANDLE exclude each other");
        if (!p || p->idx != x || p->level < 0 || p->level >= MRAG_POOL_LEVELS || !p->bits)
            return fail(MRAG_ERR_ARG, "mrag_filter: pool handle is null, belongs to another index, or has no level selected");
        const uint32_t* src = p->bits + size_t(p->level) * p->words;
        const int64_t words = ceil_div(std::max<int64_t>(x->n_docs, 1), 32);
        if (words > p->words) {                          // documents were added since the pool was built: they are not in it
            if (w->pool_bits.reserve(size_t(words))) return MRAG_ERR_OOM;
            CU(cudaMemsetAsync(w->pool_bits.p, 0, size_t(words) * 4, s));
            CU(cudaMemcpyAsync(w->pool_bits.p, src, size_t(p->words) * 4, cudaMemcpyDeviceToDevice, s));
            src = w->pool_bits.p;
        }
        pool_bits = src;
        df.flags = (df.flags & ~uint32_t(MRAG_F_DOC_POOL_HANDLE)) | MRAG_F_DOC_POOL;
    } else if (df.flags & MRAG_F_DOC_POOL) {
        if (f->n_doc_pool < 0 || (f->n_doc_pool > 0 && !f->doc_pool))
            return fail(MRAG_ERR_ARG, "mrag_filter: doc_pool is null or n_doc_pool < 0");
        const int64_t words = ceil_div(std::max<int64_t>(x->n_docs, 1), 32);
        if (w->pool_bits.reserve(size_t(words))) return MRAG_ERR_OOM;
        CU(cudaMemsetAsync(w->pool_bits.p, 0, size_t(words) * 4, s));
        if (f->n_doc_pool > 0) {
            if (w->pool.reserve(size_t(f->n_doc_pool))) return MRAG_ERR_OOM;
            CU(cudaMemcpyAsync(w->pool.p, f->doc_pool, size_t(f->n_doc_pool) * 4, cudaMemcpyHostToDevice, s));
            pool_bitmap_kernel<<<unsigned(ceil_div(f->n_doc_pool, 256)), 256, 0, s>>>(w->pool.p, f->n_doc_pool,
                                                                                     w->pool_bits.p, x->n_docs);
            LAUNCHED();
        }
        pool_bits = w->pool_bits.p;
    }
    const int64_t n32 = ceil_div(n, 32) * 32;
    filter_mask_kernel<<<unsigned(ceil_div(n32, 256)), 256, 0, s>>>(df, x->cols, n, pool_bits, x->doc_tags,
                                                                   x->n_tag_docs, mask_out, d_npass, include_null_vec ? 1 : 0);
    LAUNCHED();
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// scan dispatch
// ------------------------------------------------------------------------------------------
static const int kMaxSmem = 232448;   // 227 KB

template <int DT, int NQ, int HYB = 0>
static int launch_gemv(const ScanArgs& a, int grid, cudaStream_t s) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = gemv_smem_bytes(NQ, a.ld, a.kp, HYB != 0);
    if (smem > size_t(kMaxSmem))
        return fail(MRAG_ERR_ARG, "scan: %zu B of shared memory needed (dim too large for k)", smem);
    if (dev < 64 && !attr_set[dev]) {
        CU(cudaFuncSetAttribute((scan_gemv_kernel<DT, NQ, HYB>), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_set[dev] = true;
    }
    scan_gemv_kernel<DT, NQ, HYB><<<grid, kGemvThreads, smem, s>>>(a);
    LAUNCHED();
    return MRAG_OK;
}

static int gemv_nq_for(int nq, int ld, int kp, bool hyb = false) {
    // widest query group whose buffers fit in shared memory
    for (int g : {4, 2, 1})
        if ((nq >= g || g == 1) && gemv_smem_bytes(g, ld, kp, hyb) <= size_t(kMaxSmem)) return g;
    return 1;
}

static int run_scan_gemv(mrag_index* x, ScanArgs a, int nq, int grid, cudaStream_t s) {
    int q0 = 0;
    while (q0 < nq) {
        const int left = nq - q0;
        const int g = gemv_nq_for(left >= 3 ? 4 : left, a.ld, a.kp);
        a.q0 = q0;
        a.nq = std::min(g, left);
        int rc;
        if (x->dtype == MRAG_BF16) {
            rc = g == 4 ? launch_gemv<1, 4>(a, grid, s) : g == 2 ? launch_gemv<1, 2>(a, grid, s) : launch_gemv<1, 1>(a, grid, s);
        } else {
            rc = g == 4 ? launch_gemv<0, 4>(a, grid, s) : g == 2 ? launch_gemv<0, 2>(a, grid, s) : launch_gemv<0, 1>(a, grid, s);
        }
        if (rc != MRAG_OK) return rc;
        q0 += a.nq;
    }
    return MRAG_OK;
}

static int mma_stages_for(int cap, bool ksplit = false) {
    const size_t fixed = mma_smem_bytes(0, cap, ksplit);
    if (fixed + 4 * size_t(kMmaStageBytes) > size_t(kMaxSmem)) return 0;
    // leave 6 KB of the SM's shared memory free when the ring is deep anyway: the cross-rank exchange kernel of the previous
    // search (1.5 KB + the per-block reserve) can then be resident beside a scan CTA
    const size_t budget = (fixed + 16 * size_t(kMmaStageBytes) + 6144 <= size_t(kMaxSmem)) ? size_t(kMaxSmem) - 6144 : size_t(kMaxSmem);
    return int(std::min<size_t>(24, (budget - fixed) / kMmaStageBytes));
}

// KS: the k-split pair kernel (rows of 769 .. 1536 elements): `grid` counts CTAs and is even, clusters of 2
template <int KREG, bool SO = false, bool KS = false>
static int launch_scan_mma(mrag_index* x, MmaArgs a, int nq, int grid, cudaStream_t s) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        CU(cudaFuncSetAttribute((scan_mma_kernel<KREG, SO, KS>), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_set[dev] = true;
    }
    if (KREG > 0) a.cap = 0;                   // candidates live in registers: no shared-memory buffers
    a.stages = mma_stages_for(a.cap, KS);
    if (a.stages < 4) return fail(MRAG_ERR_ARG, "scan_mma: candidate buffers of k = %d leave no room for the TMA ring", a.k);
    const size_t smem = mma_smem_bytes(a.stages, a.cap, KS);
    for (int q0 = 0; q0 < nq; q0 += kMmaQueries) {
        a.q0 = q0;
        a.nq = std::min(kMmaQueries, nq - q0);
        if constexpr (KS) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(unsigned(grid)); cfg.blockDim = dim3(kMmaKsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CU(cudaLaunchKernelEx(&cfg, scan_mma_kernel<KREG, SO, KS>, x->tmap, a));
        } else {
            scan_mma_kernel<KREG, SO, KS><<<grid, kMmaThreads, smem, s>>>(x->tmap, a);
        }
        LAUNCHED();
    }
    return MRAG_OK;
}

// reg_topk: keep each query's top-k in registers (k <= 16).  It wins whenever most rows a CTA sees
// are still candidates (short streams: small shards and the sampling pass), because all 32 queries
// of a warp insert in lock step; on long streams the shared-memory buffers + a sampled bound win.
static int run_scan_mma(mrag_index* x, const MmaArgs& a, int nq, int grid, bool reg_topk, cudaStream_t s) {
    if (a.ld > kMmaMaxLd)
        return (reg_topk && a.k <= kMmaRegK) ? launch_scan_mma<kMmaRegK, false, true>(x, a, nq, grid, s)
                                             : launch_scan_mma<0, false, true>(x, a, nq, grid, s);
    return (reg_topk && a.k <= kMmaRegK) ? launch_scan_mma<kMmaRegK>(x, a, nq, grid, s)
                                         : launch_scan_mma<0>(x, a, nq, grid, s);
}

// Threshold sampling for large shards: scan a strided sample of the tiles first, take the k-th best
// score of that sample per query, and let the full scan admit only rows scoring at least that.
// (k rows of the sample already reach the bound, so nothing below it can be in the top-k.)  It cuts
// the candidates a CTA buffers from ~k ln(n/k), most of them in a costly warm-up, to ~k * stride / #CTAs.
//   k <= 16: two tiles per CTA with the register top-k kernel (about 1/500 of a 10M-row shard)
//   k  > 16: every 64th tile with the buffer kernel
static uint32_t mma_sleep_ns() {
    static const uint32_t v = [] {
        const char* e = getenv("MRAG_MMA_SLEEP_NS");
        return (e && *e) ? uint32_t(strtoul(e, nullptr, 10)) : 0x989680u;
    }();
    return v;
}

static int64_t sample_min_tiles(int num_sms) {
    // default: shards of >= 8 tiles per SM (~76k rows) -- measured (r1h): from there on the sampled buffer kernel
    // beats the register top-k kernel by 2x .. 3.5x; MRAG_SAMPLE_MIN_TILES overrides (tests)
    static const int64_t env = [] {
        const char* e = getenv("MRAG_SAMPLE_MIN_TILES");
        return (e && *e) ? std::max<int64_t>(1, atoll(e)) : int64_t(0);
    }();
    return env ? env : int64_t(8) * num_sms;
}


// ORDER BY .. LIMIT over the per-producer lists: one block per query, or two levels when the producer
// set is wide (many blocks in flight sorting <= 2048 keys each, then one block per query)
// (nblocks = queries served by this launch; the group lists are indexed by query id, so they are sized by m.nq)
static int launch_merge(Workspace* w, MergeArgs m, int nq, cudaStream_t s) {
    // (with a bound from the scan the survivors are few: one block per query streams all the lists through its prefilter)
    if (int64_t(m.P) * m.kp > kMergeSlots && !m.gthr_out && !m.thr_in) {      // everything fits one block's shared memory otherwise
        MergeArgs m1 = m;
        m1.Pg = std::max(2, (kMergeSlots / 2) / m.kp);
        const int groups = int(ceil_div(m.P, m1.Pg));
        if (w->part2.reserve(size_t(std::max(nq, m.nq)) * groups * m.kp)) return MRAG_ERR_OOM;
        m1.part_out = w->part2.p;
        merge_kernel<<<dim3(unsigned(nq), unsigned(groups)), kMergeThreads, 0, s>>>(m1);
        LAUNCHED();
        m.part = w->part2.p; m.P = groups;
    }
    merge_kernel<<<nq, kMergeThreads, 0, s>>>(m);
    LAUNCHED();
    return MRAG_OK;
}

// largest k of the candidate-generating path: K' = 2 k nominees + 32 slack keys must fit a warp's rank sort (256 keys)
static const int kMma128MaxK = 106;

// fp32 shards: nominate from the bf16 shadow on the CUDA cores when the shard is big enough for the halved traffic to
// pay for the extra launches (rescoring, finalize, no-op rescan): >= 128M elements (256 MB of shadow) by default;
// MRAG_SHADOW_GEMV_MIN_ELEMS overrides (0 = always, tests; a huge value pins the exact fp32 scan).
static bool shadow_gemv_wanted(int64_t n, int ld) {
    const char* e = getenv("MRAG_SHADOW_GEMV_MIN_ELEMS");
    const long long min_elems = (e && *e) ? atoll(e) : (128ll << 20);
    return n * int64_t(ld) >= min_elems;
}

static int approx_min_nq() {
    static const int v = [] {
        const char* e = getenv("MRAG_APPROX_MIN_NQ");
        return (e && *e) ? std::max(1, atoi(e)) : 5;
    }();
    return v;
}

template <int KREG>
static int launch_scan_mma128(mrag_index* x, MmaArgs a, int nq, int grid, cudaStream_t s) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        CU(cudaFuncSetAttribute(scan_mma128_kernel<KREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_set[dev] = true;
    }
    if (KREG > 0) { a.cap = 0; a.gcand = nullptr; }
    const int smem_cap = a.gcand ? 0 : a.cap;           // candidate buffers in global memory take no shared memory
    const int kblocks = a.ld / kMmaKBlock;
    a.kbs = kblocks % 4 == 0 ? 4 : kblocks % 3 == 0 ? 3 : kblocks % 2 == 0 ? 2 : 1;     // k-blocks per stage
    const size_t fixed = mma128_smem_bytes(0, smem_cap);
    const size_t stage = size_t(a.kbs) * kMmaStageBytes;
    if (fixed + 2 * stage > size_t(kMaxSmem)) { a.kbs = 1; }
    const size_t stage1 = size_t(a.kbs) * kMmaStageBytes;
    if (fixed + 2 * stage1 > size_t(kMaxSmem)) return fail(MRAG_ERR_ARG, "scan_mma128: candidate buffers do not fit");
    a.stages = int(std::min<size_t>(24, (size_t(kMaxSmem) - fixed) / stage1));
    const size_t smem = mma128_smem_bytes(a.stages, smem_cap, a.kbs);
    for (int q0 = 0; q0 < nq; q0 += kMma128Queries) {
        a.q0 = q0;
        a.nq = std::min(kMma128Queries, nq - q0);
        scan_mma128_kernel<KREG><<<grid, kMmaThreads, smem, s>>>(x->tmap, a);
        LAUNCHED();
    }
    return MRAG_OK;
}

// the CTA-pair scan: 256 queries per pass (two CTAs of a cluster share every corpus tile)
template <int KBS>
static int launch_scan_mma256_t(mrag_index* x, MmaArgs a, int nq, int npairs, cudaStream_t s) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        CU(cudaFuncSetAttribute(scan_mma256_kernel<KBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_set[dev] = true;
    }
    const int smem_cap = a.gcand ? 0 : a.cap;
    a.kbs = KBS;
    const size_t fixed = mma256_smem_bytes(0, smem_cap);
    const size_t stage = size_t(KBS) * kMma256StageBytes;
    if (fixed + 3 * stage > size_t(kMaxSmem)) return fail(MRAG_ERR_ARG, "scan_mma256: candidate buffers do not fit");
    a.stages = int(std::min<size_t>(24, (size_t(kMaxSmem) - fixed) / stage));
    const size_t smem = mma256_smem_bytes(a.stages, smem_cap, KBS);
    a.P = npairs;
    for (int q0 = 0; q0 < nq; q0 += 2 * kMma128Queries) {
        a.q0 = q0;
        a.nq = std::min(2 * kMma128Queries, nq - q0);
        scan_mma256_kernel<KBS><<<2 * npairs, kMmaThreads, smem, s>>>(x->tmap32, a);
        LAUNCHED();
    }
    return MRAG_OK;
}

static int launch_scan_mma256(mrag_index* x, const MmaArgs& a, int nq, int npairs, cudaStream_t s) {
    const int kblocks = a.ld / kMmaKBlock;              // k-blocks per stage: the largest of 4, 3, 2, 1 that divides them
    return kblocks % 4 == 0 ? launch_scan_mma256_t<4>(x, a, nq, npairs, s)
         : kblocks % 3 == 0 ? launch_scan_mma256_t<3>(x, a, nq, npairs, s)
         : kblocks % 2 == 0 ? launch_scan_mma256_t<2>(x, a, nq, npairs, s)
                            : launch_scan_mma256_t<1>(x, a, nq, npairs, s);
}

// the CTA-pair scan with 128-row tiles (scan_mma256w.cuh): one accumulator, 8 select warps per CTA, candidate buffers
// in global memory (a.gcand, [2 * npairs][2][128][cap]); writes 2 * npairs partial lists per query
template <int KBS>
static int launch_scan_mma256w_t(mrag_index* x, MmaArgs a, int nq, int npairs, cudaStream_t s) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        CU(cudaFuncSetAttribute(scan_mma256w_kernel<KBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        attr_set[dev] = true;
    }
    if (!a.gcand) return fail(MRAG_ERR_ARG, "scan_mma256w: global candidate buffers required");
    a.kbs = KBS;
    const int ks = mma256w_smem_kblocks(a.ld);          // query k-blocks kept in shared memory (beyond 8 in tensor memory)
    const size_t fixed = mma256w_smem_bytes(0, KBS, ks);
    const size_t stage = size_t(KBS) * kMmaWStageBytes;
    a.stages = int(std::min<size_t>(24, (size_t(kMaxSmem) - fixed) / stage));
    if (a.stages < 2) return fail(MRAG_ERR_ARG, "scan_mma256w: pipeline does not fit");
    const size_t smem = mma256w_smem_bytes(a.stages, KBS, ks);
    a.P = 2 * npairs;
    for (int q0 = 0; q0 < nq; q0 += 2 * kMma128Queries) {
        a.q0 = q0;
        a.nq = std::min(2 * kMma128Queries, nq - q0);
        scan_mma256w_kernel<KBS><<<2 * npairs, kMmaWThreads, smem, s>>>(x->tmap, a);
        LAUNCHED();
    }
    return MRAG_OK;
}

static int launch_scan_mma256w(mrag_index* x, const MmaArgs& a, int nq, int npairs, cudaStream_t s) {
    const int kblocks = a.ld / kMmaKBlock;
    return kblocks % 4 == 0 ? launch_scan_mma256w_t<4>(x, a, nq, npairs, s)
         : kblocks % 3 == 0 ? launch_scan_mma256w_t<3>(x, a, nq, npairs, s)
         : kblocks % 2 == 0 ? launch_scan_mma256w_t<2>(x, a, nq, npairs, s)
                            : launch_scan_mma256w_t<1>(x, a, nq, npairs, s);
}

// MRAG_EVENTS: which of a search's phase events are recorded in its stream.  2 (default): start / prepared / scanned / end;
// 1: prepared / scanned / end (the scan phase only, what the roofline needs); 0: end only (workspace reuse needs it).
// Measured r2q (1.25M x 768, 64 queries): see profiles/README.md.
static int events_level() {
    static const int v = [] { const char* e = getenv("MRAG_EVENTS"); return (e && *e) ? atoi(e) : 2; }();
    return v;
}

// Large batches: candidate generation on the tensor cores (128 queries per pass over the bf16 rows or
// the bf16 shadow), exact rescoring of the K' = k + 32 nominees from the primary rows, certificate,
// exact CUDA-core rescan of the queries that fail it.  Results are EXACT (same arithmetic as scan_gemv).
// gen_gemv: nominate with the CUDA-core scan over the bf16 shadow instead (fp32 shards of any width, small batches):
// half the bytes of the exact fp32 scan, eps = 2^-9 (the shadow's rounding; the query stays fp32).
static int search_approx_rescore(mrag_index* x, Workspace* w, EventSet& ev, int nq, int k, const uint32_t* mask,
                                 float* d_scores, int64_t* d_rows, int32_t* d_counts, cudaStream_t s, bool gen_gemv) {
    const int64_t n = x->size;
    const int ld = x->ld;
    // nominees per query: k + 32, or 1.5 k for large k (the number of rows within eps of the k-th best grows with k)
    // (k > 64: 2k -- the number of rows within eps of the k-th best grows with k; MRAG_KC_2K=0 restores 1.5k for A/B)
    static const bool kc_2k = [] { const char* e = getenv("MRAG_KC_2K"); return !(e && e[0] == '0'); }();
    const int kc = (k > 64 && kc_2k) ? 2 * k : std::max(k + 32, k + k / 2), kpc = host_next_pow2(kc);
    const int cap = kc + kMma128Slack;
    const bool global_cand = mma128_smem_bytes(12, cap) > size_t(kMaxSmem);      // keep >= 12 stages (96 KB) in flight
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>(x->num_sms, ceil_div(n, kMmaTileRows))));
    const int64_t nwords = ceil_div(n, 32);
    const int ggrid = int(std::max<int64_t>(1, std::min<int64_t>(x->num_sms, ceil_div(nwords, kGemvWarps))));
    const int kp = std::max(8, host_next_pow2(k));
    const char* eps_env = getenv("MRAG_APPROX_EPS_SCALE");     // tests: a huge scale sends every query to the rescan
    const float eps_scale = (eps_env && *eps_env) ? float(atof(eps_env)) : 1.0f;
    // |approx - exact| <= 2^-9 (bf16 query) + 2^-9 (the shadow's rounding, fp32 corpora) + accumulation slop
    const float eps = ((gen_gemv ? 0.0f : 0x1p-9f) + (x->dtype == MRAG_F32 ? 0x1p-9f : 0.0f) + 5e-5f) * eps_scale;

    if (w->part.reserve(size_t(nq) * std::max(std::max(grid + 1, ggrid) * kpc, ggrid * kp)) || w->gthr.reserve(size_t(nq)) ||
        w->cscores.reserve(size_t(nq) * kc) || w->crows.reserve(size_t(nq) * kc) || w->ccounts.reserve(size_t(nq)) ||
        w->ckeys.reserve(size_t(nq) * kc) || w->fb.reserve(size_t(nq) + 1))
        return MRAG_ERR_OOM;
    CU(cudaMemsetAsync(w->gthr.p, 0, size_t(nq) * 4, s));
    CU(cudaMemsetAsync(w->fb.p, 0, sizeof(int), s));
    int rc;
    bool use_pairs = false, sampled_bound = false;
    int npairs = 0, nparts = 0;
    if (gen_gemv) {
        ScanArgs ga{};
        ga.rows = x->shadow; ga.n = n; ga.ld = ld; ga.mask = mask; ga.q = w->qpad.p; ga.qinv = w->qinv.p; ga.ub = nullptr;
        ga.part = w->part.p; ga.k = kc; ga.kp = kpc; ga.P = ggrid;
        int q0 = 0;
        while (q0 < nq) {
            const int left = nq - q0;
            const int g = gemv_nq_for(left >= 3 ? 4 : left, ld, kpc);
            ga.q0 = q0; ga.nq = std::min(g, left);
            rc = g == 4 ? launch_gemv<1, 4>(ga, ggrid, s) : g == 2 ? launch_gemv<1, 2>(ga, ggrid, s) : launch_gemv<1, 1>(ga, ggrid, s);
            if (rc != MRAG_OK) return rc;
            q0 += ga.nq;
        }
    } else {
    MmaArgs a{};
    a.n = n; a.ld = ld; a.mask = mask; a.inv_norm = x->inv_norm; a.q = w->qpad.p; a.qinv = w->qinv.p;
    a.part = w->part.p; a.k = kc; a.kp = kpc; a.P = grid; a.cap = cap;
    a.gthr = w->gthr.p; a.tile_mul = 1; a.sleep_ns = mma_sleep_ns();
    if (global_cand) {
        if (w->gcand.reserve(size_t(grid) * kMma128Queries * cap)) return MRAG_ERR_OOM;
        a.gcand = w->gcand.p;
    }
    if (getenv("MRAG_SCAN_STATS")) {
        if (w->stats.reserve(40)) return MRAG_ERR_OOM;
        CU(cudaMemsetAsync(w->stats.p, 0, 40 * 8, s));
        a.stats = w->stats.p;
        t_stats_ptr = w->stats.p;
    }
    const int64_t tiles = ceil_div(n, kMmaTileRows);
    // (the group-maxima bound of the exact scan was tried here too, with K' <= 64 groups: r2n, 10M x 768, k = 10: 1-3 % slower
    //  than the sampling launches at 128 / 256 / 1024 queries and 0.52 against 0.35 ms on config 2, whose tag filter leaves a
    //  CTA ~10 tiles -- with K' = 42 groups of 3.5 producers the bound is 4x looser in rank than the union's and the first
    //  tile of every producer is admitted whole)
    if (tiles >= sample_min_tiles(x->num_sms)) {
        // admission bound from a strided sample, same arithmetic, the best kMmaSampleK scores per (query, CTA) in registers:
        // the K'-th best of the union of the per-CTA lists is reached by K' known rows, so the full pass may skip anything
        // below it (a CTA holds 1/#CTAs of the sample; a list that truncates only loosens the bound)
        MmaArgs sa = a;
        sa.stats = nullptr;
        // up to 32 tiles per CTA, at most ~3 % of the shard (16 and ~1.5 % for large k): the k'-th best of the sample is the admission bound of the full pass, and every row
        // above it costs a buffer append and, every `slack` appends, a compaction.  Measured r1q (10M x 768, k = 10): 4 -> 32
        // tiles per CTA = 3.18 -> 2.49 ms at 128 queries, 5.0 -> 3.36 ms at 256 (pair scan); 64 tiles cost more than they save
        const char* pc_env = getenv("MRAG_SAMPLE128_PER_CTA");
        const int per_cta = (pc_env && *pc_env) ? std::max(1, atoi(pc_env)) :
            int(std::max<int64_t>(1, std::min<int64_t>(kc > 64 ? 16 : 32, ceil_div(tiles, (kc > 64 ? 64 : 32) * int64_t(x->num_sms)))));
        sa.tile_mul = int(std::max<int64_t>(1, tiles / (int64_t(per_cta) * x->num_sms)));
        const int sgrid = int(std::min<int64_t>(x->num_sms, ceil_div(tiles, sa.tile_mul)));
        sa.P = sgrid;
        // per-CTA list length of the sample: 4 scores for k' <= 64 (the union's k'-th best is spread ~k'/148 per CTA), the
        // full 16 beyond (measured r2i, 10M x 768: k = 10, B = 256: 3.66 -> 3.38 ms with 4; k = 100: 7.33 -> 8.48 ms with 4)
        const bool short_lists = kc <= 64;
        sa.k = std::min(kc, short_lists ? kMmaSampleK : kMmaRegK);
        sa.kp = short_lists ? kMmaSampleK : kMmaRegK;
        rc = short_lists ? launch_scan_mma128<kMmaSampleK>(x, sa, nq, sgrid, s) : launch_scan_mma128<kMmaRegK>(x, sa, nq, sgrid, s);
        if (rc != MRAG_OK) return rc;
        MergeArgs sm{};
        sm.part = w->part.p; sm.P = sgrid; sm.kp = sa.kp; sm.nq = nq; sm.k = kc; sm.lk = sa.k; sm.k_total = kc;
        sm.gthr_out = w->gthr.p;     // (overwrites the looser per-CTA bounds the sampling launches left there)
        merge_kernel<<<nq, kMergeThreads, 0, s>>>(sm);
        LAUNCHED();
        sampled_bound = true;
    }
    // more than 128 queries: CTA pairs share every corpus tile (256 queries per pass)
    static const bool pairs_ok = [] { const char* e = getenv("MRAG_MMA256"); return !(e && e[0] == '0'); }();
    use_pairs = pairs_ok && nq > kMma128Queries && x->num_sms >= 2;
    // 128-row tiles (one MMA per 128 rows: half the per-MMA fixed cost per corpus byte) unless MRAG_MMA256W=0
    const char* wide_env = getenv("MRAG_MMA256W");          // read per call: the tests switch between the two pair kernels
    // (large k: two lists of K' + 32 keys per query and CTA in global memory cost more than the wider MMA saves --
    //  measured r1q, k = 100: 4.74 ms per 256 queries against 4.36 ms with the 64-row pair kernel)
    const bool wide_ok = wide_env ? wide_env[0] != '0' : kc <= 64;
    if (use_pairs && wide_ok) {
        npairs = int(std::max<int64_t>(1, std::min<int64_t>(x->num_sms / 2, ceil_div(n, kMmaWTileRows))));
        if (w->gcand.reserve(size_t(2 * npairs) * 2 * kMma128Queries * cap)) return MRAG_ERR_OOM;
        a.gcand = w->gcand.p;
        rc = launch_scan_mma256w(x, a, nq, npairs, s);
        nparts = 2 * npairs;
    } else if (use_pairs) {
        npairs = int(std::max<int64_t>(1, std::min<int64_t>(x->num_sms / 2, tiles)));
        if (global_cand && w->gcand.reserve(size_t(2 * npairs) * kMma128Queries * cap)) return MRAG_ERR_OOM;
        if (global_cand) a.gcand = w->gcand.p;
        rc = launch_scan_mma256(x, a, nq, npairs, s);
        nparts = npairs;
    } else {
        rc = launch_scan_mma128<0>(x, a, nq, grid, s);
    }
    if (rc != MRAG_OK) return rc;
    }
    if (events_level() >= 1) CU(cudaEventRecord(ev.e[2], s));
    // nominees per query, by approximate score
    MergeArgs m{};
    m.part = w->part.p; m.P = gen_gemv ? ggrid : (use_pairs ? nparts : grid); m.kp = kpc; m.nq = nq; m.k = kc; m.k_total = kc; m.k_off = 0;
    m.scores = w->cscores.p; m.rows = w->crows.p; m.counts = w->ccounts.p; m.row_base = 0;
    // sampled scans leave in gthr a score that K' rows of the query reach (the sample's K'-th best, raised by compactions):
    // the nominee merge prefilters with it and needs no first level
    if (sampled_bound) m.thr_in = w->gthr.p;
    rc = launch_merge(w, m, nq, s);
    if (rc != MRAG_OK) return rc;
    RescoreArgs ra{};
    ra.rows = x->rows; ra.ld = ld; ra.inv_norm = x->inv_norm; ra.q = w->qpad.p; ra.qinv = w->qinv.p;
    ra.cand_rows = w->crows.p; ra.cand_scores = w->cscores.p; ra.cand_counts = w->ccounts.p; ra.nq = nq; ra.kc = kc;
    ra.keys = w->ckeys.p;
    const unsigned rblocks = unsigned(ceil_div(int64_t(nq) * kc * 32, 256));
    if (x->dtype == MRAG_BF16) rescore_kernel<1><<<rblocks, 256, 0, s>>>(ra);
    else rescore_kernel<0><<<rblocks, 256, 0, s>>>(ra);
    LAUNCHED();
    FinalizeArgs fa{};
    fa.keys = w->ckeys.p; fa.cand_scores = w->cscores.p; fa.cand_counts = w->ccounts.p; fa.nq = nq; fa.kc = kc; fa.k = k;
    fa.eps = eps; fa.scores = d_scores; fa.rows = d_rows; fa.counts = d_counts; fa.row_base = x->row_base;
    fa.need_tail = w->flags.p; fa.fb_count = w->fb.p; fa.fb_list = w->fb.p + 1; fa.gthr = w->gthr.p;
    finalize_kernel<<<nq, kFinalizeThreads, 0, s>>>(fa);
    LAUNCHED();
    // exact rescan of the queries whose certificate failed (no-op launches when there are none): the first 64 of
    // them in ONE pass of the exact tensor-core scan (bf16 shards), the rest -- and fp32 shards -- on the CUDA cores
    const bool exact_mma = x->dtype == MRAG_BF16 && x->has_tmap;
    int gskip = 0;
    if (exact_mma) {
        MmaArgs e{};
        e.n = n; e.ld = ld; e.mask = mask; e.inv_norm = x->inv_norm; e.q = w->qpad.p; e.qinv = w->qinv.p; e.qhl = (getenv("MRAG_QHL") && getenv("MRAG_QHL")[0] == '0') ? nullptr : w->qhl.p;
        e.part = w->part.p; e.k = k; e.kp = kp; e.P = grid; e.cap = k + kMmaSlack; e.gthr = w->gthr.p; e.tile_mul = 1;
        e.sleep_ns = mma_sleep_ns(); e.qlist = w->fb.p + 1; e.qcount = w->fb.p;
        rc = launch_scan_mma<0>(x, e, /*nq=*/1, grid, s);        // one launch; the kernel reads the list length itself
        if (rc != MRAG_OK) return rc;
        MergeArgs em{};
        em.part = w->part.p; em.P = grid; em.kp = kp; em.nq = nq; em.k = k; em.k_total = k; em.k_off = 0;
        em.scores = d_scores; em.rows = d_rows; em.counts = d_counts; em.row_base = x->row_base;
        em.need_tail = w->flags.p; em.qlist = w->fb.p + 1; em.qcount = w->fb.p; em.q_lo = 0; em.q_hi = kMmaQueries;
        rc = launch_merge(w, em, std::min(nq, kMmaQueries), s);
        if (rc != MRAG_OK) return rc;
        gskip = kMmaQueries;
    }
    if (nq > gskip) {
        if (w->part3.reserve(size_t(nq) * ggrid * kp)) return MRAG_ERR_OOM;
        ScanArgs g{};
        g.rows = x->rows; g.n = n; g.ld = ld; g.mask = mask; g.q = w->qpad.p; g.qinv = w->qinv.p; g.ub = nullptr;
        g.part = w->part3.p; g.k = k; g.kp = kp; g.P = ggrid; g.qlist = w->fb.p + 1; g.qcount = w->fb.p; g.qskip = gskip;
        const int gq = gemv_nq_for(4, ld, kp);
        if (x->dtype == MRAG_BF16)
            rc = gq == 4 ? launch_gemv<1, 4>(g, ggrid, s) : gq == 2 ? launch_gemv<1, 2>(g, ggrid, s) : launch_gemv<1, 1>(g, ggrid, s);
        else
            rc = gq == 4 ? launch_gemv<0, 4>(g, ggrid, s) : gq == 2 ? launch_gemv<0, 2>(g, ggrid, s) : launch_gemv<0, 1>(g, ggrid, s);
        if (rc != MRAG_OK) return rc;
        MergeArgs fm{};
        fm.part = w->part3.p; fm.P = ggrid; fm.kp = kp; fm.nq = nq; fm.k = k; fm.k_total = k; fm.k_off = 0;
        fm.scores = d_scores; fm.rows = d_rows; fm.counts = d_counts; fm.row_base = x->row_base;
        fm.need_tail = w->flags.p; fm.qlist = w->fb.p + 1; fm.qcount = w->fb.p; fm.q_lo = gskip; fm.q_hi = 0;
        rc = launch_merge(w, fm, nq - gskip, s);
        if (rc != MRAG_OK) return rc;
    }
    t_fb_ptr = w->fb.p;
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// search
// ------------------------------------------------------------------------------------------
static int search_locked(mrag_index* x, Workspace* w, EventSet& ev, const float* q, int nq, int k,
                         const mrag_filter* filter, float* scores, int64_t* rows, int32_t* counts,
                         uint32_t options, cudaStream_t s) {
    const bool dev_io = options & MRAG_OPT_DEVICE_IO;
    const int64_t n = x->size;
    const int ld = x->ld;
    const size_t nk = size_t(nq) * k;

    const int evl = events_level();
    if (evl >= 2) CU(cudaEventRecord(ev.e[0], s));
    // ---- queries
    const float* d_q = q;
    if (!dev_io) {
        if (w->qraw.reserve(size_t(nq) * x->dim)) return MRAG_ERR_OOM;
        CU(cudaMemcpyAsync(w->qraw.p, q, size_t(nq) * x->dim * 4, cudaMemcpyHostToDevice, s));
        d_q = w->qraw.p;
    }
    if (w->qpad.reserve(size_t(nq) * ld) || w->qinv.reserve(size_t(nq)) || w->flags.reserve(4) ||
        w->counts.reserve(size_t(nq)) || w->scores.reserve(nk) || w->rows.reserve(nk))
        return MRAG_ERR_OOM;
    // bf16 shards: also the packed hi / lo planes the exact tensor-core scan loads into tensor memory
    uint32_t* qhl = nullptr;
    static const bool qhl_ok = [] { const char* e = getenv("MRAG_QHL"); return !(e && e[0] == '0'); }();    // MRAG_QHL=0: split in the scan (A/B)
    if (qhl_ok && x->has_tmap && x->dtype == MRAG_BF16) {
        if (w->qhl.reserve(size_t(nq) * ld)) return MRAG_ERR_OOM;
        qhl = w->qhl.p;
    }
    // cross-CTA admission bound of the exact tensor-core scan (k <= 16): group maxima (scan_mma.cuh, GroupBound) instead of a
    // threshold-sampling launch.  MRAG_GMAX: 0 = sampling launches as before, 1 = group maxima + register top-k,
    // 2 (default) = group maxima + shared-memory candidate buffers on shards of >= sample_min_tiles tiles
    // (measured r2m, 1.25M x 768, 64 queries, k = 10: 0.378 / 0.480 / 0.349 ms per step)
    static const int gmax_mode = [] { const char* e = getenv("MRAG_GMAX"); return (e && *e) ? atoi(e) : 2; }();
    // MRAG_GMAX_MAXK: largest k served by the group maxima (16 slots up to k = 16, 128 beyond; default: every single-round k).
    // Measured r2x (10M x 768, 64 queries, k = 100): 2.58-2.62 ms per step against 2.65 with the sampling launch (16);
    // the merge phase 0.08 -> 0.03 ms.
    static const int gmax_maxk = [] { const char* e = getenv("MRAG_GMAX_MAXK"); return (e && *e) ? std::min(128, atoi(e)) : 128; }();
    const bool use_gmax = gmax_mode > 0 && x->has_tmap && x->dtype == MRAG_BF16 && k <= gmax_maxk;
    const int gslots = k <= 16 ? 16 : 128;
    if (w->gthr.reserve(size_t(nq)) || (use_gmax && w->gmax.reserve(size_t(nq) * gslots))) return MRAG_ERR_OOM;
    query_prep_kernel<<<unsigned(ceil_div(int64_t(nq) * 32, 128)), 128, 0, s>>>(d_q, nq, x->dim, ld, w->qpad.p,
                                                                               w->qinv.p, nullptr, nq, qhl, w->gthr.p,
                                                                               use_gmax ? w->gmax.p : nullptr, k, gslots, w->flags.p);
    LAUNCHED();

    float* d_scores = dev_io ? scores : w->scores.p;
    int64_t* d_rows = dev_io ? rows : w->rows.p;
    int32_t* d_counts = dev_io ? counts : w->counts.p;

    // ---- WHERE
    const uint32_t* mask = x->cols.valid;     // embedding_vec IS NOT NULL
    if (n > 0 && filter && filter->flags) {
        if (w->mask.reserve(size_t(ceil_div(n, 32)) + 1)) return MRAG_ERR_OOM;
        int rc = build_mask(x, w, filter, n, w->mask.p, nullptr, s);
        if (rc != MRAG_OK) return rc;
        mask = w->mask.p;
    }
    if (evl >= 1) CU(cudaEventRecord(ev.e[1], s));
    if (cudaEvent_t pe = x->prepared_event.load(std::memory_order_acquire)) CU(cudaEventRecord(pe, s));

    // ---- scan + select, MRAG_FUSED_K results per round
    const int rounds = int(ceil_div(k, MRAG_FUSED_K));
    const int64_t nwords = ceil_div(n, 32);
    // tensor-core scan: bf16 rows of <= 768 elements; single queries stay on the CUDA-core scan,
    // which already streams at ~0.9 of the HBM peak and skips masked rows individually
    // kernels:  gemv   CUDA cores, exact, <= 4 queries per pass, any dtype / dim, skips masked rows individually
    //           mma    tcgen05, exact (hi/lo split queries), 64 queries per pass, bf16 rows, dim <= 768
    //           mma128 tcgen05 candidate generation, 128 queries per pass over bf16 rows / the bf16 shadow,
    //                  + exact rescoring with certificate (k <= 32)
    const bool can_mma = x->has_tmap && x->dtype == MRAG_BF16 && n > 0;
    const bool can_mma128 = x->has_tmap && ld <= kMmaMaxLd && n > 0 && k <= kMma128MaxK;
    const bool ksplit = ld > kMmaMaxLd;                  // rows of 769 .. 1536 elements: k-split CTA pairs
    const int mma_units = ksplit ? std::max(1, x->num_sms / 2) : x->num_sms;     // CTAs (pairs) that share the tiles
    // a single query also goes to the tensor-core scan (it streams faster than the CUDA-core kernel) unless a
    // filter is active: the CUDA-core scan skips masked rows one by one, the tensor-core scan only 64-row tiles
    bool use_mma = can_mma && (nq >= 2 || !(filter && filter->flags));
    bool use_mma128 = can_mma128 && (x->dtype == MRAG_BF16 ? nq > kMmaQueries : nq >= approx_min_nq());
    if (options & MRAG_OPT_FORCE_GEMV) use_mma = use_mma128 = false;
    if (options & MRAG_OPT_FORCE_MMA) {
        if (!can_mma) return fail(MRAG_ERR_STATE, "mrag_search: the tensor-core scan needs a bf16 index with dim <= %d", kMmaKsMaxLd);
        use_mma = true; use_mma128 = false;
    }
    if (options & MRAG_OPT_FORCE_MMA128) {
        if (!can_mma128)
            return fail(MRAG_ERR_STATE, "mrag_search: the 128-query scan needs bf16 rows or a bf16 shadow, dim <= %d, k <= %d", kMmaMaxLd, kMma128MaxK);
        use_mma128 = true;
    }
    // fp32 shards that cannot (dim > 768) or should not (few queries) take the tensor-core candidate scan still
    // halve their traffic by nominating from the bf16 shadow on the CUDA cores
    const bool use_shadow_gemv = !use_mma128 && !use_mma && x->dtype == MRAG_F32 && x->shadow && n > 0 && k <= kMma128MaxK &&
                                 !(options & (MRAG_OPT_FORCE_GEMV | MRAG_OPT_FORCE_MMA | MRAG_OPT_FORCE_MMA128)) && shadow_gemv_wanted(n, ld) &&
                                 // a document pool / single document is selective: the exact scan touches few rows and needs fewer launches
                                 !(filter && (filter->flags & (MRAG_F_DOC_EQ | MRAG_F_DOC_POOL)));
    if (use_mma128 || use_shadow_gemv) {
        int rc = search_approx_rescore(x, w, ev, nq, k, mask, d_scores, d_rows, d_counts, s, use_shadow_gemv);
        if (rc != MRAG_OK) return rc;
        t_last_kind = use_mma128 ? "mma128" : "gemv_shadow";
    }
    const int grid = use_mma
        ? int(std::max<int64_t>(1, std::min<int64_t>(mma_units, ceil_div(n, kMmaTileRows)))) * (ksplit ? 2 : 1)
        : int(std::max<int64_t>(1, std::min<int64_t>(x->num_sms, ceil_div(nwords, kGemvWarps))));
    if (rounds > 1 && w->ub.reserve(size_t(nq))) return MRAG_ERR_OOM;
    bool tail_folded = false;
    // event 2 marks the end of the LAST scan; for multi-round searches the merge time of the
    // earlier rounds is attributed to the scan phase
    for (int r = 0; r < rounds && !use_mma128 && !use_shadow_gemv; ++r) {
        const int k_off = r * MRAG_FUSED_K;
        const int kr = std::min(MRAG_FUSED_K, k - k_off);
        const int kp = std::max(8, host_next_pow2(kr));
        if (w->part.reserve(size_t(nq) * grid * std::max(kp, kMmaRegK))) return MRAG_ERR_OOM;
        if (n > 0 && use_mma) {
            MmaArgs a{};
            a.n = n; a.ld = ld; a.mask = mask; a.inv_norm = x->inv_norm; a.q = w->qpad.p; a.qinv = w->qinv.p;
            a.qhl = qhl;
            a.ub = (r > 0) ? w->ub.p : nullptr;
            a.part = w->part.p; a.k = kr; a.kp = kp; a.P = grid; a.cap = kr + kMmaSlack;
            if (r > 0) CU(cudaMemsetAsync(w->gthr.p, 0, size_t(nq) * 4, s));      // (round 0: cleared by query_prep_kernel)
            a.gthr = w->gthr.p;
            const bool gmax_on = use_gmax && rounds == 1;
            if (gmax_on) { a.gmax = w->gmax.p; a.ngroups = kr; a.gslots = gslots; }
            a.tile_mul = 1;
            a.stats = nullptr;
            a.tstamps = nullptr;
            a.sleep_ns = mma_sleep_ns();
            unsigned long long* ts_sample = nullptr;
            if (getenv("MRAG_SCAN_STATS")) {
                // [0..8) counters, [8..24) stamps of the sampling launch, [24..40) stamps of the main launch
                if (w->stats.reserve(40)) return MRAG_ERR_OOM;
                CU(cudaMemsetAsync(w->stats.p, 0, 40 * 8, s));
                a.stats = w->stats.p;
                a.tstamps = w->stats.p + 24;
                ts_sample = w->stats.p + 8;
                t_stats_ptr = w->stats.p;
            }
            const int64_t tiles = ceil_div(n, kMmaTileRows);
            const bool sampled = !gmax_on && tiles >= sample_min_tiles(mma_units);
            if (sampled) {
                // always the register top-k kernel: each CTA keeps the kMmaSampleK best scores of a few tiles per query and
                // the merge takes the k-th best of the union (k known rows reach it, so it is a valid bound; a CTA holds
                // 1/#CTAs of the sample, and a list that truncates only loosens the bound)
                MmaArgs sa = a;
                sa.stats = nullptr;
                sa.tstamps = ts_sample;
                // one sampled tile per 256 tiles a CTA will scan, up to 4 (k <= 64); one per 48, up to 12, beyond: the
                // bound has to sit high enough that a CTA admits only a handful of rows per query, and the k-th best of
                // a sample sits lower the larger k is (measured r1q, 10M rows, k = 100: 3 -> 12 tiles per CTA = -0.18 ms)
                static const int per_cta_env = [] { const char* e = getenv("MRAG_SAMPLE_PER_CTA"); return (e && *e) ? atoi(e) : 0; }();
                const int per_cta = per_cta_env > 0 ? per_cta_env :
                    int(std::max<int64_t>(1, kr <= 64 ? std::min<int64_t>(4, ceil_div(tiles, 256 * int64_t(mma_units)))
                                                      : std::min<int64_t>(12, ceil_div(tiles, 48 * int64_t(mma_units)))));
                sa.tile_mul = int(std::max<int64_t>(1, tiles / (int64_t(per_cta) * mma_units)));
                const int sgrid = int(std::min<int64_t>(mma_units, ceil_div(tiles, sa.tile_mul))) * (ksplit ? 2 : 1);
                sa.P = sgrid;
                const bool short_lists = kr <= 64;              // see search_approx_rescore
                sa.k = std::min(kr, short_lists ? kMmaSampleK : kMmaRegK);
                sa.kp = short_lists ? kMmaSampleK : kMmaRegK;
                int rc = ksplit ? (short_lists ? launch_scan_mma<kMmaSampleK, true, true>(x, sa, nq, sgrid, s)
                                               : launch_scan_mma<kMmaRegK, true, true>(x, sa, nq, sgrid, s))
                                : (short_lists ? launch_scan_mma<kMmaSampleK, true>(x, sa, nq, sgrid, s)
                                               : launch_scan_mma<kMmaRegK, true>(x, sa, nq, sgrid, s));     // scores only
                if (rc != MRAG_OK) return rc;
                MergeArgs sm{};
                sm.part = w->part.p; sm.P = sgrid; sm.kp = sa.kp; sm.nq = nq; sm.k = kr; sm.lk = sa.k; sm.k_total = k; sm.k_off = k_off;
                sm.gthr_out = w->gthr.p;
                merge_kernel<<<nq, kMergeThreads, 0, s>>>(sm);
                LAUNCHED();
            }
            // unsampled shards: register top-k by default; MRAG_UNSAMPLED_BUFFER=1 pins the shared-memory buffer kernel (tuning)
            static const bool unsampled_buffer = [] { const char* e = getenv("MRAG_UNSAMPLED_BUFFER"); return e && e[0] == '1'; }();
            const bool buffers = gmax_on ? (gmax_mode == 2 && tiles >= sample_min_tiles(mma_units)) : (sampled || unsampled_buffer);
            int rc = run_scan_mma(x, a, nq, grid, /*reg_topk=*/!buffers, s);
            if (rc != MRAG_OK) return rc;
            t_last_kind = ksplit ? "mma_ks" : "mma";
        } else if (n > 0) {
            ScanArgs a{};
            a.rows = x->rows; a.n = n; a.ld = ld; a.mask = mask; a.q = w->qpad.p; a.qinv = w->qinv.p;
            a.ub = (r > 0) ? w->ub.p : nullptr;
            a.part = w->part.p; a.k = kr; a.kp = kp; a.P = grid;
            a.unit_log2 = (filter && (filter->flags & (MRAG_F_DOC_EQ | MRAG_F_DOC_POOL))) ? 2 : 0;
            int rc = run_scan_gemv(x, a, nq, grid, s);
            if (rc != MRAG_OK) return rc;
            t_last_kind = "gemv";
        } else {
            CU(cudaMemsetAsync(w->part.p, 0, size_t(nq) * grid * kp * 8, s));
        }
        if (r == rounds - 1 && evl >= 1) CU(cudaEventRecord(ev.e[2], s));
        MergeArgs m{};
        m.part = w->part.p; m.P = grid; m.kp = kp; m.nq = nq; m.k = kr; m.k_total = k; m.k_off = k_off;
        m.scores = d_scores; m.rows = d_rows; m.counts = d_counts; m.row_base = x->row_base;
        m.ub_out = (rounds > 1) ? w->ub.p : nullptr;
        m.need_tail = w->flags.p;
        // the exact tensor-core scan leaves a bound that k rows of the query reach (group maxima / compactions) in gthr
        if (n > 0 && use_mma && rounds == 1) m.thr_in = w->gthr.p;
        // the last merge of the search also appends the NaN tail of the queries that end short (no nan_tail_kernel launch)
        if (n > 0 && r == rounds - 1) {
            m.tail_mask = mask; m.tail_inv_norm = x->inv_norm; m.tail_qinv = w->qinv.p; m.tail_n = n;
            tail_folded = true;
        }
        int rcm = launch_merge(w, m, nq, s);
        if (rcm != MRAG_OK) return rcm;
    }
    // ---- NaN tail (Postgres: NaN distances sort last)
    if (n > 0 && !tail_folded) {
        TailArgs t{};
        t.inv_norm = x->inv_norm; t.mask = mask; t.n = n; t.qinv = w->qinv.p; t.nq = nq; t.k_total = k;
        t.scores = d_scores; t.rows = d_rows; t.counts = d_counts; t.row_base = x->row_base;
        t.need_tail = w->flags.p;
        nan_tail_kernel<<<nq, 256, 0, s>>>(t);
        LAUNCHED();
    }
    if (x->row_ids && n > 0) {
        remap_rows_kernel<<<unsigned(ceil_div(int64_t(nk), 256)), 256, 0, s>>>(d_rows, nk, x->row_ids, x->row_base);
        LAUNCHED();
    }
    if (!dev_io) {
        CU(cudaMemcpyAsync(scores, d_scores, nk * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(rows, d_rows, nk * 8, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(counts, d_counts, size_t(nq) * 4, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaEventRecord(ev.e[3], s));
    ev.recorded = true;
    return MRAG_OK;
}

static int search_direct(mrag_index* x, const float* q, int nq, int k, const mrag_filter* filter, float* scores,
                         int64_t* rows, int32_t* counts, uint32_t options, void* stream);

// One thread at a time is the "leader": it drains every queued request with its own k into one batch, runs it, hands
// the results back and passes leadership on.  No timer: an idle index serves a lone request immediately; under load
// the batch is whatever arrived during the previous scan.
static int search_coalesced(mrag_index* x, const float* q, int nq, int k, float* scores, int64_t* rows, int32_t* counts,
                            uint32_t options) {
    mrag_index::Pending me;
    me.q = q; me.nq = nq; me.k = k; me.options = options; me.scores = scores; me.rows = rows; me.counts = counts;
    std::unique_lock<std::mutex> l(x->co_lock);
    x->co_queue.push_back(&me);
    x->co_requests.fetch_add(1, std::memory_order_relaxed);
    for (;;) {
        if (me.done) break;
        if (x->co_busy) { x->co_cv.wait(l); continue; }
        // become the leader for one batch
        x->co_busy = true;
        std::vector<mrag_index::Pending*> batch;
        int total = 0;
        for (auto it = x->co_queue.begin(); it != x->co_queue.end();) {
            mrag_index::Pending* p = *it;
            if (p->k == me.k && p->options == me.options && total + p->nq <= 1024) { batch.push_back(p); total += p->nq; it = x->co_queue.erase(it); }
            else ++it;
        }
        l.unlock();
        int rc = MRAG_OK;
        if (batch.size() == 1) {
            mrag_index::Pending* p = batch[0];
            rc = search_direct(x, p->q, p->nq, p->k, nullptr, p->scores, p->rows, p->counts, p->options, nullptr);
            p->rc = rc; if (rc != MRAG_OK) p->err = t_err;
        } else {
            const size_t dim = size_t(x->dim);
            std::vector<float> Q(size_t(total) * dim), S(size_t(total) * me.k);
            std::vector<int64_t> R(size_t(total) * me.k);
            std::vector<int32_t> C;
            C.resize(size_t(total));
            size_t off = 0;
            for (auto* p : batch) { memcpy(Q.data() + off * dim, p->q, size_t(p->nq) * dim * 4); off += size_t(p->nq); }
            rc = search_direct(x, Q.data(), total, me.k, nullptr, S.data(), R.data(), C.data(), me.options, nullptr);
            off = 0;
            for (auto* p : batch) {
                if (rc == MRAG_OK) {
                    memcpy(p->scores, S.data() + off * me.k, size_t(p->nq) * me.k * 4);
                    memcpy(p->rows, R.data() + off * me.k, size_t(p->nq) * me.k * 8);
                    memcpy(p->counts, C.data() + off, size_t(p->nq) * 4);
                } else {
                    p->err = t_err;
                }
                p->rc = rc;
                off += size_t(p->nq);
            }
        }
        x->co_batches.fetch_add(1, std::memory_order_relaxed);
        l.lock();
        for (auto* p : batch) p->done = true;
        x->co_busy = false;
        x->co_cv.notify_all();
        // (if this thread's own request had a different k than the batch it led -- impossible: it leads its own k)
    }
    if (me.rc != MRAG_OK) t_err = me.err;
    return me.rc;
}

extern "C" int mrag_search(mrag_index* x, const float* q, int nq, int k, const mrag_filter* filter, float* scores,
                           int64_t* rows, int32_t* counts, uint32_t options, void* stream) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_search: null index");
    if ((options & MRAG_OPT_COALESCE) && !(options & MRAG_OPT_DEVICE_IO) && !stream && !(filter && filter->flags) &&
        q && scores && rows && counts && nq >= 1 && nq <= 1024 && k >= 1 && k <= MRAG_MAX_K)
        return search_coalesced(x, q, nq, k, scores, rows, counts, options & ~uint32_t(MRAG_OPT_COALESCE));
    return search_direct(x, q, nq, k, filter, scores, rows, counts, options & ~uint32_t(MRAG_OPT_COALESCE), stream);
}

// (debugging aid, not in mrag.h) coalescing counters: out2[0] = batches run, out2[1] = requests served
extern "C" int mrag_debug_coalesce_stats(mrag_index* x, int64_t* out2) {
    if (!x || !out2) return MRAG_ERR_ARG;
    out2[0] = x->co_batches.load(); out2[1] = x->co_requests.load();
    return MRAG_OK;
}

static int search_direct(mrag_index* x, const float* q, int nq, int k, const mrag_filter* filter, float* scores,
                         int64_t* rows, int32_t* counts, uint32_t options, void* stream) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_search: null index");
    if (nq < 0) return fail(MRAG_ERR_ARG, "mrag_search: nq < 0");
    if (nq == 0) return MRAG_OK;
    if (nq > 65535) return fail(MRAG_ERR_ARG, "mrag_search: nq %d > 65535", nq);
    if (k < 1 || k > MRAG_MAX_K) return fail(MRAG_ERR_ARG, "mrag_search: k %d not in [1,%d]", k, MRAG_MAX_K);
    if (!q || !scores || !rows || !counts) return fail(MRAG_ERR_ARG, "mrag_search: null buffer");
    const bool dev_io = options & MRAG_OPT_DEVICE_IO;
    const bool no_sync = (options & MRAG_OPT_NO_SYNC) && dev_io;
    if ((options & MRAG_OPT_NO_SYNC) && !dev_io)
        return fail(MRAG_ERR_ARG, "mrag_search: MRAG_OPT_NO_SYNC needs MRAG_OPT_DEVICE_IO");
    {
        const uint32_t forced = options & (MRAG_OPT_FORCE_GEMV | MRAG_OPT_FORCE_MMA | MRAG_OPT_FORCE_MMA128);
        if (forced & (forced - 1)) return fail(MRAG_ERR_ARG, "mrag_search: the FORCE_* options exclude each other");
    }

    std::shared_lock<std::shared_mutex> rl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_search: cudaSetDevice(%d) failed (no CPU path)", x->device);
    // host buffers + NULL stream: an internal per-call stream (concurrent callers overlap);
    // device buffers + NULL stream: the CUDA default (legacy) stream, as for any CUDA library
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    if (dev_io && !user) user = cudaStreamLegacy;
    Workspace* w = acquire_ws(x, user);
    if (!w) return fail(MRAG_ERR_OOM, "mrag_search: cannot create a workspace");
    cudaStream_t s = user ? user : w->own_stream;

    EventSet* ev = &w->ev;
    if (t_ring_used < int(t_ring.size()) && t_ring_device == x->device) ev = &t_ring[size_t(t_ring_used++)];
    int rc = search_locked(x, w, *ev, q, nq, k, filter, scores, rows, counts, options, s);
    if (ev != &w->ev) {
        // keep the workspace's own end marker valid for reuse decisions
        cudaEventRecord(w->ev.e[3], s);
        w->ev.recorded = true;
    }
    if (rc == MRAG_OK && !no_sync) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_search: %s", cudaGetErrorString(e));
    } else if (rc != MRAG_OK) {
        cudaStreamSynchronize(s);
        cudaGetLastError();
    }
    t_last_ev = *ev;
    t_last_valid = (rc == MRAG_OK);
    release_ws(x, w, (rc == MRAG_OK && no_sync) ? s : nullptr);
    return rc;
}


// ------------------------------------------------------------------------------------------
// hybrid rerank (K5)
// ------------------------------------------------------------------------------------------
extern "C" int mrag_set_chunk_features(mrag_index* x, int64_t first_row, const mrag_chunkfeat* feat, int64_t n) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_set_chunk_features: null index");
    if (first_row < 0 || n < 0) return fail(MRAG_ERR_ARG, "mrag_set_chunk_features: negative range");
    if (n == 0) return MRAG_OK;
    if (!feat) return fail(MRAG_ERR_ARG, "mrag_set_chunk_features: feat is null");
    std::unique_lock<std::shared_mutex> wl(x->lock);
    if (first_row + n > x->size) return fail(MRAG_ERR_ARG, "mrag_set_chunk_features: rows [%lld, %lld) beyond size %lld",
                                             (long long)first_row, (long long)(first_row + n), (long long)x->size);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_set_chunk_features: cudaSetDevice failed");
    if (!x->feat) {
        const int64_t cap32 = ceil_div(x->capacity, 32) * 32;
        CU(cudaMalloc(&x->feat, size_t(cap32) * sizeof(mrag_chunkfeat)));
        CU(cudaMemsetAsync(x->feat, 0, size_t(cap32) * sizeof(mrag_chunkfeat), x->wstream));   // all-zero = no features
    }
    CU(cudaMemcpyAsync(x->feat + first_row, feat, size_t(n) * sizeof(mrag_chunkfeat), cudaMemcpyHostToDevice, x->wstream));
    CU(cudaStreamSynchronize(x->wstream));
    return MRAG_OK;
}

extern "C" int mrag_set_dtag_overflow(mrag_index* x, const uint32_t* rows, const uint16_t* codes, int64_t n) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_set_dtag_overflow: null index");
    if (n < 0 || (n > 0 && (!rows || !codes))) return fail(MRAG_ERR_ARG, "mrag_set_dtag_overflow: bad arrays");
    for (int64_t i = 1; i < n; ++i)
        if (rows[i] < rows[i - 1]) return fail(MRAG_ERR_ARG, "mrag_set_dtag_overflow: pairs must be sorted by row");
    std::unique_lock<std::shared_mutex> wl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_set_dtag_overflow: cudaSetDevice failed");
    uint32_t* dr = nullptr;
    uint16_t* dc = nullptr;
    if (n > 0) {
        CU(cudaMalloc(&dr, size_t(n) * 4));
        cudaError_t e = cudaMalloc(&dc, size_t(n) * 2);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dr, rows, size_t(n) * 4, cudaMemcpyHostToDevice, x->wstream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dc, codes, size_t(n) * 2, cudaMemcpyHostToDevice, x->wstream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(x->wstream);
        if (e != cudaSuccess) { cudaFree(dr); if (dc) cudaFree(dc); return fail(MRAG_ERR_CUDA, "mrag_set_dtag_overflow: %s", cudaGetErrorString(e)); }
    }
    if (x->over_rows) cudaFree(x->over_rows);
    if (x->over_codes) cudaFree(x->over_codes);
    x->over_rows = dr; x->over_codes = dc; x->n_over = n;
    return MRAG_OK;
}

extern "C" int mrag_set_doc_jtags(mrag_index* x, int64_t first_doc, const uint64_t* bits, int64_t n_docs) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_set_doc_jtags: null index");
    if (first_doc < 0 || n_docs < 0) return fail(MRAG_ERR_ARG, "mrag_set_doc_jtags: negative range");
    if (n_docs == 0) return MRAG_OK;
    if (!bits) return fail(MRAG_ERR_ARG, "mrag_set_doc_jtags: bits is null");
    std::unique_lock<std::shared_mutex> wl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_set_doc_jtags: cudaSetDevice failed");
    const int64_t need = first_doc + n_docs;
    if (need > x->jtag_docs_cap) {
        int64_t cap = std::max<int64_t>(need + need / 2, 1024);
        uint64_t* p = nullptr;
        CU(cudaMalloc(&p, size_t(cap) * MRAG_JTAG_WORDS * 8));
        CU(cudaMemsetAsync(p, 0, size_t(cap) * MRAG_JTAG_WORDS * 8, x->wstream));
        if (x->doc_jtags) {
            CU(cudaMemcpyAsync(p, x->doc_jtags, size_t(x->n_jtag_docs) * MRAG_JTAG_WORDS * 8, cudaMemcpyDeviceToDevice, x->wstream));
            CU(cudaStreamSynchronize(x->wstream));
            cudaFree(x->doc_jtags);
        }
        x->doc_jtags = p;
        x->jtag_docs_cap = cap;
    }
    CU(cudaMemcpyAsync(x->doc_jtags + size_t(first_doc) * MRAG_JTAG_WORDS, bits, size_t(n_docs) * MRAG_JTAG_WORDS * 8,
                       cudaMemcpyHostToDevice, x->wstream));
    CU(cudaStreamSynchronize(x->wstream));
    x->n_jtag_docs = std::max(x->n_jtag_docs, need);
    return MRAG_OK;
}

static void derive_hyb(const mrag_hybrid_query& q, DevHyb* d) {
    memset(d, 0, sizeof *d);
    d->q = q;
    const int np = std::max(0, std::min<int>(q.n_phrases, MRAG_HYB_MAX_PHRASES));
    d->q.n_phrases = np;
    float tw = 0.0f;
    for (int i = 0; i < np; ++i) {
        tw += q.phrase_weight[i];                       // same left-to-right order as the kernel's partial sums
        if (q.phrase_dcode[i] != 0) d->has_dcodes = 1;
        if (q.phrase_weight[i] > 0.0f && q.phrase_jbit[i] < 0) {
            const int pb = q.phrase_bit[i];
            if (pb >= 0 && pb < MRAG_PHRASE_WORDS * 64) d->need[pb >> 6] |= 1ull << (pb & 63);
            else d->impossible = 1;
        }
    }
    d->total_weight = tw != 0.0f ? tw : 1.0f;           // `sum(phrase_weights) or 1.0`
    d->max_weight = q.w_sim + q.w_auth + q.w_len + q.w_jpd + q.w_cov;
    float qs = 0.0f;
    for (int c = 0; c < MRAG_JPD_CATS; ++c) qs += q.qcat[c];
    d->qcat_sum = qs;
    if (!(qs > 0.0f)) d->q.w_jpd = 0.0f;
    for (int i = 0; i < MRAG_SMALL_WORDS; ++i) if (q.source_type_any[i]) d->src_restrict = 1;
}

// HybChunk of queries [q0, q0 + 32): the transposed per-query constants of hybrid_mask_kernel's phase 1
static void fill_hyb_chunk(const DevHyb* dh, int q0, int nq, HybChunk* c) {
    memset(c, 0, sizeof *c);
    const int nqc = std::min(32, nq - q0);
    for (int i = 0; i < nqc; ++i) {
        const DevHyb& h = dh[q0 + i];
        const uint32_t bit = 1u << i;
        c->live |= bit;
        if (h.q.n_phrases > 0) c->phr |= bit;
        if (h.impossible) c->imp |= bit;
        if (h.src_restrict) c->src |= bit;
        if (h.q.contact_query) c->contact |= bit;
        if (h.has_dcodes) c->dcodes |= bit;
        if (h.q.floor < 1.0f) c->lowfloor |= bit;
        int need_bits = 0;
        for (int p = 0; p < MRAG_PHRASE_WORDS * 64; ++p)
            if ((h.need[p >> 6] >> (p & 63)) & 1ull) { c->qneed[p] |= bit; ++need_bits; }
        for (int b = 0; b < 5; ++b) if ((need_bits >> b) & 1) c->cnt[b] |= bit;
        for (int j = 0; j < h.q.n_phrases; ++j) {
            const uint16_t dc = uint16_t(h.q.phrase_dcode[j]);
            if (dc == 0 || c->ncodes == 0xFFFFFFFFu) continue;
            uint32_t s = 0;
            while (s < c->ncodes && c->code[s] != dc) ++s;
            if (s == c->ncodes) {
                if (c->ncodes == 64) { c->ncodes = 0xFFFFFFFFu; continue; }
                c->code[c->ncodes++] = dc;
            }
            c->codemask[s] |= bit;
        }
    }
}

static int hybrid_locked(mrag_index* x, Workspace* w, EventSet& ev, const float* q, int nq, int k, const mrag_filter* filter,
                         const mrag_hybrid_query* hq, float* scores, float* cos_out, int64_t* rows, int32_t* counts,
                         cudaStream_t s) {
    const int64_t n = x->size;
    const int ld = x->ld;
    const size_t nk = size_t(nq) * k;
    CU(cudaEventRecord(ev.e[0], s));
    if (w->qraw.reserve(size_t(nq) * x->dim) || w->qpad.reserve(size_t(nq) * ld) || w->qinv.reserve(size_t(nq)) ||
        w->flags.reserve(4) || w->counts.reserve(size_t(nq)) || w->scores.reserve(nk) || w->rows.reserve(nk) ||
        w->cscores.reserve(nk) || w->hyb.reserve(size_t(nq)))
        return MRAG_ERR_OOM;
    CU(cudaMemcpyAsync(w->qraw.p, q, size_t(nq) * x->dim * 4, cudaMemcpyHostToDevice, s));
    query_prep_kernel<<<unsigned(ceil_div(int64_t(nq) * 32, 128)), 128, 0, s>>>(w->qraw.p, nq, x->dim, ld, w->qpad.p, w->qinv.p, nullptr, nq, nullptr);
    LAUNCHED();
    std::vector<DevHyb> dh;
    dh.resize(size_t(nq));
    for (int i = 0; i < nq; ++i) derive_hyb(hq[i], &dh[size_t(i)]);
    CU(cudaMemcpyAsync(w->hyb.p, dh.data(), size_t(nq) * sizeof(DevHyb), cudaMemcpyHostToDevice, s));
    std::vector<HybChunk> hc;
    hc.resize(size_t(ceil_div(nq, 32)));
    for (size_t c = 0; c < hc.size(); ++c) fill_hyb_chunk(dh.data(), int(c) * 32, nq, &hc[c]);
    if (w->hchunks.reserve(hc.size())) return MRAG_ERR_OOM;
    CU(cudaMemcpyAsync(w->hchunks.p, hc.data(), hc.size() * sizeof(HybChunk), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));                      // dh / hc are stack-lifetime host buffers
    const int kp = std::max(8, host_next_pow2(k));
    const int64_t nwords = ceil_div(n, 32);
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>(x->num_sms, ceil_div(nwords, kGemvWarps))));
    if (w->part.reserve(size_t(nq) * grid * kp)) return MRAG_ERR_OOM;
    bool pair_path = false;
    if (n > 0) {
        const uint32_t* mask = x->cols.valid;
        if (filter && filter->flags) {
            if (w->mask.reserve(size_t(nwords) + 1)) return MRAG_ERR_OOM;
            int rc = build_mask(x, w, filter, n, w->mask.p, nullptr, s);
            if (rc != MRAG_OK) return rc;
            mask = w->mask.p;
        }
        if (w->hmask.reserve(size_t(nq) * nwords)) return MRAG_ERR_OOM;
        const DtagOver ov{x->over_rows, x->over_codes, x->n_over};
        hybrid_mask_kernel<<<unsigned(ceil_div(nwords * 32, 256)), 256, 0, s>>>(w->hyb.p, w->hchunks.p, nq, x->feat, mask, x->cols.doc_idx, x->cols.source_type, x->doc_jtags,
                                                                              x->n_jtag_docs, n, w->hmask.p, nwords, ov);
        LAUNCHED();
        CU(cudaEventRecord(ev.e[1], s));
        // ---- pair path: the floors usually leave a query a small share of the rows; then the rerank runs over the list of
        //      surviving (query, row) pairs (one warp per pair) instead of 4 queries per pass over every mask word.
        //      Taken when EVERY query of the batch keeps <= 2^20 rows (a query without required phrases keeps them all and
        //      sends the batch down the scan below).  MRAG_HYB_PAIRS=0: always scan.
        static const bool pairs_ok = [] { const char* e = getenv("MRAG_HYB_PAIRS"); return !(e && e[0] == '0'); }();
        if (pairs_ok) {
            if (w->npass.reserve(size_t(2) * nq) || w->segoff.reserve(size_t(nq) + 1)) return MRAG_ERR_OOM;
            CU(cudaMemsetAsync(w->npass.p, 0, size_t(2) * nq * 8, s));
            hybrid_count_kernel<<<dim3(unsigned(std::min<int64_t>(512, ceil_div(nwords, 256))), unsigned(nq)), 256, 0, s>>>(w->hmask.p, nwords, w->npass.p);
            LAUNCHED();
            std::vector<unsigned long long> cnt;
            cnt.resize(size_t(nq));
            CU(cudaMemcpyAsync(cnt.data(), w->npass.p, size_t(nq) * 8, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            std::vector<int64_t> seg;
            seg.resize(size_t(nq) + 1);
            unsigned long long maxc = 0;
            seg[0] = 0;
            for (int i = 0; i < nq; ++i) { seg[size_t(i) + 1] = seg[size_t(i)] + int64_t(cnt[size_t(i)]); maxc = std::max(maxc, cnt[size_t(i)]); }
            const int64_t total = seg[size_t(nq)];
            const char* pm_env = getenv("MRAG_HYB_PAIR_MAX");              // read per call: the tests switch between the two paths
            const unsigned long long pair_max = (pm_env && *pm_env) ? strtoull(pm_env, nullptr, 10) : (1ull << 20);
            if (maxc <= pair_max && pair_max > 0) {
                if (w->pair_rows.reserve(size_t(std::max<int64_t>(total, 1))) || w->part.reserve(size_t(std::max<int64_t>(total, 1)))) return MRAG_ERR_OOM;
                CU(cudaMemcpyAsync(w->segoff.p, seg.data(), (size_t(nq) + 1) * 8, cudaMemcpyHostToDevice, s));
                CU(cudaStreamSynchronize(s));              // seg is a stack-lifetime host buffer
                if (total > 0) {
                    hybrid_fill_kernel<<<dim3(unsigned(std::min<int64_t>(512, ceil_div(nwords, 256))), unsigned(nq)), 256, 0, s>>>(
                        w->hmask.p, nwords, w->segoff.p, w->npass.p + nq, w->pair_rows.p);
                    LAUNCHED();
                    PairArgs pa{};
                    pa.rows = x->rows; pa.ld = ld; pa.inv_norm = x->inv_norm; pa.q = w->qpad.p; pa.qinv = w->qinv.p;
                    pa.pair_rows = w->pair_rows.p; pa.seg_off = w->segoff.p; pa.nq = nq; pa.total = total;
                    pa.feat = x->feat; pa.hyb = w->hyb.p; pa.doc_idx = x->cols.doc_idx; pa.authority = x->cols.authority;
                    pa.doc_jtags = x->doc_jtags; pa.n_jtag_docs = x->n_jtag_docs; pa.ov = ov; pa.keys = w->part.p;
                    const unsigned pb = unsigned(ceil_div(ceil_div(total, 32) * 32, 256));      // one warp per 32 pairs
                    if (x->dtype == MRAG_BF16) hybrid_pair_score_kernel<1><<<pb, 256, 0, s>>>(pa);
                    else hybrid_pair_score_kernel<0><<<pb, 256, 0, s>>>(pa);
                    LAUNCHED();
                }
                pair_path = true;
            }
        }
        ScanArgs a{};
        a.rows = x->rows; a.n = n; a.ld = ld; a.mask = mask; a.q = w->qpad.p; a.qinv = w->qinv.p; a.ub = nullptr;
        a.part = w->part.p; a.k = k; a.kp = kp; a.P = grid;
        a.hmask = w->hmask.p; a.hwords = nwords; a.feat = x->feat; a.hyb = w->hyb.p; a.doc_idx = x->cols.doc_idx;
        a.dtag_over = ov;
        a.authority = x->cols.authority; a.doc_jtags = x->doc_jtags; a.n_jtag_docs = x->n_jtag_docs;
        int q0 = pair_path ? nq : 0;
        while (q0 < nq) {
            const int left = nq - q0;
            const int g = gemv_nq_for(left >= 3 ? 4 : left, ld, kp, true);
            a.q0 = q0; a.nq = std::min(g, left);
            a.hyb = w->hyb.p;
            int rc;
            if (x->dtype == MRAG_BF16)
                rc = g == 4 ? launch_gemv<1, 4, 1>(a, grid, s) : g == 2 ? launch_gemv<1, 2, 1>(a, grid, s) : launch_gemv<1, 1, 1>(a, grid, s);
            else
                rc = g == 4 ? launch_gemv<0, 4, 1>(a, grid, s) : g == 2 ? launch_gemv<0, 2, 1>(a, grid, s) : launch_gemv<0, 1, 1>(a, grid, s);
            if (rc != MRAG_OK) return rc;
            q0 += a.nq;
        }
    } else {
        CU(cudaMemsetAsync(w->part.p, 0, size_t(nq) * grid * kp * 8, s));
        CU(cudaEventRecord(ev.e[1], s));
    }
    CU(cudaEventRecord(ev.e[2], s));
    MergeArgs m{};
    m.part = w->part.p; m.P = grid; m.kp = kp; m.nq = nq; m.k = k; m.k_total = k; m.k_off = 0;
    m.scores = w->scores.p; m.rows = w->rows.p; m.counts = w->counts.p; m.row_base = x->row_base; m.no_clamp = 1;
    t_hybrid_pairs = pair_path;
    if (pair_path) {
        // two levels: 8 blocks per query each keep the best k of a slice of the query's segment, then the usual merge of 8 lists
        // (one block per query left 126 SMs idle: 22 queries x ~34k keys took 0.85 ms)
        const int slices = 8;
        if (w->part2.reserve(size_t(nq) * slices * kp)) return MRAG_ERR_OOM;
        MergeArgs m1 = m;
        m1.seg_off = w->segoff.p; m1.P = 1; m1.kp = kp; m1.part_out = w->part2.p; m1.lk = 1;
        merge_kernel<<<dim3(unsigned(nq), unsigned(slices)), kMergeThreads, 0, s>>>(m1);
        LAUNCHED();
        m.part = w->part2.p; m.P = slices; m.kp = kp;
    }
    int rc = launch_merge(w, m, nq, s);
    if (rc != MRAG_OK) return rc;
    if (n > 0) {
        CosRowsArgs ca{};
        ca.rows = x->rows; ca.ld = ld; ca.inv_norm = x->inv_norm; ca.q = w->qpad.p; ca.qinv = w->qinv.p;
        ca.sel_rows = w->rows.p; ca.sel_counts = w->counts.p; ca.row_base = x->row_base; ca.nq = nq; ca.k = k; ca.cos_out = w->cscores.p;
        const unsigned cb = unsigned(ceil_div(int64_t(nq) * k * 32, 256));
        if (x->dtype == MRAG_BF16) cos_rows_kernel<1><<<cb, 256, 0, s>>>(ca);
        else cos_rows_kernel<0><<<cb, 256, 0, s>>>(ca);
        LAUNCHED();
        if (x->row_ids) {
            remap_rows_kernel<<<unsigned(ceil_div(int64_t(nk), 256)), 256, 0, s>>>(w->rows.p, nk, x->row_ids, x->row_base);
            LAUNCHED();
        }
    }
    CU(cudaMemcpyAsync(scores, w->scores.p, nk * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(rows, w->rows.p, nk * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(counts, w->counts.p, size_t(nq) * 4, cudaMemcpyDeviceToHost, s));
    if (cos_out && n > 0) CU(cudaMemcpyAsync(cos_out, w->cscores.p, nk * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(ev.e[3], s));
    ev.recorded = true;
    return MRAG_OK;
}

extern "C" int mrag_search_hybrid(mrag_index* x, const float* q, int nq, int k, const mrag_filter* filter,
                                  const mrag_hybrid_query* hq, float* scores, float* cos_out, int64_t* rows,
                                  int32_t* counts, void* stream) {
    if (!x) return fail(MRAG_ERR_ARG, "mrag_search_hybrid: null index");
    if (nq < 0) return fail(MRAG_ERR_ARG, "mrag_search_hybrid: nq < 0");
    if (nq == 0) return MRAG_OK;
    if (nq > 65535) return fail(MRAG_ERR_ARG, "mrag_search_hybrid: nq %d > 65535", nq);
    if (k < 1 || k > MRAG_FUSED_K) return fail(MRAG_ERR_ARG, "mrag_search_hybrid: k %d not in [1,%d]", k, MRAG_FUSED_K);
    if (!q || !hq || !scores || !rows || !counts) return fail(MRAG_ERR_ARG, "mrag_search_hybrid: null buffer");
    std::shared_lock<std::shared_mutex> rl(x->lock);
    if (!x->feat && x->size > 0) return fail(MRAG_ERR_STATE, "mrag_search_hybrid: no chunk features set (mrag_set_chunk_features)");
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_search_hybrid: cudaSetDevice(%d) failed (no CPU path)", x->device);
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    Workspace* w = acquire_ws(x, user);
    if (!w) return fail(MRAG_ERR_OOM, "mrag_search_hybrid: cannot create a workspace");
    cudaStream_t s = user ? user : w->own_stream;
    EventSet* ev = &w->ev;
    if (t_ring_used < int(t_ring.size()) && t_ring_device == x->device) ev = &t_ring[size_t(t_ring_used++)];
    int rc = hybrid_locked(x, w, *ev, q, nq, k, filter, hq, scores, cos_out, rows, counts, s);
    if (ev != &w->ev) { cudaEventRecord(w->ev.e[3], s); w->ev.recorded = true; }
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc == MRAG_OK && e != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_search_hybrid: %s", cudaGetErrorString(e));
    if (rc != MRAG_OK) cudaGetLastError();
    if (rc == MRAG_OK && x->size == 0 && cos_out)
        for (size_t i = 0; i < size_t(nq) * k; ++i) cos_out[i] = NAN;
    t_last_ev = *ev;
    t_last_valid = (rc == MRAG_OK);
    t_last_kind = t_hybrid_pairs ? "pairs_hybrid" : "gemv_hybrid";
    release_ws(x, w, nullptr);
    return rc;
}


extern "C" int mrag_rerank_candidates(mrag_index* x, const mrag_candidate* cands, int64_t n, const mrag_hybrid_query* hq,
                                      float* scores, float* coverage, uint8_t* keep) {
    if (!x || !hq || n < 0 || (n > 0 && (!cands || !scores || !coverage || !keep)))
        return fail(MRAG_ERR_ARG, "mrag_rerank_candidates: null argument");
    if (n == 0) return MRAG_OK;
    std::shared_lock<std::shared_mutex> rl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_rerank_candidates: cudaSetDevice failed (no CPU path)");
    Workspace* w = acquire_ws(x, nullptr);
    if (!w) return fail(MRAG_ERR_OOM, "mrag_rerank_candidates: cannot create a workspace");
    cudaStream_t s = w->own_stream;
    int rc = MRAG_OK;
    DevHyb dh;
    derive_hyb(*hq, &dh);
    const size_t in_bytes = size_t(n) * sizeof(mrag_candidate);
    // one staging buffer: candidates | scores | coverage | keep
    const size_t off_s = (in_bytes + 15) / 16 * 16, off_c = off_s + size_t(n) * 4, off_k = off_c + size_t(n) * 4, total = off_k + size_t(n);
    if (w->hyb.reserve(1) || w->ckeys.reserve((total + 7) / 8)) rc = MRAG_ERR_OOM;
    if (rc == MRAG_OK) {
        unsigned char* base = reinterpret_cast<unsigned char*>(w->ckeys.p);
        cudaMemcpyAsync(w->hyb.p, &dh, sizeof dh, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(base, cands, in_bytes, cudaMemcpyHostToDevice, s);
        rerank_candidates_kernel<<<unsigned(ceil_div(n, 128)), 128, 0, s>>>(reinterpret_cast<const mrag_candidate*>(base), n, w->hyb.p, x->doc_jtags,
                                                                           x->n_jtag_docs, reinterpret_cast<float*>(base + off_s),
                                                                           reinterpret_cast<float*>(base + off_c), base + off_k);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaMemcpyAsync(scores, base + off_s, size_t(n) * 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(coverage, base + off_c, size_t(n) * 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(keep, base + off_k, size_t(n), cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_rerank_candidates: %s", cudaGetErrorString(e));
    }
    release_ws(x, w, nullptr);
    return rc;
}

// The d-tag arm (corpus_search.py:1605-1701): WHERE over live rows + "chunk_d_tags ? key" for any key.  The row
// bitmap comes back to the HOST (the shim orders the matches by (authority tier, id) and cuts to k, :1674-1680).
extern "C" int mrag_dtag_mask(mrag_index* x, const mrag_filter* filter, const uint16_t* dcodes, int n_codes,
                              uint32_t* host_mask_out, int64_t* counts) {
    if (!x || !host_mask_out || !counts) return fail(MRAG_ERR_ARG, "mrag_dtag_mask: null argument");
    if (n_codes < 0 || n_codes > 32 || (n_codes > 0 && !dcodes)) return fail(MRAG_ERR_ARG, "mrag_dtag_mask: 0 <= n_codes <= 32");
    for (int i = 0; i <= n_codes; ++i) counts[i] = 0;
    std::shared_lock<std::shared_mutex> rl(x->lock);
    const int64_t n = x->size;
    if (n == 0) return MRAG_OK;
    if (!x->feat) return fail(MRAG_ERR_STATE, "mrag_dtag_mask: no chunk features set (mrag_set_chunk_features)");
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_dtag_mask: cudaSetDevice failed (no CPU path)");
    Workspace* w = acquire_ws(x, nullptr);
    if (!w) return fail(MRAG_ERR_OOM, "mrag_dtag_mask: cannot create a workspace");
    cudaStream_t s = w->own_stream;
    const int64_t nwords = ceil_div(n, 32);
    int rc = MRAG_OK;
    mrag_filter none;
    memset(&none, 0, sizeof none);
    if (w->mask.reserve(size_t(nwords) + 1) || w->hmask.reserve(size_t(nwords) + 1) || w->stats.reserve(40) || w->codes.reserve(32))
        rc = MRAG_ERR_OOM;
    if (rc == MRAG_OK) rc = build_mask(x, w, filter ? filter : &none, n, w->mask.p, nullptr, s, /*include_null_vec=*/true);
    if (rc == MRAG_OK) {
        unsigned long long hc[40];
        cudaMemsetAsync(w->stats.p, 0, 40 * 8, s);
        if (n_codes) cudaMemcpyAsync(w->codes.p, dcodes, size_t(n_codes) * 2, cudaMemcpyHostToDevice, s);
        dtag_mask_kernel<<<unsigned(ceil_div(nwords * 32, 256)), 256, 0, s>>>(x->feat, w->mask.p, n, w->codes.p, n_codes, w->hmask.p, w->stats.p,
                                                                             DtagOver{x->over_rows, x->over_codes, x->n_over});
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaMemcpyAsync(host_mask_out, w->hmask.p, size_t(nwords) * 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(hc, w->stats.p, size_t(n_codes + 1) * 8, cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_dtag_mask: %s", cudaGetErrorString(e));
        else for (int i = 0; i <= n_codes; ++i) counts[i] = int64_t(hc[i]);
    } else {
        cudaStreamSynchronize(s);
    }
    release_ws(x, w, nullptr);
    return rc;
}


// ------------------------------------------------------------------------------------------
// snapshot: one shard <-> one file (the GPU-resident copy is re-creatable from the table and keyed
// by corpus_state.corpus_version, publish.py:314; a snapshot makes the cold start one sequential read)
// ------------------------------------------------------------------------------------------
namespace {
struct SnapHeader {
    char magic[8];            // "MRAGSNP1"
    int32_t dim, ld, dtype, reserved;
    int64_t size, n_docs, n_tag_docs, n_jtag_docs, has_feat, row_base, user_version;
};

struct SnapIO {
    FILE* f = nullptr;
    void* pinned = nullptr;
    size_t chunk = size_t(64) << 20;
    cudaStream_t s = nullptr;
    ~SnapIO() { if (f) fclose(f); if (pinned) cudaFreeHost(pinned); }
    int init(const char* path, const char* mode, cudaStream_t stream) {
        f = fopen(path, mode);
        if (!f) return fail(MRAG_ERR_ARG, "snapshot: cannot open %s", path);
        if (cudaMallocHost(&pinned, chunk) != cudaSuccess) return fail(MRAG_ERR_OOM, "snapshot: pinned staging buffer");
        s = stream;
        return MRAG_OK;
    }
    int put(const void* dev, size_t bytes) {                 // device -> file
        const char* p = static_cast<const char*>(dev);
        for (size_t off = 0; off < bytes; off += chunk) {
            const size_t m = std::min(chunk, bytes - off);
            if (cudaMemcpyAsync(pinned, p + off, m, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
                return fail(MRAG_ERR_CUDA, "snapshot: D2H failed");
            if (fwrite(pinned, 1, m, f) != m) return fail(MRAG_ERR_ARG, "snapshot: short write");
        }
        return MRAG_OK;
    }
    int get(void* dev, size_t bytes) {                       // file -> device
        char* p = static_cast<char*>(dev);
        for (size_t off = 0; off < bytes; off += chunk) {
            const size_t m = std::min(chunk, bytes - off);
            if (fread(pinned, 1, m, f) != m) return fail(MRAG_ERR_ARG, "snapshot: short read (truncated file?)");
            if (cudaMemcpyAsync(p + off, pinned, m, cudaMemcpyHostToDevice, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
                return fail(MRAG_ERR_CUDA, "snapshot: H2D failed");
        }
        return MRAG_OK;
    }
};
}  // namespace

extern "C" int mrag_save(mrag_index* x, const char* path, int64_t user_version) {
    if (!x || !path) return fail(MRAG_ERR_ARG, "mrag_save: null argument");
    std::shared_lock<std::shared_mutex> rl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_save: cudaSetDevice failed (no CPU path)");
    SnapIO io;
    int rc = io.init(path, "wb", x->wstream);
    if (rc != MRAG_OK) return rc;
    SnapHeader h{};
    memcpy(h.magic, "MRAGSNP1", 8);
    h.dim = x->dim; h.ld = x->ld; h.dtype = x->dtype; h.size = x->size; h.n_docs = x->n_docs; h.n_tag_docs = x->n_tag_docs;
    h.n_jtag_docs = x->n_jtag_docs; h.has_feat = x->feat ? 1 : 0; h.row_base = x->row_base; h.user_version = user_version;
    h.reserved = x->row_ids ? 1 : 0;                       // explicit row ids follow the features
    if (fwrite(&h, sizeof h, 1, io.f) != 1) return fail(MRAG_ERR_ARG, "mrag_save: short write");
    const size_t n = size_t(x->size), words = (n + 31) / 32;
    if ((rc = io.put(x->rows, n * x->ld * elem_size(x->dtype))) || (rc = io.put(x->inv_norm, n * 4)) ||
        (rc = io.put(x->cols.doc_idx, n * 4)) || (rc = io.put(x->cols.payer, n * 2)) || (rc = io.put(x->cols.state, n)) ||
        (rc = io.put(x->cols.program, n)) || (rc = io.put(x->cols.authority, n)) || (rc = io.put(x->cols.source_type, n)) ||
        (rc = io.put(x->cols.valid, words * 4)) || (rc = io.put(x->cols.live, words * 4)))
        return rc;
    if (x->n_tag_docs && (rc = io.put(x->doc_tags, size_t(x->n_tag_docs) * MRAG_TAG_WORDS * 8))) return rc;
    if (x->n_jtag_docs && (rc = io.put(x->doc_jtags, size_t(x->n_jtag_docs) * MRAG_JTAG_WORDS * 8))) return rc;
    if (x->feat && (rc = io.put(x->feat, n * sizeof(mrag_chunkfeat)))) return rc;
    if (x->row_ids && (rc = io.put(x->row_ids, n * 8))) return rc;
    return MRAG_OK;
}

__global__ void shadow_from_rows_kernel(const float* __restrict__ rows, __nv_bfloat16* __restrict__ shadow, size_t count) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) shadow[i] = __float2bfloat16_rn(rows[i]);
}

extern "C" int mrag_load(mrag_index** out, const char* path, int device, int64_t capacity, int64_t* user_version) {
    if (!out || !path) return fail(MRAG_ERR_ARG, "mrag_load: null argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(MRAG_ERR_ARG, "mrag_load: cannot open %s", path);
    SnapHeader h{};
    const bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "MRAGSNP1", 8) == 0;
    fclose(f);
    if (!ok) return fail(MRAG_ERR_ARG, "mrag_load: %s is not an mrag snapshot", path);
    if (capacity <= 0) capacity = std::max<int64_t>(h.size, 1);
    if (capacity < h.size) return fail(MRAG_ERR_ARG, "mrag_load: capacity %lld < %lld rows in the snapshot", (long long)capacity, (long long)h.size);
    mrag_index* x = nullptr;
    int rc = mrag_create(&x, h.dim, h.dtype, device, capacity);
    if (rc != MRAG_OK) return rc;
    DeviceGuard g(device);
    SnapIO io;
    rc = io.init(path, "rb", x->wstream);
    if (rc == MRAG_OK && fseek(io.f, long(sizeof h), SEEK_SET) != 0) rc = fail(MRAG_ERR_ARG, "mrag_load: seek failed");
    const size_t n = size_t(h.size), words = (n + 31) / 32;
    if (rc == MRAG_OK) {
        (rc = io.get(x->rows, n * h.ld * elem_size(h.dtype))) || (rc = io.get(x->inv_norm, n * 4)) ||
        (rc = io.get(x->cols.doc_idx, n * 4)) || (rc = io.get(x->cols.payer, n * 2)) || (rc = io.get(x->cols.state, n)) ||
        (rc = io.get(x->cols.program, n)) || (rc = io.get(x->cols.authority, n)) || (rc = io.get(x->cols.source_type, n)) ||
        (rc = io.get(x->cols.valid, words * 4)) || (rc = io.get(x->cols.live, words * 4));
    }
    if (rc == MRAG_OK && h.n_tag_docs) {
        x->tag_docs_cap = h.n_tag_docs;
        if (cudaMalloc(&x->doc_tags, size_t(h.n_tag_docs) * MRAG_TAG_WORDS * 8) != cudaSuccess) rc = fail(MRAG_ERR_OOM, "mrag_load: doc tags");
        else rc = io.get(x->doc_tags, size_t(h.n_tag_docs) * MRAG_TAG_WORDS * 8);
    }
    if (rc == MRAG_OK && h.n_jtag_docs) {
        x->jtag_docs_cap = h.n_jtag_docs;
        if (cudaMalloc(&x->doc_jtags, size_t(h.n_jtag_docs) * MRAG_JTAG_WORDS * 8) != cudaSuccess) rc = fail(MRAG_ERR_OOM, "mrag_load: doc j-tags");
        else rc = io.get(x->doc_jtags, size_t(h.n_jtag_docs) * MRAG_JTAG_WORDS * 8);
    }
    if (rc == MRAG_OK && h.has_feat) {
        const int64_t cap32 = ceil_div(x->capacity, 32) * 32;
        if (cudaMalloc(&x->feat, size_t(cap32) * sizeof(mrag_chunkfeat)) != cudaSuccess) rc = fail(MRAG_ERR_OOM, "mrag_load: features");
        else {
            cudaMemsetAsync(x->feat, 0, size_t(cap32) * sizeof(mrag_chunkfeat), x->wstream);
            rc = io.get(x->feat, n * sizeof(mrag_chunkfeat));
        }
    }
    if (rc == MRAG_OK && h.reserved == 1) {
        std::vector<int64_t> ident(size_t(x->capacity));
        for (int64_t i = 0; i < x->capacity; ++i) ident[size_t(i)] = i + h.row_base;
        if (cudaMalloc(&x->row_ids, size_t(x->capacity) * 8) != cudaSuccess) rc = fail(MRAG_ERR_OOM, "mrag_load: row ids");
        else if (cudaMemcpy(x->row_ids, ident.data(), size_t(x->capacity) * 8, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_load: row ids");
        else rc = io.get(x->row_ids, n * 8);
    }
    if (rc == MRAG_OK && x->shadow && n) {                   // the bf16 shadow is derived data: rebuilt, not stored
        const size_t count = n * size_t(h.ld);
        shadow_from_rows_kernel<<<unsigned((count + 255) / 256), 256, 0, x->wstream>>>(static_cast<const float*>(x->rows), x->shadow, count);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaStreamSynchronize(x->wstream) != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_load: shadow rebuild failed");
    }
    if (rc != MRAG_OK) {
        std::string keep = t_err;
        mrag_destroy(x);
        t_err = keep;
        return rc;
    }
    x->size = h.size; x->n_docs = h.n_docs; x->n_tag_docs = h.n_tag_docs; x->n_jtag_docs = h.n_jtag_docs; x->row_base = h.row_base;
    if (user_version) *user_version = h.user_version;
    *out = x;
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// candidate pool on the device (build_candidate_pool, corpus_search_agent.py:1762-1888)
// ------------------------------------------------------------------------------------------
extern "C" int mrag_pool_build(mrag_index* x, const mrag_pool_query* q, mrag_pool** out, int64_t counts[MRAG_POOL_LEVELS + 1]) {
    if (!x || !q || !out) return fail(MRAG_ERR_ARG, "mrag_pool_build: null argument");
    *out = nullptr;
    std::shared_lock<std::shared_mutex> rl(x->lock);
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_pool_build: cudaSetDevice failed (no CPU path)");
    mrag_pool* p = new (std::nothrow) mrag_pool();
    if (!p) return fail(MRAG_ERR_OOM, "mrag_pool_build: host allocation failed");
    p->idx = x;
    p->n_docs = std::max<int64_t>(std::max<int64_t>(x->n_docs, 1), std::max(x->n_tag_docs, x->n_jtag_docs));   // documents may have tags before rows
    p->words = ceil_div(p->n_docs, 32);
    Workspace* w = acquire_ws(x, nullptr);
    if (!w) { delete p; return fail(MRAG_ERR_OOM, "mrag_pool_build: cannot create a workspace"); }
    cudaStream_t s = w->own_stream;
    int rc = MRAG_OK;
    unsigned long long hc[MRAG_POOL_LEVELS + 1] = {0, 0, 0, 0, 0};
    if (cudaMalloc(&p->bits, size_t(MRAG_POOL_LEVELS) * p->words * 4) != cudaSuccess) rc = fail(MRAG_ERR_OOM, "mrag_pool_build: bitmap allocation failed");
    if (rc == MRAG_OK && w->stats.reserve(40)) rc = MRAG_ERR_OOM;
    if (rc == MRAG_OK) {
        DevPoolQuery dq;
        memset(&dq, 0, sizeof dq);
        memcpy(dq.d_all, q->d_all, sizeof dq.d_all); memcpy(dq.p_all, q->p_all, sizeof dq.p_all);
        memcpy(dq.j_all, q->j_all, sizeof dq.j_all); memcpy(dq.ahca, q->ahca, sizeof dq.ahca);
        dq.has_j = q->has_j; dq.has_d = q->has_d; dq.has_p = q->has_p; dq.has_ahca = q->has_ahca;
        cudaMemsetAsync(w->stats.p, 0, 40 * 8, s);
        pool_cascade_kernel<<<unsigned(ceil_div(p->words * 32, 256)), 256, 0, s>>>(dq, x->doc_tags, x->n_tag_docs, x->doc_jtags, x->n_jtag_docs,
                                                                                  p->n_docs, p->words, p->bits, w->stats.p);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaMemcpyAsync(hc, w->stats.p, sizeof hc, cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_pool_build: %s", cudaGetErrorString(e));
    }
    release_ws(x, w, nullptr);
    if (rc != MRAG_OK) { if (p->bits) cudaFree(p->bits); delete p; return rc; }
    for (int l = 0; l < MRAG_POOL_LEVELS; ++l) { p->counts[l] = int64_t(hc[l]); if (counts) counts[l] = p->counts[l]; }
    if (counts) counts[MRAG_POOL_LEVELS] = int64_t(hc[MRAG_POOL_LEVELS]);
    *out = p;
    return MRAG_OK;
}

extern "C" int mrag_pool_select(mrag_pool* p, int level, int64_t cap, int64_t* n_kept) {
    if (!p || !p->idx) return fail(MRAG_ERR_ARG, "mrag_pool_select: null pool");
    if (level < 0 || level >= MRAG_POOL_LEVELS) return fail(MRAG_ERR_ARG, "mrag_pool_select: level %d not in [0,%d)", level, MRAG_POOL_LEVELS);
    DeviceGuard g(p->idx->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_pool_select: cudaSetDevice failed");
    p->level = level;
    int64_t kept = p->counts[level];
    if (cap >= 0 && kept > cap) {
        unsigned long long* d_kept = nullptr;
        unsigned long long h = 0;
        CU(cudaMalloc(&d_kept, 8));
        pool_cap_kernel<<<1, 1024, 0, p->idx->wstream>>>(p->bits + size_t(level) * p->words, p->words, cap, d_kept);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaMemcpyAsync(&h, d_kept, 8, cudaMemcpyDeviceToHost, p->idx->wstream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->idx->wstream);
        cudaFree(d_kept);
        if (e != cudaSuccess) return fail(MRAG_ERR_CUDA, "mrag_pool_select: %s", cudaGetErrorString(e));
        kept = int64_t(h);
        p->counts[level] = kept;
    }
    if (n_kept) *n_kept = kept;
    return MRAG_OK;
}

extern "C" int mrag_pool_add_docs(mrag_pool* p, const uint32_t* docs, int64_t n) {
    if (!p || !p->idx || p->level < 0) return fail(MRAG_ERR_ARG, "mrag_pool_add_docs: null pool or no level selected");
    if (n < 0 || (n > 0 && !docs)) return fail(MRAG_ERR_ARG, "mrag_pool_add_docs: bad list");
    if (n == 0) return MRAG_OK;
    DeviceGuard g(p->idx->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_pool_add_docs: cudaSetDevice failed");
    uint32_t* d = nullptr;
    CU(cudaMalloc(&d, size_t(n) * 4));
    cudaStream_t s = p->idx->wstream;
    cudaError_t e = cudaMemcpyAsync(d, docs, size_t(n) * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        pool_bitmap_kernel<<<unsigned(ceil_div(n, 256)), 256, 0, s>>>(d, n, p->bits + size_t(p->level) * p->words, p->n_docs);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaStreamSynchronize(s);
    }
    cudaFree(d);
    if (e != cudaSuccess) return fail(MRAG_ERR_CUDA, "mrag_pool_add_docs: %s", cudaGetErrorString(e));
    return MRAG_OK;
}

extern "C" int mrag_pool_docs(mrag_pool* p, uint32_t* out, int64_t max, int64_t* n) {
    if (!p || !p->idx || p->level < 0 || !n) return fail(MRAG_ERR_ARG, "mrag_pool_docs: null pool / no level selected / null count");
    *n = 0;
    if (max < 0 || (max > 0 && !out)) return fail(MRAG_ERR_ARG, "mrag_pool_docs: bad buffer");
    DeviceGuard g(p->idx->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_pool_docs: cudaSetDevice failed");
    std::vector<uint32_t> h(size_t(p->words));
    CU(cudaMemcpy(h.data(), p->bits + size_t(p->level) * p->words, size_t(p->words) * 4, cudaMemcpyDeviceToHost));
    int64_t k = 0;
    for (int64_t w = 0; w < p->words && k < max; ++w) {
        uint32_t x = h[size_t(w)];
        while (x && k < max) {
            const int b = __builtin_ctz(x);
            x &= x - 1;
            out[k++] = uint32_t(w * 32 + b);
        }
    }
    *n = k;
    return MRAG_OK;
}

extern "C" int mrag_pool_destroy(mrag_pool* p) {
    if (!p) return MRAG_OK;
    if (p->idx) {
        DeviceGuard g(p->idx->device);
        if (p->bits) cudaFree(p->bits);
    }
    delete p;
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// K2 alone, K4
// ------------------------------------------------------------------------------------------
extern "C" int mrag_filter_mask(mrag_index* x, const mrag_filter* filter, uint32_t* d_mask_out, int64_t* n_pass,
                                void* stream) {
    if (!x || !d_mask_out) return fail(MRAG_ERR_ARG, "mrag_filter_mask: null argument");
    if (n_pass) *n_pass = 0;
    std::shared_lock<std::shared_mutex> rl(x->lock);
    const int64_t n = x->size;
    if (n == 0) return MRAG_OK;
    DeviceGuard g(x->device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_filter_mask: cudaSetDevice failed");
    Workspace* w = acquire_ws(x, static_cast<cudaStream_t>(stream));
    if (!w) return fail(MRAG_ERR_OOM, "mrag_filter_mask: cannot create a workspace");
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : w->own_stream;
    mrag_filter none;
    memset(&none, 0, sizeof none);
    int rc = MRAG_OK;
    if (w->npass.reserve(1)) rc = MRAG_ERR_OOM;
    if (rc == MRAG_OK && cudaMemsetAsync(w->npass.p, 0, 8, s) != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "memset failed");
    if (rc == MRAG_OK) rc = build_mask(x, w, filter ? filter : &none, n, d_mask_out, w->npass.p, s);
    unsigned long long np = 0;
    if (rc == MRAG_OK) {
        cudaError_t e = cudaMemcpyAsync(&np, w->npass.p, 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(MRAG_ERR_CUDA, "mrag_filter_mask: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(s);
    }
    if (n_pass) *n_pass = int64_t(np);
    release_ws(x, w, nullptr);
    return rc;
}

extern "C" int mrag_merge_topk(int device, int n_lists, int nq, int k, const float* d_scores_in,
                               const int64_t* d_rows_in, const int32_t* d_counts_in, int64_t stride_scores,
                               int64_t stride_rows, int64_t stride_counts, float* d_scores_out,
                               int64_t* d_rows_out, int32_t* d_counts_out, void* stream) {
    if (n_lists < 1 || nq < 0 || k < 1) return fail(MRAG_ERR_ARG, "mrag_merge_topk: bad sizes");
    if (nq == 0) return MRAG_OK;
    if (int64_t(n_lists) * k > kXMergeMaxSlots)
        return fail(MRAG_ERR_ARG, "mrag_merge_topk: n_lists*k = %lld > %d", (long long)n_lists * k, kXMergeMaxSlots);
    if (!d_scores_in || !d_rows_in || !d_counts_in || !d_scores_out || !d_rows_out || !d_counts_out)
        return fail(MRAG_ERR_ARG, "mrag_merge_topk: null buffer");
    DeviceGuard g(device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_merge_topk: cudaSetDevice(%d) failed (no CPU path)", device);
    static bool attr_set[64] = {};
    if (device < 64 && !attr_set[device]) {
        CU(cudaFuncSetAttribute(xmerge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kXMergeMaxSlots * 12));
        attr_set[device] = true;
    }
    XMergeArgs a{};
    a.n_lists = n_lists; a.nq = nq; a.k = k;
    a.scores_in = d_scores_in; a.rows_in = d_rows_in; a.counts_in = d_counts_in;
    a.stride_scores = stride_scores; a.stride_rows = stride_rows; a.stride_counts = stride_counts;
    a.scores_out = d_scores_out; a.rows_out = d_rows_out; a.counts_out = d_counts_out;
    const int total = n_lists * k;
    const size_t smem = size_t(host_next_pow2(total < 2 ? 2 : total)) * 12;
    xmerge_kernel<<<nq, kMergeThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    LAUNCHED();
    return MRAG_OK;
}

extern "C" int mrag_exchange_merge(int device, int world, int rank, int nq, int k, void* const* peer_bufs,
                                   int64_t slot_bytes, int64_t scores_off, int64_t counts_off, uint32_t epoch,
                                   float* d_scores_out, int64_t* d_rows_out, int32_t* d_counts_out, void* stream) {
    if (world < 1 || world > 8 || rank < 0 || rank >= world || nq < 1 || k < 1) return fail(MRAG_ERR_ARG, "mrag_exchange_merge: bad sizes");
    if (int64_t(world) * k > kXMergeMaxSlots) return fail(MRAG_ERR_ARG, "mrag_exchange_merge: world*k = %lld > %d", (long long)world * k, kXMergeMaxSlots);
    if (!peer_bufs || !d_scores_out || !d_rows_out || !d_counts_out) return fail(MRAG_ERR_ARG, "mrag_exchange_merge: null buffer");
    if (epoch == 0) return fail(MRAG_ERR_ARG, "mrag_exchange_merge: epochs start at 1 (flags are zero-initialised)");
    DeviceGuard g(device);
    if (!g.ok) return fail(MRAG_ERR_CUDA, "mrag_exchange_merge: cudaSetDevice(%d) failed (no CPU path)", device);
    static bool attr_set[64] = {};
    if (device < 64 && !attr_set[device]) {
        CU(cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kXMergeMaxSlots * 12));
        // same shared-memory carve-out as the scan kernels, so that an exchange block and a scan CTA can share an SM
        // (an SM only changes its carve-out when it is empty)
        CU(cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set[device] = true;
    }
    XchgArgs a{};
    a.world = world; a.rank = rank; a.nq = nq; a.k = k;
    for (int p = 0; p < world; ++p) {
        if (!peer_bufs[p]) return fail(MRAG_ERR_ARG, "mrag_exchange_merge: peer buffer %d is null", p);
        a.bufs[p] = static_cast<unsigned char*>(peer_bufs[p]);
    }
    a.slot_bytes = slot_bytes; a.scores_off = scores_off; a.counts_off = counts_off;
    a.area_bytes = int64_t(world) * slot_bytes;
    a.flags_off = 2 * a.area_bytes;
    a.epoch = epoch;
    a.scores_out = d_scores_out; a.rows_out = d_rows_out; a.counts_out = d_counts_out;
    const int total = world * k;
    const size_t smem = size_t(host_next_pow2(total < 2 ? 2 : total)) * 12;
    // small merges run with 128 threads (40 registers each): such a block fits on an SM beside a resident scan CTA, so an
    // exchange issued on a side stream overlaps the NEXT search's scan (sharded.py, search_async)
    xchg_merge_kernel<<<nq, total <= 256 ? 128 : kMergeThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    LAUNCHED();
    return MRAG_OK;
}

// ------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------
static float phase_ms(const EventSet& ev, int what) {
    if (!ev.recorded) return -1.0f;
    if (cudaEventSynchronize(ev.e[3]) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    cudaError_t e;
    switch (what) {
        case 0: e = cudaEventElapsedTime(&ms, ev.e[0], ev.e[1]); break;
        case 1: e = cudaEventElapsedTime(&ms, ev.e[1], ev.e[2]); break;
        case 2: e = cudaEventElapsedTime(&ms, ev.e[2], ev.e[3]); break;
        case 3: e = cudaEventElapsedTime(&ms, ev.e[0], ev.e[3]); break;
        default: return -1.0f;
    }
    if (e != cudaSuccess) cudaGetLastError();       // (an event MRAG_EVENTS left out)
    return e == cudaSuccess ? ms : -1.0f;
}

extern "C" float mrag_last_kernel_ms(int what) {
    if (!t_last_valid) return -1.0f;
    return phase_ms(t_last_ev, what);
}

extern "C" int mrag_profile_begin(int n) {
    for (auto& e : t_ring) e.destroy();
    t_ring.clear();
    t_ring_used = 0;
    t_ring_device = -1;
    t_last_valid = false;
    if (n <= 0) return MRAG_OK;
    if (n > 4096) return fail(MRAG_ERR_ARG, "mrag_profile_begin: n %d > 4096", n);
    int dev = 0;
    CU(cudaGetDevice(&dev));
    t_ring.resize(size_t(n));
    for (auto& e : t_ring) {
        int rc = e.create();
        if (rc != 0) return MRAG_ERR_CUDA;
    }
    t_ring_device = dev;
    return MRAG_OK;
}

extern "C" int mrag_profile_read(int what, float* out_ms, int max) {
    if (!out_ms || max < 0) return fail(MRAG_ERR_ARG, "mrag_profile_read: bad buffer");
    int n = 0;
    for (int i = 0; i < t_ring_used && n < max; ++i) out_ms[n++] = phase_ms(t_ring[size_t(i)], what);
    return n;
}

// debugging aid (not in mrag.h): counters of the last tensor-core scan on this thread, if the
// environment variable MRAG_SCAN_STATS was set; the caller must have synchronised the search.
extern "C" int mrag_debug_scan_stats(unsigned long long* out40) {
    if (!t_stats_ptr || !out40) return MRAG_ERR_STATE;
    return cudaMemcpy(out40, t_stats_ptr, 40 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? MRAG_OK : MRAG_ERR_CUDA;
}

// debugging aid (not in mrag.h): queries of the last approx+rescore search on this thread that went to the exact rescan
extern "C" int mrag_debug_fallback_count(void) {
    if (!t_fb_ptr) return -1;
    int c = -1;
    return cudaMemcpy(&c, t_fb_ptr, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess ? c : -1;
}

extern "C" int64_t mrag_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* mrag_last_scan_kind(void) { return t_last_kind; }
extern "C" const char* mrag_last_error(void) { return t_err.c_str(); }
extern "C" const char* mrag_version(void) { return "mrag-b200 0.1 (sm_100a)"; }
