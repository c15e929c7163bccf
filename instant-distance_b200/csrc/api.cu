// api.cu — C ABI (include/instant_distance_b200.h) + the batched search kernel.
//
// Host side of the drop-in boundary: owns the row-major point matrix and the per-layer fixed-stride adjacency
// ("CSR with implicit row_ptr = pid * stride", rows INVALID-terminated) in HBM, and launches the sm_100a kernels.
// No PyTorch, no CPU fallback: if the CUDA runtime reports no device every compute entry point fails loudly.
#include "internal.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace idb {

thread_local char g_err[512] = "";

idb_status fail(idb_status st, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return st;
}

// one translation unit per CH (search_chN.cu)
cudaError_t dispatch_search_ch1(const SearchArgs&, int, int, int, cudaStream_t);
cudaError_t dispatch_search_ch2(const SearchArgs&, int, int, int, cudaStream_t);
cudaError_t dispatch_search_ch3(const SearchArgs&, int, int, int, cudaStream_t);
cudaError_t dispatch_search_ch4(const SearchArgs&, int, int, int, cudaStream_t);
cudaError_t dispatch_search_ch6(const SearchArgs&, int, int, int, cudaStream_t);
cudaError_t dispatch_search_ch8(const SearchArgs&, int, int, int, cudaStream_t);

cudaError_t dispatch_search(const SearchArgs& a, int ch, int row_t, int ef_t, int grid, cudaStream_t st) {
    switch (ch) {
        case 1: return dispatch_search_ch1(a, row_t, ef_t, grid, st);
        case 2: return dispatch_search_ch2(a, row_t, ef_t, grid, st);
        case 3: return dispatch_search_ch3(a, row_t, ef_t, grid, st);
        case 4: return dispatch_search_ch4(a, row_t, ef_t, grid, st);
        case 5: case 6: return dispatch_search_ch6(a, row_t, ef_t, grid, st);
        default: return dispatch_search_ch8(a, row_t, ef_t, grid, st);
    }
}

__global__ void distance_kernel(const float4* a, const float4* b, uint32_t nchunks, float* out) {
    const int lane = threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t c = lane; c < nchunks; c += 32) {
        float4 x = a[c], y = b[c];
        float d0 = __fsub_rn(x.x, y.x), d1 = __fsub_rn(x.y, y.y), d2 = __fsub_rn(x.z, y.z), d3 = __fsub_rn(x.w, y.w);
        acc.x = __fmaf_rn(d0, d0, acc.x);
        acc.y = __fmaf_rn(d1, d1, acc.y);
        acc.z = __fmaf_rn(d2, d2, acc.z);
        acc.w = __fmaf_rn(d3, d3, acc.w);
    }
    float s = butterfly_sum(__fadd_rn(__fadd_rn(acc.x, acc.y), __fadd_rn(acc.z, acc.w)));
    if (lane == 0) *out = s;
}

// f32 -> bf16 (round to nearest even) and back (exact), element-wise over the padded row matrix
__global__ void narrow_bf16_kernel(const float* src, uint16_t* dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = __float_as_uint(src[i]);
        uint32_t r;
        if ((b & 0x7fffffffu) > 0x7f800000u) r = (b >> 16) | 0x40u;           // NaN stays NaN
        else r = (b + 0x7fffu + ((b >> 16) & 1u)) >> 16;                       // RNE
        dst[i] = (uint16_t)r;
    }
}
__global__ void widen_bf16_kernel(const uint16_t* src, float* dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __uint_as_float((uint32_t)src[i] << 16);
}

// Adjacency sanity check for graphs adopted from outside (idb_index_from_graph_*, idb_index_load): every entry must be
// INVALID or a PointId below `limit`; otherwise the traversal would read out of bounds.
__global__ void validate_rows_kernel(const uint32_t* rows, size_t count, uint32_t limit, uint32_t* bad) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t v = rows[i];
        if (v != kInvalid && v >= limit) atomicAdd(bad, 1u);
    }
}

__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

cudaError_t fill_u32(uint32_t* p, size_t n, uint32_t v, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    fill_u32_kernel<<<1184, 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}

static uint32_t next_pow2(uint64_t v) {
    uint32_t p = 1;
    while (p < v && p < (1u << 30)) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------------------------
// Scratch management
// ---------------------------------------------------------------------------------------------------------
template <class T>
static cudaError_t ensure(T*& p, size_t& cap, size_t need) {
    if (need <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = need + need / 4 + 64;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
}

cudaError_t ensure_u32(uint32_t*& p, size_t& cap, size_t need) { return ensure(p, cap, need); }
cudaError_t ensure_u64(uint64_t*& p, size_t& cap, size_t need) { return ensure(p, cap, need); }
cudaError_t ensure_f32(float*& p, size_t& cap, size_t need) { return ensure(p, cap, need); }


static std::atomic<int> g_persist_users{0};  // live indexes holding bucket tables (they share the device's persisting-L2 set-aside)

idb_status Index::ensure_search_scratch(uint32_t ef, uint64_t nq, uint32_t k) {
    // visited tables: one per resident warp, sized for >= 2x the worst plausible number of visited ids (2M per expansion)
    uint32_t want_slots = std::max<uint32_t>(1024u, next_pow2((uint64_t)vis_mult * 2 * M * std::max<uint32_t>(ef, 16u)));
    if (vis_slots_override) want_slots = vis_slots_override;  // tests: force the overflow -> retry path
    const uint32_t warps = (uint32_t)search_grid() * kSearchWarps;
    // bitmap flavour of the big tier: n bits per warp, used when that is no bigger than twice the hash table
    const uint32_t bm_words = (uint32_t)std::min<uint64_t>(((n + 31) / 32 + 127) / 128 * 128, 0xFFFFFF80u);
    sc.bm_words = (vis_bitmap && !vis_slots_override && (n + 31) / 32 <= 2ull * want_slots) ? bm_words : 0u;
    const uint32_t want_stride = std::max(want_slots, sc.bm_words);
    if (want_slots > sc.gslots || want_stride > sc.vis_stride || !sc.vis_tables || (vis_slots_override && want_slots != sc.gslots)) {
        if (sc.vis_tables) cudaFree(sc.vis_tables);
        sc.vis_tables = nullptr;
        size_t words = (size_t)warps * want_stride;
        CUDA_TRY(cudaMalloc(&sc.vis_tables, words * 4));
        CUDA_TRY(fill_u32(sc.vis_tables, words, kInvalid, stream));
        sc.vis_stride = want_stride;
        sc.gslots = want_slots;
    }
    // bucket-set flavour (K1's default when it fits): ~1.25 slots per id a query can possibly visit (2M per expansion, ~ef
    // expansions), compact tables, kept in the persisting part of L2 by an access-policy window on this index's stream.
    sc.bucket_slots = 0;
    if (vis_buckets && !vis_slots_override) {
        uint32_t slots = std::max<uint32_t>(1024u, next_pow2(((uint64_t)5 * 2 * M * std::max<uint32_t>(ef, 16u) + 3) / 4));
        if (bucket_slots_override) slots = bucket_slots_override;  // tests: force the overflow -> retry path
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
        const size_t bytes = (size_t)warps * slots * 4;
        if (vis_buckets > 1 || (bytes <= (size_t)max_persist && bytes <= (size_t)max_window)) {
            if (slots != sc.bucket_cap || !sc.bucket_tables) {
                if (!sc.bucket_tables) g_persist_users.fetch_add(1);
                if (sc.bucket_tables) cudaFree(sc.bucket_tables);
                sc.bucket_tables = nullptr;
                CUDA_TRY(cudaMalloc(&sc.bucket_tables, bytes));
                CUDA_TRY(fill_u32(sc.bucket_tables, (size_t)warps * slots, kInvalid, stream));
                sc.bucket_cap = slots;
                if (max_persist > 0 && max_window > 0) {
                    const size_t carve = std::min<size_t>(bytes, (size_t)max_persist);
                    CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
                    cudaStreamAttrValue av;
                    std::memset(&av, 0, sizeof(av));
                    av.accessPolicyWindow.base_ptr = sc.bucket_tables;
                    av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)max_window);
                    av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)av.accessPolicyWindow.num_bytes);
                    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    CUDA_TRY(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
                }
            }
            sc.bucket_slots = slots;
        }
    }
    if (!sc.retry_tables) {
        size_t words = (size_t)kRetryWarps * kRetrySlots;
        CUDA_TRY(cudaMalloc(&sc.retry_tables, words * 4));
        CUDA_TRY(fill_u32(sc.retry_tables, words, kInvalid, stream));
    }
    if (!sc.tie_tables) CUDA_TRY(cudaMalloc(&sc.tie_tables, (size_t)(warps + kRetryWarps) * kTieCap * 8));
    if (!sc.ctrl) CUDA_TRY(cudaMalloc(&sc.ctrl, 64));
    CUDA_TRY(ensure(sc.status, sc.status_cap, nq));
    CUDA_TRY(ensure(sc.fail_list, sc.fail_cap, nq));
    CUDA_TRY(ensure(sc.counters, sc.counters_cap, nq * 4));
    (void)k;
    return IDB_OK;
}

int Index::search_grid() const { return num_sms * ctas_per_sm; }

// Enqueue one batched search; all pointers are device pointers, d_queries padded to nchunks*4 floats per row.
idb_status Index::enqueue_search(const float* d_queries_padded, uint64_t nq, uint32_t ef, uint32_t k, uint32_t* d_ids,
                                 float* d_dist, uint32_t* d_len) {
    if (ef > 512) return fail(IDB_ERR_UNSUPPORTED, "ef_search %u > 512 is not supported yet", ef);
    idb_status st = ensure_search_scratch(ef, nq, k);
    if (st != IDB_OK) return st;
    CUDA_TRY(cudaMemsetAsync(sc.ctrl, 0, 64, stream));

    SearchArgs a;
    a.g = view();
    a.queries = reinterpret_cast<const float4*>(d_queries_padded);
    a.n_work = nq;
    a.n_work_dev = nullptr;
    a.work_list = nullptr;
    a.ef = ef;
    a.k = k;
    a.out_ids = d_ids;
    a.out_dist = d_dist;
    a.out_len = d_len;
    a.counters = sc.counters;
    a.status = sc.status;
    a.work_counter = reinterpret_cast<unsigned long long*>(sc.ctrl);
    a.fail_count = reinterpret_cast<uint32_t*>(sc.ctrl + 16);
    a.fail_list = sc.fail_list;
    if (sc.bucket_slots) {
        a.vis_tables = sc.bucket_tables;
        a.gslots = sc.bucket_slots;
        a.gshift = 32 - (uint32_t)std::log2((double)sc.bucket_slots);
        a.vis_stride = sc.bucket_slots;
        a.vis_mode = kVisBuckets;
    } else {
        a.vis_tables = sc.vis_tables;
        a.gslots = sc.bm_words ? sc.bm_words : sc.gslots;
        a.gshift = 32 - (uint32_t)std::log2((double)sc.gslots);
        a.vis_stride = sc.vis_stride;
        a.vis_mode = sc.bm_words ? kVisBitmap : kVisHash;
    }
    a.tie_tables = sc.tie_tables;
    a.variant = variant;
    a.out_keys = pending_out_keys;
    a.id_map = d_id_map;

    const int ch = (int)((nchunks + 31) / 32);
    if (ch > 8) return fail(IDB_ERR_UNSUPPORTED, "dim %u > 1024 is not supported yet", dim);
    const int row_t = (int)((2 * M + 31) / 32);
    const int ef_t = (int)((ef + 31) / 32);
    const uint64_t warps_needed = nq;
    int grid = search_grid();
    const int min_grid = (int)std::min<uint64_t>((warps_needed + kSearchWarps - 1) / kSearchWarps, (uint64_t)grid);
    grid = std::max(1, min_grid);
    if (profiling) CUDA_TRY(cudaEventRecord(ev0, stream));
    CUDA_TRY(dispatch_search(a, ch, row_t, ef_t, grid, stream));
    if (profiling) CUDA_TRY(cudaEventRecord(ev1, stream));
    last_launches = 2;  // K1 + the (normally idle) retry pass

    // Retry pass (device-side, unconditional, normally a no-op): queries whose visited table overflowed are re-run
    // by a few warps with 2^21-slot tables.  n_work is read from fail_count on the device.
    SearchArgs r = a;
    r.work_list = sc.fail_list;
    r.n_work_dev = a.fail_count;
    r.n_work = 0;
    r.work_counter = reinterpret_cast<unsigned long long*>(sc.ctrl + 32);
    r.fail_count = reinterpret_cast<uint32_t*>(sc.ctrl + 48);
    r.fail_list = nullptr;       // failures of the retry pass are only counted (and visible in status[])
    r.vis_tables = sc.retry_tables;
    r.gslots = kRetrySlots;
    r.gshift = 32 - 21;
    r.vis_stride = kRetrySlots;
    r.vis_mode = kVisHash;
    r.tie_tables = sc.tie_tables + (size_t)search_grid() * kSearchWarps * kTieCap;
    CUDA_TRY(dispatch_search(r, ch, row_t, ef_t, kRetryWarps / kSearchWarps, stream));
    last_nq = nq;
    return IDB_OK;
}

idb_status Index::narrow_points_to_bf16() {
    const size_t total = n * (size_t)nchunks * 4;
    if (total == 0) { bf16 = true; return IDB_OK; }
    CUDA_TRY(cudaMalloc(&d_points_bf16, total * 2));
    narrow_bf16_kernel<<<num_sms * 8, 256, 0, stream>>>(d_points, d_points_bf16, total);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    cudaFree(d_points);
    d_points = nullptr;
    bf16 = true;
    return IDB_OK;
}

idb_status Index::copy_points_f32(float* host_out, uint64_t r0, uint64_t m) {
    if (m == 0) return IDB_OK;
    const size_t stride = (size_t)nchunks * 4;
    if (!bf16) {
        CUDA_TRY(cudaMemcpy2DAsync(host_out, dim * 4, d_points + r0 * stride, stride * 4, dim * 4, m, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        return IDB_OK;
    }
    float* tmp = nullptr;
    CUDA_TRY(cudaMalloc(&tmp, m * stride * 4));
    widen_bf16_kernel<<<num_sms * 8, 256, 0, stream>>>(d_points_bf16 + r0 * stride, tmp, m * stride);
    cudaError_t e = cudaMemcpy2DAsync(host_out, dim * 4, tmp, stride * 4, dim * 4, m, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp);
    CUDA_TRY(e);
    return IDB_OK;
}

GraphView Index::view() const {
    GraphView g;
    g.points = bf16 ? reinterpret_cast<const char*>(d_points_bf16) : reinterpret_cast<const char*>(d_points);
    g.bf16 = bf16 ? 1u : 0u;
    g.nchunks = nchunks;
    g.zero = d_zero;
    g.upper = d_upper_ptrs;
    g.n_upper = (uint32_t)d_upper.size();
    g.M = M;
    g.n = n;
    g.flags = opt_flags;
    return g;
}

Index::~Index() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(d_points);
    cudaFree(d_points_bf16);
    cudaFree(d_zero);
    for (auto* p : d_upper) cudaFree(p);
    cudaFree(d_upper_ptrs);
    cudaFree(sc.vis_tables);
    if (sc.bucket_tables) {  // hand the persisting lines back; the last user also returns the L2 set-aside
        cudaCtxResetPersistingL2Cache();
        if (g_persist_users.fetch_sub(1) == 1) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    cudaFree(sc.bucket_tables);
    cudaFree(sc.retry_tables);
    cudaFree(sc.tie_tables);
    cudaFree(sc.ctrl);
    cudaFree(sc.status);
    cudaFree(sc.fail_list);
    cudaFree(sc.counters);
    cudaFree(sc.q);
    cudaFree(sc.ids);
    cudaFree(sc.dist);
    cudaFree(sc.len);
    cudaFree(sc.keys_local);
    cudaFree(sc.keys_all);
    cudaFree(sc.q2);
    cudaFree(sc.ids2);
    cudaFree(d_id_map);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
}

idb_status Index::init_device(int dev) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(IDB_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (dev < 0 || dev >= count) return fail(IDB_ERR_INVALID_ARG, "device %d out of range (0..%d)", dev, count - 1);
    device = dev;
    CUDA_TRY(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(IDB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (const char* e = std::getenv("IDB_OPT")) opt_flags = (uint32_t)std::atoi(e);
    if (const char* e = std::getenv("IDB_VIS_MULT")) vis_mult = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("IDB_CTAS_PER_SM")) ctas_per_sm = std::min(kMaxCtasPerSm, std::max(1, std::atoi(e)));
    if (const char* e = std::getenv("IDB_VARIANT")) variant = std::atoi(e);
    if (const char* e = std::getenv("IDB_VIS_BITMAP")) vis_bitmap = std::atoi(e);
    if (const char* e = std::getenv("IDB_VIS_BUCKETS")) vis_buckets = std::atoi(e);
    if (const char* e = std::getenv("IDB_BUCKET_SLOTS")) bucket_slots_override = next_pow2((uint64_t)std::max(64, std::atoi(e)));
    if (const char* e = std::getenv("IDB_VIS_SLOTS")) vis_slots_override = next_pow2((uint64_t)std::max(64, std::atoi(e)));
    return IDB_OK;
}

// Upload a graph (host arrays) into HBM.
idb_status Index::upload(const float* points, uint64_t n_, uint32_t dim_, uint32_t M_, uint32_t ef, const uint32_t* zero,
                         uint32_t n_upper, const uint32_t* const* upper, const uint64_t* upper_n_) {
    n = n_;
    dim = dim_;
    M = M_;
    ef_search = ef;
    nchunks = (dim + 3) / 4;
    if (n == 0) return IDB_OK;
    const size_t stride = (size_t)nchunks * 4;
    CUDA_TRY(cudaMalloc(&d_points, n * stride * sizeof(float)));
    if (stride == dim) {
        CUDA_TRY(cudaMemcpyAsync(d_points, points, n * stride * sizeof(float), cudaMemcpyHostToDevice, stream));
    } else {
        CUDA_TRY(cudaMemsetAsync(d_points, 0, n * stride * sizeof(float), stream));
        CUDA_TRY(cudaMemcpy2DAsync(d_points, stride * sizeof(float), points, dim * sizeof(float), dim * sizeof(float), n,
                                   cudaMemcpyHostToDevice, stream));
    }
    CUDA_TRY(cudaMalloc(&d_zero, n * 2 * (size_t)M * 4));
    if (zero) CUDA_TRY(cudaMemcpyAsync(d_zero, zero, n * 2 * (size_t)M * 4, cudaMemcpyHostToDevice, stream));
    std::vector<const uint32_t*> ptrs;
    for (uint32_t l = 0; l < n_upper; ++l) {
        uint32_t* p = nullptr;
        CUDA_TRY(cudaMalloc(&p, std::max<size_t>(4, upper_n_[l] * (size_t)M * 4)));
        d_upper.push_back(p);
        upper_n.push_back(upper_n_[l]);
        if (upper && upper[l])
            CUDA_TRY(cudaMemcpyAsync(p, upper[l], upper_n_[l] * (size_t)M * 4, cudaMemcpyHostToDevice, stream));
        ptrs.push_back(p);
    }
    CUDA_TRY(cudaMalloc(&d_upper_ptrs, std::max<size_t>(1, n_upper) * sizeof(uint32_t*)));
    if (n_upper)
        CUDA_TRY(cudaMemcpyAsync(d_upper_ptrs, ptrs.data(), n_upper * sizeof(uint32_t*), cudaMemcpyHostToDevice, stream));
    // reject graphs whose adjacency points outside the layer it belongs to
    if (zero) {
        uint32_t* d_bad = nullptr;
        CUDA_TRY(cudaMalloc(&d_bad, 4));
        CUDA_TRY(cudaMemsetAsync(d_bad, 0, 4, stream));
        validate_rows_kernel<<<num_sms * 4, 256, 0, stream>>>(d_zero, n * 2 * (size_t)M, (uint32_t)n, d_bad);
        for (uint32_t l = 0; l < n_upper; ++l)
            if (upper && upper[l] && upper_n_[l])
                validate_rows_kernel<<<num_sms * 4, 256, 0, stream>>>(d_upper[l], upper_n_[l] * (size_t)M, (uint32_t)upper_n_[l], d_bad);
        uint32_t bad = 0;
        cudaError_t e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(d_bad);
        CUDA_TRY(e);
        if (bad) return fail(IDB_ERR_INVALID_ARG, "%u adjacency entries refer to PointIds outside their layer", bad);
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    return IDB_OK;
}

}  // namespace idb

using namespace idb;

// =========================================================================================================
// extern "C"
// =========================================================================================================
extern "C" {

const char* idb_last_error(void) { return g_err; }
const char* idb_version(void) { return "instant-distance-b200 0.1.0 (sm_100a)"; }

int32_t idb_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}

idb_status idb_params_default(idb_params* p) {
    if (!p) return fail(IDB_ERR_INVALID_ARG, "params is null");
    p->M = 32;                               // core:787
    p->ef_construction = 100;                // core:105
    p->ef_search = 100;                      // core:104
    p->ml = 1.0f / std::log((float)32);      // core:107
    p->seed = 0;                             // core:108 draws from entropy; the C ABI makes it explicit
    p->heuristic = 1;                        // core:106
    p->extend_candidates = 0;                // core:124
    p->keep_pruned = 1;                      // core:125
    p->insert_batch = 0;
    p->device = 0;
    p->storage = IDB_STORAGE_F32;
    p->progress = nullptr;
    p->progress_user = nullptr;
    return IDB_OK;
}

static idb_status index_from_graph(const float* points, uint64_t n, uint32_t dim, uint32_t M, uint32_t ef_search,
                                   const uint32_t* zero, uint32_t n_upper, const uint32_t* const* upper,
                                   const uint64_t* upper_n, int32_t device, bool bf16, idb_index** out_index) {
    if (!out_index) return fail(IDB_ERR_INVALID_ARG, "out_index is null");
    *out_index = nullptr;
    if (dim == 0) return fail(IDB_ERR_INVALID_ARG, "dim must be >= 1");
    if (M < 2 || M > 64) return fail(IDB_ERR_INVALID_ARG, "M = %u unsupported (2..64)", M);
    if (n >= 0xFFFFFFFFull) return fail(IDB_ERR_INVALID_ARG, "N = %llu >= u32::MAX (lib.rs:256)", (unsigned long long)n);
    if (n && (!points || !zero)) return fail(IDB_ERR_INVALID_ARG, "points/zero is null");
    if (n_upper > 31) return fail(IDB_ERR_INVALID_ARG, "too many layers");
    if (n_upper && (!upper || !upper_n)) return fail(IDB_ERR_INVALID_ARG, "upper/upper_n is null");
    auto* ix = new (std::nothrow) Index();
    if (!ix) return fail(IDB_ERR_OOM, "host allocation failed");
    idb_status st = ix->init_device(device);
    if (st == IDB_OK) st = ix->upload(points, n, dim, M, ef_search, zero, n_upper, upper, upper_n);
    if (st == IDB_OK && bf16) st = ix->narrow_points_to_bf16();
    if (st != IDB_OK) { delete ix; return st; }
    *out_index = reinterpret_cast<idb_index*>(ix);
    return IDB_OK;
}

idb_status idb_index_from_graph_f32(const float* points, uint64_t n, uint32_t dim, uint32_t M, uint32_t ef_search,
                                    const uint32_t* zero, uint32_t n_upper, const uint32_t* const* upper,
                                    const uint64_t* upper_n, int32_t device, idb_index** out_index) {
    return index_from_graph(points, n, dim, M, ef_search, zero, n_upper, upper, upper_n, device, false, out_index);
}

idb_status idb_index_from_graph_bf16(const float* points, uint64_t n, uint32_t dim, uint32_t M, uint32_t ef_search,
                                     const uint32_t* zero, uint32_t n_upper, const uint32_t* const* upper,
                                     const uint64_t* upper_n, int32_t device, idb_index** out_index) {
    return index_from_graph(points, n, dim, M, ef_search, zero, n_upper, upper, upper_n, device, true, out_index);
}

idb_status idb_search_batch_device(idb_index* index, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                   uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    if (nq == 0) return IDB_OK;
    if (!d_queries || !d_out_ids) return fail(IDB_ERR_INVALID_ARG, "queries/out_ids is null");
    if (k == 0) return fail(IDB_ERR_INVALID_ARG, "k must be >= 1");
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    const uint32_t ef = ef_search ? ef_search : ix->ef_search;
    if (ix->n == 0 || ef == 0) {  // empty index (core:359-361) / ef_search = 0: empty result lists
        CUDA_TRY(fill_u32(d_out_ids, nq * k, kInvalid, ix->stream));
        if (d_out_dist) CUDA_TRY(fill_u32(reinterpret_cast<uint32_t*>(d_out_dist), nq * k, 0x7f800000u, ix->stream));
        if (d_out_len) CUDA_TRY(cudaMemsetAsync(d_out_len, 0, nq * 4, ix->stream));
        ix->last_nq = 0;
        return IDB_OK;
    }
    const float* qp = d_queries;
    const size_t stride = (size_t)ix->nchunks * 4;
    if (stride != ix->dim || (reinterpret_cast<uintptr_t>(d_queries) & 15)) {
        CUDA_TRY(ensure(ix->sc.q, ix->sc.q_cap, nq * stride));
        CUDA_TRY(cudaMemsetAsync(ix->sc.q, 0, nq * stride * 4, ix->stream));
        CUDA_TRY(cudaMemcpy2DAsync(ix->sc.q, stride * 4, d_queries, ix->dim * 4, ix->dim * 4, nq, cudaMemcpyDeviceToDevice,
                                   ix->stream));
        qp = ix->sc.q;
    }
    return ix->enqueue_search(qp, nq, ef, k, d_out_ids, d_out_dist, d_out_len);
}

idb_status idb_search_batch_f32(idb_index* index, const float* queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                uint32_t* out_ids, float* out_dist, uint32_t* out_len) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    if (nq == 0) return IDB_OK;
    if (!queries || !out_ids) return fail(IDB_ERR_INVALID_ARG, "queries/out_ids is null");
    if (k == 0) return fail(IDB_ERR_INVALID_ARG, "k must be >= 1");
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    const uint32_t ef = ef_search ? ef_search : ix->ef_search;
    if (ix->n == 0 || ef == 0) {
        for (uint64_t i = 0; i < nq * k; ++i) out_ids[i] = IDB_INVALID;
        if (out_dist) for (uint64_t i = 0; i < nq * k; ++i) out_dist[i] = INFINITY;
        if (out_len) std::memset(out_len, 0, nq * 4);
        ix->last_nq = 0;
        return IDB_OK;
    }
    const size_t stride = (size_t)ix->nchunks * 4;
    CUDA_TRY(ensure(ix->sc.q, ix->sc.q_cap, nq * stride));
    CUDA_TRY(ensure(ix->sc.ids, ix->sc.ids_cap, nq * k));
    CUDA_TRY(ensure(ix->sc.dist, ix->sc.dist_cap, nq * k));
    CUDA_TRY(ensure(ix->sc.len, ix->sc.len_cap, nq));
    if (stride == ix->dim) {
        CUDA_TRY(cudaMemcpyAsync(ix->sc.q, queries, nq * stride * 4, cudaMemcpyHostToDevice, ix->stream));
    } else {
        CUDA_TRY(cudaMemsetAsync(ix->sc.q, 0, nq * stride * 4, ix->stream));
        CUDA_TRY(cudaMemcpy2DAsync(ix->sc.q, stride * 4, queries, ix->dim * 4, ix->dim * 4, nq, cudaMemcpyHostToDevice, ix->stream));
    }
    idb_status st = ix->enqueue_search(ix->sc.q, nq, ef, k, ix->sc.ids, ix->sc.dist, ix->sc.len);
    if (st != IDB_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(out_ids, ix->sc.ids, nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    if (out_dist) CUDA_TRY(cudaMemcpyAsync(out_dist, ix->sc.dist, nq * k * 4, cudaMemcpyDeviceToHost, ix->stream));
    if (out_len) CUDA_TRY(cudaMemcpyAsync(out_len, ix->sc.len, nq * 4, cudaMemcpyDeviceToHost, ix->stream));
    uint32_t ctrl[16];
    CUDA_TRY(cudaMemcpyAsync(ctrl, ix->sc.ctrl, 64, cudaMemcpyDeviceToHost, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    if (ctrl[12] != 0)  // failures that survived the retry pass
        return fail(IDB_ERR_CAPACITY, "%u of %llu queries overflowed an internal per-query structure (visited table / tie list)",
                    ctrl[12], (unsigned long long)nq);
    return IDB_OK;
}

idb_status idb_last_search_counters(idb_index* index, uint64_t nq, uint64_t* out) {
    if (!index || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = reinterpret_cast<Index*>(index);
    std::lock_guard<std::mutex> lk(ix->mu);
    if (nq > ix->last_nq) return fail(IDB_ERR_INVALID_ARG, "nq exceeds the last search batch (%llu)", (unsigned long long)ix->last_nq);
    CUDA_TRY(cudaSetDevice(ix->device));
    std::vector<uint32_t> tmp(nq * 4);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), ix->sc.counters, nq * 16, cudaMemcpyDeviceToHost, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    for (uint64_t i = 0; i < nq * 4; ++i) out[i] = tmp[i];
    return IDB_OK;
}

idb_status idb_index_info(const idb_index* index, idb_info* out) {
    if (!index || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    const Index* ix = reinterpret_cast<const Index*>(index);
    std::memset(out, 0, sizeof(*out));
    out->n = ix->n;
    out->dim = ix->dim;
    out->M = ix->M;
    out->ef_search = ix->ef_search;
    out->device = ix->device;
    out->storage = ix->bf16 ? IDB_STORAGE_BF16 : IDB_STORAGE_F32;
    out->n_layers = ix->n == 0 ? 0 : (uint32_t)ix->d_upper.size() + 1;
    if (ix->n) out->layer_n[0] = ix->n;
    for (size_t l = 0; l < ix->upper_n.size() && l + 1 < 32; ++l) out->layer_n[l + 1] = ix->upper_n[l];
    return IDB_OK;
}

idb_status idb_index_export_points(const idb_index* index, float* out) {
    if (!index || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = const_cast<Index*>(reinterpret_cast<const Index*>(index));
    if (ix->n == 0) return IDB_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    idb_status st = ix->copy_points_f32(out, 0, ix->n);
    if (st != IDB_OK) return st;
    return IDB_OK;
}

idb_status idb_index_export_zero(const idb_index* index, uint32_t* out) {
    if (!index || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = const_cast<Index*>(reinterpret_cast<const Index*>(index));
    if (ix->n == 0) return IDB_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaMemcpyAsync(out, ix->d_zero, ix->n * 2 * (size_t)ix->M * 4, cudaMemcpyDeviceToHost, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return IDB_OK;
}

idb_status idb_index_export_upper(const idb_index* index, uint32_t layer, uint32_t* out) {
    if (!index || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = const_cast<Index*>(reinterpret_cast<const Index*>(index));
    if (layer == 0 || layer > ix->d_upper.size()) return fail(IDB_ERR_INVALID_ARG, "layer %u out of range", layer);
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaMemcpyAsync(out, ix->d_upper[layer - 1], ix->upper_n[layer - 1] * (size_t)ix->M * 4, cudaMemcpyDeviceToHost, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return IDB_OK;
}

idb_status idb_index_set_profiling(idb_index* index, int32_t enabled) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    if (enabled && !ix->ev0) {
        CUDA_TRY(cudaEventCreate(&ix->ev0));
        CUDA_TRY(cudaEventCreate(&ix->ev1));
    }
    ix->profiling = enabled != 0;
    return IDB_OK;
}

idb_status idb_index_last_kernel_ms(idb_index* index, float* out_ms, uint32_t* out_launches) {
    if (!index || !out_ms) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = reinterpret_cast<Index*>(index);
    std::lock_guard<std::mutex> lk(ix->mu);
    if (!ix->profiling || !ix->ev0) return fail(IDB_ERR_INVALID_ARG, "profiling is not enabled on this index");
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaEventSynchronize(ix->ev1));
    CUDA_TRY(cudaEventElapsedTime(out_ms, ix->ev0, ix->ev1));
    if (out_launches) *out_launches = ix->last_launches;
    return IDB_OK;
}

void* idb_index_stream(idb_index* index) { return index ? reinterpret_cast<Index*>(index)->stream : nullptr; }

idb_status idb_index_sync(idb_index* index) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return IDB_OK;
}

void idb_index_free(idb_index* index) { delete reinterpret_cast<Index*>(index); }

idb_status idb_distance_f32(const float* a, const float* b, uint32_t dim, int32_t device, float* out) {
    if (!a || !b || !out || dim == 0) return fail(IDB_ERR_INVALID_ARG, "null argument or dim == 0");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(IDB_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    const uint32_t nchunks = (dim + 3) / 4;
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (2 * (size_t)nchunks * 4 + 4) * sizeof(float)));
    cudaMemset(d, 0, (2 * (size_t)nchunks * 4 + 4) * sizeof(float));
    cudaMemcpy(d, a, dim * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d + nchunks * 4, b, dim * 4, cudaMemcpyHostToDevice);
    distance_kernel<<<1, 32>>>(reinterpret_cast<const float4*>(d), reinterpret_cast<const float4*>(d + nchunks * 4), nchunks,
                               d + 2 * (size_t)nchunks * 4);
    cudaError_t e = cudaMemcpy(out, d + 2 * (size_t)nchunks * 4, 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(IDB_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e));
    return IDB_OK;
}

idb_status idb_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(IDB_ERR_INVALID_ARG, "out is null");
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(IDB_ERR_OOM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return IDB_OK;
}
void idb_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
