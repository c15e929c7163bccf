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
cudaError_t dispatch_search_ch1(const SearchArgs&, int, int, int, cudaStream_t, const LaunchWindow&);
cudaError_t dispatch_search_ch2(const SearchArgs&, int, int, int, cudaStream_t, const LaunchWindow&);
cudaError_t dispatch_search_ch3(const SearchArgs&, int, int, int, cudaStream_t, const LaunchWindow&);
cudaError_t dispatch_search_ch4(const SearchArgs&, int, int, int, cudaStream_t, const LaunchWindow&);
cudaError_t dispatch_search_ch6(const SearchArgs&, int, int, int, cudaStream_t, const LaunchWindow&);
cudaError_t dispatch_search_ch8(const SearchArgs&, int, int, int, cudaStream_t, const LaunchWindow&);
cudaError_t dispatch_search_long(const SearchArgs&, int, int, int, cudaStream_t, const LaunchWindow&);

cudaError_t dispatch_search(const SearchArgs& a, int ch, int row_t, int ef_t, int grid, cudaStream_t st, const LaunchWindow& win) {
    switch (ch) {
        case 1: return dispatch_search_ch1(a, row_t, ef_t, grid, st, win);
        case 2: return dispatch_search_ch2(a, row_t, ef_t, grid, st, win);
        case 3: return dispatch_search_ch3(a, row_t, ef_t, grid, st, win);
        case 4: return dispatch_search_ch4(a, row_t, ef_t, grid, st, win);
        case 5: case 6: return dispatch_search_ch6(a, row_t, ef_t, grid, st, win);
        case 7: case 8: return dispatch_search_ch8(a, row_t, ef_t, grid, st, win);
        default: return dispatch_search_long(a, row_t, ef_t, grid, st, win);
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

// Does any adjacency row list a PointId twice?  (The b16 visited flavour assumes it does not; graphs built by this library or by the
// reference never do.)  One warp per row of `width` <= 128 entries.
__global__ void repeated_ids_kernel(const uint32_t* rows, size_t n_rows, uint32_t width, uint32_t* repeats) {
    const int lane = threadIdx.x & 31;
    const size_t wpb = blockDim.x >> 5;
    for (size_t r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rows; r += (size_t)gridDim.x * wpb) {
        uint32_t e[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) e[t] = (uint32_t)(lane + 32 * t) < width ? rows[r * width + lane + 32 * t] : kInvalid;
        bool rep = false;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t peers = __match_any_sync(kFullMask, e[t] == kInvalid ? (0x80000000u | (uint32_t)lane) + 0u : e[t]);
            rep |= e[t] != kInvalid && (peers & ((1u << lane) - 1u));
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2)
                if (t2 > t)
                    for (int src = 0; src < 32; ++src) {
                        const uint32_t o = __shfl_sync(kFullMask, e[t], src);
                        rep |= o != kInvalid && o == e[t2];
                    }
        }
        if (__any_sync(kFullMask, rep) && lane == 0) atomicAdd(repeats, 1u);
    }
}

__global__ void nsmid_kernel(uint32_t* out) { asm volatile("mov.u32 %0, %%nsmid;" : "=r"(*out)); }

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


// ---------------------------------------------------------------------------------------------------------
// DeviceCtx: the per-device pool of per-warp scratch tables
// ---------------------------------------------------------------------------------------------------------
namespace {
std::mutex g_ctx_mu;
DeviceCtx* g_ctx[64] = {};
int g_l2_pref[64] = {};  // idb_device_set_persisting_l2: 0 default (allowed), -1 disabled
}  // namespace

TablePool DeviceCtx::main_pool(bool b16) const {
    TablePool tp;
    tp.slot_masks = slot_masks;
    tp.fixed_word = -1;
    tp.word_base = 0;
    tp.slots_per_word = (uint32_t)slots_per_sm;
    tp.vis_tables = b16 ? b16_tables : big_tables;
    tp.vis_stride = b16 ? b16_stride : big_stride;
    tp.vis_ext = b16 ? b16_ext : nullptr;
    tp.ext_stride = b16 ? b16_stride : 0;
    tp.tie_tables = tie_tables;
    tp.tie_cap = kTieCap;
    return tp;
}
TablePool DeviceCtx::retry_pool() const {
    TablePool tp;
    tp.slot_masks = slot_masks;
    tp.fixed_word = sm_ids;
    tp.word_base = (uint32_t)sm_ids;
    tp.slots_per_word = kRetryCtas;
    tp.vis_tables = retry_tables;
    tp.vis_stride = kRetrySlots;
    tp.vis_ext = nullptr;
    tp.ext_stride = 0;
    tp.tie_tables = retry_ties;
    tp.tie_cap = kRetryTieCap;
    return tp;
}

idb_status DeviceCtx::acquire(int device, DeviceCtx** out) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (device < 0 || device >= 64) return fail(IDB_ERR_INVALID_ARG, "device %d out of range", device);
    if (g_ctx[device]) {
        g_ctx[device]->refs++;
        *out = g_ctx[device];
        return IDB_OK;
    }
    auto* c = new (std::nothrow) DeviceCtx();
    if (!c) return fail(IDB_ERR_OOM, "host allocation failed");
    c->device = device;
    cudaDeviceProp prop;
    auto bail = [&](cudaError_t e, int line) {
        delete c;
        return fail(e == cudaErrorMemoryAllocation ? IDB_ERR_OOM : IDB_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e), __FILE__,
                    line, cudaGetErrorString(e));
    };
#define CTX_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return bail(e_, __LINE__); } while (0)
    CTX_TRY(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (const char* e = std::getenv("IDB_CTAS_PER_SM")) c->slots_per_sm = std::min(kMaxCtasPerSm, std::max(1, std::atoi(e)));
    {   // SM ids run up to %nsmid, which counts disabled SMs too
        uint32_t* d = nullptr;
        uint32_t h = 0;
        CTX_TRY(cudaMalloc(&d, 4));
        nsmid_kernel<<<1, 1>>>(d);
        cudaError_t e1 = cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
        cudaFree(d);
        CTX_TRY(e1);
        c->sm_ids = std::max<int>((int)h, c->num_sms);
    }
    c->n_tables = (uint32_t)c->sm_ids * (uint32_t)c->slots_per_sm * kSearchWarps;
    c->n_tables_live = (uint32_t)c->num_sms * (uint32_t)c->slots_per_sm * kSearchWarps;
    cudaDeviceGetAttribute(&c->max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
    cudaDeviceGetAttribute(&c->max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
    c->l2_allowed = g_l2_pref[device] >= 0;
    if (const char* e = std::getenv("IDB_L2_PERSIST")) c->l2_allowed = c->l2_allowed && std::atoi(e) != 0;
    // b16 tables: as many bytes per warp as keep ALL tables inside the persisting part of L2 (34 KB on B200: 79 MB / 2368 warps)
    // b16 tables: two segments of `b16_l2_bytes` per warp.  The first — as many bytes per warp as keep ALL live tables inside the
    // persisting part of L2 (32 KB on B200: 79 MB / 2368 warps) and the pool inside one access-policy window — is what normal
    // traversals use; the second serves large ef (up to 2040 buckets over both: the per-row tally has 2048 one-byte entries).
    size_t per = c->max_persist > 0 ? (size_t)c->max_persist / c->n_tables_live : (size_t)32 * 1024;
    per = std::min<size_t>(std::max<size_t>(per / 512 * 512, 8 * 1024), 32 * 1024);
    c->b16_l2_bytes = (uint32_t)per;
    c->b16_stride = (uint32_t)(per / 4);
    CTX_TRY(cudaMalloc(&c->slot_masks, ((size_t)c->sm_ids + 1) * 4));
    CTX_TRY(cudaMemset(c->slot_masks, 0, ((size_t)c->sm_ids + 1) * 4));
    CTX_TRY(cudaMalloc(&c->b16_tables, (size_t)c->n_tables * per));
    CTX_TRY(cudaMemset(c->b16_tables, 0xFF, (size_t)c->n_tables * per));
    CTX_TRY(cudaMalloc(&c->b16_ext, (size_t)c->n_tables * per));
    CTX_TRY(cudaMemset(c->b16_ext, 0xFF, (size_t)c->n_tables * per));
    const size_t retry_words = (size_t)kRetryCtas * kSearchWarps * kRetrySlots;
    CTX_TRY(cudaMalloc(&c->retry_tables, retry_words * 4));
    CTX_TRY(cudaMemset(c->retry_tables, 0xFF, retry_words * 4));
    CTX_TRY(cudaMalloc(&c->tie_tables, (size_t)c->n_tables * kTieCap * 8));
    CTX_TRY(cudaMalloc(&c->retry_ties, (size_t)kRetryCtas * kSearchWarps * kRetryTieCap * 8));
    CTX_TRY(cudaDeviceSynchronize());
#undef CTX_TRY
    c->refs = 1;
    g_ctx[device] = c;
    *out = c;
    return IDB_OK;
}

void DeviceCtx::release(DeviceCtx* c) {
    if (!c) return;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (--c->refs > 0) return;
    g_ctx[c->device] = nullptr;
    delete c;
}

DeviceCtx::~DeviceCtx() {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    if (l2_reserved) {  // hand the persisting lines and the device's L2 set-aside back
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    cudaFree(slot_masks);
    cudaFree(b16_tables);
    cudaFree(b16_ext);
    cudaFree(big_tables);
    cudaFree(retry_tables);
    cudaFree(tie_tables);
    cudaFree(retry_ties);
}

idb_status DeviceCtx::ensure_big(uint32_t stride_words) {
    if (big_tables && stride_words <= big_stride) return IDB_OK;
    CUDA_TRY(cudaDeviceSynchronize());  // nobody may be using the old tables (enqueues are serialised by mu)
    cudaFree(big_tables);
    big_tables = nullptr;
    big_stride = 0;
    const size_t words = (size_t)n_tables * stride_words;
    CUDA_TRY(cudaMalloc(&big_tables, words * 4));
    CUDA_TRY(cudaMemset(big_tables, 0xFF, words * 4));
    CUDA_TRY(cudaDeviceSynchronize());
    big_stride = stride_words;
    return IDB_OK;
}

idb_status DeviceCtx::reserve_l2(size_t bytes) {
    if (!l2_allowed || max_persist <= 0 || max_window <= 0) return IDB_OK;
    bytes = std::min<size_t>(bytes, (size_t)max_persist);
    if (bytes <= l2_reserved) return IDB_OK;
    CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes));
    l2_reserved = bytes;
    return IDB_OK;
}

static bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

cudaError_t HostOut::enqueue(Lane& ln) {
    lane = &ln;
    staged = false;
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        staged = staged || !host_ptr_is_pinned(parts[i].user);
        parts[i].off = total;
        total += (parts[i].bytes + 63) / 64 * 64;
    }
    if (staged && total > ln.h_out_cap) {
        if (ln.h_out) cudaFreeHost(ln.h_out);
        ln.h_out = nullptr;
        ln.h_out_cap = 0;
        const size_t want = total + total / 4;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&ln.h_out), want, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        ln.h_out_cap = want;
    }
    for (int i = 0; i < n; ++i) {
        void* dst = staged ? static_cast<void*>(ln.h_out + parts[i].off) : parts[i].user;
        cudaError_t e = cudaMemcpyAsync(dst, parts[i].dev, parts[i].bytes, cudaMemcpyDeviceToHost, ln.stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void HostOut::finish() const {
    if (!staged) return;
    for (int i = 0; i < n; ++i) std::memcpy(parts[i].user, lane->h_out + parts[i].off, parts[i].bytes);
}

void Lane::free_all() {
    cudaFree(ctrl); cudaFree(status); cudaFree(fail_list); cudaFree(counters);
    cudaFree(q); cudaFree(ids); cudaFree(dist); cudaFree(len);
    cudaFree(keys_local); cudaFree(keys_all); cudaFree(q2); cudaFree(ids2);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_ctrl) cudaEventDestroy(ev_ctrl);
    if (h_ctrl) cudaFreeHost(h_ctrl);
    if (h_out) cudaFreeHost(h_out);
    if (stream) cudaStreamDestroy(stream);
}

void Index::note_overflows(uint32_t ef, uint64_t n_work, uint32_t overflowed, int level) {
    if (level < 1 || level > 2) return;
    if ((uint64_t)overflowed * 1000 > n_work) {
        std::atomic<uint32_t>& d = b16_demote_ef[level - 1];
        uint32_t cur = d.load();
        while (ef < cur && !d.compare_exchange_weak(cur, ef)) {}
    }
}

Lane& Index::pick_lane() {
    // prefer an idle lane; otherwise queue behind the next one in rotation
    for (int i = 0; i < kLanes; ++i) {
        Lane& ln = lanes[(next_lane.load() + i) % kLanes];
        if (ln.mu.try_lock()) {
            next_lane.fetch_add(i + 1);
            return ln;
        }
    }
    Lane& ln = lanes[next_lane.fetch_add(1) % kLanes];
    ln.mu.lock();
    return ln;
}

idb_status Index::attach_window(Lane& ln, const LaunchWindow& win) {
    if (ln.win_base == win.base && ln.win_bytes == win.bytes) return IDB_OK;
    cudaStreamAttrValue av;
    std::memset(&av, 0, sizeof(av));
    if (win.base && win.bytes) {
        av.accessPolicyWindow.base_ptr = win.base;
        av.accessPolicyWindow.num_bytes = win.bytes;
        av.accessPolicyWindow.hitRatio = win.hit_ratio;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    CUDA_TRY(cudaStreamSetAttribute(ln.stream, cudaStreamAttributeAccessPolicyWindow, &av));
    ln.win_base = win.base;
    ln.win_bytes = win.bytes;
    return IDB_OK;
}

idb_status Index::ensure_lane_scratch(Lane& ln, uint64_t nq) {
    if (!ln.ctrl) {
        CUDA_TRY(cudaMalloc(&ln.ctrl, 64));
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&ln.h_ctrl), 128, cudaHostAllocDefault));  // [0,16): sampled tally, [16,32): host API
        CUDA_TRY(cudaEventCreateWithFlags(&ln.ev_ctrl, cudaEventDisableTiming));
    }
    if (ln.ctrl_pending && cudaEventQuery(ln.ev_ctrl) == cudaSuccess) {  // the previous call's tally has arrived
        note_overflows(ln.ctrl_ef, ln.ctrl_nq, ln.h_ctrl[4], ln.ctrl_b16);
        ln.ctrl_pending = false;
    }
    CUDA_TRY(ensure(ln.status, ln.status_cap, nq));
    CUDA_TRY(ensure(ln.fail_list, ln.fail_cap, nq));
    CUDA_TRY(ensure(ln.counters, ln.counters_cap, nq * 4));
    if (profiling && !ln.ev0) {
        CUDA_TRY(cudaEventCreate(&ln.ev0));
        CUDA_TRY(cudaEventCreate(&ln.ev1));
    }
    return IDB_OK;
}

// Which flavour of the big visited tier serves a traversal with this ef, and where its tables are.
//   b16 (default): exact while ceil(n / 32768) <= buckets and no adjacency row repeats an id; ~2 u16 slots per id the traversal can
//     possibly visit (2M per expansion, ~ef expansions), clamped to the per-warp stride (= what fits the persisting part of L2);
//   else bitmap (n bits per warp) when that is no bigger than 2x the hash table, else the hash set.
idb_status Index::select_visited_tier(uint32_t ef, SearchArgs& a, LaunchWindow& win) {
    DeviceCtx& c = *ctx;
    const uint32_t efx = std::max<uint32_t>(ef, 16u);
    // 2.5 u16 slots per id the traversal can possibly insert (2M per expansion, ~ef expansions): typical load 1/3, queries handed to the
    // retry pass beyond 11/16 (profiles/r02_call4_tune_*: 1M x 128 sift, ef 100 / 128 / 200 visit at most 6.0k / 7.2k / 10.7k ids).
    // Tables stay within the L2-resident 32 KB while 0.9 * 2M * ef ids fit below the hand-over point; larger ef uses up to 64 KB.
    const uint64_t want_bytes = ((uint64_t)2 * M * efx * 5 + 511) / 512 * 512;
    const uint32_t seg = c.b16_l2_bytes;                           // bytes per segment
    const uint32_t nb_lo_max = (seg - kB16Stash * 4) / 32;         // the first segment also holds the stash
    // Which flavour (profiles/r02_call13_tune_*):
    //   1. b16 inside the L2-resident first segment while a typical traversal (~0.8 * 2M * ef ids) stays below its hand-over point
    //      (M = 32: up to ef ~ 200) and this index has not been seen to overflow it at this ef;
    //   2. b16 over both segments for BIG indexes only (n > 4M, where the bitmap would be > 512 KB per warp), same conditions;
    //      at 1M points the half-DRAM-resident large table is no faster than the bitmap;
    //   3. bitmap (n bits per warp) when that is no bigger than 2x the hash table, else the hash set.
    auto cap_of = [&](uint32_t buckets) { return (uint64_t)buckets * b16_cap_16ths; };
    const uint64_t typical = (uint64_t)2 * M * efx * 4 / 5;
    int level = 0;
    if (typical <= cap_of(nb_lo_max) && ef < b16_demote_ef[0].load()) level = 1;
    else if (n > 4000000ull && typical <= cap_of(2040u) && ef < b16_demote_ef[1].load()) level = 2;
    uint32_t b16_bytes = (uint32_t)std::min<uint64_t>(want_bytes, level == 2 ? 2ull * seg : seg);
    if (b16_bytes_override) {
        b16_bytes = std::min<uint32_t>(std::max<uint32_t>(b16_bytes_override / 32 * 32, 512u), 2 * seg);
        if (ef < b16_demote_ef[0].load()) level = 1;
    }
    const uint32_t nb = std::min<uint32_t>((b16_bytes - kB16Stash * 4) / 32, 2040u);  // buckets over both segments
    const uint32_t nb_lo = std::min(nb, nb_lo_max);
    const bool b16_exact = rows_distinct && (n + 32767) / 32768 <= nb;
    int tier = vis_tier;
    if (tier == 2 && !b16_exact) tier = -1;
    if (tier < 0) tier = (b16_exact && level > 0) ? 2 : -1;
    b16_level = tier == 2 ? (level > 0 ? level : 1) : 0;
    win = LaunchWindow();
    if (tier == 2) {
        a.pool = c.main_pool(true);
        a.gslots = nb_lo * 8 + kB16Stash;
        a.b16_nb = nb;
        a.gshift = 0;
        a.vis_mode = kVisB16;
        a.b16_cap_ids = nb * b16_cap_16ths;  // <= 11 of 16 slots on average; fuller tables hand the query to the retry pass
        idb_status st = c.reserve_l2((size_t)c.n_tables_live * std::min(b16_bytes, seg));
        if (st != IDB_OK) return st;
        if (c.l2_reserved) {
            win.base = c.b16_tables;
            win.bytes = std::min<size_t>((size_t)c.n_tables * c.b16_stride * 4, (size_t)c.max_window);
            win.hit_ratio = 1.0f;  // the window covers the first segments only; of those only the bytes in use are ever touched
        }
        return IDB_OK;
    }
    uint32_t want_slots = std::max<uint32_t>(1024u, next_pow2((uint64_t)vis_mult * 2 * M * efx));
    if (vis_slots_override) want_slots = vis_slots_override;  // tests: force the overflow -> retry path
    const uint32_t bm_words = (uint32_t)std::min<uint64_t>(((n + 31) / 32 + 127) / 128 * 128, 0xFFFFFF80u);
    const bool bitmap = tier == 1 || (tier < 0 && !vis_slots_override && (n + 31) / 32 <= 2ull * want_slots);
    idb_status st = c.ensure_big(std::max(want_slots, bitmap ? bm_words : 0u));
    if (st != IDB_OK) return st;
    a.pool = c.main_pool(false);
    a.gslots = bitmap ? bm_words : want_slots;
    a.gshift = 32 - (uint32_t)std::log2((double)want_slots);
    a.vis_mode = bitmap ? kVisBitmap : kVisHash;
    a.b16_cap_ids = 0;
    a.b16_nb = 0;
    return IDB_OK;
}

int Index::search_grid() const { return num_sms * ctx->slots_per_sm; }

// Enqueue one batched search on a lane; all pointers are device pointers, d_queries padded to nchunks*4 floats per row.
// The caller holds ln.mu.
idb_status Index::enqueue_search(Lane& ln, const float* d_queries_padded, uint64_t nq, uint32_t ef, uint32_t k, uint32_t* d_ids,
                                 float* d_dist, uint32_t* d_len, uint64_t* out_keys) {
    if (n) ef = (uint32_t)std::min<uint64_t>(ef, n);  // admission is rank < ef and there are only n distinct ids: same results
    if (ef > 1024) return fail(IDB_ERR_UNSUPPORTED, "ef_search %u > 1024 (on an index of more than 1024 points) is not supported", ef);
    idb_status st = ensure_lane_scratch(ln, nq);
    if (st != IDB_OK) return st;
    CUDA_TRY(cudaMemsetAsync(ln.ctrl, 0, 64, ln.stream));

    SearchArgs a;
    std::memset(&a, 0, sizeof(a));
    a.g = view();
    a.queries = reinterpret_cast<const float4*>(d_queries_padded);
    a.n_work = nq;
    a.ef = ef;
    a.k = k;
    a.out_ids = d_ids;
    a.out_dist = d_dist;
    a.out_len = d_len;
    a.counters = ln.counters;
    a.status = ln.status;
    a.work_counter = reinterpret_cast<unsigned long long*>(ln.ctrl);
    a.fail_count = reinterpret_cast<uint32_t*>(ln.ctrl + 16);
    a.fail_list = ln.fail_list;
    a.variant = variant;
    a.out_keys = out_keys;
    a.id_map = d_id_map;

    const int ch = (int)((nchunks + 31) / 32);
    if (ch > 8 && (nchunks + 31) / 32 * 512 > 40 * 1024)
        return fail(IDB_ERR_UNSUPPORTED, "dim %u > 10240 is not supported (the query of a long-row traversal lives in shared memory)", dim);
    const int row_t = (int)((2 * M + 31) / 32);
    const int ef_t = (int)((ef + 31) / 32);
    const int grid = std::max(1, (int)std::min<uint64_t>((nq + kSearchWarps - 1) / kSearchWarps, (uint64_t)search_grid()));

    std::lock_guard<std::mutex> lk(ctx->mu);  // the tables this launch uses must not be regrown under it
    LaunchWindow win;
    st = select_visited_tier(ef, a, win);
    if (st == IDB_OK) st = attach_window(ln, win);
    if (st != IDB_OK) return st;
    if (profiling) CUDA_TRY(cudaEventRecord(ln.ev0, ln.stream));
    CUDA_TRY(dispatch_search(a, ch, row_t, ef_t, grid, ln.stream, win));
    if (profiling) CUDA_TRY(cudaEventRecord(ln.ev1, ln.stream));
    ln.last_launches = 2;  // K1 + the (normally idle) retry pass

    // Retry pass (device-side, unconditional, normally a no-op): queries whose visited table overflowed are re-run
    // by a few warps with 2^18-slot hash sets.  n_work is read from fail_count on the device.
    SearchArgs r = a;
    r.work_list = ln.fail_list;
    r.n_work_dev = a.fail_count;
    r.n_work = 0;
    r.work_counter = reinterpret_cast<unsigned long long*>(ln.ctrl + 32);
    r.fail_count = reinterpret_cast<uint32_t*>(ln.ctrl + 48);
    r.fail_list = nullptr;       // failures of the retry pass are only counted (and visible in status[])
    r.pool = ctx->retry_pool();
    r.gslots = kRetrySlots;
    r.gshift = 32 - 18;
    r.vis_mode = kVisHash;
    static_assert(kRetrySlots == 1u << 18, "gshift above");
    CUDA_TRY(dispatch_search(r, ch, row_t, ef_t, kRetryCtas, ln.stream, LaunchWindow()));
    ln.last_b16 = a.vis_mode == kVisB16 ? b16_level : 0;
    if (!ln.ctrl_pending) {  // sample this call's overflow tally (one read-back in flight per lane; evaluated by a later call)
        CUDA_TRY(cudaMemcpyAsync(ln.h_ctrl, ln.ctrl, 64, cudaMemcpyDeviceToHost, ln.stream));
        CUDA_TRY(cudaEventRecord(ln.ev_ctrl, ln.stream));
        ln.ctrl_pending = true;
        ln.ctrl_b16 = ln.last_b16;
        ln.ctrl_ef = ef;
        ln.ctrl_nq = nq;
    }
    ln.last_nq = nq;
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
    for (auto& ln : lanes)
        if (ln.stream) cudaStreamSynchronize(ln.stream);
    cudaFree(d_points);
    cudaFree(d_points_bf16);
    cudaFree(d_zero);
    for (auto* p : d_upper) cudaFree(p);
    cudaFree(d_upper_ptrs);
    cudaFree(d_id_map);
    for (auto& ln : lanes) ln.free_all();
    DeviceCtx::release(ctx);
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
    idb_status st = DeviceCtx::acquire(dev, &ctx);
    if (st != IDB_OK) return st;
    for (auto& ln : lanes) CUDA_TRY(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
    stream = lanes[0].stream;
    if (const char* e = std::getenv("IDB_OPT")) opt_flags = (uint32_t)std::atoi(e);
    if (const char* e = std::getenv("IDB_VIS_MULT")) vis_mult = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("IDB_VARIANT")) variant = std::atoi(e);
    if (const char* e = std::getenv("IDB_VIS_TIER")) vis_tier = std::atoi(e);
    if (const char* e = std::getenv("IDB_B16_CAP")) b16_cap_16ths = (uint32_t)std::min(14, std::max(1, std::atoi(e)));
    if (const char* e = std::getenv("IDB_B16_BYTES")) b16_bytes_override = (uint32_t)std::max(64, std::atoi(e));
    if (const char* e = std::getenv("IDB_VIS_SLOTS")) vis_slots_override = next_pow2((uint64_t)std::max(64, std::atoi(e)));
    return IDB_OK;
}

// Upload a graph (host arrays) into HBM.
idb_status Index::upload(const float* points, uint64_t n_, uint32_t dim_, uint32_t M_, uint32_t ef, const uint32_t* zero,
                         uint32_t n_upper, const uint32_t* const* upper, const uint64_t* upper_n_) {
    // Layer l holds PointIds [0, n_l) (lib.rs:275-281): n >= n_1 >= n_2 >= ... >= 1.  The descent carries ids found on layer l
    // into layer l-1 and seeds PointId 0 on the top layer, so anything else would read adjacency rows out of bounds.
    for (uint32_t l = 0; l < n_upper; ++l) {
        const uint64_t below = l == 0 ? n_ : upper_n_[l - 1];
        if (upper_n_[l] == 0 || upper_n_[l] > below)
            return fail(IDB_ERR_INVALID_ARG, "layer %u has %llu nodes but the layer below has %llu (need n >= n_1 >= ... >= 1)", l + 1,
                        (unsigned long long)upper_n_[l], (unsigned long long)below);
    }
    n = n_;
    dim = dim_;
    M = M_;
    ef_search = ef;
    nchunks = (dim + 3) / 4;
    if (n == 0) return IDB_OK;
    const size_t stride = (size_t)nchunks * 4;
    if (n > SIZE_MAX / (stride * sizeof(float)) || n > SIZE_MAX / (2 * (size_t)M * 4)) return fail(IDB_ERR_INVALID_ARG, "n * dim overflows size_t");
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
        if (e == cudaSuccess && bad == 0) {  // rows that list a PointId twice are legal input but rule out the b16 visited flavour
            repeated_ids_kernel<<<num_sms * 8, 128, 0, stream>>>(d_zero, n, 2 * M, d_bad);
            for (uint32_t l = 0; l < n_upper; ++l)
                if (upper && upper[l] && upper_n_[l]) repeated_ids_kernel<<<num_sms * 8, 128, 0, stream>>>(d_upper[l], upper_n_[l], M, d_bad);
            uint32_t rep = 0;
            e = cudaMemcpyAsync(&rep, d_bad, 4, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            rows_distinct = rep == 0;
        }
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

}  // extern "C"

namespace idb {
static idb_status search_device_on_lane(Index* ix, Lane& ln, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                        uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len, uint64_t* out_keys) {
    CUDA_TRY(cudaSetDevice(ix->device));
    const uint32_t ef = ef_search ? ef_search : ix->ef_search;
    if (ix->n == 0 || ef == 0) {  // empty index (core:359-361) / ef_search = 0: empty result lists
        CUDA_TRY(fill_u32(d_out_ids, nq * k, kInvalid, ln.stream));
        if (d_out_dist) CUDA_TRY(fill_u32(reinterpret_cast<uint32_t*>(d_out_dist), nq * k, 0x7f800000u, ln.stream));
        if (d_out_len) CUDA_TRY(cudaMemsetAsync(d_out_len, 0, nq * 4, ln.stream));
        if (out_keys) CUDA_TRY(cudaMemsetAsync(out_keys, 0xFF, nq * k * 8, ln.stream));  // kKeyNone everywhere
        ln.last_nq = 0;
        return IDB_OK;
    }
    const float* qp = d_queries;
    const size_t stride = (size_t)ix->nchunks * 4;
    if (stride != ix->dim || (reinterpret_cast<uintptr_t>(d_queries) & 15)) {
        CUDA_TRY(ensure(ln.q, ln.q_cap, nq * stride));
        CUDA_TRY(cudaMemsetAsync(ln.q, 0, nq * stride * 4, ln.stream));
        CUDA_TRY(cudaMemcpy2DAsync(ln.q, stride * 4, d_queries, ix->dim * 4, ix->dim * 4, nq, cudaMemcpyDeviceToDevice, ln.stream));
        qp = ln.q;
    }
    return ix->enqueue_search(ln, qp, nq, ef, k, d_out_ids, d_out_dist, d_out_len, out_keys);
}

// used by sharded.cu: lane 0, caller holds its mutex
idb_status search_device_keys(Index* ix, Lane& ln, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k, uint32_t* d_ids,
                              uint64_t* d_keys) {
    return search_device_on_lane(ix, ln, d_queries, nq, ef_search, k, d_ids, nullptr, nullptr, d_keys);
}
}  // namespace idb

extern "C" {

idb_status idb_search_batch_device_lane(idb_index* index, uint32_t lane, const float* d_queries, uint64_t nq, uint32_t ef_search,
                                        uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    if (lane >= (uint32_t)kLanes) return fail(IDB_ERR_INVALID_ARG, "lane %u out of range (0..%d)", lane, kLanes - 1);
    if (nq == 0) return IDB_OK;
    if (!d_queries || !d_out_ids) return fail(IDB_ERR_INVALID_ARG, "queries/out_ids is null");
    if (k == 0) return fail(IDB_ERR_INVALID_ARG, "k must be >= 1");
    Lane& ln = ix->lanes[lane];
    std::lock_guard<std::mutex> lk(ln.mu);
    ix->last_lane.store((int)lane);
    return search_device_on_lane(ix, ln, d_queries, nq, ef_search, k, d_out_ids, d_out_dist, d_out_len, nullptr);
}

idb_status idb_search_batch_device(idb_index* index, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                   uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len) {
    return idb_search_batch_device_lane(index, 0, d_queries, nq, ef_search, k, d_out_ids, d_out_dist, d_out_len);
}

idb_status idb_search_batch_f32(idb_index* index, const float* queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                uint32_t* out_ids, float* out_dist, uint32_t* out_len) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    if (nq == 0) return IDB_OK;
    if (!queries || !out_ids) return fail(IDB_ERR_INVALID_ARG, "queries/out_ids is null");
    if (k == 0) return fail(IDB_ERR_INVALID_ARG, "k must be >= 1");
    const uint32_t ef = ef_search ? ef_search : ix->ef_search;
    if (ix->n == 0 || ef == 0) {
        for (uint64_t i = 0; i < nq * k; ++i) out_ids[i] = IDB_INVALID;
        if (out_dist) for (uint64_t i = 0; i < nq * k; ++i) out_dist[i] = INFINITY;
        if (out_len) std::memset(out_len, 0, nq * 4);
        return IDB_OK;
    }
    // Hnsw<P>: Sync (core:352-356): any number of host threads may search at once; each call takes an idle lane (own stream and
    // control state), so concurrent callers overlap on the device instead of serialising.
    Lane& ln = ix->pick_lane();
    std::lock_guard<std::mutex> lk(ln.mu, std::adopt_lock);
    ix->last_lane.store((int)(&ln - ix->lanes));
    CUDA_TRY(cudaSetDevice(ix->device));
    const size_t stride = (size_t)ix->nchunks * 4;
    CUDA_TRY(ensure(ln.q, ln.q_cap, nq * stride));
    CUDA_TRY(ensure(ln.ids, ln.ids_cap, nq * k));
    CUDA_TRY(ensure(ln.dist, ln.dist_cap, nq * k));
    CUDA_TRY(ensure(ln.len, ln.len_cap, nq));
    if (stride == ix->dim) {
        CUDA_TRY(cudaMemcpyAsync(ln.q, queries, nq * stride * 4, cudaMemcpyHostToDevice, ln.stream));
    } else {
        CUDA_TRY(cudaMemsetAsync(ln.q, 0, nq * stride * 4, ln.stream));
        CUDA_TRY(cudaMemcpy2DAsync(ln.q, stride * 4, queries, ix->dim * 4, ix->dim * 4, nq, cudaMemcpyHostToDevice, ln.stream));
    }
    idb_status st = ix->enqueue_search(ln, ln.q, nq, ef, k, ln.ids, ln.dist, ln.len, nullptr);
    if (st != IDB_OK) return st;
    HostOut ho;
    ho.add(out_ids, ln.ids, nq * k * 4);
    ho.add(out_dist, ln.dist, nq * k * 4);
    ho.add(out_len, ln.len, nq * 4);
    CUDA_TRY(ho.enqueue(ln));
    // The control block comes back through PINNED memory: a device-to-pageable cudaMemcpyAsync blocks inside the driver until the
    // copy has run (i.e. until this call's K1 has finished) and stalls the launches of other caller threads meanwhile, so concurrent
    // callers would never have a second batch queued behind the running one.
    uint32_t* ctrl = ln.h_ctrl + 16;
    CUDA_TRY(cudaMemcpyAsync(ctrl, ln.ctrl, 64, cudaMemcpyDeviceToHost, ln.stream));
    CUDA_TRY(cudaStreamSynchronize(ln.stream));
    ho.finish();
    ix->note_overflows(ef, nq, ctrl[4], ln.last_b16);
    if (ctrl[12] != 0)  // failures that survived the retry pass
        return fail(IDB_ERR_CAPACITY, "%u of %llu queries overflowed an internal per-query structure (visited table / tie list)",
                    ctrl[12], (unsigned long long)nq);
    return IDB_OK;
}

idb_status idb_last_search_failures(idb_index* index, uint32_t lane, uint32_t* out_failed) {
    if (!index || !out_failed) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = reinterpret_cast<Index*>(index);
    if (lane >= (uint32_t)kLanes) return fail(IDB_ERR_INVALID_ARG, "lane %u out of range", lane);
    Lane& ln = ix->lanes[lane];
    std::lock_guard<std::mutex> lk(ln.mu);
    *out_failed = 0;
    if (!ln.ctrl || ln.last_nq == 0) return IDB_OK;
    CUDA_TRY(cudaSetDevice(ix->device));
    uint32_t ctrl[16];
    CUDA_TRY(cudaMemcpyAsync(ctrl, ln.ctrl, 64, cudaMemcpyDeviceToHost, ln.stream));
    CUDA_TRY(cudaStreamSynchronize(ln.stream));
    *out_failed = ctrl[12];
    return IDB_OK;
}

idb_status idb_last_search_retried(idb_index* index, uint32_t lane, uint32_t* out_retried) {
    if (!index || !out_retried) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = reinterpret_cast<Index*>(index);
    if (lane == 0xFFFFFFFFu) lane = (uint32_t)ix->last_lane.load();  // the lane of the last call issued on this index
    if (lane >= (uint32_t)kLanes) return fail(IDB_ERR_INVALID_ARG, "lane %u out of range", lane);
    Lane& ln = ix->lanes[lane];
    std::lock_guard<std::mutex> lk(ln.mu);
    *out_retried = 0;
    if (!ln.ctrl || ln.last_nq == 0) return IDB_OK;
    CUDA_TRY(cudaSetDevice(ix->device));
    uint32_t ctrl[16];
    CUDA_TRY(cudaMemcpyAsync(ctrl, ln.ctrl, 64, cudaMemcpyDeviceToHost, ln.stream));
    CUDA_TRY(cudaStreamSynchronize(ln.stream));
    *out_retried = ctrl[4];
    return IDB_OK;
}

idb_status idb_last_search_counters(idb_index* index, uint64_t nq, uint64_t* out) {
    if (!index || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = reinterpret_cast<Index*>(index);
    Lane& ln = ix->lanes[ix->last_lane.load()];
    std::lock_guard<std::mutex> lk(ln.mu);
    if (nq > ln.last_nq) return fail(IDB_ERR_INVALID_ARG, "nq exceeds the last search batch (%llu)", (unsigned long long)ln.last_nq);
    CUDA_TRY(cudaSetDevice(ix->device));
    std::vector<uint32_t> tmp(nq * 4);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), ln.counters, nq * 16, cudaMemcpyDeviceToHost, ln.stream));
    CUDA_TRY(cudaStreamSynchronize(ln.stream));
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
    ix->profiling = enabled != 0;  // the events are created by the next call on each lane
    return IDB_OK;
}

idb_status idb_index_last_kernel_ms(idb_index* index, float* out_ms, uint32_t* out_launches) {
    if (!index || !out_ms) return fail(IDB_ERR_INVALID_ARG, "null argument");
    Index* ix = reinterpret_cast<Index*>(index);
    Lane& ln = ix->lanes[ix->last_lane.load()];
    std::lock_guard<std::mutex> lk(ln.mu);
    if (!ix->profiling || !ln.ev0) return fail(IDB_ERR_INVALID_ARG, "profiling is not enabled on this index");
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaEventSynchronize(ln.ev1));
    CUDA_TRY(cudaEventElapsedTime(out_ms, ln.ev0, ln.ev1));
    if (out_launches) *out_launches = ln.last_launches;
    return IDB_OK;
}

void* idb_index_stream(idb_index* index) { return index ? reinterpret_cast<Index*>(index)->stream : nullptr; }
uint32_t idb_index_num_lanes(void) { return (uint32_t)kLanes; }
void* idb_index_lane_stream(idb_index* index, uint32_t lane) {
    return index && lane < (uint32_t)kLanes ? reinterpret_cast<Index*>(index)->lanes[lane].stream : nullptr;
}

idb_status idb_index_sync(idb_index* index) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    CUDA_TRY(cudaSetDevice(ix->device));
    for (auto& ln : ix->lanes) CUDA_TRY(cudaStreamSynchronize(ln.stream));
    return IDB_OK;
}

idb_status idb_device_set_persisting_l2(int32_t device, int32_t enabled) {
    if (device < 0 || device >= 64) return fail(IDB_ERR_INVALID_ARG, "device %d out of range", device);
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    g_l2_pref[device] = enabled ? 0 : -1;
    if (g_ctx[device]) {
        std::lock_guard<std::mutex> lk2(g_ctx[device]->mu);
        g_ctx[device]->l2_allowed = enabled != 0;
        if (!enabled && g_ctx[device]->l2_reserved) {
            CUDA_TRY(cudaSetDevice(device));
            CUDA_TRY(cudaDeviceSynchronize());
            cudaCtxResetPersistingL2Cache();
            CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
            g_ctx[device]->l2_reserved = 0;
        }
    }
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
