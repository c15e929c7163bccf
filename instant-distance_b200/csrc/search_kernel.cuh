// search_kernel.cuh — K1: batched Hnsw::search kernel (lib.rs:352-383 per query) and its launch helpers.
// Instantiated once per CH (float4 chunks per lane) in search_chN.cu so the translation units build in parallel.
#pragma once
#include <cstring>

#include "internal.cuh"

namespace idb {

// ---------------------------------------------------------------------------------------------------------
// K1: batched Hnsw::search — persistent grid, one warp per live query, queries claimed from an atomic counter.
// ---------------------------------------------------------------------------------------------------------
// FULL: dim is a multiple of 128, i.e. every lane owns a real chunk in each of its CH slots: no chunk predicates, and
// full batches of row loads carry no predicates at all (hnsw_device.cuh batch_distances_impl).
// Shared-memory carve-up of one traversal warp (K1 and the build's KA).
template <int EF_T>
struct WarpSmem {
    static constexpr int kNearBytes = 2 * 32 * EF_T * 8;
    static constexpr int kBytes = kNearBytes + kSmallVisSlots * 4 + 128 * 4 + 128 * 8;
    static __device__ __forceinline__ void carve(WarpState& s, unsigned char* base) {
        s.near_base = reinterpret_cast<uint64_t*>(base);
        s.near_len = 32 * EF_T;
        s.vis.small = reinterpret_cast<uint32_t*>(base + kNearBytes);
        s.cpid = s.vis.small + kSmallVisSlots;
        s.ckey = reinterpret_cast<uint64_t*>(s.cpid + 128);
        s.vis.hist = s.vis.small;  // the b16 tally borrows the small tier's 2 KB while the big tier is live (hnsw_device.cuh)
    }
};
// Long rows (CH == 0): the warps' query buffers follow the per-warp traversal state in dynamic shared memory.
__host__ __device__ inline uint32_t long_q_bytes(uint32_t nchunks) { return (nchunks + 31) / 32 * 32 * 16; }
template <int EF_T, int CH>
__device__ __forceinline__ void long_q_bind(QVec<CH>& q, unsigned char* smem_raw, uint32_t nchunks, int warp, int warps_per_cta) {
    if constexpr (CH == 0) {
        q.ngroups = (nchunks + 31) / 32;
        q.s = reinterpret_cast<float4*>(smem_raw + (size_t)warps_per_cta * WarpSmem<EF_T>::kBytes + (size_t)warp * long_q_bytes(nchunks));
    }
}
// Point the warp at its claimed scratch tables.
// b16 flavour: gslots = words of the first segment in use (8 * nb_lo + stash), b16_nb = buckets over both segments.
__device__ __forceinline__ void bind_tables(WarpState& s, const TablePool& tp, uint32_t table, uint32_t gslots, uint32_t gshift,
                                            uint32_t mode, uint32_t cap_ids, uint32_t b16_nb) {
    s.vis.big = tp.vis_tables + (size_t)table * tp.vis_stride;
    s.vis.gslots = gslots;
    s.vis.gshift = gshift;
    s.vis.mode = mode;
    s.vis.nb_lo = mode == kVisB16 ? (gslots - kB16Stash) >> 3 : 1u;
    s.vis.nb = mode == kVisB16 ? b16_nb : 1u;
    s.vis.big_hi = mode == kVisB16 && tp.vis_ext ? tp.vis_ext + (size_t)table * tp.ext_stride - (size_t)s.vis.nb_lo * 8 : s.vis.big;
    s.vis.nb_inv = 1.0f / (float)s.vis.nb;
    s.vis.cap_ids = cap_ids;
    s.vis.stash_cnt = 0;
    s.vis.count = 0;
    s.vis.use_big = false;
    s.ties = tp.tie_tables + (size_t)table * tp.tie_cap;
    s.tie_cap = tp.tie_cap;
}

template <int CH, int ROW_T, int EF_T, int B, int OCC, class RT = RowF32, bool FULL = false, bool TMA = false>
__global__ void __launch_bounds__(kSearchWarps * 32, OCC) search_kernel(SearchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_claim[2];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned long long n_work = a.n_work_dev ? (unsigned long long)*a.n_work_dev : a.n_work;
    if (n_work == 0) return;  // the retry pass, normally: nothing to do, no tables claimed

    WarpState s;
    WarpSmem<EF_T>::carve(s, smem_raw + (size_t)warp * WarpSmem<EF_T>::kBytes);
    const uint32_t table0 = cta_tables_acquire(a.pool, s_claim, kSearchWarps);
    bind_tables(s, a.pool, table0 + warp, a.gslots, a.gshift, a.vis_mode, a.b16_cap_ids, a.b16_nb);
    vis_clear_small(s.vis, lane);  // the big tables are handed over clean by their previous holder
    if constexpr (TMA) {  // EXPERIMENT: per-warp ring of B rows + its mbarrier behind the traversal state
        unsigned char* rb = smem_raw + (size_t)kSearchWarps * WarpSmem<EF_T>::kBytes;
        s.mbar = reinterpret_cast<uint64_t*>(rb) + warp;
        s.ring = reinterpret_cast<char*>(rb + 64 + (size_t)warp * B * a.g.nchunks * 16);
        s.mbar_phase = 0;
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(s.mbar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }

    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1ull);
        w = __shfl_sync(kFullMask, w, 0);
        if (w >= n_work) break;
        const uint64_t qi = a.work_list ? a.work_list[w] : w;

        QVec<CH> q;
        long_q_bind<EF_T>(q, smem_raw, a.g.nchunks, warp, kSearchWarps);
        q_from_f32<CH>(q, a.queries + qi * a.g.nchunks, a.g.nchunks, lane);

        descend<CH, ROW_T, EF_T, B, false, RT, FULL, TMA>(a.g, s, q, 0u, a.ef, lane, a.counters ? a.counters + qi * 4 : nullptr);

        const bool ok = s.status == kQueryOk;
        const uint64_t* near = (s.near_base + s.cur * s.near_len);
        const uint32_t len = ok ? s.cnt : 0u;
        for (uint32_t j = lane; j < a.k; j += 32) {
            uint64_t key = j < len ? near[j] : 0ull;
            const uint32_t gid = j < len ? (a.id_map ? a.id_map[key_pid(key)] : key_pid(key)) : kInvalid;
            a.out_ids[qi * a.k + j] = gid;
            if (a.out_keys) a.out_keys[qi * a.k + j] = j < len ? (((uint64_t)key_dbits(key) << 32) | gid) : kKeyNone;
            if (a.out_dist) a.out_dist[qi * a.k + j] = j < len ? __uint_as_float(key_dbits(key)) : __int_as_float(0x7f800000);
        }
        if (lane == 0) {
            if (a.out_len) a.out_len[qi] = len;
            a.status[qi] = s.status;
            if (!ok) {
                uint32_t slot = atomicAdd(a.fail_count, 1u);
                if (a.fail_list) a.fail_list[slot] = (uint32_t)qi;
            }
        }
        finish_query(s, lane);
    }
    cta_tables_release(a.pool, s_claim);
}

// Launch with (optionally) a persisting-L2 access-policy window on the b16 visited tables as a LAUNCH attribute: no stream state.
template <class Kern, class Args>
static cudaError_t launch_with_window(Kern kern, int grid, int block, int smem, cudaStream_t stream, const LaunchWindow& win,
                                      const Args& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (win.base && win.bytes) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = win.base;
        attr[0].val.accessPolicyWindow.num_bytes = win.bytes;
        attr[0].val.accessPolicyWindow.hitRatio = win.hit_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, a);
}

template <int CH, int ROW_T, int EF_T, int B, int OCC = kSearchCtasPerSm, class RT = RowF32, bool FULL = false, bool TMA = false>
static cudaError_t launch_search(const SearchArgs& a, int grid, cudaStream_t stream, const LaunchWindow& win) {
    const int smem = (WarpSmem<EF_T>::kBytes + (CH == 0 ? (int)long_q_bytes(a.g.nchunks) : 0)) * kSearchWarps +
                     (TMA ? 64 + kSearchWarps * B * (int)a.g.nchunks * 16 : 0);
    auto kern = search_kernel<CH, ROW_T, EF_T, B, OCC, RT, FULL, TMA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return launch_with_window(kern, grid, kSearchWarps * 32, smem, stream, win, a);
}

template <int CH, int ROW_T, int EF_T, int B, class RT>
cudaError_t launch_search_full(const SearchArgs& a, int grid, cudaStream_t st, const LaunchWindow& win) {
    if constexpr (CH > 0) {
        if (a.g.nchunks == 32u * CH) return launch_search<CH, ROW_T, EF_T, B, kSearchCtasPerSm, RT, true>(a, grid, st, win);
    }
    return launch_search<CH, ROW_T, EF_T, B, kSearchCtasPerSm, RT, false>(a, grid, st, win);
}
template <int CH, int B, class RT>
cudaError_t dispatch_row_ef_rt(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st, const LaunchWindow& win) {
    if (row_t <= 2) {
        if (ef_t <= 4) return launch_search_full<CH, 2, 4, B, RT>(a, grid, st, win);
        if (ef_t <= 8) return launch_search_full<CH, 2, 8, B, RT>(a, grid, st, win);
        if (ef_t <= 16) return launch_search_full<CH, 2, 16, B, RT>(a, grid, st, win);
        return launch_search_full<CH, 2, 32, B, RT>(a, grid, st, win);
    }
    if (ef_t <= 4) return launch_search_full<CH, 4, 4, B, RT>(a, grid, st, win);
    if (ef_t <= 16) return launch_search_full<CH, 4, 16, B, RT>(a, grid, st, win);
    return launch_search_full<CH, 4, 32, B, RT>(a, grid, st, win);
}
template <int CH, int B>
cudaError_t dispatch_row_ef(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st, const LaunchWindow& win) {
    // bf16 rows stay packed while in flight (half the registers per row): twice the rows in flight per lane
    if (a.g.bf16) return dispatch_row_ef_rt<CH, (2 * B <= 16 ? 2 * B : B), RowBF16>(a, row_t, ef_t, grid, st, win);
    return dispatch_row_ef_rt<CH, B, RowF32>(a, row_t, ef_t, grid, st, win);
}

}  // namespace idb
