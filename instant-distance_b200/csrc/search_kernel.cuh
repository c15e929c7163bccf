// search_kernel.cuh — K1: batched Hnsw::search kernel (lib.rs:352-383 per query) and its launch helpers.
// Instantiated once per CH (float4 chunks per lane) in search_chN.cu so the translation units build in parallel.
#pragma once
#include "internal.cuh"

namespace idb {

// ---------------------------------------------------------------------------------------------------------
// K1: batched Hnsw::search — persistent grid, one warp per live query, queries claimed from an atomic counter.
// ---------------------------------------------------------------------------------------------------------
// FULL: dim is a multiple of 128, i.e. every lane owns a real chunk in each of its CH slots: no chunk predicates, and
// full batches of row loads carry no predicates at all (hnsw_device.cuh batch_distances_impl).
template <int CH, int ROW_T, int EF_T, int B, int OCC, class RT = RowF32, bool FULL = false>
__global__ void __launch_bounds__(kSearchWarps * 32, OCC) search_kernel(SearchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t gwarp = blockIdx.x * kSearchWarps + warp;

    constexpr int kNearBytes = 2 * 32 * EF_T * 8;
    constexpr int kWarpBytes = kNearBytes + kSmallVisSlots * 4 + 128 * 4 + 128 * 8;
    unsigned char* base = smem_raw + (size_t)warp * kWarpBytes;

    WarpState s;
    s.near_base = reinterpret_cast<uint64_t*>(base);
    s.near_len = 32 * EF_T;
    s.vis.small = reinterpret_cast<uint32_t*>(base + kNearBytes);
    s.cpid = s.vis.small + kSmallVisSlots;
    s.ckey = reinterpret_cast<uint64_t*>(s.cpid + 128);
    s.vis.big = a.vis_tables + (size_t)gwarp * a.vis_stride;
    s.vis.gslots = a.gslots;
    s.vis.gshift = a.gshift;
    s.vis.mode = a.vis_mode;
    s.vis.count = 0;
    s.vis.use_big = false;
    s.ties = a.tie_tables + (size_t)gwarp * kTieCap;
    vis_clear_small(s.vis, lane);  // big tables are handed over clean by the host / previous launch
    const unsigned long long n_work = a.n_work_dev ? (unsigned long long)*a.n_work_dev : a.n_work;

    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1ull);
        w = __shfl_sync(kFullMask, w, 0);
        if (w >= n_work) break;
        const uint64_t qi = a.work_list ? a.work_list[w] : w;

        float4 q[CH];
        const float4* qrow = a.queries + qi * a.g.nchunks;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const uint32_t c = lane + 32 * j;
            q[j] = c < a.g.nchunks ? __ldg(qrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }

        descend<CH, ROW_T, EF_T, B, false, RT, FULL>(a.g, s, q, 0u, a.ef, lane, a.counters ? a.counters + qi * 4 : nullptr);

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
}

template <int CH, int ROW_T, int EF_T, int B, int OCC = kSearchCtasPerSm, class RT = RowF32, bool FULL = false>
static cudaError_t launch_search(const SearchArgs& a, int grid, cudaStream_t stream) {
    constexpr int kWarpBytes = 2 * 32 * EF_T * 8 + kSmallVisSlots * 4 + 128 * 4 + 128 * 8;
    const int smem = kWarpBytes * kSearchWarps;
    auto kern = search_kernel<CH, ROW_T, EF_T, B, OCC, RT, FULL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kSearchWarps * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int CH, int ROW_T, int EF_T, int B, class RT>
cudaError_t launch_search_full(const SearchArgs& a, int grid, cudaStream_t st) {
    if (a.g.nchunks == 32u * CH) return launch_search<CH, ROW_T, EF_T, B, kSearchCtasPerSm, RT, true>(a, grid, st);
    return launch_search<CH, ROW_T, EF_T, B, kSearchCtasPerSm, RT, false>(a, grid, st);
}
template <int CH, int B, class RT>
cudaError_t dispatch_row_ef_rt(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st) {
    if (row_t <= 2) {
        if (ef_t <= 4) return launch_search_full<CH, 2, 4, B, RT>(a, grid, st);
        return launch_search_full<CH, 2, 16, B, RT>(a, grid, st);
    }
    if (ef_t <= 4) return launch_search_full<CH, 4, 4, B, RT>(a, grid, st);
    return launch_search_full<CH, 4, 16, B, RT>(a, grid, st);
}
template <int CH, int B>
cudaError_t dispatch_row_ef(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st) {
    // bf16 rows stay packed while in flight (half the registers per row): twice the rows in flight per lane
    if (a.g.bf16) return dispatch_row_ef_rt<CH, (2 * B <= 16 ? 2 * B : B), RowBF16>(a, row_t, ef_t, grid, st);
    return dispatch_row_ef_rt<CH, B, RowF32>(a, row_t, ef_t, grid, st);
}

}  // namespace idb
