// build_dispatch.cuh — launch helpers for the construction kernels, instantiated once per CH in build_chN.cu.
#pragma once
#include "build_kernels.cuh"

namespace idb {

enum BuildOp : int { kOpInsertSearch = 0, kOpSelectNew = 1, kOpRelink = 2, kOpRelinkSimple = 3 };

struct BuildLaunch {
    BuildOp op;
    int row_t, ef_t;        // KA template selectors
    bool stage;             // K2: kept rows staged in shared memory
    int grid;
    uint32_t smem_per_warp; // K2
    LaunchWindow win;       // KA: persisting-L2 window on the b16 visited tables
};

template <int CH, int ROW_T, int EF_T, int B, class RT>
cudaError_t launch_insert_search(const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) {
    const int smem = (WarpSmem<EF_T>::kBytes + (CH == 0 ? (int)long_q_bytes(a.g.nchunks) : 0)) * kSearchWarps;
    auto kern = insert_search_kernel<CH, ROW_T, EF_T, B, RT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return launch_with_window(kern, l.grid, kSearchWarps * 32, smem, st, l.win, a);
}

template <int CH, int NB, bool kStage, class RT>
cudaError_t launch_k2(const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) {
    const int smem = (int)l.smem_per_warp * kBuildWarps;
    if (l.op == kOpSelectNew) {
        auto kern = select_new_kernel<CH, NB, kStage, RT>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<l.grid, kBuildWarps * 32, smem, st>>>(a, l.smem_per_warp);
    } else {
        auto kern = relink_kernel<CH, NB, kStage, RT>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<l.grid, kBuildWarps * 32, smem, st>>>(a, l.smem_per_warp);
    }
    return cudaGetLastError();
}

template <int CH, int B, int NB, class RT>
cudaError_t build_dispatch_rt(const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) {
    switch (l.op) {
        case kOpInsertSearch:
            if (l.row_t <= 2) {
                if (l.ef_t <= 4) return launch_insert_search<CH, 2, 4, B, RT>(a, l, st);
                if (l.ef_t <= 8) return launch_insert_search<CH, 2, 8, B, RT>(a, l, st);
                if (l.ef_t <= 16) return launch_insert_search<CH, 2, 16, B, RT>(a, l, st);
                return launch_insert_search<CH, 2, 32, B, RT>(a, l, st);
            }
            if (l.ef_t <= 4) return launch_insert_search<CH, 4, 4, B, RT>(a, l, st);
            if (l.ef_t <= 16) return launch_insert_search<CH, 4, 16, B, RT>(a, l, st);
            return launch_insert_search<CH, 4, 32, B, RT>(a, l, st);
        case kOpSelectNew:
        case kOpRelink:
            if constexpr (CH > 0) {
                if (l.stage) return launch_k2<CH, NB, true, RT>(a, l, st);
            }
            return launch_k2<CH, NB, false, RT>(a, l, st);
        case kOpRelinkSimple: {
            const int smem = CH == 0 ? (int)long_q_bytes(a.g.nchunks) * kBuildWarps : 0;
            auto kern = relink_simple_kernel<CH, RT>;
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                if (e != cudaSuccess) return e;
            }
            kern<<<l.grid, kBuildWarps * 32, smem, st>>>(a);
            return cudaGetLastError();
        }
    }
    return cudaErrorInvalidValue;
}
template <int CH, int B, int NB>
cudaError_t build_dispatch(const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) {
    if (a.g.bf16) return build_dispatch_rt<CH, B, NB, RowBF16>(a, l, st);
    return build_dispatch_rt<CH, B, NB, RowF32>(a, l, st);
}

}  // namespace idb
