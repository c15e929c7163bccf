// K1 instantiation for rows of up to 128 floats (1 float4 chunk(s) per lane, 16 row loads in flight per lane).
#include "search_kernel.cuh"
namespace idb {
cudaError_t dispatch_search_ch1(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st, const LaunchWindow& win) {
    // tuning variants of the headline shape (ROW_T=2, EF_T=4): rows in flight per lane x resident CTAs per SM
    if (a.variant && row_t <= 2 && ef_t <= 4) {
        switch (a.variant) {
            case 1: return launch_search<1, 2, 4, 8, occ_for_warps(20)>(a, grid, st, win);
            case 2: return launch_search<1, 2, 4, 8, occ_for_warps(24)>(a, grid, st, win);
            case 3: return launch_search<1, 2, 4, 4, occ_for_warps(32)>(a, grid, st, win);
            case 4: return launch_search<1, 2, 4, 16, occ_for_warps(12)>(a, grid, st, win);
            // EXPERIMENT (profiles/r02_experiment_tma_ring.md): rows via cp.async.bulk into a shared-memory ring, no register staging
            case 5: if (!a.g.bf16) return launch_search<1, 2, 4, 16, occ_for_warps(16), RowF32, false, true>(a, grid, st, win); break;
            case 6: if (!a.g.bf16) return launch_search<1, 2, 4, 8, occ_for_warps(20), RowF32, false, true>(a, grid, st, win); break;
            case 7: if (!a.g.bf16) return launch_search<1, 2, 4, 8, occ_for_warps(24), RowF32, false, true>(a, grid, st, win); break;
            case 8: if (!a.g.bf16) return launch_search<1, 2, 4, 32, occ_for_warps(8), RowF32, false, true>(a, grid, st, win); break;
            default: break;
        }
    }
    return dispatch_row_ef<1, 16>(a, row_t, ef_t, grid, st, win);
}
}  // namespace idb
