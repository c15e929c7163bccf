// K1 instantiation for rows of up to 512 floats (4 float4 chunk(s) per lane, 4 row loads in flight per lane).
#include "search_kernel.cuh"
namespace idb {
cudaError_t dispatch_search_ch4(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st, const LaunchWindow& win) {
    return dispatch_row_ef<4, 4>(a, row_t, ef_t, grid, st, win);
}
}  // namespace idb
