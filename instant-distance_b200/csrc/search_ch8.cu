// K1 instantiation for rows of up to 1024 floats (8 float4 chunk(s) per lane, 2 row loads in flight per lane).
#include "search_kernel.cuh"
namespace idb {
cudaError_t dispatch_search_ch8(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st, const LaunchWindow& win) {
    return dispatch_row_ef<8, 2>(a, row_t, ef_t, grid, st, win);
}
}  // namespace idb
