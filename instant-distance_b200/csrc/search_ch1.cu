// K1 instantiation for rows of up to 128 floats (1 float4 chunk(s) per lane, 16 row loads in flight per lane).
#include "search_kernel.cuh"
namespace idb {
cudaError_t dispatch_search_ch1(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st) {
    return dispatch_row_ef<1, 16>(a, row_t, ef_t, grid, st);
}
}  // namespace idb
