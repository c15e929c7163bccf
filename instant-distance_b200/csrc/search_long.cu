// K1 instantiation for rows of more than 1024 elements (any dim): the query lives in shared memory, distances run over groups
// of 32 chunks with 8 rows in flight per lane (hnsw_device.cuh batch_distances_long).
#include "search_kernel.cuh"
namespace idb {
cudaError_t dispatch_search_long(const SearchArgs& a, int row_t, int ef_t, int grid, cudaStream_t st, const LaunchWindow& win) {
    return dispatch_row_ef<0, kLongRowsInFlight>(a, row_t, ef_t, grid, st, win);
}
}  // namespace idb
