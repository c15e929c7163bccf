// Construction kernels (KA insert search, K2 select/relink) for rows of more than 1024 elements (any dim).
#include "build_dispatch.cuh"
namespace idb {
cudaError_t build_dispatch_long(const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) { return build_dispatch<0, kLongRowsInFlight, kLongRowsInFlight>(a, l, st); }
}  // namespace idb
