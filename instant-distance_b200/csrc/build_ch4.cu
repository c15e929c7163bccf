// Construction kernels (KA insert search, K2 select/relink) for rows of up to 512 floats.
#include "build_dispatch.cuh"
namespace idb {
cudaError_t build_dispatch_ch4(const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) { return build_dispatch<4, 4, 4>(a, l, st); }
}  // namespace idb
