// Construction kernels (KA insert search, K2 select/relink) for rows of up to 768 floats.
#include "build_dispatch.cuh"
namespace idb {
cudaError_t build_dispatch_ch6(const BuildArgs& a, const BuildLaunch& l, cudaStream_t st) { return build_dispatch<6, 2, 2>(a, l, st); }
}  // namespace idb
