// placeholder until the rasterizer lands (replaced in the next commit)
#include "a3d_host.cuh"
using namespace a3d;
extern "C" size_t a3d_raster_workspace_bytes(int, int, int, int, int64_t) { return 0; }
extern "C" int a3d_raster_forward(const a3d_raster_args*, float*, float*, float*, int32_t*, void*, size_t, int64_t, int64_t*, void*) {
  return fail(A3D_EINVAL, "rasterizer not built yet");
}
extern "C" int a3d_raster_backward(const a3d_raster_args*, const float*, const float*, const float*, const float*, const int32_t*,
                                   void*, size_t, int64_t, float*, float*, float*, float*, float*, float*, float*, void*) {
  return fail(A3D_EINVAL, "rasterizer not built yet");
}
extern "C" int a3d_raster_binning_tap(const void*, int, int, int, int, int64_t, int, uint64_t*, uint32_t*, uint32_t*, void*) {
  return fail(A3D_EINVAL, "rasterizer not built yet");
}
