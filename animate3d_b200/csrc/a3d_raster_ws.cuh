// Shared declarations of the rasterizer translation units: device argument block, workspace carving, launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/a3d.h"
#include "a3d_raster_math.h"

namespace a3d {

struct RasterDev {
  int P, H, W, num_cams;
  const a3d_raster_cam* cams;
  const float *means3D, *scales, *rotations, *opacities, *shs, *colors_precomp;
  int sh_degree, sh_coeffs, per_cam_geometry;
  float scale_modifier;
  float bg[3];
};

struct RasterWs {
  // per (camera, gaussian)
  float* depth;
  float2* xy;
  float4* conic_opac;
  float4* rgb_depth;
  int4* rect;
  uint32_t* tiles;
  uint32_t* offsets;
  uint8_t* clamped;
  // binning
  uint64_t *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  uint2* ranges;
  // per pixel
  uint32_t* n_contrib;
  float* final_T;
  long long* counters;   // [cams] pair counts, [cams] total, [cams+1] overflow flag
  // backward scratch, per (camera, gaussian)
  float *g_mean2d, *g_conic, *g_depth, *g_rgb;
  void* cub_temp;
  size_t cub_bytes;
};

// Carves `base` (may be null for a size query) and returns the bytes needed.
inline size_t carve_workspace(void* base, int P, int H, int W, int cams, long long cap, size_t cub_bytes, RasterWs* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += bytes;
    return p;
  };
  const size_t n = (size_t)cams * P;
  const int gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;
  RasterWs w;
  w.depth = (float*)take(n * 4);
  w.xy = (float2*)take(n * 8);
  w.conic_opac = (float4*)take(n * 16);
  w.rgb_depth = (float4*)take(n * 16);
  w.rect = (int4*)take(n * 16);
  w.tiles = (uint32_t*)take(n * 4);
  w.offsets = (uint32_t*)take(n * 4);
  w.clamped = (uint8_t*)take(n);
  w.keys_a = (uint64_t*)take((size_t)cap * 8);
  w.keys_b = (uint64_t*)take((size_t)cap * 8);
  w.vals_a = (uint32_t*)take((size_t)cap * 4);
  w.vals_b = (uint32_t*)take((size_t)cap * 4);
  w.ranges = (uint2*)take((size_t)cams * gx * gy * 8);
  w.n_contrib = (uint32_t*)take((size_t)cams * H * W * 4);
  w.final_T = (float*)take((size_t)cams * H * W * 4);
  w.counters = (long long*)take((size_t)(cams + 2) * 8);
  w.g_mean2d = (float*)take(n * 8);
  w.g_conic = (float*)take(n * 12);
  w.g_depth = (float*)take(n * 4);
  w.g_rgb = (float*)take(n * 12);
  w.cub_temp = take(cub_bytes);
  w.cub_bytes = cub_bytes;
  if (ws) *ws = w;
  return (off + 255) & ~(size_t)255;
}

void launch_preprocess(const RasterDev& a, const RasterWs& ws, int32_t* radii, cudaStream_t st);
void launch_counts(const RasterWs& ws, int P, int cams, long long cap, cudaStream_t st);
void launch_duplicate(const RasterDev& a, const RasterWs& ws, int gx, int num_tiles, long long cap, cudaStream_t st);
void launch_ranges(const RasterWs& ws, long long n, long long total_tiles, cudaStream_t st);
void launch_preprocess_backward(const RasterDev& a, const RasterWs& ws, const int32_t* radii, float* dmeans3D, float* dscales,
                                float* drots, float* dcolors, float* dshs, float* dmeans2D, cudaStream_t st);

}  // namespace a3d
