// Rasterizer, per-gaussian stages (compiled with --fmad=false so that the index-defining arithmetic -- radii, tile
// rectangles, depth sort keys -- is bit-identical to oracle/raster_oracle.py): batched-over-cameras preprocess, key
// duplication, tile-range identification, and the per-gaussian backward.  All cameras of a call go through ONE launch of
// each kernel (the reference loops over cameras in Python: gaussian_batch_renderer_4d.py:27).
#include "a3d_raster_ws.cuh"

namespace a3d {

__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
  b[0] = kShC0;
  if (deg > 0) {
    b[1] = -0.4886025119029199f * y; b[2] = 0.4886025119029199f * z; b[3] = -0.4886025119029199f * x;
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      b[4] = 1.0925484305920792f * xy; b[5] = -1.0925484305920792f * yz; b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
      b[7] = -1.0925484305920792f * xz; b[8] = 0.5462742152960396f * (xx - yy);
      if (deg > 2) {
        b[9] = -0.5900435899266435f * y * (3.0f * xx - yy); b[10] = 2.890611442640554f * xy * z;
        b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy); b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy); b[14] = 1.445305721320277f * z * (xx - yy);
        b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
      }
    }
  }
}

__global__ void raster_preprocess_kernel(RasterDev a, RasterWs ws, int32_t* __restrict__ radii) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.num_cams * a.P) return;
  const int cam = idx / a.P, i = idx % a.P;
  const int g = a.per_cam_geometry ? cam : 0;
  const a3d_raster_cam& c = a.cams[cam];
  PreGauss o;
  const size_t gi = (size_t)g * a.P + i;
  const bool ok = preprocess_gaussian(a.means3D + 3 * gi, a.scales + 3 * gi, a.rotations + 4 * gi, a.scale_modifier, c.viewmatrix,
                                      c.projmatrix, c.tanfovx, c.tanfovy, a.H, a.W, o);
  radii[idx] = o.radius;
  ws.tiles[idx] = ok ? (uint32_t)o.tiles : 0u;
  if (!ok) return;
  ws.depth[idx] = o.depth;
  ws.xy[idx] = make_float2(o.px, o.py);
  ws.conic_opac[idx] = make_float4(o.conA, o.conB, o.conC, a.opacities[i]);
  ws.rect[idx] = make_int4(o.rx0, o.ry0, o.rx1, o.ry1);
  float rgb[3];
  uint32_t clamped = 0;
  if (a.colors_precomp) {
    for (int ch = 0; ch < 3; ++ch) rgb[ch] = a.colors_precomp[3 * i + ch];
  } else {
    float dx = a.means3D[3 * gi] - c.campos[0], dy = a.means3D[3 * gi + 1] - c.campos[1], dz = a.means3D[3 * gi + 2] - c.campos[2];
    const float n = sqrtf((dx * dx + dy * dy) + dz * dz);
    dx /= n; dy /= n; dz /= n;
    float b[16];
    sh_basis(a.sh_degree, dx, dy, dz, b);
    const int nb = (a.sh_degree + 1) * (a.sh_degree + 1);
    const float* sh = a.shs + (size_t)i * a.sh_coeffs * 3;
    for (int ch = 0; ch < 3; ++ch) {
      float v = 0.f;
      for (int k = 0; k < nb; ++k) v += b[k] * sh[3 * k + ch];
      v += 0.5f;
      if (v < 0.f) { clamped |= 1u << ch; v = 0.f; }
      rgb[ch] = v;
    }
  }
  ws.rgb_depth[idx] = make_float4(rgb[0], rgb[1], rgb[2], o.depth);
  ws.clamped[idx] = (uint8_t)clamped;
}

__global__ void raster_counts_kernel(RasterWs ws, int P, int cams, long long cap) {
  const int cam = threadIdx.x;
  if (cam < cams) {
    const uint32_t hi = ws.offsets[(size_t)(cam + 1) * P - 1];
    const uint32_t lo = cam ? ws.offsets[(size_t)cam * P - 1] : 0u;
    ws.counters[cam] = (long long)(hi - lo);
  }
  if (cam == 0) {
    const long long total = ws.offsets[(size_t)cams * P - 1];
    ws.counters[cams] = total;
    ws.counters[cams + 1] = total > cap ? 1 : 0;
  }
}

__global__ void raster_duplicate_kernel(RasterDev a, RasterWs ws, int gx, int num_tiles, long long cap) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.num_cams * a.P) return;
  if (ws.tiles[idx] == 0) return;
  const int cam = idx / a.P, i = idx % a.P;
  long long off = idx ? ws.offsets[idx - 1] : 0;
  const int4 r = ws.rect[idx];
  const uint32_t dbits = __float_as_uint(ws.depth[idx]);
  for (int y = r.y; y < r.w; ++y)
    for (int x = r.x; x < r.z; ++x) {
      if (off < cap) {
        const uint64_t tile = (uint64_t)cam * num_tiles + (uint64_t)(y * gx + x);
        ws.keys_a[off] = (tile << 31) | dbits;     // depth > 0.2: the sign bit is always 0 and is not sorted (one radix pass less)
        ws.vals_a[off] = (uint32_t)i;
      }
      ++off;
    }
}

__global__ void raster_ranges_kernel(RasterWs ws, long long n, long long total_tiles) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint64_t tile = ws.keys_b[idx] >> 31;
  if (tile >= (uint64_t)total_tiles) return;   // padding key
  if (idx == 0 || (ws.keys_b[idx - 1] >> 31) != tile) ws.ranges[tile].x = (uint32_t)idx;
  if (idx == n - 1 || (ws.keys_b[idx + 1] >> 31) != tile) ws.ranges[tile].y = (uint32_t)(idx + 1);
}

// per-gaussian backward: (dconic, dmean2D, ddepth, drgb) of every camera -> means3D / scales / rotations / colors / SH
__global__ void raster_preprocess_backward_kernel(RasterDev a, RasterWs ws, const int32_t* __restrict__ radii,
                                                  float* __restrict__ dmeans3D, float* __restrict__ dscales,
                                                  float* __restrict__ drots, float* __restrict__ dcolors,
                                                  float* __restrict__ dshs, float* __restrict__ dmeans2D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.num_cams * a.P) return;
  if (radii[idx] <= 0) return;
  const int cam = idx / a.P, i = idx % a.P;
  const int g = a.per_cam_geometry ? cam : 0;
  const size_t gi = (size_t)g * a.P + i;
  const a3d_raster_cam& c = a.cams[cam];
  const float gpx = ws.g_mean2d[2 * (size_t)idx], gpy = ws.g_mean2d[2 * (size_t)idx + 1];
  float dm[3], ds[3], dr[4];
  preprocess_backward(a.means3D + 3 * gi, a.scales + 3 * gi, a.rotations + 4 * gi, a.scale_modifier, c.viewmatrix, c.projmatrix,
                      c.tanfovx, c.tanfovy, a.H, a.W, ws.g_conic[3 * (size_t)idx], ws.g_conic[3 * (size_t)idx + 1],
                      ws.g_conic[3 * (size_t)idx + 2], gpx, gpy, ws.g_depth[idx], dm, ds, dr);
  for (int k = 0; k < 3; ++k) {
    if (dmeans3D) atomicAdd(dmeans3D + 3 * gi + k, dm[k]);
    if (dscales) atomicAdd(dscales + 3 * gi + k, ds[k]);
  }
  if (drots) for (int k = 0; k < 4; ++k) atomicAdd(drots + 4 * gi + k, dr[k]);
  if (dmeans2D) {   // upstream convention: gradient w.r.t. NDC coordinates
    dmeans2D[3 * (size_t)idx] = gpx * 0.5f * (float)a.W;
    dmeans2D[3 * (size_t)idx + 1] = gpy * 0.5f * (float)a.H;
    dmeans2D[3 * (size_t)idx + 2] = 0.f;
  }
  const float* grgb = ws.g_rgb + 3 * (size_t)idx;
  if (a.colors_precomp) {
    if (dcolors) for (int ch = 0; ch < 3; ++ch) atomicAdd(dcolors + 3 * i + ch, grgb[ch]);
  } else if (dshs) {
    // colour = clamp(sum_k basis_k(dir) * sh_k + 0.5): linear in the coefficients.  (The dependence of the basis on the
    // view direction is not differentiated: Animate3D's features are frozen buffers, gaussian_4d.py:262-297.)
    float dx = a.means3D[3 * gi] - c.campos[0], dy = a.means3D[3 * gi + 1] - c.campos[1], dz = a.means3D[3 * gi + 2] - c.campos[2];
    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= n; dy /= n; dz /= n;
    float b[16];
    sh_basis(a.sh_degree, dx, dy, dz, b);
    const int nb = (a.sh_degree + 1) * (a.sh_degree + 1);
    const uint32_t cl = ws.clamped[idx];
    for (int ch = 0; ch < 3; ++ch) {
      if (cl & (1u << ch)) continue;
      for (int k = 0; k < nb; ++k) atomicAdd(dshs + ((size_t)i * a.sh_coeffs + k) * 3 + ch, b[k] * grgb[ch]);
    }
  }
}

void launch_preprocess(const RasterDev& a, const RasterWs& ws, int32_t* radii, cudaStream_t st) {
  const int n = a.num_cams * a.P;
  raster_preprocess_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, ws, radii);
}
void launch_counts(const RasterWs& ws, int P, int cams, long long cap, cudaStream_t st) {
  raster_counts_kernel<<<1, ((cams + 31) / 32) * 32, 0, st>>>(ws, P, cams, cap);
}
void launch_duplicate(const RasterDev& a, const RasterWs& ws, int gx, int num_tiles, long long cap, cudaStream_t st) {
  const int n = a.num_cams * a.P;
  raster_duplicate_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, ws, gx, num_tiles, cap);
}
void launch_ranges(const RasterWs& ws, long long n, long long total_tiles, cudaStream_t st) {
  raster_ranges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, n, total_tiles);
}
void launch_preprocess_backward(const RasterDev& a, const RasterWs& ws, const int32_t* radii, float* dmeans3D, float* dscales,
                                float* drots, float* dcolors, float* dshs, float* dmeans2D, cudaStream_t st) {
  const int n = a.num_cams * a.P;
  raster_preprocess_backward_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, ws, radii, dmeans3D, dscales, drots, dcolors, dshs, dmeans2D);
}

}  // namespace a3d
