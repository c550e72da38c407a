// Per-frame deformation field of the 4D gaussians, all frames of a batch in one launch:
//   k-planes (HexPlane) lookup -- 2 scales x 6 planes, bilinear, border clamp, align_corners -- product over planes, concat
//   over scales -> 32 features -> three bias-free MLPs 32 -> 32 (ReLU) -> {3, 4, 3} -> xyz + d, exp(s + d), normalize(q + d)
// forward and backward (gradients to the plane grids and MLP weights; the static gaussians are frozen buffers in
// Animate3D, gaussian_4d.py:262-297).  The `use_global_trans` branch (gaussian_4d.py:499-511, 525-539) needs the MEAN feature
// vector of a frame before any output can be formed: kMode 2 computes those means (one extra pass over the L2-resident
// planes), the host evaluates the two 32-wide global MLPs + Euler matrix on [T, 32], and the main pass takes the rotated
// base quaternions as `rot_base`; backward returns d/d rot_base and folds d/d mean-feature back into the plane gradients.  Replaces, per (frame, gaussian), the 12 grid_sample + 6 GEMM + elementwise launches the
// reference issues PER CAMERA (custom/threestudio-animate3d/geometry/gaussian_4d.py:39-64, 450-548;
// renderer/diff_gaussian_rasterizer_advanced_4d.py:77-83, 119-135) and de-duplicates the 4 views of a frame.
#include "a3d_host.cuh"

namespace a3d {

constexpr int kPlanes = 6;
constexpr int kMaxScales = 2;
constexpr int kFeat = 16;                  // channels per plane
constexpr int kHid = 32;                   // MLP width == feature width (2 scales x 16)

struct DeformGrids {
  const float* plane[kMaxScales][kPlanes]; // [C, H, W]
  float* gplane[kMaxScales][kPlanes];      // gradients (backward) or null
  int h[kMaxScales][kPlanes], w[kMaxScales][kPlanes];
  int scales;
};

struct DeformMlp {          // row-major [out, in]
  const float *w1[3], *w2[3];              // 0: xyz (3), 1: rot (4), 2: scale (3)
  float *gw1[3], *gw2[3];
};

__device__ __constant__ int c_comb[kPlanes][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
__device__ __forceinline__ int out_dim(int m) { return m == 1 ? 4 : 3; }

struct Bilerp {
  int x0, x1, y0, y1;
  float wx, wy;
};
// grid_sample(align_corners=True, padding_mode="border") coordinates for normalised (gx, gy): gx indexes W, gy indexes H
__device__ __forceinline__ Bilerp bilerp_setup(float gx, float gy, int W, int H) {
  float fx = (gx + 1.f) * 0.5f * (float)(W - 1);
  float fy = (gy + 1.f) * 0.5f * (float)(H - 1);
  fx = fminf(fmaxf(fx, 0.f), (float)(W - 1));
  fy = fminf(fmaxf(fy, 0.f), (float)(H - 1));
  Bilerp b;
  b.x0 = (int)floorf(fx); b.y0 = (int)floorf(fy);
  b.x1 = min(b.x0 + 1, W - 1); b.y1 = min(b.y0 + 1, H - 1);
  b.wx = fx - (float)b.x0; b.wy = fy - (float)b.y0;
  return b;
}

struct DeformGlobal {
  const float* rot_base;      // [T,P,4] per-frame base quaternions (replace `rotation`) or null
  float* g_rot_base;          // backward: d/d rot_base, [T,P,4], or null
  const float* g_featmean;    // backward: d/d (mean feature of frame t), [T, nfeat], or null
  float* featmean;            // kMode 2: [T, nfeat], pre-zeroed; receives sum / P
};

// kMode: 0 = forward, 1 = backward, 2 = per-frame mean of the k-planes features only
template <int kMode>
__global__ void __launch_bounds__(128)
deform_kernel(DeformGrids G, DeformMlp M, DeformGlobal X, const float* __restrict__ xyz, const float* __restrict__ scaling,
              const float* __restrict__ rotation, const float* __restrict__ times, int P, int T, int deform_scale,
              float* __restrict__ out_means, float* __restrict__ out_scales, float* __restrict__ out_rots,
              const float* __restrict__ g_means, const float* __restrict__ g_scales, const float* __restrict__ g_rots) {
  constexpr bool kBackward = kMode == 1;
  extern __shared__ float sm[];
  // shared copies of the MLP weights (and, in backward, of their gradient accumulators)
  float* sw1 = sm;                                   // [3][32][32]
  float* sw2 = sw1 + 3 * kHid * kHid;                // [3][4][32] (padded to 4 outputs)
  float* sg1 = sw2 + 3 * 4 * kHid;                   // backward only
  float* sg2 = sg1 + 3 * kHid * kHid;
  const int nfeat = G.scales * kFeat;
  if (kMode != 2) {
    for (int i = threadIdx.x; i < 3 * kHid * kHid; i += blockDim.x) {
      const int m = i / (kHid * kHid), r = i % (kHid * kHid);
      sw1[i] = (r % kHid < nfeat) ? M.w1[m][(r / kHid) * nfeat + r % kHid] : 0.f;
      if (kBackward) sg1[i] = 0.f;
    }
    for (int i = threadIdx.x; i < 3 * 4 * kHid; i += blockDim.x) {
      const int m = i / (4 * kHid), r = (i % (4 * kHid)) / kHid, c = i % kHid;
      sw2[i] = r < out_dim(m) ? M.w2[m][r * kHid + c] : 0.f;
      if (kBackward) sg2[i] = 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < 2 * kHid; i += blockDim.x) sm[i] = 0.f;   // [2 frames a block may straddle][32]
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < P * T;
  const int t = active ? idx / P : 0, i = active ? idx % P : 0;
  float pt[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) { pt[0] = xyz[3 * i]; pt[1] = xyz[3 * i + 1]; pt[2] = xyz[3 * i + 2]; pt[3] = times[t]; }
  // ---- features: product over planes of bilinear samples
  float feat[kHid];
#pragma unroll
  for (int c = 0; c < kHid; ++c) feat[c] = c < nfeat ? 1.f : 0.f;
  if (active) {
    for (int s = 0; s < G.scales; ++s)
      for (int pl = 0; pl < kPlanes; ++pl) {
        const int W = G.w[s][pl], H = G.h[s][pl];
        const Bilerp b = bilerp_setup(pt[c_comb[pl][0]], pt[c_comb[pl][1]], W, H);
        const float* g = G.plane[s][pl];
#pragma unroll
        for (int c = 0; c < kFeat; ++c) {
          const float* gc = g + (size_t)c * H * W;
          const float v00 = __ldg(gc + b.y0 * W + b.x0), v01 = __ldg(gc + b.y0 * W + b.x1);
          const float v10 = __ldg(gc + b.y1 * W + b.x0), v11 = __ldg(gc + b.y1 * W + b.x1);
          const float v = (v00 * (1.f - b.wx) + v01 * b.wx) * (1.f - b.wy) + (v10 * (1.f - b.wx) + v11 * b.wx) * b.wy;
          feat[s * kFeat + c] *= v;
        }
      }
  }
  if (kMode == 2) {
    // block-level sum per frame (a 128-thread block touches at most two frames when P >= 128), then one atomic per value
    const int t_first = (blockIdx.x * blockDim.x) / P;
    if (active) {
      const int slot = t - t_first;
      if (slot < 2) {
#pragma unroll
        for (int c = 0; c < kHid; ++c) if (c < nfeat) atomicAdd(&sm[slot * kHid + c], feat[c]);
      } else {
#pragma unroll
        for (int c = 0; c < kHid; ++c) if (c < nfeat) atomicAdd(&X.featmean[(size_t)t * nfeat + c], feat[c] / (float)P);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * kHid; k += blockDim.x) {
      const int slot = k / kHid, c = k % kHid;
      if (c < nfeat && t_first + slot < T && sm[k] != 0.f) atomicAdd(&X.featmean[(size_t)(t_first + slot) * nfeat + c], sm[k] / (float)P);
    }
    return;
  }
  // ---- three MLPs
  float hid[3][kHid];
  float outv[3][4];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
#pragma unroll
    for (int r = 0; r < kHid; ++r) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < kHid; ++c) a = fmaf(sw1[(m * kHid + r) * kHid + c], feat[c], a);
      hid[m][r] = fmaxf(a, 0.f);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < kHid; ++c) a = fmaf(sw2[(m * 4 + r) * kHid + c], hid[m][c], a);
      outv[m][r] = a;
    }
  }
  float q[4] = {0.f, 0.f, 0.f, 1.f}, sc[3] = {0.f, 0.f, 0.f};
  float qn = 1.f;
  if (active) {
    const float* qb = X.rot_base ? X.rot_base + 4 * ((size_t)t * P + i) : rotation + 4 * i;
    for (int k = 0; k < 4; ++k) q[k] = qb[k] + outv[1][k];
    qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
    for (int k = 0; k < 3; ++k) sc[k] = expf(scaling[3 * i + k] + (deform_scale ? outv[2][k] : 0.f));
  }
  if (!kBackward) {
    if (active) {
      const size_t o = (size_t)t * P + i;
      for (int k = 0; k < 3; ++k) out_means[3 * o + k] = pt[k] + outv[0][k];
      for (int k = 0; k < 3; ++k) out_scales[3 * o + k] = sc[k];
      for (int k = 0; k < 4; ++k) out_rots[4 * o + k] = q[k] / qn;
    }
    return;
  }
  // ---- backward
  float dout[3][4];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int r = 0; r < 4; ++r) dout[m][r] = 0.f;
  if (active) {
    const size_t o = (size_t)t * P + i;
    for (int k = 0; k < 3; ++k) dout[0][k] = g_means ? g_means[3 * o + k] : 0.f;
    if (deform_scale && g_scales) for (int k = 0; k < 3; ++k) dout[2][k] = g_scales[3 * o + k] * sc[k];     // d exp
    if (g_rots) {   // y = q / |q|: dq = (g - y (y.g)) / |q|
      float y[4], dot = 0.f;
      for (int k = 0; k < 4; ++k) { y[k] = q[k] / qn; dot += y[k] * g_rots[4 * o + k]; }
      for (int k = 0; k < 4; ++k) dout[1][k] = (g_rots[4 * o + k] - y[k] * dot) / qn;
    }
    if (X.g_rot_base) for (int k = 0; k < 4; ++k) X.g_rot_base[4 * o + k] = dout[1][k];   // q = rot_base + delta
  }
  float dfeat[kHid];
#pragma unroll
  for (int c = 0; c < kHid; ++c) dfeat[c] = 0.f;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    float dh[kHid];
#pragma unroll
    for (int c = 0; c < kHid; ++c) {
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) a = fmaf(sw2[(m * 4 + r) * kHid + c], dout[m][r], a);
      dh[c] = hid[m][c] > 0.f ? a : 0.f;
    }
    if (active) {
      // weight gradients: accumulate in shared memory, one global atomic per weight per block at the end
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (dout[m][r] != 0.f) {
#pragma unroll
          for (int c = 0; c < kHid; ++c) atomicAdd(&sg2[(m * 4 + r) * kHid + c], dout[m][r] * hid[m][c]);
        }
      }
#pragma unroll
      for (int r = 0; r < kHid; ++r) {
        if (dh[r] != 0.f) {
#pragma unroll
          for (int c = 0; c < kHid; ++c) atomicAdd(&sg1[(m * kHid + r) * kHid + c], dh[r] * feat[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kHid; ++c) {
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < kHid; ++r) a = fmaf(sw1[(m * kHid + r) * kHid + c], dh[r], a);
      dfeat[c] += a;
    }
  }
  if (active && X.g_featmean) {
#pragma unroll
    for (int c = 0; c < kHid; ++c) if (c < nfeat) dfeat[c] += X.g_featmean[(size_t)t * nfeat + c] / (float)P;
  }
  if (active) {
    // d sample_p = dfeat * prod_{q != p} sample_q = dfeat * feat / sample_p  (recompute samples; guard tiny values)
    for (int s = 0; s < G.scales; ++s)
      for (int pl = 0; pl < kPlanes; ++pl) {
        float* gg = G.gplane[s][pl];
        if (!gg) continue;
        const int W = G.w[s][pl], H = G.h[s][pl];
        const Bilerp b = bilerp_setup(pt[c_comb[pl][0]], pt[c_comb[pl][1]], W, H);
        const float* g = G.plane[s][pl];
        for (int c = 0; c < kFeat; ++c) {
          // product of the OTHER planes, recomputed exactly (no division)
          float other = 1.f;
          for (int p2 = 0; p2 < kPlanes; ++p2) {
            if (p2 == pl) continue;
            const int W2 = G.w[s][p2], H2 = G.h[s][p2];
            const Bilerp b2 = bilerp_setup(pt[c_comb[p2][0]], pt[c_comb[p2][1]], W2, H2);
            const float* gc = G.plane[s][p2] + (size_t)c * H2 * W2;
            other *= (__ldg(gc + b2.y0 * W2 + b2.x0) * (1.f - b2.wx) + __ldg(gc + b2.y0 * W2 + b2.x1) * b2.wx) * (1.f - b2.wy) +
                     (__ldg(gc + b2.y1 * W2 + b2.x0) * (1.f - b2.wx) + __ldg(gc + b2.y1 * W2 + b2.x1) * b2.wx) * b2.wy;
          }
          const float ds = dfeat[s * kFeat + c] * other;
          if (ds == 0.f) continue;
          float* gc = gg + (size_t)c * H * W;
          atomicAdd(gc + b.y0 * W + b.x0, ds * (1.f - b.wx) * (1.f - b.wy));
          atomicAdd(gc + b.y0 * W + b.x1, ds * b.wx * (1.f - b.wy));
          atomicAdd(gc + b.y1 * W + b.x0, ds * (1.f - b.wx) * b.wy);
          atomicAdd(gc + b.y1 * W + b.x1, ds * b.wx * b.wy);
        }
        (void)g;
      }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 3 * kHid * kHid; k += blockDim.x) {
    const int m = k / (kHid * kHid), r = k % (kHid * kHid);
    if (r % kHid < nfeat && M.gw1[m] && sg1[k] != 0.f) atomicAdd(&M.gw1[m][(r / kHid) * nfeat + r % kHid], sg1[k]);
  }
  for (int k = threadIdx.x; k < 3 * 4 * kHid; k += blockDim.x) {
    const int m = k / (4 * kHid), r = (k % (4 * kHid)) / kHid, c = k % kHid;
    if (r < out_dim(m) && M.gw2[m] && sg2[k] != 0.f) atomicAdd(&M.gw2[m][r * kHid + c], sg2[k]);
  }
}

}  // namespace a3d

using namespace a3d;

static DeformGlobal globals(const a3d_deform_args* a, float* featmean) {
  DeformGlobal X;
  X.rot_base = a->rot_base; X.g_rot_base = a->grad_rot_base; X.g_featmean = a->grad_featmean; X.featmean = featmean;
  return X;
}

static int fill(const a3d_deform_args* a, DeformGrids* G, DeformMlp* M) {
  const bool bwd = a && a->grad_w1[0] != nullptr;
  (void)bwd;
  if (!a || a->P <= 0 || a->T <= 0) return fail(A3D_EINVAL, "a3d_deform: bad sizes");
  if (a->num_scales < 1 || a->num_scales > kMaxScales || a->channels != kFeat || a->hidden != kHid)
    return fail(A3D_EINVAL, "a3d_deform: supports 1-2 scales x 16 channels and 32-wide MLPs (got %d x %d, hidden %d)", a->num_scales,
                a->channels, a->hidden);
  memset(G, 0, sizeof(*G));
  memset(M, 0, sizeof(*M));
  G->scales = a->num_scales;
  for (int s = 0; s < a->num_scales; ++s)
    for (int p = 0; p < kPlanes; ++p) {
      G->plane[s][p] = a->planes[s * kPlanes + p];
      G->gplane[s][p] = a->grad_planes[s * kPlanes + p];
      G->h[s][p] = a->plane_h[s * kPlanes + p];
      G->w[s][p] = a->plane_w[s * kPlanes + p];
      if (!G->plane[s][p] || G->h[s][p] < 1 || G->w[s][p] < 1) return fail(A3D_EINVAL, "a3d_deform: bad plane %d/%d", s, p);
    }
  for (int m = 0; m < 3; ++m) {
    M->w1[m] = a->w1[m]; M->w2[m] = a->w2[m];
    M->gw1[m] = a->grad_w1[m];
    M->gw2[m] = a->grad_w2[m];
    if (!M->w1[m] || !M->w2[m]) return fail(A3D_EINVAL, "a3d_deform: null MLP weights");
  }
  return 0;
}

extern "C" int a3d_deform_forward(const a3d_deform_args* a, float* means, float* scales, float* rots, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DeformGrids G; DeformMlp M;
  if (int r = fill(a, &G, &M)) return r;
  const int n = a->P * a->T;
  const size_t smem = (3 * kHid * kHid + 3 * 4 * kHid) * sizeof(float);
  deform_kernel<0><<<(n + 127) / 128, 128, smem, st>>>(G, M, globals(a, nullptr), a->xyz, a->scaling, a->rotation, a->times, a->P, a->T,
                                                      a->deform_scale, means, scales, rots, nullptr, nullptr, nullptr);
  A3D_LAUNCH_CHECK();
  return A3D_OK;
}

extern "C" int a3d_deform_featmean(const a3d_deform_args* a, float* featmean, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DeformGrids G; DeformMlp M;
  if (int r = fill(a, &G, &M)) return r;
  if (!featmean) return fail(A3D_EINVAL, "a3d_deform_featmean: null output");
  if (a->P < 128) return fail(A3D_EINVAL, "a3d_deform_featmean: needs at least 128 gaussians (got %d)", a->P);
  const int n = a->P * a->T;
  A3D_CUDA_CHECK(cudaMemsetAsync(featmean, 0, sizeof(float) * a->T * a->num_scales * kFeat, st));
  const size_t smem = (3 * kHid * kHid + 3 * 4 * kHid) * sizeof(float);
  deform_kernel<2><<<(n + 127) / 128, 128, smem, st>>>(G, M, globals(a, featmean), a->xyz, a->scaling, a->rotation, a->times, a->P, a->T,
                                                      a->deform_scale, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  A3D_LAUNCH_CHECK();
  return A3D_OK;
}

extern "C" int a3d_deform_backward(const a3d_deform_args* a, const float* g_means, const float* g_scales, const float* g_rots,
                                   void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DeformGrids G; DeformMlp M;
  if (int r = fill(a, &G, &M)) return r;
  const int n = a->P * a->T;
  const size_t smem = 2 * (3 * kHid * kHid + 3 * 4 * kHid) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    A3D_CUDA_CHECK(cudaFuncSetAttribute(deform_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  deform_kernel<1><<<(n + 127) / 128, 128, smem, st>>>(G, M, globals(a, nullptr), a->xyz, a->scaling, a->rotation, a->times, a->P, a->T,
                                                      a->deform_scale, nullptr, nullptr, nullptr, g_means, g_scales, g_rots);
  A3D_LAUNCH_CHECK();
  return A3D_OK;
}
