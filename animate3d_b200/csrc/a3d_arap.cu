// ARAP regulariser of the 4D-SDS / reconstruction step (custom/threestudio-animate3d/systems/util.py:58-117,
// 138-215, called from systems/animate3d.py:215-244): brute-force KNN graph of the control nodes (built once -- the
// gaussians' rest positions are frozen) and one fused kernel for all frames that fits the per-node rotation, evaluates the
// edge energy and scatters its gradient.  The arithmetic lives in a3d_arap_math.h (shared with the CPU harness that checks
// it against the reference's own functions).  STATUS: the math is validated on the CPU; this wrapper has not run on
// hardware yet (round-1 GPU budget exhausted) -- tests/test_arap_gpu.py runs it in a subprocess.
#include "a3d_host.cuh"
#include "a3d_arap_math.h"

namespace a3d {

// K+1 nearest (squared distance ascending, ties by index) of every point among all points; the first one (the point itself)
// is dropped -- pytorch3d.ops.knn_points(points, points, K=K+1)[..., 1:] as the reference uses it (util.py:79-82).
__global__ void __launch_bounds__(128) knn_kernel(const float* __restrict__ pts, int n, int K, int* __restrict__ nbr,
                                                  float* __restrict__ dist2) {
  __shared__ float sp[128 * 3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < n;
  const float px = ok ? pts[3 * i] : 0.f, py = ok ? pts[3 * i + 1] : 0.f, pz = ok ? pts[3 * i + 2] : 0.f;
  float bd[kArapMaxK + 1];
  int bi[kArapMaxK + 1];
  for (int k = 0; k <= K; ++k) { bd[k] = INFINITY; bi[k] = -1; }
  for (int base = 0; base < n; base += 128) {
    const int j = base + threadIdx.x;
    __syncthreads();
    if (j < n) { sp[3 * threadIdx.x] = pts[3 * j]; sp[3 * threadIdx.x + 1] = pts[3 * j + 1]; sp[3 * threadIdx.x + 2] = pts[3 * j + 2]; }
    __syncthreads();
    const int cnt = min(128, n - base);
    if (ok) {
      for (int c = 0; c < cnt; ++c) {
        const float dx = px - sp[3 * c], dy = py - sp[3 * c + 1], dz = pz - sp[3 * c + 2];
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < bd[K]) {                      // strict: an equal distance keeps the earlier index (stable sort order)
          int k = K;
          while (k > 0 && d < bd[k - 1]) { bd[k] = bd[k - 1]; bi[k] = bi[k - 1]; --k; }
          bd[k] = d; bi[k] = base + c;
        }
      }
    }
  }
  if (ok)
    for (int k = 0; k < K; ++k) { nbr[(int64_t)i * K + k] = bi[k + 1]; dist2[(int64_t)i * K + k] = bd[k + 1]; }
}

__global__ void __launch_bounds__(128) arap_kernel(const float* __restrict__ nodes, int Nt, int Nv, const int* __restrict__ nbr, int K,
                                                   const float* __restrict__ weight, const int* __restrict__ sample, int Ns,
                                                   float* __restrict__ err, float* __restrict__ grad) {
  __shared__ float red[4];
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)(Nt - 1) * Ns;
  float energy = 0.f;
  if (idx < total) {
    const int t = 1 + (int)(idx / Ns);
    const int si = (int)(idx % Ns);
    const int i = sample ? sample[si] : si;
    if (i >= 0 && i < Nv) {
      float e0[kArapMaxK][3], et[kArapMaxK][3], g0[kArapMaxK][3], gt[kArapMaxK][3], w[kArapMaxK];
      bool valid[kArapMaxK];
      int nb[kArapMaxK];
      const float* p0 = nodes;
      const float* pt = nodes + (int64_t)t * Nv * 3;
#pragma unroll
      for (int n = 0; n < kArapMaxK; ++n) {
        const int j = n < K ? nbr[(int64_t)i * K + n] : -1;
        nb[n] = j;
        valid[n] = j >= 0 && j < Nv;
        w[n] = (n < K) ? (weight ? weight[(int64_t)i * K + n] : (valid[n] ? 1.f : 0.f)) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          e0[n][c] = valid[n] ? p0[3 * (int64_t)i + c] - p0[3 * (int64_t)j + c] : 0.f;
          et[n][c] = valid[n] ? pt[3 * (int64_t)i + c] - pt[3 * (int64_t)j + c] : 0.f;
        }
      }
      energy = arap_node(K, e0, et, valid, w, gt, g0);
      if (grad) {
        float* gT = grad + (int64_t)t * Nv * 3;
#pragma unroll
        for (int n = 0; n < kArapMaxK; ++n) {
          if (n < K && valid[n]) {
            const int j = nb[n];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              atomicAdd(gT + 3 * (int64_t)i + c, gt[n][c]);
              atomicAdd(gT + 3 * (int64_t)j + c, -gt[n][c]);
              atomicAdd(grad + 3 * (int64_t)i + c, g0[n][c]);
              atomicAdd(grad + 3 * (int64_t)j + c, -g0[n][c]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) energy += __shfl_xor_sync(0xffffffffu, energy, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = energy;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(err, red[0] + red[1] + red[2] + red[3]);
}

}  // namespace a3d

using namespace a3d;

extern "C" int a3d_knn_graph(const float* points, int n, int K, int32_t* nbr, float* dist2, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!points || !nbr || !dist2 || n < 2 || K < 1 || K > kArapMaxK || K >= n)
    return fail(A3D_EINVAL, "a3d_knn_graph: need 1 <= K <= %d < n (got K=%d, n=%d) and non-null buffers", kArapMaxK, K, n);
  knn_kernel<<<(n + 127) / 128, 128, 0, st>>>(points, n, K, nbr, dist2);
  A3D_LAUNCH_CHECK();
  return A3D_OK;
}

extern "C" int a3d_arap(const float* nodes, int Nt, int Nv, const int32_t* nbr, int K, const float* weight, const int32_t* sample,
                        int Ns, float* err, float* grad, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!nodes || !nbr || !err || Nt < 1 || Nv < 1 || K < 1 || K > kArapMaxK)
    return fail(A3D_EINVAL, "a3d_arap: bad arguments (Nt=%d Nv=%d K=%d)", Nt, Nv, K);
  if (!sample) Ns = Nv;
  if (Ns < 1) return fail(A3D_EINVAL, "a3d_arap: empty node sample");
  A3D_CUDA_CHECK(cudaMemsetAsync(err, 0, sizeof(float), st));
  if (grad) A3D_CUDA_CHECK(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)Nt * Nv * 3, st));
  const int64_t total = (int64_t)(Nt - 1) * Ns;
  if (total > 0) {
    arap_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(nodes, Nt, Nv, nbr, K, weight, sample, Ns, err, grad);
    A3D_LAUNCH_CHECK();
  }
  return A3D_OK;
}
