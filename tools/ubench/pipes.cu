// Issue-rate microbenchmark for the SM sub-partition pipes the attention softmax leans on (MUFU.EX2, F2FP, FFMA, FFMA2,
// FMNMX3, integer ALU) and a few mixes.  One CTA per SM, W warps per CTA; every thread runs U independent dependency
// chains.  Prints warp-instructions per clock per SMSP (1.0 = the issue limit).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/ubench/pipes.cu && /tmp/pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { EX2H2, TANHH2, MUFU, F2FP, FFMA, FFMA2, FMNMX3, LEA, MUFU_FFMA, MUFU_F2FP, MUFU2_F2FP_FFMA2, POLY, MIX50, MIX25, CH_MUFU2_F2FP, CH_SOFTMAX, CH_SOFTMAX_NOMAX, NOPS };
static const char* kNames[] = {"ex2.approx.ftz.f16x2", "tanh.approx.f16x2", "MUFU.EX2", "F2FP.F16.F32.PACK", "FFMA", "FFMA2", "FMNMX3", "LEA/IADD", "MUFU+FFMA 1:1",
                               "MUFU+F2FP 2:1", "softmax pair: FFMA2+2MUFU+F2FP+FMNMX3", "poly pair (13 ops)",
                               "mix: 1 mufu pair + 1 poly pair", "mix: 3 mufu pairs + 1 poly pair",
                               "chained: 2 MUFU + 1 F2FP (same pipe -> 20 cycles, separate -> 16)", "chained softmax pair: FFMA2 + 2 MUFU + F2FP + FMNMX3",
                               "chained softmax pair without FMNMX3"};

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t tanhh2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmax2(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t lea23(uint32_t a, uint32_t b) { uint32_t d; asm volatile("{ .reg .u32 t; shl.b32 t, %1, 23; add.u32 %0, t, %2; }" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm volatile("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

// one "softmax pair" via MUFU: returns packed half2; s = scores, c = scale, m = -max
__device__ __forceinline__ uint32_t pair_mufu(uint64_t s, uint64_t c, uint64_t m, float& mx) {
  float a, b, sa, sb;
  upk(s, sa, sb);
  mx = fmax3(mx, sa, sb);
  upk(ffma2(s, c, m), a, b);
  return f2fp(ex2(a), ex2(b));
}
__device__ __forceinline__ uint32_t pair_poly(uint64_t s, uint64_t c, uint64_t m, float& mx) {
  float a, b, sa, sb;
  upk(s, sa, sb);
  mx = fmax3(mx, sa, sb);
  upk(ffma2(s, c, m), a, b);
  a = fmax2(a, -126.f); b = fmax2(b, -126.f);
  const uint64_t y = pk(a, b);
  const uint64_t magic = pk(12582912.f, 12582912.f), nmagic = pk(-12582912.f, -12582912.f);
  const uint64_t t = fadd2(y, magic);
  const uint64_t n = fadd2(t, nmagic);
  const uint64_t one = pk(1.f, 1.f), mone = pk(-1.f, -1.f);
  const uint64_t f = ffma2(n, mone, y);
  uint64_t r = ffma2(pk(0.0551716685f, 0.0551716685f), f, pk(0.2426111251f, 0.2426111251f));
  r = ffma2(r, f, pk(0.6932609677f, 0.6932609677f));
  r = ffma2(r, f, pk(0.9999280572f, 0.9999280572f));
  float r0, r1, t0, t1;
  upk(r, r0, r1); upk(t, t0, t1);
  (void)one;
  const float e0 = __uint_as_float(lea23(__float_as_uint(t0), __float_as_uint(r0)));
  const float e1 = __uint_as_float(lea23(__float_as_uint(t1), __float_as_uint(r1)));
  return f2fp(e0, e1);
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cycles, int iters, float seed) {
  constexpr int U = 8;
  float x[U];
  uint64_t xx[U];
  uint32_t acc = 0;
  float mx = 0.f;
#pragma unroll
  for (int u = 0; u < U; ++u) { x[u] = seed + u * 0.001f + threadIdx.x * 1e-6f; xx[u] = pk(x[u], x[u] + 0.5f); }
  const uint64_t c2 = pk(seed * 0.9f, seed * 0.9f), m2 = pk(-seed, -seed);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (OP == EX2H2) x[u] = __uint_as_float(ex2h2(__float_as_uint(x[u])));
      if (OP == TANHH2) x[u] = __uint_as_float(tanhh2(__float_as_uint(x[u])));
      if (OP == MUFU) x[u] = ex2(x[u]);
      if (OP == F2FP) { acc ^= f2fp(x[u], x[(u + 1) % U]); x[u] = __uint_as_float(acc); }
      if (OP == FFMA) x[u] = ffma(x[u], seed, 0.5f);
      if (OP == FFMA2) xx[u] = ffma2(xx[u], c2, m2);
      if (OP == FMNMX3) x[u] = fmax3(x[u], x[(u + 1) % U], seed);
      if (OP == LEA) x[u] = __uint_as_float(lea23(__float_as_uint(x[u]), acc));
      if (OP == MUFU_FFMA) x[u] = ex2(ffma(x[u], seed, 0.5f));
      if (OP == MUFU_F2FP) { if (u & 1) { acc ^= f2fp(ex2(x[u]), ex2(x[u - 1])); } }
      if (OP == MUFU2_F2FP_FFMA2) acc ^= pair_mufu(xx[u], c2, m2, mx);
      if (OP == POLY) acc ^= pair_poly(xx[u], c2, m2, mx);
      if (OP == MIX50) acc ^= (u & 1) ? pair_poly(xx[u], c2, m2, mx) : pair_mufu(xx[u], c2, m2, mx);
      if (OP == MIX25) acc ^= ((u & 3) == 3) ? pair_poly(xx[u], c2, m2, mx) : pair_mufu(xx[u], c2, m2, mx);
      // loop-carried versions (round 1's MUFU_F2FP / "softmax pair" lines had loop-invariant inputs and were folded by the compiler)
      if (OP == CH_MUFU2_F2FP) {
        float a, b;
        upk(xx[u], a, b);
        a = ex2(a); b = ex2(b);
        xx[u] = pk(a, b);                      // two MUFU chains per unit
        acc ^= f2fp(a, b);                     // one conversion of their results
      }
      if (OP == CH_SOFTMAX || OP == CH_SOFTMAX_NOMAX) {
        float a, b, sa, sb;
        upk(xx[u], sa, sb);
        if (OP == CH_SOFTMAX) mx = fmax3(mx, sa, sb);
        xx[u] = ffma2(xx[u], c2, m2);          // the scores keep changing: nothing is loop-invariant
        upk(xx[u], a, b);
        acc ^= f2fp(ex2(a), ex2(b));
      }
    }
    if (OP >= MUFU2_F2FP_FFMA2 && OP < CH_MUFU2_F2FP) {   // keep the inputs changing without adding work per pair
      xx[it & (U - 1)] = pk(mx, __uint_as_float(acc));
    }
  }
  const long long t1 = clock64();
  float s = mx + __uint_as_float(acc);
#pragma unroll
  for (int u = 0; u < U; ++u) { float a, b; upk(xx[u], a, b); s += x[u] + a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int warps, float* out, long long* cyc, int ops_per_unit) {
  const int iters = 2000;
  bench<OP><<<148, warps * 32>>>(out, cyc, iters, 0.37f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)h[i];
  avg /= 148;
  const double units = (double)iters * 8 * (warps / 4.0);   // per SMSP
  printf("%-46s warps/SMSP=%d  %8.2f cycles per unit per SMSP-warp-slot  (%.3f units/clk/SMSP, ~%d instr/unit)\n", kNames[OP],
         warps / 4, avg / units, units / avg, ops_per_unit);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  for (int w : {4, 8, 16}) {
    run<EX2H2>(w, out, cyc, 1);
    run<TANHH2>(w, out, cyc, 1);
    run<MUFU>(w, out, cyc, 1);
    run<F2FP>(w, out, cyc, 1);
    run<FFMA>(w, out, cyc, 1);
    run<FFMA2>(w, out, cyc, 1);
    run<FMNMX3>(w, out, cyc, 1);
    run<LEA>(w, out, cyc, 2);
    run<MUFU_FFMA>(w, out, cyc, 2);
    run<MUFU_F2FP>(w, out, cyc, 2);
    run<MUFU2_F2FP_FFMA2>(w, out, cyc, 5);
    run<POLY>(w, out, cyc, 13);
    run<MIX50>(w, out, cyc, 9);
    run<MIX25>(w, out, cyc, 7);
    run<CH_MUFU2_F2FP>(w, out, cyc, 3);
    run<CH_SOFTMAX>(w, out, cyc, 5);
    run<CH_SOFTMAX_NOMAX>(w, out, cyc, 4);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
