// a3d_gemm: C = epilogue(A * B^T) on 5th-gen tensor cores.
//   * persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..5 = epilogue
//   * operands staged by TMA into 128B-swizzled shared memory (K-major), accumulators double-buffered in TMEM so the
//     epilogue of tile i overlaps the main loop of tile i+1
//   * A3D_A_CONV3: the A operand is an implicit 3x3 im2col of an NHWC image -- each k-block is one (tap, 64-channel)
//     slice fetched with a rank-4 TMA box whose out-of-bounds rows/cols are zero-filled by the hardware (= padding 1);
//     stride-2 convolutions use the tensor map's traversal strides
// Replaces torch linear/conv2d (cuBLAS/cuDNN) under the diffusers blocks driven by
// animatediff/models/unet_motion_mv_model.py:768-859 of the reference.
#include <type_traits>

#include "a3d_common.cuh"
#include "a3d_host.cuh"

namespace a3d {

struct GemmDev {
  int64_t M, N, K;
  int num_k_blocks;
  int tiles_m, tiles_n;
  // conv A addressing
  int a_mode;
  int cpb;        // 64-channel blocks per tap (C/64)
  int conv_s;     // stride
  int tpi;        // output tiles per image (>=1) or 0 when several images share a tile
  int boh;        // output rows per tile
  int bimg;       // images per tile
  int tpr;        // tiles per output row (> 1 when an output row is wider than the 128-row tile)
  int conv_pad;   // zero padding in front of row / column 0 (1, or 0 for the asymmetric (0,1,0,1) padding of the VAE downsampler)
  // epilogue
  const float* bias;
  const float* rowbias;
  int64_t rb_ld, rb_div, rb_mod;
  int rb_stage;   // row-bias table rows a 128-row tile touches are staged in shared memory: 0 = no (global loads),
                  // 1 = slot = row_in_tile / rb_div (rb_div >= 8 divides 128, or rb_div % 128 == 0), 2 = slot = row_in_tile % rb_mod
  int rb_slots;   // distinct table rows per tile (<= 16)
  float acc_scale;
  const __half* R1; int64_t ldr1; float r1_scale;
  const __half* R2; int64_t ldr2;
  void* C; int64_t ldc;
  int geglu, out_f32;
  int64_t perm_a, perm_b;
  long long* trace;   // debug: per-tile clock64 timestamps of CTA 0 (null in production), see a3d_debug_set_gemm_trace
  int dbg_nostore;    // debug (A3D_GEMM_NOSTORE=1): compute the epilogue but skip the global stores -- timing experiments only
  int tma_store;      // epilogue writes through shared memory + TMA bulk stores (full 128-byte lines, asynchronous)
  int stages;         // depth of the TMA -> MMA operand ring
};

// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below fp16 output
// resolution): 1 rcp + 1 ex2 on the XU pipe + 11 FMA-pipe instructions instead of erff's ~30 (the GEGLU epilogue is ALU-bound)
__device__ __forceinline__ float gelu_erf(float x) {
  // gelu(x) = x/2 + |x|/2 erf(|x| / sqrt 2);  erf(z) = 1 - (a1 t + .. + a5 t^5) exp(-z^2), t = 1 / (1 + 0.3275911 z).  Constants folded so
  // that one evaluation is 7 FFMA + 4 FMUL + MUFU.RCP + MUFU.EX2 (the K = 320 GEGLU tiles are bound by epilogue instruction issue)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float xs = x * 0.84932180028801904f;              // sqrt(log2(e) / 2): exp(-x^2 / 2) = 2^(-xs^2)
  const float e = ex2_approx(-xs * xs);
  const float erf_abs = fmaf(-(poly * t), e, 1.0f);
  const float h = 0.5f * x;
  return fmaf(fabsf(h), erf_abs, h);
}

__device__ __forceinline__ int64_t perm_row(int64_t m, int64_t a, int64_t b) {
  if (a == 0) return m;
  const int64_t ab = a * b;
  return (m / ab) * ab + (m % b) * a + (m / b) % a;
}

constexpr int kBM = 128;
constexpr int kBK = 64;

template <int BN>
struct GemmCfg {
  // Ring depth is a launch parameter (GemmDev::stages): short-K GEMMs do not react to it at all (measured 2/3/4 stages at
  // BN = 256, profiles/r01_gemm_epilogue.txt) and give one stage to the TMA-store staging; the long-K implicit convolutions
  // (5-D TMA boxes, longer latency) keep the deep ring and store directly.
  static constexpr int kMaxStages = (BN == 256) ? 4 : (BN == 160 ? 5 : 6);
  static constexpr int kStagesStaged = (BN == 256) ? 3 : (BN == 160 ? 5 : 5);
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStageOut = 2048;   // 2 x 256 fp32 bias values (double-buffered with the accumulator)
  static constexpr int kRbLd = BN + 8;      // padded row of the staged row-bias slice (staggers the banks of the <=16 slots)
  static constexpr int kRbBytes = 16 * kRbLd * 4;
  // output staging for the TMA-store epilogue: one 32-row x 64-column (128 B per row, 128B-swizzled) box per epilogue warp
  static constexpr int kOutStage = (BN == 160) ? 0 : 8 * 4096;
  static constexpr int kTail = kStageOut + kRbBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int smem_bytes(bool staged) { return (staged ? kStagesStaged * kStageBytes + kOutStage : kMaxStages * kStageBytes) + kTail; }
  static constexpr int kSmemBytes = smem_bytes(true) > smem_bytes(false) ? smem_bytes(true) : smem_bytes(false);
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// EPI: 0 = bias/row-bias only, 1 = + residuals / row permutation, 2 = GEGLU, 3 = fp32 output (register pressure and dead
// code differ enough that one runtime-branched epilogue spilled)
enum { kEpiPlain = 0, kEpiRes = 1, kEpiGeglu = 2, kEpiF32 = 3 };

template <int BN, int EPI>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const GemmDev p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  const int kStages = p.stages;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_out = smem + kStages * Cfg::kStageBytes;           // [8 warps][32 rows][128 B], 1024-byte aligned
  uint8_t* smem_stage = smem_out + (p.tma_store ? Cfg::kOutStage : 0);
  float* smem_rb = reinterpret_cast<float*>(smem_stage + Cfg::kStageOut);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStageOut + Cfg::kRbBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;        // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer (highest warp ids: the scheduler
    // prefers them, so the single-thread control warps are not starved by the 8 busy epilogue warps)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.tiles_n, nt = tile % p.tiles_n;
        int img0 = 0, oh0 = 0, ow0 = 0;
        if (p.a_mode == A3D_A_CONV3) {
          if (p.tpr > 1) { img0 = mt / p.tpi; oh0 = (mt % p.tpi) / p.tpr; ow0 = ((mt % p.tpi) % p.tpr) * kBM; }
          else if (p.tpi > 0) { img0 = mt / p.tpi; oh0 = (mt % p.tpi) * p.boh; }
          else { img0 = mt * p.bimg; oh0 = 0; }
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (p.a_mode == A3D_A_CONV3) {
            const int tap = kb / p.cpb, c0 = (kb % p.cpb) * kBK;
            const int ky = tap / 3, kx = tap % 3;
            tma_load_5d(smem_a + stage * Cfg::kABytes, &mapA, &full_bar[stage], c0, ow0 * p.conv_s + kx - p.conv_pad,
                        oh0 * p.conv_s + ky - p.conv_pad, img0, 0);
          } else {
            tma_load_5d(smem_a + stage * Cfg::kABytes, &mapA, &full_bar[stage], kb * kBK, mt * kBM, 0, 0, 0);
          }
          tma_load_5d(smem_b + stage * Cfg::kBBytes, &mapB, &full_bar[stage], kb * kBK, nt * BN, 0, 0, 0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const bool tr = p.trace && blockIdx.x == 0 && tcount < 60;
        if (tr) p.trace[tcount * 16 + 8] = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        if (tr) p.trace[tcount * 16 + 9] = clock64();
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes), 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 halves (32 B) along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == 0 && tr) p.trace[tcount * 16 + 10] = clock64();
          if (kb == p.num_k_blocks - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (tr) p.trace[tcount * 16 + 11] = clock64();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps: 2 per TMEM lane quadrant)
    // TMEM -> registers (one accumulator row per thread) -> bias / row-bias / scale / residuals in fp32 -> fp16 ->
    // 256-bit global stores: every store instruction writes whole 32-byte sectors (the first version's 16-byte slivers
    // of 32 different lines ran the output at < 0.7 TB/s).  Residual tiles are fetched with 256-bit loads issued before
    // the TMEM wait.  The two warps of a quadrant split the tile's columns in units of 32.
    const int quad = warp & 3;
    const int half = warp >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tcount = 0;
    // row-bias table rows of a tile are FETCHED one tile ahead into registers (their global-load latency then hides behind the
    // previous tile's epilogue instead of standing in front of this one) and only copied to shared memory at the tile's top
    constexpr int kRbVec = BN / 4;
    constexpr int kRbPf = (16 * kRbVec + 255) / 256;      // float4 per thread for <= 16 staged rows
    float4 rb_pf[kRbPf];
    auto rb_fetch = [&](int t) {
      const int64_t tile_row0 = (int64_t)(t / p.tiles_n) * kBM;
      const int ntf = t % p.tiles_n;
#pragma unroll
      for (int k = 0; k < kRbPf; ++k) {
        const int i = threadIdx.x + 256 * k;
        rb_pf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < p.rb_slots * kRbVec) {
          const int slot = i / kRbVec, c4 = i % kRbVec;
          const int64_t rep = tile_row0 + (p.rb_stage == 1 ? (int64_t)slot * p.rb_div : (int64_t)slot);
          const int64_t trow = (rep / p.rb_div) % p.rb_mod;
          const int64_t col = (int64_t)ntf * BN + c4 * 4;
          if (col + 4 <= p.N && rep < p.M) rb_pf[k] = __ldg(reinterpret_cast<const float4*>(p.rowbias + trow * p.rb_ld + col));
        }
      }
    };
    constexpr bool kRbAhead = EPI == kEpiPlain;      // (the residual epilogue has no registers to spare; it fetches in place)
    constexpr bool rb_ahead = kRbAhead;
    if (rb_ahead && p.rb_stage && (int)blockIdx.x < num_tiles) rb_fetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile / p.tiles_n, nt = tile % p.tiles_n;
      const bool tr = p.trace && blockIdx.x == 0 && threadIdx.x == 0 && tcount < 60;
      if (tr) p.trace[tcount * 16 + 0] = clock64();
      // everything that does not read the accumulator (bias / row-bias staging, the residual tile) is started BEFORE the wait
      // for the tile's MMAs, so its latency hides behind them
      auto wait_acc = [&]() {
        mbar_wait(&tfull_bar[acc], acc_phase);
        if (tr) p.trace[tcount * 16 + 1] = clock64();
        tc_fence_after();
      };
      const int64_t row0 = (int64_t)mt * kBM + quad * 32;
      const int64_t row = row0 + lane;
      const bool row_ok = row < p.M;
      const int64_t rbrow = (p.rowbias && row_ok) ? ((row / p.rb_div) % p.rb_mod) : 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
      const int64_t n_out = (EPI == kEpiGeglu) ? p.N / 2 : p.N;
      const bool tile_full = (int64_t)(nt + 1) * BN <= p.N;       // no ragged columns in this tile: the per-unit bounds tests fold away
      // the tile's BN bias values: one coalesced load by the 256 epilogue threads, then broadcast reads from shared memory
      float* sb = reinterpret_cast<float*>(smem_stage) + acc * 256;
      // Row-bias (positional-encoding / time-embedding tables, fp32 [table rows, N]): a 128-row tile touches at most 16
      // distinct table rows.  Staged once per tile with coalesced loads instead of 128 x BN scattered 4-byte L2 reads (the
      // K=320 QKV projections with a table were 2x slower than the plain ones).  Single buffer: the barrier in front keeps
      // a fast warp from overwriting slots a slow warp still reads for the previous tile.
      const float* rbs = nullptr;
      if (p.rb_stage) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!rb_ahead) rb_fetch(tile);
#pragma unroll
        for (int k = 0; k < kRbPf; ++k) {
          const int i = threadIdx.x + 256 * k;
          if (i < p.rb_slots * kRbVec) *reinterpret_cast<float4*>(smem_rb + (i / kRbVec) * Cfg::kRbLd + (i % kRbVec) * 4) = rb_pf[k];
        }
        if (rb_ahead && tile + (int)gridDim.x < num_tiles) rb_fetch(tile + gridDim.x);
        const int rit = quad * 32 + lane;
        const int slot = p.rb_stage == 1 ? (int)(rit / p.rb_div) : (int)(rit % p.rb_mod);
        rbs = smem_rb + slot * Cfg::kRbLd;
      }
      if (p.bias) {
        const int64_t bc = (int64_t)nt * BN + threadIdx.x;
        if (threadIdx.x < BN) sb[threadIdx.x] = bc < p.N ? __ldg(p.bias + bc) : 0.f;
      }
      if (p.bias || p.rb_stage) asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tr) p.trace[tcount * 16 + 2] = clock64();

      // fp32 values of NC (16 or 32) consecutive accumulator columns -> + bias + rowbias, * scale
      auto finish = [&](float* v, int64_t col0, auto nc_tag) {
        constexpr int NC = decltype(nc_tag)::value;
        if (tile_full || col0 + NC <= p.N) {
          if (p.bias) {
            const float4* bs = reinterpret_cast<const float4*>(sb + (col0 - (int64_t)nt * BN));
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) {
              const float4 bv = bs[i];
              v[4 * i] += bv.x; v[4 * i + 1] += bv.y; v[4 * i + 2] += bv.z; v[4 * i + 3] += bv.w;
            }
          }
          if (rbs) {
            const float4* rb = reinterpret_cast<const float4*>(rbs + (col0 - (int64_t)nt * BN));
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) {
              const float4 bv = rb[i];
              v[4 * i] += bv.x; v[4 * i + 1] += bv.y; v[4 * i + 2] += bv.z; v[4 * i + 3] += bv.w;
            }
          } else if (p.rowbias) {
            const float4* rb = reinterpret_cast<const float4*>(p.rowbias + rbrow * p.rb_ld + col0);
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) {
              const float4 bv = __ldg(rb + i);
              v[4 * i] += bv.x; v[4 * i + 1] += bv.y; v[4 * i + 2] += bv.z; v[4 * i + 3] += bv.w;
            }
          }
        } else {
          for (int i = 0; i < NC; ++i) {
            if (col0 + i < p.N) {
              if (p.bias) v[i] += __ldg(p.bias + col0 + i);
              if (p.rowbias) v[i] += __ldg(p.rowbias + rbrow * p.rb_ld + col0 + i);
            }
          }
        }
        if (p.acc_scale != 1.0f) {
#pragma unroll
          for (int i = 0; i < NC; ++i) v[i] *= p.acc_scale;
        }
      };
      using N16 = std::integral_constant<int, 16>;
      using N32 = std::integral_constant<int, 32>;
      const int64_t orow = (EPI == kEpiRes && row_ok) ? perm_row(row, p.perm_a, p.perm_b) : row;
      // NC finished fp32 values of this thread's row -> (+ residuals) -> fp16 -> 256-bit stores (whole 32-byte sectors)
      auto store = [&](float* v, const uint32_t* r1, const uint32_t* r2, int64_t ocol, auto nc_tag) {
        constexpr int NC = decltype(nc_tag)::value;
        if (EPI == kEpiRes && p.R1) {
#pragma unroll
          for (int i = 0; i < NC / 2; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r1[i]));
            v[2 * i] += p.r1_scale * f.x; v[2 * i + 1] += p.r1_scale * f.y;
          }
        }
        if (EPI == kEpiRes && p.R2) {
#pragma unroll
          for (int i = 0; i < NC / 2; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r2[i]));
            v[2 * i] += f.x; v[2 * i + 1] += f.y;
          }
        }
        uint32_t h[NC / 2];
#pragma unroll
        for (int i = 0; i < NC / 2; ++i) h[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
        if (p.dbg_nostore && h[0] != 0x12345678u) return;
        if (Cfg::kOutStage && p.tma_store) {
          // this row's NC halves into the warp's staging box: 16-byte chunks (ocol % 64) / 8 .. of row `lane`
          uint8_t* box = smem_out + (warp & 7) * 4096;
          const int c16 = (int)(ocol & 63) >> 3;      // tiles start on multiples of 64 columns (BN = 128 / 256)
#pragma unroll
          for (int i = 0; i < NC / 8; ++i)
            *reinterpret_cast<uint4*>(box + sw128_offset(lane, c16 + i)) = *reinterpret_cast<const uint4*>(h + 4 * i);
          return;
        }
        __half* o = reinterpret_cast<__half*>(p.C) + orow * p.ldc + ocol;
        if (tile_full || ocol + NC <= n_out) {
#pragma unroll
          for (int i = 0; i < NC / 16; ++i) st_global_256(o + 16 * i, h + 8 * i);
        } else {
          for (int i = 0; i < NC; ++i) if (ocol + i < n_out) o[i] = reinterpret_cast<const __half*>(h)[i];
        }
      };
      // residual tiles: 16 columns (one 256-bit load each) per unit
      auto load_res = [&](uint32_t* r1, uint32_t* r2, int64_t ocol) {
        if (!tile_full && ocol + 16 > n_out) {   // ragged tail: scalar fill
          for (int i = 0; i < 16; ++i) {
            const bool ok = ocol + i < n_out;
            if (p.R1) reinterpret_cast<__half*>(r1)[i] = ok ? p.R1[row * p.ldr1 + ocol + i] : __float2half(0.f);
            if (p.R2) reinterpret_cast<__half*>(r2)[i] = ok ? p.R2[orow * p.ldr2 + ocol + i] : __float2half(0.f);
          }
          return;
        }
        if (p.R1) ld_global_256(p.R1 + row * p.ldr1 + ocol, r1);
        if (p.R2) ld_global_256(p.R2 + orow * p.ldr2 + ocol, r2);
      };
      auto load_one = [&](const __half* src, int64_t ld, int64_t r, uint32_t* dst, int64_t ocol) {   // 16 columns of one residual
        if (!tile_full && ocol + 16 > n_out) {
          for (int i = 0; i < 16; ++i) reinterpret_cast<__half*>(dst)[i] = ocol + i < n_out ? src[r * ld + ocol + i] : __float2half(0.f);
        } else {
          ld_global_256(src + r * ld + ocol, dst);
        }
      };

      if constexpr (EPI == kEpiF32) {
        // fp32 output (time-embedding table only)
        constexpr int U = BN / 32;
        const int u0 = half ? (U + 1) / 2 : 0, u1 = half ? U : (U + 1) / 2;
        wait_acc();
#pragma unroll 1
        for (int ch = u0; ch < u1; ++ch) {
          uint32_t r[32];
          tmem_ld32(taddr + ch * 32, r);
          tmem_wait_ld();
          const int64_t col0 = (int64_t)nt * BN + ch * 32;
          if (row_ok && col0 < p.N) {
            finish(reinterpret_cast<float*>(r), col0, N32{});
            float* o = reinterpret_cast<float*>(p.C) + row * p.ldc + col0;
            for (int i = 0; i < 32; ++i) if (col0 + i < p.N) o[i] = __uint_as_float(r[i]);
          }
        }
      } else {
        // Units of accumulator columns, split between the two warps of a quadrant.  Plain: 32 output columns per unit;
        // residual epilogue: 16 (keeps accumulators + two residual tiles double-buffered inside the register budget);
        // GEGLU: accumulator columns come as (u[32] | g[32]) blocks and a unit is half a block (16 u + 16 g columns -> 16
        // outputs).  The TMEM load (and the residual loads) of unit u+1 are issued BEFORE unit u is finished and stored:
        // with two epilogue warps per sub-partition nothing else hides the TMEM / L2 latency, and the K = 320 GEMMs of
        // level 0 were bound by exactly this chain.
        constexpr bool kG = EPI == kEpiGeglu;
        constexpr bool kR = EPI == kEpiRes;
        constexpr int W = kR ? 16 : 32;          // accumulator columns per unit (GEGLU: 16 + 16)
        constexpr int U = BN / W;
        const int u0 = half ? (U + 1) / 2 : 0, u1 = half ? U : (U + 1) / 2;
        auto acc_col = [&](int u) -> int64_t {   // first accumulator column of the unit's (first) column group
          return (int64_t)nt * BN + (kG ? (u >> 1) * 64 + 16 * (u & 1) : u * W);
        };
        auto issue = [&](uint32_t* r, uint32_t* r1, uint32_t* r2, int u) {
          if (kG) {
            tmem_ld16p(taddr + (u >> 1) * 64 + 16 * (u & 1), r);
            tmem_ld16p(taddr + (u >> 1) * 64 + 32 + 16 * (u & 1), r + 16);
          } else if (kR) {
            tmem_ld16p(taddr + u * 16, r);
            if (row_ok && (tile_full || acc_col(u) < p.N)) load_res(r1, r2, acc_col(u));
          } else {
            tmem_ld32p(taddr + u * 32, r);
          }
        };
        auto process = [&](uint32_t* r, uint32_t* r1, uint32_t* r2, int u) {
          const int64_t col0 = acc_col(u);
          if (!row_ok || (!tile_full && col0 >= p.N)) return;
          float* v = reinterpret_cast<float*>(r);
          if (kG) {
            finish(v, col0, N16{});
            finish(v + 16, col0 + 32, N16{});
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= gelu_erf(v[16 + i]);
            store(v, r1, r2, ((int64_t)nt * BN + (u >> 1) * 64) / 2 + 16 * (u & 1), N16{});
          } else if (kR) {
            finish(v, col0, N16{});
            store(v, r1, r2, col0, N16{});
          } else {
            finish(v, col0, N32{});
            store(v, r1, r2, col0, N32{});
          }
        };
        // TMA-store epilogue: a 64-column group of the warp's 32 rows (one 128-byte line per row) is complete after the
        // unit whose columns end on a multiple of 64; lane 0 then hands the box to the TMA engine and the warp moves on.
        // Stores leave the SM as whole lines, asynchronously (the direct 256-bit stores -- 32 different lines per
        // instruction -- cost the K = 320 GEMMs a third of their time, profiles/r01_gemm_epilogue.txt).
        const bool staged = Cfg::kOutStage && p.tma_store;
        auto group_done = [&](int u) {
          if (!staged) return;
          const int64_t cend = acc_col(u) + W;                       // one past the unit's last accumulator column
          if ((cend - (int64_t)nt * BN) % 64 != 0) return;
          const int64_t gcol = cend - 64;
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && (tile_full || gcol < n_out)) {
            tma_store_5d(&mapC, smem_out + (warp & 7) * 4096, (int)gcol, (int)row0, 0, 0, 0);
            tma_store_commit();
          }
        };
        // before the first write of a group: the TMA engine must have finished READING the warp's box (previous group)
        auto group_begin = [&](int u) {
          if (!staged || (acc_col(u) - (int64_t)nt * BN) % 64 != 0) return;
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        };
        if constexpr (kR) {
          // Residual epilogue: ALL residual (R2) columns this thread will add are requested before the accumulator wait -- they
          // do not depend on the MMAs, and with one 256-bit load in flight per 16-column unit the K = 320 / 640 projections
          // (+ residual) spent their epilogue on a chain of exposed L2 / HBM latencies (one per unit).
          constexpr int UH = (U + 1) / 2;
          uint32_t r2all[UH][8];
          const bool pre = p.R2 != nullptr && row_ok;
          if (pre) {
#pragma unroll
            for (int k = 0; k < UH; ++k) {
              const int u = u0 + k;
              if (u < u1 && (tile_full || acc_col(u) < p.N)) load_one(p.R2, p.ldr2, orow, r2all[k], acc_col(u));
              // (R1 -- unused by the model since the output GEMMs were merged -- stays on the per-unit path below)
            }
          }
          wait_acc();
          uint32_t ra[16], rb[16];
          uint32_t r1a[8], r1b[8], r2d[8];
          auto issue_r = [&](uint32_t* r, uint32_t* r1, int u) {
            tmem_ld16p(taddr + u * 16, r);
            if (p.R1 && row_ok && (tile_full || acc_col(u) < p.N)) load_one(p.R1, p.ldr1, row, r1, acc_col(u));
          };
          if (u0 < u1) issue_r(ra, r1a, u0);
#pragma unroll
          for (int k = 0; k < UH; ++k) {
            const int u = u0 + k;
            if (u < u1) {
              uint32_t* cur = (k & 1) ? rb : ra;
              uint32_t* nxt = (k & 1) ? ra : rb;
              uint32_t* cur1 = (k & 1) ? r1b : r1a;
              uint32_t* nxt1 = (k & 1) ? r1a : r1b;
              tmem_wait_ld();
              if (u + 1 < u1) issue_r(nxt, nxt1, u + 1);
              group_begin(u);
              process(cur, cur1, pre ? r2all[k] : r2d, u);
              group_done(u);
            }
          }
        } else {
          wait_acc();
          uint32_t ra[32], rb[32];
          uint32_t r1a[1], r2a[1], r1b[1], r2b[1];
          if (u0 < u1) issue(ra, r1a, r2a, u0);
#pragma unroll 1
          for (int u = u0; u < u1; u += 2) {
            tmem_wait_ld();
            if (u + 1 < u1) issue(rb, r1b, r2b, u + 1);
            group_begin(u);
            process(ra, r1a, r2a, u);
            group_done(u);
            if (u + 1 < u1) {
              tmem_wait_ld();
              if (u + 2 < u1) issue(ra, r1a, r2a, u + 2);
              group_begin(u + 1);
              process(rb, r1b, r2b, u + 1);
              group_done(u + 1);
            }
          }
        }
      }
      if (tr) p.trace[tcount * 16 + 3] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  if (Cfg::kOutStage && p.tma_store && warp < 8 && lane == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// SIMT bring-up / reference kernel: same semantics, one thread per output element (also serves odd shapes)
// ---------------------------------------------------------------------------------------------------------------
struct SimtConv { int n, h, w, c, s, oh, ow, pad; };

__global__ void gemm_simt_kernel(const GemmDev p, const __half* __restrict__ A, int64_t lda, const __half* __restrict__ B,
                                 SimtConv cv) {
  const int64_t n_out = p.geglu ? p.N / 2 : p.N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.M * n_out) return;
  const int64_t m = idx / n_out;
  const int64_t j = idx % n_out;
  auto dot = [&](int64_t n) {
    float acc = 0.f;
    const __half* b = B + n * p.K;
    if (p.a_mode == A3D_A_PLAIN) {
      const __half* a = A + m * lda;
      for (int64_t k = 0; k < p.K; ++k) acc += __half2float(a[k]) * __half2float(b[k]);
    } else {
      const int img = (int)(m / (cv.oh * cv.ow));
      const int oy = (int)((m / cv.ow) % cv.oh), ox = (int)(m % cv.ow);
      for (int tap = 0; tap < 9; ++tap) {
        const int iy = oy * cv.s + tap / 3 - cv.pad, ix = ox * cv.s + tap % 3 - cv.pad;
        if (iy < 0 || iy >= cv.h || ix < 0 || ix >= cv.w) continue;
        const __half* a = A + (((int64_t)img * cv.h + iy) * cv.w + ix) * cv.c;
        const __half* bb = b + (int64_t)tap * cv.c;
        for (int c = 0; c < cv.c; ++c) acc += __half2float(a[c]) * __half2float(bb[c]);
      }
    }
    return acc;
  };
  const int64_t orow = perm_row(m, p.perm_a, p.perm_b);
  if (p.geglu) {
    const int64_t blk = j / 32, e = j % 32;
    const int64_t nu = blk * 64 + e, ng = nu + 32;
    float u = dot(nu), g = dot(ng);
    if (p.bias) { u += p.bias[nu]; g += p.bias[ng]; }
    reinterpret_cast<__half*>(p.C)[orow * p.ldc + j] = __float2half_rn(u * gelu_erf(g));
    return;
  }
  float v = dot(j);
  if (p.bias) v += p.bias[j];
  if (p.rowbias) v += p.rowbias[((m / p.rb_div) % p.rb_mod) * p.rb_ld + j];
  v *= p.acc_scale;
  if (p.R1) v += p.r1_scale * __half2float(p.R1[m * p.ldr1 + j]);
  if (p.R2) v += __half2float(p.R2[orow * p.ldr2 + j]);
  if (p.out_f32) reinterpret_cast<float*>(p.C)[orow * p.ldc + j] = v;
  else reinterpret_cast<__half*>(p.C)[orow * p.ldc + j] = __float2half_rn(v);
}

template <int BN, int EPI>
static int launch_tc_epi(const GemmDev& dev, const CUtensorMap* mapA, const CUtensorMap* mapB, const CUtensorMap* mapC,
                         cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    A3D_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = dev.tiles_m * dev.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  GemmDev d2 = dev;
  d2.stages = dev.tma_store ? Cfg::kStagesStaged : Cfg::kMaxStages;
  gemm_tc_kernel<BN, EPI><<<grid, 320, Cfg::smem_bytes(dev.tma_store != 0), st>>>(d2, *mapA, *mapB, *mapC);
  A3D_LAUNCH_CHECK();
  return A3D_OK;
}

template <int BN>
static int launch_tc(const GemmDev& dev, const CUtensorMap* mapA, const CUtensorMap* mapB, const CUtensorMap* mapC,
                     cudaStream_t st) {
  if (dev.out_f32) return launch_tc_epi<BN, kEpiF32>(dev, mapA, mapB, mapC, st);
  if (dev.geglu) return launch_tc_epi<BN, kEpiGeglu>(dev, mapA, mapB, mapC, st);
  if (dev.R1 || dev.R2 || dev.perm_a) return launch_tc_epi<BN, kEpiRes>(dev, mapA, mapB, mapC, st);
  return launch_tc_epi<BN, kEpiPlain>(dev, mapA, mapB, mapC, st);
}

static long long* g_gemm_trace = nullptr;

}  // namespace a3d

// debug hook (not part of the product path): per-tile clock64 timestamps of CTA 0 of the following tcgen05 GEMM launches
extern "C" int a3d_debug_set_gemm_trace(void* device_buffer_1024_int64) {
  a3d::g_gemm_trace = reinterpret_cast<long long*>(device_buffer_1024_int64);
  return A3D_OK;
}

extern "C" int a3d_gemm(const a3d_gemm_args* a, void* stream) {
  using namespace a3d;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!a || !a->A || !a->B || !a->C) return fail(A3D_EINVAL, "a3d_gemm: null operand");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return fail(A3D_EINVAL, "a3d_gemm: empty problem M=%lld N=%lld K=%lld",
                                                        (long long)a->M, (long long)a->N, (long long)a->K);
  GemmDev d;
  memset(&d, 0, sizeof(d));
  d.M = a->M; d.N = a->N; d.K = a->K;
  d.a_mode = a->a_mode;
  d.bias = a->bias; d.rowbias = a->rowbias; d.rb_ld = a->rb_ld;
  d.rb_div = a->rb_div > 0 ? a->rb_div : 1; d.rb_mod = a->rb_mod > 0 ? a->rb_mod : (int64_t)1 << 40;
  d.acc_scale = a->acc_scale;   // the caller passes 1.0 when unused; 0 is a legitimate blend weight, not a sentinel
  d.R1 = reinterpret_cast<const __half*>(a->R1); d.ldr1 = a->ldr1; d.r1_scale = a->r1_scale;
  d.R2 = reinterpret_cast<const __half*>(a->R2); d.ldr2 = a->ldr2;
  d.C = a->C; d.ldc = a->ldc; d.geglu = a->geglu; d.out_f32 = a->out_f32;
  d.perm_a = a->perm_a; d.perm_b = a->perm_b;
  d.trace = g_gemm_trace;
  {
    static int nostore = -1;
    if (nostore < 0) { const char* e = getenv("A3D_GEMM_NOSTORE"); nostore = e ? atoi(e) : 0; }
    d.dbg_nostore = nostore;
  }
  if (a->geglu && (a->out_f32 || a->R1 || a->R2 || a->perm_a || (a->N % 128)))
    return fail(A3D_EINVAL, "a3d_gemm: GEGLU epilogue takes bias / row-bias only, fp16 output, N %% 128 == 0");
  if (a->out_f32 && (a->R1 || a->R2 || a->perm_a))
    return fail(A3D_EINVAL, "a3d_gemm: fp32 output supports bias / row-bias only");

  SimtConv cv{0, 0, 0, 0, 1, 0, 0};
  if (a->a_mode == A3D_A_CONV3) {
    const int s = a->conv_stride;
    if (s != 1 && s != 2) return fail(A3D_EINVAL, "a3d_gemm: conv stride must be 1 or 2");
    if (a->conv_h % s || a->conv_w % s) return fail(A3D_EINVAL, "a3d_gemm: conv H/W must be multiples of the stride");
    cv = SimtConv{a->conv_n, a->conv_h, a->conv_w, a->conv_c, s, a->conv_h / s, a->conv_w / s, a->conv_nopad_lo ? 0 : 1};
    if (a->K != 9LL * a->conv_c || a->M != (int64_t)cv.n * cv.oh * cv.ow)
      return fail(A3D_EINVAL, "a3d_gemm: conv geometry does not match M/K");
    d.conv_s = s;
  } else if (a->a_mode != A3D_A_PLAIN) {
    return fail(A3D_EINVAL, "a3d_gemm: unknown a_mode %d", a->a_mode);
  }

  // ---- can the tensor-core path take it?
  bool tc_ok = (a->K % kBK == 0) && (a->N % (a->geglu ? 16 : 8) == 0) && ((reinterpret_cast<uintptr_t>(a->A) & 15) == 0) &&
               ((reinterpret_cast<uintptr_t>(a->B) & 15) == 0) && (a->ldc % 16 == 0) &&
               ((reinterpret_cast<uintptr_t>(a->C) & 31) == 0);
  if (a->a_mode == A3D_A_PLAIN) tc_ok = tc_ok && (a->lda % 8 == 0) && a->lda >= a->K;
  int boh = 0, bimg = 1, tpi = 0, tpr = 1;
  if (a->a_mode == A3D_A_CONV3) {
    tc_ok = tc_ok && (cv.c % kBK == 0);
    const int opix = cv.oh * cv.ow;
    if (cv.ow > kBM) {           // wide images (VAE at 256^2): a tile is a 128-pixel piece of one output row
      tc_ok = tc_ok && (cv.ow % kBM == 0);
      boh = 1; tpr = cv.ow / kBM; tpi = cv.oh * tpr; bimg = 1;
    } else if (opix >= kBM) {
      tc_ok = tc_ok && (opix % kBM == 0) && (kBM % cv.ow == 0);
      boh = kBM / cv.ow; tpi = opix / kBM; bimg = 1;
    } else {
      tc_ok = tc_ok && (kBM % opix == 0);
      boh = cv.oh; tpi = 0; bimg = kBM / (opix > 0 ? opix : 1);
    }
    tc_ok = tc_ok && (tpr > 1 ? kBM : cv.ow) * cv.s <= 256 && boh * cv.s <= 256;
  }
  if (a->R1) tc_ok = tc_ok && (a->ldr1 % 16 == 0) && ((reinterpret_cast<uintptr_t>(a->R1) & 31) == 0);
  if (a->R2) tc_ok = tc_ok && (a->ldr2 % 16 == 0) && ((reinterpret_cast<uintptr_t>(a->R2) & 31) == 0);
  int impl = a->impl;
  if (impl == A3D_GEMM_AUTO) impl = tc_ok ? A3D_GEMM_TCGEN05 : A3D_GEMM_SIMT;
  if (impl == A3D_GEMM_TCGEN05 && !tc_ok)
    return fail(A3D_EINVAL, "a3d_gemm: shape/alignment not supported by the tcgen05 path (M=%lld N=%lld K=%lld)",
                (long long)a->M, (long long)a->N, (long long)a->K);

  if (impl == A3D_GEMM_SIMT) {
    const int64_t n_out = a->geglu ? a->N / 2 : a->N;
    const int64_t total = a->M * n_out;
    const int threads = 128;
    const int64_t blocks = (total + threads - 1) / threads;
    gemm_simt_kernel<<<(unsigned)blocks, threads, 0, st>>>(d, reinterpret_cast<const __half*>(a->A), a->lda,
                                                           reinterpret_cast<const __half*>(a->B), cv);
    A3D_LAUNCH_CHECK();
    return A3D_OK;
  }

  // ---- tile shape
  int BN;
  if (a->geglu) BN = (a->N % 256 == 0) ? 256 : 128;
  else if (a->N % 256 == 0) BN = 256;
  else if (a->K <= 640 && a->N > 512 && (256 * ((a->N + 255) / 256) - a->N) * 8 <= a->N) {
    // short-K GEMMs are bound by the epilogue / output stores, where the 256-wide tile amortises the per-tile handshakes
    // best: a ragged last tile (<= 12.5% idle MMA columns) costs nothing there
    BN = 256;
  } else if (a->N % 160 == 0) BN = 160;
  else BN = 128;
  {
    static int force = -1;
    if (force < 0) { const char* e = getenv("A3D_GEMM_BN"); force = e ? atoi(e) : 0; }
    if (!a->geglu && (force == 128 || force == 160 || force == 256)) BN = force;   // tuning override
  }
  if (a->geglu && a->N % BN) return fail(A3D_EINVAL, "a3d_gemm: GEGLU needs N %% 128 == 0");
  d.num_k_blocks = (int)(a->K / kBK);
  d.tiles_m = (int)((a->M + kBM - 1) / kBM);
  d.tiles_n = (int)((a->N + BN - 1) / BN);
  d.cpb = a->a_mode == A3D_A_CONV3 ? cv.c / kBK : 0;
  d.tpi = tpi; d.boh = boh; d.bimg = bimg; d.tpr = tpr; d.conv_pad = cv.pad;
  d.rb_stage = 0; d.rb_slots = 0;
  if (a->rowbias && (a->N % 4 == 0) && (a->rb_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->rowbias) & 15) == 0)) {
    if (d.rb_div % kBM == 0) { d.rb_stage = 1; d.rb_slots = 1; }                                  // one table row per tile
    else if (d.rb_div >= 8 && kBM % d.rb_div == 0) { d.rb_stage = 1; d.rb_slots = (int)(kBM / d.rb_div); }
    else if (d.rb_div == 1 && d.rb_mod <= 16 && kBM % d.rb_mod == 0) { d.rb_stage = 2; d.rb_slots = (int)d.rb_mod; }
  }

  const CUtensorMap *mapA = nullptr, *mapB = nullptr;
  {
    MapKey kb;
    const uint64_t dims[5] = {(uint64_t)a->K, (uint64_t)a->N, 1, 1, 1};
    const uint64_t str[4] = {(uint64_t)a->K, (uint64_t)a->K * a->N, (uint64_t)a->K * a->N, (uint64_t)a->K * a->N};
    const uint32_t box[5] = {kBK, (uint32_t)BN, 1, 1, 1};
    kb = make_key(a->B, dims, str, box);
    if (int r = get_tensor_map(kb, &mapB)) return r;
  }
  if (a->a_mode == A3D_A_PLAIN) {
    const uint64_t dims[5] = {(uint64_t)a->K, (uint64_t)a->M, 1, 1, 1};
    const uint64_t str[4] = {(uint64_t)a->lda, (uint64_t)a->lda * a->M, (uint64_t)a->lda * a->M, (uint64_t)a->lda * a->M};
    const uint32_t box[5] = {kBK, kBM, 1, 1, 1};
    MapKey ka = make_key(a->A, dims, str, box);
    if (int r = get_tensor_map(ka, &mapA)) return r;
  } else {
    const uint64_t img = (uint64_t)cv.h * cv.w * cv.c;
    const uint64_t dims[5] = {(uint64_t)cv.c, (uint64_t)cv.w, (uint64_t)cv.h, (uint64_t)cv.n, 1};
    const uint64_t str[4] = {(uint64_t)cv.c, (uint64_t)cv.w * cv.c, img, img * cv.n};
    const uint32_t box[5] = {kBK, (uint32_t)((tpr > 1 ? kBM : cv.ow) * cv.s), (uint32_t)(boh * cv.s), (uint32_t)bimg, 1};
    MapKey ka = make_key(a->A, dims, str, box);
    ka.estr[1] = cv.s; ka.estr[2] = cv.s;
    if (int r = get_tensor_map(ka, &mapA)) return r;
  }
  // output through shared memory + TMA bulk stores where the epilogue allows it (plain / residual, no row permutation)
  const CUtensorMap* mapC = mapA;
  d.tma_store = 0;
  {
    static int want = -1;
    if (want < 0) { const char* e = getenv("A3D_GEMM_TMA_STORE"); want = e ? atoi(e) : 1; }
    // short-K GEMMs are bound by their output stores; the long-K ones (convolutions) keep the deeper operand ring instead
    if (want && BN != 160 && a->K <= 1280 && !a->geglu && !a->out_f32 && !a->perm_a && (a->ldc % 8 == 0) &&
        ((reinterpret_cast<uintptr_t>(a->C) & 15) == 0) && (a->N % 8 == 0)) {
      const uint64_t dims[5] = {(uint64_t)a->N, (uint64_t)a->M, 1, 1, 1};
      const uint64_t str[4] = {(uint64_t)a->ldc, (uint64_t)a->ldc * a->M, (uint64_t)a->ldc * a->M, (uint64_t)a->ldc * a->M};
      const uint32_t box[5] = {64, 32, 1, 1, 1};
      MapKey kc = make_key(a->C, dims, str, box);
      if (int r = get_tensor_map(kc, &mapC)) return r;
      d.tma_store = 1;
    }
  }
  switch (BN) {
    case 256: return launch_tc<256>(d, mapA, mapB, mapC, st);
    case 160: return launch_tc<160>(d, mapA, mapB, mapC, st);
    default: return launch_tc<128>(d, mapA, mapB, mapC, st);
  }
}
