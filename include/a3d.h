/* animate3d_b200 -- C ABI of the B200-native hot path (liba3d.so).
 *
 * Plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host.  All entry points
 * enqueue work on `stream` (a cudaStream_t passed as void*) and return 0 on success or a negative A3D_E* code; the
 * message of the last failure on the calling thread is available from a3d_last_error().  Nothing here allocates
 * device memory: the caller owns every buffer (workspace sizes are queried first), which is what makes the whole
 * UNet forward capturable in one CUDA graph.
 *
 * Which reference interface each entry point replaces (reference = yanqinJiang/Animate3D @ 033a1be):
 *   a3d_gemm            torch.nn.functional.linear / conv2d (cuBLAS / cuDNN) under diffusers ResnetBlock2D, Transformer2DModel,
 *                       TransformerTemporalModel, FeedForward/GEGLU -- called from animatediff/models/unet_motion_mv_model.py:768-859
 *   a3d_attention       xformers.ops.memory_efficient_attention at animatediff/models/attention_processor.py:103,233,268,405,416,656,691
 *   a3d_temporal_attn   Attention.get_attention_scores + torch.bmm at animatediff/models/attention_processor.py:634-635
 *   a3d_group_norm      torch GroupNorm(+SiLU) in ResnetBlock2D / Transformer2DModel / TransformerTemporalModel (5-D, over frames)
 *   a3d_layer_norm      torch LayerNorm in BasicTransformerBlock
 *   a3d_conv_in/out     unet_motion_mv_model.py:767-768 (permute + conv_in) and 859-862 (conv_out + permute back)
 *   a3d_ddim_cfg_step   animatediff/pipelines/pipeline.py:1023-1031 (CFG combine + DDIMScheduler.step + frame-0 re-injection)
 *   a3d_raster_*        diff_gaussian_rasterization._C.rasterize_gaussians / rasterize_gaussians_backward, called at
 *                       custom/threestudio-animate3d/renderer/diff_gaussian_rasterizer_advanced_4d.py:161-170
 */
#ifndef A3D_H_
#define A3D_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A3D_OK 0
#define A3D_EINVAL (-1)   /* bad argument / unsupported shape */
#define A3D_ECUDA (-2)    /* CUDA runtime or driver error     */
#define A3D_ENOTSUP (-3)  /* device is not sm_100             */

const char* a3d_last_error(void);
int a3d_version(void);
/* 0 when the current device is sm_100 and the driver entry points needed for TMA are available. */
int a3d_init(void);

/* ---------------------------------------------------------------- GEMM / implicit-GEMM convolution ----------- */
/* C[M,N] = epilogue( A[M,K] * B[N,K]^T ), fp16 operands, fp32 accumulation in tensor memory (tcgen05).
 * A operand modes: A3D_A_PLAIN  row-major [M, lda]
 *                  A3D_A_CONV3  NHWC image [n_img, H, W, C] gathered as a 3x3 (pad 1, stride 1|2) im2col, K = 9*C,
 *                               k index = (ky*3+kx)*C + c ; M = n_img*OH*OW
 * Epilogue:  v = acc + bias[n] + rowbias[((m / rb_div) % rb_mod) * rb_ld + n]
 *            v = acc_scale * v + r1_scale * R1[m, n] + R2[om, n]          (om = permuted output row, see below)
 *            GEGLU: columns come in interleaved blocks of 32 (u | g); out[.., j] = u_j * gelu_erf(g_j), N_out = N/2
 *            out[om, n] stored as fp16 (or fp32 when out_f32)
 * Row permutation (perm_a, perm_b > 0): om = (m / (a*b))*(a*b) + (m % b)*a + (m / b) % a   ("(x a b) -> (x b a)").
 */
enum { A3D_A_PLAIN = 0, A3D_A_CONV3 = 1 };
enum { A3D_GEMM_AUTO = 0, A3D_GEMM_TCGEN05 = 1, A3D_GEMM_SIMT = 2 };

typedef struct a3d_gemm_args {
  const void* A;      /* fp16 */
  const void* B;      /* fp16 weights [N, K], K contiguous */
  void* C;            /* fp16 (or fp32) output */
  int64_t M, N, K;
  int64_t lda, ldc;   /* elements; lda ignored for CONV3 */
  int a_mode;
  int conv_n, conv_h, conv_w, conv_c, conv_stride; /* CONV3 geometry (input) */
  int conv_nopad_lo;          /* CONV3: 0 = zero padding 1 on every side (UNet); 1 = no padding in front of row / column 0 and 1 behind
                                 the last ones = F.pad(x, (0,1,0,1)) + conv(padding=0), the SD-VAE Downsample2D */
  const float* bias;          /* [N] or NULL */
  const float* rowbias;       /* [rb_rows, rb_ld] or NULL */
  int64_t rb_ld, rb_div, rb_mod;
  float acc_scale;            /* multiplies the accumulator: MUST be set (1.0 if unused; 0.0 is honoured, not remapped) */
  const void* R1; int64_t ldr1; float r1_scale;   /* fp16 or NULL */
  const void* R2; int64_t ldr2;                   /* fp16 or NULL */
  int geglu;
  int out_f32;
  int64_t perm_a, perm_b;     /* 0,0 = identity */
  int impl;                   /* A3D_GEMM_* */
} a3d_gemm_args;

/* Replaces every Linear / Conv2d(3x3) the reference's forward issues through torch: ResnetBlock2D conv1/conv2/
 * time_emb_proj/conv_shortcut, Down/Upsample2D convs, Transformer2DModel proj_in/proj_out + BasicTransformerBlock FF
 * (GEGLU) (diffusers 0.28, reached from animatediff/models/unet_motion_mv_model.py:768-859), and the processors'
 * to_q/to_k/to_v/to_out + *_i2v / *_ip / *_sp projections (animatediff/models/attention_processor.py:211-231, 383-403,
 * 425-441, 600-655, 686-717), with their bias, residual, positional-table and layout epilogues fused. */
int a3d_gemm(const a3d_gemm_args* args, void* stream);

/* ---------------------------------------------------------------- fused attention (tcgen05) ------------------ */
/* O = softmax(scale * Q K^T) V per (batch, head); Q/K/V are read in place from projection outputs through rank-5
 * strided views so that the reference's "(b n f) l c -> (b f) (n l) c" regroupings never materialise.
 * A view addresses element (col, i1, i2, i3, i4) at base + col + i1*s1 + i2*s2 + i3*s3 + i4*s4 (element strides).
 * The sequence index l of a (batch) is l = i2*e1 + i1 (i1 < e1, i2 < e2); the batch index is qb = i4*e3 + i3.
 * Keys: batch kb = qb / kv_div; i3 is forced to 0 when kv_i3_zero (frame-0 keys of the I2V branch).
 * Head h reads Q/K columns [h*dqk, h*dqk+dqk) (dqk = roundup16(d), zero padded by the projection) and V columns
 * [h*dv, h*dv+dv) with dv = roundup16(d+1) whose column d holds 1.0 (the row sum then falls out of the PV product).
 * Output: fp16 [.., heads*d] written through the same view geometry as Q with row stride ldo:
 *   out = (accumulate ? out : 0) + out_scale * O.
 */
typedef struct a3d_view5 {
  const void* base;
  int64_t s1, s2, s3, s4;   /* element strides of dims 1..4 (dim 0 = columns, stride 1) */
  int64_t cols;             /* extent of dim 0 (total columns addressable from base) */
  int32_t e1, e2, e3, e4;   /* extents of dims 1..4 */
} a3d_view5;

typedef struct a3d_attn_args {
  a3d_view5 q, k, v;        /* k and v share extents */
  void* out;                /* fp16 */
  int64_t os1, os2, os3, os4; /* element strides of the output rows (same extents as q) */
  int heads, d;             /* true head dim (40/80/160) */
  float scale;
  int kv_div, kv_i3_zero;
  int accumulate; float out_scale;
  int impl;                 /* A3D_GEMM_AUTO / _TCGEN05 / _SIMT */
} a3d_attn_args;

/* Replaces the xformers.ops.memory_efficient_attention calls of the processors together with the einops regroups around
 * them: animatediff/models/attention_processor.py:233 (text keys) and 268 (IP-adapter image keys, accumulated with
 * `scale`), 405 (cross-view self-attention) and 416 (I2V branch, frame-0 keys: kv_i3_zero), 656 and 691 (spatio-temporal
 * processor's cross-view / image branches). */
int a3d_attention(const a3d_attn_args* args, void* stream);

/* Temporal attention over F frames for every (pixel, head): qkv [P, F, 3*C] fp16 (q | k | v), out [P, F, C].
 * Replaces the attention of the motion modules' temporal transformer blocks (diffusers TransformerTemporalModel reached
 * from unet_motion_mv_model.py:790-836) and the temporal branch of attention_processor.py:541-723 (line 103's call). */
int a3d_temporal_attn(const void* qkv, void* out, int64_t pixels, int frames, int heads, int d, float scale, int64_t ldo,
                      void* stream);   /* ldo: output row stride in halves (0 = C): lets the result land in a column block of a
                                          wider buffer, e.g. next to the cross-view branch for the merged output projection */

/* ---------------------------------------------------------------- normalisation / elementwise ---------------- */
/* Replaces nn.GroupNorm (+ SiLU) of ResnetBlock2D.norm1/norm2, Transformer2DModel.norm, TransformerTemporalModel.norm
 * (over frames) and conv_norm_out + conv_act (unet_motion_mv_model.py:262-266, 855-857).
 * GroupNorm over (rows_per_sample x C/groups) per sample, NHWC fp16.  x = concat(x1[C1], x2[C2]) along channels
 * (x2 may be NULL).  Optional SiLU.  Output rows may be permuted (perm_a, perm_b as in a3d_gemm).
 * Statistics are reduced without atomics in a fixed order ((n, mean, M2) partials merged with Chan's formula: per-thread
 * pivoted sums -> shared-memory tree -> warp-shuffle tree), so results are bit-reproducible and free of the
 * E[x^2]-E[x]^2 cancellation.  ws_stats: caller-owned scratch of a3d_group_norm_ws_bytes(samples, rows_per_sample, c1+c2,
 * groups) bytes (no initialisation needed). */
size_t a3d_group_norm_ws_bytes(int64_t samples, int64_t rows_per_sample, int c, int groups);
int a3d_group_norm(const void* x1, int c1, const void* x2, int c2, const float* gamma, const float* beta, void* y,
                   int64_t samples, int64_t rows_per_sample, int groups, float eps, int silu, int64_t perm_a,
                   int64_t perm_b, float* ws_stats, void* stream);
/* Input gradient of GroupNorm(+SiLU) for the VAE encoder on the SDS gradient path (animatemv_guidance.py:365-373; weights are
 * frozen, 301-302).  x / dy / dx: NHWC fp16 [samples * rows_per_sample, c]; fwd_ws_stats: the first 2*groups*samples floats of the
 * forward call's ws_stats ((mean, rstd) per (sample, group)); ws: scratch of a3d_group_norm_ws_bytes(...) bytes. */
int a3d_group_norm_backward(const void* x, int c, const float* gamma, const float* beta, const float* fwd_ws_stats, const void* dy,
                            void* dx, int64_t samples, int64_t rows_per_sample, int groups, int silu, float* ws, void* stream);
/* LayerNorm over C per row: BasicTransformerBlock.norm1/2/3 of the spatial and temporal transformers (diffusers). */
int a3d_layer_norm(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int c, float eps,
                   void* stream);
/* nearest x2 upsample NHWC: Upsample2D's F.interpolate in the up blocks (unet_motion_mv_model.py:838-850) */
int a3d_upsample2x(const void* x, void* y, int64_t n, int h, int w, int c, void* stream);
/* y[r, :] = silu(x[r / rep, :]) as fp16 (time-embedding broadcast for the time_emb_proj GEMM) */
int a3d_silu_rows(const float* x, void* y, int64_t rows, int c, int rep, void* stream);
/* conv_in: sample [BN, Cin, F, H, W] (fp32) -> NHWC fp16 [(BN F), H, W, Cout], 3x3 pad 1; w [Cout, Cin, 3, 3] fp32.
 * Replaces the permute/reshape + self.conv_in of unet_motion_mv_model.py:765-768. */
int a3d_conv_in(const float* sample, const float* w, const float* b, void* y, int bn, int cin, int f, int h, int wd,
                int cout, void* stream);
/* conv_out: NHWC fp16 [(BN F), H, W, Cin] -> [BN, Cout, F, H, W] fp32.  Replaces self.conv_out + the reshape/permute back
 * to [B, C, F, H, W] of unet_motion_mv_model.py:857-862. */
int a3d_conv_out(const void* x, const float* w, const float* b, float* y, int bn, int cin, int f, int h, int wd,
                 int cout, void* stream);
/* sinusoidal timestep projection: out[r, :] = [cos(t_r f_i), sin(t_r f_i)] fp32, dim = 2*half -- diffusers `Timesteps`
 * (flip_sin_to_cos=True, freq_shift=0) as called at unet_motion_mv_model.py:723, 734 (self.time_proj) */
int a3d_timestep_proj(const float* t, float* out, int rows, int half, void* stream);
/* small fp32 linear: y[M,N] = act(x[M,K]) W[N,K]^T + b  (act: 0 none, 1 silu);  add: y += previous y when accumulate.
 * The two TimestepEmbedding MLPs (time_embedding, camera_embedding; unet_motion_mv_model.py:730, 737, 742) and the IP-adapter
 * ImageProjection (encoder_hid_proj, 754-763). */
int a3d_linear_f32(const float* x, const float* w, const float* b, float* y, int m, int n, int k, int act_in,
                   int accumulate, void* stream);
/* fp32 -> fp16 cast with optional LayerNorm-free copy (utility) */
int a3d_cast_f32_f16(const float* x, void* y, int64_t n, void* stream);

/* ---------------------------------------------------------------- scheduler ---------------------------------- */
/* One DDIM (eta=0) update with classifier-free guidance and frame-0 re-injection (pipeline.py:1023-1031):
 *   eps = e_a + g * (e_b - e_a)   with (e_a, e_b) = (uncond, cond) halves of noise_pred when uncond_first, else the
 *   guidance ordering eps = e_text + g*(e_text - e_uncond);  x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t);
 *   x' = sqrt(a_prev) x0 + sqrt(1-a_prev) eps;  frame 0 of x' := first_frame.
 * latents/noise_pred: fp32 [BN(,x2), C, F, H, W]. */
int a3d_ddim_cfg_step(float* latents, const float* noise_pred, const float* first_frame, int bn, int c, int f, int hw,
                      float guidance, float alpha_t, float alpha_prev, int uncond_first, void* stream);

/* ---------------------------------------------------------------- 4D-Gaussian rasterizer --------------------- */
typedef struct a3d_raster_cam {
  float viewmatrix[16];   /* row-vector convention, as passed by threestudio/utils/ops.py:344-359 */
  float projmatrix[16];
  float campos[3];
  float tanfovx, tanfovy;
} a3d_raster_cam;

typedef struct a3d_raster_args {
  int P;                     /* gaussians */
  int H, W;
  int num_cams;              /* cameras rendered by this call (batched) */
  const a3d_raster_cam* cams;/* device array [num_cams] */
  const float* means3D;      /* [cam_stride_geom ? num_cams : 1][P,3] */
  const float* scales;       /* [..][P,3] */
  const float* rotations;    /* [..][P,4] */
  const float* opacities;    /* [P,1] */
  const float* shs;          /* [P,(deg+1)^2,3] or NULL */
  const float* colors_precomp; /* [P,3] or NULL */
  int sh_degree, sh_coeffs;
  int per_cam_geometry;      /* 1: means/scales/rotations have a leading camera dim */
  float scale_modifier;
  float bg[3];
} a3d_raster_args;

/* workspace bytes for a forward with at most max_rendered (tile,gaussian) pairs per camera */
size_t a3d_raster_workspace_bytes(int P, int H, int W, int num_cams, int64_t max_rendered);
/* forward: color [cams,3,H,W], depth [cams,1,H,W], alpha [cams,1,H,W], radii [cams,P] int32.
 * num_rendered_host (pinned, [cams]) receives the pair counts; returns A3D_EINVAL if a camera overflowed max_rendered. */
int a3d_raster_forward(const a3d_raster_args* args, float* color, float* depth, float* alpha, int32_t* radii,
                       void* workspace, size_t workspace_bytes, int64_t max_rendered, int64_t* num_rendered_host,
                       void* stream);
/* backward: grads w.r.t. means3D/scales/rotations ([cams or 1][P,*], accumulated over cameras when geometry is shared),
 * opacities [P,1], colors/shs, means2D [cams,P,3]. Needs the workspace of the matching forward untouched. */
int a3d_raster_backward(const a3d_raster_args* args, const float* dL_dcolor, const float* dL_ddepth,
                        const float* dL_dalpha, const float* alpha, const int32_t* radii, void* workspace,
                        size_t workspace_bytes, int64_t max_rendered, float* dL_dmeans3D, float* dL_dscales,
                        float* dL_drotations, float* dL_dopacity, float* dL_dcolors, float* dL_dshs,
                        float* dL_dmeans2D, void* stream);
/* debug/parity taps of the binning stage for one camera: sorted keys [n] uint64, point_list [n] uint32, ranges [tiles] uint2 */
int a3d_raster_binning_tap(const void* workspace, int P, int H, int W, int num_cams, int64_t max_rendered, int cam,
                           uint64_t* keys_out, uint32_t* point_list_out, uint32_t* ranges_out, void* stream);

/* ---------------------------------------------------------------- 4D deformation field (k-planes + MLPs) --------- */
/* Per (frame, gaussian): 32 k-planes features (num_scales x 16 channels, product over the 6 coordinate planes of
 * bilinear samples at (x,y,z,t)) -> three bias-free MLPs 32 -> 32 (ReLU) -> {3,4,3} -> means = xyz + d,
 * scales = exp(scaling + d) (d only when deform_scale), rotations = normalize(rotation + d).
 * Replaces Gaussian4DModel.interpolate_ms_features / get_xyz / get_scaling / get_rotation
 * (custom/threestudio-animate3d/geometry/gaussian_4d.py:450-548), without the optional global rot/trans branch.
 * planes[s*6+p] is a [channels, plane_h, plane_w] fp32 grid for coordinate pair p of (0,1),(0,2),(0,3),(1,2),(1,3),(2,3)
 * with the FIRST coordinate of the pair indexing W (grid_sample convention).  w1[m] [hidden, num_scales*channels],
 * w2[m] [out_m, hidden] for m = 0 (xyz, 3), 1 (rotation, 4), 2 (scaling, 3).  Outputs are [T, P, *].
 * backward accumulates (+=) into grad_planes / grad_w1 / grad_w2 (caller zeroes them). */
typedef struct a3d_deform_args {
  int P, T;
  const float* xyz;        /* [P,3] */
  const float* scaling;    /* [P,3] raw (log) scales */
  const float* rotation;   /* [P,4] raw quaternions */
  const float* times;      /* [T] timestamps in [-1,1] */
  int num_scales, channels, hidden;
  const float* planes[12];
  int plane_h[12], plane_w[12];
  const float* w1[3];
  const float* w2[3];
  int deform_scale;
  float* grad_planes[12];  /* backward only; entries may be NULL */
  float* grad_w1[3];
  float* grad_w2[3];
  /* `use_global_trans` support (gaussian_4d.py:499-511, 525-539); all optional (NULL = off) */
  const float* rot_base;      /* [T,P,4] per-frame base quaternions, replace `rotation` (forward and backward) */
  float* grad_rot_base;       /* backward: receives dL/d rot_base, [T,P,4] */
  const float* grad_featmean; /* backward: dL/d (per-frame mean feature), [T, num_scales*channels]; folded into grad_planes */
} a3d_deform_args;

int a3d_deform_forward(const a3d_deform_args* args, float* means, float* scales, float* rotations, void* stream);
/* per-frame mean over the P gaussians of the k-planes feature vector (`hidden_feats.mean(0)`, gaussian_4d.py:501, 527):
 * featmean [T, num_scales*channels] */
int a3d_deform_featmean(const a3d_deform_args* args, float* featmean, void* stream);
/* Accumulates into args->grad_w1 / grad_w2 ([out, in] like the weights) and args->grad_planes.  grad_planes[i] is a
 * CHANNEL-LAST scratch [plane_h[i], plane_w[i], channels] (zero it first; 16-byte aligned): the 16 channels of a texel are
 * written with vector reductions; transpose it into the [1, channels, H, W] parameter gradient afterwards. */
int a3d_deform_backward(const a3d_deform_args* args, const float* dL_dmeans, const float* dL_dscales, const float* dL_drotations,
                        void* stream);

/* ---------------------------------------------------------------- ARAP regulariser -------------- */
/* K nearest neighbours of every point among the same points, self excluded, squared distances ascending, ties by index:
 * pytorch3d.ops.knn_points(p, p, K=K+1)[..., 1:] as used by cal_connectivity_from_points
 * (custom/threestudio-animate3d/systems/util.py:79-82).  nbr [n,K] int32, dist2 [n,K]; K <= 12. */
int a3d_knn_graph(const float* points, int n, int K, int32_t* nbr, float* dist2, void* stream);
/* cal_arap_error (util.py:183-215) with estimate_rotation (138-173) fused, all frames in one launch:
 *   err = sum_{t>=1} sum_{i in sample} sum_n w[i,n] | e_in(t) - R_i(t) e_in(0) |^2,   e_in(t) = x_t[i] - x_t[nbr[i,n]],
 * R_i(t) = the weighted Procrustes rotation of the node's frame-0 edges onto its frame-t edges (no gradient through R).
 * nodes [Nt,Nv,3]; nbr [Nv,K] (-1 = no edge); weight [Nv,K] or NULL (1 on existing edges, the reference's call);
 * sample [Ns] node indices or NULL (all nodes).  err: device scalar; grad [Nt,Nv,3] = d err / d nodes or NULL.  Both are
 * zeroed here.  STATUS: arithmetic validated on the CPU (tests/test_arap_cpu.py); GPU run pending. */
int a3d_arap(const float* nodes, int Nt, int Nv, const int32_t* nbr, int K, const float* weight, const int32_t* sample, int Ns,
             float* err, float* grad, void* stream);

/* debug hook: a device uint64 counter the tcgen05 attention kernels bump once per (warp, key step) that takes the
   lazy-rescale branch of the single-pass softmax (NULL = off; tests use it to prove adversarial inputs reach that branch).
   The buffer must hold 8 + 4*32*8 + 32 uint64: words 8.. receive a clock64 timeline of four softmax warps of CTA (0,0,0) of the
   head-dim-40 kernel (tools/attn_timeline.py) */
int a3d_debug_set_attn_trace(void* device_counter_u64);
/* tuning hook of the head-dim-40 attention kernel (default 15 = all on): bit 0 one-step-ahead non-blocking barrier tests, bit 1
   probabilities through tensor memory (TS-mode P V), bit 2 one 128-query tile per CTA with two CTAs per SM, bit 3 (TS) no explicit
   "P buffer consumed" wait (implied by the arrival of the step's scores: tcgen05 operations retire in issue order) */
int a3d_debug_set_attn_poly(int flags);
/* measurement hooks of the rasterizer (bench.py's splat roofline): with timing enabled every forward / backward records CUDA
   events at its stage boundaries; a3d_debug_raster_stage_ms returns the milliseconds of the last forward's stages
   [0] preprocess [1] scan + counts + duplicate [2] radix sort [3] tile ranges [4] render and the last backward's [5] render
   backward [6] preprocess backward ([7] unused) */
int a3d_debug_raster_timing(int enable);
int a3d_debug_raster_stage_ms(float* out_host8);
/* same for the tcgen05 GEMM: per-tile timestamps of CTA 0 (epilogue warp 0 and the MMA-issuing thread) */
int a3d_debug_set_gemm_trace(void* device_buffer_1024_int64);

#ifdef __cplusplus
}
#endif
#endif /* A3D_H_ */
