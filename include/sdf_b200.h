/*
 * sdf_b200.h — C ABI of libsdf_b200.so: the sm_100a kernels behind
 * stable-dreamfusion's operator surface (raymarching.*, grid_encode, sh_encode,
 * freq_encode, sd_utils.StableDiffusion.train_step).
 *
 * Conventions:
 *   - every pointer is a raw DEVICE pointer unless its name starts with host_;
 *     all arrays are contiguous, row-major;
 *   - the caller allocates every output; nothing is allocated or retained
 *     natively except small per-device scratch (documented per function);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     no entry point synchronises the host unless stated;
 *   - return value: 0 = ok, <0 = error (SDF_ERR_*), message via sdf_last_error();
 *     launch errors are checked with cudaPeekAtLastError after each launch.
 * Each declaration cites the reference interface it replaces
 * (paths relative to the reference repository root).
 */
#ifndef SDF_B200_H
#define SDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDF_OK 0
#define SDF_ERR_ARG (-1)
#define SDF_ERR_CUDA (-2)
#define SDF_ERR_UNSUPPORTED (-3)

const char* sdf_last_error(void);
int sdf_abi_version(void);

/* ------------------------------------------------------------------ raymarching
 * replaces raymarching/src/raymarching.h:6-18 (pybind: raymarching/src/bindings.cpp:5-19) */

/* near_far_from_aabb (raymarching.h:6; kernel raymarching.cu:92) */
int sdf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                           float* nears, float* fars, void* stream);
/* sph_from_ray (raymarching.h:7; raymarching.cu:163) */
int sdf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, void* stream);
/* morton3D / morton3D_invert (raymarching.h:8-9; raymarching.cu:214,237) */
int sdf_morton3D(const int* coords, uint32_t N, int* indices, void* stream);
int sdf_morton3D_invert(const int* indices, uint32_t N, int* coords, void* stream);
/* packbits (raymarching.h:10; raymarching.cu:268).  N = number of output bytes; grid 16-byte aligned. */
int sdf_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield, void* stream);
/* flatten_rays (raymarching.h:11; raymarching.cu:303) */
int sdf_flatten_rays(const int* rays, uint32_t N, uint32_t M, int* res, void* stream);

/* march_rays_train (raymarching.h:13; raymarching.cu:338,477).  The reference's
 * two kernel passes around a host sync become:
 *   _count: rays[N,2] <- (exclusive offset in ray order, count); counter[0] <- M on the device;
 *           host_M (optional, pinned host int32) receives M via a device store.
 *   _write: emits xyzs[M,3] dirs[M,3] ts[M,2] at the offsets in rays; rows beyond `capacity` are never written. */
int sdf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, int contract,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                               const float* nears, const float* fars, const float* noises /* may be NULL */,
                               int* rays, int* counter, int* host_M, void* stream);
int sdf_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, int contract,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                               const float* nears, const float* fars, const float* noises,
                               float* xyzs, float* dirs, float* ts, const int* rays, uint32_t capacity, void* stream);

/* composite_rays_train_forward / _backward (raymarching.h:14-15; raymarching.cu:501,606).
 * forward writes weights for every sample of every valid ray (0 after early termination). */
int sdf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts, const int* rays,
                                     uint32_t M, uint32_t N, float T_thresh, int binarize,
                                     float* weights, float* weights_sum, float* depth, float* image, void* stream);
int sdf_composite_rays_train_backward(const float* grad_weights /* may be NULL */, const float* grad_weights_sum /* may be NULL */,
                                      const float* grad_depth /* may be NULL */, const float* grad_image,
                                      const float* sigmas, const float* rgbs, const float* ts, const int* rays,
                                      const float* weights_sum, const float* depth, const float* image,
                                      uint32_t M, uint32_t N, float T_thresh, int binarize,
                                      float* grad_sigmas, float* grad_rgbs, void* stream);

/* march_rays / composite_rays, inference (raymarching.h:17-18; raymarching.cu:714,843) */
int sdf_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t, const float* rays_o,
                   const float* rays_d, float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                   const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* ts,
                   const float* noises /* may be NULL */, void* stream);
int sdf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int binarize, int* rays_alive, float* rays_t,
                       const float* sigmas, const float* rgbs, const float* ts, float* weights_sum, float* depth, float* image,
                       void* stream);

/* The same loop (nerf/renderer.py:759-794) with its bookkeeping on the device: `state` is 32 bytes of device memory holding
 * (n_alive, n_step = clamp(N / n_alive, 1, 8), M = n_alive * n_step, step, ...); every call is launched with capacity N and reads the
 * counts itself; &state[2] (int32 M) is the m_dev of sdf_field_forward.  sdf_infer_compact replaces `rays_alive[rays_alive >= 0]`
 * (a host sync + three PyTorch kernels) and advances the state; host_alive (optional, pinned int32) mirrors n_alive for polling. */
int sdf_infer_begin(void* state, uint32_t N, uint32_t max_steps, int* rays_alive, float* rays_t, const float* nears,
                    float* weights_sum, float* depth, float* image, int* host_alive, void* stream);
int sdf_infer_march(const void* state, uint32_t N, const int* rays_alive, const float* rays_t, const float* rays_o, const float* rays_d,
                    float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                    const float* fars, float* xyzs, float* dirs, float* ts, const float* noises /* may be NULL */, void* stream);
int sdf_infer_composite(const void* state, uint32_t N, float T_thresh, int binarize, int* rays_alive, float* rays_t, const float* sigmas,
                        const float* rgbs, const float* ts, float* weights_sum, float* depth, float* image, void* stream);
int sdf_infer_compact(void* state, uint32_t N, const int* rays_in, int* rays_out, int* host_alive, void* stream);

/* ------------------------------------------------------------------ gridencoder
 * replaces gridencoder/src/gridencoder.h:12-16 (pybind: gridencoder/src/bindings.cpp:5-10).
 * dtype: 0 = fp32 table/outputs/grads, 1 = fp16.  inputs are fp32 in [0,1].
 * Layout change vs the reference: outputs and grad are [B, L*C] (point-major, what
 * grid.py:64 / :82 produce by permuting), dy_dx is [B, L, D, C] as in the reference. */
int sdf_grid_encode_forward(const float* inputs, const void* embeddings, const int* offsets, void* outputs,
                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                            void* dy_dx /* may be NULL */, uint32_t gridtype, int align_corners, uint32_t interp, int dtype, void* stream);
int sdf_grid_encode_backward(const void* grad, const float* inputs, const int* offsets, void* grad_embeddings /* accumulated into */,
                             uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                             const void* dy_dx /* may be NULL */, void* grad_inputs /* may be NULL */,
                             uint32_t gridtype, int align_corners, uint32_t interp, int dtype /* of grad, dy_dx, grad_inputs */,
                             int acc_dtype /* of grad_embeddings: (dtype,acc) in (0,0),(1,1),(1,0) */, void* stream);
/* grad_total_variation / grad_weight_decay (gridencoder.h:15-16; gridencoder.cu:526,671), fp32 */
int sdf_grid_grad_total_variation(const float* inputs, const float* embeddings, float* grad, const int* offsets, float weight,
                                  uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                  int align_corners, void* stream);
int sdf_grid_grad_weight_decay(const float* embeddings, float* grad, const int* offsets, float weight,
                               uint32_t n_entries, uint32_t C, uint32_t L, void* stream);
/* device-evaluated ceil(exp2f(level*S)*H) per level (gridencoder.cu:133), copied to host_out[L]; synchronises `stream`. */
int sdf_grid_level_resolutions(const int* offsets, uint32_t L, float S, uint32_t H, uint32_t* host_out, void* stream);

/* ------------------------------------------------------------------ freqencoder / shencoder
 * replace freqencoder/src/freqencoder.h:6-9 and shencoder/src/shencoder.h:9-10 */
int sdf_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs, void* stream);
int sdf_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                             float* grad_inputs, void* stream);
int sdf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree,
                          float* dy_dx /* may be NULL; [B, 3, degree^2] */, void* stream);
int sdf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx,
                           float* grad_inputs /* accumulated into */, void* stream);

/* ------------------------------------------------------------------ fused radiance field
 * replaces NeRFNetwork.forward / .density / .common_forward / .normal of the -O backbone
 * (nerf/network_grid.py:68-142): hashgrid(L16, C2, smoothstep) -> MLP 32-64-64-4 -> trunc_exp(+density_blob) / sigmoid
 * -> finite-difference normal (6 extra evaluations at x +- 1e-2) -> safe_normalize -> shading, in one kernel.
 * xyzs [M,3] fp32 in [-bound,bound]; table_fp16 [n_entries,2] half; w*,b*: fp32 nn.Linear parameters of sigma_net;
 * shading 0 albedo / 1 lambertian / 2 textureless / 3 normal; light_d [3] or [M,3] (light_per_sample);
 * m_dev: optional device int32 with the live sample count (<= M).  aux [M,10] fp32: stash for the backward. */
int sdf_field_forward(const float* xyzs, uint32_t M, const int* m_dev, const void* table_fp16, const int* offsets,
                      uint32_t n_levels, uint32_t n_levels_active, float per_level_scale_log2, uint32_t base_resolution,
                      int interp_smoothstep, const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* w3, const float* b3, float bound, float blob_density, float blob_radius,
                      int shading, const float* light_d, int light_per_sample, float ambient_ratio,
                      float* sigmas, float* colors /* may be NULL */, float* normals /* may be NULL */, float* aux /* may be NULL */,
                      void* feat /* may be NULL: feature stash, sdf_field_feat_bytes(M, shading) bytes */, void* stream);
/* size of the optional feature stash (interpolated features of every stencil point, so the backward need not gather again) */
long long sdf_field_feat_bytes(uint32_t M, int shading);
/* backward: gradients are ACCUMULATED into grad_table [n_entries,2] fp32 and gw1 [64,32] gb1 [64] gw2 [64,64] gb2 [64] gw3 [4,64] gb3 [4]
 * (replaces 7 x {MLP autograd, gridencoder.cu:253 kernel_grid_backward} of the reference). */
int sdf_field_backward(const float* xyzs, uint32_t M, const int* m_dev, const void* table_fp16, const int* offsets,
                       uint32_t n_levels, uint32_t n_levels_active, float per_level_scale_log2, uint32_t base_resolution,
                       int interp_smoothstep, const float* w1, const float* b1, const float* w2, const float* b2,
                       const float* w3, const float* b3, float bound, float blob_density, float blob_radius,
                       int shading, const float* light_d, int light_per_sample, float ambient_ratio, const float* aux,
                       const float* g_sigmas /* may be NULL */, const float* g_colors /* may be NULL */, const float* g_normals /* may be NULL */,
                       float* grad_table, float* gw1, float* gb1, float* gw2, float* gb2, float* gw3, float* gb3,
                       const void* feat /* may be NULL: re-gather */, void* stream);

/* ------------------------------------------------------------------ SD-1.5-shaped UNet / VAE encoder building blocks
 * replace the cuDNN / cuBLAS calls underneath guidance/sd_utils.py:95-108 (diffusers UNet2DConditionModel / AutoencoderKL;
 * module structure per the vendored CompVis code: ldm/modules/diffusionmodules/openaimodel.py:164-277,414-778,
 * ldm/modules/attention.py:152-275, ldm/modules/diffusionmodules/model.py:82-204,368-460).
 * Activations are NHWC fp16. */

/* tcgen05 implicit-GEMM plan:
 *   out[img,y,x,n] = act(alpha * sum_{tap,c} a[img, y+dy(tap), x+dx(tap), c] * wt[(img,y,) n, tap*Cin + c] + bias[n] + temb[img,n] + residual[img,y,x,n])
 * Strides are in fp16 ELEMENTS and multiples of 8.  a: element (img,y,x,c) at img*a_simg + y*a_sy + x*a_sx + c; channels >= a_c_valid
 * read as zero (TMA fill).  wt: row n, K index k at (img*w_simg + y*w_sy) + n*w_ld + k; k >= w_k_valid / rows >= n_rows_w read as zero;
 * w_sy = w_simg = 0 -> shared weights (conv / linear); non-zero -> batched product (attention: y = head, img = batch).
 * Cin = K iterated per tap (multiple of 64); taps = 1 or 9 (3x3, stride 1, zero pad 1; tap = ky*3+kx; packed [n][tap][Cin]).
 * linear: H = 1, Nimg = 1, W = rows.  block_n in {64,128,160}.  splitk > 1 needs workspace fp32 [Nimg*H*W, N].
 * cta_pair = 1: 2-CTA clusters issue tcgen05.mma.cta_group::2 (M = 256), each CTA stages half of the weight tile;
 * block_n in {128,160,256}; not available for batched products (w_sy / w_simg != 0).
 * Returns a plan handle >= 0, or a negative error code. */
int sdf_gemm_plan_create(const void* a, long long a_sx, long long a_sy, long long a_simg, int a_c_valid,
                         const void* wt, long long w_ld, long long w_sy, long long w_simg, int w_k_valid, int n_rows_w,
                         int Nimg, int H, int W, int Cin, int taps, int N,
                         void* out, long long o_sx, long long o_sy, long long o_simg,
                         const float* bias, const void* temb, int temb_ld,
                         const void* residual, long long r_sx, long long r_sy, long long r_simg,
                         int act, float alpha, int splitk, float* workspace, int block_n, int cta_pair);
/* the same for a STRIDED 3x3 convolution (taps = 9): H, W are the OUTPUT geometry, the input a is [Nimg, H*stride, W*stride, .] and
 *   out[img,y,x,n] = ... sum a[img, y*stride + ky - pad_lo, x*stride + kx - pad_lo, c] * wt[n, (ky*3+kx)*Cin + c] ...
 * read straight out of the input with an element-strided TMA box (no im2col buffer): the UNet's Downsample (stride 2, pad 1,
 * ldm/modules/diffusionmodules/openaimodel.py:130-138) is (stride 2, pad_lo 1), the VAE encoder's (zero pad (0,1,0,1) then stride 2,
 * model.py:67-79) is (stride 2, pad_lo 0).  stride in {1, 2}; stride = 1, pad_lo = 1 is sdf_gemm_plan_create. */
int sdf_gemm_plan_create_strided(const void* a, long long a_sx, long long a_sy, long long a_simg, int a_c_valid,
                                 const void* wt, long long w_ld, long long w_sy, long long w_simg, int w_k_valid, int n_rows_w,
                                 int Nimg, int H, int W, int Cin, int taps, int N,
                                 void* out, long long o_sx, long long o_sy, long long o_simg,
                                 const float* bias, const void* temb, int temb_ld,
                                 const void* residual, long long r_sx, long long r_sy, long long r_simg,
                                 int act, float alpha, int splitk, float* workspace, int block_n, int cta_pair, int stride, int pad_lo);
/* let the plan's epilogue accumulate the GroupNorm(32, C) statistics of a consumer of its output (two consumers per plan): stats fp32
 * [Nimg, 32, 2] zeroed by the caller; channel_offset = consumer channel of this product's column 0 (concatenated inputs).  Returns
 * SDF_ERR_UNSUPPORTED for split-K / ragged-N / GEGLU plans: the caller then keeps sdf_groupnorm_forward's own statistics pass. */
int sdf_gemm_plan_set_gn_stats(int plan, int slot, float* stats, int channels_per_group, int channel_offset);
int sdf_gemm_run(int plan, void* stream);
int sdf_gemm_plan_destroy(int plan);

/* memory-bound companions (csrc/sd_ops.cu); x/y are NHWC fp16 with row strides ld* (elements, multiples of 8) */
/* GroupNorm(+SiLU), ldm/modules/diffusionmodules/util.py:214 / model.py:38.  stats: fp32 scratch [Nimg,G,2] (sums / sums of squares,
 * kept for sdf_groupnorm_backward). */
int sdf_groupnorm_forward(const void* x, int ldx, void* y, int ldy, int Nimg, int HW, int C, int G, const float* gamma, const float* beta,
                          float eps, int silu_act, float* stats, void* stream);
/* sdf_groupnorm_forward without its own memset: `stats` was zeroed by the caller (one memset for all norms of a launch list) */
int sdf_groupnorm_forward_prezeroed(const void* x, int ldx, void* y, int ldy, int Nimg, int HW, int C, int G, const float* gamma, const float* beta,
                                    float eps, int silu_act, float* stats, void* stream);
/* the normalise(+SiLU) pass alone, for statistics produced by sdf_gemm_plan_set_gn_stats */
int sdf_groupnorm_apply(const void* x, int ldx, void* y, int ldy, int Nimg, int HW, int C, int G, const float* gamma, const float* beta,
                        float eps, int silu_act, const float* stats, void* stream);
int sdf_groupnorm_backward(const void* x, int ldx, const void* dy, int ldd, void* dx, int ldo, int Nimg, int HW, int C, int G,
                           const float* gamma, const float* beta, float eps, int silu_act, const float* stats, float* bstats,
                           int accumulate, void* stream);
int sdf_layernorm_forward(const void* x, int ldx, void* y, int ldy, int rows, int C, const float* gamma, const float* beta, float eps, void* stream);
int sdf_softmax_rows(const void* x, void* y, long long rows, int cols, int ld, float scale, void* stream);
/* Fused multi-head attention forward, o[b,i,h*d:(h+1)*d] = softmax_j(scale * q[b,i,h].k[b,j,h]) v[b,j,h]; fp16, token-major
 * [B, tokens, heads*d] with row strides ldq/ldk/ldo (elements).  Replaces einsum + softmax + einsum of CrossAttention.forward,
 * ldm/modules/attention.py:170-193 (diffusers' attention processor under guidance/sd_utils.py:104).  d in {32, 40, 64, 80, 160}. */
int sdf_flash_attention(const void* q, const void* k, const void* v, void* o, int B, int heads, int n, int nkv, int d,
                        int ldq, int ldk, int ldo, float scale, void* stream);
int sdf_softmax_rows_backward(const void* p, const void* dp, void* ds, long long rows, int cols, int ld, float scale, void* stream);
int sdf_geglu(const void* x, int ldx, void* y, int ldy, long long rows, int inner, void* stream);
/* direct 3x3 convolution (stride 1, zero pad 1) with <= 4 real channels on its image side, and its data-gradient: the VAE encoder's conv_in
 * (ldm/modules/diffusionmodules/model.py:387) without the zero-padded k-blocks of the implicit GEMM.  w fp32 [Cout, Cin, 3, 3]. */
int sdf_conv3x3_small_cin_forward(const void* x, int ldx, const float* w, const float* bias /* may be NULL */, void* y, int ldy, int Nimg, int H, int W,
                                  int Cin, int Cout, void* stream);
int sdf_conv3x3_small_cin_dgrad(const void* dy, int ldd, const float* w, void* dx, int ldx, int Nimg, int H, int W, int Cin, int C, void* stream);
int sdf_upsample_nearest2(const void* x, int ldx, void* y, int ldy, int Nimg, int H, int W, int C, void* stream);
int sdf_im2col_s2(const void* x, int ldx, void* col, int Nimg, int H, int W, int C, int Ho, int Wo, int pad_top, int pad_left, void* stream);
int sdf_col2im_s2(const void* dcol, void* dx, int ldx, int Nimg, int H, int W, int C, int Ho, int Wo, int pad_top, int pad_left, void* stream);
int sdf_copy2d(const void* x, int ldx, void* y, int ldy, long long rows, int C, void* stream);
int sdf_add2d(const void* a, int lda, const void* b, int ldb, void* y, int ldy, long long rows, int C, void* stream);
int sdf_transpose2d(const void* x, int ldx, void* y, int ldy, int batch, int rows, int C, void* stream);
int sdf_timestep_embedding(const int* t, int B, int dim, void* out, int ldo, void* stream);

/* SDS glue (guidance/sd_utils.py:86-163): bilinear 64->512 resize (+ its adjoint), posterior sample + add_noise into the UNet input,
 * classifier-free guidance + w(t)(eps_hat - eps) + loss value + gradient wrt the VAE moments */
int sdf_bilinear_forward(const float* src, int B, int Cc, int h, int w, void* dst, int ldd, int H, int W, float a, float b, void* stream);
int sdf_bilinear_backward(const void* ddst, int ldd, int H, int W, float* dsrc, int B, int Cc, int h, int w, float a, void* stream);
int sdf_sds_prepare(const void* moments, int ldm, const float* latents_in, const float* eps_post, const float* noise, const int* t,
                    const float* alphas_cumprod, int Bimg, int HW, float* latents, void* x_in, int ldx, float vae_scale, void* stream);
int sdf_sds_grad(const void* eps, int lde, const float* noise, const int* t, const float* alphas_cumprod, int Bimg, int HW,
                 float guidance_scale, float grad_scale, const float* view_scale /* may be NULL: per-image factor (Zero123 angle scaling,
                 guidance/zero123_utils.py:124-126) */, const void* moments, int ldm, const float* eps_post, float vae_scale,
                 float* grad, void* d_moments, float* loss, void* stream);

/* ------------------------------------------------------------------ fused Adan + GradScaler protocol
 * replaces optimizer.py:102-258 (Adan.step, _single_tensor_adan) and the unscale / inf-check / skip of nerf/utils.py:1063-1067.
 * Per optimiser step: sdf_adan_begin(acc) ; sdf_adan_grad_norm(grad_i, ...) for every tensor ; sdf_adan_advance ; sdf_adan_step(...) for every tensor.
 * acc: device float[2] = (sum of squared unscaled gradients, non-finite flag).  Nothing is read back to the host. */
int sdf_adan_begin(float* acc, void* stream);
int sdf_adan_grad_norm(const float* grad, long long n, float inv_scale, float* acc, void* stream);
/* steps[0..n_groups) += 1 unless acc flags non-finite gradients: a skipped step advances neither the bias corrections nor the
 * first-step initialisation of neg_pre_grad (GradScaler.step does not call optimizer.step, nerf/utils.py:1066) */
int sdf_adan_advance(const float* acc, int* steps, int n_groups, void* stream);
/* step: 1-based executed-step count, read from the device word step_dev when that is not NULL.  param_half (optional): fp16 mirror of
 * the updated parameter.  ema (optional): shadow -= ema_one_minus_decay * (shadow - param_new) (torch_ema, nerf/utils.py:282-283).
 * zero_grad: clear grad after use — also on a skipped (non-finite) step, as optimizer.zero_grad() runs every iteration (nerf/utils.py:1043). */
int sdf_adan_step(float* param, float* grad, float* exp_avg, float* exp_avg_diff, float* exp_avg_sq, float* neg_pre_grad, long long n,
                  float beta1, float beta2, float beta3, int step, const int* step_dev, float lr, float weight_decay, float eps,
                  float max_grad_norm, int no_prox, float inv_scale, const float* acc, void* param_half /* may be NULL */,
                  float* ema /* may be NULL */, float ema_one_minus_decay, int zero_grad, void* stream);
/* torch_ema.ExponentialMovingAverage.update for one tensor (nerf/utils.py:1090-1091, once per epoch) */
int sdf_ema_update(float* shadow, const float* param, long long n, float one_minus_decay, void* stream);

/* ------------------------------------------------------------------ training-render glue (csrc/render_aux.cu)
 * The host-side arithmetic of nerf/renderer.py:run_cuda / nerf/utils.py:train_step that is neither an extension op nor the field:
 * replaces eager PyTorch launches of the reference, so that one SDS step has no host synchronisation (see sdf_b200/render.py). */

/* background colour + mix + layout: bg = sigmoid(bg_net(freq_encode(rays_d, 6))) (nerf/network_grid.py:141-147) or the constant bg_const[3]
 * when w1 == NULL; image = image_c + (1 - weights_sum) * bg (nerf/renderer.py:796-808); pred (optional) = [B, C, HW] planar copy, channel 3
 * = weights_sum (latent mode, nerf/utils.py:545-549).  half_round: round where fp16 autocast rounds.  bg [N,3] is kept for the backward. */
int sdf_background_forward(const float* rays_d, uint32_t N, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* bg_const, int half_round, const float* image_c, const float* weights_sum,
                           float* bg /* may be NULL */, float* image /* may be NULL */, float* pred /* may be NULL */, uint32_t HW, uint32_t C, void* stream);
/* g_image [N,3] and/or g_pred [B,C,HW] -> g_image_c [N,3], g_weights_sum [N]; bg_net gradients ACCUMULATED into gw1 [32,39] gb1 [32] gw2 [3,32] gb2 [3] */
int sdf_background_backward(const float* g_image /* may be NULL */, const float* g_pred /* may be NULL */, uint32_t HW, uint32_t C,
                            const float* rays_d, uint32_t N, const float* w1, const float* b1, const float* w2, const float* b2,
                            const float* bg_const, int half_round, const float* weights_sum, float* g_image_c, float* g_weights_sum,
                            float* gw1, float* gb1, float* gw2, float* gb2, void* stream);
/* out[0] = mean_m entropy(clamp(weights, 1e-5, 1 - 1e-5)) (nerf/utils.py:690-694); out[1] = mean_m weights * max(normal . normalize(dir), 0)^2
 * (nerf/renderer.py:741-743; 0 when normals == NULL); the mean runs over the live sample count *m_dev (<= M_cap).  scratch: 3 words. */
int sdf_render_regularizers_forward(const float* weights, const float* normals /* may be NULL */, const float* dirs, const int* m_dev,
                                    uint32_t M_cap, float* scratch, float* out, void* stream);
/* g_out: device float[2] = d loss / d out[0], d loss / d out[1] (each further scaled by its lambda); writes g_weights [M], g_normals [M,3] (weights detached there) */
int sdf_render_regularizers_backward(const float* g_out, float lambda_entropy, float lambda_orient, const float* weights,
                                     const float* normals /* may be NULL */, const float* dirs, const int* m_dev, uint32_t M_cap,
                                     float* g_weights, float* g_normals /* may be NULL */, void* stream);
/* occupancy refresh (nerf/renderer.py:1103-1149) without host round trips: jittered cell points in Morton order (noise uniform [0,1) [n,3]),
 * decayed max-update with running (sum, count) of valid cells in acc[2], bit packing against min(acc[0]/acc[1], density_thresh) read on the device */
int sdf_occupancy_points(const float* noise, uint32_t n, uint32_t grid_size, float bound_cas, float* xyzs, void* stream);
int sdf_occupancy_update(float* grid, const float* sigmas, uint32_t n, float decay, float* acc, void* stream);
int sdf_packbits_mean(const float* grid, uint32_t N, const float* acc, float density_thresh, uint8_t* bitfield, float* mean_out /* may be NULL */,
                      void* stream);
/* per-ray 3-vectors -> per-sample rows through rays[N,2] = (offset, count): what `light_d[flatten_rays]` does in nerf/renderer.py:735-737 */
int sdf_expand_ray_vec3(const float* values, const int* rays, uint32_t N, uint32_t cap, float* out, void* stream);
/* pinhole rays (nerf/utils.py:113-176, N = -1) of pixels first, first + stride, ... of each of the B poses [B,4,4] */
int sdf_get_rays(const float* poses, uint32_t B, uint32_t H, uint32_t W, float focal, float cx, float cy, uint32_t first, uint32_t stride,
                 float* rays_o, float* rays_d, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------------------
 * DMTet stage (BASELINE config C5): marching tetrahedra on a fixed lattice, mesh normals, mesh regularisers, differentiable rasterisation.
 * Replaces the eager PyTorch of nerf/renderer.py:94-174 (DMTet.__call__), :877-890 (normals), :176-254 (normal_consistency,
 * laplacian_smooth_loss) and the nvdiffrast calls of :893-903 (rasterize, interpolate).  Every output is capacity-sized with device-side
 * counts[4] = (vertices, faces, one-triangle tets, two-triangle tets): no host synchronisation.  Lattice topology: sdf_b200/tetgrid.py. */
long long sdf_dmtet_scratch_ints(int E, int F);
int sdf_dmtet_extract(const float* pos, const float* deform /* may be NULL */, float tet_grid_size, const float* sdf, const int* tets, const int* edges,
                      const int* tet_edges, int N, int F, int E, float* verts, int* vert_edge, int* faces, int* counts, int* scratch, void* stream);
int sdf_dmtet_extract_backward(const float* pos, const float* deform, float tet_grid_size, const float* sdf, const int* edges, const int* vert_edge,
                               const int* counts, int E, const float* d_verts, float* d_sdf, float* d_deform, void* stream);
int sdf_mesh_normals_forward(const float* verts, const int* faces, const int* counts, int vcap, int fcap, float* face_n, float* vert_n_raw, float* vert_n,
                             void* stream);
int sdf_mesh_normals_backward(const float* verts, const int* faces, const int* counts, int fcap, const float* vert_n_raw, const float* d_vert_n,
                              const float* d_face_n, float* d_verts, void* stream);
int sdf_mesh_halfedge_keys(const int* faces, const int* counts, int vcap, int fcap, long long* keys, int* face_of, void* stream);
int sdf_mesh_losses_forward(const long long* sorted_keys, const int* sorted_face_of, const int* counts, int vcap, int fcap, const float* face_n,
                            const float* verts, float* work, float* losses, void* stream);
int sdf_mesh_losses_backward(const long long* sorted_keys, const int* sorted_face_of, const int* counts, int vcap, int fcap, const float* face_n,
                             const float* work, const float* g /* device [2] */, float* d_face_n, float* d_verts, void* stream);
int sdf_mesh_clip_transform(const float* verts, const int* counts, int vcap, const float* mvp, float* clip, void* stream);
int sdf_mesh_rasterize(const float* clip, const int* faces, const int* counts, int fcap, const float* verts, const float* vert_n, int H, int W,
                       void* zbuf, float* rast, float* xyz, float* nrm, float* mask, void* stream);
int sdf_mesh_rasterize_backward(const float* rast, const float* clip, const int* faces, const float* verts, const float* vert_n, const float* mvp, int H,
                                int W, const float* d_xyz, const float* d_nrm, float* d_verts, float* d_vert_n, void* stream);

/* G-buffer shading (nerf/renderer.py:916-928; mode 0 albedo, 1 lambertian, 2 textureless, 3 normal; light = 3 device floats) into c4 [P,4] = (rgb, coverage),
 * and clamp(c4, 0, 1) split into sdf_background_forward's inputs (nerf/renderer.py:930-947). */
int sdf_mesh_shade_forward(const float* albedo, const float* nrm, const float* mask, const float* light, float ambient, int mode, int P, float* c4, void* stream);
int sdf_mesh_shade_backward(const float* g_c4, const float* albedo, const float* nrm, const float* mask, const float* light, float ambient, int mode, int P,
                            float* g_albedo, float* g_nrm, void* stream);
int sdf_mesh_c4_split(const float* c4, int P, float* image_c, float* weights_sum, void* stream);
int sdf_mesh_c4_split_backward(const float* g_image_c, const float* g_weights_sum, const float* c4, int P, float* g_c4, void* stream);

/* dr.antialias (nerf/renderer.py:930-931) restated: silhouette-edge coverage blending of the (rgb, coverage) image over horizontally / vertically
 * adjacent pixel pairs; face_adj [adj_faces,3] from sdf_mesh_face_adjacency (sorted half-edge keys + the sort permutation). */
int sdf_mesh_face_adjacency(const long long* sorted_keys, const int* order, const int* counts, int fcap, int* face_adj, void* stream);
int sdf_mesh_antialias_forward(const float* color, int C /* <= 8 */, const float* rast, const float* clip, const int* faces, const int* face_adj, int adj_faces,
                               int H, int W, float* out, void* stream);
/* position gradient ACCUMULATED into d_verts [.,3] through mvp, or into d_clip [.,4] when mvp == NULL */
int sdf_mesh_antialias_backward(const float* g_out, const float* color, int C, const float* rast, const float* clip, const int* faces, const int* face_adj,
                                int adj_faces, const float* mvp /* may be NULL */, int H, int W, float* g_color, float* d_pos /* may be NULL */, void* stream);
/* the unfused primitives with nvdiffrast's semantics (dr.rasterize / dr.interpolate, nerf/renderer.py:895-898), bound by the drop-in package
 * stable-dreamfusion_b200/nvdiffrast/torch so that the reference's own run_dmtet runs unchanged */
int sdf_mesh_rasterize_only(const float* clip, const int* faces, const int* counts, int fcap, int H, int W, void* zbuf, float* rast, void* stream);
int sdf_mesh_rasterize_uv_backward(const float* g_rast, const float* rast, const float* clip, const int* faces, int H, int W, float* d_clip, void* stream);
int sdf_mesh_interpolate_forward(const float* attr, int C, const float* rast, const int* faces, int P, float* out, void* stream);
int sdf_mesh_interpolate_backward(const float* g_out, const float* attr, int C, const float* rast, const int* faces, int P, float* d_attr /* may be NULL */,
                                  float* d_rast /* may be NULL */, void* stream);

/* d(albedo) / d(position) of the DMTet stage's texture lookup (nerf/renderer.py:905-912 with nerf/network_grid.py:68-79 and the grad_inputs path of
 * gridencoder/grid.py:77-100): MLP data-gradient back to the features, contracted with sdf_grid_encode_forward's dy_dx. */
int sdf_field_albedo_input_grad(const void* feat, const void* dy_dx, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                const float* b3, const float* g_albedo, const float* mask, int P, int L, float bound, float* d_xyz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDF_B200_H */
