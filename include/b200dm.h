/*
 * libb200dm — C ABI of the B200 (sm_100a) diffusion hot path.
 *
 * The reference (seungjunlee96/lightning-generative-models) has no native/FFI layer: the path sits
 * behind the Python classes Unet / GaussianDiffusion in models/generative/diffusion/ddpm.py.  Each
 * entry point below therefore cites the reference Python lines whose arithmetic it replaces; the
 * Python host mirror (lightning-generative-models_b200/b200dm) binds them with ctypes and re-creates
 * the reference's class surface on top.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named h_*;
 *   - the library never allocates, frees or synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*);
 *   - return 0 on success; <0 on error: -1 bad shape/alignment, -2 unsupported configuration,
 *     -3 workspace too small, -4 CUDA error (text via b200dm_last_error(), thread-local);
 *   - dtype: 0 = fp32 activations ("fp32 mode"), 1 = bf16 activations with fp32 accumulation;
 *   - activations are NHWC with an explicit pixel stride `ld` (elements) so that channel slices of a
 *     concat buffer are addressable without copies (replaces torch.cat, ddpm.py:459,462,468).
 */
#ifndef B200DM_H
#define B200DM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DM_OK 0
#define B200DM_ERR_SHAPE (-1)
#define B200DM_ERR_UNSUPPORTED (-2)
#define B200DM_ERR_WORKSPACE (-3)
#define B200DM_ERR_CUDA (-4)

#define B200DM_F32 0
#define B200DM_BF16 1

/* objective enum — GaussianDiffusion(objective=...), ddpm.py:540,562-566 */
#define B200DM_PRED_NOISE 0
#define B200DM_PRED_X0 1
#define B200DM_PRED_V 2

int b200dm_version(void);
const char* b200dm_last_error(void);
/* number of kernels launched by this library (all threads of the process) since the last reset */
int64_t b200dm_launch_count(void);
void b200dm_reset_launch_count(void);
/* 1 if the tcgen05/TMA path can be used on the current device (sm_100), else 0 */
int b200dm_tc_available(void);

/* ------------------------------------------------------------------------------------------------
 * Fused scheduler / loss kernels (HBM-bound, vectorised).  Images are NCHW fp32 [B, C*H*W].
 * `coef` tables are the fp32 schedule buffers registered at ddpm.py:601-662 (length T).
 * Noise: if `noise` != NULL it is read (exact-parity mode); otherwise it is generated in registers
 * with Philox4x32-10 keyed by (seed, stream_id) and counter = global element index / 4 + elem_offset/4,
 * so results do not depend on how a batch is sharded over GPUs.
 * ---------------------------------------------------------------------------------------------- */

/* Inputs of the forward-noising step, shared by q_sample and the loss (the loss kernel re-derives x0 and eps from
 * the same description instead of reading stored copies: 8 + 12 B per element in fp32 I/O).
 *   x0  = normalize ? 2*img - 1 : img                                          ddpm.py:945
 *   eps = noise ? noise[i] : Philox4x32-10 N(0,1) of element elem_offset + i    ddpm.py:881 (randn_like)
 *   eps += offset_strength * offset[b, c]      (offset != NULL)                 ddpm.py:889-891 (offset noise)
 * hw = pixels per channel plane (only needed with offset noise; hw % 4 == 0). */
typedef struct {
  const float* img;        /* [B, chw] fp32 */
  const int64_t* t;        /* [B] */
  const float* noise;      /* [B, chw] injected normals, or NULL */
  const float* offset;     /* [B, chw/hw] normals of the offset noise, or NULL */
  const float* sqrt_ac;    /* sqrt_alphas_cumprod [T] */
  const float* sqrt_1mac;  /* sqrt_one_minus_alphas_cumprod [T] */
  float offset_strength;
  int32_t normalize;
  int32_t B;
  int32_t reserved;
  int64_t chw, hw;
  uint64_t seed, stream_id, elem_offset;
} b200dm_noise_desc;

/* normalize + q_sample: x_t = sqrt_ac[t]*x0 + sqrt_1mac[t]*eps.
 * Replaces ddpm.py:945 (normalize), :881 (randn_like), :889-891 (offset noise), :869-876 (q_sample).
 * Optionally writes eps (noise_out) and the normalised x0 (x0_out). */
int b200dm_q_sample(const b200dm_noise_desc* d, float* x_t, float* noise_out, float* x0_out, void* stream);

/* target + MSE + loss weight + mean, and dL/d(model_out).  Replaces ddpm.py:911-925, :684-688.
 * `d` is the descriptor q_sample was called with.  loss_acc: fp32[1] accumulator (must be zeroed by the caller);
 * per-element grad d_out = 2*w[t]*(out-target)/(B*chw) written if d_out != NULL. */
int b200dm_loss_fwd_bwd(const b200dm_noise_desc* d, const float* model_out, const float* loss_weight,
                        float* loss_acc, float* d_out, int32_t objective, void* stream);

/* DDIM update for one (time, time_next) pair given the UNet output.
 * Replaces model_predictions(clip_x_start=True, rederive_pred_noise=True) ddpm.py:707-734 and the
 * update at :812-827.  coefficient scalars are computed on the host from the fp32 buffers exactly as
 * the reference does (`alpha_next.sqrt()`, `c`, `sigma`).  last != 0: x_next = x0 (time_next < 0).
 * x0_out optional.  x_next may be the same buffer as x_t (in-place update of the sampler state). */
int b200dm_ddim_step(const float* x_t, const float* model_out, const float* noise, float* x_next,
                     float* x0_out, float c_sqrt_ac, float c_sqrt_1mac, float c_sqrt_recip,
                     float c_sqrt_recipm1, float sqrt_alpha_next, float c, float sigma, int32_t last,
                     int32_t objective, int64_t n, uint64_t seed, uint64_t stream_id,
                     uint64_t elem_offset, void* stream);

/* DDPM ancestral step.  Replaces p_mean_variance + p_sample, ddpm.py:736-757 (x0 clamped to [-1,1],
 * posterior mean, + noise_std*z with noise_std = exp(0.5*logvar) computed by the host in fp32,
 * z = 0 when add_noise == 0 i.e. t == 0).  x_prev may be the same buffer as x_t. */
int b200dm_ddpm_step(const float* x_t, const float* model_out, const float* noise, float* x_prev,
                     float* x0_out, float c_sqrt_ac, float c_sqrt_1mac, float c_sqrt_recip,
                     float c_sqrt_recipm1, float coef1, float coef2, float noise_std,
                     int32_t add_noise, int32_t objective, int64_t n, uint64_t seed,
                     uint64_t stream_id, uint64_t elem_offset, void* stream);

/* N(0,1) fill with the same Philox stream (initial image of the samplers, ddpm.py:763,800). */
int b200dm_randn(float* out, int64_t n, uint64_t seed, uint64_t stream_id, uint64_t elem_offset,
                 void* stream);

/* unnormalize_to_zero_to_one, ddpm.py:86-87 */
int b200dm_unnormalize(const float* x, float* y, int64_t n, void* stream);

/* out[b] = [a[b] | b[b]] for NCHW fp32 tensors with chw_a / chw_b elements per sample (multiples of 4):
 * torch.cat((x_self_cond, x), dim=1) of a self-conditioned Unet.forward, ddpm.py:433-435. */
int b200dm_concat2_nchw(const float* a, const float* b, float* out, int32_t B, int64_t chw_a, int64_t chw_b,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Convolutions as implicit GEMM (ddpm.py:96,103,160,187,213-215,252-253,377,413).
 *   mode 0: k x k, stride 1, 'same' padding (k = 1 or 3); input [B,H,W,Cin]
 *   mode 1: pixel-unshuffle(2) + 1x1 == 2x2 stride-2 conv; input [B,2H,2W,Cin], 4 taps (p1,p2)
 *   mode 2: transpose of mode 1 (its data gradient): input [B,H,W,Cin], output [B,2H,2W,Cout],
 *           weight taps (p1,p2) of shape [Cout][Cin]
 *   mode 3: nn.Upsample(scale_factor=2, mode="nearest") followed by the 3x3 conv (ddpm.py:93-97) as ONE launch
 *           (tcgen05 path only): input [B,H,W,Cin] at the LOW resolution, output [B,2H,2W,Cout].
 *           Output phase (a,b) = (oy&1, ox&1) is a 2x2 conv over the source image (rows y+r+a-1, columns
 *           x+c+b-1 for tap (r,c)) whose weights are sums of the 3x3 taps that land on the same source pixel:
 *           16 instead of 36 multiply-adds per output.  Packed weights: [4 taps (r,c)][4 phases (a,b)][Cout][Cin]
 * Packed weights: [taps][Cout][Cin] in the activation dtype (Cin contiguous).
 * y = conv(x) + bias (+ res) (+ y if accumulate).  impl: 0 = SIMT (any shape, fp32 or bf16),
 * 1 = tcgen05/TMEM/TMA (bf16; Cin % 64 == 0, Cout % 64 == 0).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t dtype, mode, ksize, impl;
  int32_t B, H, W; /* OUTPUT spatial size for modes 0/1; INPUT spatial size for modes 2/3 */
  int32_t Cin, Cout;
  const void* x;
  int32_t x_ld;
  const void* w;
  const float* bias;
  void* y;
  int32_t y_ld;
  const void* res;
  int32_t res_ld;
  int32_t accumulate;
  /* tcgen05 path, mode 0: GroupNorm partial statistics of y fused into the epilogue (deterministic, no atomics).
   * gn_part: fp32 [B*H*W/seg][Cout/8][2] = (sum, sum of squares) per pixel slot of seg = min(32, H*W) pixels
   * and 8-channel chunk; b200dm_gn_fwd_pre adds the H*W/seg slots of a sample and the chunks of a group.
   * gn_groups > 0 with Cout/gn_groups a multiple of 8.  NULL = off. */
  float* gn_part;
  int32_t gn_groups;
} b200dm_conv_desc;

int b200dm_conv_fwd(const b200dm_conv_desc* d, void* stream);

/* Block.forward in one launch (ddpm.py:164-173, :189-200), tcgen05 path:
 *   y = SiLU( GroupNorm_groups(conv3x3(x) + bias) * (scale + 1) + shift ) (+ res)
 * `d` describes the 3x3 'same' conv (mode 0, ksize 3, impl 1; d->y receives the ACTIVATED output, d->res is added after
 * the SiLU, d->gn_part / d->accumulate must be 0).  A CTA (or thread-block cluster) owns one sample, the conv
 * accumulators stay in TMEM until the sample's statistics are known, and the norm is applied straight out of TMEM:
 * the conv -> norm intermediate never makes an HBM round trip and is never rounded to bf16 before the norm.
 *   gamma, beta [Cout]; film = FiLM (scale | shift) rows [B][film_ld] with scale at column c and shift at Cout + c,
 *   or NULL; stats (optional, out) [B][groups][2] = (mean, rstd); raw (optional, out) = conv(x) + bias in bf16
 *   [B,H,W,raw_ld] (what b200dm_gn_apply_bwd reads in training).
 * Images of 8x8 / 4x4 pixels take a second kernel: a 128-pixel tile holds whole samples and a 64-column sub-tile whole
 * groups, so the statistics are tile-local (no cluster).
 * b200dm_conv_gn_supported returns 1 when the layer fits (bf16, Cin % 64 == 0, 8 groups, and either 16 <= W <= 128,
 * H % 16 == 0, W % 8 == 0 with Cout in {64,128,256}, or 8x8 / 4x4 images with Cout in {128,256,512}), else 0: use
 * b200dm_conv_fwd + b200dm_gn_fwd_pre there. */
typedef struct {
  const float* gamma;
  const float* beta;
  const float* film;
  int32_t film_ld;
  int32_t groups;
  float eps;
  int32_t raw_ld;
  float* stats;
  void* raw;
} b200dm_gn_desc;

int b200dm_conv_gn_supported(const b200dm_conv_desc* d, const b200dm_gn_desc* gn);
int b200dm_conv_gn_fwd(const b200dm_conv_desc* d, const b200dm_gn_desc* gn, void* stream);

/* weight gradient of a mode-0/mode-1 conv:  dW[tap][co][ci] (+)= sum_pix dY[pix,co] * X[pix+tap,ci]
 * (fp32, same packed order as the master weights).  Uses `ws` (fp32, ws_bytes) for split-K partials
 * when needed.  impl as above. */
typedef struct {
  int32_t dtype, mode, ksize, impl;
  int32_t B, H, W;
  int32_t Cin, Cout;
  const void* x;
  int32_t x_ld;
  const void* dy;
  int32_t dy_ld;
  float* dw;
  int32_t accumulate;
  int32_t cin_valid; /* tcgen05 path: columns ci >= cin_valid are padding and are skipped (0 = all Cin) */
  /* element strides of dw: dw[tap*s_tap + co*s_co + ci*s_ci]; all 0 => tap-major packed
   * (s_tap = Cout*Cin, s_co = Cin, s_ci = 1).  The pixel-unshuffle conv keeps the reference layout
   * [Cout][c*4 + tap] (s_tap = 1, s_co = 4*Cin, s_ci = 4). */
  int64_t s_tap, s_co, s_ci;
} b200dm_wgrad_desc;

int b200dm_conv_wgrad(const b200dm_wgrad_desc* d, void* stream);

/* column sums: out[c] (+)= sum_rows x[row*ld + c]  (bias gradients) */
int b200dm_colsum(int32_t dtype, const void* x, int32_t ld, int64_t rows, int32_t C, float* out,
                  int32_t accumulate, void* stream);

/* n (1..16) independent column sums in one launch: items[i].out[c] += sum_rows items[i].x[row*ld + c] (always
 * accumulating; all items of one dtype; C and ld multiples of 8, x 16-byte aligned).  The bias gradients of the
 * convs of one gradient bucket (autograd of the `bias=True` nn.Conv2d's, ddpm.py:96,103,160,252-253,377,413). */
typedef struct {
  const void* x;
  float* out;
  int64_t rows;
  int32_t ld;
  int32_t C;
} b200dm_colsum_item;
int b200dm_colsum_batched(int32_t dtype, const b200dm_colsum_item* items, int32_t n, void* stream);

/* init_conv 7x7, pad 3 (ddpm.py:304,437): NCHW fp32 in -> NHWC out; weight OIHW fp32. */
/* Stem on tensor cores (bf16): P [B*H*W][KP] bf16 = im2col of the 7x7 patches of x (NCHW fp32), columns in the
 * OIHW order of init_conv.weight, zero-padded to KP (multiple of 64); wp [Cout][KP] bf16 = zero-padded weight
 * rows.  conv_fwd (mode 0, ksize 1, Cin = KP) over P is then the 7x7 conv and conv_wgrad (cin_valid = C*49,
 * s_co = C*49) its weight gradient. */
int b200dm_im2col7(const float* x, void* P, int32_t B, int32_t C, int32_t H, int32_t W, int32_t KP, void* stream);
int b200dm_pack_stem_weight(const float* w, void* wp, int32_t Cout, int32_t K, int32_t KP, void* stream);

/* The 7x7 stem as ONE tensor-core launch (csrc/stem_tc.cu): the im2col patches are built in shared memory instead of
 * HBM.  x NCHW fp32 [B][C][H][W], y NHWC bf16 (row pitch y_ld).
 * wp = b200dm_pack_stem_rows' bf16 rows [64][KP]: K order (channel, ky) filter rows of 7 taps + one zero, KP >= C*56.
 * init_conv, ddpm.py:304,437.  Needs W in {8,16,32,64,128}, H*W a multiple of 128, its buffers within 227 KiB. */
int b200dm_pack_stem_rows(const float* w, void* wp, int32_t Cout, int32_t C, int32_t KP, void* stream);
int b200dm_stem7_supported(int32_t B, int32_t C, int32_t H, int32_t W, int32_t KP, int32_t y_ld);
int b200dm_stem7_fwd(const float* x, const void* wp, const float* bias, void* y, int32_t y_ld, int32_t B, int32_t C,
                     int32_t H, int32_t W, int32_t KP, void* stream);
/* Weights of conv_fwd mode 3 (nearest-2x upsample + 3x3 conv in one launch, Upsample ddpm.py:93-97) from the fp32
 * master weight in [ky*3+kx][Cout][Cin] order: out = bf16 [4 taps][4 phases][Cout][Cin] (see b200dm_conv_desc). */
int b200dm_pack_upconv_weight(const float* w, void* out, int32_t Cout, int32_t Cin, void* stream);

int b200dm_init_conv_fwd(int32_t dtype, const float* x, const float* w, const float* bias, void* y,
                         int32_t y_ld, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cout,
                         void* stream);
int b200dm_init_conv_wgrad(int32_t dtype, const float* x, const void* dy, int32_t dy_ld, float* dw,
                           int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cout, void* stream);

/* final_conv 1x1 Cin -> C (ddpm.py:422,471): NHWC in -> NCHW fp32 out; weight [C][Cin] fp32. */
int b200dm_final_conv_fwd(int32_t dtype, const void* x, int32_t x_ld, const float* w,
                          const float* bias, float* y, int32_t B, int32_t HW, int32_t Cin, int32_t C,
                          void* stream);
/* dx NHWC (dtype), dw/db fp32 accumulated into (dw, db must be zeroed or hold prior grads) */
int b200dm_final_conv_bwd(int32_t dtype, const void* x, int32_t x_ld, const float* w,
                          const float* dy, void* dx, int32_t dx_ld, float* dw, float* db, int32_t B,
                          int32_t HW, int32_t Cin, int32_t C, void* stream);

/* nearest x2 upsample (ddpm.py:95) and its gradient (2x2 sum). NHWC. */
int b200dm_upsample2x_fwd(int32_t dtype, const void* x, int32_t x_ld, void* y, int32_t y_ld,
                          int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int b200dm_upsample2x_bwd(int32_t dtype, const void* dy, int32_t dy_ld, void* dx, int32_t dx_ld,
                          int32_t B, int32_t H, int32_t W, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm(8) + FiLM + SiLU (+ residual)  — Block.forward ddpm.py:164-173, ResnetBlock :189-200
 * ---------------------------------------------------------------------------------------------- */
/* per-(sample, group) mean and rstd (eps 1e-5) of x [B,HW,C]; stats = fp32 [B][G][2] */
int b200dm_gn_stats(int32_t dtype, const void* x, int32_t x_ld, float* stats, int32_t B, int32_t HW,
                    int32_t C, int32_t G, float eps, void* stream);
/* y = silu(((x-mean)*rstd*gamma+beta)*(1+scale)+shift) (+res).  film: fp32, scale at
 * film[b*film_ld + c], shift at film[b*film_ld + C + c]; NULL => no FiLM. */
int b200dm_gn_apply_fwd(int32_t dtype, const void* x, int32_t x_ld, const float* stats,
                        const float* gamma, const float* beta, const float* film, int32_t film_ld,
                        const void* res, int32_t res_ld, void* y, int32_t y_ld, int32_t B, int32_t HW,
                        int32_t C, int32_t G, void* stream);
/* stats + apply in ONE launch: a thread-block cluster per sample reduces (sum, sum of squares) through
 * distributed shared memory, writes stats [B][G][2] = (mean, rstd) for the backward pass and applies the
 * norm while the chunk is still in L2.  C <= 512; other shapes take the two-launch path internally. */
int b200dm_gn_fwd(int32_t dtype, const void* x, int32_t x_ld, float* stats, const float* gamma,
                  const float* beta, const float* film, int32_t film_ld, const void* res, int32_t res_ld,
                  void* y, int32_t y_ld, int32_t B, int32_t HW, int32_t C, int32_t G, float eps,
                  void* stream);
/* apply only: the statistics come from the conv epilogue (b200dm_conv_desc.gn_part): part = fp32
 * [B][slots][C/8][2] (sum, sum of squares per 8-channel chunk) with slots = H*W / min(32, H*W) per sample.  One pass, no cluster;
 * also writes stats [B][G][2] = (mean, rstd) for the backward pass. */
int b200dm_gn_fwd_pre(int32_t dtype, const void* x, int32_t x_ld, const float* part, int32_t slots, float* stats,
                      const float* gamma, const float* beta, const float* film, int32_t film_ld, const void* res,
                      int32_t res_ld, void* y, int32_t y_ld, int32_t B, int32_t HW, int32_t C, int32_t G, float eps,
                      void* stream);
/* backward.  C <= 512: one cluster launch per call (per-channel sums exchanged through distributed shared
 * memory; dgamma/dbeta/dbias reduced over the batch with fp32 atomics; sums/gmeans unused).  Otherwise
 * three launches inside:  (1) per-(b,c) sums of dz, dz*xnorm and x  (2) parameter / FiLM
 * grads + group means (+ the bias gradient of the conv that produced x, if dbias != NULL: the pixel
 * sum of dx follows in closed form from the sums)  (3) dx.  sums: fp32 workspace of
 * b200dm_gn_bwd_ws_floats(B, HW, C) floats (per-CTA partials, reduced deterministically);
 * gmeans: fp32 workspace [B][G][2].  dgamma/dbeta/dbias accumulate (+=); dfilm (same addressing as
 * film) is overwritten. */
int64_t b200dm_gn_bwd_ws_floats(int32_t B, int32_t HW, int32_t C);
int b200dm_gn_apply_bwd(int32_t dtype, const void* dy, int32_t dy_ld, const void* x, int32_t x_ld,
                        const float* stats, const float* gamma, const float* beta, const float* film,
                        int32_t film_ld, void* dx, int32_t dx_ld, float* dgamma, float* dbeta,
                        float* dfilm, float* dbias, float* sums, float* gmeans, int32_t B, int32_t HW,
                        int32_t C, int32_t G, void* stream);

/* RMSNorm (ddpm.py:107-113): y = x / max(||x||_2, 1e-12) * g * sqrt(C) (+ res) */
int b200dm_rmsnorm_fwd(int32_t dtype, const void* x, int32_t x_ld, const float* g, const void* res,
                       int32_t res_ld, void* y, int32_t y_ld, int64_t rows, int32_t C, void* stream);
/* dx = rmsnorm'(dy) (+ res);  dg += ... */
int b200dm_rmsnorm_bwd(int32_t dtype, const void* dy, int32_t dy_ld, const void* x, int32_t x_ld,
                       const float* g, const void* res, int32_t res_ld, void* dx, int32_t dx_ld,
                       float* dg, int64_t rows, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention cores; qkv is [B, n, 384] (q | k | v, each heads*32 channels, head-major), heads = 4.
 * ---------------------------------------------------------------------------------------------- */
/* LinearAttention core, ddpm.py:222-238.  mem_kv fp32 [2][4][32][4].  ctx: fp32 [B][4][32][32];
 * kstat: fp32 [B][4][32][2] (max, sum of exp) saved for backward.  out [B,n,128]. */
int b200dm_linattn_fwd(int32_t dtype, const void* qkv, int32_t qkv_ld, const float* mem_kv,
                       float* ctx, float* kstat, void* out, int32_t out_ld, int32_t B, int32_t n,
                       void* stream);
/* dctx: fp32 workspace [B][4][33][32] (the 32x32 context gradient plus one row of per-d dot products);
 * dmem_kv accumulates (+=) */
int b200dm_linattn_bwd(int32_t dtype, const void* dout, int32_t dout_ld, const void* qkv,
                       int32_t qkv_ld, const float* mem_kv, const float* ctx, const float* kstat,
                       float* dctx, void* dqkv, int32_t dqkv_ld, float* dmem_kv, int32_t B,
                       int32_t n, void* stream);

/* LinearAttention block for inference in four launches that never store q, k, v (csrc/linattn_tc.cu):
 *   y = RMSNorm(to_out(LinearAttention(to_qkv(RMSNorm(x))))) + x
 * reference ddpm.py:205-238 (LinearAttention.forward), :184-191 (RMSNorm), :449,:464 (`attn(x) + x`).  bf16 only.
 * wqkv is to_qkv.weight with the first RMSNorm's gain folded in (b200dm_pack_linattn_qkv); wout is to_out.0.weight as
 * bf16 [C][128]; gout is to_out.1.g.  n must be a multiple of 128 and C 64 or 128 (b200dm_linattn_block_supported);
 * ws holds b200dm_linattn_block_ws_floats(B, n, C) floats of scratch (16-byte aligned). */
typedef struct {
  int32_t B, n, C;
  int32_t x_ld, y_ld;
  int32_t reserved;
  const void* x;
  void* y;
  const void* wqkv;
  const void* wout;
  const float* bout;
  const float* gout;
  const float* mem_kv;
  float* ws;
} b200dm_linattn_block_desc;
int b200dm_pack_linattn_qkv(const float* w, const float* g, void* out, int32_t C, void* stream);
int64_t b200dm_linattn_block_ws_floats(int32_t B, int32_t n, int32_t C);
int b200dm_linattn_block_supported(const b200dm_linattn_block_desc* d);
int b200dm_linattn_block_fwd(const b200dm_linattn_block_desc* d, void* stream);
/* Attention + Attend.forward (math branch), ddpm.py:255-271, models/modules/attend.py:111-126.
 * mem_kv fp32 [2][4][4][32]; n <= 64. */
int b200dm_attn_fwd(int32_t dtype, const void* qkv, int32_t qkv_ld, const float* mem_kv, void* out,
                    int32_t out_ld, int32_t B, int32_t n, void* stream);
int b200dm_attn_bwd(int32_t dtype, const void* dout, int32_t dout_ld, const void* qkv,
                    int32_t qkv_ld, const float* mem_kv, void* dqkv, int32_t dqkv_ld,
                    float* dmem_kv, int32_t B, int32_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Time embedding (ddpm.py:119-132, :328-333) and small fp32 linears (ddpm.py:179-183)
 * ---------------------------------------------------------------------------------------------- */
/* emb[b, :] = [sin(t*f_i) | cos(t*f_i)], f_i = exp(-i*ln(theta)/(dim/2-1)) */
int b200dm_sinusoidal(const int64_t* t, float* emb, int32_t B, int32_t dim, float theta,
                      void* stream);
/* Y = act(X W^T + b); X [M,K], W [N,K], Y [M,N] fp32.  act: 0 none, 1 GELU(erf), 2 SiLU.
 * pre (optional) receives the pre-activation. */
int b200dm_linear_fwd(const float* X, const float* W, const float* b, float* Y, float* pre, int32_t M,
                      int32_t N, int32_t K, int32_t act, void* stream);
/* dPre = dY * act'(pre) (in place into dY when act != 0), dX = dPre W, dW += dPre^T X, db += colsum */
int b200dm_linear_bwd(const float* X, const float* W, const float* pre, float* dY, float* dX,
                      float* dW, float* db, int32_t M, int32_t N, int32_t K, int32_t act,
                      void* stream);

/* The same for a column range of a layer without activation: dY = [M, N] window of a [M, ldy] matrix, W / dW = the
 * matching N rows ([N, K]); dX = dY W is overwritten, or added to when dx_accumulate != 0; no bias gradient. */
int b200dm_linear_bwd_cols(const float* X, const float* W, const float* dY, int32_t ldy, float* dX,
                           int32_t dx_accumulate, float* dW, int32_t M, int32_t N, int32_t K, void* stream);

/* Diagnostic (scripts/umma_rate.py): cycles for `iters` x 8 tcgen05.mma (M=128, N=n_tile, K=16) issued by one
 * thread per CTA on resident shared-memory operands; mode 0 = K-major SW128, 1 = shifted halo view,
 * 2 = MN-major (wgrad).  out_cycles: int64 [ctas]. */
int b200dm_debug_umma_rate(int32_t n_tile, int32_t iters, int32_t mode, int32_t ctas, long long* out_cycles,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight packing, optimiser, EMA
 * ---------------------------------------------------------------------------------------------- */
/* master fp32 w[tap*s_tap + co*s_co + ci*s_ci] -> fwd-packed [taps][Cout][Cin] (dtype) and, if
 * wt != NULL, the dgrad operand [taps][Cin][Cout] with taps reversed when flip != 0. */
int b200dm_pack_conv_weight(int32_t dtype, const float* w, void* wf, void* wt, int32_t taps,
                            int32_t Cout, int32_t Cin, int32_t flip, int64_t s_tap, int64_t s_co,
                            int64_t s_ci, void* stream);
/* the same for every conv of the network in one launch.  `table` is a DEVICE array of entries; entry i
 * owns the CTAs [tile_begin, tile_begin + taps*tiles_co*tiles_ci) with tiles_* = ceil(C* / 64). */
typedef struct {
  const float* w;
  void* wf;
  void* wt;
  int32_t taps, Cout, Cin, flip;
  int64_t s_tap, s_co, s_ci;
  int32_t tile_begin, tiles_ci, tiles_co, reserved;
} b200dm_pack_entry;
int b200dm_pack_conv_weights_batched(int32_t dtype, const b200dm_pack_entry* table, int32_t n_entries,
                                     int32_t total_tiles, void* stream);
/* a contiguous run of table entries only (`entries` = &table[first]; their tile_begin keep the numbering of the
 * full table, `tile_first` = entries[0].tile_begin, `n_tiles` = tiles owned by the run): lets the optimiser
 * re-pack one gradient bucket's convs as soon as their weights are updated. */
int b200dm_pack_conv_weights_range(int32_t dtype, const b200dm_pack_entry* entries, int32_t n_entries,
                                   int32_t tile_first, int32_t n_tiles, void* stream);
/* fused Adam over a flat fp32 arena (torch.optim.Adam semantics, ddpm.py:1053-1059):
 * grad_scale multiplies g first (DDP mean).  step is the 1-based step count. */
int b200dm_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                     void* stream);
/* the same with a bounded grid (`ctas_per_sm` x SMs CTAs of 256 threads): for running the update of one gradient
 * bucket on a second stream behind backward without taking every thread slot of the machine. */
int b200dm_adam_step_bg(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                        int32_t ctas_per_sm, void* stream);
/* ema = ema + (1-decay)*(online-ema)  (ema_pytorch lerp), or copy when decay == 0 */
int b200dm_ema_update(float* ema, const float* online, int64_t n, float decay, void* stream);
int b200dm_fill_f32(float* p, int64_t n, float value, void* stream);
/* Launches issued after this call size their persistent grids for (SM count - n) SMs: set while a collective with n
 * CTAs (NCCL_MAX_CTAS / ProcessGroupNCCL max_ctas) runs next to the backward pass, 0 otherwise.  Process-wide. */
int b200dm_set_reserved_sms(int32_t n);
/* fp32 <-> bf16 copies of a gradient bucket for the opt-in bf16 all-reduce (n % 4 == 0). */
int b200dm_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);
int b200dm_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DM_H */
