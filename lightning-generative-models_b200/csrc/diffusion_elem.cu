// Fused scheduler / loss kernels: q_sample, loss(+grad), DDIM step, DDPM step, Philox randn.
// HBM-bound; one thread = 4 consecutive elements = one float4 per tensor = one Philox block.
// Arithmetic mirrors the reference's op order (mul, mul, add — no FMA contraction) so the fp32
// mode agrees with torch to the last bits.  Reference: models/generative/diffusion/ddpm.py
// :673-694 (predict_*), :707-757 (model_predictions, p_sample), :812-827 (DDIM), :869-925.
#include "common.cuh"

namespace b200dm {

constexpr int kElemThreads = 256;

static inline int elem_grid(int64_t nvec) {
  int64_t blocks = (nvec + kElemThreads - 1) / kElemThreads;
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float4 ld4(const float* p, int64_t i) {
  return __ldg(reinterpret_cast<const float4*>(p) + i);
}
__device__ __forceinline__ void st4(float* p, int64_t i, float4 v) {
  reinterpret_cast<float4*>(p)[i] = v;
}
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }


// ---------------------------------------------------------------------------------------------
// Forward noising inputs shared by q_sample and the loss: the SAME descriptor is given to both, so the loss kernel
// re-derives x0 (normalize) and eps (Philox block of the element, or the injected tensor, + offset noise) instead
// of reading copies that q_sample would have had to store: q_sample moves 8 B and the loss 12 B per element in
// fp32 I/O (SURVEY 8d's fp32-mode byte budgets) instead of 16 + 16.
struct NoiseIn {
  const float* img;
  const int64_t* t;
  const float* noise;      // injected eps, or nullptr -> Philox
  const float* offset;     // [B * C] per-(sample, channel) normals of the offset noise (ddpm.py:889-891), or nullptr
  const float* sqrt_ac;
  const float* sqrt_1mac;
  float offset_strength;
  int normalize;
  int64_t chw4, hw4;       // float4 vectors per sample / per channel plane
  uint64_t seed, stream_id, vec_offset;
};

__device__ __forceinline__ void noise_in_load(const NoiseIn& d, int64_t i, int b, float4& x, float4& e) {
  x = ld4(d.img, i);
  if (d.normalize) {
    x.x = sub_(mul_(x.x, 2.f), 1.f); x.y = sub_(mul_(x.y, 2.f), 1.f);
    x.z = sub_(mul_(x.z, 2.f), 1.f); x.w = sub_(mul_(x.w, 2.f), 1.f);
  }
  e = d.noise ? ld4(d.noise, i) : Philox::normal4(d.seed, d.vec_offset + (uint64_t)i, d.stream_id);
  if (d.offset) {          // noise += strength * offset[b, c]   (one channel per float4: hw % 4 == 0)
    const int64_t c = (i - (int64_t)b * d.chw4) / d.hw4;
    const float o = mul_(d.offset_strength, __ldg(d.offset + (int64_t)b * (d.chw4 / d.hw4) + c));
    e.x = add_(e.x, o); e.y = add_(e.y, o); e.z = add_(e.z, o); e.w = add_(e.w, o);
  }
}

__global__ void __launch_bounds__(kElemThreads)
q_sample_kernel(const NoiseIn d, float* __restrict__ x_t, float* __restrict__ noise_out,
                float* __restrict__ x0_out, int64_t nvec) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / d.chw4);
    int64_t tb = d.t[b];
    float ca = __ldg(d.sqrt_ac + tb), cb = __ldg(d.sqrt_1mac + tb);
    float4 x, e;
    noise_in_load(d, i, b, x, e);
    float4 o;
    o.x = add_(mul_(ca, x.x), mul_(cb, e.x));
    o.y = add_(mul_(ca, x.y), mul_(cb, e.y));
    o.z = add_(mul_(ca, x.z), mul_(cb, e.z));
    o.w = add_(mul_(ca, x.w), mul_(cb, e.w));
    st4(x_t, i, o);
    if (noise_out) st4(noise_out, i, e);
    if (x0_out) st4(x0_out, i, x);
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float target_of(int objective, float x0, float e, float ca, float cb) {
  if (objective == B200DM_PRED_NOISE) return e;
  if (objective == B200DM_PRED_X0) return x0;
  return sub_(mul_(ca, e), mul_(cb, x0));  // predict_v, ddpm.py:684-688
}

__global__ void __launch_bounds__(kElemThreads)
loss_kernel(const NoiseIn dsc, const float* __restrict__ out, const float* __restrict__ loss_weight,
            float* __restrict__ loss_acc, float* __restrict__ d_out, int64_t nvec, int objective,
            float inv_count) {
  pdl_prologue();
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / dsc.chw4);
    int64_t tb = dsc.t[b];
    float ca = __ldg(dsc.sqrt_ac + tb), cb = __ldg(dsc.sqrt_1mac + tb), w = __ldg(loss_weight + tb);
    float4 o = ld4(out, i), x, e;
    noise_in_load(dsc, i, b, x, e);
    float4 d;
    d.x = o.x - target_of(objective, x.x, e.x, ca, cb);
    d.y = o.y - target_of(objective, x.y, e.y, ca, cb);
    d.z = o.z - target_of(objective, x.z, e.z, ca, cb);
    d.w = o.w - target_of(objective, x.w, e.w, ca, cb);
    acc += w * (d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w);
    if (d_out) {
      float s = 2.f * w * inv_count;
      st4(d_out, i, make_float4(s * d.x, s * d.y, s * d.z, s * d.w));
    }
  }
  __shared__ float red[kElemThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kElemThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss_acc, v * inv_count);
  }
}

// ---------------------------------------------------------------------------------------------
struct StepCoef {
  float sac, s1mac, sr, srm1;  // sqrt_alphas_cumprod[t], sqrt_one_minus..., sqrt_recip..., sqrt_recipm1...
};

// model_predictions (ddpm.py:707-734): x0 (optionally clamped) and eps re-derived from it.
__device__ __forceinline__ void predict(int objective, const StepCoef& c, float x, float out,
                                        bool clip, bool rederive, float& x0, float& eps) {
  if (objective == B200DM_PRED_NOISE) {
    eps = out;
    x0 = sub_(mul_(c.sr, x), mul_(c.srm1, out));
    if (clip) {
      x0 = clamp1(x0);
      if (rederive) eps = __fdiv_rn(sub_(mul_(c.sr, x), x0), c.srm1);
    }
  } else {
    x0 = (objective == B200DM_PRED_X0) ? out : sub_(mul_(c.sac, x), mul_(c.s1mac, out));
    if (clip) x0 = clamp1(x0);
    eps = __fdiv_rn(sub_(mul_(c.sr, x), x0), c.srm1);
  }
}

// x_t and x_next may be the SAME buffer (the samplers update their state in place): neither is __restrict__ and x_t is
// read with a plain (coherent) load; every thread reads vector i before it writes vector i.
__global__ void __launch_bounds__(kElemThreads)
ddim_step_kernel(const float* x_t, const float* __restrict__ out,
                 const float* __restrict__ noise, float* x_next,
                 float* __restrict__ x0_out, StepCoef c, float sqrt_an, float cc, float sigma,
                 int last, int objective, int64_t nvec, uint64_t seed, uint64_t stream_id,
                 uint64_t vec_offset) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(x_t)[i], o = ld4(out, i);
    float xs[4] = {x.x, x.y, x.z, x.w}, os[4] = {o.x, o.y, o.z, o.w}, r[4], x0s[4];
    float zs[4] = {0.f, 0.f, 0.f, 0.f};
    if (!last && sigma != 0.f) {
      float4 z = noise ? ld4(noise, i) : Philox::normal4(seed, vec_offset + (uint64_t)i, stream_id);
      zs[0] = z.x; zs[1] = z.y; zs[2] = z.z; zs[3] = z.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x0, eps;
      predict(objective, c, xs[k], os[k], true, true, x0, eps);
      x0s[k] = x0;
      // img = x_start*alpha_next.sqrt() + c*pred_noise + sigma*noise   (ddpm.py:827)
      r[k] = last ? x0 : add_(add_(mul_(x0, sqrt_an), mul_(cc, eps)), mul_(sigma, zs[k]));
    }
    st4(x_next, i, make_float4(r[0], r[1], r[2], r[3]));
    if (x0_out) st4(x0_out, i, make_float4(x0s[0], x0s[1], x0s[2], x0s[3]));
  }
}

__global__ void __launch_bounds__(kElemThreads)
ddpm_step_kernel(const float* x_t, const float* __restrict__ out,
                 const float* __restrict__ noise, float* x_prev,
                 float* __restrict__ x0_out, StepCoef c, float coef1, float coef2, float noise_std,
                 int add_noise, int objective, int64_t nvec, uint64_t seed, uint64_t stream_id,
                 uint64_t vec_offset) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(x_t)[i], o = ld4(out, i);
    float xs[4] = {x.x, x.y, x.z, x.w}, os[4] = {o.x, o.y, o.z, o.w}, r[4], x0s[4];
    float zs[4] = {0.f, 0.f, 0.f, 0.f};
    if (add_noise) {
      float4 z = noise ? ld4(noise, i) : Philox::normal4(seed, vec_offset + (uint64_t)i, stream_id);
      zs[0] = z.x; zs[1] = z.y; zs[2] = z.z; zs[3] = z.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x0, eps;
      // p_mean_variance: model_predictions without clipping, then x_start.clamp_(-1, 1)  (:736-746)
      predict(objective, c, xs[k], os[k], false, false, x0, eps);
      x0 = clamp1(x0);
      x0s[k] = x0;
      float mean = add_(mul_(coef1, x0), mul_(coef2, xs[k]));  // q_posterior :696-705
      r[k] = add_noise ? add_(mean, mul_(noise_std, zs[k])) : mean;
    }
    st4(x_prev, i, make_float4(r[0], r[1], r[2], r[3]));
    if (x0_out) st4(x0_out, i, make_float4(x0s[0], x0s[1], x0s[2], x0s[3]));
  }
}

__global__ void __launch_bounds__(kElemThreads)
randn_kernel(float* __restrict__ out, int64_t nvec, uint64_t seed, uint64_t stream_id,
             uint64_t vec_offset) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x)
    st4(out, i, Philox::normal4(seed, vec_offset + (uint64_t)i, stream_id));
}

__global__ void __launch_bounds__(kElemThreads)
unnormalize_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t nvec) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = ld4(x, i);
    st4(y, i, make_float4(mul_(add_(v.x, 1.f), 0.5f), mul_(add_(v.y, 1.f), 0.5f),
                          mul_(add_(v.z, 1.f), 0.5f), mul_(add_(v.w, 1.f), 0.5f)));
  }
}

// out[b] = [a[b] | b[b]] along the channel axis of NCHW tensors (torch.cat((x_self_cond, x), dim=1), ddpm.py:435)
__global__ void concat2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                               int64_t n4, int64_t qa, int64_t qb) {
  pdl_prologue();
  const int64_t tot = qa + qb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / tot, r = i - s * tot;
    const float4 v = r < qa ? reinterpret_cast<const float4*>(a)[s * qa + r]
                            : reinterpret_cast<const float4*>(b)[s * qb + (r - qa)];
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

__global__ void fill_kernel(float* __restrict__ p, int64_t n, float v) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

}  // namespace b200dm

using namespace b200dm;

#define CHECK_VEC(n, what) \
  B200DM_REQUIRE((n) > 0 && (n) % 4 == 0, B200DM_ERR_SHAPE, what ": element count %lld must be a positive multiple of 4", (long long)(n))
#define CHECK_ALIGN16(p, what) \
  B200DM_REQUIRE(((uintptr_t)(p) & 15) == 0, B200DM_ERR_SHAPE, what ": pointer not 16-byte aligned")

static int noise_in_from(const b200dm_noise_desc* d, NoiseIn* o, const char* what) {
  B200DM_REQUIRE(d != nullptr && d->B > 0, B200DM_ERR_SHAPE, "%s: empty batch", what);
  B200DM_REQUIRE(d->chw > 0 && d->chw % 4 == 0, B200DM_ERR_SHAPE, "%s: chw=%lld must be a positive multiple of 4", what,
                 (long long)d->chw);
  B200DM_REQUIRE(d->elem_offset % 4 == 0, B200DM_ERR_SHAPE, "%s: elem_offset must be a multiple of 4", what);
  B200DM_REQUIRE(((uintptr_t)d->img & 15) == 0 && ((uintptr_t)d->noise & 15) == 0, B200DM_ERR_SHAPE,
                 "%s: pointers must be 16-byte aligned", what);
  B200DM_REQUIRE(d->img && d->t && d->sqrt_ac && d->sqrt_1mac, B200DM_ERR_SHAPE, "%s: null input", what);
  if (d->offset)
    B200DM_REQUIRE(d->hw > 0 && d->hw % 4 == 0 && d->chw % d->hw == 0, B200DM_ERR_SHAPE,
                   "%s: offset noise needs hw %% 4 == 0 and chw %% hw == 0 (hw=%lld)", what, (long long)d->hw);
  o->img = d->img; o->t = d->t; o->noise = d->noise;
  o->offset = (d->offset && d->offset_strength != 0.f) ? d->offset : nullptr;
  o->sqrt_ac = d->sqrt_ac; o->sqrt_1mac = d->sqrt_1mac;
  o->offset_strength = d->offset_strength; o->normalize = d->normalize;
  o->chw4 = d->chw / 4; o->hw4 = d->hw > 0 ? d->hw / 4 : d->chw / 4;
  o->seed = d->seed; o->stream_id = d->stream_id; o->vec_offset = d->elem_offset / 4;
  return B200DM_OK;
}

extern "C" int b200dm_q_sample(const b200dm_noise_desc* d, float* x_t, float* noise_out, float* x0_out,
                               void* stream) {
  NoiseIn in;
  int rc = noise_in_from(d, &in, "q_sample");
  if (rc) return rc;
  CHECK_ALIGN16(x_t, "q_sample x_t"); CHECK_ALIGN16(noise_out, "q_sample noise_out");
  CHECK_ALIGN16(x0_out, "q_sample x0_out");
  B200DM_REQUIRE(x_t != nullptr, B200DM_ERR_SHAPE, "q_sample: x_t is null");
  int64_t nvec = (int64_t)d->B * in.chw4;
  launch_k(q_sample_kernel, elem_grid(nvec), kElemThreads, 0, (cudaStream_t)stream, in, x_t, noise_out, x0_out, nvec);
  count_launch();
  return check_launch("q_sample");
}

extern "C" int b200dm_loss_fwd_bwd(const b200dm_noise_desc* d, const float* model_out, const float* loss_weight,
                                   float* loss_acc, float* d_out, int32_t objective, void* stream) {
  NoiseIn in;
  int rc = noise_in_from(d, &in, "loss");
  if (rc) return rc;
  B200DM_REQUIRE(objective >= 0 && objective <= 2, B200DM_ERR_UNSUPPORTED, "loss: unknown objective %d", objective);
  B200DM_REQUIRE(model_out && loss_weight && loss_acc, B200DM_ERR_SHAPE, "loss: null input");
  CHECK_ALIGN16(model_out, "loss out"); CHECK_ALIGN16(d_out, "loss d_out");
  int64_t nvec = (int64_t)d->B * in.chw4;
  float inv_count = 1.f / ((float)d->B * (float)d->chw);
  launch_k(loss_kernel, elem_grid(nvec), kElemThreads, 0, (cudaStream_t)stream, in, model_out, loss_weight, loss_acc,
           d_out, nvec, objective, inv_count);
  count_launch();
  return check_launch("loss_fwd_bwd");
}

extern "C" int b200dm_ddim_step(const float* x_t, const float* model_out, const float* noise,
                                float* x_next, float* x0_out, float c_sqrt_ac, float c_sqrt_1mac,
                                float c_sqrt_recip, float c_sqrt_recipm1, float sqrt_alpha_next,
                                float c, float sigma, int32_t last, int32_t objective, int64_t n,
                                uint64_t seed, uint64_t stream_id, uint64_t elem_offset, void* stream) {
  CHECK_VEC(n, "ddim_step");
  B200DM_REQUIRE(objective >= 0 && objective <= 2, B200DM_ERR_UNSUPPORTED, "ddim_step: unknown objective %d", objective);
  B200DM_REQUIRE(elem_offset % 4 == 0, B200DM_ERR_SHAPE, "ddim_step: elem_offset must be a multiple of 4");
  CHECK_ALIGN16(x_t, "ddim x_t"); CHECK_ALIGN16(model_out, "ddim out"); CHECK_ALIGN16(x_next, "ddim x_next");
  CHECK_ALIGN16(noise, "ddim noise"); CHECK_ALIGN16(x0_out, "ddim x0_out");
  StepCoef sc{c_sqrt_ac, c_sqrt_1mac, c_sqrt_recip, c_sqrt_recipm1};
  launch_k(ddim_step_kernel, elem_grid(n / 4), kElemThreads, 0, (cudaStream_t)stream, 
      x_t, model_out, noise, x_next, x0_out, sc, sqrt_alpha_next, c, sigma, last, objective, n / 4,
      seed, stream_id, elem_offset / 4);
  count_launch();
  return check_launch("ddim_step");
}

extern "C" int b200dm_ddpm_step(const float* x_t, const float* model_out, const float* noise,
                                float* x_prev, float* x0_out, float c_sqrt_ac, float c_sqrt_1mac,
                                float c_sqrt_recip, float c_sqrt_recipm1, float coef1, float coef2,
                                float noise_std, int32_t add_noise, int32_t objective, int64_t n,
                                uint64_t seed, uint64_t stream_id, uint64_t elem_offset, void* stream) {
  CHECK_VEC(n, "ddpm_step");
  B200DM_REQUIRE(objective >= 0 && objective <= 2, B200DM_ERR_UNSUPPORTED, "ddpm_step: unknown objective %d", objective);
  B200DM_REQUIRE(elem_offset % 4 == 0, B200DM_ERR_SHAPE, "ddpm_step: elem_offset must be a multiple of 4");
  CHECK_ALIGN16(x_t, "ddpm x_t"); CHECK_ALIGN16(model_out, "ddpm out"); CHECK_ALIGN16(x_prev, "ddpm x_prev");
  CHECK_ALIGN16(noise, "ddpm noise"); CHECK_ALIGN16(x0_out, "ddpm x0_out");
  StepCoef sc{c_sqrt_ac, c_sqrt_1mac, c_sqrt_recip, c_sqrt_recipm1};
  launch_k(ddpm_step_kernel, elem_grid(n / 4), kElemThreads, 0, (cudaStream_t)stream, 
      x_t, model_out, noise, x_prev, x0_out, sc, coef1, coef2, noise_std, add_noise, objective, n / 4,
      seed, stream_id, elem_offset / 4);
  count_launch();
  return check_launch("ddpm_step");
}

extern "C" int b200dm_randn(float* out, int64_t n, uint64_t seed, uint64_t stream_id,
                            uint64_t elem_offset, void* stream) {
  CHECK_VEC(n, "randn");
  B200DM_REQUIRE(elem_offset % 4 == 0, B200DM_ERR_SHAPE, "randn: elem_offset must be a multiple of 4");
  CHECK_ALIGN16(out, "randn out");
  launch_k(randn_kernel, elem_grid(n / 4), kElemThreads, 0, (cudaStream_t)stream, out, n / 4, seed, stream_id,
                                                                           elem_offset / 4);
  count_launch();
  return check_launch("randn");
}

extern "C" int b200dm_unnormalize(const float* x, float* y, int64_t n, void* stream) {
  CHECK_VEC(n, "unnormalize");
  CHECK_ALIGN16(x, "unnormalize x"); CHECK_ALIGN16(y, "unnormalize y");
  launch_k(unnormalize_kernel, elem_grid(n / 4), kElemThreads, 0, (cudaStream_t)stream, x, y, n / 4);
  count_launch();
  return check_launch("unnormalize");
}

extern "C" int b200dm_concat2_nchw(const float* a, const float* b, float* out, int32_t B, int64_t chw_a,
                                   int64_t chw_b, void* stream) {
  B200DM_REQUIRE(a && b && out && B > 0 && chw_a > 0 && chw_b > 0, B200DM_ERR_SHAPE, "concat2_nchw: empty input");
  B200DM_REQUIRE(chw_a % 4 == 0 && chw_b % 4 == 0, B200DM_ERR_SHAPE, "concat2_nchw: C*H*W must be a multiple of 4");
  CHECK_ALIGN16(a, "concat2_nchw a"); CHECK_ALIGN16(b, "concat2_nchw b"); CHECK_ALIGN16(out, "concat2_nchw out");
  const int64_t n4 = (int64_t)B * (chw_a + chw_b) / 4;
  launch_k(concat2_kernel, elem_grid(n4), kElemThreads, 0, (cudaStream_t)stream, a, b, out, n4, chw_a / 4, chw_b / 4);
  count_launch();
  return check_launch("concat2_nchw");
}

extern "C" int b200dm_fill_f32(float* p, int64_t n, float value, void* stream) {
  if (n <= 0) return B200DM_OK;
  launch_k(fill_kernel, elem_grid(n), kElemThreads, 0, (cudaStream_t)stream, p, n, value);
  count_launch();
  return check_launch("fill_f32");
}
