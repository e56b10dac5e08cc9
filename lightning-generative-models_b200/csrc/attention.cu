// Attention cores on NHWC qkv tensors [B, n, 384] (q | k | v; each 4 heads x 32, head-major).
//   LinearAttention   ddpm.py:222-238   (softmax over d for q, over n+4 for k; 32x32 context per head)
//   Attention/Attend  ddpm.py:255-271, models/modules/attend.py:111-126 (n <= 64, 4 memory kv)
// Everything inside a (sample, head) is fp32; tensors are in the activation dtype.
#include "common.cuh"

namespace b200dm {

constexpr int HEADS = 4, DH = 32, HID = HEADS * DH, NMEM = 4;
constexpr float kScale = 0.17677669529663687f;  // 32^-0.5

// ================================ LinearAttention ====================================================
// Kernel A: per (b,h): kmax[d], ksum[d], ctx[d][e] = sum_j softmax_j(k)[d,j] * v[e,j]  (j over mem + pixels)
constexpr int LA_CHUNK = 64;
template <typename T>
__global__ void __launch_bounds__(256)
linattn_ctx_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ mem_kv,
                   float* __restrict__ ctx, float* __restrict__ kstat, int n) {
  const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS;
  const int tid = threadIdx.x;
  const T* kbase = qkv + (int64_t)b * n * ld + HID + h * DH;
  const T* vbase = qkv + (int64_t)b * n * ld + 2 * HID + h * DH;
  const float* mk = mem_kv + (0 * HEADS + h) * DH * NMEM;  // [d][m]
  const float* mv = mem_kv + (1 * HEADS + h) * DH * NMEM;  // [e][m]
  __shared__ float red[8][DH];
  __shared__ float kmax[DH];
  __shared__ float P[LA_CHUNK][DH + 1];
  __shared__ float V[LA_CHUNK][DH];
  // ---- pass 1: max over j
  {
    const int d = tid & 31, jl = tid >> 5;
    float m = -INFINITY;
    for (int j = jl; j < n; j += 8) m = fmaxf(m, Elem<T>::ld(kbase + (int64_t)j * ld + d));
    if (jl < NMEM) m = fmaxf(m, mk[d * NMEM + jl]);
    red[jl][d] = m;
    __syncthreads();
    if (tid < DH) {
      float mm = red[0][tid];
      for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i][tid]);
      kmax[tid] = mm;
    }
    __syncthreads();
  }
  // ---- pass 2: accumulate
  const int d = tid >> 3, e0 = (tid & 7) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, psum = 0.f;
  const int total = n + NMEM;
  for (int c0 = 0; c0 < total; c0 += LA_CHUNK) {
    for (int i = tid; i < LA_CHUNK * DH; i += 256) {
      int jj = i >> 5, dd = i & 31;
      int j = c0 + jj;  // global index: [0,NMEM) memory, then pixels
      float pv = 0.f, vv = 0.f;
      if (j < total) {
        float kv_ = j < NMEM ? mk[dd * NMEM + j] : Elem<T>::ld(kbase + (int64_t)(j - NMEM) * ld + dd);
        vv = j < NMEM ? mv[dd * NMEM + j] : Elem<T>::ld(vbase + (int64_t)(j - NMEM) * ld + dd);
        pv = __expf(kv_ - kmax[dd]);
      }
      P[jj][dd] = pv;
      V[jj][dd] = vv;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < LA_CHUNK; ++jj) {
      float p = P[jj][d];
      float4 v4 = *reinterpret_cast<const float4*>(&V[jj][e0]);
      acc[0] = fmaf(p, v4.x, acc[0]);
      acc[1] = fmaf(p, v4.y, acc[1]);
      acc[2] = fmaf(p, v4.z, acc[2]);
      acc[3] = fmaf(p, v4.w, acc[3]);
      psum += p;
    }
    __syncthreads();
  }
  float inv = 1.f / psum;
  float* cp = ctx + (((int64_t)b * HEADS + h) * DH + d) * DH + e0;
  *reinterpret_cast<float4*>(cp) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  if ((tid & 7) == 0) {
    kstat[(((int64_t)b * HEADS + h) * DH + d) * 2] = kmax[d];
    kstat[(((int64_t)b * HEADS + h) * DH + d) * 2 + 1] = psum;
  }
}

// Kernel B: out[j, h*32+e] = sum_d ctx[d][e] * softmax_d(q[:,j])[d] * scale ; one warp per pixel
template <typename T>
__global__ void __launch_bounds__(256)
linattn_out_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ ctx,
                   T* __restrict__ out, int out_ld, int n, int pix_per_block) {
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float col[DH];  // ctx[d][e = lane]
  const float* cp = ctx + (int64_t)bh * DH * DH;
#pragma unroll
  for (int d = 0; d < DH; ++d) col[d] = cp[d * DH + lane];
  const int j0 = blockIdx.x * pix_per_block;
  const int j1 = min(j0 + pix_per_block, n);
  for (int j = j0 + warp; j < j1; j += 8) {
    const int64_t row = (int64_t)b * n + j;
    float q = Elem<T>::ld(qkv + row * ld + h * DH + lane);
    float m = warp_max(q);
    float p = __expf(q - m);
    float s = warp_sum(p);
    p = p / s * kScale;
    float o = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) o = fmaf(col[d], __shfl_sync(0xffffffffu, p, d), o);
    Elem<T>::st(out + row * out_ld + h * DH + lane, o);
  }
}

// Backward kernel C: per (b,h): dq for every pixel and dctx[d][e] = sum_j qs[d,j]*dout[e,j]
template <typename T>
__global__ void __launch_bounds__(256)
linattn_bwd_q_kernel(const T* __restrict__ dout, int dout_ld, const T* __restrict__ qkv, int ld,
                     const float* __restrict__ ctx, float* __restrict__ dctx, T* __restrict__ dqkv,
                     int dld, int n) {
  const int bh = blockIdx.x, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ float QS[LA_CHUNK][DH + 1];
  __shared__ float DO[LA_CHUNK][DH];
  float rowc[DH];  // ctx[d = lane][e]
  const float* cp = ctx + (int64_t)bh * DH * DH;
#pragma unroll
  for (int e = 0; e < DH; ++e) rowc[e] = cp[lane * DH + e];
  const int d = tid >> 3, e0 = (tid & 7) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < n; c0 += LA_CHUNK) {
    for (int jj = warp; jj < LA_CHUNK; jj += 8) {
      int j = c0 + jj;
      float qs = 0.f, g = 0.f;
      if (j < n) {
        const int64_t row = (int64_t)b * n + j;
        float q = Elem<T>::ld(qkv + row * ld + h * DH + lane);
        g = Elem<T>::ld(dout + row * dout_ld + h * DH + lane);  // dout[e = lane]
        float m = warp_max(q);
        float p = __expf(q - m);
        p = p / warp_sum(p);
        qs = p * kScale;
        float dqs = 0.f;  // sum_e ctx[lane][e] * dout[e]
#pragma unroll
        for (int e = 0; e < DH; ++e) dqs = fmaf(rowc[e], __shfl_sync(0xffffffffu, g, e), dqs);
        float t = warp_sum(p * dqs);
        Elem<T>::st(dqkv + row * dld + h * DH + lane, kScale * p * (dqs - t));
      }
      QS[jj][lane] = qs;
      DO[jj][lane] = g;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < LA_CHUNK; ++jj) {
      float p = QS[jj][d];
      float4 v4 = *reinterpret_cast<const float4*>(&DO[jj][e0]);
      acc[0] = fmaf(p, v4.x, acc[0]);
      acc[1] = fmaf(p, v4.y, acc[1]);
      acc[2] = fmaf(p, v4.z, acc[2]);
      acc[3] = fmaf(p, v4.w, acc[3]);
    }
    __syncthreads();
  }
  *reinterpret_cast<float4*>(dctx + ((int64_t)bh * DH + d) * DH + e0) =
      make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// Backward kernel D: dk, dv for every pixel (+ memory kv grads); one warp per pixel
template <typename T>
__global__ void __launch_bounds__(256)
linattn_bwd_kv_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ mem_kv,
                      const float* __restrict__ ctx, const float* __restrict__ kstat,
                      const float* __restrict__ dctx, T* __restrict__ dqkv, int dld,
                      float* __restrict__ dmem_kv, int n, int pix_per_block) {
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float drow[DH], dcol[DH];  // dctx[lane][e], dctx[d][lane]
  const float* dp = dctx + (int64_t)bh * DH * DH;
  const float* cp = ctx + (int64_t)bh * DH * DH;
  float Dd = 0.f;  // sum_e dctx[lane][e]*ctx[lane][e]
#pragma unroll
  for (int e = 0; e < DH; ++e) {
    drow[e] = dp[lane * DH + e];
    dcol[e] = dp[e * DH + lane];
    Dd = fmaf(drow[e], cp[lane * DH + e], Dd);
  }
  const float kmax = kstat[((int64_t)bh * DH + lane) * 2], kinv = 1.f / kstat[((int64_t)bh * DH + lane) * 2 + 1];
  const int total = n + NMEM;
  const int j0 = blockIdx.x * pix_per_block;
  const int j1 = min(j0 + pix_per_block, total);
  for (int j = j0 + warp; j < j1; j += 8) {
    float kv_, vv;
    if (j < NMEM) {
      kv_ = mem_kv[((0 * HEADS + h) * DH + lane) * NMEM + j];
      vv = mem_kv[((1 * HEADS + h) * DH + lane) * NMEM + j];
    } else {
      const int64_t row = (int64_t)b * n + (j - NMEM);
      kv_ = Elem<T>::ld(qkv + row * ld + HID + h * DH + lane);
      vv = Elem<T>::ld(qkv + row * ld + 2 * HID + h * DH + lane);
    }
    float ks = __expf(kv_ - kmax) * kinv;  // softmax_j(k)[d = lane, j]
    float dks = 0.f, dv = 0.f;
#pragma unroll
    for (int e = 0; e < DH; ++e) {
      dks = fmaf(drow[e], __shfl_sync(0xffffffffu, vv, e), dks);  // sum_e dctx[lane][e]*v[e]
      dv = fmaf(dcol[e], __shfl_sync(0xffffffffu, ks, e), dv);    // sum_d ks[d]*dctx[d][lane]
    }
    float dk = ks * (dks - Dd);
    if (j < NMEM) {
      atomicAdd(dmem_kv + ((0 * HEADS + h) * DH + lane) * NMEM + j, dk);
      atomicAdd(dmem_kv + ((1 * HEADS + h) * DH + lane) * NMEM + j, dv);
    } else {
      const int64_t row = (int64_t)b * n + (j - NMEM);
      Elem<T>::st(dqkv + row * dld + HID + h * DH + lane, dk);
      Elem<T>::st(dqkv + row * dld + 2 * HID + h * DH + lane, dv);
    }
  }
}

// ================================ full softmax attention (n <= 64) ==================================
constexpr int FA_MAXN = 64, FA_MAXKV = FA_MAXN + NMEM;
struct FaSmem {
  float Q[FA_MAXN][DH + 1];
  float K[FA_MAXKV][DH + 1];
  float V[FA_MAXKV][DH + 1];
  float S[FA_MAXN][FA_MAXKV + 1];
};

template <typename T>
__device__ __forceinline__ void fa_load(FaSmem& s, const T* qkv, int ld, const float* mem_kv, int b,
                                        int h, int n) {
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    int r = i >> 5, d = i & 31;
    const T* p = qkv + ((int64_t)b * n + r) * ld + h * DH + d;
    s.Q[r][d] = Elem<T>::ld(p);
    s.K[r + NMEM][d] = Elem<T>::ld(p + HID);
    s.V[r + NMEM][d] = Elem<T>::ld(p + 2 * HID);
  }
  for (int i = threadIdx.x; i < NMEM * DH; i += blockDim.x) {
    int r = i >> 5, d = i & 31;
    s.K[r][d] = mem_kv[((0 * HEADS + h) * NMEM + r) * DH + d];
    s.V[r][d] = mem_kv[((1 * HEADS + h) * NMEM + r) * DH + d];
  }
}

// S <- softmax(Q K^T * scale) rows
__device__ __forceinline__ void fa_probs(FaSmem& s, int n, int kv) {
  for (int i = threadIdx.x; i < n * kv; i += blockDim.x) {
    int r = i / kv, c = i - r * kv;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) a = fmaf(s.Q[r][d], s.K[c][d], a);
    s.S[r][c] = a * kScale;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < n; r += nw) {
    float m = -INFINITY;
    for (int c = lane; c < kv; c += 32) m = fmaxf(m, s.S[r][c]);
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < kv; c += 32) {
      float p = __expf(s.S[r][c] - m);
      s.S[r][c] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int c = lane; c < kv; c += 32) s.S[r][c] *= inv;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ mem_kv,
                T* __restrict__ out, int out_ld, int n) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FaSmem& s = *reinterpret_cast<FaSmem*>(smraw);
  const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS, kv = n + NMEM;
  fa_load(s, qkv, ld, mem_kv, b, h, n);
  __syncthreads();
  fa_probs(s, n, kv);
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    int r = i >> 5, d = i & 31;
    float a = 0.f;
    for (int c = 0; c < kv; ++c) a = fmaf(s.S[r][c], s.V[c][d], a);
    Elem<T>::st(out + ((int64_t)b * n + r) * out_ld + h * DH + d, a);
  }
}

struct FaBwdSmem {
  FaSmem f;
  float dO[FA_MAXN][DH + 1];
  float dS[FA_MAXN][FA_MAXKV + 1];
};

template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const T* __restrict__ dout, int dout_ld, const T* __restrict__ qkv, int ld,
                const float* __restrict__ mem_kv, T* __restrict__ dqkv, int dld,
                float* __restrict__ dmem_kv, int n) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FaBwdSmem& s = *reinterpret_cast<FaBwdSmem*>(smraw);
  const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS, kv = n + NMEM;
  fa_load(s.f, qkv, ld, mem_kv, b, h, n);
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    int r = i >> 5, d = i & 31;
    s.dO[r][d] = Elem<T>::ld(dout + ((int64_t)b * n + r) * dout_ld + h * DH + d);
  }
  __syncthreads();
  fa_probs(s.f, n, kv);
  // dP = dO V^T
  for (int i = threadIdx.x; i < n * kv; i += blockDim.x) {
    int r = i / kv, c = i - r * kv;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) a = fmaf(s.dO[r][d], s.f.V[c][d], a);
    s.dS[r][c] = a;
  }
  __syncthreads();
  // dS = P .* (dP - rowsum(P .* dP))
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = warp; r < n; r += nw) {
      float t = 0.f;
      for (int c = lane; c < kv; c += 32) t = fmaf(s.f.S[r][c], s.dS[r][c], t);
      t = warp_sum(t);
      for (int c = lane; c < kv; c += 32) s.dS[r][c] = s.f.S[r][c] * (s.dS[r][c] - t);
    }
  }
  __syncthreads();
  // dQ = dS K * scale
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    int r = i >> 5, d = i & 31;
    float a = 0.f;
    for (int c = 0; c < kv; ++c) a = fmaf(s.dS[r][c], s.f.K[c][d], a);
    Elem<T>::st(dqkv + ((int64_t)b * n + r) * dld + h * DH + d, a * kScale);
  }
  // dK = dS^T Q * scale ; dV = P^T dO
  for (int i = threadIdx.x; i < kv * DH; i += blockDim.x) {
    int c = i >> 5, d = i & 31;
    float ak = 0.f, av = 0.f;
    for (int r = 0; r < n; ++r) {
      ak = fmaf(s.dS[r][c], s.f.Q[r][d], ak);
      av = fmaf(s.f.S[r][c], s.dO[r][d], av);
    }
    ak *= kScale;
    if (c < NMEM) {
      atomicAdd(dmem_kv + ((0 * HEADS + h) * NMEM + c) * DH + d, ak);
      atomicAdd(dmem_kv + ((1 * HEADS + h) * NMEM + c) * DH + d, av);
    } else {
      T* p = dqkv + ((int64_t)b * n + (c - NMEM)) * dld + h * DH + d;
      Elem<T>::st(p + HID, ak);
      Elem<T>::st(p + 2 * HID, av);
    }
  }
}

}  // namespace b200dm

using namespace b200dm;
typedef __nv_bfloat16 bf16;

static inline int la_chunks(int B, int total, int* ppb) {
  int chunks = (4 * num_sms() + B * HEADS - 1) / (B * HEADS);
  int maxc = (total + 63) / 64;
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  *ppb = (total + chunks - 1) / chunks;
  return (total + *ppb - 1) / *ppb;
}

extern "C" int b200dm_linattn_fwd(int32_t dtype, const void* qkv, int32_t qkv_ld, const float* mem_kv,
                                  float* ctx, float* kstat, void* out, int32_t out_ld, int32_t B,
                                  int32_t n, void* stream) {
  B200DM_REQUIRE(B > 0 && n > 0, B200DM_ERR_SHAPE, "linattn_fwd: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  int ppb, chunks = la_chunks(B, n, &ppb);
  dim3 g2(chunks, B * HEADS);
  if (dtype == B200DM_F32) {
    linattn_ctx_kernel<float><<<B * HEADS, 256, 0, st>>>((const float*)qkv, qkv_ld, mem_kv, ctx, kstat, n);
    linattn_out_kernel<float><<<g2, 256, 0, st>>>((const float*)qkv, qkv_ld, ctx, (float*)out, out_ld, n, ppb);
  } else {
    linattn_ctx_kernel<bf16><<<B * HEADS, 256, 0, st>>>((const bf16*)qkv, qkv_ld, mem_kv, ctx, kstat, n);
    linattn_out_kernel<bf16><<<g2, 256, 0, st>>>((const bf16*)qkv, qkv_ld, ctx, (bf16*)out, out_ld, n, ppb);
  }
  count_launch(2);
  return check_launch("linattn_fwd");
}

extern "C" int b200dm_linattn_bwd(int32_t dtype, const void* dout, int32_t dout_ld, const void* qkv,
                                  int32_t qkv_ld, const float* mem_kv, const float* ctx,
                                  const float* kstat, float* dctx, void* dqkv, int32_t dqkv_ld,
                                  float* dmem_kv, int32_t B, int32_t n, void* stream) {
  B200DM_REQUIRE(B > 0 && n > 0, B200DM_ERR_SHAPE, "linattn_bwd: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  int ppb, chunks = la_chunks(B, n + NMEM, &ppb);
  dim3 g2(chunks, B * HEADS);
  if (dtype == B200DM_F32) {
    linattn_bwd_q_kernel<float><<<B * HEADS, 256, 0, st>>>((const float*)dout, dout_ld, (const float*)qkv, qkv_ld, ctx, dctx, (float*)dqkv, dqkv_ld, n);
    linattn_bwd_kv_kernel<float><<<g2, 256, 0, st>>>((const float*)qkv, qkv_ld, mem_kv, ctx, kstat, dctx, (float*)dqkv, dqkv_ld, dmem_kv, n, ppb);
  } else {
    linattn_bwd_q_kernel<bf16><<<B * HEADS, 256, 0, st>>>((const bf16*)dout, dout_ld, (const bf16*)qkv, qkv_ld, ctx, dctx, (bf16*)dqkv, dqkv_ld, n);
    linattn_bwd_kv_kernel<bf16><<<g2, 256, 0, st>>>((const bf16*)qkv, qkv_ld, mem_kv, ctx, kstat, dctx, (bf16*)dqkv, dqkv_ld, dmem_kv, n, ppb);
  }
  count_launch(2);
  return check_launch("linattn_bwd");
}

extern "C" int b200dm_attn_fwd(int32_t dtype, const void* qkv, int32_t qkv_ld, const float* mem_kv,
                               void* out, int32_t out_ld, int32_t B, int32_t n, void* stream) {
  B200DM_REQUIRE(B > 0 && n > 0 && n <= FA_MAXN, B200DM_ERR_UNSUPPORTED,
                 "attn_fwd: n=%d (softmax attention is built for n <= %d tokens)", n, FA_MAXN);
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = sizeof(FaSmem);
  if (dtype == B200DM_F32) {
    cudaFuncSetAttribute(attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_fwd_kernel<float><<<B * HEADS, 256, smem, st>>>((const float*)qkv, qkv_ld, mem_kv, (float*)out, out_ld, n);
  } else {
    cudaFuncSetAttribute(attn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_fwd_kernel<bf16><<<B * HEADS, 256, smem, st>>>((const bf16*)qkv, qkv_ld, mem_kv, (bf16*)out, out_ld, n);
  }
  count_launch();
  return check_launch("attn_fwd");
}

extern "C" int b200dm_attn_bwd(int32_t dtype, const void* dout, int32_t dout_ld, const void* qkv,
                               int32_t qkv_ld, const float* mem_kv, void* dqkv, int32_t dqkv_ld,
                               float* dmem_kv, int32_t B, int32_t n, void* stream) {
  B200DM_REQUIRE(B > 0 && n > 0 && n <= FA_MAXN, B200DM_ERR_UNSUPPORTED,
                 "attn_bwd: n=%d (softmax attention is built for n <= %d tokens)", n, FA_MAXN);
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = sizeof(FaBwdSmem);
  if (dtype == B200DM_F32) {
    cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_bwd_kernel<float><<<B * HEADS, 256, smem, st>>>((const float*)dout, dout_ld, (const float*)qkv, qkv_ld, mem_kv, (float*)dqkv, dqkv_ld, dmem_kv, n);
  } else {
    cudaFuncSetAttribute(attn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_bwd_kernel<bf16><<<B * HEADS, 256, smem, st>>>((const bf16*)dout, dout_ld, (const bf16*)qkv, qkv_ld, mem_kv, (bf16*)dqkv, dqkv_ld, dmem_kv, n);
  }
  count_launch();
  return check_launch("attn_bwd");
}
