// Attention cores on NHWC qkv tensors [B, n, 384] (q | k | v; each 4 heads x 32, head-major).
//   LinearAttention   ddpm.py:222-238   (softmax over d for q, over n+4 for k; 32x32 context per head)
//   Attention/Attend  ddpm.py:255-271, models/modules/attend.py:111-126 (n <= 64, 4 memory kv)
// Everything inside a (sample, head) is fp32; tensors are in the activation dtype.
#include "common.cuh"

namespace b200dm {

constexpr int HEADS = 4, DH = 32, HID = HEADS * DH, NMEM = 4;
constexpr float kScale = 0.17677669529663687f;  // 32^-0.5

// ================================ LinearAttention ====================================================
// All four kernels work on 64-pixel chunks staged in shared memory and compute the 32x32 products as
// register micro-tiles (4x4 or 2x4 per thread) instead of one-value-per-lane shuffles.
constexpr int LA_CHUNK = 64;
constexpr int LA_CPB = 4;   // chunks per CTA in the per-pixel kernels (amortises the 32x32 operand loads)

// vectorised tile load: 256 threads, thread = (pixel = tid/4, 8 channels = (tid%4)*8)
template <typename T>
__device__ __forceinline__ void la_load8(const T* base, int ld, int j, int n, int part, float (&v)[8]) {
  if (j < n) {
    ld8(base + (int64_t)j * ld + part * 8, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
}

// Kernel A: per (b,h): kmax[d], ksum[d], ctx[d][e] = sum_j softmax_j(k)[d,j] * v[e,j]  (j over mem + pixels)
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
  __shared__ float kmax[DH];
  __shared__ __align__(16) float P[LA_CHUNK][DH];
  __shared__ __align__(16) float V[LA_CHUNK][DH];
  __shared__ __align__(16) float red[4][DH][DH];   // also used for the max reduction
  const int pix = tid >> 2, part = tid & 3;
  // ---- pass 1: max over j
  {
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int j = pix; j < n; j += 64) {
      float v[8];
      ld8(kbase + (int64_t)j * ld + part * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
    }
    float* mred = &red[0][0][0];                   // [64 pixel lanes][32]
#pragma unroll
    for (int i = 0; i < 8; ++i) mred[pix * DH + part * 8 + i] = m[i];
    __syncthreads();
    if (tid < DH) {
      float mm = -INFINITY;
      for (int i = 0; i < 64; ++i) mm = fmaxf(mm, mred[i * DH + tid]);
      for (int i = 0; i < NMEM; ++i) mm = fmaxf(mm, mk[tid * NMEM + i]);
      kmax[tid] = mm;
    }
    __syncthreads();
  }
  // ---- pass 2: accumulate; thread = (pixel quarter, 4 d x 4 e micro-tile)
  const int grp = tid >> 6, tt = tid & 63;
  const int d0 = (tt >> 3) * 4, e0 = (tt & 7) * 4;
  float acc[4][4] = {}, psum[4] = {0.f, 0.f, 0.f, 0.f};
  const int total = n + NMEM;
  for (int c0 = 0; c0 < total; c0 += LA_CHUNK) {
    {
      const int j = c0 + pix;          // global index: [0,NMEM) memory, then pixels
      float kv[8], vv[8];
      if (j < NMEM) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          kv[i] = mk[(part * 8 + i) * NMEM + j];
          vv[i] = mv[(part * 8 + i) * NMEM + j];
        }
      } else {
        la_load8(kbase, ld, j - NMEM, n, part, kv);
        la_load8(vbase, ld, j - NMEM, n, part, vv);
      }
      const bool ok = j < total;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        P[pix][part * 8 + i] = ok ? __expf(kv[i] - kmax[part * 8 + i]) : 0.f;
        V[pix][part * 8 + i] = ok ? vv[i] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = grp * 16; jj < grp * 16 + 16; ++jj) {
      const float4 p4 = *reinterpret_cast<const float4*>(&P[jj][d0]);
      const float4 v4 = *reinterpret_cast<const float4*>(&V[jj][e0]);
      const float pa[4] = {p4.x, p4.y, p4.z, p4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        psum[i] += pa[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(pa[i], va[k], acc[i][k]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(&red[grp][d0 + i][e0]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __shared__ float ps[4][DH];
  if ((tt & 7) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) ps[grp][d0 + i] = psum[i];
  }
  __syncthreads();
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5, e = i & 31;
    const float sum = ps[0][d] + ps[1][d] + ps[2][d] + ps[3][d];
    const float a = red[0][d][e] + red[1][d][e] + red[2][d][e] + red[3][d][e];
    ctx[((int64_t)b * HEADS + h) * DH * DH + i] = a / sum;
    if (e == 0) {
      kstat[(((int64_t)b * HEADS + h) * DH + d) * 2] = kmax[d];
      kstat[(((int64_t)b * HEADS + h) * DH + d) * 2 + 1] = sum;
    }
  }
}

// softmax over the 32 channels of a pixel held by 4 consecutive lanes (8 each); returns probabilities
__device__ __forceinline__ void softmax32_quad(float (&v)[8]) {
  float m = v[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, v[i]);
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = __expf(v[i] - m);
    s += v[i];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  const float inv = 1.f / s;
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= inv;
}

// Kernel B: out[j, h*32+e] = sum_d ctx[d][e] * softmax_d(q[:,j])[d] * scale; CTA = 64 pixels of one (b,h)
template <typename T>
__global__ void __launch_bounds__(256)
linattn_out_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ ctx,
                   T* __restrict__ out, int out_ld, int n) {
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x;
  __shared__ __align__(16) float C_s[DH][DH];        // ctx[d][e]
  __shared__ __align__(16) float Q_s[LA_CHUNK][DH];  // scale * softmax(q)
  for (int i = tid; i < DH * DH; i += 256) C_s[i >> 5][i & 31] = ctx[(int64_t)bh * DH * DH + i];
  const int pix = tid >> 2, part = tid & 3;
  for (int cc = 0; cc < LA_CPB; ++cc) {
  const int j0 = (blockIdx.x * LA_CPB + cc) * LA_CHUNK;
  if (j0 >= n) break;
  __syncthreads();
  {
    float v[8];
    la_load8(qkv + (int64_t)b * n * ld + h * DH, ld, j0 + pix, n, part, v);
    softmax32_quad(v);
#pragma unroll
    for (int i = 0; i < 8; ++i) Q_s[pix][part * 8 + i] = v[i] * kScale;
  }
  __syncthreads();
  const int pp = tid >> 3, e0 = (tid & 7) * 4;
  float a0[4] = {}, a1[4] = {};
#pragma unroll 8
  for (int d = 0; d < DH; ++d) {
    const float4 c4 = *reinterpret_cast<const float4*>(&C_s[d][e0]);
    const float q0 = Q_s[2 * pp][d], q1 = Q_s[2 * pp + 1][d];
    a0[0] = fmaf(c4.x, q0, a0[0]); a0[1] = fmaf(c4.y, q0, a0[1]);
    a0[2] = fmaf(c4.z, q0, a0[2]); a0[3] = fmaf(c4.w, q0, a0[3]);
    a1[0] = fmaf(c4.x, q1, a1[0]); a1[1] = fmaf(c4.y, q1, a1[1]);
    a1[2] = fmaf(c4.z, q1, a1[2]); a1[3] = fmaf(c4.w, q1, a1[3]);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = j0 + 2 * pp + r;
    if (j < n) {
      T* o = out + ((int64_t)b * n + j) * out_ld + h * DH + e0;
      const float* a = r ? a1 : a0;
#pragma unroll
      for (int i = 0; i < 4; ++i) Elem<T>::st(o + i, a[i]);
    }
  }
  }  // chunk loop
}

// Backward kernel C: per (b,h): dq for every pixel and dctx[d][e] = sum_j qs[d,j]*dout[e,j]
template <typename T>
__global__ void __launch_bounds__(256)
linattn_bwd_q_kernel(const T* __restrict__ dout, int dout_ld, const T* __restrict__ qkv, int ld,
                     const float* __restrict__ ctx, float* __restrict__ dctx, T* __restrict__ dqkv,
                     int dld, int n) {
  const int bh = blockIdx.x, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x;
  __shared__ __align__(16) float CT_s[DH][DH];        // ctx transposed: CT_s[e][d]
  __shared__ __align__(16) float P_s[LA_CHUNK][DH];   // softmax(q) (without the scale)
  __shared__ __align__(16) float DO_s[LA_CHUNK][DH];  // dout[pixel][e]
  __shared__ __align__(16) float red[4][DH][DH];
  for (int i = tid; i < DH * DH; i += 256) CT_s[i & 31][i >> 5] = ctx[(int64_t)bh * DH * DH + i];
  const int pix = tid >> 2, part = tid & 3;
  const int grp = tid >> 6, tt = tid & 63;
  const int d0 = (tt >> 3) * 4, e0 = (tt & 7) * 4;
  const int pp = tid >> 3, dq0 = (tid & 7) * 4;       // dq mapping: pixel pair, 4 d's
  float acc[4][4] = {};
  for (int c0 = 0; c0 < n; c0 += LA_CHUNK) {
    __syncthreads();
    {
      float v[8], g[8];
      la_load8(qkv + (int64_t)b * n * ld + h * DH, ld, c0 + pix, n, part, v);
      la_load8(dout + (int64_t)b * n * dout_ld + h * DH, dout_ld, c0 + pix, n, part, g);
      softmax32_quad(v);
      const bool ok = c0 + pix < n;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        P_s[pix][part * 8 + i] = ok ? v[i] : 0.f;
        DO_s[pix][part * 8 + i] = g[i];
      }
    }
    __syncthreads();
    // dqs[pixel][d] = sum_e ctx[d][e] * dout[pixel][e]; two pixels x four d per thread
    float s0[4] = {}, s1[4] = {};
#pragma unroll 8
    for (int e = 0; e < DH; ++e) {
      const float4 c4 = *reinterpret_cast<const float4*>(&CT_s[e][dq0]);
      const float g0 = DO_s[2 * pp][e], g1 = DO_s[2 * pp + 1][e];
      s0[0] = fmaf(c4.x, g0, s0[0]); s0[1] = fmaf(c4.y, g0, s0[1]);
      s0[2] = fmaf(c4.z, g0, s0[2]); s0[3] = fmaf(c4.w, g0, s0[3]);
      s1[0] = fmaf(c4.x, g1, s1[0]); s1[1] = fmaf(c4.y, g1, s1[1]);
      s1[2] = fmaf(c4.z, g1, s1[2]); s1[3] = fmaf(c4.w, g1, s1[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float* sr = r ? s1 : s0;
      const float4 p4 = *reinterpret_cast<const float4*>(&P_s[2 * pp + r][dq0]);
      const float pa[4] = {p4.x, p4.y, p4.z, p4.w};
      float t = pa[0] * sr[0] + pa[1] * sr[1] + pa[2] * sr[2] + pa[3] * sr[3];
      t += __shfl_xor_sync(0xffffffffu, t, 1);   // the 8 lanes of a pixel pair cover the 32 d's
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      const int j = c0 + 2 * pp + r;
      if (j < n) {
        T* o = dqkv + ((int64_t)b * n + j) * dld + h * DH + dq0;
#pragma unroll
        for (int i = 0; i < 4; ++i) Elem<T>::st(o + i, kScale * pa[i] * (sr[i] - t));
      }
    }
    // dctx[d][e] += sum_pixels (scale*p[d]) * dout[e]
#pragma unroll 4
    for (int jj = grp * 16; jj < grp * 16 + 16; ++jj) {
      const float4 p4 = *reinterpret_cast<const float4*>(&P_s[jj][d0]);
      const float4 v4 = *reinterpret_cast<const float4*>(&DO_s[jj][e0]);
      const float pa[4] = {p4.x, p4.y, p4.z, p4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(pa[i], va[k], acc[i][k]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(&red[grp][d0 + i][e0]) =
        make_float4(acc[i][0] * kScale, acc[i][1] * kScale, acc[i][2] * kScale, acc[i][3] * kScale);
  __syncthreads();
  float* dws = dctx + (int64_t)bh * (DH + 1) * DH;     // [33][32]: dctx rows, then Dd
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5, e = i & 31;
    const float v = red[0][d][e] + red[1][d][e] + red[2][d][e] + red[3][d][e];
    dws[i] = v;
    red[0][d][e] = v * CT_s[e][d];                      // dctx[d][e]*ctx[d][e]
  }
  __syncthreads();
  if (tid < DH) {
    float a = 0.f;
    for (int e = 0; e < DH; ++e) a += red[0][tid][(e + tid) & 31];
    dws[DH * DH + tid] = a;
  }
}

// Backward kernel D: dk, dv for 64 pixels of one (b,h) (+ the memory kv gradients in chunk 0)
template <typename T>
__global__ void __launch_bounds__(256)
linattn_bwd_kv_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ mem_kv,
                      const float* __restrict__ ctx, const float* __restrict__ kstat,
                      const float* __restrict__ dctx, T* __restrict__ dqkv, int dld,
                      float* __restrict__ dmem_kv, int n) {
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x;
  __shared__ __align__(16) float D_s[DH][DH];         // dctx[d][e]
  __shared__ __align__(16) float DT_s[DH][DH];        // dctx^T: DT_s[e][d]
  __shared__ __align__(16) float KS_s[LA_CHUNK][DH];  // softmax_j(k)[d, pixel]
  __shared__ __align__(16) float V_s[LA_CHUNK][DH];
  __shared__ float Dd[DH], kmx[DH], kinv[DH];
  const float* dws = dctx + (int64_t)bh * (DH + 1) * DH;
  for (int i = tid; i < DH * DH; i += 256) {
    const float v = dws[i];
    D_s[i >> 5][i & 31] = v;
    DT_s[i & 31][i >> 5] = v;
  }
  if (tid < DH) {
    kmx[tid] = kstat[((int64_t)bh * DH + tid) * 2];
    kinv[tid] = 1.f / kstat[((int64_t)bh * DH + tid) * 2 + 1];
    Dd[tid] = dws[DH * DH + tid];
  }
  const int pix = tid >> 2, part = tid & 3;
  const int total = n + NMEM;
  for (int cc = 0; cc < LA_CPB; ++cc) {
  const int c0 = (blockIdx.x * LA_CPB + cc) * LA_CHUNK;  // global index: [0,NMEM) memory, then pixels
  if (c0 >= total) break;
  __syncthreads();
  {
    const int j = c0 + pix;
    float kv[8], vv[8];
    if (j < NMEM) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        kv[i] = mem_kv[((0 * HEADS + h) * DH + part * 8 + i) * NMEM + j];
        vv[i] = mem_kv[((1 * HEADS + h) * DH + part * 8 + i) * NMEM + j];
      }
    } else {
      la_load8(qkv + (int64_t)b * n * ld + HID + h * DH, ld, j - NMEM, n, part, kv);
      la_load8(qkv + (int64_t)b * n * ld + 2 * HID + h * DH, ld, j - NMEM, n, part, vv);
    }
    const bool ok = j < total;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int d = part * 8 + i;
      KS_s[pix][d] = ok ? __expf(kv[i] - kmx[d]) * kinv[d] : 0.f;
      V_s[pix][d] = ok ? vv[i] : 0.f;
    }
  }
  __syncthreads();
  const int pp = tid >> 3, c4 = (tid & 7) * 4;   // pixel pair, 4 channels (e for dv, d for dk)
  float dv0[4] = {}, dv1[4] = {}, dk0[4] = {}, dk1[4] = {};
#pragma unroll 8
  for (int k = 0; k < DH; ++k) {
    const float4 a4 = *reinterpret_cast<const float4*>(&D_s[k][c4]);    // dctx[d=k][e..]
    const float4 b4 = *reinterpret_cast<const float4*>(&DT_s[k][c4]);   // dctx[d..][e=k]
    const float ks0 = KS_s[2 * pp][k], ks1 = KS_s[2 * pp + 1][k];
    const float v0 = V_s[2 * pp][k], v1 = V_s[2 * pp + 1][k];
    dv0[0] = fmaf(a4.x, ks0, dv0[0]); dv0[1] = fmaf(a4.y, ks0, dv0[1]);
    dv0[2] = fmaf(a4.z, ks0, dv0[2]); dv0[3] = fmaf(a4.w, ks0, dv0[3]);
    dv1[0] = fmaf(a4.x, ks1, dv1[0]); dv1[1] = fmaf(a4.y, ks1, dv1[1]);
    dv1[2] = fmaf(a4.z, ks1, dv1[2]); dv1[3] = fmaf(a4.w, ks1, dv1[3]);
    dk0[0] = fmaf(b4.x, v0, dk0[0]); dk0[1] = fmaf(b4.y, v0, dk0[1]);
    dk0[2] = fmaf(b4.z, v0, dk0[2]); dk0[3] = fmaf(b4.w, v0, dk0[3]);
    dk1[0] = fmaf(b4.x, v1, dk1[0]); dk1[1] = fmaf(b4.y, v1, dk1[1]);
    dk1[2] = fmaf(b4.z, v1, dk1[2]); dk1[3] = fmaf(b4.w, v1, dk1[3]);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = c0 + 2 * pp + r;
    if (j >= total) continue;
    const float* dks = r ? dk1 : dk0;
    const float* dv = r ? dv1 : dv0;
    float dk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dk[i] = KS_s[2 * pp + r][c4 + i] * (dks[i] - Dd[c4 + i]);
    if (j < NMEM) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        atomicAdd(dmem_kv + ((0 * HEADS + h) * DH + c4 + i) * NMEM + j, dk[i]);
        atomicAdd(dmem_kv + ((1 * HEADS + h) * DH + c4 + i) * NMEM + j, dv[i]);
      }
    } else {
      T* o = dqkv + ((int64_t)b * n + (j - NMEM)) * dld + h * DH + c4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        Elem<T>::st(o + HID + i, dk[i]);
        Elem<T>::st(o + 2 * HID + i, dv[i]);
      }
    }
  }
  }  // chunk loop
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


extern "C" int b200dm_linattn_fwd(int32_t dtype, const void* qkv, int32_t qkv_ld, const float* mem_kv,
                                  float* ctx, float* kstat, void* out, int32_t out_ld, int32_t B,
                                  int32_t n, void* stream) {
  B200DM_REQUIRE(B > 0 && n > 0, B200DM_ERR_SHAPE, "linattn_fwd: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  B200DM_REQUIRE(qkv_ld % 8 == 0 && ((uintptr_t)qkv & 15) == 0, B200DM_ERR_SHAPE, "linattn_fwd: qkv must be 16-byte aligned, ld %% 8 == 0");
  dim3 g2((n + LA_CHUNK * LA_CPB - 1) / (LA_CHUNK * LA_CPB), B * HEADS);
  if (dtype == B200DM_F32) {
    linattn_ctx_kernel<float><<<B * HEADS, 256, 0, st>>>((const float*)qkv, qkv_ld, mem_kv, ctx, kstat, n);
    linattn_out_kernel<float><<<g2, 256, 0, st>>>((const float*)qkv, qkv_ld, ctx, (float*)out, out_ld, n);
  } else {
    linattn_ctx_kernel<bf16><<<B * HEADS, 256, 0, st>>>((const bf16*)qkv, qkv_ld, mem_kv, ctx, kstat, n);
    linattn_out_kernel<bf16><<<g2, 256, 0, st>>>((const bf16*)qkv, qkv_ld, ctx, (bf16*)out, out_ld, n);
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
  B200DM_REQUIRE(qkv_ld % 8 == 0 && dout_ld % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)dout & 15) == 0,
                 B200DM_ERR_SHAPE, "linattn_bwd: tensors must be 16-byte aligned, ld %% 8 == 0");
  dim3 g2((n + NMEM + LA_CHUNK * LA_CPB - 1) / (LA_CHUNK * LA_CPB), B * HEADS);
  if (dtype == B200DM_F32) {
    linattn_bwd_q_kernel<float><<<B * HEADS, 256, 0, st>>>((const float*)dout, dout_ld, (const float*)qkv, qkv_ld, ctx, dctx, (float*)dqkv, dqkv_ld, n);
    linattn_bwd_kv_kernel<float><<<g2, 256, 0, st>>>((const float*)qkv, qkv_ld, mem_kv, ctx, kstat, dctx, (float*)dqkv, dqkv_ld, dmem_kv, n);
  } else {
    linattn_bwd_q_kernel<bf16><<<B * HEADS, 256, 0, st>>>((const bf16*)dout, dout_ld, (const bf16*)qkv, qkv_ld, ctx, dctx, (bf16*)dqkv, dqkv_ld, n);
    linattn_bwd_kv_kernel<bf16><<<g2, 256, 0, st>>>((const bf16*)qkv, qkv_ld, mem_kv, ctx, kstat, dctx, (bf16*)dqkv, dqkv_ld, dmem_kv, n);
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
