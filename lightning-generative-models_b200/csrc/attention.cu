// Attention cores on NHWC qkv tensors [B, n, 384] (q | k | v; each 4 heads x 32, head-major).
//   LinearAttention   ddpm.py:222-238   (softmax over d for q, over n+4 for k; 32x32 context per head)
//   Attention/Attend  ddpm.py:255-271, models/modules/attend.py:111-126 (n <= 64, 4 memory kv)
// Everything inside a (sample, head) is fp32; tensors are in the activation dtype.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200dm {

constexpr int HEADS = 4, DH = 32, HID = HEADS * DH, NMEM = 4;
constexpr float kScale = 0.17677669529663687f;  // 32^-0.5

// ================================ LinearAttention ====================================================
// All four kernels work on 64-pixel chunks staged in shared memory and compute the 32x32 products as
// register micro-tiles (4x4 or 2x4 per thread) instead of one-value-per-lane shuffles.
constexpr int LA_CHUNK = 64;
constexpr float kLog2e = 1.4426950408889634f;
// single-instruction exp2 / reciprocal (MUFU) without the denormal-range fix-up code of expf / division
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
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
  pdl_prologue();
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

// the same with exp as one FFMA + ex2.approx and a single approximate reciprocal; probabilities times `scale`
__device__ __forceinline__ void softmax32_quad_fast(float (&v)[8], float scale) {
  float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  const float nm = -m * kLog2e;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = ex2_ftz(fmaf(v[i], kLog2e, nm));
    s += v[i];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  const float inv = rcp_ftz(s) * scale;                  // s >= 1: the maximum contributes exp(0)
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= inv;
}

// Kernel B: out[j, h*32+e] = sum_d ctx[d][e] * softmax_d(q[:,j])[d] * scale; CTA = 64 pixels of one (b,h)
template <typename T>
__global__ void __launch_bounds__(256)
linattn_out_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ ctx,
                   T* __restrict__ out, int out_ld, int n) {
  pdl_prologue();
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
  pdl_prologue();
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
  pdl_prologue();
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

// =====================================================================================================
// Cluster-fused LinearAttention: ONE launch per direction.  A cluster of CL CTAs owns one (sample, head);
// each CTA handles a contiguous pixel range.  The 32x32 context (forward) / its gradient (backward) is
// reduced across the cluster through distributed shared memory, so q/k/v (and dout) are streamed from HBM
// exactly once and nothing but the final context/statistics (saved for backward) goes through global
// memory.  Global loads of the next 64-pixel chunk are issued before the current chunk is processed
// (register double buffering) — the previous kernels exposed one DRAM round trip per chunk.
// =====================================================================================================
template <typename T> struct LaRaw;
template <> struct LaRaw<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void zero() { a = make_float4(0, 0, 0, 0); b = a; }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct LaRaw<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

struct LaFwdSmem {
  float P[LA_CHUNK][DH];       // exp(k - m_loc) of the current chunk / scale*softmax(q) in the output phase
  float V[LA_CHUNK][DH];
  float red[4][DH][DH];        // per pixel-quarter partial contexts; also the max reduction scratch
  float ctx_loc[DH][DH];       // this CTA's un-normalised partial context   (read by the cluster peers)
  float m_loc[DH], l_loc[DH];  // its running max and exp-sum per key channel (read by the cluster peers)
  float ctx[DH][DH];           // merged, normalised context
  float ps[4][DH];
};

template <typename T>
__global__ void __launch_bounds__(256)
linattn_fwd_cluster_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ mem_kv,
                           float* __restrict__ ctx_out, float* __restrict__ kstat, T* __restrict__ out,
                           int out_ld, int n) {
  pdl_prologue();
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char la_smem_raw[];
  LaFwdSmem& s = *reinterpret_cast<LaFwdSmem*>(la_smem_raw);
  const T* qbase = qkv + (int64_t)b * n * ld + h * DH;
  const T* kbase = qbase + HID;
  const T* vbase = qbase + 2 * HID;
  const float* mk = mem_kv + (0 * HEADS + h) * DH * NMEM;  // [d][m]
  const float* mv = mem_kv + (1 * HEADS + h) * DH * NMEM;  // [e][m]
  // pixel range of this CTA (whole 64-pixel chunks); rank 0 also owns the NMEM memory keys
  const int chunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  const int cpr = (chunks + CL - 1) / CL;
  const int p0 = min(rank * cpr * LA_CHUNK, n), p1 = min((rank + 1) * cpr * LA_CHUNK, n);
  const int pix = tid >> 2, part = tid & 3;
  // ---- phase 1a: local max of k over the range (4 independent loads in flight per thread)
  {
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int j = p0 + pix; j < p1; j += 4 * LA_CHUNK) {
      LaRaw<T> raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j + u * LA_CHUNK < p1) raw[u].load(kbase + (int64_t)(j + u * LA_CHUNK) * ld + part * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u * LA_CHUNK < p1) {
          float v[8];
          raw[u].unpack(v);
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
        }
      }
    }
    float* mred = &s.red[0][0][0];                   // [64 pixel lanes][32]
#pragma unroll
    for (int i = 0; i < 8; ++i) mred[pix * DH + part * 8 + i] = m[i];
    __syncthreads();
    if (tid < DH) {
      float mm = -INFINITY;
      for (int i = 0; i < 64; ++i) mm = fmaxf(mm, mred[i * DH + tid]);
      if (rank == 0)
        for (int i = 0; i < NMEM; ++i) mm = fmaxf(mm, mk[tid * NMEM + i]);
      s.m_loc[tid] = mm;
    }
    __syncthreads();
  }
  // ---- phase 1b: partial context; thread = (pixel quarter, 4 d x 4 e micro-tile)
  const int grp = tid >> 6, tt = tid & 63;
  const int d0 = (tt >> 3) * 4, e0 = (tt & 7) * 4;
  {
    float acc[4][4] = {}, psum[4] = {0.f, 0.f, 0.f, 0.f};
    // local index space: rank 0 -> [0,NMEM) memory then its pixels; other ranks -> pixels only
    const int lead = rank == 0 ? NMEM : 0;
    const int total = (p1 - p0) + lead;
    LaRaw<T> rk, rv;
    auto fetch = [&](int c0) {
      const int j = c0 + pix - lead;                  // pixel offset inside the range (negative: memory key)
      if (j >= 0 && j < p1 - p0) {
        rk.load(kbase + (int64_t)(p0 + j) * ld + part * 8);
        rv.load(vbase + (int64_t)(p0 + j) * ld + part * 8);
      } else {
        rk.zero();
        rv.zero();
      }
    };
    if (total > 0) fetch(0);
    for (int c0 = 0; c0 < total; c0 += LA_CHUNK) {
      float kv[8], vv[8];
      rk.unpack(kv);
      rv.unpack(vv);
      const int jl = c0 + pix;                        // local index of this thread's row
      if (jl < lead) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          kv[i] = mk[(part * 8 + i) * NMEM + jl];
          vv[i] = mv[(part * 8 + i) * NMEM + jl];
        }
      }
      if (c0 + LA_CHUNK < total) fetch(c0 + LA_CHUNK);   // next chunk's loads fly during this chunk's math
      const bool ok = jl < total;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s.P[pix][part * 8 + i] = ok ? __expf(kv[i] - s.m_loc[part * 8 + i]) : 0.f;
        s.V[pix][part * 8 + i] = ok ? vv[i] : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int jj = grp * 16; jj < grp * 16 + 16; ++jj) {
        const float4 p4 = *reinterpret_cast<const float4*>(&s.P[jj][d0]);
        const float4 v4 = *reinterpret_cast<const float4*>(&s.V[jj][e0]);
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
      *reinterpret_cast<float4*>(&s.red[grp][d0 + i][e0]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if ((tt & 7) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s.ps[grp][d0 + i] = psum[i];
    }
    __syncthreads();
    for (int i = tid; i < DH * DH; i += 256) {
      const int d = i >> 5, e = i & 31;
      s.ctx_loc[d][e] = s.red[0][d][e] + s.red[1][d][e] + s.red[2][d][e] + s.red[3][d][e];
      if (e == 0) s.l_loc[d] = s.ps[0][d] + s.ps[1][d] + s.ps[2][d] + s.ps[3][d];
    }
  }
  cluster.sync();
  // ---- merge the CL partial softmax states (every CTA computes the full context it needs)
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5, e = i & 31;
    float M = -INFINITY;
    for (int r = 0; r < CL; ++r) M = fmaxf(M, cluster.map_shared_rank(&s.m_loc[0], r)[d]);
    float L = 0.f, a = 0.f;
    for (int r = 0; r < CL; ++r) {
      const float w = __expf(cluster.map_shared_rank(&s.m_loc[0], r)[d] - M);
      L = fmaf(cluster.map_shared_rank(&s.l_loc[0], r)[d], w, L);
      a = fmaf(cluster.map_shared_rank(&s.ctx_loc[0][0], r)[i], w, a);
    }
    const float c = a / L;
    s.ctx[d][e] = c;
    if (rank == 0) {
      ctx_out[(int64_t)bh * DH * DH + i] = c;
      if (e == 0) {
        kstat[((int64_t)bh * DH + d) * 2] = M;
        kstat[((int64_t)bh * DH + d) * 2 + 1] = L;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: out[j, e] = sum_d ctx[d][e] * scale * softmax_d(q[j, :])[d]  for the own pixels
  {
    LaRaw<T> rq;
    auto fetchq = [&](int j0) {
      if (j0 + pix < p1) rq.load(qbase + (int64_t)(j0 + pix) * ld + part * 8);
      else rq.zero();
    };
    if (p0 < p1) fetchq(p0);
    const int pp = tid >> 3, eo = (tid & 7) * 4;
    for (int j0 = p0; j0 < p1; j0 += LA_CHUNK) {
      float v[8];
      rq.unpack(v);
      if (j0 + LA_CHUNK < p1) fetchq(j0 + LA_CHUNK);
      softmax32_quad(v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s.P[pix][part * 8 + i] = v[i] * kScale;
      __syncthreads();
      float a0[4] = {}, a1[4] = {};
#pragma unroll 8
      for (int d = 0; d < DH; ++d) {
        const float4 c4 = *reinterpret_cast<const float4*>(&s.ctx[d][eo]);
        const float q0 = s.P[2 * pp][d], q1 = s.P[2 * pp + 1][d];
        a0[0] = fmaf(c4.x, q0, a0[0]); a0[1] = fmaf(c4.y, q0, a0[1]);
        a0[2] = fmaf(c4.z, q0, a0[2]); a0[3] = fmaf(c4.w, q0, a0[3]);
        a1[0] = fmaf(c4.x, q1, a1[0]); a1[1] = fmaf(c4.y, q1, a1[1]);
        a1[2] = fmaf(c4.z, q1, a1[2]); a1[3] = fmaf(c4.w, q1, a1[3]);
      }
      if (j0 + 2 * pp < p1) st4(out + ((int64_t)b * n + j0 + 2 * pp) * out_ld + h * DH + eo, a0);
      if (j0 + 2 * pp + 1 < p1) st4(out + ((int64_t)b * n + j0 + 2 * pp + 1) * out_ld + h * DH + eo, a1);
      __syncthreads();
    }
  }
  cluster.sync();     // peers may still be reading this CTA's partials
}

struct LaBwdSmem {
  float CT[DH][DH];            // ctx transposed: CT[e][d]
  float A[LA_CHUNK][DH];       // softmax(q) in phase 1, softmax_n(k) in phase 2
  float Bm[LA_CHUNK][DH];      // dout in phase 1, v in phase 2
  float red[4][DH][DH];
  float dctx_loc[DH][DH];      // this CTA's partial d(context)   (read by the cluster peers)
  float D[DH][DH];             // merged dctx[d][e]
  float DT[DH][DH];            // merged dctx^T
  float Dd[DH], kmx[DH], kinv[DH];
};

template <typename T>
__global__ void __launch_bounds__(256)
linattn_bwd_cluster_kernel(const T* __restrict__ dout, int dout_ld, const T* __restrict__ qkv, int ld,
                           const float* __restrict__ mem_kv, const float* __restrict__ ctx,
                           const float* __restrict__ kstat, T* __restrict__ dqkv, int dld,
                           float* __restrict__ dmem_kv, int n) {
  pdl_prologue();
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char la_smem_raw[];
  LaBwdSmem& s = *reinterpret_cast<LaBwdSmem*>(la_smem_raw);
  const T* qbase = qkv + (int64_t)b * n * ld + h * DH;
  const T* kbase = qbase + HID;
  const T* vbase = qbase + 2 * HID;
  const T* gbase = dout + (int64_t)b * n * dout_ld + h * DH;
  T* dqbase = dqkv + (int64_t)b * n * dld + h * DH;
  const int chunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  const int cpr = (chunks + CL - 1) / CL;
  const int p0 = min(rank * cpr * LA_CHUNK, n), p1 = min((rank + 1) * cpr * LA_CHUNK, n);
  for (int i = tid; i < DH * DH; i += 256) s.CT[i & 31][i >> 5] = ctx[(int64_t)bh * DH * DH + i];
  if (tid < DH) {
    s.kmx[tid] = kstat[((int64_t)bh * DH + tid) * 2];
    s.kinv[tid] = 1.f / kstat[((int64_t)bh * DH + tid) * 2 + 1];
  }
  const int pix = tid >> 2, part = tid & 3;
  const int grp = tid >> 6, tt = tid & 63;
  const int d0 = (tt >> 3) * 4, e0 = (tt & 7) * 4;
  const int pp = tid >> 3, c4 = (tid & 7) * 4;          // pixel pair, 4 channels
  // ---- phase 1: dq for the own pixels, partial dctx[d][e] = sum_j scale*softmax(q)[j,d]*dout[j,e]
  {
    float acc[4][4] = {};
    LaRaw<T> rq, rg;
    auto fetch = [&](int j0) {
      if (j0 + pix < p1) {
        rq.load(qbase + (int64_t)(j0 + pix) * ld + part * 8);
        rg.load(gbase + (int64_t)(j0 + pix) * dout_ld + part * 8);
      } else {
        rq.zero();
        rg.zero();
      }
    };
    if (p0 < p1) fetch(p0);
    for (int j0 = p0; j0 < p1; j0 += LA_CHUNK) {
      float v[8], g[8];
      rq.unpack(v);
      rg.unpack(g);
      if (j0 + LA_CHUNK < p1) fetch(j0 + LA_CHUNK);
      softmax32_quad(v);
      const bool ok = j0 + pix < p1;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s.A[pix][part * 8 + i] = ok ? v[i] : 0.f;
        s.Bm[pix][part * 8 + i] = g[i];
      }
      __syncthreads();
      // dqs[pixel][d] = sum_e ctx[d][e] * dout[pixel][e]; two pixels x four d per thread
      float s0[4] = {}, s1[4] = {};
#pragma unroll 8
      for (int e = 0; e < DH; ++e) {
        const float4 cc = *reinterpret_cast<const float4*>(&s.CT[e][c4]);
        const float g0 = s.Bm[2 * pp][e], g1 = s.Bm[2 * pp + 1][e];
        s0[0] = fmaf(cc.x, g0, s0[0]); s0[1] = fmaf(cc.y, g0, s0[1]);
        s0[2] = fmaf(cc.z, g0, s0[2]); s0[3] = fmaf(cc.w, g0, s0[3]);
        s1[0] = fmaf(cc.x, g1, s1[0]); s1[1] = fmaf(cc.y, g1, s1[1]);
        s1[2] = fmaf(cc.z, g1, s1[2]); s1[3] = fmaf(cc.w, g1, s1[3]);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float* sr = r ? s1 : s0;
        const float4 p4 = *reinterpret_cast<const float4*>(&s.A[2 * pp + r][c4]);
        const float pa[4] = {p4.x, p4.y, p4.z, p4.w};
        float t = pa[0] * sr[0] + pa[1] * sr[1] + pa[2] * sr[2] + pa[3] * sr[3];
        t += __shfl_xor_sync(0xffffffffu, t, 1);   // the 8 lanes of a pixel pair cover the 32 d's
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        const int j = j0 + 2 * pp + r;
        if (j < p1) {
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = kScale * pa[i] * (sr[i] - t);
          st4(dqbase + (int64_t)j * dld + c4, o);
        }
      }
#pragma unroll 4
      for (int jj = grp * 16; jj < grp * 16 + 16; ++jj) {
        const float4 p4 = *reinterpret_cast<const float4*>(&s.A[jj][d0]);
        const float4 v4 = *reinterpret_cast<const float4*>(&s.Bm[jj][e0]);
        const float pa[4] = {p4.x, p4.y, p4.z, p4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(pa[i], va[k], acc[i][k]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&s.red[grp][d0 + i][e0]) =
          make_float4(acc[i][0] * kScale, acc[i][1] * kScale, acc[i][2] * kScale, acc[i][3] * kScale);
    __syncthreads();
    for (int i = tid; i < DH * DH; i += 256) {
      const int d = i >> 5, e = i & 31;
      s.dctx_loc[d][e] = s.red[0][d][e] + s.red[1][d][e] + s.red[2][d][e] + s.red[3][d][e];
    }
  }
  cluster.sync();
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5, e = i & 31;
    float a = 0.f;
    for (int r = 0; r < CL; ++r) a += cluster.map_shared_rank(&s.dctx_loc[0][0], r)[i];
    s.D[d][e] = a;
    s.DT[e][d] = a;
    s.red[0][d][e] = a * s.CT[e][d];                    // dctx[d][e]*ctx[d][e]
  }
  __syncthreads();
  if (tid < DH) {
    float a = 0.f;
    for (int e = 0; e < DH; ++e) a += s.red[0][tid][(e + tid) & 31];
    s.Dd[tid] = a;
  }
  // ---- phase 2: dk, dv for the own pixels (+ the memory keys on rank 0)
  {
    const int lead = rank == 0 ? NMEM : 0;
    const int total = (p1 - p0) + lead;
    LaRaw<T> rk, rv;
    auto fetch = [&](int c0) {
      const int j = c0 + pix - lead;
      if (j >= 0 && j < p1 - p0) {
        rk.load(kbase + (int64_t)(p0 + j) * ld + part * 8);
        rv.load(vbase + (int64_t)(p0 + j) * ld + part * 8);
      } else {
        rk.zero();
        rv.zero();
      }
    };
    if (total > 0) fetch(0);
    for (int c0 = 0; c0 < total; c0 += LA_CHUNK) {
      float kv[8], vv[8];
      rk.unpack(kv);
      rv.unpack(vv);
      const int jl = c0 + pix;
      if (jl < lead) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          kv[i] = mem_kv[((0 * HEADS + h) * DH + part * 8 + i) * NMEM + jl];
          vv[i] = mem_kv[((1 * HEADS + h) * DH + part * 8 + i) * NMEM + jl];
        }
      }
      if (c0 + LA_CHUNK < total) fetch(c0 + LA_CHUNK);
      const bool ok = jl < total;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int d = part * 8 + i;
        s.A[pix][d] = ok ? __expf(kv[i] - s.kmx[d]) * s.kinv[d] : 0.f;
        s.Bm[pix][d] = ok ? vv[i] : 0.f;
      }
      __syncthreads();
      float dv0[4] = {}, dv1[4] = {}, dk0[4] = {}, dk1[4] = {};
#pragma unroll 8
      for (int k = 0; k < DH; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&s.D[k][c4]);    // dctx[d=k][e..]
        const float4 b4 = *reinterpret_cast<const float4*>(&s.DT[k][c4]);   // dctx[d..][e=k]
        const float ks0 = s.A[2 * pp][k], ks1 = s.A[2 * pp + 1][k];
        const float v0 = s.Bm[2 * pp][k], v1 = s.Bm[2 * pp + 1][k];
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
        const int jr = c0 + 2 * pp + r;                 // local index
        if (jr >= total) continue;
        const float* dks = r ? dk1 : dk0;
        const float* dv = r ? dv1 : dv0;
        float dk[4], dvv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dk[i] = s.A[2 * pp + r][c4 + i] * (dks[i] - s.Dd[c4 + i]);
          dvv[i] = dv[i];
        }
        if (jr < lead) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            atomicAdd(dmem_kv + ((0 * HEADS + h) * DH + c4 + i) * NMEM + jr, dk[i]);
            atomicAdd(dmem_kv + ((1 * HEADS + h) * DH + c4 + i) * NMEM + jr, dvv[i]);
          }
        } else {
          T* o = dqbase + (int64_t)(p0 + jr - lead) * dld + c4;
          st4(o + HID, dk);
          st4(o + 2 * HID, dvv);
        }
      }
    }
  }
  cluster.sync();
}

// =====================================================================================================
// bf16 path: the same cluster algorithm with the four 32-wide contractions on tensor cores
// (mma.sync.m16n8k16, bf16 operands staged in shared memory, fp32 accumulation).  The fp32 SIMT version
// above issues ~400 instructions per thread per 64-pixel chunk and is instruction-bound; this version is
// HBM-bound.  (tcgen05 needs M >= 64 per instruction; the per-head problems are 32x32xn, so the warp-level
// MMA is the fitting tensor instruction here — <1% of the network's FLOPs.)
// Operand rounding to bf16 matches the reference's autocast einsum (ddpm.py:232-237 under bf16-mixed).
// =====================================================================================================
constexpr int LP = 40;   // row pitch (bf16 elements) of the staged operand tiles: 80 B -> conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
typedef __nv_bfloat16 lbf;

// acc[32 x 32 tile of this warp: rows mt*16.., cols nt*8..] += A^T B over the 64 staged rows.
//   A, B: [64][LP] bf16 (row = pixel).  warp w: mt = w >> 2, nt = w & 3.
__device__ __forceinline__ void mma_atb(const lbf* A, const lbf* B, float (&acc)[4], int warp, int lane) {
  const int mt = warp >> 2, nt = warp & 3;
  const int mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4], b[2];
    ldsm_x4_t(a, (uint32_t)__cvta_generic_to_shared(A + (ks * 16 + (mi >> 1) * 8 + r) * LP + mt * 16 + (mi & 1) * 8));
    ldsm_x2_t(b, (uint32_t)__cvta_generic_to_shared(B + (ks * 16 + (mi & 1) * 8 + r) * LP + nt * 8));
    mma_bf16_16816(acc, a, b);
  }
}
// acc[j][n] = sum_k X[j][k] * W(k, n) for 64 rows j, 32 k, 32 n.  warp w: rows (w >> 1)*16.., n-tiles
// (w & 1)*2 + {0,1}.  W_KN = true: W stored [k][LP] (row = k); false: W stored [n][LP] (row = n).
template <bool W_KN>
__device__ __forceinline__ void mma_xw(const lbf* X, const lbf* W, float (&acc)[2][4], int warp, int lane) {
  const int mt = warp >> 1, nt0 = (warp & 1) * 2;
  const int mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, (uint32_t)__cvta_generic_to_shared(X + (mt * 16 + (mi & 1) * 8 + r) * LP + ks * 16 + (mi >> 1) * 8));
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      uint32_t b[2];
      const int n0 = (nt0 + t) * 8;
      if (W_KN) ldsm_x2_t(b, (uint32_t)__cvta_generic_to_shared(W + (ks * 16 + (mi & 1) * 8 + r) * LP + n0));
      else ldsm_x2(b, (uint32_t)__cvta_generic_to_shared(W + (n0 + r) * LP + ks * 16 + (mi & 1) * 8));
      mma_bf16_16816(acc[t], a, b);
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// column reduction over the 64 pixel lanes: thread (pix = tid >> 2, part = tid & 3) holds 8 channels;
// result[32] = op over all pixels.  scratch: [8 warps][32].
template <bool MAX>
__device__ __forceinline__ void la_col_reduce(float (&v)[8], float (*scratch)[DH], float* result, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, v[i], o);
      v[i] = MAX ? fmaxf(v[i], t) : v[i] + t;
    }
  }
  if ((tid & 31) < 4) {
#pragma unroll
    for (int i = 0; i < 8; ++i) scratch[tid >> 5][(tid & 3) * 8 + i] = v[i];
  }
  __syncthreads();
  if (tid < DH) {
    float r = scratch[0][tid];
#pragma unroll
    for (int w = 1; w < 8; ++w) r = MAX ? fmaxf(r, scratch[w][tid]) : r + scratch[w][tid];
    result[tid] = r;
  }
  __syncthreads();
}

struct LaFwdTc {
  lbf Kb[2][LA_CHUNK][LP];     // double-buffered: exp(k - m_loc) / scale*softmax(q) in the output phase
  lbf Vb[2][LA_CHUNK][LP];     // double-buffered: v / output staging in the output phase
  lbf Cb[DH][LP];              // merged context, bf16, [d][e]
  float mred[8][DH];           // reduction scratch (max, exp sums)
  float ctx_loc[DH][DH];       // partial context (read by the cluster peers)
  float m_loc[DH], l_loc[DH];
};

__global__ void __launch_bounds__(256)
linattn_fwd_tc_kernel(const lbf* __restrict__ qkv, int ld, const float* __restrict__ mem_kv,
                      float* __restrict__ ctx_out, float* __restrict__ kstat, lbf* __restrict__ out,
                      int out_ld, int n) {
  pdl_prologue();
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ __align__(16) LaFwdTc s;
  const lbf* qbase = qkv + (int64_t)b * n * ld + h * DH;
  const lbf* kbase = qbase + HID;
  const lbf* vbase = qbase + 2 * HID;
  const float* mk = mem_kv + (0 * HEADS + h) * DH * NMEM;  // [d][m]
  const float* mv = mem_kv + (1 * HEADS + h) * DH * NMEM;  // [e][m]
  const int chunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  const int cpr = (chunks + CL - 1) / CL;
  const int p0 = min(rank * cpr * LA_CHUNK, n), p1 = min((rank + 1) * cpr * LA_CHUNK, n);
  const int pix = tid >> 2, part = tid & 3;
  // ---- phase 1a: local max of k
  {
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int j = p0 + pix; j < p1; j += 4 * LA_CHUNK) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j + u * LA_CHUNK < p1) raw[u] = *reinterpret_cast<const uint4*>(kbase + (int64_t)(j + u * LA_CHUNK) * ld + part * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u * LA_CHUNK < p1) {
          float v[8];
          unpack8(raw[u], v);
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
        }
      }
    }
    if (rank == 0 && pix < NMEM) {
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], mk[(part * 8 + i) * NMEM + pix]);
    }
    la_col_reduce<true>(m, s.mred, s.m_loc, tid);
  }
  // ---- phase 1b: partial context on tensor cores.  The loop is issue-bound (ncu: 55 % issue slots at 43 %
  //      occupancy), so it is written for instruction count: exp as one FFMA + ex2.approx (the max is folded into
  //      the addend), walking pointers, pixels past the end loaded as k = -inf / v = 0 instead of being selected
  //      away, and the memory key/values of rank 0 as a chunk of their own in front of the loop.
  {
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, psum[8] = {};
    float nm2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) nm2[i] = -s.m_loc[part * 8 + i] * kLog2e;
    lbf* const kst = &s.Kb[0][pix][part * 8];
    lbf* const vst = &s.Vb[0][pix][part * 8];
    constexpr int BUF = LA_CHUNK * LP;                 // elements between the two buffers of a tile
    int cb = 0;
    if (rank == 0) {
      float pv[8], vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = pix < NMEM;
        pv[i] = ok ? ex2_ftz(fmaf(mk[(part * 8 + i) * NMEM + (ok ? pix : 0)], kLog2e, nm2[i])) : 0.f;
        vv[i] = ok ? mv[(part * 8 + i) * NMEM + pix] : 0.f;
        psum[i] += pv[i];
      }
      *reinterpret_cast<uint4*>(kst) = pack8(pv);
      *reinterpret_cast<uint4*>(vst) = pack8(vv);
      __syncthreads();
      mma_atb(&s.Kb[0][0][0], &s.Vb[0][0][0], acc, warp, lane);
      cb = 1;
    }
    const uint4 kNegInf = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf x 8
    const lbf* kp = kbase + (int64_t)(p0 + pix) * ld + part * 8;
    const int64_t step = (int64_t)LA_CHUNK * ld;
    uint4 rk = kNegInf, rv = make_uint4(0, 0, 0, 0);
    if (p0 + pix < p1) {
      rk = *reinterpret_cast<const uint4*>(kp);
      rv = *reinterpret_cast<const uint4*>(kp + HID);
    }
    // one barrier per chunk: the operand tiles are double-buffered, so staging chunk c+1 may overlap the MMAs
    // of chunk c (buffer c&1 was last read two barriers ago)
    for (int j0 = p0; j0 < p1; j0 += LA_CHUNK, cb ^= 1) {
      float kv[8];
      unpack8(rk, kv);
      const uint4 vraw = rv;
      kp += step;
      if (j0 + LA_CHUNK + pix < p1) {
        rk = *reinterpret_cast<const uint4*>(kp);
        rv = *reinterpret_cast<const uint4*>(kp + HID);
      } else {
        rk = kNegInf;
        rv = make_uint4(0, 0, 0, 0);
      }
      float pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        pv[i] = ex2_ftz(fmaf(kv[i], kLog2e, nm2[i]));   // exp(k - m); 0 for the -inf padding
        psum[i] += pv[i];
      }
      *reinterpret_cast<uint4*>(kst + cb * BUF) = pack8(pv);
      *reinterpret_cast<uint4*>(vst + cb * BUF) = vraw;
      __syncthreads();
      mma_atb(&s.Kb[cb][0][0], &s.Vb[cb][0][0], acc, warp, lane);
    }
    __syncthreads();
    {
      const int mt = warp >> 2, nt = warp & 3, row = lane >> 2, col = 2 * (lane & 3);
      s.ctx_loc[mt * 16 + row][nt * 8 + col] = acc[0];
      s.ctx_loc[mt * 16 + row][nt * 8 + col + 1] = acc[1];
      s.ctx_loc[mt * 16 + row + 8][nt * 8 + col] = acc[2];
      s.ctx_loc[mt * 16 + row + 8][nt * 8 + col + 1] = acc[3];
    }
    la_col_reduce<false>(psum, s.mred, s.l_loc, tid);
  }
  cluster.sync();
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5, e = i & 31;
    float M = -INFINITY;
    for (int r = 0; r < CL; ++r) M = fmaxf(M, cluster.map_shared_rank(&s.m_loc[0], r)[d]);
    float Lsum = 0.f, a = 0.f;
    for (int r = 0; r < CL; ++r) {
      const float w = __expf(cluster.map_shared_rank(&s.m_loc[0], r)[d] - M);
      Lsum = fmaf(cluster.map_shared_rank(&s.l_loc[0], r)[d], w, Lsum);
      a = fmaf(cluster.map_shared_rank(&s.ctx_loc[0][0], r)[i], w, a);
    }
    const float c = a / Lsum;
    s.Cb[d][e] = __float2bfloat16_rn(c);
    if (rank == 0) {
      ctx_out[(int64_t)bh * DH * DH + i] = c;
      if (e == 0) {
        kstat[((int64_t)bh * DH + d) * 2] = M;
        kstat[((int64_t)bh * DH + d) * 2 + 1] = Lsum;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: out = (scale * softmax_d(q)) @ ctx for the own pixels
  {
    const lbf* qp = qbase + (int64_t)(p0 + pix) * ld + part * 8;
    const int64_t step = (int64_t)LA_CHUNK * ld;
    lbf* op = out + ((int64_t)b * n + p0 + pix) * out_ld + h * DH + part * 8;     // where chunk c-1 is written
    const int64_t ostep = (int64_t)LA_CHUNK * out_ld;
    lbf* const qst = &s.Kb[0][pix][part * 8];
    const lbf* const ost = &s.Vb[0][pix][part * 8];
    constexpr int BUF = LA_CHUNK * LP;
    uint4 rq = make_uint4(0, 0, 0, 0);
    if (p0 + pix < p1) rq = *reinterpret_cast<const uint4*>(qp);
    // software pipeline with one barrier per chunk: stage q(c) | barrier | write out(c-1) | MMA(c) -> staging(c)
    int cb = 0;
    for (int j0 = p0; j0 < p1; j0 += LA_CHUNK, cb ^= 1) {
      float v[8];
      unpack8(rq, v);
      qp += step;
      if (j0 + LA_CHUNK + pix < p1) rq = *reinterpret_cast<const uint4*>(qp);
      softmax32_quad_fast(v, kScale);
      *reinterpret_cast<uint4*>(qst + cb * BUF) = pack8(v);
      __syncthreads();
      if (j0 > p0) {                                     // every pixel of a non-final chunk is in range
        *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(ost + (cb ^ 1) * BUF);
        op += ostep;
      }
      float acc[2][4] = {};
      mma_xw<true>(&s.Kb[cb][0][0], &s.Cb[0][0], acc, warp, lane);
      {
        const int mt = warp >> 1, nt0 = (warp & 1) * 2, row = lane >> 2, col = 2 * (lane & 3);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          *reinterpret_cast<__nv_bfloat162*>(&s.Vb[cb][mt * 16 + row][(nt0 + t) * 8 + col]) = __floats2bfloat162_rn(acc[t][0], acc[t][1]);
          *reinterpret_cast<__nv_bfloat162*>(&s.Vb[cb][mt * 16 + row + 8][(nt0 + t) * 8 + col]) = __floats2bfloat162_rn(acc[t][2], acc[t][3]);
        }
      }
    }
    __syncthreads();
    if (p0 < p1) {
      const int jl = p0 + ((p1 - p0 - 1) / LA_CHUNK) * LA_CHUNK;      // first pixel of the last chunk
      if (jl + pix < p1)
        *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(ost + (cb ^ 1) * BUF);
    }
  }
  cluster.sync();
}

struct LaBwdTc {                // operand / staging tiles are double-buffered: one barrier per chunk
  lbf Ab[2][LA_CHUNK][LP];     // softmax(q) (phase 1) / softmax_n(k) (phase 2)
  lbf Bb[2][LA_CHUNK][LP];     // dout (phase 1) / v (phase 2)
  lbf Cb[DH][LP];              // ctx, bf16, [d][e]
  lbf Db[DH][LP];              // merged dctx, bf16, [d][e]
  lbf Ob[2][LA_CHUNK][LP];     // dv staging
  float F[2][LA_CHUNK][DH + 1];  // fp32 product staging (dqs / dks)
  float dctx_loc[DH][DH];      // partial dctx (read by the cluster peers)
  float D[DH][DH];             // merged dctx * ctx scratch
  float Dd[DH], kmx[DH], kinv[DH];
};

__global__ void __launch_bounds__(256, 3)
linattn_bwd_tc_kernel(const lbf* __restrict__ dout, int dout_ld, const lbf* __restrict__ qkv, int ld,
                      const float* __restrict__ mem_kv, const float* __restrict__ ctx,
                      const float* __restrict__ kstat, lbf* __restrict__ dqkv, int dld,
                      float* __restrict__ dmem_kv, int n) {
  pdl_prologue();
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int bh = blockIdx.y, b = bh / HEADS, h = bh % HEADS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  extern __shared__ __align__(16) unsigned char la_smem_raw[];
  LaBwdTc& s = *reinterpret_cast<LaBwdTc*>(la_smem_raw);
  const lbf* qbase = qkv + (int64_t)b * n * ld + h * DH;
  const lbf* kbase = qbase + HID;
  const lbf* vbase = qbase + 2 * HID;
  const lbf* gbase = dout + (int64_t)b * n * dout_ld + h * DH;
  lbf* dqbase = dqkv + (int64_t)b * n * dld + h * DH;
  const int chunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  const int cpr = (chunks + CL - 1) / CL;
  const int p0 = min(rank * cpr * LA_CHUNK, n), p1 = min((rank + 1) * cpr * LA_CHUNK, n);
  for (int i = tid; i < DH * DH; i += 256) s.Cb[i >> 5][i & 31] = __float2bfloat16_rn(ctx[(int64_t)bh * DH * DH + i]);
  if (tid < DH) {
    s.kmx[tid] = kstat[((int64_t)bh * DH + tid) * 2];
    s.kinv[tid] = 1.f / kstat[((int64_t)bh * DH + tid) * 2 + 1];
  }
  const int pix = tid >> 2, part = tid & 3;
  // ---- phase 1: dq for the own pixels; partial dctx = P^T dout
  {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // (issue-bound loop: walking pointers, exp as FFMA + ex2.approx, see the forward kernel)
    uint4 rq = make_uint4(0, 0, 0, 0), rg = rq;
    const lbf* qp = qbase + (int64_t)(p0 + pix) * ld + part * 8;
    const lbf* gp = gbase + (int64_t)(p0 + pix) * dout_ld + part * 8;
    lbf* dqp = dqbase + (int64_t)(p0 + pix) * dld + part * 8;        // where chunk c-1's dq goes
    const int64_t qstep = (int64_t)LA_CHUNK * ld, gstep = (int64_t)LA_CHUNK * dout_ld, dstep = (int64_t)LA_CHUNK * dld;
    auto fetch = [&](int j0) {
      if (j0 + pix < p1) {
        rq = *reinterpret_cast<const uint4*>(qp);
        rg = *reinterpret_cast<const uint4*>(gp);
      } else {
        rq = make_uint4(0, 0, 0, 0);
        rg = rq;
      }
    };
    if (p0 < p1) fetch(p0);
    __syncthreads();
    // software pipeline, one barrier per chunk:  stage(c) | barrier | finish dq(c-1) from F[(c-1)&1] | MMA(c)
    float pvp[8];                      // softmax(q) of the previous chunk (this thread's 8 channels)
    bool okp = false;
    auto finish_dq = [&](int cbp) {
      float t = 0.f, dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dv[i] = s.F[cbp][pix][part * 8 + i];
        t = fmaf(pvp[i], dv[i], t);
      }
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
#pragma unroll
      for (int i = 0; i < 8; ++i) dv[i] = kScale * pvp[i] * (dv[i] - t);
      if (okp) *reinterpret_cast<uint4*>(dqp) = pack8(dv);
      dqp += dstep;
    };
    int cb = 0;
    for (int j0 = p0; j0 < p1; j0 += LA_CHUNK, cb ^= 1) {
      float pv[8];
      unpack8(rq, pv);
      const uint4 graw = rg;
      const bool ok = j0 + pix < p1;
      qp += qstep;
      gp += gstep;
      if (j0 + LA_CHUNK < p1) fetch(j0 + LA_CHUNK);
      softmax32_quad_fast(pv, ok ? 1.f : 0.f);           // rows past the end become zeros
      *reinterpret_cast<uint4*>(&s.Ab[cb][pix][part * 8]) = pack8(pv);
      *reinterpret_cast<uint4*>(&s.Bb[cb][pix][part * 8]) = graw;
      __syncthreads();
      if (j0 > p0) finish_dq(cb ^ 1);
      // dqs[j][d] = sum_e dout[j][e] * ctx[d][e]   (W stored [n = d][k = e])
      float dqs[2][4] = {};
      mma_xw<false>(&s.Bb[cb][0][0], &s.Cb[0][0], dqs, warp, lane);
      {
        const int mt = warp >> 1, nt0 = (warp & 1) * 2, row = lane >> 2, col = 2 * (lane & 3);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          s.F[cb][mt * 16 + row][(nt0 + t) * 8 + col] = dqs[t][0];
          s.F[cb][mt * 16 + row][(nt0 + t) * 8 + col + 1] = dqs[t][1];
          s.F[cb][mt * 16 + row + 8][(nt0 + t) * 8 + col] = dqs[t][2];
          s.F[cb][mt * 16 + row + 8][(nt0 + t) * 8 + col + 1] = dqs[t][3];
        }
      }
      mma_atb(&s.Ab[cb][0][0], &s.Bb[cb][0][0], acc, warp, lane);     // dctx += P^T dout
#pragma unroll
      for (int i = 0; i < 8; ++i) pvp[i] = pv[i];
      okp = ok;
    }
    __syncthreads();
    if (p0 < p1) finish_dq(cb ^ 1);
    {
      const int mt = warp >> 2, nt = warp & 3, row = lane >> 2, col = 2 * (lane & 3);
      s.dctx_loc[mt * 16 + row][nt * 8 + col] = acc[0] * kScale;
      s.dctx_loc[mt * 16 + row][nt * 8 + col + 1] = acc[1] * kScale;
      s.dctx_loc[mt * 16 + row + 8][nt * 8 + col] = acc[2] * kScale;
      s.dctx_loc[mt * 16 + row + 8][nt * 8 + col + 1] = acc[3] * kScale;
    }
  }
  cluster.sync();
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5, e = i & 31;
    float a = 0.f;
    for (int r = 0; r < CL; ++r) a += cluster.map_shared_rank(&s.dctx_loc[0][0], r)[i];
    s.Db[d][e] = __float2bfloat16_rn(a);
    s.D[d][e] = a * ctx[(int64_t)bh * DH * DH + i];          // dctx[d][e]*ctx[d][e]
  }
  __syncthreads();
  if (tid < DH) {
    float a = 0.f;
    for (int e = 0; e < DH; ++e) a += s.D[tid][(e + tid) & 31];
    s.Dd[tid] = a;
  }
  __syncthreads();
  // ---- phase 2: dk, dv for the own pixels (+ the memory keys on rank 0)
  {
    const int lead = rank == 0 ? NMEM : 0;
    const int total = (p1 - p0) + lead;
    float nkm2[8], kinv[8], Dd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      nkm2[i] = -s.kmx[part * 8 + i] * kLog2e;
      kinv[i] = s.kinv[part * 8 + i];
      Dd[i] = s.Dd[part * 8 + i];
    }
    const uint4 kNegInf = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf x 8
    uint4 rk = kNegInf, rv = make_uint4(0, 0, 0, 0);
    // row c0 + pix of the (memory rows, own pixels) sequence; never dereferenced outside [p0, p1)
    const lbf* kp = kbase + ((int64_t)p0 + pix - lead) * ld + part * 8;
    const int64_t kstep = (int64_t)LA_CHUNK * ld;
    auto fetch = [&](int c0) {
      const int j = c0 + pix - lead;
      if (j >= 0 && j < p1 - p0) {
        rk = *reinterpret_cast<const uint4*>(kp);
        rv = *reinterpret_cast<const uint4*>(kp + HID);
      } else {
        rk = kNegInf;                                   // exp(-inf - m) = 0: padding rows need no select
        rv = make_uint4(0, 0, 0, 0);
      }
    };
    if (total > 0) fetch(0);
    // software pipeline, one barrier per chunk:  stage(c) | barrier | finish dk/dv(c-1) | MMA(c)
    float ksp[8];
    int jlp = 0;
    bool okp = false;
    auto finish_kv = [&](int cbp) {
      if (!okp) return;
      float dk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dk[i] = ksp[i] * (s.F[cbp][pix][part * 8 + i] - Dd[i]);
      if (jlp < lead) {
        float dvf[8];
        unpack8(*reinterpret_cast<const uint4*>(&s.Ob[cbp][pix][part * 8]), dvf);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          atomicAdd(dmem_kv + ((0 * HEADS + h) * DH + part * 8 + i) * NMEM + jlp, dk[i]);
          atomicAdd(dmem_kv + ((1 * HEADS + h) * DH + part * 8 + i) * NMEM + jlp, dvf[i]);
        }
      } else {
        lbf* o = dqbase + (int64_t)(p0 + jlp - lead) * dld + part * 8;
        *reinterpret_cast<uint4*>(o + HID) = pack8(dk);
        *reinterpret_cast<uint4*>(o + 2 * HID) = *reinterpret_cast<const uint4*>(&s.Ob[cbp][pix][part * 8]);
      }
    };
    int cb = 0;
    for (int c0 = 0; c0 < total; c0 += LA_CHUNK, cb ^= 1) {
      float ks[8];
      unpack8(rk, ks);
      uint4 vraw = rv;
      const int jl = c0 + pix;
      if (jl < lead) {
        float vv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ks[i] = mem_kv[((0 * HEADS + h) * DH + part * 8 + i) * NMEM + jl];
          vv[i] = mem_kv[((1 * HEADS + h) * DH + part * 8 + i) * NMEM + jl];
        }
        vraw = pack8(vv);
      }
      kp += kstep;
      if (c0 + LA_CHUNK < total) fetch(c0 + LA_CHUNK);
      const bool ok = jl < total;
#pragma unroll
      for (int i = 0; i < 8; ++i) ks[i] = ex2_ftz(fmaf(ks[i], kLog2e, nkm2[i])) * kinv[i];
      *reinterpret_cast<uint4*>(&s.Ab[cb][pix][part * 8]) = pack8(ks);
      *reinterpret_cast<uint4*>(&s.Bb[cb][pix][part * 8]) = vraw;
      __syncthreads();
      if (c0 > 0) finish_kv(cb ^ 1);
      float dvv[2][4] = {}, dks[2][4] = {};
      mma_xw<true>(&s.Ab[cb][0][0], &s.Db[0][0], dvv, warp, lane);    // dv[j][e] = sum_d ks[j][d] dctx[d][e]
      mma_xw<false>(&s.Bb[cb][0][0], &s.Db[0][0], dks, warp, lane);   // dks[j][d] = sum_e v[j][e] dctx[d][e]
      {
        const int mt = warp >> 1, nt0 = (warp & 1) * 2, row = lane >> 2, col = 2 * (lane & 3);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int c = (nt0 + t) * 8 + col;
          *reinterpret_cast<__nv_bfloat162*>(&s.Ob[cb][mt * 16 + row][c]) = __floats2bfloat162_rn(dvv[t][0], dvv[t][1]);
          *reinterpret_cast<__nv_bfloat162*>(&s.Ob[cb][mt * 16 + row + 8][c]) = __floats2bfloat162_rn(dvv[t][2], dvv[t][3]);
          s.F[cb][mt * 16 + row][c] = dks[t][0];
          s.F[cb][mt * 16 + row][c + 1] = dks[t][1];
          s.F[cb][mt * 16 + row + 8][c] = dks[t][2];
          s.F[cb][mt * 16 + row + 8][c + 1] = dks[t][3];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) ksp[i] = ks[i];
      jlp = jl;
      okp = ok;
    }
    __syncthreads();
    if (total > 0) finish_kv(cb ^ 1);
  }
  cluster.sync();
}

// cluster size for n pixels: about two waves of CTAs on the machine (each CTA has a fixed cost of two
// reductions and two cluster barriers, so more, smaller CTAs lose), at most 8, at least one chunk each
static int la_cluster_size(int BH, int n) {
  const int chunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  int cl = (int)((2LL * num_sms() * 4 + BH / 2) / BH);
  if (cl > 8) cl = 8;
  if (cl > chunks) cl = chunks;
  return cl < 1 ? 1 : cl;
}

template <typename K, typename... Args>
static cudaError_t la_launch_cluster(K kernel, dim3 grid, int cl, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
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
  pdl_prologue();
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

// ---- bf16 forward on tensor cores (mma.sync m16n8k16, the flash-attention register flow): one CTA per (b, head),
//      one warp per 16 query rows.  S = Q K^T stays in the accumulator registers, the row softmax is done there
//      (quad shuffles), and the probabilities are re-packed in place as the A operand of P V.
constexpr int FA_KVP = 80;     // key/value rows (4 memory rows + up to 64 tokens) padded to a multiple of 16
struct FaTcSmem {
  lbf Q[FA_MAXN][LP];
  lbf K[FA_KVP][LP];
  lbf V[FA_KVP][LP];
};

__global__ void __launch_bounds__(128)
attn_fwd_tc_kernel(const lbf* __restrict__ qkv, int ld, const float* __restrict__ mem_kv,
                   lbf* __restrict__ out, int out_ld, int n) {
  pdl_prologue();
  __shared__ __align__(16) FaTcSmem s;
  const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS, kv = n + NMEM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const lbf* base = qkv + (int64_t)b * n * ld + h * DH;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < FA_MAXN * 4; i += 128) {            // q, k, v rows as raw 16-byte vectors; zero padding
    const int r = i >> 2, part = i & 3;
    uint4 q = zero, k = zero, v = zero;
    if (r < n) {
      const lbf* p = base + (int64_t)r * ld + part * 8;
      q = *reinterpret_cast<const uint4*>(p);
      k = *reinterpret_cast<const uint4*>(p + HID);
      v = *reinterpret_cast<const uint4*>(p + 2 * HID);
    }
    *reinterpret_cast<uint4*>(&s.Q[r][part * 8]) = q;
    *reinterpret_cast<uint4*>(&s.K[r + NMEM][part * 8]) = k;
    *reinterpret_cast<uint4*>(&s.V[r + NMEM][part * 8]) = v;
  }
  for (int i = tid; i < (FA_KVP - FA_MAXN - NMEM) * 4; i += 128) {
    const int r = FA_MAXN + NMEM + (i >> 2), part = i & 3;
    *reinterpret_cast<uint4*>(&s.K[r][part * 8]) = zero;
    *reinterpret_cast<uint4*>(&s.V[r][part * 8]) = zero;
  }
  for (int i = tid; i < NMEM * DH; i += 128) {              // memory key/values, [2][heads][4][32] fp32
    const int r = i >> 5, d = i & 31;
    s.K[r][d] = __float2bfloat16_rn(mem_kv[((0 * HEADS + h) * NMEM + r) * DH + d]);
    s.V[r][d] = __float2bfloat16_rn(mem_kv[((1 * HEADS + h) * NMEM + r) * DH + d]);
  }
  __syncthreads();
  if (warp * 16 >= n) return;
  const int mi = lane >> 3, r8 = lane & 7;
  // S[16][80] = Q K^T
  float sacc[FA_KVP / 8][4] = {};
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, (uint32_t)__cvta_generic_to_shared(&s.Q[warp * 16 + (mi & 1) * 8 + r8][ks * 16 + (mi >> 1) * 8]));
#pragma unroll
    for (int nt = 0; nt < FA_KVP / 8; ++nt) {
      uint32_t bb[2];
      ldsm_x2(bb, (uint32_t)__cvta_generic_to_shared(&s.K[nt * 8 + r8][ks * 16 + (mi & 1) * 8]));
      mma_bf16_16816(sacc[nt], a, bb);
    }
  }
  // softmax over the kv columns of the two rows this thread holds (lane >> 2 and + 8); exp2 with the scale folded in
  const float sc = kScale * kLog2e;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < FA_KVP / 8; ++nt) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = nt * 8 + 2 * (lane & 3) + (j & 1);
      sacc[nt][j] = col < kv ? sacc[nt][j] * sc : -INFINITY;
    }
    m0 = fmaxf(m0, fmaxf(sacc[nt][0], sacc[nt][1]));
    m1 = fmaxf(m1, fmaxf(sacc[nt][2], sacc[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
  uint32_t pa[FA_KVP / 16][4];                               // probabilities as A fragments of P V
#pragma unroll
  for (int nt = 0; nt < FA_KVP / 8; ++nt) {
    const float p0 = ex2_ftz(sacc[nt][0] - m0), p1 = ex2_ftz(sacc[nt][1] - m0);
    const float p2 = ex2_ftz(sacc[nt][2] - m1), p3 = ex2_ftz(sacc[nt][3] - m1);
    l0 += p0 + p1;
    l1 += p2 + p3;
    __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
    pa[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<uint32_t*>(&lo);
    pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // O[16][32] = P V
  float oacc[4][4] = {};
#pragma unroll
  for (int j = 0; j < FA_KVP / 16; ++j) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t bb[2];
      ldsm_x2_t(bb, (uint32_t)__cvta_generic_to_shared(&s.V[j * 16 + (mi & 1) * 8 + r8][nt * 8]));
      mma_bf16_16816(oacc[nt], pa[j], bb);
    }
  }
  const float i0 = rcp_ftz(l0), i1 = rcp_ftz(l1);
  const int row0 = warp * 16 + (lane >> 2), col = 2 * (lane & 3);
  lbf* o = out + ((int64_t)b * n + row0) * out_ld + h * DH + col;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    if (row0 < n) *reinterpret_cast<__nv_bfloat162*>(o + nt * 8) = __floats2bfloat162_rn(oacc[nt][0] * i0, oacc[nt][1] * i0);
    if (row0 + 8 < n)
      *reinterpret_cast<__nv_bfloat162*>(o + (int64_t)8 * out_ld + nt * 8) = __floats2bfloat162_rn(oacc[nt][2] * i1, oacc[nt][3] * i1);
  }
}

// ---- bf16 backward on tensor cores.  Per (b, head): every warp recomputes P for its 16 query rows as in the forward,
//      then dP = dO V^T, dS = P .* (dP - rowsum(P .* dP)) in registers and dQ = scale dS K from the re-packed
//      fragments; P and dS (bf16) are parked in shared memory and the two products that reduce over the query rows,
//      dV = P^T dO and dK = scale dS^T Q, are computed as A^T B (ldmatrix.trans), warp w owning output columns 8w..
constexpr int FA_LPK = FA_KVP + 8;          // row pitch of the parked P / dS tiles: 176 B, conflict-free ldmatrix
struct FaBwdTcSmem {
  lbf Q[FA_MAXN][LP];
  lbf K[FA_KVP][LP];
  lbf V[FA_KVP][LP];
  lbf dO[FA_MAXN][LP];
  lbf P[FA_MAXN][FA_LPK];
  lbf dS[FA_MAXN][FA_LPK];
};

__global__ void __launch_bounds__(128)
attn_bwd_tc_kernel(const lbf* __restrict__ dout, int dout_ld, const lbf* __restrict__ qkv, int ld,
                   const float* __restrict__ mem_kv, lbf* __restrict__ dqkv, int dld,
                   float* __restrict__ dmem_kv, int n) {
  pdl_prologue();
  __shared__ __align__(16) FaBwdTcSmem s;
  const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS, kv = n + NMEM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const lbf* base = qkv + (int64_t)b * n * ld + h * DH;
  const lbf* gbase = dout + (int64_t)b * n * dout_ld + h * DH;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < FA_MAXN * 4; i += 128) {
    const int r = i >> 2, part = i & 3;
    uint4 q = zero, k = zero, v = zero, g = zero;
    if (r < n) {
      const lbf* p = base + (int64_t)r * ld + part * 8;
      q = *reinterpret_cast<const uint4*>(p);
      k = *reinterpret_cast<const uint4*>(p + HID);
      v = *reinterpret_cast<const uint4*>(p + 2 * HID);
      g = *reinterpret_cast<const uint4*>(gbase + (int64_t)r * dout_ld + part * 8);
    }
    *reinterpret_cast<uint4*>(&s.Q[r][part * 8]) = q;
    *reinterpret_cast<uint4*>(&s.K[r + NMEM][part * 8]) = k;
    *reinterpret_cast<uint4*>(&s.V[r + NMEM][part * 8]) = v;
    *reinterpret_cast<uint4*>(&s.dO[r][part * 8]) = g;
  }
  for (int i = tid; i < (FA_KVP - FA_MAXN - NMEM) * 4; i += 128) {
    const int r = FA_MAXN + NMEM + (i >> 2), part = i & 3;
    *reinterpret_cast<uint4*>(&s.K[r][part * 8]) = zero;
    *reinterpret_cast<uint4*>(&s.V[r][part * 8]) = zero;
  }
  for (int i = tid; i < NMEM * DH; i += 128) {
    const int r = i >> 5, d = i & 31;
    s.K[r][d] = __float2bfloat16_rn(mem_kv[((0 * HEADS + h) * NMEM + r) * DH + d]);
    s.V[r][d] = __float2bfloat16_rn(mem_kv[((1 * HEADS + h) * NMEM + r) * DH + d]);
  }
  __syncthreads();
  const int mi = lane >> 3, r8 = lane & 7;
  const int qrow = lane >> 2, qcol = 2 * (lane & 3);
  constexpr int NT = FA_KVP / 8;
  // ---- S = Q K^T, P = softmax(scale S) for this warp's 16 query rows (registers)
  float pr[NT][4] = {};
  uint32_t qa[2][4], ga[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    ldsm_x4(qa[ks], (uint32_t)__cvta_generic_to_shared(&s.Q[warp * 16 + (mi & 1) * 8 + r8][ks * 16 + (mi >> 1) * 8]));
    ldsm_x4(ga[ks], (uint32_t)__cvta_generic_to_shared(&s.dO[warp * 16 + (mi & 1) * 8 + r8][ks * 16 + (mi >> 1) * 8]));
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t bb[2];
      ldsm_x2(bb, (uint32_t)__cvta_generic_to_shared(&s.K[nt * 8 + r8][ks * 16 + (mi & 1) * 8]));
      mma_bf16_16816(pr[nt], qa[ks], bb);
    }
  }
  {
    const float sc = kScale * kLog2e;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) pr[nt][j] = (nt * 8 + qcol + (j & 1)) < kv ? pr[nt][j] * sc : -INFINITY;
      m0 = fmaxf(m0, fmaxf(pr[nt][0], pr[nt][1]));
      m1 = fmaxf(m1, fmaxf(pr[nt][2], pr[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      pr[nt][0] = ex2_ftz(pr[nt][0] - m0);
      pr[nt][1] = ex2_ftz(pr[nt][1] - m0);
      pr[nt][2] = ex2_ftz(pr[nt][2] - m1);
      pr[nt][3] = ex2_ftz(pr[nt][3] - m1);
      l0 += pr[nt][0] + pr[nt][1];
      l1 += pr[nt][2] + pr[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = rcp_ftz(l0), i1 = rcp_ftz(l1);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      pr[nt][0] *= i0;
      pr[nt][1] *= i0;
      pr[nt][2] *= i1;
      pr[nt][3] *= i1;
    }
  }
  // ---- dP = dO V^T; dS = P .* (dP - rowsum(P .* dP))
  float dp[NT][4] = {};
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t bb[2];
      ldsm_x2(bb, (uint32_t)__cvta_generic_to_shared(&s.V[nt * 8 + r8][ks * 16 + (mi & 1) * 8]));
      mma_bf16_16816(dp[nt], ga[ks], bb);
    }
  }
  float t0 = 0.f, t1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    t0 = fmaf(pr[nt][0], dp[nt][0], fmaf(pr[nt][1], dp[nt][1], t0));
    t1 = fmaf(pr[nt][2], dp[nt][2], fmaf(pr[nt][3], dp[nt][3], t1));
  }
  t0 += __shfl_xor_sync(0xffffffffu, t0, 1);
  t0 += __shfl_xor_sync(0xffffffffu, t0, 2);
  t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
  t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
  uint32_t dsa[NT / 2][4];                               // dS as A fragments of dS K
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const float d0 = pr[nt][0] * (dp[nt][0] - t0), d1 = pr[nt][1] * (dp[nt][1] - t0);
    const float d2 = pr[nt][2] * (dp[nt][2] - t1), d3 = pr[nt][3] * (dp[nt][3] - t1);
    __nv_bfloat162 lo = __floats2bfloat162_rn(d0, d1), hi = __floats2bfloat162_rn(d2, d3);
    dsa[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<uint32_t*>(&lo);
    dsa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
    // park P and dS (bf16) for the products over the query rows
    *reinterpret_cast<__nv_bfloat162*>(&s.dS[warp * 16 + qrow][nt * 8 + qcol]) = lo;
    *reinterpret_cast<__nv_bfloat162*>(&s.dS[warp * 16 + qrow + 8][nt * 8 + qcol]) = hi;
    *reinterpret_cast<__nv_bfloat162*>(&s.P[warp * 16 + qrow][nt * 8 + qcol]) = __floats2bfloat162_rn(pr[nt][0], pr[nt][1]);
    *reinterpret_cast<__nv_bfloat162*>(&s.P[warp * 16 + qrow + 8][nt * 8 + qcol]) = __floats2bfloat162_rn(pr[nt][2], pr[nt][3]);
  }
  // ---- dQ = scale * dS K
  {
    float dq[4][4] = {};
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bb[2];
        ldsm_x2_t(bb, (uint32_t)__cvta_generic_to_shared(&s.K[j * 16 + (mi & 1) * 8 + r8][nt * 8]));
        mma_bf16_16816(dq[nt], dsa[j], bb);
      }
    }
    const int row0 = warp * 16 + qrow;
    lbf* o = dqkv + ((int64_t)b * n + row0) * dld + h * DH + qcol;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (row0 < n)
        *reinterpret_cast<__nv_bfloat162*>(o + nt * 8) = __floats2bfloat162_rn(dq[nt][0] * kScale, dq[nt][1] * kScale);
      if (row0 + 8 < n)
        *reinterpret_cast<__nv_bfloat162*>(o + (int64_t)8 * dld + nt * 8) =
            __floats2bfloat162_rn(dq[nt][2] * kScale, dq[nt][3] * kScale);
    }
  }
  __syncthreads();
  // ---- dV = P^T dO, dK = scale * dS^T Q: [80 x 32] each, K = 64 query rows; warp w owns columns 8w .. 8w+7
  {
    float dv[FA_KVP / 16][4] = {}, dk[FA_KVP / 16][4] = {};
#pragma unroll
    for (int ks = 0; ks < FA_MAXN / 16; ++ks) {
      uint32_t bg[2], bq[2];
      ldsm_x2_t(bg, (uint32_t)__cvta_generic_to_shared(&s.dO[ks * 16 + (mi & 1) * 8 + r8][warp * 8]));
      ldsm_x2_t(bq, (uint32_t)__cvta_generic_to_shared(&s.Q[ks * 16 + (mi & 1) * 8 + r8][warp * 8]));
#pragma unroll
      for (int mt = 0; mt < FA_KVP / 16; ++mt) {
        uint32_t a[4];
        ldsm_x4_t(a, (uint32_t)__cvta_generic_to_shared(&s.P[ks * 16 + (mi >> 1) * 8 + r8][mt * 16 + (mi & 1) * 8]));
        mma_bf16_16816(dv[mt], a, bg);
        ldsm_x4_t(a, (uint32_t)__cvta_generic_to_shared(&s.dS[ks * 16 + (mi >> 1) * 8 + r8][mt * 16 + (mi & 1) * 8]));
        mma_bf16_16816(dk[mt], a, bq);
      }
    }
#pragma unroll
    for (int mt = 0; mt < FA_KVP / 16; ++mt) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = mt * 16 + qrow + half * 8;            // key/value row
        const int d = warp * 8 + qcol;
        const float k0 = dk[mt][half * 2] * kScale, k1 = dk[mt][half * 2 + 1] * kScale;
        const float v0 = dv[mt][half * 2], v1 = dv[mt][half * 2 + 1];
        if (c < NMEM) {
          float* mk = dmem_kv + ((0 * HEADS + h) * NMEM + c) * DH + d;
          float* mv = dmem_kv + ((1 * HEADS + h) * NMEM + c) * DH + d;
          atomicAdd(mk, k0);
          atomicAdd(mk + 1, k1);
          atomicAdd(mv, v0);
          atomicAdd(mv + 1, v1);
        } else if (c < kv) {
          lbf* p = dqkv + ((int64_t)b * n + (c - NMEM)) * dld + h * DH + d;
          *reinterpret_cast<__nv_bfloat162*>(p + HID) = __floats2bfloat162_rn(k0, k1);
          *reinterpret_cast<__nv_bfloat162*>(p + 2 * HID) = __floats2bfloat162_rn(v0, v1);
        }
      }
    }
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
  pdl_prologue();
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
  {
    const int cl = la_cluster_size(B * HEADS, n);
    dim3 grid(cl, B * HEADS);
    cudaError_t e;
    if (dtype == B200DM_F32) {
      const size_t smem = sizeof(LaFwdSmem);
      cudaFuncSetAttribute(linattn_fwd_cluster_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e = la_launch_cluster(linattn_fwd_cluster_kernel<float>, grid, cl, smem, st, (const float*)qkv, (int)qkv_ld,
                            mem_kv, ctx, kstat, (float*)out, (int)out_ld, (int)n);
    } else {
      e = la_launch_cluster(linattn_fwd_tc_kernel, grid, cl, 0, st, (const bf16*)qkv, (int)qkv_ld, mem_kv, ctx,
                            kstat, (bf16*)out, (int)out_ld, (int)n);
    }
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "linattn_fwd: launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return check_launch("linattn_fwd");
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
  {
    (void)dctx;
    const int cl = la_cluster_size(B * HEADS, n);
    dim3 grid(cl, B * HEADS);
    cudaError_t e;
    if (dtype == B200DM_F32) {
      const size_t smem = sizeof(LaBwdSmem);
      cudaFuncSetAttribute(linattn_bwd_cluster_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e = la_launch_cluster(linattn_bwd_cluster_kernel<float>, grid, cl, smem, st, (const float*)dout, (int)dout_ld,
                            (const float*)qkv, (int)qkv_ld, mem_kv, ctx, kstat, (float*)dqkv, (int)dqkv_ld, dmem_kv,
                            (int)n);
    } else {
      const size_t smem = sizeof(LaBwdTc);
      cudaFuncSetAttribute(linattn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e = la_launch_cluster(linattn_bwd_tc_kernel, grid, cl, smem, st, (const bf16*)dout, (int)dout_ld,
                            (const bf16*)qkv, (int)qkv_ld, mem_kv, ctx, kstat, (bf16*)dqkv, (int)dqkv_ld, dmem_kv,
                            (int)n);
    }
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "linattn_bwd: launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return check_launch("linattn_bwd");
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
    launch_k(attn_fwd_kernel<float>, B * HEADS, 256, smem, st, (const float*)qkv, qkv_ld, mem_kv, (float*)out, out_ld, n);
  } else if (qkv_ld % 8 == 0 && out_ld % 2 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 3) == 0) {
    launch_k(attn_fwd_tc_kernel, B * HEADS, 128, 0, st, (const bf16*)qkv, qkv_ld, mem_kv, (bf16*)out, out_ld, n);
  } else {
    cudaFuncSetAttribute(attn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(attn_fwd_kernel<bf16>, B * HEADS, 256, smem, st, (const bf16*)qkv, qkv_ld, mem_kv, (bf16*)out, out_ld, n);
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
    launch_k(attn_bwd_kernel<float>, B * HEADS, 256, smem, st, (const float*)dout, dout_ld, (const float*)qkv, qkv_ld, mem_kv, (float*)dqkv, dqkv_ld, dmem_kv, n);
  } else if (qkv_ld % 8 == 0 && dout_ld % 8 == 0 && dqkv_ld % 2 == 0 && ((uintptr_t)qkv & 15) == 0 &&
             ((uintptr_t)dout & 15) == 0 && ((uintptr_t)dqkv & 3) == 0) {
    launch_k(attn_bwd_tc_kernel, B * HEADS, 128, 0, st, (const bf16*)dout, dout_ld, (const bf16*)qkv, qkv_ld, mem_kv,
             (bf16*)dqkv, dqkv_ld, dmem_kv, n);
  } else {
    cudaFuncSetAttribute(attn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(attn_bwd_kernel<bf16>, B * HEADS, 256, smem, st, (const bf16*)dout, dout_ld, (const bf16*)qkv, qkv_ld, mem_kv, (bf16*)dqkv, dqkv_ld, dmem_kv, n);
  }
  count_launch();
  return check_launch("attn_bwd");
}
