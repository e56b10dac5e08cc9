// Weight packing (fp32 master -> GEMM operand layouts), fused Adam, EMA lerp.
// Reference: torch.optim.Adam configured at ddpm.py:1053-1059; ema_pytorch.EMA.update at :1047-1048.
#include "common.cuh"

namespace b200dm {

// master w: [taps][Cout][Cin] fp32.  wf: same order in T.  wt: [taps'][Cin][Cout] in T with
// taps' = flip ? taps-1-t : t  (the data-gradient operand: rotate the filter 180 deg, swap in/out).
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wt,
                                   int taps, int Cout, int Cin, int flip, long long s_tap,
                                   long long s_co, long long s_ci) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const float* wsrc = w + (int64_t)t * s_tap;
  for (int r = ty; r < 32; r += 8) {
    int co = co0 + r, ci = ci0 + tx;
    float v = (co < Cout && ci < Cin) ? wsrc[(int64_t)co * s_co + (int64_t)ci * s_ci] : 0.f;
    tile[r][tx] = v;
    if (wf && co < Cout && ci < Cin) Elem<T>::st(wf + ((int64_t)t * Cout + co) * Cin + ci, v);
  }
  __syncthreads();
  if (wt) {
    const int tt = flip ? taps - 1 - t : t;
    for (int r = ty; r < 32; r += 8) {
      int ci = ci0 + r, co = co0 + tx;
      if (co < Cout && ci < Cin) Elem<T>::st(wt + ((int64_t)tt * Cin + ci) * Cout + co, tile[tx][r]);
    }
  }
}

// All conv weights of the network in ONE launch: `table` lists the tensors, each CTA finds its
// (tensor, tap, 64x64 tile) by binary search over the tile prefix sums.  256 threads; float4 loads along Cin,
// 8-byte (4 x bf16) / 16-byte (4 x fp32) stores for both the forward copy and the transposed data-gradient copy.
constexpr int PK = 64;
template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(a, b);
  h[1] = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const b200dm_pack_entry* __restrict__ table, int n, int tile_first) {
  pdl_prologue();
  __shared__ float tile[PK][PK + 1];
  int lo = 0, hi = n - 1;
  const int bid = blockIdx.x + tile_first;             // tile index in the numbering of the FULL table
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (table[mid].tile_begin <= bid) lo = mid; else hi = mid - 1;
  }
  const b200dm_pack_entry e = table[lo];
  int local = bid - e.tile_begin;
  const int per_tap = e.tiles_ci * e.tiles_co;
  const int t = local / per_tap;
  local -= t * per_tap;
  const int ci0 = (local % e.tiles_ci) * PK, co0 = (local / e.tiles_ci) * PK;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int c4 = (tid & 15) * 4, r0 = tid >> 4;          // 4 consecutive columns, rows r0 + 16 i
  const float* wsrc = e.w + (int64_t)t * e.s_tap;
  T* wf = (T*)e.wf;
  T* wt = (T*)e.wt;
  const bool full = co0 + PK <= e.Cout && ci0 + PK <= e.Cin;
  const bool vec = full && e.s_ci == 1 && (e.s_co & 3) == 0 && (e.s_tap & 3) == 0 && (e.Cin & 3) == 0 && (e.Cout & 3) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + 16 * i, co = co0 + r, ci = ci0 + c4;
    float v[4];
    if (vec) {
      const float4 f = *reinterpret_cast<const float4*>(wsrc + (int64_t)co * e.s_co + ci);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = (co < e.Cout && ci + j < e.Cin) ? wsrc[(int64_t)co * e.s_co + (int64_t)(ci + j) * e.s_ci] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[r][c4 + j] = v[j];
    if (wf) {
      T* dst = wf + ((int64_t)t * e.Cout + co) * e.Cin + ci;
      if (vec) {
        store4<T>(dst, v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (co < e.Cout && ci + j < e.Cin) Elem<T>::st(dst + j, v[j]);
      }
    }
  }
  __syncthreads();
  if (wt) {
    const int tt = e.flip ? e.taps - 1 - t : t;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 16 * i, ci = ci0 + r, co = co0 + c4;     // row of the transposed tile = input channel
      T* dst = wt + ((int64_t)tt * e.Cin + ci) * e.Cout + co;
      if (vec) {
        store4<T>(dst, tile[c4][r], tile[c4 + 1][r], tile[c4 + 2][r], tile[c4 + 3][r]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ci < e.Cin && co + j < e.Cout) Elem<T>::st(dst + j, tile[c4 + j][r]);
      }
    }
  }
}

// Weights of the fused nearest-2x upsample + 3x3 conv (conv_fwd mode 3): for output phase (a, b) and 2x2 tap
// (r, c) the sum of the 3x3 taps that read the same source pixel.  w: master layout [ky*3+kx][Cout][Cin] fp32;
// out: [tap r*2+c][phase a*2+b][Cout][Cin] bf16.
__global__ void pack_upconv_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin) {
  pdl_prologue();
  const int n = Cout * Cin;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16 * n) return;
  const int e = i % n, ph = (i / n) & 3, t = i / (4 * n);
  const int a = ph >> 1, b = ph & 1, r = t >> 1, c = t & 1;
  // phase 0: tap 0 <- k {0}, tap 1 <- k {1, 2};  phase 1: tap 0 <- k {0, 1}, tap 1 <- k {2}
  const int ky0 = a == 0 ? (r == 0 ? 0 : 1) : (r == 0 ? 0 : 2), ky1 = a == 0 ? (r == 0 ? 0 : 2) : (r == 0 ? 1 : 2);
  const int kx0 = b == 0 ? (c == 0 ? 0 : 1) : (c == 0 ? 0 : 2), kx1 = b == 0 ? (c == 0 ? 0 : 2) : (c == 0 ? 1 : 2);
  float acc = 0.f;
  for (int ky = ky0; ky <= ky1; ++ky)
    for (int kx = kx0; kx <= kx1; ++kx) acc += w[(ky * 3 + kx) * n + e];
  out[i] = __float2bfloat16_rn(acc);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, int64_t n, float step_size, float beta1, float beta2, float eps,
            float weight_decay, float bc2_sqrt, float grad_scale) {
  pdl_prologue();
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= grad_scale;
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    mi = mi + (1.f - beta1) * (gi - mi);          // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * beta2 + (1.f - beta2) * gi * gi;    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);           // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  // one element per thread and iteration: measured 5.05 TB/s on the flat arenas; a float4 / streaming-policy version
  // was slower (4.4 TB/s)
  for (int64_t i = tid; i < n; i += nthr) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[i], mi, vi);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

__global__ void __launch_bounds__(256)
ema_kernel(float* __restrict__ ema, const float* __restrict__ online, int64_t n, float w) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float e = ema[i];
    ema[i] = e + w * (online[i] - e);  // ema.lerp_(online, 1 - decay)
  }
}

}  // namespace b200dm

using namespace b200dm;

extern "C" int b200dm_pack_conv_weight(int32_t dtype, const float* w, void* wf, void* wt, int32_t taps,
                                       int32_t Cout, int32_t Cin, int32_t flip, int64_t s_tap,
                                       int64_t s_co, int64_t s_ci, void* stream) {
  B200DM_REQUIRE(taps > 0 && Cout > 0 && Cin > 0, B200DM_ERR_SHAPE, "pack_conv_weight: bad shape");
  if (s_tap == 0 && s_co == 0 && s_ci == 0) { s_tap = (int64_t)Cout * Cin; s_co = Cin; s_ci = 1; }
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, taps), block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(pack_weight_kernel<float>, grid, block, 0, st, w, (float*)wf, (float*)wt, taps, Cout, Cin, flip, s_tap, s_co, s_ci);
  else
    launch_k(pack_weight_kernel<__nv_bfloat16>, grid, block, 0, st, w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wt, taps, Cout, Cin, flip, s_tap, s_co, s_ci);
  count_launch();
  return check_launch("pack_conv_weight");
}

extern "C" int b200dm_pack_conv_weights_batched(int32_t dtype, const b200dm_pack_entry* table,
                                                int32_t n_entries, int32_t total_tiles, void* stream) {
  B200DM_REQUIRE(table && n_entries > 0 && total_tiles > 0, B200DM_ERR_SHAPE, "pack_conv_weights_batched: empty table");
  dim3 block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(pack_weights_batched_kernel<float>, total_tiles, block, 0, st, table, n_entries, 0);
  else
    launch_k(pack_weights_batched_kernel<__nv_bfloat16>, total_tiles, block, 0, st, table, n_entries, 0);
  count_launch();
  return check_launch("pack_conv_weights_batched");
}

extern "C" int b200dm_pack_conv_weights_range(int32_t dtype, const b200dm_pack_entry* entries, int32_t n_entries,
                                              int32_t tile_first, int32_t n_tiles, void* stream) {
  B200DM_REQUIRE(entries && n_entries > 0 && n_tiles > 0 && tile_first >= 0, B200DM_ERR_SHAPE,
                 "pack_conv_weights_range: empty range");
  dim3 block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(pack_weights_batched_kernel<float>, n_tiles, block, 0, st, entries, n_entries, tile_first);
  else
    launch_k(pack_weights_batched_kernel<__nv_bfloat16>, n_tiles, block, 0, st, entries, n_entries, tile_first);
  count_launch();
  return check_launch("pack_conv_weights_range");
}

extern "C" int b200dm_pack_upconv_weight(const float* w, void* out, int32_t Cout, int32_t Cin, void* stream) {
  B200DM_REQUIRE(w && out && Cout > 0 && Cin > 0, B200DM_ERR_SHAPE, "pack_upconv_weight: bad arguments");
  const int total = 16 * Cout * Cin;
  launch_k(pack_upconv_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)out, Cout, Cin);
  count_launch();
  return check_launch("pack_upconv_weight");
}

extern "C" int b200dm_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                float grad_scale, void* stream) {
  return b200dm_adam_step_bg(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, 16, stream);
}

extern "C" int b200dm_adam_step_bg(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                   float grad_scale, int32_t ctas_per_sm, void* stream) {
  B200DM_REQUIRE(n > 0 && step >= 1 && ctas_per_sm >= 1, B200DM_ERR_SHAPE, "adam_step: n=%lld step=%d", (long long)n, step);
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  float step_size = (float)((double)lr / bc1);
  float bc2_sqrt = (float)sqrt(bc2);
  int64_t blocks = (n + 255) / 256, cap = (int64_t)num_sms() * ctas_per_sm;
  launch_k(adam_kernel, (unsigned)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream, 
      p, g, m, v, n, step_size, beta1, beta2, eps, weight_decay, bc2_sqrt, grad_scale);
  count_launch();
  return check_launch("adam_step");
}

// ---- gradient compression for the wire (opt-in bf16 all-reduce, b200dm.distributed.GradSync(wire="bf16")) ----------
namespace b200dm {
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            int64_t n4) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 o;
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                                            int64_t n4) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 u = reinterpret_cast<const uint2*>(x)[i];
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    reinterpret_cast<float4*>(y)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
}
}  // namespace b200dm

extern "C" int b200dm_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
  B200DM_REQUIRE(n >= 0 && n % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 7) == 0, B200DM_ERR_SHAPE,
                 "cast_f32_bf16: n %% 4 == 0 and aligned pointers required");
  if (n == 0) return B200DM_OK;
  int64_t blocks = (n / 4 + 255) / 256, cap = (int64_t)b200dm::num_sms() * 8;
  launch_k(b200dm::cast_f32_bf16_kernel, (int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream, x,
           (__nv_bfloat16*)y, n / 4);
  b200dm::count_launch();
  return b200dm::check_launch("cast_f32_bf16");
}

extern "C" int b200dm_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream) {
  B200DM_REQUIRE(n >= 0 && n % 4 == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)x & 7) == 0, B200DM_ERR_SHAPE,
                 "cast_bf16_f32: n %% 4 == 0 and aligned pointers required");
  if (n == 0) return B200DM_OK;
  int64_t blocks = (n / 4 + 255) / 256, cap = (int64_t)b200dm::num_sms() * 8;
  launch_k(b200dm::cast_bf16_f32_kernel, (int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream,
           (const __nv_bfloat16*)x, y, n / 4);
  b200dm::count_launch();
  return b200dm::check_launch("cast_bf16_f32");
}

extern "C" int b200dm_ema_update(float* ema, const float* online, int64_t n, float decay, void* stream) {
  B200DM_REQUIRE(n > 0, B200DM_ERR_SHAPE, "ema_update: empty");
  int64_t blocks = (n + 255) / 256, cap = (int64_t)num_sms() * 16;
  launch_k(ema_kernel, (unsigned)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream, ema, online, n, 1.f - decay);
  count_launch();
  return check_launch("ema_update");
}
