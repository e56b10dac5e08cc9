// Shared device/host helpers for libb200dm (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200dm.h"

namespace b200dm {

// ---- error plumbing (thread-local, see b200dm_last_error) -----------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError -> B200DM_ERR_CUDA

#define B200DM_REQUIRE(cond, code, ...)  \
  do {                                   \
    if (!(cond)) {                       \
      b200dm::set_error(__VA_ARGS__);    \
      return (code);                     \
    }                                    \
  } while (0)

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// Every kernel is launched with the programmatic-stream-serialization attribute and starts with
// griddepcontrol.launch_dependents + griddepcontrol.wait: the next kernel of the stream may be scheduled while
// this one drains, runs its prologue (barrier / TMEM / descriptor setup) and blocks in `pdl_wait()` until this
// kernel has completed and its writes are visible.  Because EVERY kernel waits, completion is transitive along
// the stream.  Nothing before `pdl_wait()` may touch global memory.  B200DM_PDL=0 disables the attribute.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// SMs the persistent grids size themselves for: the device's SM count minus the SMs set aside for a concurrent
// collective (b200dm_set_reserved_sms).  A conv CTA needs a whole SM (>= 200 KiB of shared memory), so while NCCL's
// CTAs hold SMs a 148-CTA grid runs as two waves; sizing it to the free SMs keeps it at one.
int reserved_sms();
inline int device_sms() {      // the device's SM count, whatever is reserved (sizes that must not change between calls)
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
inline int num_sms() {
  const int n = device_sms(), r = reserved_sms();
  return n - r > 16 ? n - r : n;
}

// ---- element access templated on the activation dtype -----------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 8-element vector load/store (32 B fp32 / 16 B bf16); pointers must be aligned accordingly.
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float silu_f(float z) { return z / (1.f + __expf(-z)); }
__device__ __forceinline__ float silu_grad_f(float z) {
  float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}

// ---- Philox4x32-10 (Salmon et al.), counter-based: key=(seed), counter=(idx, stream) ------------
struct Philox {
  static __device__ __forceinline__ uint4 round(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  }
  static __device__ __forceinline__ uint4 gen(uint64_t seed, uint64_t idx, uint64_t stream) {
    uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 c = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream,
                         (uint32_t)(stream >> 32));
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      c = round(c, k);
      k.x += 0x9E3779B9u;
      k.y += 0xBB67AE85u;
    }
    return c;
  }
  // four N(0,1) samples from one 128-bit block (Box-Muller on two pairs)
  static __device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t idx, uint64_t stream) {
    uint4 r = gen(seed, idx, stream);
    const float S = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)r.x + 0.5f) * S, u1 = ((float)r.y + 0.5f) * S;
    float u2 = ((float)r.z + 0.5f) * S, u3 = ((float)r.w + 0.5f) * S;
    u0 = fminf(fmaxf(u0, 1e-12f), 1.f);
    u2 = fminf(fmaxf(u2, 1e-12f), 1.f);
    float ra = sqrtf(-2.f * __logf(u0)), rb = sqrtf(-2.f * __logf(u2));
    float sa, ca, sb, cb;
    __sincosf(6.283185307179586f * u1, &sa, &ca);
    __sincosf(6.283185307179586f * u3, &sb, &cb);
    return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
  }
};

}  // namespace b200dm
