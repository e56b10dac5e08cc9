// GroupNorm(8)+FiLM+SiLU(+residual) forward/backward and RMSNorm forward/backward, NHWC.
// Reference: Block.forward ddpm.py:164-173, ResnetBlock.forward :189-200, RMSNorm :107-113.
// All statistics and arithmetic in fp32; tensors in the activation dtype.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200dm {

// ---- GroupNorm statistics: one CTA per (sample, group), two passes (mean, then centred variance) ----
template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T* __restrict__ x, int ld, float* __restrict__ stats, int HW, int C, int G,
                float eps) {
  pdl_prologue();
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int gs = C / G, vpp = gs / 8;  // 8-wide vectors per pixel within the group
  const int64_t nvec = (int64_t)HW * vpp;
  const T* base = x + (int64_t)b * HW * ld + g * gs;
  __shared__ float red[8];
  __shared__ float s_mean;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
    int64_t p = i / vpp;
    int v = (int)(i - p * vpp);
    float f[8];
    ld8(base + p * ld + v * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += f[j];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    s_mean = s / ((float)HW * (float)gs);
  }
  __syncthreads();
  const float mean = s_mean;
  acc = 0.f;
  for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
    int64_t p = i / vpp;
    int v = (int)(i - p * vpp);
    float f[8];
    ld8(base + p * ld + v * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d = f[j] - mean;
      acc = fmaf(d, d, acc);
    }
  }
  acc = warp_sum(acc);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    float var = s / ((float)HW * (float)gs);
    stats[(b * G + g) * 2 + 0] = mean;
    stats[(b * G + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

// ---- apply: y = silu(((x-mean)*rstd*gamma+beta)*(1+scale)+shift) (+res) -------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gn_apply_fwd_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ stats,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ film, int film_ld, const T* __restrict__ res,
                    int res_ld, T* __restrict__ y, int y_ld, int64_t total8, int HW, int C, int G) {
  pdl_prologue();
  const int C8 = C / 8, gs = C / G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c0 = (int)(i % C8) * 8;
    int64_t p = i / C8;
    int b = (int)(p / HW);
    int g = c0 / gs;
    float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
    float v[8], r[8];
    ld8(x + p * x_ld + c0, v);
    if (res) ld8(res + p * res_ld + c0, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float n = (v[j] - mean) * rstd * gamma[c0 + j] + beta[c0 + j];
      if (film) n = n * (film[(int64_t)b * film_ld + c0 + j] + 1.f) + film[(int64_t)b * film_ld + C + c0 + j];
      float o = silu_f(n);
      v[j] = res ? o + r[j] : o;
    }
    st8(y + p * y_ld + c0, v);
  }
}

// ---- backward pass 1: sums[b][c] = (sum_p dz, sum_p dz*xnorm, sum_p x) ---------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_reduce_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ x, int x_ld,
                     const float* __restrict__ stats, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ film, int film_ld,
                     float* __restrict__ sums, int HW, int C, int G, int pix_per_block) {
  pdl_prologue();
  extern __shared__ float sred[];  // [lanes][C][3]
  const int b = blockIdx.y;
  const int C8 = C / 8, gs = C / G;
  const int lanes = blockDim.x / C8;          // pixel lanes
  const int cv = threadIdx.x % C8, lane = threadIdx.x / C8;
  const int c0 = cv * 8;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, HW);
  float s1[8] = {}, s2[8] = {}, s0[8] = {};
  if (lane < lanes) {
    const int g = c0 / gs;
    const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
    float ga[8], be[8], sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ga[j] = gamma[c0 + j];
      be[j] = beta[c0 + j];
      sc[j] = film ? film[(int64_t)b * film_ld + c0 + j] + 1.f : 1.f;
      sh[j] = film ? film[(int64_t)b * film_ld + C + c0 + j] : 0.f;
    }
    for (int p = p0 + lane; p < p1; p += lanes) {
      int64_t row = (int64_t)b * HW + p;
      float xv[8], gv[8];
      ld8(x + row * x_ld + c0, xv);
      ld8(dy + row * dy_ld + c0, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float xn = (xv[j] - mean) * rstd;
        float z = (xn * ga[j] + be[j]) * sc[j] + sh[j];
        float dz = gv[j] * silu_grad_f(z);
        s1[j] += dz;
        s2[j] = fmaf(dz, xn, s2[j]);
        s0[j] += xv[j];
      }
    }
  }
  if (lane < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sred[(lane * C + c0 + j) * 3] = s1[j];
      sred[(lane * C + c0 + j) * 3 + 1] = s2[j];
      sred[(lane * C + c0 + j) * 3 + 2] = s0[j];
    }
  }
  __syncthreads();
  // deterministic: each CTA writes its own partial [b][chunk][c][3]; pass 2 sums the chunks
  float* part = sums + ((int64_t)b * gridDim.x + blockIdx.x) * C * 3;
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sred[l * C * 3 + i];
    part[i] = s;
  }
}

// ---- backward pass 2: FiLM grads, group means, dgamma/dbeta; one CTA per sample ---------------------
__global__ void __launch_bounds__(256)
gn_bwd_params_kernel(const float* __restrict__ sums, const float* __restrict__ stats,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ film, int film_ld, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dfilm, float* __restrict__ dbias,
                     float* __restrict__ gmeans, int B, int HW, int C, int G, int chunks) {
  pdl_prologue();
  // one CTA per group: thread = (channel of the group, sample lane); every parameter gradient of the
  // group's channels is reduced over the batch inside the CTA -> plain (+=) stores, no global atomics
  extern __shared__ float sh[];            // tot[gs][3] per sample lane, then reductions
  const int g = blockIdx.x, gs = C / G;
  const int cl = threadIdx.x % gs, bl = threadIdx.x / gs, BL = blockDim.x / gs;
  const int c = g * gs + cl;
  float* accg = sh;                        // [BL][gs][3] : dgamma, dbeta, dbias partials
  float* gsum = sh + BL * gs * 3;          // [BL][2] group sums of the current sample per lane
  const float ga = gamma[c], be = beta[c];
  const float inv = 1.f / ((float)gs * (float)HW);
  float a_dg = 0.f, a_db = 0.f, a_dbias = 0.f;
  // gridDim.y slices of the batch: at most 4 sample-lane iterations per CTA
  const int per = (B + gridDim.y - 1) / gridDim.y;
  const int bbeg = blockIdx.y * per, bend = min(bbeg + per, B);
  for (int b0 = bbeg; b0 < bend; b0 += BL) {
    const int b = b0 + bl < bend ? b0 + bl : B;     // B = "no sample"
    float S1 = 0.f, S2 = 0.f, S0 = 0.f, sc = 1.f;
    if (b < B) {
      for (int k = 0; k < chunks; ++k) {
        const float* p = sums + (((int64_t)b * chunks + k) * C + c) * 3;
        S1 += p[0]; S2 += p[1]; S0 += p[2];
      }
      sc = film ? film[(int64_t)b * film_ld + c] + 1.f : 1.f;
      if (dfilm) {
        dfilm[(int64_t)b * film_ld + c] = ga * S2 + be * S1;  // d scale
        dfilm[(int64_t)b * film_ld + C + c] = S1;             // d shift
      }
    }
    const float a = sc * ga;
    // group sums over the gs channels of this sample lane
    __syncthreads();
    if (cl == 0) { gsum[bl * 2] = 0.f; gsum[bl * 2 + 1] = 0.f; }
    __syncthreads();
    atomicAdd(&gsum[bl * 2], a * S1);
    atomicAdd(&gsum[bl * 2 + 1], a * S2);
    __syncthreads();
    if (b < B) {
      const float M1 = gsum[bl * 2] * inv, M2 = gsum[bl * 2 + 1] * inv;
      if (cl == 0) {
        gmeans[(b * G + g) * 2] = M1;
        gmeans[(b * G + g) * 2 + 1] = M2;
      }
      a_dg += sc * S2;
      a_db += sc * S1;
      // sum_p dx[b,p,c] = rstd*(a*S1 - HW*M1 - M2*sum_p xn),  sum_p xn = (S0 - HW*mean)*rstd
      const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
      const float sum_xn = (S0 - (float)HW * mean) * rstd;
      a_dbias += rstd * (a * S1 - (float)HW * M1 - M2 * sum_xn);
    }
  }
  __syncthreads();
  accg[(bl * gs + cl) * 3] = a_dg;
  accg[(bl * gs + cl) * 3 + 1] = a_db;
  accg[(bl * gs + cl) * 3 + 2] = a_dbias;
  __syncthreads();
  if (bl == 0) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < BL; ++k) {
      t0 += accg[(k * gs + cl) * 3];
      t1 += accg[(k * gs + cl) * 3 + 1];
      t2 += accg[(k * gs + cl) * 3 + 2];
    }
    if (gridDim.y == 1) {
      dgamma[c] += t0;
      dbeta[c] += t1;
      if (dbias) dbias[c] += t2;
    } else {
      atomicAdd(dgamma + c, t0);
      atomicAdd(dbeta + c, t1);
      if (dbias) atomicAdd(dbias + c, t2);
    }
  }
}

// ---- backward pass 3: dx = rstd*(a*dz - M1 - xnorm*M2) ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ x, int x_ld,
                    const float* __restrict__ stats, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ film, int film_ld,
                    const float* __restrict__ gmeans, T* __restrict__ dx, int dx_ld, int64_t total8,
                    int HW, int C, int G) {
  pdl_prologue();
  const int C8 = C / 8, gs = C / G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c0 = (int)(i % C8) * 8;
    int64_t p = i / C8;
    int b = (int)(p / HW);
    int g = c0 / gs;
    float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
    float M1 = gmeans[(b * G + g) * 2], M2 = gmeans[(b * G + g) * 2 + 1];
    float xv[8], gv[8];
    ld8(x + p * x_ld + c0, xv);
    ld8(dy + p * dy_ld + c0, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ga = gamma[c0 + j], be = beta[c0 + j];
      float sc = film ? film[(int64_t)b * film_ld + c0 + j] + 1.f : 1.f;
      float sh = film ? film[(int64_t)b * film_ld + C + c0 + j] : 0.f;
      float xn = (xv[j] - mean) * rstd;
      float z = (xn * ga + be) * sc + sh;
      float dz = gv[j] * silu_grad_f(z);
      xv[j] = rstd * (sc * ga * dz - M1 - xn * M2);
    }
    st8(dx + p * dx_ld + c0, xv);
  }
}

// =====================================================================================================
// Cluster-fused GroupNorm: ONE launch per norm in each direction.  A thread-block cluster of CL CTAs owns
// one sample; every CTA streams its pixel chunk once from HBM for the reductions, the partial sums are
// exchanged through distributed shared memory (cluster.sync), and the chunk is streamed a second time
// (L2/L1 hits: it was read microseconds earlier by the same SM) to produce the output.  Replaces the
// stats -> apply (forward) and reduce -> params -> apply (backward) launch chains.
//   thread = (8-channel vector cv = tid % C8, pixel lane = tid / C8); all per-channel coefficients are folded
//   into two registers per channel before the pixel loop, so the loop body has no divisions or lookups.
// =====================================================================================================
constexpr int GNC_THREADS = 256;
constexpr int GNC_MAXC = 512;          // channels handled by the cluster kernels (C8 <= 64 -> >= 4 pixel lanes)

__device__ __forceinline__ float fast_sigmoid(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
// bf16 path: sigmoid(z) = 0.5*tanh(0.5 z) + 0.5 with the single-instruction tanh.approx (abs. error ~5e-4, below
// bf16 resolution) — ONE SFU op per element instead of two (ex2 + rcp); the SFU (16 ops/clk/SM) was a co-limiter
// of the GroupNorm kernels, which evaluate SiLU (forward) / SiLU' (twice, backward) for every element.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T> __device__ __forceinline__ float sigmoid_t(float z);
template <> __device__ __forceinline__ float sigmoid_t<float>(float z) { return fast_sigmoid(z); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float z) {
  return fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);
}

// raw 8-element vectors: issue the loads of several pixels first, convert afterwards (memory-level parallelism)
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
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

// STAGE: the chunk's raw vectors are parked in (dynamic) shared memory during phase 1 and phase 2 reads them
// from there instead of going back to L2 (chunk bytes = ceil(HW/CL) * C * sizeof(T), <= GNC_STAGE_MAX).
constexpr int GNC_STAGE_MAX = 40 * 1024;
template <typename T, bool STAGE>
__global__ void __launch_bounds__(GNC_THREADS, sizeof(T) == 2 ? 4 : 2)
gn_fwd_cluster_kernel(const T* __restrict__ x, int x_ld, float* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ film, int film_ld, const T* __restrict__ res, int res_ld,
                      T* __restrict__ y, int y_ld, int HW, int C, int G, float eps, int write_stats,
                      const float* __restrict__ part, int slots) {
  pdl_prologue();
  cg::cluster_group cluster = cg::this_cluster();
  // part != nullptr: (sum, sumsq) partials [slots per sample][G][2] were produced by the conv epilogue; the kernel
  // is then a plain grid of pixel chunks (launched without a cluster): one pass, no barriers between CTAs
  const int CL = part ? (int)gridDim.x : (int)cluster.num_blocks();
  const int rank = part ? (int)blockIdx.x : (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int C8 = C >> 3, gs8 = C8 / G;                 // 8-wide vectors per group
  const int NT = (int)blockDim.x;                      // 256, or 128 in the pre-statistics mode (finer grid)
  const int lanes = NT / C8;
  const int cv = threadIdx.x % C8, lane = threadIdx.x / C8;
  const int c0 = cv * 8, g = cv / gs8;
  const int ppb = (HW + CL - 1) / CL;
  const int p0 = rank * ppb, p1 = min(p0 + ppb, HW);
  __shared__ float tpart[GNC_THREADS][2];
  float* part_s = &tpart[0][0];                        // reused as [slot lanes][2G] in the pre-statistics mode
  __shared__ float gpart[64][2];                       // this CTA's per-group (sum, sumsq): read by the peers
  __shared__ float gstat[64][2];                       // (mean, rstd) per group
  const T* xp = x + (int64_t)b * HW * x_ld + c0;
  extern __shared__ __align__(16) unsigned char gn_stage_raw[];
  Raw8<T>* stage = reinterpret_cast<Raw8<T>*>(gn_stage_raw);     // [pixel of the chunk][C8]
  const T* rp = res ? res + (int64_t)b * HW * res_ld + c0 : nullptr;
  T* yp = y + (int64_t)b * HW * y_ld + c0;
  // ---- phase 1: per-group sum / sum of squares of the chunk
  float s = 0.f, ss = 0.f;
  if (part) {
    // partials: [sample][slot][C/8 chunks][2].  thread = (slot lane, (group, which)): add the sample's slots and the
    // chunks of the group, then the slot lanes through shared memory
    const int gk = threadIdx.x % (2 * G), sl = threadIdx.x / (2 * G), SL = NT / (2 * G);
    const int gq = gk >> 1, which = gk & 1, cpg = C8 / G;   // chunks per group
    float a = 0.f;
    const float* pp = part + (int64_t)b * slots * 2 * C8 + (gq * cpg) * 2 + which;
    for (int i = sl; i < slots; i += SL)
      for (int c = 0; c < cpg; ++c) a += pp[((int64_t)i * C8 + c) * 2];
    part_s[sl * 2 * G + gk] = a;
    __syncthreads();
    if (threadIdx.x < G) {
      float sum = 0.f, sq = 0.f;
      for (int i = 0; i < SL; ++i) {
        sum += part_s[i * 2 * G + 2 * threadIdx.x];
        sq += part_s[i * 2 * G + 2 * threadIdx.x + 1];
      }
      const float inv = 1.f / ((float)HW * (float)(C / G));
      const float mean = sum * inv;
      const float var = fmaxf(sq * inv - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      gstat[threadIdx.x][0] = mean;
      gstat[threadIdx.x][1] = rstd;
      if (rank == 0 && write_stats) {
        stats[(b * G + threadIdx.x) * 2] = mean;
        stats[(b * G + threadIdx.x) * 2 + 1] = rstd;
      }
    }
    __syncthreads();
  } else {
  for (int p = p0 + lane; p < p1; p += 4 * lanes) {
    Raw8<T> raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * lanes < p1) raw[u].load(xp + (int64_t)(p + u * lanes) * x_ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (p + u * lanes < p1) {
        float v[8];
        raw[u].unpack(v);
        if (STAGE) stage[(p + u * lanes - p0) * C8 + cv] = raw[u];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s += v[j];
          ss = fmaf(v[j], v[j], ss);
        }
      }
    }
  }
  tpart[threadIdx.x][0] = s;
  tpart[threadIdx.x][1] = ss;
  __syncthreads();
  if (threadIdx.x < 2 * G) {
    const int gg = threadIdx.x >> 1, k = threadIdx.x & 1;
    float a = 0.f;
    for (int l = 0; l < lanes; ++l)
      for (int v = 0; v < gs8; ++v) a += tpart[l * C8 + gg * gs8 + v][k];
    gpart[gg][k] = a;
  }
  cluster.sync();
  if (threadIdx.x < G) {
    float a = 0.f, q = 0.f;
    for (int r = 0; r < CL; ++r) {
      const float* rp = cluster.map_shared_rank(&gpart[0][0], r);
      a += rp[threadIdx.x * 2];
      q += rp[threadIdx.x * 2 + 1];
    }
    const float inv = 1.f / ((float)HW * (float)(C / G));
    const float mean = a * inv;
    const float var = fmaxf(q * inv - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    gstat[threadIdx.x][0] = mean;
    gstat[threadIdx.x][1] = rstd;
    if (rank == 0 && write_stats) {
      stats[(b * G + threadIdx.x) * 2] = mean;
      stats[(b * G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  }   // statistics from this kernel's own reduction pass
  // ---- phase 2: y = silu(A*x + Bc) (+ res)
  float A[8], Bc[8];
  {
    const float mean = gstat[g][0], rstd = gstat[g][1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ga = gamma[c0 + j] * rstd;
      float be = beta[c0 + j] - mean * ga;
      if (film) {
        const float sc = film[(int64_t)b * film_ld + c0 + j] + 1.f;
        const float sh = film[(int64_t)b * film_ld + C + c0 + j];
        ga *= sc;
        be = be * sc + sh;
      }
      // bf16 mode: silu(z) = h + h * tanh(h) with h = z / 2, so the halves are folded into the coefficients
      A[j] = sizeof(T) == 2 ? 0.5f * ga : ga;
      Bc[j] = sizeof(T) == 2 ? 0.5f * be : be;
    }
  }
  for (int p = p0 + lane; p < p1; p += 2 * lanes) {
    Raw8<T> rx[2], rr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p + u * lanes < p1) {
        if (STAGE) rx[u] = stage[(p + u * lanes - p0) * C8 + cv];
        else rx[u].load(xp + (int64_t)(p + u * lanes) * x_ld);
        if (rp) rr[u].load(rp + (int64_t)(p + u * lanes) * res_ld);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p + u * lanes < p1) {
        float v[8], r[8];
        rx[u].unpack(v);
        if (rp) rr[u].unpack(r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(A[j], v[j], Bc[j]);
          float o;
          if (sizeof(T) == 2) o = fmaf(z, tanh_approx(z), z);
          else o = z * sigmoid_t<T>(z);
          v[j] = rp ? o + r[j] : o;
        }
        st8(yp + (int64_t)(p + u * lanes) * y_ld, v);
      }
    }
  }
  if (!part) cluster.sync();        // peers may still be reading this CTA's gpart
}

// UB pixels per thread and iteration; PIPE: the next UB are requested before the current UB are processed
// (software pipeline, twice the raw registers); MINB: resident CTAs per SM the register budget is set for.
// DZ (bf16 only): phase 1 parks dz = dy * silu'(z), rounded to bf16, in dynamic shared memory — every thread its own
// vectors, [iteration][thread] x 16 B, no synchronisation — and phase 2 only reads x again: dx = P*dz + Q + R*x.
// ncu at the training shape showed the kernel instruction-bound (61 % issue slots, DRAM 17 % of peak): the second
// evaluation of the sigmoid and its derivative was a third of the instructions.
template <typename T, int UB, bool PIPE, int MINB, bool DZ>
__global__ void __launch_bounds__(GNC_THREADS, MINB)
gn_bwd_cluster_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ x, int x_ld,
                      const float* __restrict__ stats, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ film, int film_ld,
                      T* __restrict__ dx, int dx_ld, float* __restrict__ dgamma, float* __restrict__ dbeta,
                      float* __restrict__ dfilm, float* __restrict__ dbias, int HW, int C, int G) {
  pdl_prologue();
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int C8 = C >> 3, gs8 = C8 / G, gs = C / G;
  const int lanes = GNC_THREADS / C8;
  const int cv = threadIdx.x % C8, lane = threadIdx.x / C8;
  const int c0 = cv * 8, g = cv / gs8;
  const int ppb = (HW + CL - 1) / CL;
  const int p0 = rank * ppb, p1 = min(p0 + ppb, HW);
  __shared__ float red[GNC_THREADS * 8 * 3];           // [lanes][C][3], 24 KiB
  __shared__ float cpart[GNC_MAXC * 3];                // this CTA's per-channel (S1, S2, S0): read by the peers
  __shared__ float ctot[GNC_MAXC * 3];                 // cluster totals
  __shared__ float gm[64][2];                          // (M1, M2) per group
  __shared__ float coef_s[GNC_MAXC * 3];               // per channel: gamma, beta, FiLM scale + 1
  extern __shared__ uint4 dz_park[];                   // DZ: [iteration * UB + u][thread]
  const T* xp = x + (int64_t)b * HW * x_ld + c0;
  const T* gp = dy + (int64_t)b * HW * dy_ld + c0;
  // the first pixels of the chunk are requested before anything else: their latency overlaps the coefficient setup
  Raw8<T> rx[UB], rg[UB];
  auto fetch = [&](int p) {
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (p + u * lanes < p1) {
        rx[u].load(xp + (int64_t)(p + u * lanes) * x_ld);
        rg[u].load(gp + (int64_t)(p + u * lanes) * dy_ld);
      }
    }
  };
  fetch(p0 + lane);
  const float mean = __ldg(stats + (b * G + g) * 2), rstd = __ldg(stats + (b * G + g) * 2 + 1);
  float A[8], Bc[8], sc[8];
  {
    float gv[8], bv[8], sv[8], hv[8];
    ld8(gamma + c0, gv);
    ld8(beta + c0, bv);
    if (film) {      // FiLM rows may start at any float offset (the C ABI only asks for fp32 alignment)
      const float* fs = film + (int64_t)b * film_ld + c0;
      if ((((uintptr_t)fs | (uintptr_t)(fs + C)) & 15) == 0) {
        ld8(fs, sv);
        ld8(fs + C, hv);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sv[j] = fs[j];
          hv[j] = fs[C + j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ga = gv[j] * rstd;
      float be = bv[j] - mean * ga;
      sc[j] = 1.f;
      if (film) {
        sc[j] = sv[j] + 1.f;
        ga *= sc[j];
        be = be * sc[j] + hv[j];
      }
      A[j] = ga;
      Bc[j] = be;
      if (lane == 0) {      // one copy of the per-channel coefficients for the (few-thread) reductions below
        coef_s[(c0 + j) * 3] = gv[j];
        coef_s[(c0 + j) * 3 + 1] = bv[j];
        coef_s[(c0 + j) * 3 + 2] = sc[j];
      }
    }
  }
  if (threadIdx.x < 2 * G) (&gm[0][0])[threadIdx.x] = 0.f;
  // ---- phase 1: S1 = sum dz, S2 = sum dz*xn, S0 = sum x   (per channel, over the chunk); the loads of the
  //      next two pixels are in flight while the current two are processed
  float s1[8] = {}, s2[8] = {}, s0[8] = {};
  int slot = 0;
  for (int p = p0 + lane; p < p1; p += UB * lanes, slot += UB) {
    Raw8<T> cx[UB], cg2[UB];
    if (!PIPE && p != p0 + lane) fetch(p);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      cx[u] = rx[u];
      cg2[u] = rg[u];
    }
    if (PIPE && p + UB * lanes < p1) fetch(p + UB * lanes);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (p + u * lanes < p1) {
        float xv[8], gv[8];
        cx[u].unpack(xv);
        cg2[u].unpack(gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(A[j], xv[j], Bc[j]);
          const float sg = sigmoid_t<T>(z);
          const float dz = gv[j] * sg * (1.f + z * (1.f - sg));
          const float xn = (xv[j] - mean) * rstd;
          s1[j] += dz;
          s2[j] = fmaf(dz, xn, s2[j]);
          s0[j] += xv[j];
          gv[j] = dz;
        }
        if (DZ) {
          __nv_bfloat162 h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(gv[2 * j], gv[2 * j + 1]);
          dz_park[(slot + u) * GNC_THREADS + threadIdx.x] = *reinterpret_cast<const uint4*>(h);
        }
      }
    }
  }
  // phase 2 starts from the first pixels again: request them now, the reductions below hide the latency
  auto fetchx = [&](int p) {
#pragma unroll
    for (int u = 0; u < UB; ++u)
      if (p + u * lanes < p1) rx[u].load(xp + (int64_t)(p + u * lanes) * x_ld);
  };
  if (DZ) fetchx(p0 + lane);
  else fetch(p0 + lane);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float* r = red + ((lane * C + c0 + j) * 3);
    r[0] = s1[j];
    r[1] = s2[j];
    r[2] = s0[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 3; i += GNC_THREADS) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += red[l * C * 3 + i];
    cpart[i] = a;
  }
  cluster.sync();
  for (int i = threadIdx.x; i < C * 3; i += GNC_THREADS) {
    float a = 0.f;
    for (int r = 0; r < CL; ++r) a += cluster.map_shared_rank(&cpart[0], r)[i];
    ctot[i] = a;
  }
  __syncthreads();
  // group means of a*S1 and a*S2 (a = scale * gamma): every thread takes channels, shared-memory atomics per group
  {
    const float inv = 1.f / ((float)gs * (float)HW);
    const int seg = gs < 32 ? gs : 32;                 // lanes of a warp that belong to the same group
    for (int c = threadIdx.x; c < C; c += GNC_THREADS) {   // C is a multiple of 32: whole warps iterate together
      const float a = coef_s[c * 3 + 2] * coef_s[c * 3] * inv;
      float v1 = a * ctot[c * 3], v2 = a * ctot[c * 3 + 1];
      for (int o = 1; o < seg; o <<= 1) {
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      }
      if ((c & (seg - 1)) == 0) {                      // at most gs/32 adders per group
        atomicAdd(&gm[c / gs][0], v1);
        atomicAdd(&gm[c / gs][1], v2);
      }
    }
  }
  __syncthreads();
  // parameter gradients (one CTA of the cluster): batch reduction through fp32 reductions in L2
  if (rank == 0 && dgamma != nullptr) {
    for (int c = threadIdx.x; c < C; c += GNC_THREADS) {
      const int gg = c / gs;
      const float S1 = ctot[c * 3], S2 = ctot[c * 3 + 1], S0 = ctot[c * 3 + 2];
      const float ga = coef_s[c * 3], be = coef_s[c * 3 + 1], scl = coef_s[c * 3 + 2];
      if (dfilm) {
        dfilm[(int64_t)b * film_ld + c] = ga * S2 + be * S1;
        dfilm[(int64_t)b * film_ld + C + c] = S1;
      }
      atomicAdd(dgamma + c, scl * S2);
      atomicAdd(dbeta + c, scl * S1);
      if (dbias) {
        const float mu = stats[(b * G + gg) * 2], rs = stats[(b * G + gg) * 2 + 1];
        const float sum_xn = (S0 - (float)HW * mu) * rs;
        atomicAdd(dbias + c, rs * (scl * ga * S1 - (float)HW * gm[gg][0] - gm[gg][1] * sum_xn));
      }
    }
  }
  // ---- phase 2: dx = P*dz + Q + R*x,  P = rstd*scale*gamma, R = -rstd^2*M2, Q = -rstd*M1 - R*mean
  const float M1 = gm[g][0], M2 = gm[g][1];
  const float R = -rstd * rstd * M2, Q = -rstd * M1 - R * mean;
  float P[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) P[j] = rstd * sc[j] * gamma[c0 + j];
  T* dp = dx + (int64_t)b * HW * dx_ld + c0;
  slot = 0;
  for (int p = p0 + lane; p < p1; p += UB * lanes, slot += UB) {
    Raw8<T> cx[UB], cg2[UB];
    if (DZ) {
      if (!PIPE && p != p0 + lane) fetchx(p);
#pragma unroll
      for (int u = 0; u < UB; ++u) cx[u] = rx[u];
      if (PIPE && p + UB * lanes < p1) fetchx(p + UB * lanes);
    } else {
      if (!PIPE && p != p0 + lane) fetch(p);
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        cx[u] = rx[u];
        cg2[u] = rg[u];
      }
      if (PIPE && p + UB * lanes < p1) fetch(p + UB * lanes);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (p + u * lanes < p1) {
        float xv[8], gv[8];
        cx[u].unpack(xv);
        if (DZ) {
          const uint4 d4 = dz_park[(slot + u) * GNC_THREADS + threadIdx.x];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&d4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h[j]);
            xv[2 * j] = fmaf(P[2 * j], f.x, fmaf(R, xv[2 * j], Q));
            xv[2 * j + 1] = fmaf(P[2 * j + 1], f.y, fmaf(R, xv[2 * j + 1], Q));
          }
        } else {
          cg2[u].unpack(gv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float z = fmaf(A[j], xv[j], Bc[j]);
            const float sg = sigmoid_t<T>(z);
            const float dz = gv[j] * sg * (1.f + z * (1.f - sg));
            xv[j] = fmaf(P[j], dz, fmaf(R, xv[j], Q));
          }
        }
        st8(dp + (int64_t)(p + u * lanes) * dx_ld, xv);
      }
    }
  }
  cluster.sync();
}

// cluster size: as many CTAs per sample as still fit on the machine in ONE wave (a second, partial wave costs
// more than the smaller chunks save), each chunk at least two passes of the pixel lanes; any size 1..8.
static int gn_cluster_size(int B, int HW, int C, int ctas_per_sm) {
  const int lanes = GNC_THREADS / (C / 8);
  const long long slots = (long long)num_sms() * ctas_per_sm;
  int cl = (int)(slots / B);
  if (cl > 8) cl = 8;
  while (cl > 1 && HW / cl < 2 * lanes) --cl;
  return cl < 1 ? 1 : cl;
}
static bool gn_cluster_ok(int C, int G) {
  const int C8 = C / 8;
  return C <= GNC_MAXC && GNC_THREADS % C8 == 0 && C8 % G == 0 && G <= 64;
}

template <typename K, typename... Args>
static cudaError_t launch_cluster(K kernel, dim3 grid, int cl, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GNC_THREADS);
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

// ---- RMSNorm: a row (pixel) is handled by L = min(32, C/8) lanes, 32/L rows per warp; C <= 512 ----------
// These kernels are issue-bound on B200 (ncu: 69 % issue slots at 45 % occupancy, 78 instructions per 16-byte
// vector before this version), so they are written for instruction count: L is a template constant (LT; 0 = the
// run-time value, for channel counts the UNet does not use), rows are 32-bit, the row pointers walk, and
// 1 / max(||x||, 1e-12) is one rsqrt.approx of max(ss, 1e-24).

template <int LT>
__device__ __forceinline__ float seg_sum(float v, int L) {   // sum over aligned groups of L lanes
  if (LT) {
#pragma unroll
    for (int o = LT >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  } else {
    for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, int MAXV, int LT>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ g,
                   const T* __restrict__ res, int res_ld, T* __restrict__ y, int y_ld, int rows,
                   int C, int Lrt) {
  pdl_prologue();
  const int L = LT ? LT : Lrt;
  const int lane = threadIdx.x & 31;
  const int sub = lane / L, sl = lane % L, rpw = 32 / L;      // row slot inside the warp, lane in the row
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int C8 = C / 8;
  const float sqrtC = sqrtf((float)C);
  // this thread's gains (its channel vectors are the same for every row it visits)
  float gs[MAXV][8];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int cv = sl + L * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) gs[k][j] = cv < C8 ? g[cv * 8 + j] * sqrtC : 0.f;
  }
  constexpr int U = 2;                                         // rows in flight per thread
  const int stride = nwarps * rpw;                             // rows between two visits of a row slot
  const T* xp = x + (int64_t)(warp * rpw + sub) * x_ld + sl * 8;
  const T* rp = res ? res + (int64_t)(warp * rpw + sub) * res_ld + sl * 8 : nullptr;
  T* yp = y + (int64_t)(warp * rpw + sub) * y_ld + sl * 8;
  const int64_t xs = (int64_t)stride * x_ld, rs = (int64_t)stride * res_ld, ys = (int64_t)stride * y_ld;
  for (int r0 = warp * rpw; r0 < rows; r0 += stride * U, xp += U * xs, rp += U * rs, yp += U * ys) {
    Raw8<T> rx[U][MAXV], rr[U][MAXV];
    // all loads of the U rows (input and residual) are issued before any arithmetic
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = r0 + u * stride + sub < rows;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (ok && sl + L * k < C8) {
          rx[u][k].load(xp + u * xs + k * L * 8);
          if (res) rr[u][k].load(rp + u * rs + k * L * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = r0 + u * stride + sub < rows;
      float v[MAXV][8];
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (ok && sl + L * k < C8) {
          rx[u][k].unpack(v[k]);
#pragma unroll
          for (int j = 0; j < 8; ++j) ss = fmaf(v[k][j], v[k][j], ss);
        }
      }
      ss = seg_sum<LT>(ss, L);
      const float rn = rsqrt_ftz(fmaxf(ss, 1e-24f));            // 1 / max(||x||, 1e-12)
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (ok && sl + L * k < C8) {
          float o[8], rv[8];
          if (res) rr[u][k].unpack(rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = v[k][j] * rn;
            o[j] = res ? fmaf(t, gs[k][j], rv[j]) : t * gs[k][j];
          }
          st8(yp + u * ys + k * L * 8, o);
        }
      }
    }
  }
}

template <typename T, int MAXV, int LT>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const T* __restrict__ dy, int dy_ld, const T* __restrict__ x, int x_ld,
                   const float* __restrict__ g, const T* __restrict__ res, int res_ld,
                   T* __restrict__ dx, int dx_ld, float* __restrict__ dg, int rows, int C, int Lrt) {
  pdl_prologue();
  const int L = LT ? LT : Lrt;
  const int lane = threadIdx.x & 31;
  const int sub = lane / L, sl = lane % L, rpw = 32 / L;
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int C8 = C / 8;
  const float sqrtC = sqrtf((float)C);
  float gs[MAXV][8], dgacc[MAXV][8];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int cv = sl + L * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gs[k][j] = cv < C8 ? g[cv * 8 + j] : 0.f;
      dgacc[k][j] = 0.f;
    }
  }
  const int stride = nwarps * rpw;
  const int64_t row0 = warp * rpw + sub;
  const T* xp = x + row0 * x_ld + sl * 8;
  const T* gp = dy + row0 * dy_ld + sl * 8;
  const T* rp = res ? res + row0 * res_ld + sl * 8 : nullptr;
  T* op = dx + row0 * dx_ld + sl * 8;
  const int64_t xs = (int64_t)stride * x_ld, gsd = (int64_t)stride * dy_ld, rs = (int64_t)stride * res_ld,
                os = (int64_t)stride * dx_ld;
  // the kernel runs as ~2 CTAs per SM (every CTA ends with one dg atomic per channel), so it is latency-bound:
  // U rows per thread are requested before the first reduction
  constexpr int U = MAXV == 1 ? 2 : 1;
  for (int r0 = warp * rpw; r0 < rows; r0 += U * stride, xp += U * xs, gp += U * gsd, rp += U * rs, op += U * os) {
    Raw8<T> rx[U][MAXV], rg[U][MAXV], rr[U][MAXV];
#pragma unroll
    for (int w = 0; w < U; ++w) {
      const bool ok = r0 + w * stride + sub < rows;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (ok && sl + L * k < C8) {
          rx[w][k].load(xp + w * xs + k * L * 8);
          rg[w][k].load(gp + w * gsd + k * L * 8);
          if (res) rr[w][k].load(rp + w * rs + k * L * 8);
        }
      }
    }
#pragma unroll
    for (int w = 0; w < U; ++w) {
      const bool ok = r0 + w * stride + sub < rows;
      float u[MAXV][8], gd[MAXV][8];
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (ok && sl + L * k < C8) {
          rx[w][k].unpack(u[k]);
          rg[w][k].unpack(gd[k]);
#pragma unroll
          for (int j = 0; j < 8; ++j) ss = fmaf(u[k][j], u[k][j], ss);
        }
      }
      ss = seg_sum<LT>(ss, L);
      const float rn = rsqrt_ftz(fmaxf(ss, 1e-24f));
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (ok && sl + L * k < C8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            u[k][j] *= rn;                                       // unit vector
            dgacc[k][j] = fmaf(gd[k][j], u[k][j], dgacc[k][j]);  // dg_c += dy_c*u_c (x sqrtC at the end)
            gd[k][j] *= gs[k][j];                                // g .* dy
            dot = fmaf(gd[k][j], u[k][j], dot);
          }
        }
      }
      dot = seg_sum<LT>(dot, L);
      const float a = rn * sqrtC, nb = -a * dot;                 // dx = a * gd + nb * u (+ skip gradient)
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        if (ok && sl + L * k < C8) {
          float o[8], rv[8];
          if (res) rr[w][k].unpack(rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = fmaf(nb, u[k][j], a * gd[k][j]);
            o[j] = res ? t + rv[j] : t;
          }
          st8(op + w * os + k * L * 8, o);
        }
      }
    }
  }
  // dg: reduce all (warp, row-slot) partials of the CTA in shared memory, one atomic per channel per CTA
  __shared__ float dgs[512];
  for (int c = threadIdx.x; c < C; c += blockDim.x) dgs[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int cv = sl + L * k;
    if (cv < C8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&dgs[cv * 8 + j], dgacc[k][j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dg + c, dgs[c] * sqrtC);
}

static inline unsigned ew_grid(int64_t n) {
  int64_t blocks = (n + 255) / 256, cap = (int64_t)num_sms() * 16;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace b200dm

using namespace b200dm;
typedef __nv_bfloat16 bf16;

#define GN_CHECKS(name)                                                                            \
  B200DM_REQUIRE(B > 0 && HW > 0 && C > 0 && G > 0 && G <= 64 && C % G == 0 && (C / G) % 8 == 0,   \
                 B200DM_ERR_SHAPE, name ": need C %% G == 0 and (C/G) %% 8 == 0 (C=%d G=%d)", C, G)

extern "C" int b200dm_gn_stats(int32_t dtype, const void* x, int32_t x_ld, float* stats, int32_t B,
                               int32_t HW, int32_t C, int32_t G, float eps, void* stream) {
  GN_CHECKS("gn_stats");
  B200DM_REQUIRE(x_ld % 8 == 0, B200DM_ERR_SHAPE, "gn_stats: ld must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(gn_stats_kernel<float>, B * G, 256, 0, st, (const float*)x, x_ld, stats, HW, C, G, eps);
  else
    launch_k(gn_stats_kernel<bf16>, B * G, 256, 0, st, (const bf16*)x, x_ld, stats, HW, C, G, eps);
  count_launch();
  return check_launch("gn_stats");
}

extern "C" int b200dm_gn_apply_fwd(int32_t dtype, const void* x, int32_t x_ld, const float* stats,
                                   const float* gamma, const float* beta, const float* film,
                                   int32_t film_ld, const void* res, int32_t res_ld, void* y,
                                   int32_t y_ld, int32_t B, int32_t HW, int32_t C, int32_t G,
                                   void* stream) {
  GN_CHECKS("gn_apply_fwd");
  B200DM_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && (!res || res_ld % 8 == 0), B200DM_ERR_SHAPE,
                 "gn_apply_fwd: ld must be a multiple of 8");
  int64_t total8 = (int64_t)B * HW * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(gn_apply_fwd_kernel<float>, ew_grid(total8), 256, 0, st, 
        (const float*)x, x_ld, stats, gamma, beta, film, film_ld, (const float*)res, res_ld, (float*)y,
        y_ld, total8, HW, C, G);
  else
    launch_k(gn_apply_fwd_kernel<bf16>, ew_grid(total8), 256, 0, st, 
        (const bf16*)x, x_ld, stats, gamma, beta, film, film_ld, (const bf16*)res, res_ld, (bf16*)y,
        y_ld, total8, HW, C, G);
  count_launch();
  return check_launch("gn_apply_fwd");
}

extern "C" int b200dm_gn_fwd(int32_t dtype, const void* x, int32_t x_ld, float* stats, const float* gamma,
                             const float* beta, const float* film, int32_t film_ld, const void* res,
                             int32_t res_ld, void* y, int32_t y_ld, int32_t B, int32_t HW, int32_t C,
                             int32_t G, float eps, void* stream) {
  GN_CHECKS("gn_fwd");
  B200DM_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && (!res || res_ld % 8 == 0), B200DM_ERR_SHAPE,
                 "gn_fwd: ld must be a multiple of 8");
  B200DM_REQUIRE(stats != nullptr, B200DM_ERR_SHAPE, "gn_fwd: stats output is required");
  cudaStream_t st = (cudaStream_t)stream;
  if (!gn_cluster_ok(C, G)) {      // shapes outside the cluster kernel: the two-launch path
    int rc = b200dm_gn_stats(dtype, x, x_ld, stats, B, HW, C, G, eps, stream);
    if (rc) return rc;
    return b200dm_gn_apply_fwd(dtype, x, x_ld, stats, gamma, beta, film, film_ld, res, res_ld, y, y_ld, B,
                               HW, C, G, stream);
  }
  const int cl = gn_cluster_size(B, HW, C, 4);
  dim3 grid(cl, B);
  const size_t chunk_bytes = (size_t)((HW + cl - 1) / cl) * C * (dtype == B200DM_F32 ? 4 : 2);
  const bool stage = chunk_bytes <= (size_t)GNC_STAGE_MAX;
  const size_t smem = stage ? chunk_bytes : 0;
  cudaError_t e;
#define GN_FWD_LAUNCH(TT, ST)                                                                                 \
  launch_cluster(gn_fwd_cluster_kernel<TT, ST>, grid, cl, smem, st, (const TT*)x, (int)x_ld, stats, gamma, beta, \
                 film, (int)film_ld, (const TT*)res, (int)res_ld, (TT*)y, (int)y_ld, (int)HW, (int)C, (int)G, eps, 1,   \
                 (const float*)nullptr, 0)
  if (dtype == B200DM_F32)
    e = stage ? GN_FWD_LAUNCH(float, true) : GN_FWD_LAUNCH(float, false);
  else
    e = stage ? GN_FWD_LAUNCH(bf16, true) : GN_FWD_LAUNCH(bf16, false);
#undef GN_FWD_LAUNCH
  B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "gn_fwd: launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("gn_fwd");
}

extern "C" int b200dm_gn_fwd_pre(int32_t dtype, const void* x, int32_t x_ld, const float* part, int32_t slots,
                                 float* stats, const float* gamma, const float* beta, const float* film,
                                 int32_t film_ld, const void* res, int32_t res_ld, void* y, int32_t y_ld, int32_t B,
                                 int32_t HW, int32_t C, int32_t G, float eps, void* stream) {
  GN_CHECKS("gn_fwd_pre");
  B200DM_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && (!res || res_ld % 8 == 0), B200DM_ERR_SHAPE,
                 "gn_fwd_pre: ld must be a multiple of 8");
  B200DM_REQUIRE(part != nullptr && slots > 0 && stats != nullptr, B200DM_ERR_SHAPE, "gn_fwd_pre: partials and stats required");
  B200DM_REQUIRE(gn_cluster_ok(C, G) && GNC_THREADS % (2 * G) == 0, B200DM_ERR_UNSUPPORTED,
                 "gn_fwd_pre: C=%d G=%d not supported", C, G);
  cudaStream_t st = (cudaStream_t)stream;
  // plain grid of pixel chunks: about one wave of CTAs, every chunk at least two passes of the pixel lanes
  // (measured: 256-thread CTAs, 4 per SM; finer grids of 128-thread CTAs were within 1 %, grids x2 / x4 slower)
  const int threads = GNC_THREADS;
  const int lanes = threads / (C / 8);
  int chunks = (int)(((long long)num_sms() * 4) / B);
  if (chunks > 32) chunks = 32;
  while (chunks > 1 && HW / chunks < 2 * lanes) --chunks;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, B);
  if (dtype == B200DM_F32)
    launch_k(gn_fwd_cluster_kernel<float, false>, grid, threads, 0, st, (const float*)x, (int)x_ld, stats, gamma,
             beta, film, (int)film_ld, (const float*)res, (int)res_ld, (float*)y, (int)y_ld, (int)HW, (int)C, (int)G,
             eps, 1, part, (int)slots);
  else
    launch_k(gn_fwd_cluster_kernel<bf16, false>, grid, threads, 0, st, (const bf16*)x, (int)x_ld, stats, gamma,
             beta, film, (int)film_ld, (const bf16*)res, (int)res_ld, (bf16*)y, (int)y_ld, (int)HW, (int)C, (int)G,
             eps, 1, part, (int)slots);
  count_launch();
  return check_launch("gn_fwd_pre");
}

static int gn_bwd_chunks(int B, int HW, int C) {
  const int C8 = C / 8;
  int threads = 256;
  if (C8 > threads) threads = C8;
  int lanes = threads / C8;
  int chunks = (2 * num_sms() + B - 1) / B;
  int maxchunks = (HW + lanes - 1) / lanes;
  if (chunks > maxchunks) chunks = maxchunks;
  if (chunks < 1) chunks = 1;
  int ppb = (HW + chunks - 1) / chunks;
  return (HW + ppb - 1) / ppb;
}

extern "C" int64_t b200dm_gn_bwd_ws_floats(int32_t B, int32_t HW, int32_t C) {
  if (B <= 0 || HW <= 0 || C <= 0 || C % 8) return 0;
  return (int64_t)B * gn_bwd_chunks(B, HW, C) * C * 3;
}

extern "C" int b200dm_gn_apply_bwd(int32_t dtype, const void* dy, int32_t dy_ld, const void* x,
                                   int32_t x_ld, const float* stats, const float* gamma,
                                   const float* beta, const float* film, int32_t film_ld, void* dx,
                                   int32_t dx_ld, float* dgamma, float* dbeta, float* dfilm,
                                   float* dbias, float* sums, float* gmeans, int32_t B, int32_t HW,
                                   int32_t C, int32_t G, void* stream) {
  GN_CHECKS("gn_apply_bwd");
  B200DM_REQUIRE(x_ld % 8 == 0 && dy_ld % 8 == 0 && dx_ld % 8 == 0, B200DM_ERR_SHAPE,
                 "gn_apply_bwd: ld must be a multiple of 8");
  B200DM_REQUIRE(C <= 2048, B200DM_ERR_UNSUPPORTED, "gn_apply_bwd: C=%d too large", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (gn_cluster_ok(C, G)) {
    // measured on the training step (B = 128, 32x32, 38 launches): three pixels in flight, pipelined, cluster sized for
    // two CTAs per SM: 0.69 ms; two pixels / three CTAs per SM (cluster of 3): 0.85; four pixels unpipelined: 0.71;
    // four pipelined (spills): 0.79; eight unpipelined: 0.76.
    const int minb = dtype == B200DM_F32 ? 3 : 2;
    const int cl = gn_cluster_size(B, HW, C, minb);
    dim3 grid(cl, B);
    cudaError_t e;
#define GN_BWD_LAUNCH(TT, UB, PIPE, MINB, DZ, SMEM)                                                                 \
  e = launch_cluster(gn_bwd_cluster_kernel<TT, UB, PIPE, MINB, DZ>, grid, cl, SMEM, st, (const TT*)dy, (int)dy_ld,  \
                     (const TT*)x, (int)x_ld, stats, gamma, beta, film, (int)film_ld, (TT*)dx, (int)dx_ld, dgamma, \
                     dbeta, dfilm, dbias, (int)HW, (int)C, (int)G)
    if (dtype == B200DM_F32) {
      GN_BWD_LAUNCH(float, 2, true, 3, false, 0);
    } else {
      // dz parked in shared memory when the chunk's vectors fit next to a second resident CTA: one 16-byte slot per
      // (pixel of the thread, thread); 2 x (43.5 KiB static + 68 KiB) stays under the SM's 227 KiB
      const int lanes = GNC_THREADS / (C / 8), ppb = (HW + cl - 1) / cl;
      const size_t park = (size_t)((ppb + lanes - 1) / lanes) * GNC_THREADS * 16;
      static const bool dz_on = []() { const char* v = getenv("B200DM_GNB_DZ"); return !(v && v[0] == '0'); }();
      if (dz_on && park <= 68 * 1024) {
        static bool configured = false;
        if (!configured) {
          cudaFuncSetAttribute(gn_bwd_cluster_kernel<bf16, 3, true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               68 * 1024);
          configured = true;
        }
        GN_BWD_LAUNCH(bf16, 3, true, 2, true, park);
      } else {
        GN_BWD_LAUNCH(bf16, 3, true, 2, false, 0);
      }
    }
#undef GN_BWD_LAUNCH
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "gn_apply_bwd: launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return check_launch("gn_apply_bwd");
  }
  const int C8 = C / 8;
  int threads = 256;
  if (C8 > threads) threads = C8;  // C <= 2048 -> <= 256 anyway
  int lanes = threads / C8;
  int chunks = (2 * num_sms() + B - 1) / B;
  int maxchunks = (HW + lanes - 1) / lanes;
  if (chunks > maxchunks) chunks = maxchunks;
  if (chunks < 1) chunks = 1;
  int ppb = (HW + chunks - 1) / chunks;
  chunks = (HW + ppb - 1) / ppb;
  size_t smem = (size_t)lanes * C * 3 * sizeof(float);
  dim3 grid(chunks, B);
  int64_t total8 = (int64_t)B * HW * C8;
  if (dtype == B200DM_F32) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(gn_bwd_reduce_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(gn_bwd_reduce_kernel<float>, grid, threads, smem, st, 
        (const float*)dy, dy_ld, (const float*)x, x_ld, stats, gamma, beta, film, film_ld, sums, HW, C, G, ppb);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(gn_bwd_reduce_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(gn_bwd_reduce_kernel<bf16>, grid, threads, smem, st, 
        (const bf16*)dy, dy_ld, (const bf16*)x, x_ld, stats, gamma, beta, film, film_ld, sums, HW, C, G, ppb);
  }
  {
    const int gs = C / G;
    B200DM_REQUIRE(gs <= 256 && 256 % gs == 0, B200DM_ERR_UNSUPPORTED, "gn_apply_bwd: group size %d", gs);
    const int BL = 256 / gs;
    dim3 pgrid(G, (B + 4 * BL - 1) / (4 * BL));
    launch_k(gn_bwd_params_kernel, pgrid, 256, (size_t)(BL * gs * 3 + BL * 2) * sizeof(float), st, 
        sums, stats, gamma, beta, film, film_ld, dgamma, dbeta, dfilm, dbias, gmeans, B, HW, C, G, chunks);
  }
  if (dtype == B200DM_F32)
    launch_k(gn_bwd_apply_kernel<float>, ew_grid(total8), 256, 0, st, 
        (const float*)dy, dy_ld, (const float*)x, x_ld, stats, gamma, beta, film, film_ld, gmeans,
        (float*)dx, dx_ld, total8, HW, C, G);
  else
    launch_k(gn_bwd_apply_kernel<bf16>, ew_grid(total8), 256, 0, st, 
        (const bf16*)dy, dy_ld, (const bf16*)x, x_ld, stats, gamma, beta, film, film_ld, gmeans,
        (bf16*)dx, dx_ld, total8, HW, C, G);
  count_launch(3);
  return check_launch("gn_apply_bwd");
}

// dispatch on (two vectors per lane, lanes per row): the UNet's channel counts get compile-time L
#define RMS_DISPATCH(CALL)                                   \
  if (two) { CALL(2, 32); }                                  \
  else if (L == 32 && C == 256) { CALL(1, 32); }             \
  else if (L == 16 && C == 128) { CALL(1, 16); }             \
  else if (L == 8 && C == 64) { CALL(1, 8); }                \
  else { CALL(1, 0); }

extern "C" int b200dm_rmsnorm_fwd(int32_t dtype, const void* x, int32_t x_ld, const float* g,
                                  const void* res, int32_t res_ld, void* y, int32_t y_ld, int64_t rows,
                                  int32_t C, void* stream) {
  B200DM_REQUIRE(rows > 0 && rows < (1LL << 31) && C % 8 == 0 && C <= 512 && x_ld % 8 == 0 && y_ld % 8 == 0,
                 B200DM_ERR_SHAPE, "rmsnorm_fwd: need C %% 8 == 0, C <= 512 (C=%d), rows < 2^31", C);
  cudaStream_t st = (cudaStream_t)stream;
  int L = 1;
  while (L < C / 8 && L < 32) L <<= 1;
  unsigned grid = ew_grid(rows * L);
  const bool two = C / 8 > 32;      // 512 channels: two 8-wide vectors per lane
#define RMS_FWD_T(TT, MV, LT)                                                                                      \
  launch_k(rmsnorm_fwd_kernel<TT, MV, LT>, grid, 256, 0, st, (const TT*)x, x_ld, g, (const TT*)res, res_ld, (TT*)y, \
           y_ld, (int)rows, C, L)
#define RMS_FWD_F32(MV, LT) RMS_FWD_T(float, MV, LT)
#define RMS_FWD_BF16(MV, LT) RMS_FWD_T(bf16, MV, LT)
  if (dtype == B200DM_F32) {
    RMS_DISPATCH(RMS_FWD_F32)
  } else {
    RMS_DISPATCH(RMS_FWD_BF16)
  }
#undef RMS_FWD_T
#undef RMS_FWD_F32
#undef RMS_FWD_BF16
  count_launch();
  return check_launch("rmsnorm_fwd");
}

extern "C" int b200dm_rmsnorm_bwd(int32_t dtype, const void* dy, int32_t dy_ld, const void* x,
                                  int32_t x_ld, const float* g, const void* res, int32_t res_ld,
                                  void* dx, int32_t dx_ld, float* dg, int64_t rows, int32_t C,
                                  void* stream) {
  B200DM_REQUIRE(rows > 0 && rows < (1LL << 31) && C % 8 == 0 && C <= 512 && x_ld % 8 == 0 && dy_ld % 8 == 0 &&
                     dx_ld % 8 == 0,
                 B200DM_ERR_SHAPE, "rmsnorm_bwd: need C %% 8 == 0, C <= 512 (C=%d), rows < 2^31", C);
  cudaStream_t st = (cudaStream_t)stream;
  int L = 1;
  while (L < C / 8 && L < 32) L <<= 1;
  constexpr int rmult = 2;     // measured: 1x / 3x / 4x SMs CTAs are 13-26 % slower
  int64_t blocks = (rows * L + 255) / 256, cap = (int64_t)num_sms() * rmult;   // few CTAs: every CTA ends with one dg atomic per channel
  unsigned grid = (unsigned)(blocks > cap ? cap : blocks);
  const bool two = C / 8 > 32;
#define RMS_BWD_T(TT, MV, LT)                                                                                \
  launch_k(rmsnorm_bwd_kernel<TT, MV, LT>, grid, 256, 0, st, (const TT*)dy, dy_ld, (const TT*)x, x_ld, g,   \
           (const TT*)res, res_ld, (TT*)dx, dx_ld, dg, (int)rows, C, L)
#define RMS_BWD_F32(MV, LT) RMS_BWD_T(float, MV, LT)
#define RMS_BWD_BF16(MV, LT) RMS_BWD_T(bf16, MV, LT)
  if (dtype == B200DM_F32) {
    RMS_DISPATCH(RMS_BWD_F32)
  } else {
    RMS_DISPATCH(RMS_BWD_BF16)
  }
#undef RMS_BWD_T
#undef RMS_BWD_F32
#undef RMS_BWD_BF16
  count_launch();
  return check_launch("rmsnorm_bwd");
}
