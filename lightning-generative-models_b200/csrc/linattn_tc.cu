// LinearAttention block of the UNet for INFERENCE as three tcgen05 launches that never write q, k, v:
//
//     y = RMSNorm_out( to_out( LinearAttention( to_qkv( RMSNorm(x) ) ) ) ) + x
//         (reference ddpm.py:205-238 LinearAttention.forward, :184-191 RMSNorm, :449/:464 the `attn(x) + x` skip)
//
// The unfused path (rmsnorm_fwd -> 1x1 conv -> linattn_fwd -> 1x1 conv -> rmsnorm_fwd) moves the [pixels][384] qkv
// tensor through HBM twice (write-bound conv, then the attention core); at the DDIM benchmark's first level that is
// 1.6 GB per block for 134 MB of input.  Here every pixel tile is projected on the tensor cores where it is needed:
//
//   pass 0  la_ctx_kernel<0>   kmax[b][h,d]  = max_n k[h,d,n]                       (k = Wk' x / |x|, not stored)
//   pass 1  la_ctx_kernel<1>   ctx[b][h][d][e] = sum_n exp(k - kmax) v[e,n],  s[b][h,d] = sum_n exp(k - kmax)
//   pass 2  la_out_kernel      q = softmax_d(Wq' x / |x|);  o = ctx^T q / s;  y = RMSNorm(Wout o + b) + x
//
// RMSNorm(x) = x / max(|x|, 1e-12) * g * sqrt(C): the gain g * sqrt(C) is folded into the projection weights
// (b200dm_pack_linattn_qkv) and the per-pixel 1/|x| is applied to the accumulators, so the kernels read the raw x.
// Orientation: passes 0/1 compute K^T and V^T (rows = channels, columns = pixels), so that a thread owns one channel
// row and the softmax over pixels and the bf16 operand rows of the second product (P V^T, K = pixels) are thread-local;
// pass 2 computes rows = pixels, so the softmax over d, the RMSNorm over channels and the residual are thread-local.
// All products are tcgen05.mma (M = 128) with TMEM accumulators; the four heads' 32x32 contexts are the diagonal
// blocks of one 128x128 product (the off-diagonal blocks are computed and ignored / multiplied by zeros: < 2 % of
// the UNet's FLOPs).  Softmax over n uses the exact maximum (pass 0), as the reference does.
#include "tc_common.cuh"

namespace b200dm {

using namespace tc;

int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const char* what);   // conv_tc.cu
bool tc_supported();

namespace {

constexpr int LB_TILE = 128;               // pixels per tile
constexpr int LB_BLK = 128 * 64 * 2;       // one [128 rows][64 bf16] SWIZZLE_128B operand block: 16 KiB
constexpr int LB_THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2..9 transform
constexpr int LB_XF = 256;                 // transform threads
constexpr int LB_NMEM = 4;
constexpr float LB_LOG2E = 1.4426950408889634f;
constexpr float LB_SCALE = 0.17677669529663687f;   // 32^-0.5
// workspace per (sample, split): kmax[128] | s[2 column halves][128] | ctx[h][e][d]
constexpr int LB_WS_S = 128, LB_WS_CTX = 384, LB_WS = 384 + 4096;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-byte chunk `chunk` (8 bf16) of row `row` of a [rows][64] SWIZZLE_128B block
__device__ __forceinline__ void sw128_store(uint8_t* block, int row, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(block + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}
__device__ __forceinline__ float sumsq8(uint4 v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    s = fmaf(f.x, f.x, fmaf(f.y, f.y, s));
  }
  return s;
}

struct LaCtxParams {
  int B, n, C, split, x_ld;
  const __nv_bfloat16* x;
  const float* mem_kv;
  float* ws;
};

// ---------------------------------------------------------------------------------------------------------------
// passes 0 and 1.  Work item = (sample, split): a contiguous range of 128-pixel tiles of one sample.
//   warp 0     TMA: per tile and 64-channel block one stage = {x block [128 px][64], Wk block [128][64] (, Wv block)}
//   warp 1     MMA1: D1k[128 ch][128 px] = Wk' X^T (, D1v = Wv' X^T);   MMA2 (pass 1): D2[128][128] += P V^T
//   warps 2..9 transform: 1/|x| per pixel; row r = TMEM lane, 64 pixel columns each:
//              pass 0: running row maximum of k;  pass 1: P = exp(k - m) and V as bf16 operand rows of MMA2, row sums
// ---------------------------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(LB_THREADS, 1)
la_ctx_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const LaCtxParams p) {
  constexpr int NW = PASS == 0 ? 1 : 2;
  constexpr int STAGE_BYTES = (1 + NW) * LB_BLK;
  constexpr int NST = PASS == 0 ? 4 : 3;
  constexpr int TMEM_COLS = PASS == 0 ? 128 : 512;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  constexpr int PV_BYTES = PASS == 1 ? 4 * LB_BLK : 0;        // P: 2 blocks (pixels 0..63 | 64..127), V: 2 blocks
  const uint32_t pv = base + NST * STAGE_BYTES;
  uint8_t* pv_ptr = base_ptr + NST * STAGE_BYTES;
  const uint32_t bar_base = pv + PV_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NST + s); };
  const uint32_t d1_full = bar_base + 8u * (2 * NST), d1_empty = d1_full + 8u, pv_full = d1_full + 16u,
                 pv_empty = d1_full + 24u, d2_full = d1_full + 32u, d2_empty = d1_full + 40u;
  const uint32_t tmem_slot = d1_full + 48u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  float* fs = reinterpret_cast<float*>(base_ptr + (bar_base - base) + 256);
  float* rn = fs;              // [2 tile parities][128]: 1 / |x| of the tile's pixels
  float* mx = fs + 256;        // [2][128]: row maxima of the two column halves (pass 0)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.C >> 6;
  const int tps = p.n / LB_TILE, per = tps / p.split;
  const int items = p.B * p.split;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < NST; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(d1_full, 1);
    mbar_init(d1_empty, 8);
    mbar_init(pv_full, 8);
    mbar_init(pv_empty, 1);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();   // after the TMEM allocation: a dependent CTA must not take the columns first
  pdl_wait();

  if (warp == 0) {
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int b = item / p.split, sp = item - b * p.split;
      for (int t = sp * per; t < (sp + 1) * per; ++t) {
        const int row0 = b * p.n + t * LB_TILE;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) {
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            const uint32_t dst = base + stage * STAGE_BYTES;
            tma_load_3d(dst, &tmX, full_bar(stage), kb * 64, row0, 0);
            tma_load_3d(dst + LB_BLK, &tmW, full_bar(stage), kb * 64, 128, 0);           // Wk' rows
            if (PASS == 1) tma_load_3d(dst + 2 * LB_BLK, &tmW, full_bar(stage), kb * 64, 256, 0);   // Wv' rows
          }
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint64_t desc0 = make_smem_desc(base, 16, 1024);
    const uint32_t hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
    const uint32_t pv_lo0 = (uint32_t)make_smem_desc(pv, 16, 1024);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, d1e = 0, pvf = 0, d2e = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int sp = item % p.split;
      const int t0 = sp * per, t1 = t0 + per;
      bool first2 = true;
      auto mma2 = [&]() {
        mbar_wait(pv_full, pvf);
        pvf ^= 1u;
        if (first2) {            // D2 of the previous item has been read out
          mbar_wait(d2_empty, d2e ^ 1u);
          d2e ^= 1u;
        }
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lohi(tmem_base + 256u, pv_lo0 + (uint32_t)((kk * LB_BLK + k * 32) >> 4), hi,
                             pv_lo0 + (uint32_t)(((2 + kk) * LB_BLK + k * 32) >> 4), hi, idesc,
                             (first2 && kk == 0 && k == 0) ? 0u : 1u);
          umma_commit(pv_empty);
        }
        __syncwarp();
        first2 = false;
      };
      for (int t = t0; t < t1; ++t) {
        mbar_wait(d1_empty, d1e ^ 1u);      // the transform warps have read D1 of the previous tile
        d1e ^= 1u;
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (leader) {
            const uint32_t so = (uint32_t)(stage * STAGE_BYTES) >> 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
              umma_bf16_lohi(tmem_base, lo0 + so + (uint32_t)((LB_BLK + k * 32) >> 4), hi,
                             lo0 + so + (uint32_t)((k * 32) >> 4), hi, idesc, acc);
              if (PASS == 1)
                umma_bf16_lohi(tmem_base + 128u, lo0 + so + (uint32_t)((2 * LB_BLK + k * 32) >> 4), hi,
                               lo0 + so + (uint32_t)((k * 32) >> 4), hi, idesc, acc);
            }
            umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(d1_full);
        __syncwarp();
        if (PASS == 1 && t > t0) mma2();      // P V^T of the previous tile, behind this tile's projections
      }
      if (PASS == 1) {
        mma2();
        if (leader) umma_commit(d2_full);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                       // TMEM lane = channel row (h = q, d = lane)
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    uint32_t d1f = 0, pve = 0, d2f = 0;
    int tcount = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int b = item / p.split, sp = item - b * p.split;
      float* wsi = p.ws + (size_t)item * LB_WS;
      float m = 0.f, mneg = 0.f;
      if (PASS == 1) {
        m = -INFINITY;
        for (int s = 0; s < p.split; ++s) m = fmaxf(m, p.ws[(size_t)(b * p.split + s) * LB_WS + r]);
#pragma unroll
        for (int j = 0; j < LB_NMEM; ++j) m = fmaxf(m, p.mem_kv[(q * 32 + lane) * LB_NMEM + j]);
        mneg = -m * LB_LOG2E;
      }
      float run_max = -INFINITY, run_sum = 0.f;
      for (int t = sp * per; t < (sp + 1) * per; ++t, ++tcount) {
        float* rnt = rn + (tcount & 1) * 128;
        if (half == 0) {       // 1 / |x| of pixel r of the tile
          const uint4* xr = reinterpret_cast<const uint4*>(p.x + (size_t)(b * p.n + t * LB_TILE + r) * p.x_ld);
          float ss = 0.f;
          for (int c = 0; c < (p.C >> 3); c += 4) {
            const uint4 v0 = xr[c], v1 = xr[c + 1], v2 = xr[c + 2], v3 = xr[c + 3];
            ss += (sumsq8(v0) + sumsq8(v1)) + (sumsq8(v2) + sumsq8(v3));
          }
          rnt[r] = rsqrtf(fmaxf(ss, 1e-24f));
        }
        named_bar_sync(1, LB_XF);
        mbar_wait(d1_full, d1f);
        d1f ^= 1u;
        tc_fence_after();
        if (PASS == 1) {
          mbar_wait(pv_empty, pve ^ 1u);     // MMA2 of the previous tile has read P and V
          pve ^= 1u;
        }
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int col = half * 64 + cc * 32;
          uint32_t kr[32];
          tmem_ld32(tmem_base + tlane + (uint32_t)col, kr);
          tmem_ld_wait();
          const float4* rn4 = reinterpret_cast<const float4*>(rnt + col);
          if (PASS == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 s4 = rn4[i];
              run_max = fmaxf(run_max, fmaxf(fmaxf(__uint_as_float(kr[4 * i]) * s4.x, __uint_as_float(kr[4 * i + 1]) * s4.y),
                                             fmaxf(__uint_as_float(kr[4 * i + 2]) * s4.z, __uint_as_float(kr[4 * i + 3]) * s4.w)));
            }
          } else {
            uint8_t* pblk = pv_ptr + half * LB_BLK;          // pixels [half*64, +64) = K block `half`
            uint8_t* vblk = pv_ptr + (2 + half) * LB_BLK;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 sa = rn4[2 * g], sb = rn4[2 * g + 1];
              float e[8];
              e[0] = ex2f(fmaf(__uint_as_float(kr[8 * g]) * sa.x, LB_LOG2E, mneg));
              e[1] = ex2f(fmaf(__uint_as_float(kr[8 * g + 1]) * sa.y, LB_LOG2E, mneg));
              e[2] = ex2f(fmaf(__uint_as_float(kr[8 * g + 2]) * sa.z, LB_LOG2E, mneg));
              e[3] = ex2f(fmaf(__uint_as_float(kr[8 * g + 3]) * sa.w, LB_LOG2E, mneg));
              e[4] = ex2f(fmaf(__uint_as_float(kr[8 * g + 4]) * sb.x, LB_LOG2E, mneg));
              e[5] = ex2f(fmaf(__uint_as_float(kr[8 * g + 5]) * sb.y, LB_LOG2E, mneg));
              e[6] = ex2f(fmaf(__uint_as_float(kr[8 * g + 6]) * sb.z, LB_LOG2E, mneg));
              e[7] = ex2f(fmaf(__uint_as_float(kr[8 * g + 7]) * sb.w, LB_LOG2E, mneg));
              run_sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
              sw128_store(pblk, r, cc * 4 + g,
                          make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7])));
            }
            uint32_t vr[32];
            tmem_ld32(tmem_base + tlane + 128u + (uint32_t)col, vr);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 sa = rn4[2 * g], sb = rn4[2 * g + 1];
              sw128_store(vblk, r, cc * 4 + g,
                          make_uint4(pack_bf16(__uint_as_float(vr[8 * g]) * sa.x, __uint_as_float(vr[8 * g + 1]) * sa.y),
                                     pack_bf16(__uint_as_float(vr[8 * g + 2]) * sa.z, __uint_as_float(vr[8 * g + 3]) * sa.w),
                                     pack_bf16(__uint_as_float(vr[8 * g + 4]) * sb.x, __uint_as_float(vr[8 * g + 5]) * sb.y),
                                     pack_bf16(__uint_as_float(vr[8 * g + 6]) * sb.z, __uint_as_float(vr[8 * g + 7]) * sb.w)));
            }
          }
        }
        tc_fence_before();
        if (PASS == 1) fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(d1_empty);
          if (PASS == 1) mbar_arrive(pv_full);
        }
      }
      if (PASS == 0) {
        mx[half * 128 + r] = run_max;
        named_bar_sync(2, LB_XF);
        if (half == 0) wsi[r] = fmaxf(mx[r], mx[128 + r]);
        named_bar_sync(2, LB_XF);
      } else {
        wsi[LB_WS_S + half * 128 + r] = run_sum;
        mbar_wait(d2_full, d2f);
        d2f ^= 1u;
        tc_fence_after();
        if (half == 0) {       // head q's block of D2: rows (q, d = lane), columns (q, e)
          uint32_t cr[32];
          tmem_ld32(tmem_base + tlane + 256u + (uint32_t)(q * 32), cr);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) wsi[LB_WS_CTX + (q * 32 + e) * 32 + lane] = __uint_as_float(cr[e]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

struct LaOutParams {
  int B, n, C, split, x_ld, y_ld;
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* mem_kv;
  const float* ws;
  const float* bout;
  const float* gout;
};

// ---------------------------------------------------------------------------------------------------------------
// pass 2.  Tiles of 128 pixels, a contiguous run per CTA (the per-sample context operand is rebuilt on a sample change).
//   warp 0     TMA: Wq' / Wout blocks once, then the x blocks of the tiles through a ring
//   warp 1     MMA: Dq[128 px][128] = X Wq'^T;  Do = Qs CtxBD^T;  Dy[128 px][C] = O Wout^T
//   warps 2..9 row = pixel = TMEM lane, two heads / half of the channels each:
//              T1 softmax_d(q / |x|) -> bf16 rows;  T2 Do -> bf16 rows;  epilogue +bias, RMSNorm, gain, + x, store
// CtxBD[(h,e)][(h',d)] = [h == h'] * scale * ctx[h][d][e] / s[h,d]  (memory key/values included), bf16.
// ---------------------------------------------------------------------------------------------------------------
template <int KB>
__global__ void __launch_bounds__(LB_THREADS, 1)
la_out_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
              const __grid_constant__ CUtensorMap tmWo, const LaOutParams p) {
  constexpr int C = 64 * KB;
  constexpr int NSX = 4;
  constexpr int WO_BLK = C * 128;               // [C rows][64] block of Wout
  constexpr int NCH = C / 2;                    // output channels per thread
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  constexpr int OFF_WQ = 0, OFF_WO = OFF_WQ + KB * LB_BLK, OFF_CTX = OFF_WO + 2 * WO_BLK,
                OFF_QO = OFF_CTX + 2 * LB_BLK, OFF_X = OFF_QO + 2 * LB_BLK, OFF_BAR = OFF_X + NSX * LB_BLK;
  auto xfull = [&](int s) { return base + OFF_BAR + 8u * s; };
  auto xempty = [&](int s) { return base + OFF_BAR + 8u * (NSX + s); };
  const uint32_t w_full = base + OFF_BAR + 8u * (2 * NSX), dq_full = w_full + 8u, q_full = w_full + 16u,
                 do_full = w_full + 24u, o_full = w_full + 32u, dy_full = w_full + 40u;
  const uint32_t tmem_slot = w_full + 48u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + OFF_BAR + 8 * (2 * NSX) + 48);
  float* fs = reinterpret_cast<float*>(base_ptr + OFF_BAR + 256);
  float* rn = fs;               // [128]
  float* ssy = fs + 128;        // [2][128]
  float* mrow = fs + 384;       // [128]  -kmax * log2e per (h, d)
  float* sinv = fs + 512;       // [128]  scale / s
  float* bias_s = fs + 640;     // [C]
  float* gain_s = fs + 640 + C; // [C]   g * sqrt(C)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tps = p.n / LB_TILE;
  const long long T = (long long)p.B * tps;
  const int tile0 = (int)(T * blockIdx.x / gridDim.x), tile1 = (int)(T * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmWo);
    for (int s = 0; s < NSX; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    mbar_init(w_full, 1);
    mbar_init(dq_full, 1);
    mbar_init(q_full, 8);
    mbar_init(do_full, 1);
    mbar_init(o_full, 8);
    mbar_init(dy_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();   // after the TMEM allocation: a dependent CTA must not take the columns first
  pdl_wait();

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader && tile1 > tile0) {
      mbar_expect_tx(w_full, KB * LB_BLK + 2 * WO_BLK);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(base + OFF_WQ + kb * LB_BLK, &tmW, w_full, kb * 64, 0, 0);
      for (int kk = 0; kk < 2; ++kk) tma_load_3d(base + OFF_WO + kk * WO_BLK, &tmWo, w_full, kk * 64, 0, 0);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(xempty(stage), phase ^ 1u);
        if (leader) {
          mbar_expect_tx(xfull(stage), LB_BLK);
          tma_load_3d(base + OFF_X + stage * LB_BLK, &tmX, xfull(stage), kb * 64, tile * LB_TILE, 0);
        }
        if (++stage == NSX) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescY = make_idesc_bf16(128, C, 0, 0);
    const uint64_t desc0 = make_smem_desc(base, 16, 1024);
    const uint32_t hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, qf = 0, of = 0;
    auto mma_q = [&]() {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(xfull(stage), phase);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base, lo0 + (uint32_t)((OFF_X + stage * LB_BLK + k * 32) >> 4), hi,
                           lo0 + (uint32_t)((OFF_WQ + kb * LB_BLK + k * 32) >> 4), hi, idesc128,
                           (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(xempty(stage));
        }
        __syncwarp();
        if (++stage == NSX) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(dq_full);
      __syncwarp();
    };
    if (tile1 > tile0) {
      mbar_wait(w_full, 0);
      tc_fence_after();
      mma_q();
    }
    for (int tile = tile0; tile < tile1; ++tile) {
      mbar_wait(q_full, qf);          // Qs rows (and, on a sample change, CtxBD) are in shared memory; Dq has been read
      qf ^= 1u;
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + 128u, lo0 + (uint32_t)((OFF_QO + kk * LB_BLK + k * 32) >> 4), hi,
                           lo0 + (uint32_t)((OFF_CTX + kk * LB_BLK + k * 32) >> 4), hi, idesc128,
                           (kk > 0 || k > 0) ? 1u : 0u);
        umma_commit(do_full);
      }
      __syncwarp();
      if (tile + 1 < tile1) mma_q();  // the next tile's projection runs behind this tile's transforms
      mbar_wait(o_full, of);
      of ^= 1u;
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + 256u, lo0 + (uint32_t)((OFF_QO + kk * LB_BLK + k * 32) >> 4), hi,
                           lo0 + (uint32_t)((OFF_WO + kk * WO_BLK + k * 32) >> 4), hi, idescY,
                           (kk > 0 || k > 0) ? 1u : 0u);
        umma_commit(dy_full);
      }
      __syncwarp();
    }
  } else {
    const int tt = threadIdx.x - 64;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    uint8_t* ctx_ptr = base_ptr + OFF_CTX;
    uint8_t* qo_ptr = base_ptr + OFF_QO + half * LB_BLK;     // columns [half*64, +64) = K block `half`
    const float sqrtC = sqrtf((float)C);
    for (int i = tt; i < C; i += LB_XF) {
      bias_s[i] = p.bout[i];
      gain_s[i] = p.gout[i] * sqrtC;
    }
    uint32_t dqf = 0, dof = 0, dyf = 0;
    int cur_b = -1;
    for (int tile = tile0; tile < tile1; ++tile) {
      const int b = tile / tps;
      if (b != cur_b) {
        // ---- context operand of sample b (every MMA that read the previous one has completed: dy_full was waited)
        cur_b = b;
        const float* wsb = p.ws + (size_t)b * p.split * LB_WS;
        if (tt < 128) {
          const float* mk = p.mem_kv + tt * LB_NMEM;       // [h][d][j], tt = h*32 + d
          float m = -INFINITY, s = 0.f;
          for (int sp = 0; sp < p.split; ++sp) {
            m = fmaxf(m, wsb[sp * LB_WS + tt]);
            s += wsb[sp * LB_WS + LB_WS_S + tt] + wsb[sp * LB_WS + LB_WS_S + 128 + tt];
          }
#pragma unroll
          for (int j = 0; j < LB_NMEM; ++j) m = fmaxf(m, mk[j]);
#pragma unroll
          for (int j = 0; j < LB_NMEM; ++j) s += ex2f((mk[j] - m) * LB_LOG2E);
          mrow[tt] = -m * LB_LOG2E;
          sinv[tt] = LB_SCALE / s;
        }
        named_bar_sync(3, LB_XF);
        {
          const int row = tt & 127, kblk = tt >> 7;        // row = (h, e); K block kblk holds heads 2*kblk, 2*kblk+1
          const int h = row >> 5, e = row & 31;
          const bool mine = kblk == (h >> 1);
          const float* mv = p.mem_kv + (128 + h * 32 + e) * LB_NMEM;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            uint4 out = make_uint4(0u, 0u, 0u, 0u);
            if (mine && (c >> 2) == (h & 1)) {
              const int d0 = (c & 3) * 8;
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = 0.f;
              for (int sp = 0; sp < p.split; ++sp) {
                const float4* src = reinterpret_cast<const float4*>(wsb + sp * LB_WS + LB_WS_CTX + row * 32 + d0);
                const float4 a = src[0], bq = src[1];
                v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
                v[4] += bq.x; v[5] += bq.y; v[6] += bq.z; v[7] += bq.w;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int hd = h * 32 + d0 + i;
                const float* mk = p.mem_kv + hd * LB_NMEM;
#pragma unroll
                for (int j = 0; j < LB_NMEM; ++j) v[i] = fmaf(ex2f(fmaf(mk[j], LB_LOG2E, mrow[hd])), mv[j], v[i]);
                v[i] *= sinv[hd];
              }
              out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
            sw128_store(ctx_ptr + kblk * LB_BLK, row, c, out);
          }
        }
        fence_proxy_async();
      }
      // ---- 1 / |x| of pixel r; this thread's half of the row is kept for the residual
      const __nv_bfloat16* xrow = p.x + (size_t)(tile * LB_TILE + r) * p.x_ld;
      uint4 res[NCH / 8];
      {
        const uint4* xr = reinterpret_cast<const uint4*>(xrow + half * NCH);
#pragma unroll
        for (int c = 0; c < NCH / 8; ++c) res[c] = xr[c];
        if (half == 0) {
          const uint4* xo = reinterpret_cast<const uint4*>(xrow + NCH);
          float ss = 0.f;
#pragma unroll
          for (int c = 0; c < NCH / 8; ++c) ss += sumsq8(res[c]) + sumsq8(xo[c]);
          rn[r] = rsqrtf(fmaxf(ss, 1e-24f));
        }
      }
      named_bar_sync(1, LB_XF);
      const float myrn = rn[r];
      // ---- T1: softmax over d of the two heads of this half
      mbar_wait(dq_full, dqf);
      dqf ^= 1u;
      tc_fence_after();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t qr[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(half * 64 + hh * 32), qr);
        tmem_ld_wait();
        const float sc = myrn * LB_LOG2E;
        float mxv = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) mxv = fmaxf(mxv, __uint_as_float(qr[i]) * sc);
        float e[32], s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[i] = ex2f(fmaf(__uint_as_float(qr[i]), sc, -mxv));
          s += e[i];
        }
        const float inv = 1.f / s;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          sw128_store(qo_ptr, r, hh * 4 + g,
                      make_uint4(pack_bf16(e[8 * g] * inv, e[8 * g + 1] * inv), pack_bf16(e[8 * g + 2] * inv, e[8 * g + 3] * inv),
                                 pack_bf16(e[8 * g + 4] * inv, e[8 * g + 5] * inv), pack_bf16(e[8 * g + 6] * inv, e[8 * g + 7] * inv)));
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
      // ---- T2: attention output rows as the operand of to_out
      mbar_wait(do_full, dof);
      dof ^= 1u;
      tc_fence_after();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t orr[32];
        tmem_ld32(tmem_base + tlane + 128u + (uint32_t)(half * 64 + hh * 32), orr);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g)
          sw128_store(qo_ptr, r, hh * 4 + g,
                      make_uint4(pack_bf16(__uint_as_float(orr[8 * g]), __uint_as_float(orr[8 * g + 1])),
                                 pack_bf16(__uint_as_float(orr[8 * g + 2]), __uint_as_float(orr[8 * g + 3])),
                                 pack_bf16(__uint_as_float(orr[8 * g + 4]), __uint_as_float(orr[8 * g + 5])),
                                 pack_bf16(__uint_as_float(orr[8 * g + 6]), __uint_as_float(orr[8 * g + 7]))));
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_full);
      // ---- epilogue: + bias, RMSNorm over the C channels (two threads per pixel), gain, + x
      mbar_wait(dy_full, dyf);
      dyf ^= 1u;
      tc_fence_after();
      float yv[NCH];
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; c += 32) {
        uint32_t yr[32];
        tmem_ld32(tmem_base + tlane + 256u + (uint32_t)(half * NCH + c), yr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          yv[c + i] = __uint_as_float(yr[i]) + bias_s[half * NCH + c + i];
          ss = fmaf(yv[c + i], yv[c + i], ss);
        }
      }
      tc_fence_before();
      ssy[half * 128 + r] = ss;
      named_bar_sync(2, LB_XF);
      const float inv = rsqrtf(fmaxf(ssy[r] + ssy[128 + r], 1e-24f));
      __nv_bfloat16* yrow = p.y + (size_t)(tile * LB_TILE + r) * p.y_ld + half * NCH;
#pragma unroll
      for (int c = 0; c < NCH / 8; ++c) {
        const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&res[c]);
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xr2 = __bfloat1622float2(rh[i]);
          const int ch = c * 8 + 2 * i;
          o[i] = pack_bf16(fmaf(yv[ch] * inv, gain_s[half * NCH + ch], xr2.x),
                           fmaf(yv[ch + 1] * inv, gain_s[half * NCH + ch + 1], xr2.y));
        }
        reinterpret_cast<uint4*>(yrow)[c] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[o][c] = bf16(w[o][c] * g[c] * sqrt(C))   (RMSNorm gain folded into to_qkv)
__global__ void la_pack_qkv_kernel(const float* __restrict__ w, const float* __restrict__ g, __nv_bfloat16* __restrict__ out,
                                   int C, int total) {
  pdl_prologue();
  const float sc = sqrtf((float)C);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(w[i] * g[i % C] * sc);
}

int la_split(int B, int n) {
  const int tps = n / LB_TILE;
  int split = 1;
  while (split * 2 <= 8 && tps % (split * 2) == 0 && (long long)B * split < 4ll * num_sms()) split *= 2;
  return split;
}

int make_map2d(CUtensorMap* m, const void* ptr, long long rows, int cols, long long ld, int box_rows, const char* what) {
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  return encode_map(m, ptr, 3, dims, str, box, what);
}

}  // namespace
}  // namespace b200dm

using namespace b200dm;

extern "C" int b200dm_pack_linattn_qkv(const float* w, const float* g, void* out, int32_t C, void* stream) {
  B200DM_REQUIRE(w && g && out && C > 0, B200DM_ERR_SHAPE, "pack_linattn_qkv: null argument");
  const int total = 384 * C;
  launch_k(la_pack_qkv_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, w, g, (__nv_bfloat16*)out, (int)C, total);
  count_launch();
  return check_launch("pack_linattn_qkv");
}

extern "C" int64_t b200dm_linattn_block_ws_floats(int32_t B, int32_t n) {
  if (B <= 0 || n <= 0 || n % LB_TILE) return 0;
  return (int64_t)B * la_split(B, n) * LB_WS;
}

extern "C" int b200dm_linattn_block_supported(const b200dm_linattn_block_desc* d) {
  if (!d || !tc_supported()) return 0;
  if (d->B <= 0 || d->n <= 0 || d->n % LB_TILE) return 0;
  if (d->C != 64 && d->C != 128) return 0;
  if (d->x_ld % 8 || d->y_ld % 8 || d->x_ld < d->C || d->y_ld < d->C) return 0;
  if (((uintptr_t)d->x | (uintptr_t)d->y | (uintptr_t)d->wqkv | (uintptr_t)d->wout) & 15) return 0;
  if ((long long)d->B * d->n >= (1ll << 31)) return 0;
  return 1;
}

extern "C" int b200dm_linattn_block_fwd(const b200dm_linattn_block_desc* d, void* stream) {
  B200DM_REQUIRE(d != nullptr, B200DM_ERR_SHAPE, "linattn_block_fwd: null descriptor");
  B200DM_REQUIRE(b200dm_linattn_block_supported(d) == 1, B200DM_ERR_UNSUPPORTED,
                 "linattn_block_fwd: needs sm_100, bf16, n %% 128 == 0 (n=%d), C in {64,128} (C=%d), 16-byte aligned tensors",
                 d->n, d->C);
  B200DM_REQUIRE(d->x && d->y && d->wqkv && d->wout && d->bout && d->gout && d->mem_kv && d->ws, B200DM_ERR_SHAPE,
                 "linattn_block_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int split = la_split(d->B, d->n);
  const long long rows = (long long)d->B * d->n;
  CUtensorMap tmX, tmW, tmWo;
  int rc = make_map2d(&tmX, d->x, rows, d->C, d->x_ld, 128, "linattn_block x");
  if (rc) return rc;
  rc = make_map2d(&tmW, d->wqkv, 384, d->C, d->C, 128, "linattn_block wqkv");
  if (rc) return rc;
  rc = make_map2d(&tmWo, d->wout, d->C, 128, 128, d->C, "linattn_block wout");
  if (rc) return rc;
  LaCtxParams pc{};
  pc.B = d->B; pc.n = d->n; pc.C = d->C; pc.split = split; pc.x_ld = d->x_ld;
  pc.x = (const __nv_bfloat16*)d->x; pc.mem_kv = d->mem_kv; pc.ws = d->ws;
  const int items = d->B * split;
  const int grid_a = items < num_sms() ? items : num_sms();
  constexpr int smem0 = 4 * 2 * LB_BLK + 256 + 4096 + 1024;
  constexpr int smem1 = 3 * 3 * LB_BLK + 4 * LB_BLK + 256 + 4096 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e0 = cudaFuncSetAttribute(la_ctx_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0);
    cudaError_t e1 = cudaFuncSetAttribute(la_ctx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    cudaError_t e2 = cudaFuncSetAttribute(la_out_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          1 * LB_BLK + 2 * 64 * 128 + 8 * LB_BLK + 256 + 8192 + 1024);
    cudaError_t e3 = cudaFuncSetAttribute(la_out_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          2 * LB_BLK + 2 * 128 * 128 + 8 * LB_BLK + 256 + 8192 + 1024);
    B200DM_REQUIRE(e0 == cudaSuccess && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess, B200DM_ERR_CUDA,
                   "linattn_block_fwd: cudaFuncSetAttribute failed");
    configured = true;
  }
  launch_k(la_ctx_kernel<0>, grid_a, LB_THREADS, smem0, st, tmX, tmW, pc);
  launch_k(la_ctx_kernel<1>, grid_a, LB_THREADS, smem1, st, tmX, tmW, pc);
  LaOutParams po{};
  po.B = d->B; po.n = d->n; po.C = d->C; po.split = split; po.x_ld = d->x_ld; po.y_ld = d->y_ld;
  po.x = (const __nv_bfloat16*)d->x; po.y = (__nv_bfloat16*)d->y; po.mem_kv = d->mem_kv; po.ws = d->ws;
  po.bout = d->bout; po.gout = d->gout;
  const long long tiles = rows / LB_TILE;
  const int grid_b = tiles < num_sms() ? (int)tiles : num_sms();
  if (d->C == 64)
    launch_k(la_out_kernel<1>, grid_b, LB_THREADS, 1 * LB_BLK + 2 * 64 * 128 + 8 * LB_BLK + 256 + 8192 + 1024, st, tmX, tmW,
             tmWo, po);
  else
    launch_k(la_out_kernel<2>, grid_b, LB_THREADS, 2 * LB_BLK + 2 * 128 * 128 + 8 * LB_BLK + 256 + 8192 + 1024, st, tmX, tmW,
             tmWo, po);
  count_launch(3);
  return check_launch("linattn_block_fwd");
}
