// LinearAttention block of the UNet for INFERENCE as four launches that never write q, k, v:
//
//     y = RMSNorm_out( to_out( LinearAttention( to_qkv( RMSNorm(x) ) ) ) ) + x
//         (reference ddpm.py:205-238 LinearAttention.forward, :107-113 RMSNorm, :449/:464 the `attn(x) + x` skip)
//
// The unfused path (rmsnorm_fwd -> 1x1 conv -> linattn_fwd -> 1x1 conv -> rmsnorm_fwd) moves the [pixels][384] qkv
// tensor through HBM twice (write-bound conv, then the attention core); at the DDIM benchmark's first level that is
// 1.6 GB per block for 134 MB of input.  Here every pixel tile is projected on the tensor cores where it is needed:
//
//   pass 0  la_kmax_kernel  rn[pixel] = 1 / |x|;  kmax[b][h,d] = max_n k[h,d,n]        (k = Wk' x / |x|, not stored)
//   pass 1  la_ctx_kernel   ctx[b][h][d][e] = sum_n exp(k - kmax) v[e,n],  s[b][h,d] = sum_n exp(k - kmax)
//   mid     la_mid_kernel   Mb[b] = Wout * blockdiag_h(scale * (ctx + memory kv)^T / s)              ([C][128] bf16)
//   pass 2  la_out_kernel   q = softmax_d(Wq' x / |x|);  y = RMSNorm(Mb q + bias) + x
//
// RMSNorm(x) = x / max(|x|, 1e-12) * g * sqrt(C): the gain g * sqrt(C) is folded into the projection weights
// (b200dm_pack_linattn_qkv) and the per-pixel 1/|x| is applied to the accumulators, so the kernels read the raw x.
// Orientation: passes 0/1 compute K^T and V^T (rows = channels, columns = pixels), so that a thread owns one channel
// row and the softmax over pixels and the bf16 operand rows of the second product (P V^T, K = pixels) are thread-local;
// pass 2 computes rows = pixels, so the softmax over d, the RMSNorm over channels and the residual are thread-local.
// All products are tcgen05.mma (M = 128) with TMEM accumulators; the four heads' 32x32 contexts are the diagonal
// blocks of one 128x128 product (the off-diagonal blocks are computed and ignored: < 2 % of the UNet's FLOPs), and
// to_out is applied to the per-sample context once (mid) instead of to every pixel.  Softmax over n uses the exact
// maximum (pass 0), as the reference does.  18 warps per CTA: TMA producer, MMA issuer, 16 transform warps.
#include "tc_common.cuh"

namespace b200dm {

using namespace tc;

int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const char* what);   // conv_tc.cu
bool tc_supported();

#ifdef B200DM_PHASE_TIMING
__device__ long long* g_la_tbuf = nullptr;
#define LA_TS(slot) do { if (g_la_tbuf && blockIdx.x == 0) g_la_tbuf[(slot)] = clock64(); } while (0)
#else
#define LA_TS(slot) do { } while (0)
#endif

namespace {

constexpr int LB_TILE = 128;               // pixels per tile
constexpr int LB_BLK = 128 * 64 * 2;       // one [128 rows][64 bf16] SWIZZLE_128B operand block: 16 KiB
constexpr int LB_THREADS = 576;            // warp 0 TMA, warp 1 MMA, warps 2..17 transform (four per TMEM lane quarter)
constexpr int LB_XF = 512;                 // transform threads
constexpr int LB_NMEM = 4;
constexpr float LB_LOG2E = 1.4426950408889634f;
constexpr float LB_SCALE = 0.17677669529663687f;   // 32^-0.5
// workspace per (sample, split): kmax[128] | s[4 column parts][128] | ctx[h][e][d]
constexpr int LB_WS_S = 128, LB_WS_CTX = 640, LB_WS = 640 + 4096;

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

struct LaParams {
  int B, n, C, split, x_ld, y_ld;
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const __nv_bfloat16* wout;   // [C][128] bf16
  const float* mem_kv;
  const float* bout;
  const float* gout;
  float* ws;                   // per (sample, split) partials
  float* rn;                   // [B*n]: 1 / |x| per pixel (written by pass 0)
  __nv_bfloat16* mb;           // [B][C][128]: per-sample operand of pass 2 (written by la_mid_kernel)
};

// Tiles of this CTA in order: work items (sample, split) are dealt round-robin, `per` consecutive tiles each.
// Walks incrementally (the divisions happen once per item, not once per tile).
struct LaTileIter {
  int item, b, sp, tt, per, split, n, tile_px;
  __device__ __forceinline__ LaTileIter(int per_, int split_, int n_, int tile_px_)
      : item((int)blockIdx.x), tt(0), per(per_), split(split_), n(n_), tile_px(tile_px_) {
    b = item / split;
    sp = item - b * split;
  }
  __device__ __forceinline__ int row0() const { return b * n + (sp * per + tt) * tile_px; }
  __device__ __forceinline__ bool first() const { return tt == 0; }
  __device__ __forceinline__ bool last() const { return tt == per - 1; }
  __device__ __forceinline__ void next() {
    if (++tt == per) {
      tt = 0;
      item += (int)gridDim.x;
      b = item / split;
      sp = item - b * split;
    }
  }
};
__device__ __forceinline__ int la_num_items(int items) {
  return ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// pass 0: 1/|x| of every pixel (kept for passes 1 and 2) and the row maxima of k = Wk' x / |x| per (sample, split).
//   warp 0 TMA (Wk' once, x blocks through a ring) | warp 1 MMA: D1[slot][128 ch][128 px] = Wk' X^T, two TMEM
//   accumulators | warps 2..17: row r = channel (TMEM lane), 32 pixel columns each; warps 2..5 also compute 1/|x| of
//   pixel r of the tile from its global row, requested one tile ahead.
// ---------------------------------------------------------------------------------------------------------------
template <int KB>
__global__ void __launch_bounds__(LB_THREADS, 1)
la_kmax_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const LaParams p) {
  constexpr int NSX = 4;
  constexpr int OFF_X = KB * LB_BLK, OFF_BAR = OFF_X + NSX * LB_BLK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto xfull = [&](int s) { return base + OFF_BAR + 8u * s; };
  auto xempty = [&](int s) { return base + OFF_BAR + 8u * (NSX + s); };
  const uint32_t w_full = base + OFF_BAR + 8u * (2 * NSX);
  auto d1_full = [&](int s) { return w_full + 8u + 8u * s; };
  auto d1_empty = [&](int s) { return w_full + 24u + 8u * s; };
  const uint32_t tmem_slot = w_full + 40u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + OFF_BAR + 8 * (2 * NSX) + 40);
  float* fs = reinterpret_cast<float*>(base_ptr + OFF_BAR + 256);
  float* rns = fs;             // [2 tile parities][128]
  float* mx = fs + 256;        // [4 column parts][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (p.n / LB_TILE) / p.split;
  const int ntiles = la_num_items(p.B * p.split) * per;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < NSX; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();   // after the TMEM allocation: a dependent CTA must not take the columns first
  pdl_wait();

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader && ntiles > 0) {
      mbar_expect_tx(w_full, KB * LB_BLK);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(base + kb * LB_BLK, &tmW, w_full, kb * 64, 128, 0);   // Wk' rows
    }
    int stage = 0;
    uint32_t phase = 0;
    LaTileIter it(per, p.split, p.n, LB_TILE);
    for (int j = 0; j < ntiles; ++j, it.next()) {
      const int row0 = it.row0();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(xempty(stage), phase ^ 1u);
        if (leader) {
          mbar_expect_tx(xfull(stage), LB_BLK);
          tma_load_3d(base + OFF_X + stage * LB_BLK, &tmX, xfull(stage), kb * 64, row0, 0);
        }
        if (++stage == NSX) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint64_t desc0 = make_smem_desc(base, 16, 1024);
    const uint32_t hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    if (ntiles > 0) {
      mbar_wait(w_full, 0);
      tc_fence_after();
    }
    for (int j = 0; j < ntiles; ++j) {
      const int slot = j & 1;
      mbar_wait(d1_empty(slot), (uint32_t)((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(xfull(stage), phase);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + (uint32_t)(slot * 128), lo0 + (uint32_t)((kb * LB_BLK + k * 32) >> 4), hi,
                           lo0 + (uint32_t)((OFF_X + stage * LB_BLK + k * 32) >> 4), hi, idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(xempty(stage));
        }
        __syncwarp();
        if (++stage == NSX) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(d1_full(slot));
      __syncwarp();
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;     // half = column part 0..3
    const int r = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    constexpr int NV = KB * 8;       // 16-byte vectors per pixel row
    uint4 xr[NV];
    LaTileIter it(per, p.split, p.n, LB_TILE), itn(per, p.split, p.n, LB_TILE);    // current tile / the one fetched ahead
    auto fetch = [&](int j) {
      if (half == 0 && j < ntiles) {
        const uint4* src = reinterpret_cast<const uint4*>(p.x + (size_t)(itn.row0() + r) * p.x_ld);
#pragma unroll
        for (int c = 0; c < NV; ++c) xr[c] = src[c];
      }
      itn.next();
    };
    fetch(0);
    float run_max = -INFINITY;
    for (int j = 0; j < ntiles; ++j, it.next()) {
      const int row0 = it.row0();
      const int slot = j & 1;
      float* rnt = rns + slot * 128;
      if (half == 0) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < NV; ++c) ss += sumsq8(xr[c]);
        const float rv = rsqrtf(fmaxf(ss, 1e-24f));
        rnt[r] = rv;
        p.rn[row0 + r] = rv;
      }
      fetch(j + 1);
      named_bar_sync(1, LB_XF);
      mbar_wait(d1_full(slot), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      {
        const int col = half * 32;
        uint32_t kr[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(slot * 128 + col), kr);
        tmem_ld_wait();
        const float4* rn4 = reinterpret_cast<const float4*>(rnt + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 s4 = rn4[i];
          run_max = fmaxf(run_max, fmaxf(fmaxf(__uint_as_float(kr[4 * i]) * s4.x, __uint_as_float(kr[4 * i + 1]) * s4.y),
                                         fmaxf(__uint_as_float(kr[4 * i + 2]) * s4.z, __uint_as_float(kr[4 * i + 3]) * s4.w)));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d1_empty(slot));
      if (it.last()) {
        mx[half * 128 + r] = run_max;
        named_bar_sync(2, LB_XF);
        if (half == 0) p.ws[(size_t)it.item * LB_WS + r] = fmaxf(fmaxf(mx[r], mx[128 + r]), fmaxf(mx[256 + r], mx[384 + r]));
        named_bar_sync(2, LB_XF);
        run_max = -INFINITY;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 1: contexts and softmax denominators per (sample, split), 64-pixel tiles, everything double-buffered.
//   warp 0 TMA: Wk', Wv' once; x blocks [64 px][64 ch] through a ring
//   warp 1 MMA1: D1k[slot][128 ch][64 px] = Wk' X^T, D1v[slot] = Wv' X^T;  MMA2: D2[128][128] += P[slot] V[slot]^T
//   warps 2..17 row r = channel, 16 pixel columns each: P = exp(k / |x| - m), V = v / |x| as bf16 operand rows
// While the transform warps work on tile j, the tensor pipe runs MMA1 of tile j+1 and MMA2 of tile j-1.
// ---------------------------------------------------------------------------------------------------------------
template <int KB>
__global__ void __launch_bounds__(LB_THREADS, 1)
la_ctx_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const LaParams p) {
  constexpr int TPX = 64;
  constexpr int XB = TPX * 64 * 2;       // 8 KiB
  constexpr int NSX = 6;
  constexpr int OFF_WV = KB * LB_BLK, OFF_X = 2 * KB * LB_BLK, OFF_PV = OFF_X + NSX * XB, OFF_BAR = OFF_PV + 4 * LB_BLK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto xfull = [&](int s) { return base + OFF_BAR + 8u * s; };
  auto xempty = [&](int s) { return base + OFF_BAR + 8u * (NSX + s); };
  const uint32_t w_full = base + OFF_BAR + 8u * (2 * NSX);
  auto d1_full = [&](int s) { return w_full + 8u + 8u * s; };
  auto d1_empty = [&](int s) { return w_full + 24u + 8u * s; };
  auto pv_full = [&](int s) { return w_full + 40u + 8u * s; };
  auto pv_empty = [&](int s) { return w_full + 56u + 8u * s; };
  const uint32_t d2_full = w_full + 72u, d2_empty = w_full + 80u;
  const uint32_t tmem_slot = w_full + 88u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + OFF_BAR + 8 * (2 * NSX) + 88);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (p.n / TPX) / p.split;
  const int ntiles = la_num_items(p.B * p.split) * per;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < NSX; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), 16);
      mbar_init(pv_full(s), 16);
      mbar_init(pv_empty(s), 1);
    }
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader && ntiles > 0) {
      mbar_expect_tx(w_full, 2 * KB * LB_BLK);
      for (int kb = 0; kb < KB; ++kb) {
        tma_load_3d(base + kb * LB_BLK, &tmW, w_full, kb * 64, 128, 0);            // Wk' rows
        tma_load_3d(base + OFF_WV + kb * LB_BLK, &tmW, w_full, kb * 64, 256, 0);   // Wv' rows
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    LaTileIter it(per, p.split, p.n, TPX);
    for (int j = 0; j < ntiles; ++j, it.next()) {
      const int row0 = it.row0();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(xempty(stage), phase ^ 1u);
        if (leader) {
          mbar_expect_tx(xfull(stage), XB);
          tma_load_3d(base + OFF_X + stage * XB, &tmX, xfull(stage), kb * 64, row0, 0);
        }
        if (++stage == NSX) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc1 = make_idesc_bf16(128, TPX, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_bf16(128, 128, 0, 0);
    const uint64_t desc0 = make_smem_desc(base, 16, 1024);
    const uint32_t hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, d2e = 0;
    if (ntiles > 0) {
      mbar_wait(w_full, 0);
      tc_fence_after();
    }
    int tt2 = 0;                // position of MMA2's tile inside its item
    auto mma2 = [&](int j) {
      const bool first = tt2 == 0, last = tt2 == per - 1;
      if (++tt2 == per) tt2 = 0;
      const int slot = j & 1;
      mbar_wait(pv_full(slot), (uint32_t)((j >> 1) & 1));
      if (first) {              // D2 of the previous item has been read out
        mbar_wait(d2_empty, d2e ^ 1u);
        d2e ^= 1u;
      }
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_lohi(tmem_base + 256u, lo0 + (uint32_t)((OFF_PV + slot * 2 * LB_BLK + k * 32) >> 4), hi,
                         lo0 + (uint32_t)((OFF_PV + slot * 2 * LB_BLK + LB_BLK + k * 32) >> 4), hi, idesc2,
                         (first && k == 0) ? 0u : 1u);
        umma_commit(pv_empty(slot));
        if (last) umma_commit(d2_full);
      }
      __syncwarp();
    };
    for (int j = 0; j < ntiles; ++j) {
      const int slot = j & 1;
      mbar_wait(d1_empty(slot), (uint32_t)((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(xfull(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint32_t xo = lo0 + (uint32_t)((OFF_X + stage * XB) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + (uint32_t)(slot * 128), lo0 + (uint32_t)((kb * LB_BLK + k * 32) >> 4), hi,
                           xo + (uint32_t)((k * 32) >> 4), hi, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + (uint32_t)(slot * 128 + 64), lo0 + (uint32_t)((OFF_WV + kb * LB_BLK + k * 32) >> 4), hi,
                           xo + (uint32_t)((k * 32) >> 4), hi, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(xempty(stage));
        }
        __syncwarp();
        if (++stage == NSX) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(d1_full(slot));
      __syncwarp();
      if (j > 0) mma2(j - 1);
    }
    if (ntiles > 0) mma2(ntiles - 1);
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                       // channel row: h = q, d = lane
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    uint32_t d2f = 0;
    float mneg = 0.f, run_sum = 0.f;
    LaTileIter it(per, p.split, p.n, TPX);
    for (int j = 0; j < ntiles; ++j, it.next()) {
      const int slot = j & 1;
      const uint32_t par = (uint32_t)((j >> 1) & 1);
      if (it.first()) {
        float m = -INFINITY;
        for (int s = 0; s < p.split; ++s) m = fmaxf(m, p.ws[(size_t)(it.b * p.split + s) * LB_WS + r]);
#pragma unroll
        for (int jj = 0; jj < LB_NMEM; ++jj) m = fmaxf(m, p.mem_kv[r * LB_NMEM + jj]);
        mneg = -m * LB_LOG2E;
        run_sum = 0.f;
      }
      float4 rnv[4];                                   // 1 / |x| of this thread's 16 pixel columns
      {
        const float4* src = reinterpret_cast<const float4*>(p.rn + it.row0() + half * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) rnv[i] = src[i];
      }
      uint8_t* pblk = base_ptr + OFF_PV + slot * 2 * LB_BLK;
      uint8_t* vblk = pblk + LB_BLK;
      if (warp == 2 && lane == 0 && j < 16) LA_TS(200 + j * 4);
      mbar_wait(d1_full(slot), par);
      tc_fence_after();
      if (warp == 2 && lane == 0 && j < 16) LA_TS(200 + j * 4 + 1);
      uint32_t kr[16], vr[16];
      tmem_ld16(tmem_base + tlane + (uint32_t)(slot * 128 + half * 16), kr);
      tmem_ld16(tmem_base + tlane + (uint32_t)(slot * 128 + 64 + half * 16), vr);
      tmem_ld_wait();
      mbar_wait(pv_empty(slot), par ^ 1u);             // MMA2 of tile j-2 has read this P / V slot
      if (warp == 2 && lane == 0 && j < 16) LA_TS(200 + j * 4 + 2);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4 sa = rnv[2 * g], sb = rnv[2 * g + 1];
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
        sw128_store(pblk, r, half * 2 + g,
                    make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7])));
        sw128_store(vblk, r, half * 2 + g,
                    make_uint4(pack_bf16(__uint_as_float(vr[8 * g]) * sa.x, __uint_as_float(vr[8 * g + 1]) * sa.y),
                               pack_bf16(__uint_as_float(vr[8 * g + 2]) * sa.z, __uint_as_float(vr[8 * g + 3]) * sa.w),
                               pack_bf16(__uint_as_float(vr[8 * g + 4]) * sb.x, __uint_as_float(vr[8 * g + 5]) * sb.y),
                               pack_bf16(__uint_as_float(vr[8 * g + 6]) * sb.z, __uint_as_float(vr[8 * g + 7]) * sb.w)));
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(d1_empty(slot));
        mbar_arrive(pv_full(slot));
      }
      if (warp == 2 && lane == 0 && j < 16) LA_TS(200 + j * 4 + 3);
      if (it.last()) {
        float* wsi = p.ws + (size_t)it.item * LB_WS;
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
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// between passes 1 and 2, one CTA per sample: the normalised contexts (memory key/values included) times to_out:
//   Mb[c][(h,d)] = sum_e Wout[c][(h,e)] * scale * ctx[h][d][e] / s[h,d]          (bf16 [C][128])
// so that pass 2 goes from the softmaxed q straight to the to_out pre-activation:  y_pre = Mb q + bias.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
la_mid_kernel(const LaParams p) {
  pdl_prologue();
  __shared__ float mrow[128], sinv[128];
  __shared__ __align__(16) float ctxn[128 * 32];     // [h][e][d]
  const int b = blockIdx.x, tt = threadIdx.x;
  const float* wsb = p.ws + (size_t)b * p.split * LB_WS;
  if (tt < 128) {
    const float* mk = p.mem_kv + tt * LB_NMEM;
    float m = -INFINITY, s = 0.f;
    for (int sp = 0; sp < p.split; ++sp) {
      m = fmaxf(m, wsb[sp * LB_WS + tt]);
      s += (wsb[sp * LB_WS + LB_WS_S + tt] + wsb[sp * LB_WS + LB_WS_S + 128 + tt]) +
           (wsb[sp * LB_WS + LB_WS_S + 256 + tt] + wsb[sp * LB_WS + LB_WS_S + 384 + tt]);
    }
#pragma unroll
    for (int j = 0; j < LB_NMEM; ++j) m = fmaxf(m, mk[j]);
#pragma unroll
    for (int j = 0; j < LB_NMEM; ++j) s += ex2f((mk[j] - m) * LB_LOG2E);
    mrow[tt] = -m * LB_LOG2E;
    sinv[tt] = LB_SCALE / s;
  }
  __syncthreads();
  for (int i = tt; i < 4096; i += 256) {
    const int d = i & 31, he = i >> 5, h = he >> 5;
    const int hd = h * 32 + d;
    float v = 0.f;
    for (int sp = 0; sp < p.split; ++sp) v += wsb[sp * LB_WS + LB_WS_CTX + i];
    const float* mk = p.mem_kv + hd * LB_NMEM;
    const float* mv = p.mem_kv + (128 + he) * LB_NMEM;
#pragma unroll
    for (int j = 0; j < LB_NMEM; ++j) v = fmaf(ex2f(fmaf(mk[j], LB_LOG2E, mrow[hd])), mv[j], v);
    ctxn[i] = v * sinv[hd];
  }
  __syncthreads();
  // warp task = (row group of 32 output channels, head h, 8 consecutive d); lane = output channel
  const int warp = tt >> 5, lane = tt & 31;
  const int tasks = (p.C >> 5) * 16;
  for (int task = warp; task < tasks; task += 8) {
    const int rg = task >> 4, hc = task & 15, h = hc >> 2, d0 = (hc & 3) * 8;
    const int c = rg * 32 + lane;
    const uint4* wrow = reinterpret_cast<const uint4*>(p.wout + (size_t)c * 128 + h * 32);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int e8 = 0; e8 < 4; ++e8) {
      const uint4 wv = wrow[e8];
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
#pragma unroll
      for (int i2 = 0; i2 < 4; ++i2) {
        const float2 w2 = __bfloat1622float2(wh[i2]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float w = u == 0 ? w2.x : w2.y;
          const int e = e8 * 8 + i2 * 2 + u;
          const float4* cp = reinterpret_cast<const float4*>(ctxn + (h * 32 + e) * 32 + d0);
          const float4 a = cp[0], bq = cp[1];
          acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
          acc[4] = fmaf(w, bq.x, acc[4]); acc[5] = fmaf(w, bq.y, acc[5]); acc[6] = fmaf(w, bq.z, acc[6]); acc[7] = fmaf(w, bq.w, acc[7]);
        }
      }
    }
    *reinterpret_cast<uint4*>(p.mb + ((size_t)b * p.C + c) * 128 + h * 32 + d0) =
        make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pass 2.  Tiles of 128 pixels, a contiguous run per CTA, two tiles in flight (TMEM / operand slots j & 1).
//   warp 0 TMA: Wq' once; per tile the x blocks; on a sample change that sample's Mb (two buffers)
//   warp 1 MMA: Dq[slot][128 px][128] = X Wq'^T;   Dy[slot][128 px][C] = Qs[slot] Mb^T
//   warps 2..9   T1(j): row = pixel = TMEM lane, two heads each: softmax over d of q / |x| -> bf16 operand rows
//   warps 10..17 E(j):  row = pixel, half of the channels each: + bias, RMSNorm, gain, + x, staged and stored coalesced
// The softmax warps (MUFU / ALU bound) and the epilogue warps (TMEM / memory bound) work on different tiles at the same
// time; the tensor pipe runs ahead of both.
// ---------------------------------------------------------------------------------------------------------------
template <int KB>
__global__ void __launch_bounds__(LB_THREADS, 1)
la_out_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
              const __grid_constant__ CUtensorMap tmM, const LaParams p) {
  constexpr int C = 64 * KB;
  constexpr int NSX = KB == 1 ? 4 : 3;
  constexpr int MB_BLK = C * 128;               // [C rows][64] block of Mb
  constexpr int NCH = C / 2;                    // output channels per thread
  constexpr int OFF_MB = KB * LB_BLK, OFF_QS = OFF_MB + 4 * MB_BLK, OFF_X = OFF_QS + 4 * LB_BLK,
                OFF_BAR = OFF_X + NSX * LB_BLK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto xfull = [&](int s) { return base + OFF_BAR + 8u * s; };
  auto xempty = [&](int s) { return base + OFF_BAR + 8u * (NSX + s); };
  const uint32_t w_full = base + OFF_BAR + 8u * (2 * NSX);
  auto mb_full = [&](int s) { return w_full + 8u + 8u * s; };
  auto mb_empty = [&](int s) { return w_full + 24u + 8u * s; };
  auto dq_full = [&](int s) { return w_full + 40u + 8u * s; };
  auto q_full = [&](int s) { return w_full + 56u + 8u * s; };
  auto dy_full = [&](int s) { return w_full + 72u + 8u * s; };
  auto e_done = [&](int s) { return w_full + 88u + 8u * s; };
  auto stg_free = [&](int s) { return w_full + 104u + 8u * s; };
  const uint32_t tmem_slot = w_full + 120u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + OFF_BAR + 8 * (2 * NSX) + 120);
  float* fs = reinterpret_cast<float*>(base_ptr + OFF_BAR + 256);
  float* ssy = fs;              // [2 slots][2 halves][128]
  float* bias_s = fs + 512;     // [C]
  float* gain_s = fs + 512 + C; // [C]   g * sqrt(C)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tps = p.n / LB_TILE;
  const long long T = (long long)p.B * tps;
  const int tile0 = (int)(T * blockIdx.x / gridDim.x), tile1 = (int)(T * (blockIdx.x + 1) / gridDim.x);
  const int nt = tile1 - tile0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmM);
    for (int s = 0; s < NSX; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(mb_full(s), 1);
      mbar_init(mb_empty(s), 1);
      mbar_init(dq_full(s), 1);
      mbar_init(q_full(s), 8);
      mbar_init(dy_full(s), 1);
      mbar_init(e_done(s), 8);
      mbar_init(stg_free(s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader && nt > 0) {
      mbar_expect_tx(w_full, KB * LB_BLK);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(base + kb * LB_BLK, &tmW, w_full, kb * 64, 0, 0);      // Wq' rows
    }
    int stage = 0, nmb = 0, cur_b = -1;
    uint32_t phase = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      const int b = tile / tps;
      if (b != cur_b) {          // this sample's Mb into buffer nmb & 1
        cur_b = b;
        const int buf = nmb & 1;
        mbar_wait(mb_empty(buf), (uint32_t)((nmb >> 1) & 1) ^ 1u);
        if (leader) {
          mbar_expect_tx(mb_full(buf), 2 * MB_BLK);
          for (int kk = 0; kk < 2; ++kk)
            tma_load_3d(base + OFF_MB + (buf * 2 + kk) * MB_BLK, &tmM, mb_full(buf), kk * 64, b * C, 0);
        }
        ++nmb;
      }
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
    constexpr uint32_t idescQ = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescY = make_idesc_bf16(128, C, 0, 0);
    const uint64_t desc0 = make_smem_desc(base, 16, 1024);
    const uint32_t hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    auto mma_q = [&](int j) {
      const int slot = j & 1;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(xfull(stage), phase);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + (uint32_t)(slot * 256), lo0 + (uint32_t)((OFF_X + stage * LB_BLK + k * 32) >> 4), hi,
                           lo0 + (uint32_t)((kb * LB_BLK + k * 32) >> 4), hi, idescQ, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(xempty(stage));
        }
        __syncwarp();
        if (++stage == NSX) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(dq_full(slot));
      __syncwarp();
    };
    if (nt > 0) {
      mbar_wait(w_full, 0);
      tc_fence_after();
      mma_q(0);
      if (nt > 1) mma_q(1);
    }
    int nmb = 0, cur_b = -1;
    for (int j = 0; j < nt; ++j) {
      const int slot = j & 1;
      const uint32_t par = (uint32_t)((j >> 1) & 1);
      const int b = (tile0 + j) / tps;
      if (b != cur_b) {
        cur_b = b;
        mbar_wait(mb_full(nmb & 1), (uint32_t)((nmb >> 1) & 1));
        ++nmb;
      }
      const int buf = (nmb - 1) & 1;
      if (leader && j < 8) LA_TS(100 + j * 4);
      mbar_wait(q_full(slot), par);            // Qs[slot] written, Dq[slot] read
      mbar_wait(e_done(slot), par ^ 1u);       // Dy[slot] of tile j-2 read out
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + (uint32_t)(slot * 256 + 128),
                           lo0 + (uint32_t)((OFF_QS + (slot * 2 + kk) * LB_BLK + k * 32) >> 4), hi,
                           lo0 + (uint32_t)((OFF_MB + (buf * 2 + kk) * MB_BLK + k * 32) >> 4), hi, idescY,
                           (kk > 0 || k > 0) ? 1u : 0u);
        umma_commit(dy_full(slot));
        const bool last_of_sample = (j + 1 == nt) || ((tile0 + j + 1) / tps != b);
        if (last_of_sample) umma_commit(mb_empty(buf));
        if (j < 8) LA_TS(100 + j * 4 + 1);
      }
      __syncwarp();
      if (j + 2 < nt) mma_q(j + 2);
      if (leader && j < 8) LA_TS(100 + j * 4 + 2);
    }
  } else if (warp < 10) {
    // ===================== softmax warps: T1(j) for every tile =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    if (warp == 2 && lane == 0) LA_TS(0);
    for (int j = 0; j < nt; ++j) {
      const int slot = j & 1;
      const uint32_t par = (uint32_t)((j >> 1) & 1);
      const float myrn = p.rn[(size_t)(tile0 + j) * LB_TILE + r];
      uint8_t* qs = base_ptr + OFF_QS + (slot * 2 + half) * LB_BLK;    // columns [half*64, +64) = K block `half`
      if (warp == 2 && lane == 0 && j < 8) LA_TS(16 + j * 8);
      mbar_wait(dq_full(slot), par);
      mbar_wait(stg_free(slot), par ^ 1u);     // the epilogue of tile j-2 has copied its rows out of this slot
      tc_fence_after();
      if (warp == 2 && lane == 0 && j < 8) LA_TS(16 + j * 8 + 1);
      const float sc = myrn * LB_LOG2E;
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t qr[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(slot * 256 + half * 64 + hh * 32), qr);
        tmem_ld_wait();
        if (warp == 2 && lane == 0 && j < 8 && hh == 0) LA_TS(300 + j * 4);
        float mxv = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) mxv = fmaxf(mxv, __uint_as_float(qr[i]) * sc);
        if (warp == 2 && lane == 0 && j < 8 && hh == 0) LA_TS(300 + j * 4 + 1);
        float e[32], sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[i] = ex2f(fmaf(__uint_as_float(qr[i]), sc, -mxv));
          sum += e[i];
        }
        if (warp == 2 && lane == 0 && j < 8 && hh == 0) LA_TS(300 + j * 4 + 2);
        const float inv = 1.f / sum;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          sw128_store(qs, r, hh * 4 + g,
                      make_uint4(pack_bf16(e[8 * g] * inv, e[8 * g + 1] * inv), pack_bf16(e[8 * g + 2] * inv, e[8 * g + 3] * inv),
                                 pack_bf16(e[8 * g + 4] * inv, e[8 * g + 5] * inv), pack_bf16(e[8 * g + 6] * inv, e[8 * g + 7] * inv)));
        if (warp == 2 && lane == 0 && j < 8 && hh == 0) LA_TS(300 + j * 4 + 3);
      }
      tc_fence_before();
      if (warp == 2 && lane == 0 && j < 8) LA_TS(340 + j * 2);
      fence_proxy_async();
      if (warp == 2 && lane == 0 && j < 8) LA_TS(340 + j * 2 + 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full(slot));
      if (warp == 2 && lane == 0 && j < 8) LA_TS(16 + j * 8 + 2);
    }
  } else {
    // ===================== epilogue warps: E(j) for every tile =====================
    const int te = threadIdx.x - 320;
    const int q = warp & 3, half = (warp - 10) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const float sqrtC = sqrtf((float)C);
    for (int i = te; i < C; i += 256) {
      bias_s[i] = p.bout[i];
      gain_s[i] = p.gout[i] * sqrtC;
    }
    named_bar_sync(1, 256);
    for (int j = 0; j < nt; ++j) {
      const int slot = j & 1;
      const uint32_t par = (uint32_t)((j >> 1) & 1);
      const size_t row = (size_t)(tile0 + j) * LB_TILE + r;
      uint4 res[NCH / 8];
      {
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + row * p.x_ld + half * NCH);
#pragma unroll
        for (int c = 0; c < NCH / 8; ++c) res[c] = xr[c];
      }
      if (warp == 10 && lane == 0 && j < 8) LA_TS(16 + j * 8 + 3);
      mbar_wait(dy_full(slot), par);
      tc_fence_after();
      if (warp == 10 && lane == 0 && j < 8) LA_TS(16 + j * 8 + 4);
      // sum of squares of (y + bias) over this thread's channels.  C = 64: the 32 values stay in registers; C = 128:
      // a second sweep re-reads the accumulator (64 values per thread do not fit next to the softmax warps' budget)
      float ss = 0.f;
      float yk[32];
#pragma unroll
      for (int c = 0; c < NCH; c += 32) {
        uint32_t yr[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(slot * 256 + 128 + half * NCH + c), yr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = __uint_as_float(yr[i]) + bias_s[half * NCH + c + i];
          if (NCH == 32) yk[i] = v;
          ss = fmaf(v, v, ss);
        }
      }
      float* ssl = ssy + slot * 256;
      ssl[half * 128 + r] = ss;
      named_bar_sync(4 + q, 64);                 // the two warps that share these 32 pixel rows
      const float inv = rsqrtf(fmaxf(ssl[r] + ssl[128 + r], 1e-24f));
      // normalise, gain, + x; the rows are staged in this tile's operand slot (free: Dy is complete)
      uint8_t* stg = base_ptr + OFF_QS + slot * 2 * LB_BLK;
#pragma unroll
      for (int c = 0; c < NCH; c += 32) {
        if (NCH != 32) {
          uint32_t yr[32];
          tmem_ld32(tmem_base + tlane + (uint32_t)(slot * 256 + 128 + half * NCH + c), yr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) yk[i] = __uint_as_float(yr[i]) + bias_s[half * NCH + c + i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&res[c / 8 + g]);
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 xr2 = __bfloat1622float2(rh[i]);
            const int ch = half * NCH + c + g * 8 + 2 * i;
            o[i] = pack_bf16(fmaf(yk[g * 8 + 2 * i] * inv, gain_s[ch], xr2.x),
                             fmaf(yk[g * 8 + 2 * i + 1] * inv, gain_s[ch + 1], xr2.y));
          }
          const int cc = (half * NCH + c) / 8 + g;   // 16-byte chunk of the row; 64-channel blocks are 16 KiB apart
          sw128_store(stg + (cc >> 3) * LB_BLK, r, cc & 7, make_uint4(o[0], o[1], o[2], o[3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(e_done(slot));  // Dy[slot] may be overwritten
      if (warp == 10 && lane == 0 && j < 8) LA_TS(16 + j * 8 + 5);
      named_bar_sync(4 + q, 64);
      if (warp == 10 && lane == 0 && j < 8) LA_TS(16 + j * 8 + 6);
      {
        constexpr int CPR = C / 8;               // chunks per row
        const int t64 = half * 32 + lane;
        __nv_bfloat16* ybase = p.y + (size_t)(tile0 + j) * LB_TILE * p.y_ld;
#pragma unroll
        for (int i = 0; i < (32 * CPR) / 64; ++i) {
          const int id = t64 + 64 * i;
          const int rl = q * 32 + id / CPR, cc = id % CPR;
          const uint4 v = *reinterpret_cast<const uint4*>(stg + (cc >> 3) * LB_BLK + rl * 128 + (((cc & 7) ^ (rl & 7)) << 4));
          *reinterpret_cast<uint4*>(ybase + (size_t)rl * p.y_ld + cc * 8) = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(stg_free(slot));   // (generic-proxy reads done) the softmax warps may rewrite the slot
      if (warp == 10 && lane == 0 && j < 8) LA_TS(16 + j * 8 + 7);
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
  // device_sms(), not num_sms(): the workspace layout must not depend on SMs reserved for a collective at call time
  while (split * 2 <= 8 && tps % (split * 2) == 0 && (long long)B * split < 4ll * device_sms()) split *= 2;
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

#ifdef B200DM_PHASE_TIMING
extern "C" int b200dm_debug_set_la_timing_buf(long long* buf) {
  return cudaMemcpyToSymbol(g_la_tbuf, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int b200dm_pack_linattn_qkv(const float* w, const float* g, void* out, int32_t C, void* stream) {
  B200DM_REQUIRE(w && g && out && C > 0, B200DM_ERR_SHAPE, "pack_linattn_qkv: null argument");
  const int total = 384 * C;
  launch_k(la_pack_qkv_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, w, g, (__nv_bfloat16*)out, (int)C, total);
  count_launch();
  return check_launch("pack_linattn_qkv");
}

extern "C" int64_t b200dm_linattn_block_ws_floats(int32_t B, int32_t n, int32_t C) {
  if (B <= 0 || n <= 0 || n % LB_TILE || C <= 0) return 0;
  return (int64_t)B * la_split(B, n) * LB_WS + (int64_t)B * n + (int64_t)B * C * 64;
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

template <int KB>
static int la_launch(const b200dm_linattn_block_desc* d, cudaStream_t st) {
  constexpr int C = 64 * KB;
  const int split = la_split(d->B, d->n);
  const long long rows = (long long)d->B * d->n;
  LaParams p{};
  p.B = d->B; p.n = d->n; p.C = C; p.split = split; p.x_ld = d->x_ld; p.y_ld = d->y_ld;
  p.x = (const __nv_bfloat16*)d->x; p.y = (__nv_bfloat16*)d->y; p.wout = (const __nv_bfloat16*)d->wout;
  p.mem_kv = d->mem_kv; p.bout = d->bout; p.gout = d->gout;
  p.ws = d->ws;
  p.rn = d->ws + (size_t)d->B * split * LB_WS;
  p.mb = reinterpret_cast<__nv_bfloat16*>(p.rn + rows);
  CUtensorMap tmX128, tmX64, tmW, tmM;
  int rc = make_map2d(&tmX128, d->x, rows, C, d->x_ld, 128, "linattn_block x (128)");
  if (rc) return rc;
  rc = make_map2d(&tmX64, d->x, rows, C, d->x_ld, 64, "linattn_block x (64)");
  if (rc) return rc;
  rc = make_map2d(&tmW, d->wqkv, 384, C, C, 128, "linattn_block wqkv");
  if (rc) return rc;
  rc = make_map2d(&tmM, p.mb, (long long)d->B * C, 128, 128, C, "linattn_block Mb");
  if (rc) return rc;
  constexpr int smem0 = KB * LB_BLK + 4 * LB_BLK + 256 + 4096 + 1024;
  constexpr int smem1 = 2 * KB * LB_BLK + 6 * 8192 + 4 * LB_BLK + 256 + 1024;
  constexpr int smem2 = KB * LB_BLK + 4 * C * 128 + 4 * LB_BLK + (KB == 1 ? 4 : 3) * LB_BLK + 256 + 4096 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e0 = cudaFuncSetAttribute(la_kmax_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0);
    cudaError_t e1 = cudaFuncSetAttribute(la_ctx_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    cudaError_t e2 = cudaFuncSetAttribute(la_out_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    B200DM_REQUIRE(e0 == cudaSuccess && e1 == cudaSuccess && e2 == cudaSuccess, B200DM_ERR_CUDA,
                   "linattn_block_fwd: cudaFuncSetAttribute failed");
    configured = true;
  }
  const int items = d->B * split;
  const int grid_a = items < num_sms() ? items : num_sms();
  launch_k(la_kmax_kernel<KB>, grid_a, LB_THREADS, smem0, st, tmX128, tmW, p);
  launch_k(la_ctx_kernel<KB>, grid_a, LB_THREADS, smem1, st, tmX64, tmW, p);
  launch_k(la_mid_kernel, d->B, 256, 0, st, p);
  const long long tiles = rows / LB_TILE;
  const int grid_b = tiles < num_sms() ? (int)tiles : num_sms();
  launch_k(la_out_kernel<KB>, grid_b, LB_THREADS, smem2, st, tmX128, tmW, tmM, p);
  count_launch(4);
  return check_launch("linattn_block_fwd");
}

extern "C" int b200dm_linattn_block_fwd(const b200dm_linattn_block_desc* d, void* stream) {
  B200DM_REQUIRE(d != nullptr, B200DM_ERR_SHAPE, "linattn_block_fwd: null descriptor");
  B200DM_REQUIRE(b200dm_linattn_block_supported(d) == 1, B200DM_ERR_UNSUPPORTED,
                 "linattn_block_fwd: needs sm_100, bf16, n %% 128 == 0 (n=%d), C in {64,128} (C=%d), 16-byte aligned tensors",
                 d->n, d->C);
  B200DM_REQUIRE(d->x && d->y && d->wqkv && d->wout && d->bout && d->gout && d->mem_kv && d->ws, B200DM_ERR_SHAPE,
                 "linattn_block_fwd: null pointer");
  B200DM_REQUIRE(((uintptr_t)d->ws & 15) == 0, B200DM_ERR_SHAPE, "linattn_block_fwd: ws must be 16-byte aligned");
  return d->C == 64 ? la_launch<1>(d, (cudaStream_t)stream) : la_launch<2>(d, (cudaStream_t)stream);
}
