// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
//
//   D[128 pixels x N_TILE channels] = sum_{tap, 64-channel block}  A_tap[128 x 64] * W_tap[N_TILE x 64]^T
//
// A is never materialised (no im2col): each K-block is ONE 5-D TMA box over the NHWC activation whose
// (w, h) start coordinate is shifted by the filter tap; out-of-image elements are zero-filled by the
// TMA unit, which implements padding=1.  The pixel-unshuffle of Downsample (ddpm.py:100-104) is a
// strided 5-D view of the same tensor, and torch.cat inputs are channel slices of a wider buffer
// (x_ld).  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> +bias/+residual/+accumulate -> bf16 -> global).
//
// Serves the forward convs (ddpm.py:96,103,160,187,213-215,252-253,377,413) and, with the
// tap-reversed/transposed weight pack, their data gradients.
#include <stdlib.h>

#include "tc_common.cuh"

namespace b200dm {

using namespace tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    cudaGetLastError();
  }
  return fn;
}

bool tc_supported() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return major == 10 && get_encoder() != nullptr;
}

constexpr int TC_BM = 128;      // pixels per CTA tile  (UMMA M)
constexpr int TC_BK = 64;       // channels per K-block (128 B of bf16 = one swizzle row)
constexpr int TC_EPI_WARPS = 8; // two warps per TMEM lane quarter (each takes half of the columns)
constexpr int TC_THREADS = 96 + 32 * TC_EPI_WARPS;   // warp 0 TMA (A), warp 1 MMA, warps 2..9 epilogue, warp 10 TMA (B)
constexpr int TC_BWARP = 2 + TC_EPI_WARPS;            // the weight-tile producer warp
constexpr int WG_THREADS = 192;                      // wgrad kernel: 4 epilogue warps
constexpr int A_STAGE_BYTES = TC_BM * TC_BK * 2;     // 16 KiB
constexpr int OUT_STAGE_BYTES = TC_BM * 64 * 2;      // 16 KiB: [128 rows][64 bf16], 128-B swizzled

struct TcParams {
  int mode, ksize, taps, kblocks;  // kblocks = Cin / 64
  int H, W, bh, bn;                // box geometry: bw == W
  int Cout, Ncols;                 // Ncols = GEMM N (Cout, or 4*Cout in mode 2)
  int x_ld, y_ld;
  int m_tiles, n_tiles;
  long long M;                     // valid GEMM rows
  const __nv_bfloat16* y_read;     // accumulate source (== output tensor) or nullptr
  const __nv_bfloat16* res;
  int res_ld;
  const float* bias;
  float* gn_part;                  // GroupNorm partial sums [pixel slot][G][2] of the output, or nullptr
  int gn_groups, gn_seg;           // groups; pixels per slot (= lanes of a warp that belong to one image)
};

// GroupNorm statistics of a conv output, fused into the epilogue (deterministic: no atomics).
// A warp owns 32 pixels x 64 channels of a sub-tile.  S/Q hold this thread's (= pixel's) sum and sum of squares of
// the eight 8-channel chunks.  The 16 values are reduced over the `seg` lanes (32, or 16 when two 4x4 images share
// a warp) that belong to the same image by recursive halving - 16 shuffles instead of 16 x log2(seg) - after
// which every lane of a segment holds the total of ONE value and writes it:
//   part[(slot * C/8 + chunk) * 2 + {0: sum, 1: sum of squares}],  slot = pixel index / seg.
// The norm kernel adds the H*W/seg slots of an image and the chunks of a group.
__device__ __forceinline__ void gn_part_store(const float (&S)[8], const float (&Q)[8], int seg, int lane, bool ok,
                                              float* part, long long slot, int chunks_total, int chunk0) {
  float v[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] = S[k]; v[8 + k] = Q[k]; }
  int idx = 0;
  // halving steps: the lane keeps the half of its values selected by one of its index bits, sends the other half
#pragma unroll
  for (int n = 8, step = 0; n >= 1; n >>= 1, ++step) {
    const int m = (seg >> 1) >> step;                    // partner distance: 16, 8, 4, 2 (seg 32) / 8, 4, 2, 1 (seg 16)
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float keep = up ? v[i + n] : v[i];
      const float send = up ? v[i] : v[i + n];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
    idx = idx * 2 + (up ? 1 : 0);
  }
  if (seg == 32) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);   // the last pair of lanes holds the same value index
  const bool writer = seg == 32 ? (lane & 1) == 0 : true;
  if (ok && writer) part[(slot * chunks_total + chunk0 + (idx & 7)) * 2 + (idx >> 3)] = v[0];
}

// Persistent, warp-specialised implicit-GEMM convolution.  Each CTA walks tiles
// (tile = blockIdx.x + i*gridDim.x, N fastest so CTAs running together share the activation tile in L2):
//   warp 0      TMA producer : STAGES-deep ring of {A box 128x64, W box N_TILEx64}, runs ahead across tiles
//   warp 1      MMA issuer   : 4 x tcgen05.mma (128 x N_TILE x 16) per K-block into one of TWO TMEM
//                              accumulators, so tile i+1 accumulates while tile i is being drained
//   warps 2..9  epilogue     : tcgen05.ld -> +bias (+residual / +previous value) -> bf16 -> swizzled smem
//                              staging -> TMA store (coalesced 128-B rows, clipped at the tensor edge)
template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmY, const TcParams p) {
  constexpr int B_STAGE_BYTES = N_TILE * TC_BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int TMEM_COLS = 2 * N_TILE <= 128 ? 128 : 2 * N_TILE <= 256 ? 256 : 512;
  constexpr int SUBTILES = N_TILE / 64;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-B alignment
  const uint32_t out_stage = base + STAGES * STAGE_BYTES;        // 2 x 16 KiB staging for TMA stores
  const uint32_t bar_base = out_stage + 2 * OUT_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* out_stage_ptr = smem_raw + (out_stage - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));   // [Ncols]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.taps * p.kblocks;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), SUBTILES == 1 ? TC_EPI_WARPS / 2 : TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // everything above is on-chip setup; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Warp-uniform loop, one elected lane issues.  No divisions inside the K loop: taps and channel blocks are
    // nested counters (measured: decoding (tap, kc, dy, dx) from a flat index with runtime divisors cost the
    // single producer thread ~900 cycles per K block and starved the tensor pipe on the small-M layers).
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const int pad = p.ksize >> 1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * N_TILE;
      const long long first = (long long)m_tile * TC_BM;  // first pixel (linear NHW order)
      const int n_first = (int)(first / ((long long)p.H * p.W));
      const int y_first = (int)((first / p.W) % p.H);
      int dy = -pad, dx = -pad;                           // mode 0 filter offsets of the current tap
      for (int tap = 0; tap < p.taps; ++tap) {
        for (int kc = 0; kc < p.kblocks; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) {
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);     // activation box (here) + weight box (warp 10)
            const uint32_t a_dst = base + stage * STAGE_BYTES;
            if (p.mode == 1)   // unshuffle view (c' = p2*ld + c, ox, p1, oy, b)
              tma_load_5d(a_dst, &tmA, full_bar(stage), (tap & 1) * p.x_ld + kc * TC_BK, 0, tap >> 1, y_first,
                          n_first);
            else if (p.mode == 0)
              tma_load_5d(a_dst, &tmA, full_bar(stage), kc * TC_BK, dx, y_first + dy, n_first, 0);
            else if (p.mode == 3) {
              // nearest-2x upsample + 3x3 conv as four 2x2 convs over the SOURCE image: output phase (a, b) =
              // n0 / Cout reads source rows y + r + a - 1 and columns x + c + b - 1 for tap (r, c)
              const int ph = n0 / p.Cout;
              tma_load_5d(a_dst, &tmA, full_bar(stage), kc * TC_BK, (tap & 1) + (ph & 1) - 1,
                          y_first + (tap >> 1) + (ph >> 1) - 1, n_first, 0);
            } else
              tma_load_5d(a_dst, &tmA, full_bar(stage), kc * TC_BK, 0, y_first, n_first, 0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++dx > pad) { dx = -pad; ++dy; }
      }
    }
  } else if (warp == TC_BWARP) {
    // ===================== TMA producer, weight tiles =====================
    // A second issuing thread: one thread setting up both boxes of a K block (~45 dependent uniform-datapath
    // instructions) took longer than the 4 MMAs of an N=64 block (192 cycles) and starved the tensor pipe.
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * N_TILE;
      for (int tap = 0; tap < p.taps; ++tap) {
        for (int kc = 0; kc < p.kblocks; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader)
            tma_load_3d(base + stage * STAGE_BYTES + A_STAGE_BYTES, &tmB, full_bar(stage), kc * TC_BK, n0,
                        (p.mode == 2) ? 0 : tap);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop, one elected lane issues; descriptor = per-stage base + compile-time k offset
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, N_TILE, 0, 0);
    // K-major, SWIZZLE_128B: 8-row groups are 1024 B apart; +32 B per 16-element K step
    const uint64_t a_desc0 = make_smem_desc(base, 16, 1024);
    const uint64_t b_desc0 = make_smem_desc(base + A_STAGE_BYTES, 16, 1024);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
    const bool leader = elect_one();
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N_TILE);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint32_t so = (uint32_t)stage * (STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16_lohi(d_tmem, a_lo0 + so + (uint32_t)((k * 32) >> 4), a_hi,
                           b_lo0 + so + (uint32_t)((k * 32) >> 4), b_hi, idesc, (k > 0) ? 1u : (kb > 0 ? 1u : 0u));
          umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
        }
        __syncwarp();   // the other lanes wait here, not in the next try_wait (see conv3x3_halo_kernel)
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(tmem_full_bar(acc));  // accumulator complete
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Every warp is self-contained: it owns 32 pixel rows (its TMEM lane quarter) x the 64 channels of a
    // sub-tile, stages them as one 4 KiB swizzled block in its PRIVATE buffer and issues its own TMA store
    // (box = 32 pixels x 64 channels) - no CTA-wide barriers in the loop.  The two warps of a quarter
    // alternate: sub-tiles (N_TILE >= 128) or whole tiles (N_TILE = 64, accumulator = warp set).
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int wset = (warp - 2) >> 2;        // 0 / 1
    const int row = quarter * 32 + lane;
    const uint32_t my_stage = out_stage + (uint32_t)(warp - 2) * 4096u;
    uint8_t* my_row = out_stage_ptr + (warp - 2) * 4096 + lane * 128;
    // bias of all N columns, once per CTA
    for (int i = threadIdx.x - 64; i < p.Ncols; i += 32 * TC_EPI_WARPS)
      bias_s[i] = p.bias ? p.bias[p.mode == 3 ? i % p.Cout : i] : 0.f;      // mode 3: one bias per output phase
    named_bar_sync(1, 32 * TC_EPI_WARPS);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (SUBTILES == 1 && (it & 1) != wset) continue;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * N_TILE;
      const long long first = (long long)m_tile * TC_BM;
      const long long pix = first + row;
      const bool valid = pix < p.M;
      long long opix = pix;
      int co0 = n0, tap2 = 0;
      if (p.mode >= 2) {  // output pixel (2y+p1, 2x+p2) of the [2H,2W] tensor; one tap / phase per N tile
        tap2 = n0 / p.Cout;
        co0 = n0 - tap2 * p.Cout;
        const int ox = (int)(pix % p.W);
        const long long q = pix / p.W;
        const int oy = (int)(q % p.H);
        const long long b = q / p.H;
        opix = (b * (2 * p.H) + 2 * oy + (tap2 >> 1)) * (long long)(2 * p.W) + 2 * ox + (tap2 & 1);
      }
      // store coordinates of this warp's 32-pixel slice
      const unsigned slice = (unsigned)first + 32u * (unsigned)quarter;
      const unsigned hw = (unsigned)(p.H * p.W);
      const int sn = (int)(slice / hw);
      const unsigned srem = slice - (unsigned)sn * hw;
      const int sy = (int)(srem / (unsigned)p.W), sx = (int)(srem - (unsigned)sy * (unsigned)p.W);
      // residual / accumulate sources are fetched BEFORE waiting for the accumulator, so their L2 latency
      // hides behind the MMAs
      uint4 src_res[4], src_acc[4];
      auto prefetch_src = [&](int c) {
        if (valid && p.res) {
          const uint4* rr = reinterpret_cast<const uint4*>(p.res + opix * p.res_ld + co0 + c);
#pragma unroll
          for (int g = 0; g < 4; ++g) src_res[g] = rr[g];
        }
        if (valid && p.y_read) {
          const uint4* yr = reinterpret_cast<const uint4*>(p.y_read + opix * p.y_ld + co0 + c);
#pragma unroll
          for (int g = 0; g < 4; ++g) src_acc[g] = yr[g];
        }
      };
      const int s_first = SUBTILES == 1 ? 0 : wset;
      prefetch_src(s_first * 64);
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int s = s_first; s < SUBTILES; s += 2) {
        // the private staging block was handed to the TMA engine one sub-tile ago
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        float gS[8], gQ[8];
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int c = s * 64 + hh * 32;     // column offset inside the N tile
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE + c), r);
          tmem_ld_wait();
          if (hh == 1 && s + 2 >= SUBTILES) {   // last read of this accumulator by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
          }
          float v[32];
          {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + n0 + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b4[j];
              v[4 * j] = __uint_as_float(r[4 * j]) + bb.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
            }
          }
          if (p.gn_part) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float a = 0.f, q = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float t = valid ? v[8 * k + j] : 0.f;
                a += t;
                q = fmaf(t, t, q);
              }
              if (hh == 0) { gS[k] = a; gQ[k] = q; } else { gS[4 + k] = a; gQ[4 + k] = q; }   // static register indices
            }
          }
          if (valid && p.res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&src_res[g]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                v[g * 8 + 2 * j] += f.x;
                v[g * 8 + 2 * j + 1] += f.y;
              }
            }
          }
          if (valid && p.y_read) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&src_acc[g]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                v[g * 8 + 2 * j] += f.x;
                v[g * 8 + 2 * j + 1] += f.y;
              }
            }
          }
          // next 32 columns this warp will need: second half of this sub-tile, or its next sub-tile
          if (hh == 0) prefetch_src(c + 32);
          else if (s + 2 < SUBTILES) prefetch_src((s + 2) * 64);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
            const int chunk = (hh * 4 + g) ^ (lane & 7);   // 128-B swizzle: 16-B chunk index XOR (row mod 8)
            *reinterpret_cast<uint4*>(my_row + chunk * 16) = u;
          }
        }
        if (p.gn_part)
          gn_part_store(gS, gQ, p.gn_seg, lane, valid, p.gn_part, (long long)((unsigned)first + 32u * (unsigned)quarter + (unsigned)lane) / p.gn_seg,
                        p.Cout / 8, (co0 + s * 64) / 8);
        fence_proxy_async();                  // generic-proxy smem writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if (p.mode >= 2)
            tma_store_5d(&tmY, my_stage, (tap2 & 1) * p.y_ld + co0 + s * 64, sx, tap2 >> 1, sy, sn);
          else
            tma_store_5d(&tmY, my_stage, co0 + s * 64, sx, sy, sn, 0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================
// 3x3 convolution with an on-chip halo: the implicit GEMM above re-reads the activation tile once per
// filter tap (9x) through L2, which is what bounds it (measured ~6.5 TB/s L2->SM).  Here a tile is
// 8 x 16 output pixels; ONE TMA box {64 ch, 16 cols, 18 rows} brings the tile plus its 1-pixel halo into
// shared memory (row pitch 16 pixels so that image rows stay 8-row-group aligned), and the nine taps are
// nine *views* of that buffer: the UMMA A descriptor starts (dx + 16*dy) rows into the buffer, uses a
// stride of 2048 B between 8-row groups (= next image row); the 128-B swizzle phase follows the absolute
// shared-memory address (measured), so the shifted rows read back exactly what the TMA unit wrote.  For 64->64 layers the whole 3x3 weight
// (72 KiB) stays resident in shared memory for the life of the persistent CTA.
// =====================================================================================================
// Row pitch of the staged halo = the box width: 10 pixels (tile 8 + 2 halo).  The 8-row groups of the A operand are
// then 1280 B apart — not a multiple of the 1024-B swizzle atom — which is fine because both the TMA unit and the
// tensor core derive the swizzle phase from the absolute shared-memory address (measured, see conv3x3_halo below).
// Against the 16-pixel pitch of round 1 this moves 23 KB instead of 36 KB per tile through L2 -> SM, the link that
// bounds the 64-channel layers (one 128-pixel tile every 0.88 us per SM = 6 TB/s at 36 KB).  -DB200DM_HALO_W=16
// restores the padded pitch for A/B runs.
#ifndef B200DM_HALO_W
#define B200DM_HALO_W 10
#endif
constexpr int HALO_W = B200DM_HALO_W, HALO_H = 18;
constexpr int HALO_TX = HALO_W * HALO_H * 128;                 // bytes one halo box brings (mbarrier expect_tx)
constexpr int HALO_BYTES = (HALO_TX + 1023) / 1024 * 1024;     // buffer stride: 23 KiB (36 KiB at pitch 16)

// -DB200DM_PHASE_TIMING (scripts/phase_timing.py, a separate libb200dm_timing.so): CTA 0 of the halo kernel records
// SM-clock stamps of its pipeline phases into a global buffer.  Not compiled into the production library.
#ifdef B200DM_PHASE_TIMING
__device__ long long* g_tbuf = nullptr;
#define TSTAMP(slot) do { if (g_tbuf && blockIdx.x == 0) g_tbuf[(slot)] = clock64(); } while (0)
#else
#define TSTAMP(slot) do { } while (0)
#endif

struct TcHaloParams {
  int kblocks;                 // Cin / 64
  int H, W, tiles_x, tiles_y;  // 8x16 tiles per image
  int Cout, y_ld;
  int m_tiles, n_tiles;
  const __nv_bfloat16* y_read;
  const __nv_bfloat16* res;
  int res_ld;
  const float* bias;
  float* gn_part;              // GroupNorm partial sums [pixel slot][G][2] of the output, or nullptr
  int gn_groups;
  int dbg;                     // -DB200DM_HALO_DEBUG builds only (scripts/halo_debug.py): bit 0 skip epilogue stores,
                               // 1 skip A loads, 2 skip MMAs, 3 return at once, 4 no tiles, 5 no resident-weight load
};
#ifdef B200DM_HALO_DEBUG
#define HALO_DBG(p) ((p).dbg)
#else
#define HALO_DBG(p) 0          // the production kernel carries none of the experiment branches
#endif

template <int N_TILE, int A_BUFS, int B_STAGES, bool B_RESIDENT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const TcHaloParams p) {
  constexpr int B_BYTES = N_TILE * TC_BK * 2;
  constexpr int TMEM_COLS = 2 * N_TILE <= 128 ? 128 : 2 * N_TILE <= 256 ? 256 : 512;
  constexpr int SUBTILES = N_TILE / 64;
  pdl_launch_dependents();
  if (threadIdx.x == 0) TSTAMP(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + A_BUFS * HALO_BYTES;
  const int b_slots = B_RESIDENT ? 9 * p.kblocks : B_STAGES;
  const uint32_t out_stage = b_base + b_slots * B_BYTES;
  const uint32_t bar_base = out_stage + 2 * OUT_STAGE_BYTES;
  auto afull = [&](int a) { return bar_base + 8u * a; };
  auto aempty = [&](int a) { return bar_base + 8u * (A_BUFS + a); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * A_BUFS + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * A_BUFS + B_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + 2 + a); };
  const uint32_t bres_bar = bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* out_stage_ptr = smem_raw + (out_stage - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));   // [Cout]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (HALO_DBG(p) & 8) return;
  const int num_tiles = (HALO_DBG(p) & 16) ? 0 : p.m_tiles * p.n_tiles;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    for (int a = 0; a < A_BUFS; ++a) { mbar_init(afull(a), 1); mbar_init(aempty(a), 1); }
    for (int s2 = 0; s2 < B_STAGES; ++s2) { mbar_init(bfull(s2), 1); mbar_init(bempty(s2), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), SUBTILES == 1 ? TC_EPI_WARPS / 2 : TC_EPI_WARPS);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) TSTAMP(1);
  pdl_wait();   // everything above is on-chip setup; global memory is touched only below
  if (threadIdx.x == 0) TSTAMP(2);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (B_RESIDENT && !(HALO_DBG(p) & 32)) {   // the whole [9][N_TILE][Cin] weight, once
        mbar_expect_tx(bres_bar, (uint32_t)(9 * p.kblocks * B_BYTES));
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < p.kblocks; ++kc)
            tma_load_3d(b_base + (tap * p.kblocks + kc) * B_BYTES, &tmB, bres_bar, kc * TC_BK, 0, tap);
      }
      int ab = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int b = m_tile / tiles_per_img, rem = m_tile - b * tiles_per_img;
        const int y0 = (rem / p.tiles_x) * 16, x0 = (rem % p.tiles_x) * 8;
        for (int kc = 0; kc < p.kblocks; ++kc) {
          mbar_wait(aempty(ab), aph ^ 1u);
          if (kc == 0 && (tile - (int)blockIdx.x) / (int)gridDim.x < 16) TSTAMP(150 + (tile - (int)blockIdx.x) / (int)gridDim.x);
          if (HALO_DBG(p) & 2) {
            mbar_arrive(afull(ab));
          } else {
            mbar_expect_tx(afull(ab), HALO_TX);
            tma_load_5d(a_base + ab * HALO_BYTES, &tmA, afull(ab), kc * TC_BK, x0 - 1, y0 - 1, b, 0);
          }
          if (++ab == A_BUFS) { ab = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == TC_BWARP) {
    // ===================== TMA producer, weight tiles (streamed variant) =====================
    if (lane == 0 && !B_RESIDENT) {
      int bs = 0;
      uint32_t bph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * N_TILE;
        for (int kc = 0; kc < p.kblocks; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bempty(bs), bph ^ 1u);
            mbar_expect_tx(bfull(bs), B_BYTES);
            tma_load_3d(b_base + bs * B_BYTES, &tmB, bfull(bs), kc * TC_BK, n0, tap);
            if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (warp-uniform control flow and operands, so ptxas keeps the descriptors in
    // uniform registers); one elected lane issues.  Descriptors: base (lo, hi) per operand buffer + a
    // compile-time constant per (tap, k-step) added to the start-address field.
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, N_TILE, 0, 0);
    const uint64_t a_desc0 = make_smem_desc(a_base, 16, HALO_W * 128);
    const uint64_t b_desc0 = make_smem_desc(b_base, 16, 1024);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
    const bool leader = elect_one();
    int ab = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0;
    if (B_RESIDENT && !(HALO_DBG(p) & 32)) mbar_wait(bres_bar, 0);
    // `pre`: the barriers of the coming tile (its accumulator, its first halo) were already waited for while the last
    // MMAs of the previous tile were still queued, so the tensor pipe does not drain between tiles (measured with the
    // phase-timing build: ~0.4 us of waits per 1.5 us tile on the 64-channel layers otherwise)
    bool pre = false;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if (!pre) {
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
      }
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N_TILE);
      const int ti_ = (tile - (int)blockIdx.x) / (int)gridDim.x;
      if (leader && ti_ < 16) TSTAMP(200 + ti_);          // accumulator free
      for (int kc = 0; kc < p.kblocks; ++kc) {
        // (no tcgen05 fence after a TMA-full wait: the operands were written by the async proxy and complete_tx
        // orders them for the MMA; the fence is for barriers signalled by other threads' tcgen05 operations)
        if (!(pre && kc == 0)) mbar_wait(afull(ab), aph);
        if (leader && kc == 0 && ti_ < 16) TSTAMP(10 + 2 * ti_);   // first halo of the tile has landed
        const uint32_t a_lo = a_lo0 + (uint32_t)ab * (HALO_BYTES >> 4);
        if (B_RESIDENT) {
          const bool more = kc == p.kblocks - 1 && tile + (int)gridDim.x < num_tiles;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint32_t b_lo = b_lo0 + (uint32_t)(tap * p.kblocks + kc) * (B_BYTES >> 4);
            if (tap == 8) {
              pre = more;
              if (more) {     // the coming tile's barriers, behind the queue of the 32 MMAs just issued
                const int nacc = acc ^ 1, nab = ab + 1 == A_BUFS ? 0 : ab + 1;
                mbar_wait(tmem_empty_bar(nacc), (nacc == 0 ? acc_phase ^ 1u : acc_phase) ^ 1u);
                tc_fence_after();
                mbar_wait(afull(nab), nab == 0 ? aph ^ 1u : aph);
              }
            }
            if (leader && !(HALO_DBG(p) & 4)) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                umma_bf16_lohi(d_tmem, a_lo + (uint32_t)(((dx + HALO_W * dy) * 128 + k * 32) >> 4), a_hi,
                               b_lo + (uint32_t)((k * 32) >> 4), b_hi, idesc,
                               (tap > 0 || k > 0) ? 1u : (kc > 0 ? 1u : 0u));
            }
            // the 31 other lanes wait HERE for the issuing lane: left to run ahead they would sit in the next
            // mbarrier try_wait of the loop, and the warp would alternate between that spin and the MMA issue
            __syncwarp();
          }
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            // tap window: rows (tx + dx) + pitch*(ty + dy) of the halo buffer; 8-row groups = image rows
            const int dy = tap / 3, dx = tap - dy * 3;
            mbar_wait(bfull(bs), bph);
            const uint32_t b_lo = b_lo0 + (uint32_t)bs * (B_BYTES >> 4);
            if (leader) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                umma_bf16_lohi(d_tmem, a_lo + (uint32_t)(((dx + HALO_W * dy) * 128 + k * 32) >> 4), a_hi,
                               b_lo + (uint32_t)((k * 32) >> 4), b_hi, idesc,
                               (tap > 0 || k > 0) ? 1u : (kc > 0 ? 1u : 0u));
              umma_commit(bempty(bs));
            }
            __syncwarp();
            if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
          }
        }
        if (leader) umma_commit(aempty(ab));
        if (++ab == A_BUFS) { ab = 0; aph ^= 1u; }
      }
      if (leader) umma_commit(tmem_full_bar(acc));
      if (leader && ti_ < 16) TSTAMP(11 + 2 * ti_);       // all MMAs of the tile issued
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // self-contained warps (see conv_tc_kernel): 32 rows = 4 image rows x 8 pixels of the tile, private 4 KiB
    // staging block, own TMA store (box 64 ch x 8 x 4); the two warps of a quarter alternate sub-tiles / tiles
    const int quarter = warp & 3;
    const int wset = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int tx = row & 7, ty = row >> 3;
    const uint32_t my_stage = out_stage + (uint32_t)(warp - 2) * 4096u;
    uint8_t* my_row = out_stage_ptr + (warp - 2) * 4096 + lane * 128;
    for (int i = threadIdx.x - 64; i < p.Cout; i += 32 * TC_EPI_WARPS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    named_bar_sync(1, 32 * TC_EPI_WARPS);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (SUBTILES == 1 && (it & 1) != wset) continue;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * N_TILE;
      const int b = m_tile / tiles_per_img, rem = m_tile - b * tiles_per_img;
      const int y0 = (rem / p.tiles_x) * 16, x0 = (rem % p.tiles_x) * 8;
      const long long opix = ((long long)b * p.H + y0 + ty) * p.W + x0 + tx;
      // residual / accumulate sources: fetched before the accumulator wait (latency hidden behind the MMAs)
      uint4 src_res[4], src_acc[4];
      auto prefetch_src = [&](int c) {
        if (p.res) {
          const uint4* rr = reinterpret_cast<const uint4*>(p.res + opix * p.res_ld + n0 + c);
#pragma unroll
          for (int g = 0; g < 4; ++g) src_res[g] = rr[g];
        }
        if (p.y_read) {
          const uint4* yr = reinterpret_cast<const uint4*>(p.y_read + opix * p.y_ld + n0 + c);
#pragma unroll
          for (int g = 0; g < 4; ++g) src_acc[g] = yr[g];
        }
      };
      const int s_first = SUBTILES == 1 ? 0 : wset;
      prefetch_src(s_first * 64);
      const int eb_ = (quarter == 0 && lane == 0 && (it >> 1) < 8) ? 50 + wset * 50 + 4 * (it >> 1) : -1;
      if (eb_ >= 0) TSTAMP(eb_);                            // starts waiting for the accumulator
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      if (eb_ >= 0) TSTAMP(eb_ + 1);                        // accumulator complete
#pragma unroll 1
      for (int s2 = s_first; s2 < SUBTILES; s2 += 2) {
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        if (eb_ >= 0 && s2 == s_first) TSTAMP(eb_ + 2);     // staging block free
        float gS[8], gQ[8];
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int c = s2 * 64 + hh * 32;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE + c), r);
          tmem_ld_wait();
          if (hh == 1 && s2 + 2 >= SUBTILES) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
          }
          if (HALO_DBG(p) & 1) continue;
          float v[32];
          {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + n0 + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b4[j];
              v[4 * j] = __uint_as_float(r[4 * j]) + bb.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
            }
          }
          if (p.gn_part) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float a = 0.f, q = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                a += v[8 * k + j];
                q = fmaf(v[8 * k + j], v[8 * k + j], q);
              }
              if (hh == 0) { gS[k] = a; gQ[k] = q; } else { gS[4 + k] = a; gQ[4 + k] = q; }   // static register indices
            }
          }
          if (p.res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&src_res[g]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                v[g * 8 + 2 * j] += f.x;
                v[g * 8 + 2 * j + 1] += f.y;
              }
            }
          }
          if (p.y_read) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&src_acc[g]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                v[g * 8 + 2 * j] += f.x;
                v[g * 8 + 2 * j + 1] += f.y;
              }
            }
          }
          if (hh == 0) prefetch_src(c + 32);
          else if (s2 + 2 < SUBTILES) prefetch_src((s2 + 2) * 64);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
            const int chunk = (hh * 4 + g) ^ (lane & 7);
            *reinterpret_cast<uint4*>(my_row + chunk * 16) = u;
          }
        }
        if (HALO_DBG(p) & 1) continue;
        if (p.gn_part)
          gn_part_store(gS, gQ, 32, lane, true, p.gn_part, (long long)m_tile * 4 + quarter, p.Cout / 8, (n0 + s2 * 64) / 8);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_5d(&tmY, my_stage, n0 + s2 * 64, x0, y0 + 4 * quarter, b, 0);
          tma_store_commit();
        }
        if (eb_ >= 0) TSTAMP(eb_ + 3);                      // store issued
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    if (warp == 2 && lane == 0) TSTAMP(3);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TSTAMP(4);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what) {
  EncodeTiledFn enc = get_encoder();
  B200DM_REQUIRE(enc != nullptr, B200DM_ERR_UNSUPPORTED, "%s: cuTensorMapEncodeTiled unavailable", what);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims,
                   strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200DM_REQUIRE(r == CUDA_SUCCESS, B200DM_ERR_CUDA,
                 "%s: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]",
                 what, (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                 (unsigned long long)dims[2], rank > 3 ? (unsigned long long)dims[3] : 0ULL,
                 rank > 4 ? (unsigned long long)dims[4] : 0ULL, box[0], box[1], box[2],
                 rank > 3 ? box[3] : 0u, rank > 4 ? box[4] : 0u);
  return B200DM_OK;
}

// Activation map shared by the forward/dgrad and wgrad kernels.  Box = {64 ch, W, bh, bn} pixels.
int make_act_map(CUtensorMap* m, int mode, const void* x, int x_ld, int C, int B, int H, int W, int bh,
                 int bn, const char* what) {
  const cuuint64_t e = 2;  // bytes per element
  if (mode == 1) {
    // input [B, 2H, 2W, ld] viewed as (c' = p2*ld + c, ox, p1, oy, b)
    cuuint64_t dims[5] = {(cuuint64_t)x_ld + C, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[4] = {2ull * x_ld * e, 2ull * W * x_ld * e, 4ull * W * x_ld * e,
                         4ull * H * W * x_ld * e};
    cuuint32_t box[5] = {TC_BK, (cuuint32_t)W, 1, (cuuint32_t)bh, (cuuint32_t)bn};
    return encode_map(m, x, 5, dims, str, box, what);
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 1};
  cuuint64_t str[4] = {(cuuint64_t)x_ld * e, (cuuint64_t)W * x_ld * e, (cuuint64_t)H * W * x_ld * e,
                       (cuuint64_t)B * H * W * x_ld * e};
  cuuint32_t box[5] = {TC_BK, (cuuint32_t)W, (cuuint32_t)bh, (cuuint32_t)bn, 1};
  return encode_map(m, x, 5, dims, str, box, what);
}

// Output map of the per-warp epilogue: box = 32 consecutive pixels (NHW order) x 64 channels.
//   W >= 32: 32 pixels of one image row; otherwise 32/W rows (of one image) or 32/(W*H) whole images.
static int make_out_map32(CUtensorMap* m, int mode, const void* y, int y_ld, int C, int B, int H, int W,
                          const char* what) {
  const cuuint64_t e = 2;
  const int bw = W >= 32 ? 32 : W;
  const int rows = 32 / bw;
  int bh, bn;
  if (rows <= H) { bh = rows; bn = 1; } else { bh = H; bn = rows / H; }
  B200DM_REQUIRE(W % bw == 0 && H % bh == 0 && bh * bn * bw == 32, B200DM_ERR_UNSUPPORTED,
                 "%s: cannot cut %dx%d images into 32-pixel store boxes", what, H, W);
  if (mode == 1) {   // [B, 2H, 2W, ld] viewed as (c' = p2*ld + c, ox, p1, oy, b)
    cuuint64_t dims[5] = {(cuuint64_t)y_ld + C, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[4] = {2ull * y_ld * e, 2ull * W * y_ld * e, 4ull * W * y_ld * e, 4ull * H * W * y_ld * e};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, 1, (cuuint32_t)bh, (cuuint32_t)bn};
    return encode_map(m, y, 5, dims, str, box, what);
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 1};
  cuuint64_t str[4] = {(cuuint64_t)y_ld * e, (cuuint64_t)W * y_ld * e, (cuuint64_t)H * W * y_ld * e,
                       (cuuint64_t)B * H * W * y_ld * e};
  cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, 1};
  return encode_map(m, y, 5, dims, str, box, what);
}

int tile_geometry(int H, int W, int* bh, int* bn, const char* what) {
  B200DM_REQUIRE(W >= 4 && W <= 128 && (W & (W - 1)) == 0, B200DM_ERR_UNSUPPORTED,
                 "%s: W=%d must be a power of two in [4,128]", what, W);
  int rows = TC_BM / W;  // image rows per tile if H is large enough
  if (rows <= H) {
    B200DM_REQUIRE(H % rows == 0, B200DM_ERR_UNSUPPORTED, "%s: H=%d not a multiple of %d", what, H, rows);
    *bh = rows;
    *bn = 1;
  } else {
    B200DM_REQUIRE(rows % H == 0, B200DM_ERR_UNSUPPORTED, "%s: H=%d does not divide %d", what, H, rows);
    *bh = H;
    *bn = rows / H;
  }
  return B200DM_OK;
}

template <int N_TILE, int STAGES>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                     const TcParams& p, cudaStream_t st) {
  const int smem = STAGES * (A_STAGE_BYTES + N_TILE * TC_BK * 2) + 2 * OUT_STAGE_BYTES + 1024 + 256 + p.Ncols * 4;
  B200DM_REQUIRE(smem <= 227 * 1024, B200DM_ERR_UNSUPPORTED, "conv_tc: %d B of shared memory", smem);
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<N_TILE, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  launch_k(conv_tc_kernel<N_TILE, STAGES>, grid, TC_THREADS, smem, st, tmA, tmB, tmY, p);
  count_launch();
  return check_launch("conv_tc");
}

template <int N_TILE, int A_BUFS, int B_STAGES, bool B_RESIDENT>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                       const TcHaloParams& p, cudaStream_t st) {
  const int b_slots = B_RESIDENT ? 9 * p.kblocks : B_STAGES;
  const int smem = A_BUFS * HALO_BYTES + b_slots * (N_TILE * TC_BK * 2) + 2 * OUT_STAGE_BYTES + 1024 + 512 + p.Cout * 4;
  B200DM_REQUIRE(smem <= 227 * 1024, B200DM_ERR_UNSUPPORTED, "conv3x3_halo: %d B of shared memory", smem);
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel<N_TILE, A_BUFS, B_STAGES, B_RESIDENT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "conv3x3_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  launch_k(conv3x3_halo_kernel<N_TILE, A_BUFS, B_STAGES, B_RESIDENT>, grid, TC_THREADS, smem, st, tmA, tmB, tmY, p);
  count_launch();
  return check_launch("conv3x3_halo");
}

static bool halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200DM_NO_HALO");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// 3x3 'same' conv on images of at least 16x16: halo kernel
static int conv3x3_halo(const b200dm_conv_desc* d, cudaStream_t st) {
  TcHaloParams p{};
  p.kblocks = d->Cin / TC_BK;
  p.H = d->H; p.W = d->W; p.tiles_x = d->W / 8; p.tiles_y = d->H / 16;
  p.Cout = d->Cout; p.y_ld = d->y_ld;
  p.m_tiles = d->B * p.tiles_x * p.tiles_y;
  p.y_read = d->accumulate ? (const __nv_bfloat16*)d->y : nullptr;
  p.res = (const __nv_bfloat16*)d->res; p.res_ld = d->res_ld;
  p.bias = d->bias;
  p.gn_part = d->gn_part; p.gn_groups = d->gn_groups;
  {
    // MEASURED on B200 (scripts/halo_debug.py): the tensor core derives the 128-B swizzle phase from the
    // absolute shared-memory address bits [7,10), exactly like the TMA unit that wrote the tile, so a
    // window that starts at an arbitrary 128-B row needs base_offset = 0 in the UMMA descriptor.
#ifdef B200DM_HALO_DEBUG
    const char* e = getenv("B200DM_HALO_DEBUG");
    p.dbg = e ? atoi(e) : 0;
#endif
  }
  const int sms = num_sms();
  int n_tile = 64;
  if (d->Cout % 128 == 0) {
    // N = 128 unless that leaves most of the machine idle
    const long long t128 = (long long)p.m_tiles * (d->Cout / 128);
    if (t128 >= sms / 2) n_tile = 128;
  }
  p.n_tiles = d->Cout / n_tile;
  const bool resident = (d->Cout == 64 && d->Cin == 64);

  CUtensorMap tmA, tmB, tmY;
  const cuuint64_t e = 2;
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B, 1};
    cuuint64_t str[4] = {(cuuint64_t)d->x_ld * e, (cuuint64_t)d->W * d->x_ld * e,
                         (cuuint64_t)d->H * d->W * d->x_ld * e, (cuuint64_t)d->B * d->H * d->W * d->x_ld * e};
    cuuint32_t box[5] = {TC_BK, HALO_W, HALO_H, 1, 1};
    int rc = encode_map(&tmA, d->x, 5, dims, str, box, "conv3x3_halo A");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B, 1};
    cuuint64_t str[4] = {(cuuint64_t)d->y_ld * e, (cuuint64_t)d->W * d->y_ld * e,
                         (cuuint64_t)d->H * d->W * d->y_ld * e, (cuuint64_t)d->B * d->H * d->W * d->y_ld * e};
    cuuint32_t box[5] = {64, 8, 4, 1, 1};      // one epilogue warp: 4 image rows x 8 pixels x 64 channels
    int rc = encode_map(&tmY, d->y, 5, dims, str, box, "conv3x3_halo Y");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Cout, 9};
    cuuint64_t str[2] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cout * d->Cin * 2};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)n_tile, 1};
    int rc = encode_map(&tmB, d->w, 3, dims, str, box, "conv3x3_halo B");
    if (rc) return rc;
  }
  // halo buffers in flight (23 KiB each): as many as shared memory holds next to the weights and the store staging
  if (resident) return launch_halo<64, HALO_W <= 10 ? 5 : 3, 1, true>(tmA, tmB, tmY, p, st);
  if (n_tile == 128) return launch_halo<128, HALO_W <= 10 ? 4 : 2, 6, false>(tmA, tmB, tmY, p, st);
  return launch_halo<64, HALO_W <= 10 ? 5 : 3, 6, false>(tmA, tmB, tmY, p, st);
}

int conv_fwd_tc(const b200dm_conv_desc* d, void* stream) {
  B200DM_REQUIRE(tc_supported(), B200DM_ERR_UNSUPPORTED, "conv_fwd(tc): needs an sm_100 device and a TMA-capable driver");
  B200DM_REQUIRE(d->Cin % TC_BK == 0, B200DM_ERR_SHAPE, "conv_fwd(tc): Cin=%d must be a multiple of 64", d->Cin);
  B200DM_REQUIRE(d->Cout % 64 == 0, B200DM_ERR_SHAPE, "conv_fwd(tc): Cout=%d must be a multiple of 64", d->Cout);
  B200DM_REQUIRE(d->x_ld % 8 == 0 && d->y_ld % 8 == 0 && (!d->res || d->res_ld % 8 == 0), B200DM_ERR_SHAPE,
                 "conv_fwd(tc): ld must be a multiple of 8 elements");
  B200DM_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->y & 15) == 0 && ((uintptr_t)d->w & 15) == 0 &&
                     ((uintptr_t)d->res & 15) == 0,
                 B200DM_ERR_SHAPE, "conv_fwd(tc): pointers must be 16-byte aligned");
  if (d->gn_part) {
    const int gs = d->gn_groups > 0 ? d->Cout / d->gn_groups : 0;
    B200DM_REQUIRE(d->mode == 0 && d->gn_groups > 0 && d->Cout % d->gn_groups == 0 &&
                       gs % 8 == 0 && !d->res && !d->accumulate,
                   B200DM_ERR_UNSUPPORTED, "conv_fwd(tc): fused GroupNorm statistics need mode 0, no residual and "
                   "a multiple of 8 channels per group (Cout=%d groups=%d)", d->Cout, d->gn_groups);
  }
  const int ksize = d->mode == 0 ? d->ksize : 1;
  if (d->mode == 0 && ksize == 3 && d->W >= 16 && d->W % 8 == 0 && d->H % 16 == 0 && halo_enabled())
    return conv3x3_halo(d, (cudaStream_t)stream);
  int bh, bn;
  int rc = tile_geometry(d->H, d->W, &bh, &bn, "conv_fwd(tc)");
  if (rc) return rc;
  TcParams p{};
  p.mode = d->mode; p.ksize = ksize;
  p.taps = d->mode == 0 ? ksize * ksize : (d->mode == 1 || d->mode == 3 ? 4 : 1);
  p.kblocks = d->Cin / TC_BK;
  p.H = d->H; p.W = d->W; p.bh = bh; p.bn = bn;
  p.Cout = d->Cout;
  p.Ncols = d->mode >= 2 ? 4 * d->Cout : d->Cout;
  p.x_ld = d->x_ld; p.y_ld = d->y_ld;
  p.M = (long long)d->B * d->H * d->W;
  p.y_read = d->accumulate ? (const __nv_bfloat16*)d->y : nullptr;
  p.res = (const __nv_bfloat16*)d->res; p.res_ld = d->res_ld;
  p.bias = d->bias;
  p.gn_part = d->gn_part; p.gn_groups = d->gn_groups;
  p.gn_seg = d->H * d->W >= 32 ? 32 : d->H * d->W;
  p.m_tiles = (int)((p.M + TC_BM - 1) / TC_BM);

  // N tile: maximise (MMA efficiency of the tile shape) x (fill of the last wave of the persistent grid)
  const int sms = num_sms();
  int n_tile = 64;
  double best = -1.0;
  const int cand[3] = {256, 128, 64};
  const double eff[3] = {1.0, 0.85, 0.6};
  for (int i = 0; i < 3; ++i) {
    if (d->Cout % cand[i]) continue;
    const long long tiles = (long long)p.m_tiles * (p.Ncols / cand[i]);
    const long long waves = (tiles + sms - 1) / sms;
    const double score = eff[i] * (double)tiles / (double)(waves * sms);
    if (score > best) { best = score; n_tile = cand[i]; }
  }
  p.n_tiles = p.Ncols / n_tile;

  CUtensorMap tmA, tmB, tmY;
  rc = make_act_map(&tmA, d->mode == 1 ? 1 : 0, d->x, d->x_ld, d->Cin, d->B, d->H, d->W, bh, bn, "conv_fwd(tc) A");
  if (rc) return rc;
  // output: NHWC tensor (modes 0/1) or the unshuffle view of the [2H,2W] tensor (mode 2)
  rc = make_out_map32(&tmY, d->mode >= 2 ? 1 : 0, d->y, d->y_ld, d->Cout, d->B, d->H, d->W, "conv_fwd(tc) Y");
  if (rc) return rc;
  {
    const int wt = d->mode == 2 ? 1 : p.taps;
    cuuint64_t dims[3] = {(cuuint64_t)d->Cin, (cuuint64_t)p.Ncols, (cuuint64_t)wt};
    cuuint64_t str[2] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)p.Ncols * d->Cin * 2};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)n_tile, 1};
    rc = encode_map(&tmB, d->w, 3, dims, str, box, "conv_fwd(tc) B");
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n_tile == 256) return launch_tc<256, 3>(tmA, tmB, tmY, p, st);
  if (n_tile == 128) return launch_tc<128, 5>(tmA, tmB, tmY, p, st);
  return launch_tc<64, 6>(tmA, tmB, tmY, p, st);
}

// =====================================================================================================
// Block.forward in ONE launch (ddpm.py:164-173, :189-200): 3x3 conv -> GroupNorm -> FiLM -> SiLU (+ residual).
//
// The halo kernel above, re-organised so that the GroupNorm never sees HBM: a CTA - or a thread-block cluster of
// CL CTAs - owns ONE sample, and every accumulator tile of the sample stays resident in TMEM (512 columns x 128
// lanes = the conv output of 1024 pixels x 64 channels in fp32) until the sample's statistics are known:
//   MMA warp      all tiles of the CTA back to back, one TMEM accumulator each (no drain / reuse)
//   epilogue 1    as each accumulator completes: tcgen05.ld -> + bias -> per-(8-channel chunk) sum / sum of squares
//                 (overlaps the MMAs of the following tiles); training also stores the raw conv output (bf16) that
//                 the GroupNorm backward wants
//   reduce        warp shuffles -> shared memory -> (clusters) distributed shared memory; deterministic, no atomics
//   epilogue 2    tcgen05.ld again -> z = A_c * acc + B_c (gamma, beta, mean, rstd, FiLM scale/shift and the conv
//                 bias folded into two per-channel coefficients) -> SiLU via one tanh.approx -> (+ residual) ->
//                 bf16 -> swizzled staging -> TMA store
// The statistics are taken from the fp32 accumulators, i.e. the conv output is never rounded to bf16 before the norm,
// and the conv -> norm intermediate makes no HBM round trip (inference) / is written once and never re-read (training).
// =====================================================================================================
constexpr int GN_STAGE_BYTES = TC_EPI_WARPS * 2 * 4096;     // 64 KiB
struct TcGnParams {
  int kblocks;                       // Cin / 64
  int H, W, tiles_x, tpc;            // 8x16-pixel tiles per image row; pixel tiles per CTA
  int Cout, n_tiles, gs_shift;       // channels per group = 1 << gs_shift
  int res_ld, film_ld;
  int store_raw, tmem_cols, B, nbuf;
  const __nv_bfloat16* res;
  const float *bias, *gamma, *beta, *film;
  float* stats;                      // [B][8][2] (mean, rstd) or nullptr
  float eps, inv_count;              // 1 / (H*W*channels per group)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// acquire-wait (cluster scope) on a local mbarrier whose transactions come from the peers (st.async); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 8000000000LL) {
      printf("b200dm: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// one fp32 into the shared memory of CTA `rank` (same offsets as here) + complete_tx of 4 bytes on its mbarrier: data
// and signal travel together, no fence, and the receiver reads its own shared memory once the phase completes
__device__ __forceinline__ void st_async_f32_cluster(uint32_t local_addr, uint32_t local_bar, uint32_t rank, float v) {
  uint32_t raddr, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(local_bar), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ float tanh_approx_f(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 16 per-thread values (8 chunk sums, 8 chunk sums of squares) reduced over the 32 lanes of a warp by recursive
// halving (16 shuffles + 1): afterwards lane l holds the warp total of value index (idx & 7) + 8 * (idx >> 3), where
// idx is returned; lanes with an odd lane id hold a duplicate.
__device__ __forceinline__ float warp_reduce16(const float (&S)[8], const float (&Q)[8], int lane, int& idx_out) {
  float v[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] = S[k]; v[8 + k] = Q[k]; }
  int idx = 0;
#pragma unroll
  for (int n = 8, step = 0; n >= 1; n >>= 1, ++step) {
    const int m = 16 >> step;
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float keep = up ? v[i + n] : v[i];
      const float send = up ? v[i] : v[i + n];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
    idx = idx * 2 + (up ? 1 : 0);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  idx_out = idx;
  return v[0];
}

template <int N_TILE, int A_BUFS, int B_STAGES, bool B_RESIDENT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_gn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmRaw,
                  const TcGnParams p) {
  constexpr int B_BYTES = N_TILE * TC_BK * 2;
  constexpr int SUBTILES = N_TILE / 64;
  constexpr int MAX_ACC = 8;
  pdl_launch_dependents();
  if (threadIdx.x == 0) TSTAMP(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + A_BUFS * HALO_BYTES;
  const int b_slots = B_RESIDENT ? 9 * p.kblocks : B_STAGES;
  const uint32_t out_stage = b_base + b_slots * B_BYTES;        // two private 4 KiB staging blocks per epilogue warp
  const uint32_t bar_base = out_stage + GN_STAGE_BYTES;
  auto afull = [&](int a) { return bar_base + 8u * a; };
  auto aempty = [&](int a) { return bar_base + 8u * (A_BUFS + a); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * A_BUFS + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * A_BUFS + B_STAGES + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + MAX_ACC + a); };
  const uint32_t bres_bar = bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + 2 * MAX_ACC);
  // cluster exchange of the group sums: two barriers (sample parity), transaction-counted
  auto xbar = [&](uint32_t par) { return bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + 2 * MAX_ACC + 1 + par); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * A_BUFS + 2 * B_STAGES + 2 * MAX_ACC + 3);
  uint8_t* const sm = smem_raw + (base - smem_u32(smem_raw));         // generic pointer to `base`
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));
  uint8_t* out_stage_ptr = sm + (out_stage - base);
  // fp32 scratch after the barriers: bias | coefA | coefB [Cout each] | wpart[2][4][4][8][2] | gpart[2][16] | gall[8][16]
  const uint32_t f_base = bar_base + 512u;
  float* bias_s = reinterpret_cast<float*>(sm + (f_base - base));
  float* coefA = bias_s + p.Cout;
  float* coefB = coefA + p.Cout;
  float* wpart = coefB + p.Cout;                 // [wset][quarter][unit slot][chunk][sum, sumsq]
  float* gpart = wpart + 2 * 4 * 4 * 8 * 2;      // [group][sum, sumsq] of this CTA (single-CTA samples)
  float* gall = gpart + 32;                      // [sample parity][cluster rank][16]: every CTA's sums, sent by the owners
  float* part_s = gall + 256;                    // [group][which][quarter][tile parity]: second reduction stage
  const uint32_t gall_addr = f_base + (uint32_t)((3 * p.Cout + 2 * 4 * 4 * 8 * 2 + 32) * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank(), CL = (int)cluster_nctarank();
  const int n_clusters = (int)gridDim.x / CL, cluster_id = (int)blockIdx.x / CL;
  const int NA = p.tpc * p.n_tiles;              // accumulators of one sample in this CTA (<= 8)
  const int NU = NA * SUBTILES;                  // 64-column units
  // two accumulator sets (even / odd samples) when one sample needs at most half of TMEM: the MMAs of the next sample
  // then do not wait for the apply pass of the current one at all
  const int nbuf = p.nbuf;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    if (p.store_raw) prefetch_tmap(&tmRaw);
    for (int a = 0; a < A_BUFS; ++a) { mbar_init(afull(a), 1); mbar_init(aempty(a), 1); }
    for (int s2 = 0; s2 < B_STAGES; ++s2) { mbar_init(bfull(s2), 1); mbar_init(bempty(s2), 1); }
    for (int a = 0; a < MAX_ACC; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), 4 * SUBTILES); }
    mbar_init(bres_bar, 1);
    mbar_init(xbar(0), 1);                       // one local arrive.expect_tx per sample; the data arrive as transactions
    mbar_init(xbar(1), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  // the peers' exchange barriers must be initialised before anybody arrives on them
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) TSTAMP(1);
  pdl_wait();   // everything above is on-chip setup; global memory is touched only below
  if (threadIdx.x == 0) TSTAMP(2);

  if (warp == 0) {
    // ===================== TMA producer: activation halos of every sample of this cluster =====================
    if (lane == 0) {
      if (B_RESIDENT) {
        mbar_expect_tx(bres_bar, (uint32_t)(9 * p.kblocks * B_BYTES));
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < p.kblocks; ++kc)
            tma_load_3d(b_base + (tap * p.kblocks + kc) * B_BYTES, &tmB, bres_bar, kc * TC_BK, 0, tap);
      }
      int ab = 0;
      uint32_t aph = 0;
      for (int b = cluster_id; b < p.B; b += n_clusters)
        for (int mi = 0; mi < p.tpc; ++mi) {
          const int m = rank * p.tpc + mi;
          const int y0 = (m / p.tiles_x) * 16, x0 = (m % p.tiles_x) * 8;
          for (int n = 0; n < p.n_tiles; ++n)
            for (int kc = 0; kc < p.kblocks; ++kc) {
              mbar_wait(aempty(ab), aph ^ 1u);
              mbar_expect_tx(afull(ab), HALO_TX);
              tma_load_5d(a_base + ab * HALO_BYTES, &tmA, afull(ab), kc * TC_BK, x0 - 1, y0 - 1, b, 0);
              if (++ab == A_BUFS) { ab = 0; aph ^= 1u; }
            }
        }
    }
  } else if (warp == TC_BWARP) {
    // ===================== TMA producer: weight tiles (streamed variant) =====================
    if (lane == 0 && !B_RESIDENT) {
      int bs = 0;
      uint32_t bph = 0;
      for (int b = cluster_id; b < p.B; b += n_clusters)
        for (int mi = 0; mi < p.tpc; ++mi)
          for (int n = 0; n < p.n_tiles; ++n)
            for (int kc = 0; kc < p.kblocks; ++kc)
              for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(bempty(bs), bph ^ 1u);
                mbar_expect_tx(bfull(bs), B_BYTES);
                tma_load_3d(b_base + bs * B_BYTES, &tmB, bfull(bs), kc * TC_BK, n * N_TILE, tap);
                if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
              }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Accumulator a of the NEXT sample is started as soon as epilogue 2 of the current sample has read it
    // (acc_empty), so the apply pass of one sample overlaps the MMAs of the following one, tile by tile.
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, N_TILE, 0, 0);
    const uint64_t a_desc0 = make_smem_desc(a_base, 16, HALO_W * 128);
    const uint64_t b_desc0 = make_smem_desc(b_base, 16, 1024);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
    const bool leader = elect_one();
    int ab = 0, bs = 0;
    uint32_t aph = 0, bph = 0, sph = 0;          // sph: parity of the sample iteration
    if (B_RESIDENT) mbar_wait(bres_bar, 0);
    int si_ = 0;
    for (int b = cluster_id; b < p.B; b += n_clusters, sph ^= 1u, ++si_) {
      const int set0 = (si_ & (nbuf - 1)) * NA;                      // first accumulator of this sample's set
      const uint32_t bph2 = (uint32_t)(nbuf == 2 ? (si_ >> 1) : si_) & 1u;
      for (int a = 0; a < NA; ++a) {
        mbar_wait(acc_empty(set0 + a), bph2 ^ 1u);
        tc_fence_after();
        if (leader && si_ < 4) TSTAMP(200 + si_ * 16 + 2 * a);      // accumulator a handed back
        const uint32_t d_tmem = tmem_base + (uint32_t)((set0 + a) * N_TILE);
        for (int kc = 0; kc < p.kblocks; ++kc) {
          mbar_wait(afull(ab), aph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)ab * (HALO_BYTES >> 4);
          if (B_RESIDENT) {
            if (leader) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap - dy * 3;
                const uint32_t b_lo = b_lo0 + (uint32_t)(tap * p.kblocks + kc) * (B_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                  umma_bf16_lohi(d_tmem, a_lo + (uint32_t)(((dx + HALO_W * dy) * 128 + k * 32) >> 4), a_hi,
                                 b_lo + (uint32_t)((k * 32) >> 4), b_hi, idesc,
                                 (tap > 0 || k > 0) ? 1u : (kc > 0 ? 1u : 0u));
              }
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3, dx = tap - dy * 3;
              mbar_wait(bfull(bs), bph);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + (uint32_t)bs * (B_BYTES >> 4);
              if (leader) {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                  umma_bf16_lohi(d_tmem, a_lo + (uint32_t)(((dx + HALO_W * dy) * 128 + k * 32) >> 4), a_hi,
                                 b_lo + (uint32_t)((k * 32) >> 4), b_hi, idesc,
                                 (tap > 0 || k > 0) ? 1u : (kc > 0 ? 1u : 0u));
                umma_commit(bempty(bs));
              }
              __syncwarp();
              if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
            }
          }
          if (leader) umma_commit(aempty(ab));
          if (++ab == A_BUFS) { ab = 0; aph ^= 1u; }
        }
        if (leader) umma_commit(acc_full(set0 + a));
        if (leader && si_ < 4) TSTAMP(200 + si_ * 16 + 2 * a + 1);  // MMAs of accumulator a issued
      }
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    const int quarter = warp & 3;
    const int wset = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int tx = row & 7, ty = row >> 3;
    const int te = threadIdx.x - 64;
    uint8_t* const stage_ptr0 = out_stage_ptr + (warp - 2) * 8192;
    const uint32_t stage0 = out_stage + (uint32_t)(warp - 2) * 8192u;
    uint32_t nst = 0;                            // TMA stores issued by this warp: staging block = nst & 1
    for (int i = te; i < p.Cout; i += 32 * TC_EPI_WARPS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    named_bar_sync(1, 32 * TC_EPI_WARPS);
    uint32_t sph = 0;
    int si_ = 0;
    const bool st_ = warp == 2 && lane == 0;
    // one channel per thread for the coefficient step: gamma / beta are the same for every sample
    const bool has_c = te < p.Cout;
    const float gam = has_c ? p.gamma[te] : 0.f, bet = has_c ? p.beta[te] : 0.f;
    for (int b = cluster_id; b < p.B; b += n_clusters, sph ^= 1u, ++si_) {
      if (st_ && si_ < 4) TSTAMP(10 + si_ * 40);                     // sample starts (epilogue view)
      const int set0 = (si_ & (nbuf - 1)) * NA;
      const uint32_t bph2 = (uint32_t)(nbuf == 2 ? (si_ >> 1) : si_) & 1u;
      if (CL > 1 && te == 0) mbar_expect_tx(xbar(sph), (uint32_t)(16 * CL * 4));   // the sums of every CTA, 4 B each
      // FiLM scale / shift of this sample: requested now, needed after the statistics
      float fsc = 1.f, fsh = 0.f;
      if (has_c && p.film) {
        fsc = p.film[(long long)b * p.film_ld + te] + 1.f;
        fsh = p.film[(long long)b * p.film_ld + p.Cout + te];
      }
      // ---------- epilogue 1: statistics (+ raw conv output when training), as the accumulators complete ----------
      for (int u = wset; u < NU; u += 2) {
        const int a = u / SUBTILES, s2 = u - a * SUBTILES;
        const int mi = a / p.n_tiles, n = a - mi * p.n_tiles;
        const int m = rank * p.tpc + mi;
        const int y0 = (m / p.tiles_x) * 16, x0 = (m % p.tiles_x) * 8;
        uint8_t* my_row = stage_ptr0 + (nst & 1u) * 4096 + lane * 128;
        mbar_wait(acc_full(set0 + a), bph2);
        tc_fence_after();
        if (p.store_raw) {
          if (lane == 0) tma_store_wait_read<1>();                  // the block used two stores ago is free again
          __syncwarp();
        }
        float gS[8], gQ[8];
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int c = s2 * 64 + hh * 32;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((set0 + a) * N_TILE + c), r);
          tmem_ld_wait();
          float v[32];
          {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + n * N_TILE + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b4[j];
              v[4 * j] = __uint_as_float(r[4 * j]) + bb.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float sa = 0.f, q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              sa += v[8 * k + j];
              q = fmaf(v[8 * k + j], v[8 * k + j], q);
            }
            if (hh == 0) { gS[k] = sa; gQ[k] = q; } else { gS[4 + k] = sa; gQ[4 + k] = q; }
          }
          if (p.store_raw) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 uu;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&uu);
#pragma unroll
              for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
              const int chunk = (hh * 4 + g) ^ (lane & 7);
              *reinterpret_cast<uint4*>(my_row + chunk * 16) = uu;
            }
          }
        }
        if (p.store_raw) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tmRaw, stage0 + (nst & 1u) * 4096u, n * N_TILE + s2 * 64, x0, y0 + 4 * quarter, b, 0);
            tma_store_commit();
          }
          ++nst;
        }
        int idx;
        const float tot = warp_reduce16(gS, gQ, lane, idx);
        if ((lane & 1) == 0)
          wpart[((((wset * 4 + quarter) * 4 + (u >> 1)) * 8) + (idx & 7)) * 2 + (idx >> 3)] = tot;
        if (st_ && si_ < 4) TSTAMP(10 + si_ * 40 + 1 + (u >> 1));    // statistics of unit done (1..4)
      }
      named_bar_sync(1, 32 * TC_EPI_WARPS);
      if (st_ && si_ < 4) TSTAMP(10 + si_ * 40 + 5);                 // all warps through epilogue 1
      // CTA totals per group in a fixed order, two stages: 128 threads = (group, which, quarter, tile parity) add their
      // <= 4 entries (enumerated directly, no search), then 16 threads = (group, which) add the 8 partials
      float* gmine = gpart;
      if (te < 128) {
        const int g = te >> 4, which = (te >> 3) & 1, q = (te >> 1) & 3, par = te & 1;
        const int gs = 1 << p.gs_shift;
        float acc = 0.f;
        for (int mi = par; mi < p.tpc; mi += 2)
          for (int c = g * gs; c < (g + 1) * gs; c += 8) {
            const int n = c / N_TILE, cc = c - n * N_TILE;
            const int u = (mi * p.n_tiles + n) * SUBTILES + (cc >> 6);
            const int k = (cc & 63) >> 3;
            acc += wpart[(((((u & 1) * 4 + q) * 4 + (u >> 1)) * 8) + k) * 2 + which];
          }
        part_s[te] = acc;
      }
      named_bar_sync(1, 32 * TC_EPI_WARPS);
      if (te < 16) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += part_s[te * 8 + i];
        if (CL > 1) {
          // send this CTA's sum to slot [parity][rank][te] of EVERY CTA of the cluster (itself included)
          const uint32_t slot = gall_addr + (uint32_t)((sph * 8 + rank) * 16 + te) * 4u;
          for (int r = 0; r < CL; ++r) st_async_f32_cluster(slot, xbar(sph), (uint32_t)r, acc);
        } else {
          gmine[te] = acc;
        }
      }
      if (st_ && si_ < 4) TSTAMP(10 + si_ * 40 + 6);                 // CTA sums published
      if (CL > 1) {
        mbar_wait_cluster(xbar(sph), (uint32_t)(si_ >> 1) & 1u);
        if (st_ && si_ < 4) TSTAMP(10 + si_ * 40 + 7);               // every CTA's sums have landed here
      }
      named_bar_sync(1, 32 * TC_EPI_WARPS);          // gmine (CL == 1) / gall (CL > 1) complete
      if (st_ && si_ < 4) TSTAMP(10 + si_ * 40 + 8);
      // ---------- statistics -> per-channel coefficients ----------
      const float* gsrc = CL > 1 ? gall + sph * 128 : gmine;
      if (has_c) {
        const int c = te, g = c >> p.gs_shift;
        float sum = 0.f, sq = 0.f;
        for (int r = 0; r < CL; ++r) {
          sum += gsrc[r * 16 + g * 2];
          sq += gsrc[r * 16 + g * 2 + 1];
        }
        const float mean = sum * p.inv_count;
        const float var = fmaxf(sq * p.inv_count - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.eps);
        float ga = gam * rstd;
        float be = bet - mean * ga;
        ga *= fsc;                       // FiLM: (scale + 1), shift (1, 0 without FiLM)
        be = be * fsc + fsh;
        // z = ga * (acc + bias) + be; silu(z) = h + h * tanh(h) with h = z / 2: the halves are folded in
        coefA[c] = 0.5f * ga;
        coefB[c] = 0.5f * fmaf(ga, bias_s[c], be);
        if (p.stats && rank == 0 && (c & ((1 << p.gs_shift) - 1)) == 0) {
          p.stats[((long long)b * 8 + g) * 2] = mean;
          p.stats[((long long)b * 8 + g) * 2 + 1] = rstd;
        }
      }
      named_bar_sync(1, 32 * TC_EPI_WARPS);
      if (st_ && si_ < 4) TSTAMP(10 + si_ * 40 + 9);                 // coefficients ready
      // ---------- epilogue 2: apply straight out of TMEM; each accumulator is handed back to the MMA warp ----------
      for (int u = wset; u < NU; u += 2) {
        const int a = u / SUBTILES, s2 = u - a * SUBTILES;
        const int mi = a / p.n_tiles, n = a - mi * p.n_tiles;
        const int m = rank * p.tpc + mi;
        const int y0 = (m / p.tiles_x) * 16, x0 = (m % p.tiles_x) * 8;
        const long long opix = ((long long)b * p.H + y0 + ty) * p.W + x0 + tx;
        const int cbase = n * N_TILE + s2 * 64;
        uint8_t* my_row = stage_ptr0 + (nst & 1u) * 4096 + lane * 128;
        // the residual of the whole unit (this pixel's 64 channels = 128 B) is requested before anything else, so the
        // second half's loads have the first half's arithmetic to land behind
        uint4 src_res[8];
        if (p.res) {
          const uint4* rr = reinterpret_cast<const uint4*>(p.res + opix * p.res_ld + cbase);
#pragma unroll
          for (int g = 0; g < 8; ++g) src_res[g] = rr[g];
        }
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int c = cbase + hh * 32;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((set0 + a) * N_TILE + s2 * 64 + hh * 32), r);
          tmem_ld_wait();
          if (hh == 1) {                 // last read of this unit: the accumulator may be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(set0 + a));
          }
          float v[32];
          {
            const float4* a4 = reinterpret_cast<const float4*>(coefA + c);
            const float4* b4 = reinterpret_cast<const float4*>(coefB + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 aa = a4[j], bb = b4[j];
              const float z0 = fmaf(aa.x, __uint_as_float(r[4 * j]), bb.x);
              const float z1 = fmaf(aa.y, __uint_as_float(r[4 * j + 1]), bb.y);
              const float z2 = fmaf(aa.z, __uint_as_float(r[4 * j + 2]), bb.z);
              const float z3 = fmaf(aa.w, __uint_as_float(r[4 * j + 3]), bb.w);
              v[4 * j] = fmaf(z0, tanh_approx_f(z0), z0);
              v[4 * j + 1] = fmaf(z1, tanh_approx_f(z1), z1);
              v[4 * j + 2] = fmaf(z2, tanh_approx_f(z2), z2);
              v[4 * j + 3] = fmaf(z3, tanh_approx_f(z3), z3);
            }
          }
          if (p.res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&src_res[hh * 4 + g]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                v[g * 8 + 2 * j] += f.x;
                v[g * 8 + 2 * j + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 uu;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&uu);
#pragma unroll
            for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
            const int chunk = (hh * 4 + g) ^ (lane & 7);
            *reinterpret_cast<uint4*>(my_row + chunk * 16) = uu;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_5d(&tmY, stage0 + (nst & 1u) * 4096u, cbase, x0, y0 + 4 * quarter, b, 0);
          tma_store_commit();
        }
        ++nst;
        if (st_ && si_ < 4) TSTAMP(10 + si_ * 40 + 10 + (u >> 1));   // apply of unit done (10..13)
      }
    }
    // the staging blocks must outlive the TMA engine's reads; the global writes themselves complete before the grid
    // does (bulk-async stores are flushed at kernel completion), so the CTA does not wait for them
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  // nobody leaves while a peer may still read its group sums or arrive on its exchange barrier
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) TSTAMP(4);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// =====================================================================================================
// Block.forward in one launch for the SMALL levels (8x8 and 4x4 images): a 128-pixel M tile of conv_tc_kernel holds
// whole samples (2 or 8) and a 64-column sub-tile holds whole GroupNorm groups (16, 32 or 64 channels each), so the
// statistics are tile-local: no cluster, no second kernel.  Same producer / MMA pipeline as conv_tc_kernel (mode 0,
// 3x3); the epilogue reads its 32 x 64 accumulator block twice (statistics, then apply), reduces over the lanes of a
// sample by shuffles (and, at 8x8 where a sample spans two warps, through shared memory between the two warps), turns
// (gamma, beta, mean, rstd, FiLM, bias) into two coefficients per (sample, channel) in a per-warp shared-memory table
// and applies z = A*acc + B, SiLU, + residual straight out of TMEM.
// =====================================================================================================
struct TcGnSmallParams {
  int kblocks, H, W, bh, bn;
  int Cout, n_tiles, m_tiles, gs_shift, hw;      // hw = pixels per sample (16 or 64)
  int res_ld, film_ld, store_raw, B;
  long long M;
  const __nv_bfloat16* res;
  const float *bias, *gamma, *beta, *film;
  float* stats;
  float eps, inv_count;
};

template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_gn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmRaw,
                  const TcGnSmallParams p) {
  constexpr int B_STAGE_BYTES = N_TILE * TC_BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int TMEM_COLS = 2 * N_TILE <= 128 ? 128 : 2 * N_TILE <= 256 ? 256 : 512;
  constexpr int SUBTILES = N_TILE / 64;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_stage = base + STAGES * STAGE_BYTES;        // one 4 KiB staging block per epilogue warp
  const uint32_t bar_base = out_stage + 2 * OUT_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* const sm = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));
  uint8_t* out_stage_ptr = sm + (out_stage - base);
  // fp32 scratch: bias[Cout] | coef[8 warps][2 samples][64][2] | xch[2 parities][2 warp sets][4 quarters][8]
  float* bias_s = reinterpret_cast<float*>(sm + (bar_base + 256u - base));
  float* coef_s = bias_s + p.Cout;
  float* xch = coef_s + TC_EPI_WARPS * 2 * 64 * 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = 9 * p.kblocks;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    if (p.store_raw) prefetch_tmap(&tmRaw);
    for (int s2 = 0; s2 < STAGES; ++s2) { mbar_init(full_bar(s2), 1); mbar_init(empty_bar(s2), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), SUBTILES == 1 ? TC_EPI_WARPS / 2 : TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: activation boxes, one per (tap, channel block) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const long long first = (long long)m_tile * TC_BM;
      const int n_first = (int)(first / ((long long)p.H * p.W));
      const int y_first = (int)((first / p.W) % p.H);
      int dy = -1, dx = -1;
      for (int tap = 0; tap < 9; ++tap) {
        for (int kc = 0; kc < p.kblocks; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) {
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_5d(base + stage * STAGE_BYTES, &tmA, full_bar(stage), kc * TC_BK, dx, y_first + dy, n_first, 0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++dx > 1) { dx = -1; ++dy; }
      }
    }
  } else if (warp == TC_BWARP) {
    // ===================== TMA producer: weight tiles =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * N_TILE;
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < p.kblocks; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader)
            tma_load_3d(base + stage * STAGE_BYTES + A_STAGE_BYTES, &tmB, full_bar(stage), kc * TC_BK, n0, tap);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, N_TILE, 0, 0);
    const uint64_t a_desc0 = make_smem_desc(base, 16, 1024);
    const uint64_t b_desc0 = make_smem_desc(base + A_STAGE_BYTES, 16, 1024);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
    const bool leader = elect_one();
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N_TILE);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full_bar(stage), phase);
        if (leader) {
          const uint32_t so = (uint32_t)stage * (STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16_lohi(d_tmem, a_lo0 + so + (uint32_t)((k * 32) >> 4), a_hi,
                           b_lo0 + so + (uint32_t)((k * 32) >> 4), b_hi, idesc, (k > 0) ? 1u : (kb > 0 ? 1u : 0u));
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(tmem_full_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;
    const int wset = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t my_stage = out_stage + (uint32_t)(warp - 2) * 4096u;
    uint8_t* my_row = out_stage_ptr + (warp - 2) * 4096 + lane * 128;
    float* my_coef = coef_s + (warp - 2) * 256;           // [sample in warp][64][A, B]
    const int R = p.hw;                                   // rows (pixels) per sample: 16 or 64
    const int seg = R >= 32 ? 32 : R;                     // lanes of this warp that belong to one sample
    const int sloc = lane / seg, lseg = lane - sloc * seg;
    const int gs = 1 << p.gs_shift;
    for (int i = threadIdx.x - 64; i < p.Cout; i += 32 * TC_EPI_WARPS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    named_bar_sync(1, 32 * TC_EPI_WARPS);
    int it = 0, unit = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (SUBTILES == 1 && (it & 1) != wset) continue;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * N_TILE;
      const long long first = (long long)m_tile * TC_BM;
      const long long pix = first + row;
      const int bsm = (int)(pix / R);                     // this row's sample
      const bool valid = pix < p.M;
      const unsigned slice = (unsigned)first + 32u * (unsigned)quarter;
      const unsigned hw = (unsigned)(p.H * p.W);
      const int sn = (int)(slice / hw);
      const unsigned srem = slice - (unsigned)sn * hw;
      const int sy = (int)(srem / (unsigned)p.W), sx = (int)(srem - (unsigned)sy * (unsigned)p.W);
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      const int s_first = SUBTILES == 1 ? 0 : wset;
#pragma unroll 1
      for (int s = s_first; s < SUBTILES; s += 2, ++unit) {
        const int cb = n0 + s * 64;                       // first output channel of this 64-column unit
        // ---------- pass 1: statistics (+ raw conv output when training) ----------
        if (p.store_raw) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        float Sg[4] = {0.f, 0.f, 0.f, 0.f}, Qg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE + s * 64 + hh * 32), r);
          tmem_ld_wait();
          float v[32];
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + cb + hh * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            v[4 * j] = __uint_as_float(r[4 * j]) + bb.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float sa = 0.f, q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              sa += v[8 * k + j];
              q = fmaf(v[8 * k + j], v[8 * k + j], q);
            }
            const int gi = (hh * 32 + 8 * k) >> p.gs_shift;          // group of this 8-channel chunk inside the unit
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              Sg[g] += gi == g ? sa : 0.f;
              Qg[g] += gi == g ? q : 0.f;
            }
          }
          if (p.store_raw) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 uu;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&uu);
#pragma unroll
              for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
              const int chunk = (hh * 4 + g) ^ (lane & 7);
              *reinterpret_cast<uint4*>(my_row + chunk * 16) = uu;
            }
          }
        }
        if (p.store_raw) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tmRaw, my_stage, cb, sx, sy, sn, 0);
            tma_store_commit();
          }
        }
        // sums over the lanes of a sample (16 or 32 lanes), then over the two warps of a 64-pixel sample
        for (int off = seg >> 1; off > 0; off >>= 1) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            Sg[g] += __shfl_xor_sync(0xffffffffu, Sg[g], off);
            Qg[g] += __shfl_xor_sync(0xffffffffu, Qg[g], off);
          }
        }
        if (R > 32) {
          float* mine = xch + (((unit & 1) * 2 + wset) * 4 + quarter) * 8;
          if (lane == 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g) { mine[g] = Sg[g]; mine[4 + g] = Qg[g]; }
          }
          named_bar_sync(2 + wset, 128);                  // the four quarter-warps of this warp set, once per unit
          const float* other = xch + (((unit & 1) * 2 + wset) * 4 + (quarter ^ 1)) * 8;
#pragma unroll
          for (int g = 0; g < 4; ++g) { Sg[g] += other[g]; Qg[g] += other[4 + g]; }
        }
        float mean[4], rstd[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          mean[g] = Sg[g] * p.inv_count;
          rstd[g] = rsqrtf(fmaxf(Qg[g] * p.inv_count - mean[g] * mean[g], 0.f) + p.eps);
        }
        const int ngrp = 64 >> p.gs_shift;                // groups inside this unit (1, 2 or 4)
        if (p.stats && valid && lseg == 0 && (R <= 32 || (quarter & 1) == 0)) {
          for (int g = 0; g < ngrp; ++g) {
            float* st = p.stats + ((long long)bsm * 8 + (cb >> p.gs_shift) + g) * 2;
            st[0] = mean[g];
            st[1] = rstd[g];
          }
        }
        // coefficients of this lane's sample for the 64 channels of the unit, split over the lanes of the sample
        __syncwarp();
        for (int cc = lseg; cc < 64; cc += seg) {
          const int c = cb + cc, gi = cc >> p.gs_shift;
          float mu = mean[0], rs = rstd[0];
#pragma unroll
          for (int g = 1; g < 4; ++g) { mu = gi == g ? mean[g] : mu; rs = gi == g ? rstd[g] : rs; }
          float ga = p.gamma[c] * rs;
          float be = p.beta[c] - mu * ga;
          if (p.film && valid) {
            const float sc = p.film[(long long)bsm * p.film_ld + c] + 1.f;
            const float sh = p.film[(long long)bsm * p.film_ld + p.Cout + c];
            ga *= sc;
            be = be * sc + sh;
          }
          my_coef[(sloc * 64 + cc) * 2] = 0.5f * ga;
          my_coef[(sloc * 64 + cc) * 2 + 1] = 0.5f * fmaf(ga, bias_s[c], be);
        }
        __syncwarp();
        // ---------- pass 2: apply straight out of TMEM ----------
        const long long opix = pix;
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint4 src_res[4];
          if (valid && p.res) {
            const uint4* rr = reinterpret_cast<const uint4*>(p.res + opix * p.res_ld + cb + hh * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) src_res[g] = rr[g];
          }
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE + s * 64 + hh * 32), r);
          tmem_ld_wait();
          if (hh == 1 && s + 2 >= SUBTILES) {   // last read of this accumulator by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
          }
          float v[32];
          const float4* c4 = reinterpret_cast<const float4*>(my_coef + (sloc * 64 + hh * 32) * 2);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 ab = c4[j];              // (A, B) of two channels
            const float z0 = fmaf(ab.x, __uint_as_float(r[2 * j]), ab.y);
            const float z1 = fmaf(ab.z, __uint_as_float(r[2 * j + 1]), ab.w);
            v[2 * j] = fmaf(z0, tanh_approx_f(z0), z0);
            v[2 * j + 1] = fmaf(z1, tanh_approx_f(z1), z1);
          }
          if (valid && p.res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&src_res[g]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                v[g * 8 + 2 * j] += f.x;
                v[g * 8 + 2 * j + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 uu;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&uu);
#pragma unroll
            for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
            const int chunk = (hh * 4 + g) ^ (lane & 7);
            *reinterpret_cast<uint4*>(my_row + chunk * 16) = uu;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_5d(&tmY, my_stage, cb, sx, sy, sn, 0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int N_TILE, int STAGES>
static int launch_tc_gn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                        const CUtensorMap& tmRaw, const TcGnSmallParams& p, cudaStream_t st) {
  const int smem = STAGES * (A_STAGE_BYTES + N_TILE * TC_BK * 2) + 2 * OUT_STAGE_BYTES + 1024 + 256 +
                   (p.Cout + TC_EPI_WARPS * 256 + 2 * 2 * 4 * 8) * 4;
  B200DM_REQUIRE(smem <= 227 * 1024, B200DM_ERR_UNSUPPORTED, "conv_gn(small): %d B of shared memory", smem);
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_gn_kernel<N_TILE, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "conv_gn(small): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  launch_k(conv_tc_gn_kernel<N_TILE, STAGES>, grid, TC_THREADS, smem, st, tmA, tmB, tmY, tmRaw, p);
  count_launch();
  return check_launch("conv_gn(small)");
}

template <int N_TILE, int A_BUFS, int B_STAGES, bool B_RESIDENT>
static int launch_gn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmRaw,
                     const TcGnParams& p, int B, int CL, cudaStream_t st) {
  const int b_slots = B_RESIDENT ? 9 * p.kblocks : B_STAGES;
  const int smem = A_BUFS * HALO_BYTES + b_slots * (N_TILE * TC_BK * 2) + GN_STAGE_BYTES + 1024 + 512 +
                   (3 * p.Cout + 2 * 4 * 4 * 8 * 2 + 32 + 256 + 128) * 4;
  B200DM_REQUIRE(smem <= 227 * 1024, B200DM_ERR_UNSUPPORTED, "conv_gn: %d B of shared memory", smem);
  static int configured = 0;
  auto kernel = conv3x3_gn_kernel<N_TILE, A_BUFS, B_STAGES, B_RESIDENT>;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "conv_gn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // persistent grid: as many clusters as the machine holds at once (one CTA per SM; a cluster lives inside one GPC, so
  // this is not simply SMs / CL), each walking samples cluster_id, cluster_id + n_clusters, ...
  static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (max_clusters[CL] == 0) {
    cfg.gridDim = dim3(CL * 64);
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms() / CL;
    }
    max_clusters[CL] = n;
  }
  const int n_clusters = B < max_clusters[CL] ? B : max_clusters[CL];
  cfg.gridDim = dim3(n_clusters * CL);
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmY, tmRaw, p);
  B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "conv_gn: launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("conv_gn");
}

// Geometry of the fused launch: tile shape, cluster size and TMEM columns; false if the layer does not fit.
struct GnGeom { int n_tile, n_tiles, tiles, cl, tpc, tmem_cols, nbuf; };
static bool conv_gn_geometry(const b200dm_conv_desc* d, int groups, GnGeom* g) {
  if (d->dtype != B200DM_BF16 || d->mode != 0 || d->ksize != 3 || d->accumulate) return false;
  if (d->Cin % 64 || d->Cout % 64 || d->Cout > 256 || groups != 8) return false;
  if (d->W < 16 || d->W % 8 || d->H % 16 || d->W > 128 || d->H > 128) return false;
  const int gs = d->Cout / groups;
  if (gs < 8 || (gs & (gs - 1))) return false;
  g->n_tile = d->Cout % 128 == 0 ? 128 : 64;
  g->n_tiles = d->Cout / g->n_tile;
  g->tiles = (d->W / 8) * (d->H / 16);
  // cluster size: the smallest power of two for which the sample's accumulators fit the CTAs' TMEM (512 columns
  // each); then doubled while that still divides the tiles and the machine is less than ~85 % full
  int cl = 1;
  while (cl <= 8 && (g->tiles % cl || (g->tiles / cl) * d->Cout > 512 || (g->tiles / cl) * g->n_tiles > 8)) cl *= 2;
  if (cl > 8) return false;
  const int sms = num_sms();
  while (cl * 2 <= 8 && g->tiles % (cl * 2) == 0 && (long long)d->B * cl * 2 <= sms && g->tiles / (cl * 2) >= 1 &&
         (g->tiles / (cl * 2)) * g->n_tiles * (g->n_tile / 64) >= 2)
    cl *= 2;
  g->cl = cl;
  g->tpc = g->tiles / cl;
  g->nbuf = (2 * g->tpc * d->Cout <= 512 && 2 * g->tpc * g->n_tiles <= 8) ? 2 : 1;
  const int cols = g->nbuf * g->tpc * d->Cout;
  g->tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  return true;
}

// MEASURED (scripts/conv_microbench.py --what gn, profiles/r2_conv_gn_micro.log): against conv (statistics in the
// epilogue) + one-pass norm, the fused launch is 10-30 % faster when one wave of CTAs covers the batch (training and
// 8-way sharded sampling) and within +-7 % when the persistent clusters walk several samples (batch 256 at 64x64), so the
// launch plan uses it wherever the layer fits: results then do not depend on how the batch is sharded either.
// small levels (conv_tc_gn_kernel): 8x8 or 4x4 images, groups of 16 / 32 / 64 channels
static bool conv_gn_small_geometry(const b200dm_conv_desc* d, int groups) {
  if (d->dtype != B200DM_BF16 || d->mode != 0 || d->ksize != 3 || d->accumulate || groups != 8) return false;
  if (d->Cin % 64 || d->Cout % 64 || d->Cout > 512) return false;
  if (d->H != d->W || (d->H != 4 && d->H != 8)) return false;
  const int gs = d->Cout / groups;
  return gs == 16 || gs == 32 || gs == 64;
}

int conv_gn_supported_tc(const b200dm_conv_desc* d, const b200dm_gn_desc* gn) {
  GnGeom g;
  if (!tc_supported() || !gn) return 0;
  if (halo_enabled() && conv_gn_geometry(d, gn->groups, &g)) return 1;
  return conv_gn_small_geometry(d, gn->groups) ? 1 : 0;
}

static int conv_gn_fwd_small(const b200dm_conv_desc* d, const b200dm_gn_desc* gn, cudaStream_t st) {
  int bh, bn;
  int rc = tile_geometry(d->H, d->W, &bh, &bn, "conv_gn_fwd");
  if (rc) return rc;
  TcGnSmallParams p{};
  p.kblocks = d->Cin / TC_BK;
  p.H = d->H; p.W = d->W; p.bh = bh; p.bn = bn;
  p.Cout = d->Cout; p.hw = d->H * d->W; p.B = d->B;
  int gs = d->Cout / gn->groups, sh = 0;
  while ((1 << sh) < gs) ++sh;
  p.gs_shift = sh;
  p.M = (long long)d->B * d->H * d->W;
  p.m_tiles = (int)((p.M + TC_BM - 1) / TC_BM);
  p.res = (const __nv_bfloat16*)d->res; p.res_ld = d->res_ld;
  p.film = gn->film; p.film_ld = gn->film_ld;
  p.store_raw = gn->raw ? 1 : 0;
  p.bias = d->bias; p.gamma = gn->gamma; p.beta = gn->beta; p.stats = gn->stats;
  p.eps = gn->eps;
  p.inv_count = 1.f / ((float)p.hw * (float)gs);
  // N tile as in conv_fwd_tc: tile-shape efficiency x fill of the last wave
  const int sms = num_sms();
  int n_tile = 64;
  double best = -1.0;
  const int cand[3] = {256, 128, 64};
  const double eff[3] = {1.0, 0.85, 0.6};
  for (int i = 0; i < 3; ++i) {
    if (d->Cout % cand[i]) continue;
    const long long tiles = (long long)p.m_tiles * (d->Cout / cand[i]);
    const long long waves = (tiles + sms - 1) / sms;
    const double score = eff[i] * (double)tiles / (double)(waves * sms);
    if (score > best) { best = score; n_tile = cand[i]; }
  }
  p.n_tiles = d->Cout / n_tile;
  CUtensorMap tmA, tmB, tmY, tmRaw;
  rc = make_act_map(&tmA, 0, d->x, d->x_ld, d->Cin, d->B, d->H, d->W, bh, bn, "conv_gn_fwd A");
  if (rc) return rc;
  rc = make_out_map32(&tmY, 0, d->y, d->y_ld, d->Cout, d->B, d->H, d->W, "conv_gn_fwd Y");
  if (rc) return rc;
  rc = make_out_map32(&tmRaw, 0, gn->raw ? gn->raw : d->y, gn->raw ? gn->raw_ld : d->y_ld, d->Cout, d->B, d->H, d->W,
                      "conv_gn_fwd raw");
  if (rc) return rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Cout, 9};
    cuuint64_t str[2] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cout * d->Cin * 2};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)n_tile, 1};
    rc = encode_map(&tmB, d->w, 3, dims, str, box, "conv_gn_fwd B");
    if (rc) return rc;
  }
  if (n_tile == 256) return launch_tc_gn<256, 3>(tmA, tmB, tmY, tmRaw, p, st);
  if (n_tile == 128) return launch_tc_gn<128, 5>(tmA, tmB, tmY, tmRaw, p, st);
  return launch_tc_gn<64, 6>(tmA, tmB, tmY, tmRaw, p, st);
}

int conv_gn_fwd_tc(const b200dm_conv_desc* d, const b200dm_gn_desc* gn, void* stream) {
  B200DM_REQUIRE(tc_supported(), B200DM_ERR_UNSUPPORTED, "conv_gn_fwd: needs an sm_100 device and a TMA-capable driver");
  GnGeom g;
  B200DM_REQUIRE(gn != nullptr, B200DM_ERR_SHAPE, "conv_gn_fwd: null norm descriptor");
  const bool big = halo_enabled() && conv_gn_geometry(d, gn->groups, &g);
  B200DM_REQUIRE(big || conv_gn_small_geometry(d, gn->groups), B200DM_ERR_UNSUPPORTED,
                 "conv_gn_fwd: unsupported layer (bf16 3x3 'same', Cin %% 64 == 0, 8 groups; 16 <= W <= 128 with Cout in "
                 "{64,128,256}, or 8x8 / 4x4 images with Cout in {128,256,512}): Cin=%d Cout=%d H=%d W=%d", d->Cin,
                 d->Cout, d->H, d->W);
  B200DM_REQUIRE(d->x_ld % 8 == 0 && d->y_ld % 8 == 0 && (!d->res || d->res_ld % 8 == 0) &&
                     (!gn->raw || gn->raw_ld % 8 == 0), B200DM_ERR_SHAPE, "conv_gn_fwd: ld must be a multiple of 8");
  B200DM_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->y & 15) == 0 && ((uintptr_t)d->w & 15) == 0 &&
                     ((uintptr_t)d->res & 15) == 0 && ((uintptr_t)gn->raw & 15) == 0,
                 B200DM_ERR_SHAPE, "conv_gn_fwd: pointers must be 16-byte aligned");
  B200DM_REQUIRE(gn->gamma && gn->beta, B200DM_ERR_SHAPE, "conv_gn_fwd: gamma / beta required");
  if (!big) return conv_gn_fwd_small(d, gn, (cudaStream_t)stream);
  TcGnParams p{};
  p.kblocks = d->Cin / TC_BK;
  p.H = d->H; p.W = d->W; p.tiles_x = d->W / 8; p.tpc = g.tpc;
  p.Cout = d->Cout; p.n_tiles = g.n_tiles;
  int gs = d->Cout / gn->groups, sh = 0;
  while ((1 << sh) < gs) ++sh;
  p.gs_shift = sh;
  p.res = (const __nv_bfloat16*)d->res; p.res_ld = d->res_ld;
  p.film = gn->film; p.film_ld = gn->film_ld;
  p.store_raw = gn->raw ? 1 : 0;
  p.B = d->B;
  p.tmem_cols = g.tmem_cols;
  p.nbuf = g.nbuf;
  p.bias = d->bias; p.gamma = gn->gamma; p.beta = gn->beta;
  p.stats = gn->stats;
  p.eps = gn->eps;
  p.inv_count = 1.f / ((float)d->H * (float)d->W * (float)gs);

  CUtensorMap tmA, tmB, tmY, tmRaw;
  const cuuint64_t e = 2;
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B, 1};
    cuuint64_t str[4] = {(cuuint64_t)d->x_ld * e, (cuuint64_t)d->W * d->x_ld * e,
                         (cuuint64_t)d->H * d->W * d->x_ld * e, (cuuint64_t)d->B * d->H * d->W * d->x_ld * e};
    cuuint32_t box[5] = {TC_BK, HALO_W, HALO_H, 1, 1};
    int rc = encode_map(&tmA, d->x, 5, dims, str, box, "conv_gn A");
    if (rc) return rc;
  }
  auto out_map = [&](CUtensorMap* m, const void* ptr, int ld, const char* what) {
    cuuint64_t dims[5] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B, 1};
    cuuint64_t str[4] = {(cuuint64_t)ld * e, (cuuint64_t)d->W * ld * e, (cuuint64_t)d->H * d->W * ld * e,
                         (cuuint64_t)d->B * d->H * d->W * ld * e};
    cuuint32_t box[5] = {64, 8, 4, 1, 1};
    return encode_map(m, ptr, 5, dims, str, box, what);
  };
  int rc = out_map(&tmY, d->y, d->y_ld, "conv_gn Y");
  if (rc) return rc;
  rc = out_map(&tmRaw, gn->raw ? gn->raw : d->y, gn->raw ? gn->raw_ld : d->y_ld, "conv_gn raw");
  if (rc) return rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Cout, 9};
    cuuint64_t str[2] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cout * d->Cin * 2};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)g.n_tile, 1};
    rc = encode_map(&tmB, d->w, 3, dims, str, box, "conv_gn B");
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool resident = (d->Cout == 64 && d->Cin == 64);
  if (resident) return launch_gn<64, HALO_W <= 10 ? 3 : 2, 1, true>(tmA, tmB, tmY, tmRaw, p, d->B, g.cl, st);
  if (g.n_tile == 128) return launch_gn<128, HALO_W <= 10 ? 3 : 2, HALO_W <= 10 ? 5 : 4, false>(tmA, tmB, tmY, tmRaw, p, d->B, g.cl, st);
  return launch_gn<64, HALO_W <= 10 ? 4 : 2, 6, false>(tmA, tmB, tmY, tmRaw, p, d->B, g.cl, st);
}

// =====================================================================================================
// Weight gradient:  dW[tap][co][ci] += sum_pixels dY[pix, co] * X[pix (+) tap, ci]
// GEMM with M = co (128), N = ci (N_TILE), K = pixels.  Both operands are "MN-major" in shared memory:
// a TMA box is [128 pixels][64 channels] (128-B swizzled rows), i.e. K runs over rows.  M = 128 spans
// two such boxes (LBO = 16 KiB apart).  Split-K over pixel tiles; fp32 partials are reduced with
// red.global.add into the (pre-zeroed or accumulating) master-gradient arena.
// =====================================================================================================
struct TcWgradParams {
  int mode, ksize;
  int H, W;
  int Cout, Cin;
  int x_ld;
  int k_tiles;        // number of 128-pixel tiles
  int splits;
  float* dw;
  long long s_tap, s_co, s_ci;
  int cin_valid;      // columns ci >= cin_valid are padding (stem im2col) and are not written
  int vec_ok;         // 16-byte vector reductions allowed (alignment of every row start)
};

template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(WG_THREADS)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                const TcWgradParams p) {
  constexpr int BOX_BYTES = TC_BM * TC_BK * 2;  // 16 KiB: [128 pixels][64 channels]
  constexpr int NB = N_TILE / 64;
  constexpr int STAGE_BYTES = (2 + NB) * BOX_BYTES;
  constexpr int TMEM_COLS = N_TILE <= 64 ? 64 : N_TILE <= 128 ? 128 : 256;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * TC_BM, ci0 = blockIdx.y * N_TILE;
  const int tap = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int per = (p.k_tiles + p.splits - 1) / p.splits;
  const int kt0 = split * per, kt1 = min(kt0 + per, p.k_tiles);
  const bool has_work = kt1 > kt0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // everything above is on-chip setup; global memory is touched only below

  if (warp == 0) {
    if (lane == 0 && has_work) {
      int stage = 0;
      uint32_t phase = 0;
      const int pad = p.ksize >> 1;
      const int dy = (p.mode == 0) ? tap / p.ksize - pad : 0;
      const int dx = (p.mode == 0) ? tap % p.ksize - pad : 0;
      // (image, row) of the first pixel of the K tile, advanced incrementally (no divisions in the loop)
      const int rows_per_tile = TC_BM / p.W;                       // image rows covered by one 128-pixel tile
      const int step_y = rows_per_tile <= p.H ? rows_per_tile : 0;
      const int step_n = rows_per_tile <= p.H ? 0 : rows_per_tile / p.H;
      const long long first0 = (long long)kt0 * TC_BM;
      int n_first = (int)(first0 / ((long long)p.H * p.W));
      int y_first = (int)((first0 / p.W) % p.H);
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), STAGE_BYTES);
        const uint32_t a_dst = base + stage * STAGE_BYTES;
        const uint32_t b_dst = a_dst + 2 * BOX_BYTES;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          tma_load_5d(a_dst + i * BOX_BYTES, &tmDY, full_bar(stage), co0 + 64 * i, 0, y_first, n_first, 0);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          if (p.mode == 1)
            tma_load_5d(b_dst + j * BOX_BYTES, &tmX, full_bar(stage), (tap & 1) * p.x_ld + ci0 + 64 * j, 0,
                        tap >> 1, y_first, n_first);
          else
            tma_load_5d(b_dst + j * BOX_BYTES, &tmX, full_bar(stage), ci0 + 64 * j, dx, y_first + dy,
                        n_first, 0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        y_first += step_y;
        n_first += step_n;
        if (y_first >= p.H) { y_first = 0; ++n_first; }
      }
    }
  } else if (warp == 1) {
    if (has_work) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, N_TILE, 1, 1);
      // MN-major, SWIZZLE_128B: 16 pixels (K) per MMA = 2 groups of 8 rows, 1024 B apart (SBO);
      // the next 64-channel slab of M/N lives one box further (LBO)
      const uint64_t a_desc0 = make_smem_desc(base, BOX_BYTES, 1024);
      const uint64_t b_desc0 = make_smem_desc(base + 2 * BOX_BYTES, BOX_BYTES, 1024);
      const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint32_t so = (uint32_t)stage * (STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < TC_BM / 16; ++k)
            umma_bf16_lohi(tmem_base, a_lo0 + so + (uint32_t)((k * 2048) >> 4), a_hi,
                           b_lo0 + so + (uint32_t)((k * 2048) >> 4), b_hi, idesc,
                           (k > 0) ? 1u : (kt > kt0 ? 1u : 0u));
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(tmem_full_bar);
    }
  } else if (has_work) {
    const int quarter = warp & 3;
    const int co = co0 + quarter * 32 + lane;
    const bool valid = co < p.Cout;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* drow = p.dw + (long long)tap * p.s_tap + (long long)co * p.s_co + (long long)ci0 * p.s_ci;
#pragma unroll 1
    for (int c = 0; c < N_TILE; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
      if (valid) {
        if (p.vec_ok) {      // contiguous along ci: 16-byte vector reductions (4x fewer L2 atomic ops)
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c + j),
                         "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])),
                         "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (ci0 + c + j < p.cin_valid) atomicAdd(drow + (long long)(c + j) * p.s_ci, __uint_as_float(r[j]));
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

template <int N_TILE, int STAGES>
static int launch_wgrad_tc(const CUtensorMap& tmDY, const CUtensorMap& tmX, const TcWgradParams& p,
                           dim3 grid, cudaStream_t st) {
  constexpr int smem = STAGES * (2 + N_TILE / 64) * (TC_BM * TC_BK * 2) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<N_TILE, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  launch_k(wgrad_tc_kernel<N_TILE, STAGES>, grid, WG_THREADS, smem, st, tmDY, tmX, p);
  count_launch();
  return check_launch("wgrad_tc");
}

// =====================================================================================================
// 3x3 weight gradient with an on-chip halo.  The generic kernel above runs one GEMM per filter tap, i.e. it
// streams dY and X through L2 nine times (measured: L2-bandwidth-bound, ~7 TB/s, 230 TFLOP/s on 64->64).
// Here a CTA walks 8x16-pixel tiles; per tile ONE TMA box brings the 18x16 halo of X (64 input channels) and
// one box the 8x16 tile of dY (64 output channels), and all nine taps are computed from them:
//   D[(tap slot, ci)][co] += sum_pixels X[pixel + tap, ci] * dY[pixel, co]
// A = X^T is MN-major straight from the NHWC halo; its 128 rows are TWO taps x 64 channels — the second
// 64-row slab is the same buffer shifted by the tap distance (the descriptor's leading-dimension byte
// offset), so the tensor core's M = 128 is filled even for 64-channel layers.  K = 16 pixels per MMA = two
// image rows of the tile (8-row groups 2048 B apart in the halo, 1024 B apart in the dY tile).  Six
// accumulators (taps 01, 2-, 34, 5-, 67, 8-) of 64 fp32 columns live in TMEM for the whole pixel range of
// the CTA (split-K over tiles across CTAs); the epilogue adds them into dW with coalesced fp32 reductions.
// =====================================================================================================
struct TcWgradHaloParams {
  int H, W, tiles_x, tiles_y, m_tiles;
  int ci_blocks, co_blocks, splits;
  float* dw;
  long long s_tap, s_co, s_ci;
};

// tap grouping: slot-0 tap / slot-1 tap (9 = none) of each accumulator: {0,1} {2,-} {3,4} {5,-} {6,7} {8,-}.
// Every pair is one pixel apart (LBO = 128 B).  MEASURED on B200: pairing taps one halo ROW apart (LBO = 2048 B)
// or across rows (LBO = 1792 B) gives a wrong second slab — a leading-dimension offset of a full swizzle atom
// or more is not a plain byte offset for MN-major SWIZZLE_128B operands — so only dx-neighbours are paired.
template <int PAIRING> struct WgPairs;
template <> struct WgPairs<1> {
  static constexpr int G = 6;
  __host__ __device__ static constexpr int t0(int g) { return g == 0 ? 0 : g == 1 ? 2 : g == 2 ? 3 : g == 3 ? 5 : g == 4 ? 6 : 8; }
  __host__ __device__ static constexpr int t1(int g) { return g == 0 ? 1 : g == 2 ? 4 : g == 4 ? 7 : 9; }
};

template <int STAGES, int PAIRING>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                     const TcWgradHaloParams p) {
  constexpr int DY_BYTES = TC_BM * TC_BK * 2;                  // 16 KiB: [128 pixels][64 co]
  constexpr int STAGE_BYTES = HALO_BYTES + DY_BYTES;           // 52 KiB
  constexpr int TMEM_COLS = 512;                               // 6 x 64 accumulator columns
  pdl_launch_dependents();
  if (threadIdx.x == 0) TSTAMP(300);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item: (ci block, co block, split of the pixel tiles)
  const int pair = blockIdx.x / p.splits, split = blockIdx.x - pair * p.splits;
  const int cib = pair / p.co_blocks, cob = pair - cib * p.co_blocks;
  const int per = (p.m_tiles + p.splits - 1) / p.splits;
  const int t0 = split * per, t1 = min(t0 + per, p.m_tiles);
  const bool has_work = t1 > t0;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) TSTAMP(301);
  pdl_wait();   // everything above is on-chip setup; global memory is touched only below
  if (threadIdx.x == 0) TSTAMP(302);

  if (warp == 0) {
    if (lane == 0 && has_work) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int b = t / tiles_per_img, rem = t - b * tiles_per_img;
        const int y0 = (rem / p.tiles_x) * 16, x0 = (rem % p.tiles_x) * 8;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), HALO_TX + DY_BYTES);
        const uint32_t x_dst = base + stage * STAGE_BYTES;
        tma_load_5d(x_dst, &tmX, full_bar(stage), cib * TC_BK, x0 - 1, y0 - 1, b, 0);
        tma_load_5d(x_dst + HALO_BYTES, &tmDY, full_bar(stage), cob * TC_BK, x0, y0, b, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (has_work) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, 64, 1, 1);
      // A: MN-major halo view.  SBO = 2048 B (next image row of the tile = next 8-row group);
      //    LBO = byte distance between the two taps of a pair (second 64-row slab of M).
      // B: MN-major dY tile, 8-row groups 1024 B apart; N = 64 needs no second slab.
      const uint64_t b_desc0 = make_smem_desc(base + HALO_BYTES, DY_BYTES, 1024);
      const uint32_t b_hi = (uint32_t)(b_desc0 >> 32), b_lo0 = (uint32_t)b_desc0;
      const uint64_t a_desc_near = make_smem_desc(base, 128, HALO_W * 128);                    // taps dx, dx+1
      const uint32_t a_hi = (uint32_t)(a_desc_near >> 32);
      const uint32_t a_lo0 = (uint32_t)a_desc_near;
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (leader && t - t0 < 8) TSTAMP(310 + 2 * (t - t0));
        if (leader) {
          const uint32_t so = (uint32_t)stage * (STAGE_BYTES >> 4);
#pragma unroll
          for (int g = 0; g < WgPairs<PAIRING>::G; ++g) {
            const int tap = WgPairs<PAIRING>::t0(g);
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
            for (int k = 0; k < TC_BM / 16; ++k)
              umma_bf16_lohi(tmem_base + (uint32_t)(g * 64),
                             a_lo0 + so + (uint32_t)((((2 * k + dy) * HALO_W + dx) * 128) >> 4), a_hi,
                             b_lo0 + so + (uint32_t)((k * 2048) >> 4), b_hi, idesc, (k > 0) ? 1u : (t > t0 ? 1u : 0u));
          }
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit(tmem_full_bar);
    }
  } else if (has_work) {
    // epilogue: TMEM lane = (tap slot, ci), column = (pair, co)
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane, slot = m >> 6, ci = cib * TC_BK + (m & 63);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) TSTAMP(304);
#pragma unroll 1
    for (int g = 0; g < WgPairs<PAIRING>::G; ++g) {
      const int tap = slot ? WgPairs<PAIRING>::t1(g) : WgPairs<PAIRING>::t0(g);
      float* dcol = p.dw + (long long)tap * p.s_tap + (long long)(cob * 64) * p.s_co + (long long)ci * p.s_ci;
#pragma unroll 1
      for (int c = 0; c < 64; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 64 + c), r);
        tmem_ld_wait();
        if (tap < 9) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dcol + (long long)(c + j) * p.s_co),
                         "f"(__uint_as_float(r[j]))
                         : "memory");
        }
      }
    }
  }

  if (warp == 2 && lane == 0) TSTAMP(305);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TSTAMP(306);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static bool wgrad_halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200DM_NO_WGRAD_HALO");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static int conv_wgrad_halo(const b200dm_wgrad_desc* d, cudaStream_t st) {
  constexpr int STAGES = HALO_W <= 10 ? 5 : 4;
  TcWgradHaloParams p{};
  p.H = d->H; p.W = d->W; p.tiles_x = d->W / 8; p.tiles_y = d->H / 16;
  p.m_tiles = d->B * p.tiles_x * p.tiles_y;
  p.ci_blocks = d->Cin / 64; p.co_blocks = d->Cout / 64;
  const int pairs = p.ci_blocks * p.co_blocks;
  int splits = num_sms() / pairs;
  if (splits < 1) splits = 1;
  if (splits > p.m_tiles) splits = p.m_tiles;
  p.splits = splits;
  p.dw = d->dw;
  p.s_tap = d->s_tap; p.s_co = d->s_co; p.s_ci = d->s_ci;
  if (p.s_tap == 0 && p.s_co == 0 && p.s_ci == 0) { p.s_tap = (long long)d->Cout * d->Cin; p.s_co = d->Cin; p.s_ci = 1; }
  CUtensorMap tmX, tmDY;
  const cuuint64_t e = 2;
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B, 1};
    cuuint64_t str[4] = {(cuuint64_t)d->x_ld * e, (cuuint64_t)d->W * d->x_ld * e,
                         (cuuint64_t)d->H * d->W * d->x_ld * e, (cuuint64_t)d->B * d->H * d->W * d->x_ld * e};
    cuuint32_t box[5] = {TC_BK, HALO_W, HALO_H, 1, 1};
    int rc = encode_map(&tmX, d->x, 5, dims, str, box, "conv_wgrad_halo X");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B, 1};
    cuuint64_t str[4] = {(cuuint64_t)d->dy_ld * e, (cuuint64_t)d->W * d->dy_ld * e,
                         (cuuint64_t)d->H * d->W * d->dy_ld * e, (cuuint64_t)d->B * d->H * d->W * d->dy_ld * e};
    cuuint32_t box[5] = {TC_BK, 8, 16, 1, 1};
    int rc = encode_map(&tmDY, d->dy, 5, dims, str, box, "conv_wgrad_halo dY");
    if (rc) return rc;
  }
  constexpr int smem = STAGES * (HALO_BYTES + TC_BM * TC_BK * 2) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(wgrad3x3_halo_kernel<STAGES, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    B200DM_REQUIRE(err == cudaSuccess, B200DM_ERR_CUDA, "wgrad3x3_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    configured = true;
  }
  launch_k(wgrad3x3_halo_kernel<STAGES, 1>, pairs * splits, WG_THREADS, smem, st, tmX, tmDY, p);
  count_launch();
  return check_launch("wgrad3x3_halo");
}

int conv_wgrad_tc(const b200dm_wgrad_desc* d, void* stream) {
  B200DM_REQUIRE(tc_supported(), B200DM_ERR_UNSUPPORTED, "conv_wgrad(tc): needs an sm_100 device and a TMA-capable driver");
  B200DM_REQUIRE(d->Cin % 64 == 0 && d->Cout % 64 == 0, B200DM_ERR_SHAPE,
                 "conv_wgrad(tc): Cin=%d, Cout=%d must be multiples of 64", d->Cin, d->Cout);
  B200DM_REQUIRE(d->x_ld % 8 == 0 && d->dy_ld % 8 == 0, B200DM_ERR_SHAPE, "conv_wgrad(tc): ld must be a multiple of 8");
  B200DM_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->dy & 15) == 0, B200DM_ERR_SHAPE,
                 "conv_wgrad(tc): pointers must be 16-byte aligned");
  if (d->mode == 0 && d->ksize == 3 && d->W >= 16 && d->W % 8 == 0 && d->H % 16 == 0 &&
      !(d->cin_valid > 0 && d->cin_valid < d->Cin) && wgrad_halo_enabled())
    return conv_wgrad_halo(d, (cudaStream_t)stream);
  int bh, bn;
  int rc = tile_geometry(d->H, d->W, &bh, &bn, "conv_wgrad(tc)");
  if (rc) return rc;
  const int taps = d->mode == 0 ? d->ksize * d->ksize : 4;
  TcWgradParams p{};
  p.mode = d->mode; p.ksize = d->mode == 0 ? d->ksize : 1;
  p.H = d->H; p.W = d->W; p.Cout = d->Cout; p.Cin = d->Cin; p.x_ld = d->x_ld;
  const long long M = (long long)d->B * d->H * d->W;
  p.k_tiles = (int)((M + TC_BM - 1) / TC_BM);
  p.dw = d->dw;
  p.s_tap = d->s_tap; p.s_co = d->s_co; p.s_ci = d->s_ci;
  if (p.s_tap == 0 && p.s_co == 0 && p.s_ci == 0) { p.s_tap = (long long)d->Cout * d->Cin; p.s_co = d->Cin; p.s_ci = 1; }
  p.cin_valid = (d->cin_valid > 0 && d->cin_valid < d->Cin) ? d->cin_valid : d->Cin;
  p.vec_ok = (p.s_ci == 1 && p.s_co % 4 == 0 && p.s_tap % 4 == 0 && p.cin_valid == d->Cin &&
              ((uintptr_t)d->dw & 15) == 0) ? 1 : 0;
  const int n_tile = (d->Cin % 128 == 0) ? 128 : 64;
  const int co_tiles = (d->Cout + TC_BM - 1) / TC_BM, ci_tiles = d->Cin / n_tile;
  const int base_ctas = co_tiles * ci_tiles * taps;
  int splits = (num_sms() + base_ctas / 2) / base_ctas;   // about one wave of CTAs
  if (splits > p.k_tiles) splits = p.k_tiles;
  if (splits < 1) splits = 1;
  while (taps * splits > 65535) --splits;
  p.splits = splits;

  CUtensorMap tmDY, tmX;
  rc = make_act_map(&tmDY, 0, d->dy, d->dy_ld, d->Cout, d->B, d->H, d->W, bh, bn, "conv_wgrad(tc) dY");
  if (rc) return rc;
  rc = make_act_map(&tmX, d->mode == 1 ? 1 : 0, d->x, d->x_ld, d->Cin, d->B, d->H, d->W, bh, bn, "conv_wgrad(tc) X");
  if (rc) return rc;
  dim3 grid(co_tiles, ci_tiles, taps * splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_tile == 128) return launch_wgrad_tc<128, 3>(tmDY, tmX, p, grid, st);
  return launch_wgrad_tc<64, 4>(tmDY, tmX, p, grid, st);
}

}  // namespace b200dm

// =====================================================================================================
// Diagnostic: issue rate of tcgen05.mma (M=128, N in {64,128,256}, K=16) on resident shared-memory
// operands, no TMA, no epilogue — the ceiling the conv kernels' main loops are measured against
// (scripts/umma_rate.py).  mode 0: K-major SW128 operands as in conv_tc_kernel; mode 1: A window shifted by
// one 128-B row with a 2048-B group stride as in conv3x3_halo_kernel; mode 2: MN-major as in wgrad.
// =====================================================================================================
namespace b200dm {
template <int N_TILE>
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int iters, int mode, long long* __restrict__ out_cycles) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 64 * 1024;     // 64 KiB A region, up to 32 KiB B region
  const uint32_t bar = b_base + 64 * 1024, tmem_slot = bar + 8;      // bar + 16: scratch barrier of modes 3 / 4
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x * 16; i < 128 * 1024; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)) + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 16, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // everything above is on-chip setup; global memory is touched only below
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = mode == 2 ? make_idesc_bf16(TC_BM, N_TILE, 1, 1) : make_idesc_bf16(TC_BM, N_TILE, 0, 0);
    const uint64_t a_desc0 = mode == 2 ? make_smem_desc(a_base, 16384, 1024)
                                       : make_smem_desc(a_base + (mode == 1 ? 128 : 0), 16, mode == 1 ? 2048 : 1024);
    const uint64_t b_desc0 = mode == 2 ? make_smem_desc(b_base, 16384, 1024) : make_smem_desc(b_base, 16, 1024);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    const uint32_t a_lo = (uint32_t)a_desc0, b_lo = (uint32_t)b_desc0;
    const long long t0 = clock64();
    if (leader && mode >= 3) {
      // modes 3 / 4: the issue pattern of one conv3x3_halo tile with resident weights (N = 64): 36 MMAs over nine
      // shifted halo windows and nine weight tiles, followed (mode 3) by the two tcgen05.commit of a tile boundary;
      // `iters` tiles back to back.  Cycles are reported per MMA like the other modes.
      const uint64_t ah = make_smem_desc(a_base, 16, HALO_W * 128);
      const uint32_t ah_lo = (uint32_t)ah, ah_hi = (uint32_t)(ah >> 32);
      const uint32_t bar2 = bar + 16;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + (uint32_t)((it & 1) * 64), ah_lo + (uint32_t)(((dx + HALO_W * dy) * 128 + k * 32) >> 4),
                           ah_hi, b_lo + (uint32_t)(((tap % 7) * 8192 + k * 32) >> 4), b_hi, idesc, (tap | k) ? 1u : 0u);
        }
        if (mode == 3) { umma_commit(bar2); umma_commit(bar2); }
      }
      umma_commit(bar);
    } else if (leader) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = mode == 2 ? (uint32_t)((k * 2048) >> 4) : (uint32_t)(((k & 3) * 32 + (k >> 2) * 16384) >> 4);
          umma_bf16_lohi(tmem_base, a_lo + off, a_hi, b_lo + off, b_hi, idesc, (it | k) ? 1u : 0u);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (leader && out_cycles) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}
}  // namespace b200dm

#ifdef B200DM_PHASE_TIMING
extern "C" int b200dm_debug_set_timing_buf(long long* buf) {
  return cudaMemcpyToSymbol(b200dm::g_tbuf, &buf, sizeof(buf)) == cudaSuccess ? 0 : -4;
}
#endif

extern "C" int b200dm_debug_umma_rate(int32_t n_tile, int32_t iters, int32_t mode, int32_t ctas, long long* out_cycles,
                                      void* stream) {
  using namespace b200dm;
  const int smem = 128 * 1024 + 1024 + 64;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_RATE(N)                                                                                       \
  do {                                                                                                       \
    cudaFuncSetAttribute(umma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);            \
    launch_k(umma_rate_kernel<N>, ctas, 128, smem, st, iters, mode, out_cycles);                                   \
  } while (0)
  if (n_tile == 64) LAUNCH_RATE(64);
  else if (n_tile == 128) LAUNCH_RATE(128);
  else if (n_tile == 256) LAUNCH_RATE(256);
  else { set_error("debug_umma_rate: n_tile must be 64, 128 or 256"); return B200DM_ERR_UNSUPPORTED; }
#undef LAUNCH_RATE
  return check_launch("debug_umma_rate");
}
