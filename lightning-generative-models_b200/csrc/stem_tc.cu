// The 7x7 stem (init_conv, ddpm.py:304,437; padding 3) as ONE tensor-core launch: the im2col patches are built in
// shared memory, never in HBM.
//
// The GEMM formulation of round 1 (b200dm_im2col7 + conv_tc_kernel) writes a [pixels][K] bf16 patch matrix
// (K = C*49 padded to a multiple of 64: 384 B per pixel for RGB) and reads it back: 0.8 GB per UNet evaluation at the
// DDIM benchmark shape, 102 + 91 us.  Here a CTA walks 128-pixel tiles (128 / W image rows):
//   warps 1..8  builders : stage the (rows + 6) x (W + 6) x C fp32 input window (zero padded), then write the tile's
//                          A operand [128 px][K] as 128-byte-swizzled bf16 blocks by hand (fence.proxy.async)
//   warp 0      MMA      : D[128 px][64] = A W^T, K / 16 tcgen05.mma, two A buffers and two TMEM accumulators
//   warps 9..12 epilogue : tcgen05.ld -> + bias -> bf16 -> staged rows -> coalesced 16-byte stores
// K order: one 16-byte chunk of the A operand = one filter ROW (channel, ky) = 7 taps + a zero: k = (c*7 + ky)*8 + kx
// (b200dm_pack_stem_rows packs the weights the same way, K = C*56 padded to a multiple of 64).  A builder thread owns a
// pixel, so the 32 lanes of a warp read 32 consecutive window elements per tap: conflict-free (the OIHW order of the
// im2col GEMM made every chunk straddle filter rows: per-tap offset tables and 2-3-way bank conflicts, 166 us).
// The weights are TMA-loaded once per CTA.  Used by inference plans; training needs the patch matrix for the weight
// gradient and keeps the GEMM path (im2col 103 + GEMM 87 us at the DDIM shape).
#include "tc_common.cuh"

namespace b200dm {

using namespace tc;

int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const char* what);   // conv_tc.cu
bool tc_supported();

namespace {

constexpr int ST_BLK = 128 * 64 * 2;       // A block [128 px][64 k] bf16, SWIZZLE_128B: 16 KiB
constexpr int ST_WBLK = 64 * 64 * 2;       // W block [64 co][64 k]: 8 KiB
constexpr int ST_BUILD = 256;              // builder threads
constexpr int ST_THREADS = 32 + ST_BUILD + 128;
constexpr int ST_MAXKB = 5;                // K <= 320 (6 input channels)

struct StemParams {
  const float* x;
  const float* bias;
  __nv_bfloat16* y;
  int y_ld, B, C, H, W, KB, rows;          // rows = 128 / W image rows per tile
};

__device__ __forceinline__ uint32_t st_pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(ST_THREADS, 1)
stem7_tc_kernel(const __grid_constant__ CUtensorMap tmW, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int KB = p.KB;
  const int OFF_A = KB * ST_WBLK;                       // two A buffers of KB blocks
  const int OFF_STG = OFF_A + 2 * KB * ST_BLK;          // output staging [128 rows][128 B]
  const int OFF_BAR = OFF_STG + ST_BLK;
  const uint32_t w_full = base + OFF_BAR;
  auto a_full = [&](int s) { return w_full + 8u + 8u * s; };
  auto a_empty = [&](int s) { return w_full + 24u + 8u * s; };
  auto d_full = [&](int s) { return w_full + 40u + 8u * s; };
  auto d_empty = [&](int s) { return w_full + 56u + 8u * s; };
  const uint32_t tmem_slot = w_full + 72u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + OFF_BAR + 72);
  float* bias_s = reinterpret_cast<float*>(base_ptr + OFF_BAR + 128);       // [64]
  float* win = bias_s + 64;                                                  // [2 buffers][C][rows + 6][W + 6]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int PW = p.W + 6, PH = p.rows + 6;
  const int tiles_per_img = p.H / p.rows;
  const long long T = (long long)p.B * tiles_per_img;
  const int tile0 = (int)(T * blockIdx.x / gridDim.x), tile1 = (int)(T * (blockIdx.x + 1) / gridDim.x);
  const int nt = tile1 - tile0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), ST_BUILD / 32);
      mbar_init(a_empty(s), 1);
      mbar_init(d_full(s), 1);
      mbar_init(d_empty(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== weights + MMA issue =====================
    const bool leader = elect_one();
    if (leader && nt > 0) {
      mbar_expect_tx(w_full, KB * ST_WBLK);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(base + kb * ST_WBLK, &tmW, w_full, kb * 64, 0, 0);
    }
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint64_t desc0 = make_smem_desc(base, 16, 1024);
    const uint32_t hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
    if (nt > 0) {
      mbar_wait(w_full, 0);
      tc_fence_after();
    }
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      const uint32_t par = (uint32_t)((j >> 1) & 1);
      mbar_wait(a_full(s), par);
      mbar_wait(d_empty(s), par ^ 1u);
      tc_fence_after();
      if (leader) {
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base + (uint32_t)(s * 64), lo0 + (uint32_t)((OFF_A + (s * KB + kb) * ST_BLK + k * 32) >> 4), hi,
                           lo0 + (uint32_t)((kb * ST_WBLK + k * 32) >> 4), hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(a_empty(s));
        umma_commit(d_full(s));
      }
      __syncwarp();
    }
  } else if (warp <= ST_BUILD / 32) {
    // ===================== builders =====================
    // thread = (pixel of the tile, chunk parity): chunk c = filter row (channel c / 7, ky = c % 7), chunks >= 7 C are padding
    const int tb = threadIdx.x - 32;
    const int px = tb & 127, chalf = tb >> 7;
    const int nchunk = KB * 8, nrows = p.C * 7;
    const int wsh = 31 - __clz(p.W);                     // W is a power of two
    const int pbase = (px >> wsh) * PW + (px & (p.W - 1));
    const int bw = tb >> 5, bl = tb & 31;
    // input window of tile j into window buffer j & 1: one warp per (channel, window row), 4-byte cp.async (the rows
    // start 3 pixels left of the image: no wider alignment), zero padding by plain stores
    const int wsz = p.C * PH * PW;
    auto load_window = [&](int j) {
      const int tile = tile0 + j;
      const int b = tile / tiles_per_img, y0 = (tile - b * tiles_per_img) * p.rows;
      float* wb = win + (j & 1) * wsz;
      const float* xb = p.x + (size_t)b * p.C * p.H * p.W;
      for (int rr = bw; rr < p.C * PH; rr += ST_BUILD / 32) {
        const int ch = rr / PH, r = rr - ch * PH;
        const int iy = y0 + r - 3;
        const bool rowok = iy >= 0 && iy < p.H;
        const float* src = xb + ((size_t)ch * p.H + (rowok ? iy : 0)) * p.W;
        for (int cx = bl; cx < PW; cx += 32) {
          const int ix = cx - 3;
          if (rowok && ix >= 0 && ix < p.W)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(wb + rr * PW + cx)), "l"(src + ix) : "memory");
          else
            wb[rr * PW + cx] = 0.f;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nt > 0) load_window(0);
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      const float* wcur = win + (j & 1) * wsz;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      named_bar_sync(1, ST_BUILD);                       // window j complete; the gathers of tile j-1 are done
      if (j + 1 < nt) load_window(j + 1);                // in flight while this tile is gathered
      mbar_wait(a_empty(s), (uint32_t)((j >> 1) & 1) ^ 1u);      // the MMAs of tile j-2 have read this buffer
      uint8_t* abuf = base_ptr + OFF_A + s * KB * ST_BLK;
      for (int c = chalf; c < nchunk; c += 2) {
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (c < nrows) {
          const int ch = c / 7, ky = c - ch * 7;
          const float* wr = wcur + (ch * PH + ky) * PW + pbase;
          out = make_uint4(st_pack(wr[0], wr[1]), st_pack(wr[2], wr[3]), st_pack(wr[4], wr[5]), st_pack(wr[6], 0.f));
        }
        *reinterpret_cast<uint4*>(abuf + (c >> 3) * ST_BLK + px * 128 + (((c & 7) ^ (px & 7)) << 4)) = out;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(s));
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    uint8_t* stg = base_ptr + OFF_STG;
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      mbar_wait(d_full(s), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(s * 64 + c), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = st_pack(__uint_as_float(v[g * 8 + 2 * i]) + bias_s[c + g * 8 + 2 * i],
                           __uint_as_float(v[g * 8 + 2 * i + 1]) + bias_s[c + g * 8 + 2 * i + 1]);
          const int cc = c / 8 + g;
          *reinterpret_cast<uint4*>(stg + r * 128 + ((cc ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty(s));
      // this warp's 32 rows, written out coalesced (4 rows x 128 B per instruction); rows of a tile are consecutive pixels
      __nv_bfloat16* ybase = p.y + (size_t)(tile0 + j) * 128 * p.y_ld;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int id = lane + 32 * i;
        const int rl = q * 32 + (id >> 3), cc = id & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((cc ^ (rl & 7)) << 4));
        *reinterpret_cast<uint4*>(ybase + (size_t)rl * p.y_ld + cc * 8) = v;
      }
      __syncwarp();                                      // the rows are re-staged by the same warp for the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// wp[co][(c*7 + ky)*8 + kx] = bf16(w[co][c][ky][kx]) for kx < 7, zero elsewhere (row padding and K padding)
__global__ void stem_pack_rows_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int C, int KP) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Cout * KP) {
    const int co = i / KP, k = i - co * KP;
    const int row = k >> 3, kx = k & 7;
    wp[i] = __float2bfloat16_rn((kx < 7 && row < C * 7) ? w[(co * C * 7 + row) * 7 + kx] : 0.f);
  }
}

}  // namespace
}  // namespace b200dm

using namespace b200dm;

static int stem7_smem_bytes(int C, int W, int KP) {
  const int KB = KP / 64, win = C * (128 / W + 6) * (W + 6);
  return KB * ST_WBLK + 2 * KB * ST_BLK + ST_BLK + 128 + 64 * 4 + 2 * win * 4 + 1024;
}

extern "C" int b200dm_stem7_supported(int32_t B, int32_t C, int32_t H, int32_t W, int32_t KP, int32_t y_ld) {
  if (!tc_supported()) return 0;
  if (B <= 0 || C < 1 || C * 56 > KP || KP % 64 || KP / 64 > ST_MAXKB) return 0;
  if (W < 8 || W > 128 || 128 % W || (H * W) % 128 || H % (128 / W)) return 0;
  if (y_ld % 8 || y_ld < 64) return 0;
  if ((long long)B * H * W >= (1ll << 31)) return 0;
  if (stem7_smem_bytes(C, W, KP) > 227 * 1024) return 0;      // K = 320 (six input channels): two A buffers do not fit
  return 1;
}

extern "C" int b200dm_pack_stem_rows(const float* w, void* wp, int32_t Cout, int32_t C, int32_t KP, void* stream) {
  B200DM_REQUIRE(w && wp && Cout > 0 && C > 0 && C * 56 <= KP, B200DM_ERR_SHAPE, "pack_stem_rows: C=%d KP=%d", C, KP);
  launch_k(stem_pack_rows_kernel, (Cout * KP + 255) / 256, 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)wp, (int)Cout,
           (int)C, (int)KP);
  count_launch();
  return check_launch("pack_stem_rows");
}

extern "C" int b200dm_stem7_fwd(const float* x, const void* wp, const float* bias, void* y, int32_t y_ld, int32_t B,
                                int32_t C, int32_t H, int32_t W, int32_t KP, void* stream) {
  B200DM_REQUIRE(b200dm_stem7_supported(B, C, H, W, KP, y_ld) == 1, B200DM_ERR_UNSUPPORTED,
                 "stem7_fwd: needs sm_100, C*56 <= KP, KP %% 64 == 0, W in {8..128} dividing 128, H*W %% 128 == 0, buffers in 227 KiB "
                 "(C=%d KP=%d H=%d W=%d)", C, KP, H, W);
  B200DM_REQUIRE(x && wp && y && ((uintptr_t)y & 15) == 0 && ((uintptr_t)wp & 15) == 0, B200DM_ERR_SHAPE,
                 "stem7_fwd: null or misaligned pointer");
  const int KB = KP / 64;
  CUtensorMap tmW;
  {
    cuuint64_t dims[3] = {(cuuint64_t)KP, 64, 1};
    cuuint64_t str[2] = {(cuuint64_t)KP * 2, (cuuint64_t)64 * KP * 2};
    cuuint32_t box[3] = {64, 64, 1};
    int rc = encode_map(&tmW, wp, 3, dims, str, box, "stem7 weights");
    if (rc) return rc;
  }
  StemParams p{};
  p.x = x; p.bias = bias; p.y = (__nv_bfloat16*)y; p.y_ld = y_ld; p.B = B; p.C = C; p.H = H; p.W = W;
  p.KB = KB; p.rows = 128 / W;
  const int smem = stem7_smem_bytes(C, W, KP);
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(stem7_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    B200DM_REQUIRE(e == cudaSuccess, B200DM_ERR_CUDA, "stem7_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const long long tiles = (long long)B * H * W / 128;
  const int grid = tiles < num_sms() ? (int)tiles : num_sms();
  launch_k(stem7_tc_kernel, grid, ST_THREADS, smem, (cudaStream_t)stream, tmW, p);
  count_launch();
  return check_launch("stem7_fwd");
}
