// CUDA-core implicit-GEMM convolutions: the fp32-mode path, and the path for shapes the tcgen05
// kernel does not take (init 7x7 conv, final 1x1 conv to C channels).  fp32 accumulation always.
// Reference ops replaced: nn.Conv2d at ddpm.py:96,103,160,187,213-215,252-253,304,377,413,422 and
// their autograd data/weight gradients.
#include "common.cuh"

namespace b200dm {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;
constexpr int kConvThreads = 256;

struct ConvP {
  int mode, ksize, B, H, W, Cin, Cout;  // H,W: output size (modes 0,1) / input size (mode 2)
  int x_ld, y_ld, res_ld, accumulate;
  int64_t M;  // number of GEMM rows (pixels of the iteration space)
  int N;      // GEMM columns (Cout, or 4*Cout for mode 2)
  int taps;
};

// Input pixel offset for GEMM row `p`, tap `tap`; returns -1 when the tap falls in the zero padding.
__device__ __forceinline__ int64_t in_pixel(const ConvP& c, int64_t p, int tap) {
  int ox = (int)(p % c.W);
  int64_t q = p / c.W;
  int oy = (int)(q % c.H);
  int64_t b = q / c.H;
  if (c.mode == 1) {  // 2x2 stride 2 over a [2H, 2W] input, tap = p1*2 + p2
    int iy = 2 * oy + (tap >> 1), ix = 2 * ox + (tap & 1);
    return (b * (2 * c.H) + iy) * (int64_t)(2 * c.W) + ix;
  }
  int pad = c.ksize >> 1;
  int iy = oy + tap / c.ksize - pad, ix = ox + tap % c.ksize - pad;
  if (iy < 0 || iy >= c.H || ix < 0 || ix >= c.W) return -1;
  return (b * c.H + iy) * (int64_t)c.W + ix;
}

template <typename T>
__global__ void __launch_bounds__(kConvThreads)
conv_simt_kernel(ConvP c, const T* __restrict__ x, const T* __restrict__ w,
                 const float* __restrict__ bias, T* __restrict__ y, const T* __restrict__ res) {
  pdl_prologue();
  __shared__ float As[BK][BM + PADM];
  __shared__ float Bs[BK][BN + PADM];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader: one row, 4 consecutive k
  const int ty = tid >> 4, tx = tid & 15;         // compute: 4x4 micro-tile
  float acc[4][4] = {};

  const int64_t prow = m0 + lrow;
  const int wrow = n0 + lrow;  // weight row (GEMM column index)
  // for mode 2 the GEMM column j = tap*Cout + co indexes the packed weight rows directly
  const int kblocks = c.Cin / BK;
  const int gtaps = (c.mode == 2) ? 1 : c.taps;

  for (int tap = 0; tap < gtaps; ++tap) {
    int64_t ipix = (prow < c.M) ? in_pixel(c, prow, tap) : -1;
    const T* xrow = (ipix >= 0) ? x + ipix * c.x_ld : nullptr;
    const T* wr = (wrow < c.N) ? w + ((int64_t)tap * c.N + wrow) * c.Cin : nullptr;
    for (int kb = 0; kb < kblocks; ++kb) {
      const int k0 = kb * BK + lk;
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
      if (xrow) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = Elem<T>::ld(xrow + k0 + j);
      }
      if (wr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Elem<T>::ld(wr + k0 + j);
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[lk + j][lrow] = a[j];
        Bs[lk + j][lrow] = b[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t p = m0 + ty * 4 + i;
    if (p >= c.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col >= c.N) continue;
      int co = col;
      int64_t opix = p;
      if (c.mode == 2) {  // scatter to the (2y+p1, 2x+p2) pixel of a [2H, 2W] output
        int tap = col / c.Cout;
        co = col - tap * c.Cout;
        int ox = (int)(p % c.W);
        int64_t q = p / c.W;
        int oy = (int)(q % c.H);
        int64_t b = q / c.H;
        opix = (b * (2 * c.H) + 2 * oy + (tap >> 1)) * (int64_t)(2 * c.W) + 2 * ox + (tap & 1);
      }
      float v = acc[i][j];
      if (bias) v += bias[co];
      if (res) v += Elem<T>::ld(res + opix * c.res_ld + co);
      T* yp = y + opix * c.y_ld + co;
      if (c.accumulate) v += Elem<T>::ld(yp);
      Elem<T>::st(yp, v);
    }
  }
}

// ---- weight gradient: dW[tap][co][ci] += sum_p dY[p,co] * X[in_pixel(p,tap), ci] -----------------
template <typename T>
__global__ void __launch_bounds__(kConvThreads)
wgrad_simt_kernel(ConvP c, const T* __restrict__ x, const T* __restrict__ dy, int dy_ld,
                  float* __restrict__ dw, int splits, long long s_tap, long long s_co, long long s_ci) {
  pdl_prologue();
  __shared__ float As[BK][BM + PADM];  // dY chunk: [pixel][co]
  __shared__ float Bs[BK][BN + PADM];  // X chunk:  [pixel][ci]
  const int tid = threadIdx.x;
  const int co_tiles = (c.Cout + BM - 1) / BM;
  const int co0 = (blockIdx.x % co_tiles) * BM;
  const int ci0 = (blockIdx.x / co_tiles) * BN;
  const int tap = blockIdx.y;
  const int split = blockIdx.z;
  const int lrow = tid >> 4, lc = (tid & 15) * 4;  // loader: pixel row (16), 4 consecutive channels
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};

  int64_t chunks = (c.M + BK - 1) / BK;
  int64_t per = (chunks + splits - 1) / splits;
  int64_t cbeg = split * per, cend = cbeg + per < chunks ? cbeg + per : chunks;
  for (int64_t ch = cbeg; ch < cend; ++ch) {
    int64_t p = ch * BK + lrow;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < c.M) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co0 + lc + j < c.Cout) a[j] = Elem<T>::ld(dy + p * dy_ld + co0 + lc + j);
      int64_t ipix = in_pixel(c, p, tap);
      if (ipix >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ci0 + lc + j < c.Cin) b[j] = Elem<T>::ld(x + ipix * c.x_ld + ci0 + lc + j);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lrow][lc]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(&Bs[lrow][lc]) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= c.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = ci0 + tx * 4 + j;
      if (ci >= c.Cin) continue;
      atomicAdd(dw + (int64_t)tap * s_tap + (int64_t)co * s_co + (int64_t)ci * s_ci, acc[i][j]);
    }
  }
}

// ---- column sums -------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int ld, int64_t rows, int C,
                              float* __restrict__ out, int64_t rows_per_block) {
  pdl_prologue();
  __shared__ float red[4][64];
  int c = blockIdx.x * 64 + threadIdx.x;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  if (c < C)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 4) acc += Elem<T>::ld(x + r * ld + c);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float v = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

// vectorised variant: one thread = 8 channels (16 B of bf16 / 32 B of fp32) of a strided subset of rows.
// 256 threads = VL vector lanes x (256 / VL) row lanes, VL = min(C/8, 64) rounded up to a power of two.
template <typename T>
__global__ void __launch_bounds__(256)
colsum8_kernel(const T* __restrict__ x, int ld, int64_t rows, int C8, float* __restrict__ out,
               int64_t rows_per_block, int VL) {
  pdl_prologue();
  __shared__ float red[256][8];
  const int vl = threadIdx.x % VL, rl = threadIdx.x / VL, RL = 256 / VL;
  const int cv = blockIdx.x * VL + vl;  // 8-channel vector index
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8] = {};
  if (cv < C8)
    for (int64_t r = r0 + rl; r < r1; r += 8 * RL) {   // eight independent loads in flight per thread
      float v[8][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r + u * RL < r1) {
          ld8(x + (r + u * RL) * ld + cv * 8, v[u]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc[j] += ((v[0][j] + v[1][j]) + (v[2][j] + v[3][j])) + ((v[4][j] + v[5][j]) + (v[6][j] + v[7][j]));
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (rl == 0 && cv < C8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int k = 0; k < RL; ++k) s += red[k * VL + vl][j];
      atomicAdd(out + cv * 8 + j, s);
    }
  }
}

// batched variant: up to COLSUM_MAX_ITEMS independent column sums in ONE launch (the bias gradients of a gradient
// bucket: each is a few microseconds of HBM traffic, so one launch per tensor is all launch latency).  The block
// range [first[i], first[i+1]) belongs to item i and is split into column blocks x row blocks like colsum8.
constexpr int COLSUM_MAX_ITEMS = 16;
struct ColsumBatch {
  const void* x[COLSUM_MAX_ITEMS];
  float* out[COLSUM_MAX_ITEMS];
  long long rows[COLSUM_MAX_ITEMS];
  int per[COLSUM_MAX_ITEMS];    // rows per row block
  int ld[COLSUM_MAX_ITEMS];
  int C8[COLSUM_MAX_ITEMS];
  int VL[COLSUM_MAX_ITEMS];
  int cb[COLSUM_MAX_ITEMS];     // column blocks
  int first[COLSUM_MAX_ITEMS + 1];
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256)
colsum8_batched_kernel(const __grid_constant__ ColsumBatch bt) {
  pdl_prologue();
  __shared__ float red[256][8];
  int it = 0;
  while (it + 1 < bt.n && (int)blockIdx.x >= bt.first[it + 1]) ++it;
  const int local = (int)blockIdx.x - bt.first[it];
  const int cbs = bt.cb[it], VL = bt.VL[it], C8 = bt.C8[it], ld = bt.ld[it];
  const int bx = local % cbs, by = local / cbs;
  const T* x = reinterpret_cast<const T*>(bt.x[it]);
  float* out = bt.out[it];
  const long long rows = bt.rows[it];
  const int vl = threadIdx.x % VL, rl = threadIdx.x / VL, RL = 256 / VL;
  const int cv = bx * VL + vl;
  const long long r0 = (long long)by * bt.per[it];
  const long long r1 = r0 + bt.per[it] < rows ? r0 + bt.per[it] : rows;
  float acc[8] = {};
  if (cv < C8)
    for (long long r = r0 + rl; r < r1; r += 8 * RL) {
      float v[8][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r + u * RL < r1) {
          ld8(x + (r + u * RL) * ld + cv * 8, v[u]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc[j] += ((v[0][j] + v[1][j]) + (v[2][j] + v[3][j])) + ((v[4][j] + v[5][j]) + (v[6][j] + v[7][j]));
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (rl == 0 && cv < C8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sum = 0.f;
      for (int k = 0; k < RL; ++k) sum += red[k * VL + vl][j];
      atomicAdd(out + cv * 8 + j, sum);
    }
  }
}

// ---- init conv 7x7 (C -> Cout=64), NCHW fp32 in, NHWC out -----------------------------------------
constexpr int IC_ROWS = 4;   // output rows per CTA
constexpr int IC_SEG = 32;   // output columns per CTA
template <typename T>
__global__ void __launch_bounds__(256)
init_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                 const float* __restrict__ bias, T* __restrict__ y, int y_ld, int B, int C, int H,
                 int W) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int K = C * 49;
  float* wsm = sm;                          // [K][64]
  float* patch = sm + K * 64;               // [C][IC_ROWS+6][IC_SEG+6]
  const int PW = IC_SEG + 6, PH = IC_ROWS + 6;
  const int tid = threadIdx.x;
  const int segs = (W + IC_SEG - 1) / IC_SEG;
  const int seg = blockIdx.x % segs;
  const int rowblk = blockIdx.x / segs;
  const int b = blockIdx.y;
  const int y0 = rowblk * IC_ROWS, x0 = seg * IC_SEG;
  for (int i = tid; i < K * 64; i += 256) {
    int co = i & 63, k = i >> 6;
    wsm[i] = w[co * K + k];
  }
  for (int i = tid; i < C * PH * PW; i += 256) {
    int px = i % PW, py = (i / PW) % PH, ch = i / (PW * PH);
    int iy = y0 + py - 3, ix = x0 + px - 3;
    patch[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                   ? x[(((int64_t)b * C + ch) * H + iy) * W + ix] : 0.f;
  }
  __syncthreads();
  const int co = tid & 63, oct = tid >> 6;  // 4 octets of 8 pixels per row segment
  const float bv = bias ? bias[co] : 0.f;
  for (int r = 0; r < IC_ROWS; ++r) {
    if (y0 + r >= H) break;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bv;
    for (int ch = 0; ch < C; ++ch)
      for (int ky = 0; ky < 7; ++ky) {
        const float* pr = patch + (ch * PH + r + ky) * PW + oct * 8;
        float pv[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) pv[i] = pr[i];
        const float* wk = wsm + ((ch * 7 + ky) * 7) * 64 + co;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          float wv = wk[kx * 64];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(pv[i + kx], wv, acc[i]);
        }
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int ox = x0 + oct * 8 + i;
      if (ox < W) Elem<T>::st(y + (((int64_t)b * H + y0 + r) * W + ox) * y_ld + co, acc[i]);
    }
  }
}

// dW[co][c][ky][kx] += sum dY[b,y,x,co] * X[b,c,y+ky-3,x+kx-3].  Persistent CTAs walk (sample, row-block)
// items; thread = 4 output channels x up to 10 filter taps (register micro-tile: 40 FMA per 11 smem
// loads); one atomic flush per CTA at the end.
template <typename T>
__global__ void __launch_bounds__(256)
init_conv_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, int dy_ld,
                       float* __restrict__ dw, int B, int C, int H, int W, int Ctot, int c0) {
  // handles input channels [c0, c0 + C) of a Ctot-channel stem (C <= 3 per launch)
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int PW = W + 6, PH = IC_ROWS + 6;
  float* gsm = sm;                            // [IC_ROWS][W][64] output gradients of the item
  float* patch = sm + IC_ROWS * W * 64;       // [C][PH][PW]
  const int K = C * 49;
  const int tid = threadIdx.x, co0 = (tid & 15) * 4, kq = tid >> 4;
  const int rowblocks = (H + IC_ROWS - 1) / IC_ROWS;
  const int items = B * rowblocks;
  constexpr int MAXK = 10;  // ceil(147 / 16); C <= 3
  int koff[MAXK];           // patch offset of tap k relative to the output pixel
#pragma unroll
  for (int i = 0; i < MAXK; ++i) {
    int k = kq + 16 * i;
    if (k < K) {
      int kx = k % 7, ky = (k / 7) % 7, ch = k / 49;
      koff[i] = (ch * PH + ky) * PW + kx;
    } else {
      koff[i] = -1;
    }
  }
  float acc[MAXK][4];
#pragma unroll
  for (int i = 0; i < MAXK; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / rowblocks, y0 = (item - b * rowblocks) * IC_ROWS;
    __syncthreads();
    for (int i = tid; i < C * PH * PW; i += 256) {
      int px = i % PW, py = (i / PW) % PH, ch = i / (PW * PH);
      int iy = y0 + py - 3, ix = px - 3;
      patch[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                     ? x[(((int64_t)b * Ctot + c0 + ch) * H + iy) * W + ix] : 0.f;
    }
    for (int i = tid; i < IC_ROWS * W * 64; i += 256) {
      int c = i & 63, ox = (i >> 6) % W, r = (i >> 6) / W;
      gsm[i] = (y0 + r < H) ? Elem<T>::ld(dy + (((int64_t)b * H + y0 + r) * W + ox) * dy_ld + c) : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < IC_ROWS; ++r)
      for (int ox = 0; ox < W; ++ox) {
        const float4 g = *reinterpret_cast<const float4*>(&gsm[(r * W + ox) * 64 + co0]);
        const float* pb = patch + r * PW + ox;
#pragma unroll
        for (int i = 0; i < MAXK; ++i) {
          if (koff[i] >= 0) {
            const float pv = pb[koff[i]];
            acc[i][0] = fmaf(g.x, pv, acc[i][0]);
            acc[i][1] = fmaf(g.y, pv, acc[i][1]);
            acc[i][2] = fmaf(g.z, pv, acc[i][2]);
            acc[i][3] = fmaf(g.w, pv, acc[i][3]);
          }
        }
      }
  }
#pragma unroll
  for (int i = 0; i < MAXK; ++i) {
    int k = kq + 16 * i;
    if (k < K) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dw + (co0 + j) * (Ctot * 49) + c0 * 49 + k, acc[i][j]);
    }
  }
}

// ---- final 1x1 conv (Cin -> C<=4), NHWC in, NCHW fp32 out ------------------------------------------
// Forward and data gradient: one thread per pixel (its Cin-channel row is read / written as 16-byte vectors).
// Parameter gradient: a thread owns one 8-channel vector (16 B of bf16) of a pixel, the Cin/8 vector lanes of a pixel
// sit in one warp (coalesced rows), FC_U pixels in flight per thread.  (Measured: the vector-lane layout is 4x faster
// for the reduction, but slower than one-thread-per-pixel for the forward / dx streams at the sampling batch.)
constexpr int FC_U = 4;

template <typename T, bool CIN64>
__global__ void __launch_bounds__(256)
final_conv_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ w,
                  const float* __restrict__ bias, float* __restrict__ y, int64_t total, int HW,
                  int Cin, int C) {
  pdl_prologue();
  extern __shared__ float wsm[];  // [C][Cin]
  for (int i = threadIdx.x; i < C * Cin; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const T* xr = x + p * x_ld;
  if (CIN64) {       // the UNet's head (dim = 64): the whole 64-channel row is requested before the first use
    float v[8][8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) ld8(xr + kk * 8, v[kk]);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c] = fmaf(v[kk][j], wsm[c * 64 + kk * 8 + j], acc[c]);
        }
  } else {
    for (int k = 0; k < Cin; k += 8) {
      float v[8];
      ld8(xr + k, v);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c] = fmaf(v[j], wsm[c * Cin + k + j], acc[c]);
        }
    }
  }
  int64_t b = p / HW;
  int pix = (int)(p - b * HW);
  for (int c = 0; c < C; ++c) y[(b * C + c) * HW + pix] = acc[c] + (bias ? bias[c] : 0.f);
}

template <typename T>
__global__ void __launch_bounds__(256)
final_conv_dx_kernel(const float* __restrict__ w, const float* __restrict__ dy, T* __restrict__ dx,
                     int dx_ld, int64_t total, int HW, int Cin, int C) {
  pdl_prologue();
  extern __shared__ float wsm[];
  for (int i = threadIdx.x; i < C * Cin; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int64_t b = p / HW;
  int pix = (int)(p - b * HW);
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) g[c] = dy[(b * C + c) * HW + pix];
  T* dr = dx + p * dx_ld;
  for (int k = 0; k < Cin; k += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) s = fmaf(g[c], wsm[c * Cin + k + j], s);
      v[j] = s;
    }
    st8(dr + k, v);
  }
}

// dW[c][ci] += sum_p dy[b,c,p]*x[p,ci];  db[c] += sum dy.  One wave of CTAs, each over a contiguous pixel range.
template <typename T>
__global__ void __launch_bounds__(256)
final_conv_dw_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ dy, float* __restrict__ dw,
                     float* __restrict__ db, int total, int HW, int Cin, int C, int per_block) {
  pdl_prologue();
  extern __shared__ float red[];  // [PL][4][Cin] + [PL][4]
  const int C8 = Cin >> 3, PL = 256 / C8;
  const int vl = threadIdx.x % C8, pl = threadIdx.x / C8;
  const int p0 = blockIdx.x * per_block, p1 = min(p0 + per_block, total);
  float acc[4][8] = {}, accb[4] = {};
  for (int pb = p0 + pl; pb < p1; pb += PL * FC_U) {
    float v[FC_U][8], g[FC_U][4];
#pragma unroll
    for (int u = 0; u < FC_U; ++u) {
      const int p = pb + u * PL;
      const bool ok = p < p1;
      const int b = p / HW, pix = p - b * HW;
      if (ok) {
        ld8(x + (int64_t)p * x_ld + vl * 8, v[u]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) g[u][c] = (ok && c < C) ? dy[((int64_t)b * C + c) * HW + pix] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < FC_U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        accb[c] += g[u][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c][j] = fmaf(g[u][c], v[u][j], acc[c][j]);
      }
  }
  float* redb = red + PL * 4 * Cin;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(pl * 4 + c) * Cin + vl * 8 + j] = acc[c][j];
    if (vl == 0) redb[pl * 4 + c] = accb[c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cin; i += 256) {
    const int c = i / Cin, ci = i - c * Cin;
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += red[(l * 4 + c) * Cin + ci];
    atomicAdd(dw + i, s);
  }
  if (db != nullptr && threadIdx.x < C) {
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += redb[l * 4 + threadIdx.x];
    atomicAdd(db + threadIdx.x, s);
  }
}

// ---- nearest x2 upsample and its gradient ------------------------------------------------------------
template <typename T>
__global__ void upsample_fwd_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld,
                                    int64_t total8, int H, int W, int C8) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % C8);
    int64_t p = i / C8;  // output pixel over [B, 2H, 2W]
    int ox = (int)(p % (2 * W));
    int64_t q = p / (2 * W);
    int oy = (int)(q % (2 * H));
    int64_t b = q / (2 * H);
    float v[8];
    ld8(x + ((b * H + (oy >> 1)) * (int64_t)W + (ox >> 1)) * x_ld + c8 * 8, v);
    st8(y + p * y_ld + c8 * 8, v);
  }
}
template <typename T>
__global__ void upsample_bwd_kernel(const T* __restrict__ dy, int dy_ld, T* __restrict__ dx,
                                    int dx_ld, int64_t total8, int H, int W, int C8) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % C8);
    int64_t p = i / C8;  // input pixel over [B, H, W]
    int ix = (int)(p % W);
    int64_t q = p / W;
    int iy = (int)(q % H);
    int64_t b = q / H;
    float s[8] = {};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float v[8];
      int64_t op = (b * (2 * H) + 2 * iy + (d >> 1)) * (int64_t)(2 * W) + 2 * ix + (d & 1);
      ld8(dy + op * dy_ld + c8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
    st8(dx + p * dx_ld + c8 * 8, s);
  }
}

template <typename T>
static int launch_conv(const b200dm_conv_desc* d, cudaStream_t st) {
  ConvP c{};
  c.mode = d->mode; c.ksize = d->ksize; c.B = d->B; c.H = d->H; c.W = d->W;
  c.Cin = d->Cin; c.Cout = d->Cout; c.x_ld = d->x_ld; c.y_ld = d->y_ld; c.res_ld = d->res_ld;
  c.accumulate = d->accumulate;
  c.M = (int64_t)d->B * d->H * d->W;
  c.taps = d->mode == 0 ? d->ksize * d->ksize : 4;
  c.N = d->mode == 2 ? 4 * d->Cout : d->Cout;
  dim3 grid((unsigned)((c.M + BM - 1) / BM), (unsigned)((c.N + BN - 1) / BN));
  launch_k(conv_simt_kernel<T>, grid, kConvThreads, 0, st, c, (const T*)d->x, (const T*)d->w, d->bias,
                                                      (T*)d->y, (const T*)d->res);
  count_launch();
  return check_launch("conv_simt");
}

int conv_fwd_simt(const b200dm_conv_desc* d, void* stream) {
  B200DM_REQUIRE(d->Cin % BK == 0, B200DM_ERR_SHAPE, "conv(simt): Cin=%d must be a multiple of %d", d->Cin, BK);
  if (d->dtype == B200DM_F32) return launch_conv<float>(d, (cudaStream_t)stream);
  return launch_conv<__nv_bfloat16>(d, (cudaStream_t)stream);
}

template <typename T>
static int launch_wgrad(const b200dm_wgrad_desc* d, cudaStream_t st) {
  ConvP c{};
  c.mode = d->mode; c.ksize = d->ksize; c.B = d->B; c.H = d->H; c.W = d->W;
  c.Cin = d->Cin; c.Cout = d->Cout; c.x_ld = d->x_ld;
  c.M = (int64_t)d->B * d->H * d->W;
  c.taps = d->mode == 0 ? d->ksize * d->ksize : 4;
  c.N = d->Cout;
  int tiles = ((c.Cout + BM - 1) / BM) * ((c.Cin + BN - 1) / BN);
  int64_t chunks = (c.M + BK - 1) / BK;
  int splits = (4 * num_sms() + tiles * c.taps - 1) / (tiles * c.taps);
  if (splits > chunks / 8) splits = (int)(chunks / 8);
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  dim3 grid(tiles, c.taps, splits);
  long long s_tap = d->s_tap, s_co = d->s_co, s_ci = d->s_ci;
  if (s_tap == 0 && s_co == 0 && s_ci == 0) { s_tap = (long long)c.Cout * c.Cin; s_co = c.Cin; s_ci = 1; }
  launch_k(wgrad_simt_kernel<T>, grid, kConvThreads, 0, st, c, (const T*)d->x, (const T*)d->dy, d->dy_ld,
                                                       d->dw, splits, s_tap, s_co, s_ci);
  count_launch();
  return check_launch("wgrad_simt");
}

int conv_wgrad_simt(const b200dm_wgrad_desc* d, void* stream) {
  if (d->dtype == B200DM_F32) return launch_wgrad<float>(d, (cudaStream_t)stream);
  return launch_wgrad<__nv_bfloat16>(d, (cudaStream_t)stream);
}

}  // namespace b200dm

using namespace b200dm;

extern "C" int b200dm_colsum(int32_t dtype, const void* x, int32_t ld, int64_t rows, int32_t C,
                             float* out, int32_t accumulate, void* stream) {
  B200DM_REQUIRE(rows > 0 && C > 0, B200DM_ERR_SHAPE, "colsum: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    int rc = b200dm_fill_f32(out, C, 0.f, stream);
    if (rc) return rc;
  }
  if (C % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0) {
    const int C8 = C / 8;
    int VL = 1;
    while (VL < C8 && VL < 64) VL <<= 1;
    int cb = (C8 + VL - 1) / VL;
    int64_t rb = (num_sms() + cb - 1) / cb;   // one wave: the final atomics of every CTA hit the same C addresses
    const int RL = 256 / VL;
    if (rb > (rows + 4 * RL - 1) / (4 * RL)) rb = (rows + 4 * RL - 1) / (4 * RL);
    if (rb < 1) rb = 1;
    int64_t per8 = (rows + rb - 1) / rb;
    dim3 grid8(cb, (unsigned)rb);
    if (dtype == B200DM_F32)
      launch_k(colsum8_kernel<float>, grid8, 256, 0, st, (const float*)x, ld, rows, C8, out, per8, VL);
    else
      launch_k(colsum8_kernel<__nv_bfloat16>, grid8, 256, 0, st, (const __nv_bfloat16*)x, ld, rows, C8, out, per8, VL);
    count_launch();
    return check_launch("colsum8");
  }
  int cblocks = (C + 63) / 64;
  int64_t rblocks = (2 * num_sms() + cblocks - 1) / cblocks;
  if (rblocks > (rows + 31) / 32) rblocks = (rows + 31) / 32;
  if (rblocks < 1) rblocks = 1;
  int64_t per = (rows + rblocks - 1) / rblocks;
  dim3 grid(cblocks, (unsigned)rblocks), block(64, 4);
  if (dtype == B200DM_F32)
    launch_k(colsum_kernel<float>, grid, block, 0, st, (const float*)x, ld, rows, C, out, per);
  else
    launch_k(colsum_kernel<__nv_bfloat16>, grid, block, 0, st, (const __nv_bfloat16*)x, ld, rows, C, out, per);
  count_launch();
  return check_launch("colsum");
}

extern "C" int b200dm_colsum_batched(int32_t dtype, const b200dm_colsum_item* items, int32_t n, void* stream) {
  B200DM_REQUIRE(items != nullptr && n >= 1 && n <= COLSUM_MAX_ITEMS, B200DM_ERR_SHAPE,
                 "colsum_batched: n=%d must be 1..%d", n, COLSUM_MAX_ITEMS);
  B200DM_REQUIRE(dtype == B200DM_F32 || dtype == B200DM_BF16, B200DM_ERR_UNSUPPORTED, "colsum_batched: dtype");
  ColsumBatch bt{};
  bt.n = n;
  double total = 0.0;
  for (int i = 0; i < n; ++i) {
    const b200dm_colsum_item& q = items[i];
    B200DM_REQUIRE(q.x != nullptr && q.out != nullptr && q.rows > 0 && q.C > 0, B200DM_ERR_SHAPE,
                   "colsum_batched: item %d is empty", i);
    B200DM_REQUIRE(q.C % 8 == 0 && q.ld % 8 == 0 && q.ld >= q.C && ((uintptr_t)q.x & 15) == 0, B200DM_ERR_SHAPE,
                   "colsum_batched: item %d: C, ld must be multiples of 8, x 16-byte aligned", i);
    total += (double)q.rows * q.C;
  }
  // about two waves of CTAs, shared out by size (at least one each); every CTA wants at least 4 passes of its row lanes
  const int budget = 2 * num_sms();
  int nblk = 0;
  for (int i = 0; i < n; ++i) {
    const b200dm_colsum_item& q = items[i];
    const int C8 = q.C / 8;
    int VL = 1;
    while (VL < C8 && VL < 64) VL <<= 1;
    const int cb = (C8 + VL - 1) / VL, RL = 256 / VL;
    long long rb = (long long)((double)budget * ((double)q.rows * q.C / total) / cb + 0.5);
    const long long rbmax = (q.rows + 4 * RL - 1) / (4 * RL);
    if (rb > rbmax) rb = rbmax;
    if (rb < 1) rb = 1;
    const long long per = (q.rows + rb - 1) / rb;
    B200DM_REQUIRE(per < (1ll << 31), B200DM_ERR_SHAPE, "colsum_batched: item %d has too many rows", i);
    rb = (q.rows + per - 1) / per;
    bt.x[i] = q.x; bt.out[i] = q.out; bt.rows[i] = q.rows; bt.per[i] = (int)per; bt.ld[i] = q.ld;
    bt.C8[i] = C8; bt.VL[i] = VL; bt.cb[i] = cb;
    bt.first[i] = nblk;
    nblk += (int)(rb * cb);
  }
  bt.first[n] = nblk;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32) launch_k(colsum8_batched_kernel<float>, nblk, 256, 0, st, bt);
  else launch_k(colsum8_batched_kernel<__nv_bfloat16>, nblk, 256, 0, st, bt);
  count_launch();
  return check_launch("colsum_batched");
}

// ---- stem on tensor cores: im2col of the 7x7 patches + zero-padded weight rows -------------------------------
// P[pix][k] = x[b, ch, y + ky - 3, x + kx - 3], k = (ch*7 + ky)*7 + kx (the OIHW order of the master weight),
// columns [C*49, KP) are zero.  The 7x7 conv (ddpm.py:304) then IS a 1x1 conv with Cin = KP over P, and its
// weight gradient a plain wgrad over P — both run on the tcgen05 kernels.
namespace b200dm {
// One CTA = four output rows of one image.  The 10 input rows those need are staged once in shared memory
// (zero padding included), then thread (vector lane v, pixel slot) gathers its eight taps from the staged tile and
// writes one 16-byte vector; the CTA's output is one contiguous span of P.
constexpr int I2C_ROWS = 4;
__global__ void im2col7_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ P, int C, int H, int W, int KP) {
  pdl_prologue();
  extern __shared__ float i2c_tile[];                  // [C][I2C_ROWS + 6][W + 6]
  const int pitch = W + 6, rows = I2C_ROWS + 6, K = C * 49;
  const int tiles_per_img = (H + I2C_ROWS - 1) / I2C_ROWS;
  const int b = blockIdx.x / tiles_per_img, oy0 = (blockIdx.x - b * tiles_per_img) * I2C_ROWS;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const float* xb = x + (size_t)b * C * H * W;
  for (int i = tid; i < C * rows * pitch; i += nthr) {
    const int cx = i % pitch, t = i / pitch, r = t % rows, ch = t / rows;
    const int iy = oy0 + r - 3, ix = cx - 3;
    i2c_tile[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xb + (ch * H + iy) * W + ix) : 0.f;
  }
  const int v = threadIdx.x;
  int off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = v * 8 + j;
    const int ch = k / 49, r = k - ch * 49, ky = r / 7, kx = r - ky * 7;
    off[j] = k < K ? (ch * rows + ky) * pitch + kx : -1;
  }
  __syncthreads();
  const int npx = min(I2C_ROWS, H - oy0) * W;
  __nv_bfloat16* Pb = P + ((size_t)(b * H + oy0) * W) * KP + v * 8;
  int cx = threadIdx.y, r = 0;                          // blockDim.y (8) divides W: a slot never straddles rows
  for (int p = threadIdx.y; p < npx; p += blockDim.y) {
    const int base = r * pitch + cx;
    float val[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) val[j] = off[j] >= 0 ? i2c_tile[off[j] + base] : 0.f;
    st8(Pb + (size_t)p * KP, val);
    cx += blockDim.y;
    if (cx >= W) {
      cx -= W;
      ++r;
    }
  }
}
__global__ void pack_stem_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int K, int KP) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Cout * KP) {
    const int co = i / KP, k = i - co * KP;
    wp[i] = __float2bfloat16_rn(k < K ? w[co * K + k] : 0.f);
  }
}
}  // namespace b200dm

extern "C" int b200dm_im2col7(const float* x, void* P, int32_t B, int32_t C, int32_t H, int32_t W, int32_t KP,
                              void* stream) {
  B200DM_REQUIRE(B > 0 && C >= 1 && C * 49 <= KP && KP % 64 == 0 && ((uintptr_t)P & 15) == 0, B200DM_ERR_SHAPE,
                 "im2col7: need C*49 <= KP, KP %% 64 == 0 (C=%d KP=%d)", C, KP);
  B200DM_REQUIRE((int64_t)B * H * W < (1LL << 31) && KP / 8 <= 128, B200DM_ERR_UNSUPPORTED, "im2col7: tensor too large");
  B200DM_REQUIRE(W % 8 == 0 && W <= 1024, B200DM_ERR_SHAPE, "im2col7: W=%d must be a multiple of 8 (<= 1024)", W);
  dim3 block(KP / 8, 8);
  const size_t smem = (size_t)C * (b200dm::I2C_ROWS + 6) * (W + 6) * sizeof(float);
  const int tiles = B * ((H + b200dm::I2C_ROWS - 1) / b200dm::I2C_ROWS);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(b200dm::im2col7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_k(b200dm::im2col7_kernel, tiles, block, smem, (cudaStream_t)stream, x, (__nv_bfloat16*)P, C, H, W, KP);
  b200dm::count_launch();
  return b200dm::check_launch("im2col7");
}

extern "C" int b200dm_pack_stem_weight(const float* w, void* wp, int32_t Cout, int32_t K, int32_t KP, void* stream) {
  B200DM_REQUIRE(Cout > 0 && K > 0 && K <= KP, B200DM_ERR_SHAPE, "pack_stem_weight: K=%d KP=%d", K, KP);
  launch_k(b200dm::pack_stem_kernel, (Cout * KP + 255) / 256, 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)wp, Cout, K, KP);
  b200dm::count_launch();
  return b200dm::check_launch("pack_stem_weight");
}

extern "C" int b200dm_init_conv_fwd(int32_t dtype, const float* x, const float* w, const float* bias,
                                    void* y, int32_t y_ld, int32_t B, int32_t C, int32_t H, int32_t W,
                                    int32_t Cout, void* stream) {
  B200DM_REQUIRE(Cout == 64, B200DM_ERR_UNSUPPORTED, "init_conv: Cout=%d (only dim=64 is built)", Cout);
  B200DM_REQUIRE(C >= 1 && C <= 6, B200DM_ERR_UNSUPPORTED, "init_conv: channels=%d (1..6 supported)", C);
  B200DM_REQUIRE(W % 8 == 0 && H % 8 == 0, B200DM_ERR_SHAPE, "init_conv: H,W must be multiples of 8");
  size_t smem = ((size_t)C * 49 * 64 + (size_t)C * (IC_ROWS + 6) * (IC_SEG + 6)) * sizeof(float);
  int segs = (W + IC_SEG - 1) / IC_SEG;
  dim3 grid(segs * ((H + IC_ROWS - 1) / IC_ROWS), B);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32) {
    cudaFuncSetAttribute(init_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(init_conv_kernel<float>, grid, 256, smem, st, x, w, bias, (float*)y, y_ld, B, C, H, W);
  } else {
    cudaFuncSetAttribute(init_conv_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(init_conv_kernel<__nv_bfloat16>, grid, 256, smem, st, x, w, bias, (__nv_bfloat16*)y, y_ld, B, C, H, W);
  }
  count_launch();
  return check_launch("init_conv_fwd");
}

extern "C" int b200dm_init_conv_wgrad(int32_t dtype, const float* x, const void* dy, int32_t dy_ld,
                                      float* dw, int32_t B, int32_t C, int32_t H, int32_t W,
                                      int32_t Cout, void* stream) {
  B200DM_REQUIRE(Cout == 64, B200DM_ERR_UNSUPPORTED, "init_conv_wgrad: Cout=%d (only dim=64 is built)", Cout);
  B200DM_REQUIRE(C >= 1 && C <= 6, B200DM_ERR_UNSUPPORTED, "init_conv_wgrad: channels=%d (1..6)", C);
  B200DM_REQUIRE(W <= 128, B200DM_ERR_UNSUPPORTED, "init_conv_wgrad: W=%d too large", W);
  int items = B * ((H + IC_ROWS - 1) / IC_ROWS);
  int grid = items < 2 * num_sms() ? items : 2 * num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  // one launch per group of up to three input channels (a self-conditioned stem has 2 * channels of them)
  for (int c0 = 0; c0 < C; c0 += 3) {
    const int cg = C - c0 < 3 ? C - c0 : 3;
    size_t smem = ((size_t)cg * (IC_ROWS + 6) * (W + 6) + (size_t)IC_ROWS * W * 64) * sizeof(float);
    if (dtype == B200DM_F32) {
      cudaFuncSetAttribute(init_conv_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_k(init_conv_wgrad_kernel<float>, grid, 256, smem, st, x, (const float*)dy, dy_ld, dw, B, cg, H, W, C, c0);
    } else {
      cudaFuncSetAttribute(init_conv_wgrad_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_k(init_conv_wgrad_kernel<__nv_bfloat16>, grid, 256, smem, st, x, (const __nv_bfloat16*)dy, dy_ld, dw, B, cg, H, W, C, c0);
    }
    count_launch();
  }
  return check_launch("init_conv_wgrad");
}

// Cin/8 vector lanes must tile a warp: Cin in {8, 16, 32, 64, 128, 256}
static bool final_conv_cin_ok(int Cin) { return Cin >= 8 && Cin <= 256 && (Cin & (Cin - 1)) == 0; }

extern "C" int b200dm_final_conv_fwd(int32_t dtype, const void* x, int32_t x_ld, const float* w,
                                     const float* bias, float* y, int32_t B, int32_t HW, int32_t Cin,
                                     int32_t C, void* stream) {
  B200DM_REQUIRE(C >= 1 && C <= 4 && Cin % 8 == 0, B200DM_ERR_UNSUPPORTED, "final_conv: C=%d Cin=%d", C, Cin);
  int64_t total = (int64_t)B * HW;
  size_t smem = (size_t)C * Cin * sizeof(float);
  unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(final_conv_kernel<float, false>, grid, 256, smem, st, (const float*)x, x_ld, w, bias, y, total, HW, Cin, C);
  else if (Cin == 64)
    launch_k(final_conv_kernel<__nv_bfloat16, true>, grid, 256, smem, st, (const __nv_bfloat16*)x, x_ld, w, bias, y, total, HW, Cin, C);
  else
    launch_k(final_conv_kernel<__nv_bfloat16, false>, grid, 256, smem, st, (const __nv_bfloat16*)x, x_ld, w, bias, y, total, HW, Cin, C);
  count_launch();
  return check_launch("final_conv_fwd");
}

extern "C" int b200dm_final_conv_bwd(int32_t dtype, const void* x, int32_t x_ld, const float* w,
                                     const float* dy, void* dx, int32_t dx_ld, float* dw, float* db,
                                     int32_t B, int32_t HW, int32_t Cin, int32_t C, void* stream) {
  B200DM_REQUIRE(C >= 1 && C <= 4 && final_conv_cin_ok(Cin), B200DM_ERR_UNSUPPORTED,
                 "final_conv_bwd: C=%d Cin=%d", C, Cin);
  B200DM_REQUIRE(!dw || (x_ld % 8 == 0 && ((uintptr_t)x & 15) == 0), B200DM_ERR_SHAPE,
                 "final_conv_bwd: x must be 16-byte aligned, ld a multiple of 8");
  const int64_t total = (int64_t)B * HW;
  B200DM_REQUIRE(total > 0 && total < (1ll << 31), B200DM_ERR_SHAPE, "final_conv_bwd: B*HW out of range");
  size_t smem = (size_t)C * Cin * sizeof(float);
  unsigned grid = (unsigned)((total + 255) / 256);
  const int PL = 256 / (Cin / 8);
  cudaStream_t st = (cudaStream_t)stream;
  // parameter gradients: one wave of CTAs (the atomics of every CTA hit the same C*Cin addresses)
  int64_t nblk = num_sms();
  if (nblk > (total + PL * FC_U - 1) / (PL * FC_U)) nblk = (total + PL * FC_U - 1) / (PL * FC_U);
  const int per = (int)((total + nblk - 1) / nblk);
  nblk = (total + per - 1) / per;
  size_t smem2 = (size_t)(PL * 4 * Cin + PL * 4) * sizeof(float);
  // dx == nullptr or dw == nullptr skips that kernel (the plan issues the parameter gradients on its side stream)
  if (dtype == B200DM_F32) {
    if (dx) launch_k(final_conv_dx_kernel<float>, grid, 256, smem, st, w, dy, (float*)dx, dx_ld, total, HW, Cin, C);
    if (dw) launch_k(final_conv_dw_kernel<float>, (unsigned)nblk, 256, smem2, st, (const float*)x, x_ld, dy, dw, db, (int)total, HW, Cin, C, per);
  } else {
    if (dx) launch_k(final_conv_dx_kernel<__nv_bfloat16>, grid, 256, smem, st, w, dy, (__nv_bfloat16*)dx, dx_ld, total, HW, Cin, C);
    if (dw) launch_k(final_conv_dw_kernel<__nv_bfloat16>, (unsigned)nblk, 256, smem2, st, (const __nv_bfloat16*)x, x_ld, dy, dw, db, (int)total, HW, Cin, C, per);
  }
  count_launch((dx ? 1 : 0) + (dw ? 1 : 0));
  return check_launch("final_conv_bwd");
}

extern "C" int b200dm_upsample2x_fwd(int32_t dtype, const void* x, int32_t x_ld, void* y, int32_t y_ld,
                                     int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  B200DM_REQUIRE(C % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0, B200DM_ERR_SHAPE, "upsample: C, ld must be multiples of 8");
  int64_t total8 = (int64_t)B * 2 * H * 2 * W * (C / 8);
  int64_t blocks = (total8 + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(upsample_fwd_kernel<float>, (unsigned)blocks, 256, 0, st, (const float*)x, x_ld, (float*)y, y_ld, total8, H, W, C / 8);
  else
    launch_k(upsample_fwd_kernel<__nv_bfloat16>, (unsigned)blocks, 256, 0, st, (const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, total8, H, W, C / 8);
  count_launch();
  return check_launch("upsample2x_fwd");
}

extern "C" int b200dm_upsample2x_bwd(int32_t dtype, const void* dy, int32_t dy_ld, void* dx,
                                     int32_t dx_ld, int32_t B, int32_t H, int32_t W, int32_t C,
                                     void* stream) {
  B200DM_REQUIRE(C % 8 == 0 && dx_ld % 8 == 0 && dy_ld % 8 == 0, B200DM_ERR_SHAPE, "upsample_bwd: C, ld must be multiples of 8");
  int64_t total8 = (int64_t)B * H * W * (C / 8);
  int64_t blocks = (total8 + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200DM_F32)
    launch_k(upsample_bwd_kernel<float>, (unsigned)blocks, 256, 0, st, (const float*)dy, dy_ld, (float*)dx, dx_ld, total8, H, W, C / 8);
  else
    launch_k(upsample_bwd_kernel<__nv_bfloat16>, (unsigned)blocks, 256, 0, st, (const __nv_bfloat16*)dy, dy_ld, (__nv_bfloat16*)dx, dx_ld, total8, H, W, C / 8);
  count_launch();
  return check_launch("upsample2x_bwd");
}
