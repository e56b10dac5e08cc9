// Time embedding and the small fp32 linears: SinusoidalPosEmb ddpm.py:119-132, time_mlp :328-333,
// per-block SiLU->Linear(time_dim, 2*Cout) :179-183 (run as ONE concatenated GEMM over all blocks).
// M = batch only, so these are latency-, not throughput-, critical: CUDA-core fp32 GEMM.
#include "common.cuh"

namespace b200dm {

__global__ void sinusoidal_kernel(const int64_t* __restrict__ t, float* __restrict__ emb, int B,
                                  int dim, float neg_log_theta_over) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int half = dim / 2;
  if (i >= B * half) return;
  int b = i / half, k = i - b * half;
  // torch: exp(arange(half) * -emb) in fp32, then t[:, None] * freqs
  float f = expf((float)k * neg_log_theta_over);
  float a = (float)t[b] * f;
  emb[b * dim + k] = sinf(a);
  emb[b * dim + half + k] = cosf(a);
}

constexpr int GM = 64, GN = 64, GK = 32, GPAD = 4;

__device__ __forceinline__ float act_fwd(int act, float v) {
  if (act == 1) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
  if (act == 2) return v / (1.f + expf(-v));
  return v;
}
__device__ __forceinline__ float act_grad(int act, float v) {
  if (act == 1) {
    float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
    float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
    return cdf + v * pdf;
  }
  if (act == 2) {
    float s = 1.f / (1.f + expf(-v));
    return s * (1.f + v * (1.f - s));
  }
  return 1.f;
}

// C[M,N] (+)= A(M,K) * B(K,N).  AT: A stored [K][M] (else [M][K]);  BT: B stored [N][K] (else [K][N]).
// M is the batch, so every launch is a handful of CTAs and latency-bound: 32-deep K chunks fetched as float4 and
// requested one chunk ahead (registers) while the current chunk is multiplied out of shared memory.

// one 64 x 32 operand tile = 512 float4, two per thread.  ROWS_ARE_K: storage rows run along k (64 contiguous
// m or n per row), else storage rows run along m / n (32 contiguous k per row).
template <bool ROWS_ARE_K>
struct TileLoader {
  const float* base;
  int ld, lim_row, lim_col;      // extents along the storage row index and the contiguous index
  bool vec;
  __device__ __forceinline__ void load(int row0, int col0, int tid, float4 (&r)[2]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = row0 + (ROWS_ARE_K ? (tid >> 4) + 16 * i : (tid >> 3) + 32 * i);
      const int col = col0 + (ROWS_ARE_K ? (tid & 15) * 4 : (tid & 7) * 4);
      const float* p = base + (int64_t)row * ld + col;
      if (row < lim_row && col + 3 < lim_col && vec) {
        r[i] = *reinterpret_cast<const float4*>(p);
      } else {
        r[i].x = (row < lim_row && col < lim_col) ? p[0] : 0.f;
        r[i].y = (row < lim_row && col + 1 < lim_col) ? p[1] : 0.f;
        r[i].z = (row < lim_row && col + 2 < lim_col) ? p[2] : 0.f;
        r[i].w = (row < lim_row && col + 3 < lim_col) ? p[3] : 0.f;
      }
    }
  }
  // shared tile is [k][64 + pad]
  __device__ __forceinline__ void store(float (*S)[GM + GPAD], int tid, const float4 (&r)[2]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (ROWS_ARE_K) {
        *reinterpret_cast<float4*>(&S[(tid >> 4) + 16 * i][(tid & 15) * 4]) = r[i];
      } else {
        const int mn = (tid >> 3) + 32 * i, k = (tid & 7) * 4;
        S[k][mn] = r[i].x;
        S[k + 1][mn] = r[i].y;
        S[k + 2][mn] = r[i].z;
        S[k + 3][mn] = r[i].w;
      }
    }
  }
};

template <bool AT, bool BT>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
             float* __restrict__ C, int ldc, const float* __restrict__ bias, float* __restrict__ pre,
             int M, int N, int K, int act, int beta, int splits) {
  pdl_prologue();
  static_assert(GM == GN, "one tile loader serves both operands");
  __shared__ __align__(16) float As[GK][GM + GPAD];
  __shared__ __align__(16) float Bs[GK][GN + GPAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  const int kchunks = (K + GK - 1) / GK;
  const int per = (kchunks + splits - 1) / splits;
  const int kb0 = blockIdx.z * per, kb1 = min(kb0 + per, kchunks);
  // A: AT -> rows are k (extent K), contiguous m (extent M); else rows m, contiguous k
  const TileLoader<AT> la{A, lda, AT ? K : M, AT ? M : K, (lda & 3) == 0 && ((uintptr_t)A & 15) == 0};
  // B: BT -> storage [N][K]: rows n, contiguous k; else rows k, contiguous n
  const TileLoader<!BT> lb{Bm, ldb, BT ? N : K, BT ? K : N, (ldb & 3) == 0 && ((uintptr_t)Bm & 15) == 0};
  float4 ra[2], rb[2];
  auto fetch = [&](int kb) {
    const int k0 = kb * GK;
    if (AT) la.load(k0, m0, tid, ra); else la.load(m0, k0, tid, ra);
    if (BT) lb.load(n0, k0, tid, rb); else lb.load(k0, n0, tid, rb);
  };
  if (kb0 < kb1) fetch(kb0);
  for (int kb = kb0; kb < kb1; ++kb) {
    __syncthreads();                       // the previous chunk has been multiplied out
    la.store(As, tid, ra);
    lb.store(Bs, tid, rb);
    __syncthreads();
    if (kb + 1 < kb1) fetch(kb + 1);       // in flight during the multiply
#pragma unroll
    for (int k = 0; k < GK; ++k) {
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
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* cp = C + (int64_t)m * ldc + n;
      if (splits > 1) {
        atomicAdd(cp, acc[i][j]);
      } else {
        float v = acc[i][j] + (bias ? bias[n] : 0.f);
        if (pre) pre[(int64_t)m * ldc + n] = v;
        v = act_fwd(act, v);
        *cp = beta ? *cp + v : v;
      }
    }
  }
}

__global__ void act_bwd_kernel(float* __restrict__ dy, const float* __restrict__ pre, int64_t n,
                               int act) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dy[i] *= act_grad(act, pre[i]);
}

}  // namespace b200dm

using namespace b200dm;

extern "C" int b200dm_sinusoidal(const int64_t* t, float* emb, int32_t B, int32_t dim, float theta,
                                 void* stream) {
  B200DM_REQUIRE(B > 0 && dim >= 4 && dim % 2 == 0, B200DM_ERR_SHAPE, "sinusoidal: bad dim %d", dim);
  int half = dim / 2;
  // math.log(theta) / (half_dim - 1) is a Python double; torch multiplies arange(int64) by the
  // negated double, producing fp32: emulate by rounding the scalar to fp32 first.
  float neg = (float)(-(log((double)theta) / (double)(half - 1)));
  int n = B * half;
  launch_k(sinusoidal_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, t, emb, B, dim, neg);
  count_launch();
  return check_launch("sinusoidal");
}

extern "C" int b200dm_linear_fwd(const float* X, const float* W, const float* b, float* Y, float* pre,
                                 int32_t M, int32_t N, int32_t K, int32_t act, void* stream) {
  B200DM_REQUIRE(M > 0 && N > 0 && K > 0 && act >= 0 && act <= 2, B200DM_ERR_SHAPE, "linear_fwd: bad shape");
  dim3 grid((M + GM - 1) / GM, (N + GN - 1) / GN, 1);
  launch_k(sgemm_kernel<false, true>, grid, 256, 0, (cudaStream_t)stream, X, K, W, K, Y, N, b, pre, M, N, K,
                                                                     act, 0, 1);
  count_launch();
  return check_launch("linear_fwd");
}

// Backward of a COLUMN RANGE of a linear layer without activation: dY is a [M, N] window of a wider [M, ldy] matrix,
// W / dW the matching N rows.  dX is overwritten or (dx_accumulate) added to.  Lets the FiLM projection of the Unet
// finish the gradients of most of its 8064 rows long before the last two ResnetBlocks of the backward pass are done.
extern "C" int b200dm_linear_bwd_cols(const float* X, const float* W, const float* dY, int32_t ldy, float* dX,
                                      int32_t dx_accumulate, float* dW, int32_t M, int32_t N, int32_t K,
                                      void* stream) {
  B200DM_REQUIRE(M > 0 && N > 0 && K > 0 && ldy >= N, B200DM_ERR_SHAPE, "linear_bwd_cols: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  if (dX) {
    int tiles = ((M + GM - 1) / GM) * ((K + GN - 1) / GN);
    int chunks = (N + GK - 1) / GK;
    int splits = 1;
    if (tiles < num_sms() && chunks >= 8) {
      splits = (2 * num_sms() + tiles - 1) / tiles;
      if (splits > chunks / 4) splits = chunks / 4;
      if (splits < 1) splits = 1;
    }
    if (splits > 1 && !dx_accumulate) {
      int rc = b200dm_fill_f32(dX, (int64_t)M * K, 0.f, stream);
      if (rc) return rc;
    }
    dim3 grid((M + GM - 1) / GM, (K + GN - 1) / GN, splits);
    launch_k(sgemm_kernel<false, false>, grid, 256, 0, st, dY, ldy, W, K, dX, K, nullptr, nullptr, M, K, N, 0,
             dx_accumulate ? 1 : 0, splits);
    ++launches;
  }
  if (dW) {
    dim3 grid((N + GM - 1) / GM, (K + GN - 1) / GN, 1);
    launch_k(sgemm_kernel<true, false>, grid, 256, 0, st, dY, ldy, X, K, dW, K, nullptr, nullptr, N, K, M, 0, 1, 1);
    ++launches;
  }
  count_launch(launches);
  return check_launch("linear_bwd_cols");
}

extern "C" int b200dm_linear_bwd(const float* X, const float* W, const float* pre, float* dY,
                                 float* dX, float* dW, float* db, int32_t M, int32_t N, int32_t K,
                                 int32_t act, void* stream) {
  B200DM_REQUIRE(M > 0 && N > 0 && K > 0 && act >= 0 && act <= 2, B200DM_ERR_SHAPE, "linear_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  if (act != 0) {
    B200DM_REQUIRE(pre != nullptr, B200DM_ERR_SHAPE, "linear_bwd: pre-activation required for act=%d", act);
    int64_t n = (int64_t)M * N;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
    launch_k(act_bwd_kernel, (unsigned)blocks, 256, 0, st, dY, pre, n, act);
    ++launches;
  }
  if (dX) {  // dX[M,K] = dY[M,N] * W[N,K]; split the N reduction when the output grid is tiny
    int tiles = ((M + GM - 1) / GM) * ((K + GN - 1) / GN);
    int chunks = (N + GK - 1) / GK;
    int splits = 1;
    if (tiles < num_sms() && chunks >= 8) {
      splits = (2 * num_sms() + tiles - 1) / tiles;          // about two CTAs per SM
      if (splits > chunks / 4) splits = chunks / 4;
      if (splits < 1) splits = 1;
    }
    if (splits > 1) {
      int rc = b200dm_fill_f32(dX, (int64_t)M * K, 0.f, stream);
      if (rc) return rc;
    }
    dim3 grid((M + GM - 1) / GM, (K + GN - 1) / GN, splits);
    launch_k(sgemm_kernel<false, false>, grid, 256, 0, st, dY, N, W, K, dX, K, nullptr, nullptr, M, K, N, 0, 0, splits);
    ++launches;
  }
  if (dW) {  // dW[N,K] += dY^T[N,M] * X[M,K]
    dim3 grid((N + GM - 1) / GM, (K + GN - 1) / GN, 1);
    launch_k(sgemm_kernel<true, false>, grid, 256, 0, st, dY, N, X, K, dW, K, nullptr, nullptr, N, K, M, 0, 1, 1);
    ++launches;
  }
  count_launch(launches);
  int rc = check_launch("linear_bwd");
  if (rc) return rc;
  if (db) return b200dm_colsum(B200DM_F32, dY, N, M, N, db, 1, stream);
  return B200DM_OK;
}
