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

constexpr int GM = 64, GN = 64, GK = 16, GPAD = 4;

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
template <bool AT, bool BT>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
             float* __restrict__ C, int ldc, const float* __restrict__ bias, float* __restrict__ pre,
             int M, int N, int K, int act, int beta, int splits) {
  pdl_prologue();
  __shared__ float As[GK][GM + GPAD];
  __shared__ float Bs[GK][GN + GPAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  int kchunks = (K + GK - 1) / GK;
  int per = (kchunks + splits - 1) / splits;
  int kb0 = blockIdx.z * per, kb1 = min(kb0 + per, kchunks);
  for (int kb = kb0; kb < kb1; ++kb) {
    const int k0 = kb * GK;
    float a[4], b[4];
    if (AT) {  // rows of storage are k; 64 consecutive m
      int kk = tid >> 4, mm = (tid & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        a[j] = (k0 + kk < K && m0 + mm + j < M) ? A[(int64_t)(k0 + kk) * lda + m0 + mm + j] : 0.f;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) As[kk][mm + j] = a[j];
    } else {   // rows of storage are m; 16 consecutive k
      int mm = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        a[j] = (m0 + mm < M && k0 + kk + j < K) ? A[(int64_t)(m0 + mm) * lda + k0 + kk + j] : 0.f;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) As[kk + j][mm] = a[j];
    }
    if (BT) {  // storage [N][K]
      int nn = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b[j] = (n0 + nn < N && k0 + kk + j < K) ? Bm[(int64_t)(n0 + nn) * ldb + k0 + kk + j] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[kk + j][nn] = b[j];
    } else {   // storage [K][N]
      int kk = tid >> 4, nn = (tid & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b[j] = (k0 + kk < K && n0 + nn + j < N) ? Bm[(int64_t)(k0 + kk) * ldb + n0 + nn + j] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[kk][nn + j] = b[j];
    }
    __syncthreads();
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
    __syncthreads();
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
    if (tiles < num_sms() && chunks >= 16) {
      splits = (num_sms() + tiles - 1) / tiles;
      if (splits > chunks / 8) splits = chunks / 8;
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
