// Discrete-graph-learning trunk, dense part:  feat = BN3(relu(y2n W^T + b))   and its backward.
//   y2n [N, K] (K = 16 * (train_len - 18): 217k ... 583k), W = fc.weight [100, K] (the largest parameter of STEP:
//   22 - 58 M values), N = nodes.  Reference: step/step_arch/discrete_graph_learning.py:66,134-135
//   (`self.fc`, `self.bn3` on the flattened conv trunk), training-mode BatchNorm1d over the N nodes.
//
// All three GEMMs stream one huge fp32 matrix exactly once and are HBM-bound on tensor cores:
//   forward      z  [N,100] = X  [N,K]  W^T          reads X (318 MB @ METR-LA) + W (153 MB)
//   backward dX     [N,K]   = g  [N,100] W           reads W, writes dX
//   backward dW     [100,K] = g^T        X           reads X, writes dW
// They run on tcgen05 with fp32-class accuracy: every fp32 operand is split in the CTA into bf16 hi + lo
// (x = hi + lo, |lo| <= 2^-9 |x|) and hi*hi + hi*lo + lo*hi is accumulated in fp32 in TMEM (relative error
// ~2^-17 per product), the scheme of the GWNet node mixes (tc_mix.cuh).  The 3x MMA work (48 GFLOP) stays far
// under the HBM time, so one code path serves the fp32 parity mode and the bf16 performance mode.
//
// Operands are converted by the CTA itself: worker warps read fp32 rows from global memory with full-line
// coalescing (a warp instruction covers 8 rows x 128 B), split, and store UMMA no-swizzle canonical images into
// shared memory (8 lanes of a store phase write one contiguous 128-byte core matrix: conflict-free).  A single
// thread issues the MMAs; stages are double-buffered through mbarriers.  The K axis is partitioned across the grid
// (and, when the trunk is sharded over ranks, across GPUs: the caller passes its [k_begin, k_end) range).
#include "common.cuh"
#include "tc_common.cuh"

namespace stepk {
using namespace tc;

constexpr int FC_F = 100;            // fc output features
constexpr int FC_FP = 112;           // padded to the MMA N / K granularity (multiple of 16)
constexpr int FC_WORKERS = 16;       // worker warps (operand conversion + epilogue)
constexpr int FC_THREADS = (2 + FC_WORKERS) * 32;   // warp 0 idle (keeps the library's role convention), warp 1 MMA

__device__ __forceinline__ void split8(const float4 &x0, const float4 &x1, uint4 &hi, uint4 &lo) {
  const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  float h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
    l[j] = x[j] - h[j];
  }
  hi = pack8_bf16(h);
  lo = pack8_bf16(l);
}

__device__ __forceinline__ void split8(const f8 &x, uint4 &hi, uint4 &lo) {
  float h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __bfloat162float(__float2bfloat16_rn(x.v[j]));
    l[j] = x.v[j] - h[j];
  }
  hi = pack8_bf16(h);
  lo = pack8_bf16(l);
}
__device__ __forceinline__ f8 f8_zero() {
  f8 r;
#pragma unroll
  for (int j = 0; j < 8; ++j) r.v[j] = 0.f;
  return r;
}

struct FcRange { long long k_begin, k_end; };     // element range of the K axis this launch covers

// stage range of CTA `i` of `n` over `total` stages
__device__ __forceinline__ void cta_stages(long long total, int i, int n, long long &s0, long long &s1) {
  s0 = total * i / n;
  s1 = total * (i + 1) / n;
}

// ===========================================================================
// K1 forward: zp[split][n][0..100) = sum over the split's K range of X[n][k] W[j][k]
//   A = X stage  (K-major [4 chunks][ROWS][8]),  B = W stage (K-major [4 chunks][112][8]),  32 K elements per stage
//   grid = (splits, row groups of MT*128 rows)
// ===========================================================================
constexpr int FCF_KC = 32;

template <int MT>
__global__ void __launch_bounds__(FC_THREADS, 1) fc_fwd_kernel(const float *__restrict__ X, const float *__restrict__ W,
                                                               long long ldk, FcRange rg, int Nn, float *__restrict__ zp) {
  constexpr int ROWS = MT * 128;
  constexpr uint32_t A_IMG = 4u * ROWS * 16, B_IMG = 4u * FC_FP * 16;
  constexpr uint32_t STAGE = 2 * A_IMG + 2 * B_IMG;
  extern __shared__ __align__(1024) uint8_t fc_smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(fc_smem + 2 * STAGE);
  uint64_t *built = bars, *consumed = bars + 2, *d_full = bars + 4;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.y * ROWS;
  const int rows_here = min(ROWS, Nn - row0);
  const int mt_here = (rows_here + 127) / 128;

  // zero both stages once: padding rows (>= Nn, features >= 100) are never written afterwards
  for (uint32_t i = threadIdx.x; i < 2 * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4 *>(fc_smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&built[i], FC_WORKERS); mbar_init(&consumed[i], 1); }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, MT * 128 >= 512 ? 512 : (MT * 128 <= 128 ? 128 : (MT * 128 <= 256 ? 256 : 512)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long total = (rg.k_end - rg.k_begin + FCF_KC - 1) / FCF_KC;
  long long s0, s1;
  cta_stages(total, blockIdx.x, gridDim.x, s0, s1);
  const int nst = (int)(s1 - s0);

  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, FC_FP, 0, 0);
      for (int i = 0; i < nst; ++i) {
        const int st = i & 1;
        mbar_wait(&built[st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t ah = smem_u32(fc_smem + st * STAGE), al = ah + A_IMG, bh = al + A_IMG, bl = bh + B_IMG;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t dbh = umma_desc(bh + kk * 2 * FC_FP * 16, FC_FP * 16, 128);
          const uint64_t dbl = umma_desc(bl + kk * 2 * FC_FP * 16, FC_FP * 16, 128);
          for (int m = 0; m < mt_here; ++m) {
            const uint64_t dah = umma_desc(ah + kk * 2 * ROWS * 16 + m * 2048, ROWS * 16, 128);
            const uint64_t dal = umma_desc(al + kk * 2 * ROWS * 16 + m * 2048, ROWS * 16, 128);
            const uint32_t d = tmem + m * 128;
            umma_bf16(d, dah, dbh, idesc, (i | kk) != 0 ? 1u : 0u);
            umma_bf16(d, dah, dbl, idesc, 1u);
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
        }
        umma_commit(&consumed[st]);
      }
      umma_commit(d_full);
    }
  } else if (warp >= 2) {
    const int ww = warp - 2;
    // row blocks of 8 rows: [0, ROWS/8) -> X rows, [ROWS/8, ROWS/8 + 13) -> W rows; lane = (chunk c, row-in-block)
    constexpr int NBLK = ROWS / 8 + 13;
    constexpr int PER = (NBLK + FC_WORKERS - 1) / FC_WORKERS;
    const int c = lane >> 3, rr = lane & 7;
    const float *src[PER];
    uint32_t dst[PER];          // unit index inside the stage: A units [0, 4*ROWS), B units follow
    bool isA[PER];
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int blk = ww + p * FC_WORKERS;
      src[p] = nullptr; dst[p] = 0; isA[p] = true;
      if (blk < ROWS / 8) {
        const int r = blk * 8 + rr;
        if (r < rows_here) { src[p] = X + (size_t)(row0 + r) * ldk + c * 8; dst[p] = (uint32_t)c * ROWS + r; }
      } else if (blk < NBLK) {
        const int j = (blk - ROWS / 8) * 8 + rr;
        if (j < FC_F) { src[p] = W + (size_t)j * ldk + c * 8; dst[p] = (uint32_t)c * FC_FP + j; isA[p] = false; }
      }
    }
    // two register sets: the loads of stages i+1 and i+2 are in flight while stage i is converted (HBM latency x
    // per-SM bandwidth needs ~2 stages = 90 KB outstanding)
    f8 va[PER], vb[PER];
    auto load_stage = [&](f8 (&v)[PER], long long s) {
      const long long k0 = rg.k_begin + s * FCF_KC;
      const bool kval = k0 + c * 8 < rg.k_end;           // K is a multiple of 8: a chunk is all-valid or all-padding
#pragma unroll
      for (int p = 0; p < PER; ++p) v[p] = (src[p] != nullptr && kval) ? ld256_nc(src[p] + k0) : f8_zero();
    };
    auto store_stage = [&](const f8 (&v)[PER], int i) {
      const int st = i & 1;
      mbar_wait(&consumed[st], ((i >> 1) & 1) ^ 1);
      uint8_t *base = fc_smem + st * STAGE;
#pragma unroll
      for (int p = 0; p < PER; ++p) {
        if (src[p] != nullptr) {
          uint4 hi, lo;
          split8(v[p], hi, lo);
          uint8_t *img = isA[p] ? base : base + 2 * A_IMG;
          const uint32_t half = isA[p] ? A_IMG : B_IMG;
          reinterpret_cast<uint4 *>(img)[dst[p]] = hi;
          reinterpret_cast<uint4 *>(img + half)[dst[p]] = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&built[st]);
    };
    if (MT <= 2) {
      if (nst > 0) load_stage(va, s0);
      if (nst > 1) load_stage(vb, s0 + 1);
      for (int i = 0; i < nst; i += 2) {
        store_stage(va, i);
        if (i + 2 < nst) load_stage(va, s0 + i + 2);
        if (i + 1 < nst) {
          store_stage(vb, i + 1);
          if (i + 3 < nst) load_stage(vb, s0 + i + 3);
        }
      }
    } else {      // 3-4 row tiles: one stage already keeps 64-80 KB in flight per SM; a second register set would spill
      if (nst > 0) load_stage(va, s0);
      for (int i = 0; i < nst; ++i) {
        store_stage(va, i);
        if (i + 1 < nst) load_stage(va, s0 + i + 1);
      }
    }
    // epilogue: TMEM -> zp[split][row][0..100)
    mbar_wait(d_full, 0);
    tc_fence_after();
    const int q = warp & 3, m = ww >> 2;
    if (m < mt_here) {
      const int r = m * 128 + q * 32 + lane;
      float *o = zp + ((size_t)blockIdx.x * Nn + row0 + r) * FC_F;
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + m * 128;
      float t[32];
#pragma unroll 1
      for (int c0 = 0; c0 < 96; c0 += 32) {
        if (nst > 0) tmem_ld32(ta + c0, t);
        else {
#pragma unroll
          for (int x = 0; x < 32; ++x) t[x] = 0.f;
        }
        if (r < rows_here) {
#pragma unroll
          for (int x = 0; x < 32; x += 4) *reinterpret_cast<float4 *>(o + c0 + x) = make_float4(t[x], t[x + 1], t[x + 2], t[x + 3]);
        }
      }
      float t16[16];
      if (nst > 0) tmem_ld16(ta + 96, t16);
      else {
#pragma unroll
        for (int x = 0; x < 16; ++x) t16[x] = 0.f;
      }
      if (r < rows_here) *reinterpret_cast<float4 *>(o + 96) = make_float4(t16[0], t16[1], t16[2], t16[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, MT * 128 <= 128 ? 128 : (MT * 128 <= 256 ? 256 : 512));
}

// z_raw[n][j] = sum_s zp[s][n][j]     (deterministic second stage of the split-K reduction)
__global__ void fc_reduce_kernel(const float *__restrict__ zp, int splits, long long per, float *__restrict__ z) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    a0 += zp[(size_t)s * per + e]; a1 += zp[(size_t)(s + 1) * per + e];
    a2 += zp[(size_t)(s + 2) * per + e]; a3 += zp[(size_t)(s + 3) * per + e];
  }
  for (; s < splits; ++s) a0 += zp[(size_t)s * per + e];
  z[e] = (a0 + a1) + (a2 + a3);
}

// bias + ReLU + BatchNorm1d over the N nodes (one block per feature):
//   z = z_raw + b (stored in place: backward needs the ReLU mask), r = relu(z),
//   training: mean/var of r over nodes (biased var normalises; stats[0]=mean, [1]=biased var, [2]=rstd), else the
//   given running statistics; feat = (r - mean) * rstd * gamma + beta.
__global__ void __launch_bounds__(256) fc_bn_fwd_kernel(float *__restrict__ z, const float *__restrict__ bias,
                                                        const float *__restrict__ gamma, const float *__restrict__ beta, int Nn,
                                                        float eps, int training, float *__restrict__ stats /*[3][100]*/,
                                                        float *__restrict__ feat) {
  __shared__ double red[2][8];
  __shared__ float s_mean, s_rstd;
  const int j = blockIdx.x, tid = threadIdx.x;
  const float b = bias[j];
  double s = 0.0, ss = 0.0;
  for (int n = tid; n < Nn; n += 256) {
    const float zz = z[(size_t)n * FC_F + j] + b;
    z[(size_t)n * FC_F + j] = zz;
    const float r = fmaxf(zz, 0.f);
    s += r; ss += (double)r * r;
  }
  if (training) {
    s = warp_sum_d(s); ss = warp_sum_d(ss);
    if ((tid & 31) == 0) { red[0][tid >> 5] = s; red[1][tid >> 5] = ss; }
    __syncthreads();
    if (tid == 0) {
      double a = 0, q = 0;
      for (int w = 0; w < 8; ++w) { a += red[0][w]; q += red[1][w]; }
      const double mean = a / Nn, var = fmax(q / Nn - mean * mean, 0.0);
      stats[j] = (float)mean; stats[FC_F + j] = (float)var; stats[2 * FC_F + j] = (float)(1.0 / sqrt(var + (double)eps));
      s_mean = (float)mean; s_rstd = stats[2 * FC_F + j];
    }
  } else if (tid == 0) {
    s_mean = stats[j]; s_rstd = rsqrtf(stats[FC_F + j] + eps);
    stats[2 * FC_F + j] = s_rstd;
  }
  __syncthreads();
  const float mean = s_mean, sc = s_rstd * gamma[j], be = beta[j];
  for (int n = tid; n < Nn; n += 256) feat[(size_t)n * FC_F + j] = (fmaxf(z[(size_t)n * FC_F + j], 0.f) - mean) * sc + be;
}

// backward of fc_bn_fwd_kernel (training mode): dfeat -> g = dL/dz, dgamma, dbeta, dbias
__global__ void __launch_bounds__(256) fc_bn_bwd_kernel(const float *__restrict__ dfeat, const float *__restrict__ z,
                                                        const float *__restrict__ gamma, const float *__restrict__ stats, int Nn,
                                                        float *__restrict__ g, float *__restrict__ dgamma,
                                                        float *__restrict__ dbeta, float *__restrict__ dbias) {
  __shared__ double red[2][8];
  __shared__ float s_a, s_b;
  const int j = blockIdx.x, tid = threadIdx.x;
  const float mean = stats[j], rstd = stats[2 * FC_F + j], ga = gamma[j];
  double sd = 0.0, sx = 0.0;
  for (int n = tid; n < Nn; n += 256) {
    const float dy = dfeat[(size_t)n * FC_F + j];
    const float xh = (fmaxf(z[(size_t)n * FC_F + j], 0.f) - mean) * rstd;
    sd += dy; sx += (double)dy * xh;
  }
  sd = warp_sum_d(sd); sx = warp_sum_d(sx);
  if ((tid & 31) == 0) { red[0][tid >> 5] = sd; red[1][tid >> 5] = sx; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, q = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; q += red[1][w]; }
    dbeta[j] = (float)a; dgamma[j] = (float)q;
    s_a = (float)(a / Nn); s_b = (float)(q / Nn);
  }
  __syncthreads();
  const float ma = s_a, mb = s_b;
  double sb = 0.0;
  for (int n = tid; n < Nn; n += 256) {
    const float zz = z[(size_t)n * FC_F + j];
    const float xh = (fmaxf(zz, 0.f) - mean) * rstd;
    const float dr = ga * rstd * (dfeat[(size_t)n * FC_F + j] - ma - xh * mb);
    const float dz = zz > 0.f ? dr : 0.f;
    g[(size_t)n * FC_F + j] = dz;
    sb += dz;
  }
  sb = warp_sum_d(sb);
  __syncthreads();
  if ((tid & 31) == 0) red[0][tid >> 5] = sb;
  __syncthreads();
  if (tid == 0) {
    double a = 0;
    for (int w = 0; w < 8; ++w) a += red[0][w];
    dbias[j] = (float)a;
  }
}

// resident g images shared by the two backward kernels: unit (row r, chunk c) = g[r][8c .. 8c+8) (zeros past column 100)
__device__ __forceinline__ void load_g_unit(const float *__restrict__ g, int r, int c, bool row_ok, float4 &x0, float4 &x1) {
  x0 = make_float4(0.f, 0.f, 0.f, 0.f); x1 = x0;
  if (row_ok && c < 13) {
    const float4 *p = reinterpret_cast<const float4 *>(g + (size_t)r * FC_F + c * 8);
    x0 = __ldg(p);
    if (c < 12) x1 = __ldg(p + 1);
  }
}

// ===========================================================================
// K2 backward to the trunk activations: dX[n][k] = sum_j g[n][j] W[j][k]
//   A = g images, resident (K-major [14 chunks][256 rows][8]); B = W stage, MN-major [8 column groups][112][8];
//   64 K columns per stage, double-buffered accumulators [2 row tiles x 64 columns].
//   grid = (splits, row groups of 256 rows).  Workers: warps 2-9 convert W, warps 10-17 drain TMEM.
// ===========================================================================
constexpr int FCX_KC = 64, FCX_ROWS = 256;
constexpr uint32_t FCX_G_IMG = 14u * FCX_ROWS * 16;                 // one of (hi, lo)
constexpr uint32_t FCX_W_IMG = (FCX_KC / 8) * FC_FP * 16;           // one of (hi, lo)

__global__ void __launch_bounds__(FC_THREADS, 1) fc_dx_kernel(const float *__restrict__ g, const float *__restrict__ W,
                                                              long long ldk, FcRange rg, int Nn, float *__restrict__ dX) {
  extern __shared__ __align__(1024) uint8_t fc_smem[];
  uint8_t *sG = fc_smem;                         // hi | lo
  uint8_t *sW = sG + 2 * FCX_G_IMG;              // 2 stages x (hi | lo)
  float *sStage = reinterpret_cast<float *>(sW + 4 * FCX_W_IMG);        // 8 epilogue warps x [32][33] transpose tiles
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStage + 8 * 32 * 33);
  uint64_t *built = bars, *consumed = bars + 2, *acc_full = bars + 4, *acc_empty = bars + 6, *g_ready = bars + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.y * FCX_ROWS;
  const int rows_here = min(FCX_ROWS, Nn - row0);
  const int mt_here = (rows_here + 127) / 128;

  for (uint32_t i = threadIdx.x; i < 4 * FCX_W_IMG / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sW)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();       // the padding features (rows 100..111) stay zero and are read by the tensor core
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&built[i], 8); mbar_init(&consumed[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8);
    }
    mbar_init(g_ready, FC_WORKERS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long total = (rg.k_end - rg.k_begin + FCX_KC - 1) / FCX_KC;
  long long s0, s1;
  cta_stages(total, blockIdx.x, gridDim.x, s0, s1);
  const int nst = (int)(s1 - s0);

  if (warp >= 2) {
    // ---- resident g images: 14 chunks x 256 rows (all 16 worker warps) ----
    const int ww = warp - 2;
    for (int blk = ww; blk < 4 * (FCX_ROWS / 8); blk += FC_WORKERS) {      // 4 chunk quads (14 chunks) x 32 row blocks
      // a warp instruction covers 8 rows x 4 consecutive chunks (128 B of a g row)
      const int rb = blk % (FCX_ROWS / 8), cb = blk / (FCX_ROWS / 8);
      const int r = rb * 8 + (lane & 7), c = cb * 4 + (lane >> 3);
      if (c < 14) {
        float4 x0, x1;
        load_g_unit(g, row0 + r, c, r < rows_here, x0, x1);
        uint4 hi, lo;
        split8(x0, x1, hi, lo);
        reinterpret_cast<uint4 *>(sG)[c * FCX_ROWS + r] = hi;
        reinterpret_cast<uint4 *>(sG + FCX_G_IMG)[c * FCX_ROWS + r] = lo;
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(g_ready);
  }

  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, FCX_KC, 0, 1);
      mbar_wait(g_ready, 0);
      tc_fence_after();
      const uint32_t gh = smem_u32(sG), gl = gh + FCX_G_IMG;
      for (int i = 0; i < nst; ++i) {
        const int st = i & 1;
        mbar_wait(&built[st], (i >> 1) & 1);
        mbar_wait(&acc_empty[st], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t wh = smem_u32(sW + st * 2 * FCX_W_IMG), wl = wh + FCX_W_IMG;
#pragma unroll 1
        for (int ks = 0; ks < FC_FP / 16; ++ks) {
          const uint64_t dbh = umma_desc(wh + ks * 256, 128, FC_FP * 16), dbl = umma_desc(wl + ks * 256, 128, FC_FP * 16);
          for (int m = 0; m < mt_here; ++m) {
            const uint64_t dah = umma_desc(gh + ks * 2 * FCX_ROWS * 16 + m * 2048, FCX_ROWS * 16, 128);
            const uint64_t dal = umma_desc(gl + ks * 2 * FCX_ROWS * 16 + m * 2048, FCX_ROWS * 16, 128);
            const uint32_t d = tmem + st * 128 + m * FCX_KC;
            umma_bf16(d, dah, dbh, idesc, ks != 0 ? 1u : 0u);
            umma_bf16(d, dah, dbl, idesc, 1u);
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
        }
        umma_commit(&consumed[st]);
        umma_commit(&acc_full[st]);
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ---- W stage converters: unit (column group n8, feature j) = W[j][k0 + 8 n8 .. +8) -> image index n8*112 + j ----
    const int ww = warp - 2;
    // a warp instruction: 8 features x 4 column groups (128 contiguous bytes per feature row); 13 feature blocks x 2 halves
    constexpr int NIT = 26, PER = (NIT + 7) / 8;
    f8 va[PER], vb[PER];
    auto load_stage = [&](f8 (&v)[PER], long long s) {
      const long long k0 = rg.k_begin + s * FCX_KC;
#pragma unroll
      for (int p = 0; p < PER; ++p) {
        const int it = ww + p * 8;
        const int j = (it >> 1) * 8 + (lane & 7), n8 = (it & 1) * 4 + (lane >> 3);
        v[p] = (it < NIT && j < FC_F && k0 + n8 * 8 < rg.k_end) ? ld256_nc(W + (size_t)j * ldk + k0 + n8 * 8) : f8_zero();
      }
    };
    auto store_stage = [&](const f8 (&v)[PER], int i) {
      const int st = i & 1;
      mbar_wait(&consumed[st], ((i >> 1) & 1) ^ 1);
      uint8_t *base = sW + st * 2 * FCX_W_IMG;
#pragma unroll
      for (int p = 0; p < PER; ++p) {
        const int it = ww + p * 8;
        const int j = (it >> 1) * 8 + (lane & 7), n8 = (it & 1) * 4 + (lane >> 3);
        if (it < NIT && j < FC_F) {
          uint4 hi, lo;
          split8(v[p], hi, lo);
          reinterpret_cast<uint4 *>(base)[n8 * FC_FP + j] = hi;
          reinterpret_cast<uint4 *>(base + FCX_W_IMG)[n8 * FC_FP + j] = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&built[st]);
    };
    if (nst > 0) load_stage(va, s0);
    if (nst > 1) load_stage(vb, s0 + 1);
    for (int i = 0; i < nst; i += 2) {
      store_stage(va, i);
      if (i + 2 < nst) load_stage(va, s0 + i + 2);
      if (i + 1 < nst) {
        store_stage(vb, i + 1);
        if (i + 3 < nst) load_stage(vb, s0 + i + 3);
      }
    }
  } else if (warp >= 10) {
    // ---- epilogue: accumulator [row tile m][64 columns] -> dX rows.  TMEM gives a thread one row; the 32 x 32 block is
    // transposed through a private shared-memory tile so that every store instruction writes one full 128-byte line ----
    const int q = warp & 3, m = (warp - 10) >> 2;
    const int rbase = m * 128 + q * 32;
    float *stg = sStage + (warp - 10) * (32 * 33);
    for (int i = 0; i < nst; ++i) {
      const int st = i & 1;
      mbar_wait(&acc_full[st], (i >> 1) & 1);
      tc_fence_after();
      const long long k0 = rg.k_begin + (s0 + i) * FCX_KC;
      if (m < mt_here) {
#pragma unroll 1
        for (int c0 = 0; c0 < FCX_KC; c0 += 32) {
          float t[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + st * 128 + m * FCX_KC + c0, t);
#pragma unroll
          for (int x = 0; x < 32; ++x) stg[lane * 33 + x] = t[x];
          __syncwarp();
          const bool col_ok = k0 + c0 + lane < rg.k_end;
          const int nrows = min(32, rows_here - rbase);
          float *o = dX + (size_t)(row0 + rbase) * ldk + k0 + c0 + lane;
          for (int rr = 0; rr < nrows; ++rr)
            if (col_ok) o[(size_t)rr * ldk] = stg[rr * 33 + lane];
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[st]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ===========================================================================
// K3 backward to the weight: dW[j][k] = sum_n g[n][j] X[n][k]
//   A = g^T images, resident per pass of 256 nodes (MN-major [16 feature groups][256 nodes][8]);
//   B = X stage, MN-major [4 column groups][256 nodes][8]; 32 K columns per stage; accumulator [128 x 32] double-buffered.
//   More than 256 nodes: further passes over the CTA's K range accumulate into dW (the CTA owns its slice).
//   grid = (splits).  Workers: warps 2-13 convert X, warps 14-17 drain TMEM.
// ===========================================================================
constexpr int FCW_KC = 32, FCW_NODES = 256;
constexpr uint32_t FCW_G_IMG = 16u * FCW_NODES * 16;                // one of (hi, lo)
constexpr uint32_t FCW_X_IMG = (FCW_KC / 8) * FCW_NODES * 16;       // one of (hi, lo)
constexpr int FCW_CONV = 12;                                         // converter warps

__global__ void __launch_bounds__(FC_THREADS, 1) fc_dw_kernel(const float *__restrict__ g, const float *__restrict__ X,
                                                              long long ldk, FcRange rg, int Nn, float scale,
                                                              float *__restrict__ dW) {
  extern __shared__ __align__(1024) uint8_t fc_smem[];
  uint8_t *sG = fc_smem;
  uint8_t *sX = sG + 2 * FCW_G_IMG;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sX + 4 * FCW_X_IMG);
  uint64_t *built = bars, *consumed = bars + 2, *acc_full = bars + 4, *acc_empty = bars + 6, *g_ready = bars + 8, *g_free = bars + 9;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&built[i], FCW_CONV); mbar_init(&consumed[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4);
    }
    mbar_init(g_ready, FC_WORKERS); mbar_init(g_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long total = (rg.k_end - rg.k_begin + FCW_KC - 1) / FCW_KC;
  long long s0, s1;
  cta_stages(total, blockIdx.x, gridDim.x, s0, s1);
  const int nst = (int)(s1 - s0);
  const int npass = (Nn + FCW_NODES - 1) / FCW_NODES;

  // one flat sequence of (pass, stage) iterations keeps every ring's phase arithmetic uniform
  for (int pass = 0; pass < npass; ++pass) {
    const int n0 = pass * FCW_NODES;
    const int nodes_here = min(FCW_NODES, Nn - n0);
    const int ksteps = (nodes_here + 15) / 16;
    const int it0 = pass * nst;

    if (warp >= 2) {
      // ---- g^T images of this pass: unit (feature group j8, node) = g[node][8 j8 .. +8) -> index j8*256 + node ----
      if (pass > 0) mbar_wait(g_free, (pass - 1) & 1);          // the previous pass's MMAs have read the old images
      const int ww = warp - 2;
      for (int blk = ww; blk < 16 * (FCW_NODES / 8) / 4; blk += FC_WORKERS) {
        const int rb = blk % (FCW_NODES / 8), cb = blk / (FCW_NODES / 8);
        const int r = rb * 8 + (lane & 7), c = cb * 4 + (lane >> 3);
        float4 x0, x1;
        load_g_unit(g, n0 + r, c, r < nodes_here, x0, x1);
        uint4 hi, lo;
        split8(x0, x1, hi, lo);
        reinterpret_cast<uint4 *>(sG)[c * FCW_NODES + r] = hi;
        reinterpret_cast<uint4 *>(sG + FCW_G_IMG)[c * FCW_NODES + r] = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_ready);
    }

    if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, FCW_KC, 1, 1);
        mbar_wait(g_ready, pass & 1);
        tc_fence_after();
        const uint32_t gh = smem_u32(sG), gl = gh + FCW_G_IMG;
        for (int i = 0; i < nst; ++i) {
          const int it = it0 + i, st = it & 1;
          mbar_wait(&built[st], (it >> 1) & 1);
          mbar_wait(&acc_empty[st], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t xh = smem_u32(sX + st * 2 * FCW_X_IMG), xl = xh + FCW_X_IMG;
          const uint32_t d = tmem + st * FCW_KC;
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t dah = umma_desc(gh + ks * 256, 128, FCW_NODES * 16), dal = umma_desc(gl + ks * 256, 128, FCW_NODES * 16);
            const uint64_t dbh = umma_desc(xh + ks * 256, 128, FCW_NODES * 16), dbl = umma_desc(xl + ks * 256, 128, FCW_NODES * 16);
            umma_bf16(d, dah, dbh, idesc, ks != 0 ? 1u : 0u);
            umma_bf16(d, dah, dbl, idesc, 1u);
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
          umma_commit(&consumed[st]);
          umma_commit(&acc_full[st]);
        }
        umma_commit(g_free);
      }
    } else if (warp >= 2 && warp < 2 + FCW_CONV) {
      // ---- X stage converters: unit (column group c, node r) = X[n0 + r][k0 + 8c .. +8) -> index c*256 + r ----
      const int ww = warp - 2;
      constexpr int NBLK = FCW_NODES / 8, PER = (NBLK + FCW_CONV - 1) / FCW_CONV;
      const int c = lane >> 3, rr = lane & 7;
      f8 va[PER], vb[PER];
      auto load_stage = [&](f8 (&v)[PER], long long s) {
        const long long k0 = rg.k_begin + s * FCW_KC + c * 8;
#pragma unroll
        for (int p = 0; p < PER; ++p) {
          const int blk = ww + p * FCW_CONV, r = blk * 8 + rr;
          v[p] = (blk < NBLK && r < nodes_here && k0 < rg.k_end) ? ld256_nc(X + (size_t)(n0 + r) * ldk + k0) : f8_zero();
        }
      };
      auto store_stage = [&](const f8 (&v)[PER], int i) {
        const int it = it0 + i, st = it & 1;
        mbar_wait(&consumed[st], ((it >> 1) & 1) ^ 1);
        uint8_t *base = sX + st * 2 * FCW_X_IMG;
#pragma unroll
        for (int p = 0; p < PER; ++p) {
          const int blk = ww + p * FCW_CONV, r = blk * 8 + rr;
          if (blk < NBLK) {                                   // rows past the pass's nodes are written as zeros
            uint4 hi, lo;
            split8(v[p], hi, lo);
            reinterpret_cast<uint4 *>(base)[c * FCW_NODES + r] = hi;
            reinterpret_cast<uint4 *>(base + FCW_X_IMG)[c * FCW_NODES + r] = lo;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&built[st]);
      };
      if (nst > 0) load_stage(va, s0);
      if (nst > 1) load_stage(vb, s0 + 1);
      for (int i = 0; i < nst; i += 2) {
        store_stage(va, i);
        if (i + 2 < nst) load_stage(va, s0 + i + 2);
        if (i + 1 < nst) {
          store_stage(vb, i + 1);
          if (i + 3 < nst) load_stage(vb, s0 + i + 3);
        }
      }
    } else if (warp >= 2 + FCW_CONV) {
      // ---- epilogue: lane = feature j; dW[j][k0 .. k0+32) (+)= scale * acc ----
      const int q = warp & 3, j = q * 32 + lane;
      for (int i = 0; i < nst; ++i) {
        const int it = it0 + i, st = it & 1;
        mbar_wait(&acc_full[st], (it >> 1) & 1);
        tc_fence_after();
        float t[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + st * FCW_KC, t);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[st]);
        const long long k0 = rg.k_begin + (s0 + i) * FCW_KC;
        if (j < FC_F) {
          float *o = dW + (size_t)j * ldk + k0;
#pragma unroll
          for (int x = 0; x < 32; x += 4) {
            if (k0 + x < rg.k_end) {
              float4 v = make_float4(t[x] * scale, t[x + 1] * scale, t[x + 2] * scale, t[x + 3] * scale);
              if (pass > 0) {
                const float4 old = *reinterpret_cast<const float4 *>(o + x);
                v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
              }
              *reinterpret_cast<float4 *>(o + x) = v;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}

static int fc_grid_splits(long long stages) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(stages < sms ? (stages > 0 ? stages : 1) : sms);
}

}  // namespace stepk

using namespace stepk;

extern "C" int step_dgl_fc_splits(int N, long long k_begin, long long k_end) {
  if (N <= 0 || k_end <= k_begin) return 0;
  const int groups = (N + 511) / 512;
  int s = fc_grid_splits((k_end - k_begin + FCF_KC - 1) / FCF_KC) / groups;
  return s < 1 ? 1 : s;
}

extern "C" int step_dgl_fc_fwd(const float *x, const float *w, int N, long long K, long long k_begin, long long k_end,
                               float *partial, float *z_raw, void *stream) {
  STEP_REQUIRE(x && w && partial && z_raw && N > 0 && K > 0, "dgl_fc_fwd: bad argument");
  STEP_REQUIRE(K % 8 == 0 && k_begin >= 0 && k_begin % FCF_KC == 0 && k_end > k_begin && k_end <= K && k_end % 8 == 0,
               "dgl_fc_fwd: K must be a multiple of 8 and the range must start on a multiple of 32");
  STEP_REQUIRE(((uintptr_t)x & 31) == 0 && ((uintptr_t)w & 31) == 0, "dgl_fc_fwd: operands must be 32-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = step_dgl_fc_splits(N, k_begin, k_end);
  const FcRange rg{k_begin, k_end};
  int rc;
  const int mt_all = (N + 127) / 128;
#define FCF_LAUNCH(MT)                                                                              \
  do {                                                                                              \
    const size_t smem = 2 * (size_t)(2 * 4 * (MT * 128) * 16 + 2 * 4 * FC_FP * 16) + 8 * 8 + 16;    \
    if ((rc = allow_smem(fc_fwd_kernel<MT>, smem))) return rc;                                      \
    fc_fwd_kernel<MT><<<dim3(splits, (N + MT * 128 - 1) / (MT * 128)), FC_THREADS, smem, st>>>(x, w, K, rg, N, partial); \
  } while (0)
  if (mt_all == 1) FCF_LAUNCH(1);
  else if (mt_all == 2) FCF_LAUNCH(2);
  else if (mt_all == 3) FCF_LAUNCH(3);
  else FCF_LAUNCH(4);
#undef FCF_LAUNCH
  STEP_LAUNCH_CHECK("fc_fwd_kernel");
  const long long per = (long long)N * FC_F;
  fc_reduce_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(partial, splits, per, z_raw);
  return check_launch("fc_reduce_kernel");
}

extern "C" int step_dgl_fc_bn_fwd(float *z, const float *bias, const float *gamma, const float *beta, int N, float eps,
                                  int training, float *stats, float *feat, void *stream) {
  STEP_REQUIRE(z && bias && gamma && beta && stats && feat && N > 0, "dgl_fc_bn_fwd: bad argument");
  fc_bn_fwd_kernel<<<FC_F, 256, 0, (cudaStream_t)stream>>>(z, bias, gamma, beta, N, eps, training, stats, feat);
  return check_launch("fc_bn_fwd_kernel");
}

extern "C" int step_dgl_fc_bn_bwd(const float *dfeat, const float *z, const float *gamma, const float *stats, int N, float *g,
                                  float *dgamma, float *dbeta, float *dbias, void *stream) {
  STEP_REQUIRE(dfeat && z && gamma && stats && g && dgamma && dbeta && dbias && N > 0, "dgl_fc_bn_bwd: bad argument");
  fc_bn_bwd_kernel<<<FC_F, 256, 0, (cudaStream_t)stream>>>(dfeat, z, gamma, stats, N, g, dgamma, dbeta, dbias);
  return check_launch("fc_bn_bwd_kernel");
}

extern "C" int step_dgl_fc_bwd(const float *g, const float *x, const float *w, int N, long long K, long long k_begin,
                               long long k_end, float dw_scale, float *dx, float *dw, void *stream) {
  STEP_REQUIRE(g && x && w && dx && dw && N > 0 && K > 0, "dgl_fc_bwd: bad argument");
  STEP_REQUIRE(K % 16 == 0 && k_begin >= 0 && k_begin % FCX_KC == 0 && k_end > k_begin && k_end <= K && k_end % 8 == 0,
               "dgl_fc_bwd: K must be a multiple of 16 and the range must start on a multiple of 64");
  STEP_REQUIRE(((uintptr_t)x & 31) == 0 && ((uintptr_t)w & 31) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)dx & 15) == 0 &&
                   ((uintptr_t)dw & 15) == 0, "dgl_fc_bwd: x / w must be 32-byte aligned, g / dx / dw 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const FcRange rg{k_begin, k_end};
  int rc;
  {
    const int groups = (N + FCX_ROWS - 1) / FCX_ROWS;
    int splits = fc_grid_splits((k_end - k_begin + FCX_KC - 1) / FCX_KC) / groups;
    if (splits < 1) splits = 1;
    const size_t smem = 2 * (size_t)FCX_G_IMG + 4 * (size_t)FCX_W_IMG + 8 * 32 * 33 * sizeof(float) + 12 * 8 + 16;
    if ((rc = allow_smem(fc_dx_kernel, smem))) return rc;
    fc_dx_kernel<<<dim3(splits, groups), FC_THREADS, smem, st>>>(g, w, K, rg, N, dx);
    STEP_LAUNCH_CHECK("fc_dx_kernel");
  }
  {
    const int splits = fc_grid_splits((k_end - k_begin + FCW_KC - 1) / FCW_KC);
    const size_t smem = 2 * (size_t)FCW_G_IMG + 4 * (size_t)FCW_X_IMG + 12 * 8 + 16;
    if ((rc = allow_smem(fc_dw_kernel, smem))) return rc;
    fc_dw_kernel<<<splits, FC_THREADS, smem, st>>>(g, x, K, rg, N, dw_scale, dw);
  }
  return check_launch("fc_dw_kernel");
}
