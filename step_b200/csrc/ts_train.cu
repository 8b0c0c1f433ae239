// TSFormer pre-training (stage 1 of STEP) - the kernels the masked auto-encoder's BACKWARD needs, fp32:
//   attention backward (dQ, dK, dV of softmax(QK^T/sqrt(24))V with the forward's dropout mask regenerated from its counters),
//   residual + LayerNorm forward (keeping the row statistics) and backward, counter-based dropout (forward == backward).
// Reference: step/step_arch/tsformer/tsformer.py:71-160 (encoding with the 75 % mask, decoding, reconstruction),
// transformer_layers.py:10-20 (nn.TransformerEncoderLayer, post-norm); the reference differentiates these through
// autograd over cuBLAS / ATen kernels.  The dense layers' gradients run on the split-bf16 tcgen05 GEMM (tc_gemm.cu).
#include <math.h>
#include "common.cuh"

namespace stepk {

constexpr int TT_D = 96, TT_H = 4, TT_HD = 24, TT_CHUNK = 8;   // TT_CHUNK = the forward kernel's key chunk (dropout counters)

// keep-mask of attention probability (row i, key j) exactly as attn_fwd_kernel draws it (ts_encoder.cu)
__device__ __forceinline__ bool attn_keep(uint64_t rowbase, int j, uint32_t thr, uint64_t key) {
  const uint4 r = philox4x32(rowbase + (uint64_t)(j & ~3), key);
  const uint32_t w = (j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w));
  return w >= thr;
}

// ---------------------------------------------------------------------------
// pass A: one thread per query row.  L_i = log2-sum-exp of the scaled scores, D_i = dO_i . O_i,
// dQ_i = sum_j dS_ij K_j / sqrt(d),  dS_ij = P_ij (dPd_ij m_ij / (1-p) - D_i),  dPd_ij = dO_i . V_j
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const float *__restrict__ qkv, const float *__restrict__ out, const float *__restrict__ dout,
                                  int P, int Ppad, uint32_t thr, float dscale, uint64_t key, float *__restrict__ dqkv,
                                  float *__restrict__ LD /*[S,4,P,2]*/) {
  extern __shared__ __align__(16) float sm[];
  float *Ks = sm, *Vs = sm + (size_t)P * TT_HD;
  const int s = blockIdx.x, h = blockIdx.y;
  const float *base = qkv + (size_t)s * P * (3 * TT_D);
  for (int idx = threadIdx.x; idx < P * 6; idx += blockDim.x) {
    const int r = idx / 6, part = idx % 6;
    const float *rowp = base + (size_t)r * (3 * TT_D) + h * TT_HD + part * 4;
    *reinterpret_cast<float4 *>(Ks + r * TT_HD + part * 4) = *reinterpret_cast<const float4 *>(rowp + TT_D);
    *reinterpret_cast<float4 *>(Vs + r * TT_HD + part * 4) = *reinterpret_cast<const float4 *>(rowp + 2 * TT_D);
  }
  __syncthreads();
  const float c = 0.20412414523193154f * 1.4426950408889634f;       // log2(e) / sqrt(24)
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    float q[TT_HD], go[TT_HD], dq[TT_HD];
    const float *qp = base + (size_t)i * (3 * TT_D) + h * TT_HD;
    const float *op = out + ((size_t)s * P + i) * TT_D + h * TT_HD, *gp = dout + ((size_t)s * P + i) * TT_D + h * TT_HD;
    float D = 0.f;
#pragma unroll
    for (int d = 0; d < TT_HD; ++d) { q[d] = qp[d] * c; go[d] = gp[d]; D = fmaf(go[d], op[d], D); dq[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < P; ++j) {
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < TT_HD; ++d) sc = fmaf(q[d], Ks[j * TT_HD + d], sc);
      const float mn = fmaxf(m, sc);
      l = l * exp2f(m - mn) + exp2f(sc - mn);
      m = mn;
    }
    const float L = m + log2f(l);
    const uint64_t rowbase = ((uint64_t)(s * TT_H + h) * P + i) * Ppad;
    for (int j = 0; j < P; ++j) {
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < TT_HD; ++d) { sc = fmaf(q[d], Ks[j * TT_HD + d], sc); dp = fmaf(go[d], Vs[j * TT_HD + d], dp); }
      const float p = exp2f(sc - L);
      if (thr) dp = attn_keep(rowbase, j, thr, key) ? dp * dscale : 0.f;
      const float ds = p * (dp - D);
#pragma unroll
      for (int d = 0; d < TT_HD; ++d) dq[d] = fmaf(ds, Ks[j * TT_HD + d], dq[d]);
    }
    float *o = dqkv + ((size_t)s * P + i) * (3 * TT_D) + h * TT_HD;
#pragma unroll
    for (int d = 0; d < TT_HD; ++d) o[d] = dq[d] * 0.20412414523193154f;
    float *ld = LD + (((size_t)s * TT_H + h) * P + i) * 2;
    ld[0] = L; ld[1] = D;
  }
}

// ---------------------------------------------------------------------------
// pass B: one thread per key row.  dV_j = sum_i Pd_ij dO_i,  dK_j = sum_i dS_ij Q_i / sqrt(d)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const float *__restrict__ qkv, const float *__restrict__ dout, const float *__restrict__ LD,
                                   int P, int Ppad, uint32_t thr, float dscale, uint64_t key, float *__restrict__ dqkv) {
  extern __shared__ __align__(16) float sm[];
  float *Qs = sm, *Gs = sm + (size_t)P * TT_HD, *Ls = Gs + (size_t)P * TT_HD;      // Q (scaled), dO, (L, D)
  const int s = blockIdx.x, h = blockIdx.y;
  const float *base = qkv + (size_t)s * P * (3 * TT_D);
  const float c = 0.20412414523193154f * 1.4426950408889634f;
  for (int idx = threadIdx.x; idx < P * 6; idx += blockDim.x) {
    const int r = idx / 6, part = idx % 6;
    float4 qv = *reinterpret_cast<const float4 *>(base + (size_t)r * (3 * TT_D) + h * TT_HD + part * 4);
    qv.x *= c; qv.y *= c; qv.z *= c; qv.w *= c;
    *reinterpret_cast<float4 *>(Qs + r * TT_HD + part * 4) = qv;
    *reinterpret_cast<float4 *>(Gs + r * TT_HD + part * 4) =
        *reinterpret_cast<const float4 *>(dout + ((size_t)s * P + r) * TT_D + h * TT_HD + part * 4);
  }
  for (int idx = threadIdx.x; idx < 2 * P; idx += blockDim.x) Ls[idx] = LD[((size_t)s * TT_H + h) * P * 2 + idx];
  __syncthreads();
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    float k[TT_HD], v[TT_HD], dk[TT_HD], dv[TT_HD];
    const float *kp = base + (size_t)j * (3 * TT_D) + TT_D + h * TT_HD;
#pragma unroll
    for (int d = 0; d < TT_HD; ++d) { k[d] = kp[d]; v[d] = kp[TT_D + d]; dk[d] = 0.f; dv[d] = 0.f; }
    for (int i = 0; i < P; ++i) {
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < TT_HD; ++d) { sc = fmaf(Qs[i * TT_HD + d], k[d], sc); dp = fmaf(Gs[i * TT_HD + d], v[d], dp); }
      const float p = exp2f(sc - Ls[2 * i]);
      float pd = p;
      if (thr) {
        const bool keep = attn_keep(((uint64_t)(s * TT_H + h) * P + i) * Ppad, j, thr, key);
        pd = keep ? p * dscale : 0.f;
        dp = keep ? dp * dscale : 0.f;
      }
      const float ds = p * (dp - Ls[2 * i + 1]);
#pragma unroll
      for (int d = 0; d < TT_HD; ++d) {
        dv[d] = fmaf(pd, Gs[i * TT_HD + d], dv[d]);
        dk[d] = fmaf(ds, Qs[i * TT_HD + d], dk[d]);        // Qs carries log2(e)/sqrt(d): divide the log2(e) back out below
      }
    }
    float *o = dqkv + ((size_t)s * P + j) * (3 * TT_D) + TT_D + h * TT_HD;
#pragma unroll
    for (int d = 0; d < TT_HD; ++d) { o[d] = dk[d] * 0.6931471805599453f; o[TT_D + d] = dv[d]; }
  }
}

// ---------------------------------------------------------------------------
// y = LayerNorm(x (+ r)) * w + b over 96 features, one warp per row; keeps s = x + r and (mean, rstd) for backward
// ---------------------------------------------------------------------------
__global__ void add_ln_fwd_kernel(const float *__restrict__ x, const float *__restrict__ r, const float *__restrict__ w,
                                  const float *__restrict__ b, long long M, float *__restrict__ sum, float *__restrict__ stat,
                                  float *__restrict__ y) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = x[row * TT_D + lane + 32 * i] + (r ? r[row * TT_D + lane + 32 * i] : 0.f);
    if (sum) sum[row * TT_D + lane + 32 * i] = v[i];
  }
  const float mean = warp_sum(v[0] + v[1] + v[2]) * (1.f / 96.f);
  const float d0 = v[0] - mean, d1 = v[1] - mean, d2 = v[2] - mean;
  const float rstd = 1.0f / sqrtf(warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.f / 96.f) + 1e-5f);
  if (lane == 0 && stat) { stat[2 * row] = mean; stat[2 * row + 1] = rstd; }
  y[row * TT_D + lane] = d0 * rstd * w[lane] + b[lane];
  y[row * TT_D + lane + 32] = d1 * rstd * w[lane + 32] + b[lane + 32];
  y[row * TT_D + lane + 64] = d2 * rstd * w[lane + 64] + b[lane + 64];
}

// dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dy * w;  dw += dy xhat, db += dy  (per-block partials, then atomics)
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ sum,
                                                         const float *__restrict__ stat, const float *__restrict__ w, long long M,
                                                         float *__restrict__ dx, float *__restrict__ dw, float *__restrict__ db) {
  __shared__ float red[2][8][TT_D];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  float aw[3] = {0.f, 0.f, 0.f}, ab[3] = {0.f, 0.f, 0.f};
  for (long long row = (long long)blockIdx.x * 8 + wp; row < M; row += 8LL * gridDim.x) {
    const float mean = stat[2 * row], rstd = stat[2 * row + 1];
    float g[3], xh[3], dyv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dyv[i] = dy[row * TT_D + lane + 32 * i];
      xh[i] = (sum[row * TT_D + lane + 32 * i] - mean) * rstd;
      g[i] = dyv[i] * w[lane + 32 * i];
      aw[i] = fmaf(dyv[i], xh[i], aw[i]);
      ab[i] += dyv[i];
    }
    const float mg = warp_sum(g[0] + g[1] + g[2]) * (1.f / 96.f);
    const float mgx = warp_sum(g[0] * xh[0] + g[1] * xh[1] + g[2] * xh[2]) * (1.f / 96.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) dx[row * TT_D + lane + 32 * i] = rstd * (g[i] - mg - xh[i] * mgx);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) { red[0][wp][lane + 32 * i] = aw[i]; red[1][wp][lane + 32 * i] = ab[i]; }
  __syncthreads();
  if (threadIdx.x < 2 * TT_D) {
    const int which = threadIdx.x / TT_D, f = threadIdx.x % TT_D;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[which][i][f];
    atomicAdd((which ? db : dw) + f, s);
  }
}

// y = x * keep / (1 - p), keep drawn from Philox(seed, site) per group of 4 elements: the same call is its own backward
__global__ void dropout_kernel(const float *__restrict__ x, long long n4, uint32_t thr, float scale, uint64_t key,
                               float *__restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uint4 r = philox4x32((uint64_t)i, key);
  float4 v = reinterpret_cast<const float4 *>(x)[i];
  v.x = r.x >= thr ? v.x * scale : 0.f; v.y = r.y >= thr ? v.y * scale : 0.f;
  v.z = r.z >= thr ? v.z * scale : 0.f; v.w = r.w >= thr ? v.w * scale : 0.f;
  reinterpret_cast<float4 *>(y)[i] = v;
}

}  // namespace stepk

using namespace stepk;

extern "C" int step_attn_bwd_f32(const float *qkv, const float *out, const float *dout, int S, int P, float drop_p,
                                 unsigned long long seed, unsigned drop_site, float *scratch, float *dqkv, void *stream) {
  STEP_REQUIRE(qkv && out && dout && scratch && dqkv && S > 0 && P > 0, "attn_bwd: bad argument");
  STEP_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attn_bwd: drop_p must be in [0, 1)");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ppad = (P + TT_CHUNK - 1) / TT_CHUNK * TT_CHUNK;
  uint32_t thr = 0; float scale = 1.f;
  if (drop_p > 0.f) { thr = drop_threshold(drop_p); scale = 1.0f / (1.0f - drop_p); }
  const uint64_t key = rng_key(seed, drop_site);
  int threads = (P + 31) / 32 * 32;
  if (threads > 256) threads = 256;       // ~200 registers per thread (24-wide q/k/v/gradient rows live in registers)
  const size_t smA = (size_t)P * TT_HD * 2 * sizeof(float), smB = ((size_t)P * TT_HD * 2 + 2 * (size_t)P) * sizeof(float);
  if (smB > 220 * 1024) return fail(STEP_EUNSUPPORTED, "attn_bwd: P=%lld does not fit shared memory", P);
  int rc;
  if ((rc = allow_smem(attn_bwd_q_kernel, 220 * 1024))) return rc;
  if ((rc = allow_smem(attn_bwd_kv_kernel, 220 * 1024))) return rc;
  attn_bwd_q_kernel<<<dim3(S, TT_H), threads, smA, st>>>(qkv, out, dout, P, Ppad, thr, scale, key, dqkv, scratch);
  STEP_LAUNCH_CHECK("attn_bwd_q_kernel");
  attn_bwd_kv_kernel<<<dim3(S, TT_H), threads, smB, st>>>(qkv, dout, scratch, P, Ppad, thr, scale, key, dqkv);
  return check_launch("attn_bwd_kv_kernel");
}

extern "C" int step_add_layernorm96_fwd(const float *x, const float *r, const float *w, const float *b, long long M, float *sum,
                                        float *stat, float *y, void *stream) {
  STEP_REQUIRE(x && w && b && y && M > 0, "add_layernorm_fwd: bad argument");
  add_ln_fwd_kernel<<<(unsigned)((M + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, r, w, b, M, sum, stat, y);
  return check_launch("add_ln_fwd_kernel");
}

extern "C" int step_add_layernorm96_bwd(const float *dy, const float *sum, const float *stat, const float *w, long long M,
                                        float *dx, float *dw, float *db, void *stream) {
  STEP_REQUIRE(dy && sum && stat && w && dx && dw && db && M > 0, "add_layernorm_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dw, 0, TT_D * sizeof(float), st);
  cudaMemsetAsync(db, 0, TT_D * sizeof(float), st);
  long long grid = (M + 7) / 8;
  if (grid > 592) grid = 592;
  add_ln_bwd_kernel<<<(unsigned)grid, 256, 0, st>>>(dy, sum, stat, w, M, dx, dw, db);
  return check_launch("add_ln_bwd_kernel");
}

extern "C" int step_dropout_f32(const float *x, long long n, float drop_p, unsigned long long seed, unsigned site, float *y,
                                void *stream) {
  STEP_REQUIRE(x && y && n > 0 && n % 4 == 0, "dropout: n must be a positive multiple of 4");
  STEP_REQUIRE(drop_p > 0.f && drop_p < 1.f, "dropout: drop_p must be in (0, 1)");
  dropout_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n / 4, drop_threshold(drop_p),
                                                                                   1.0f / (1.0f - drop_p), rng_key(seed, site), y);
  return check_launch("dropout_kernel");
}
