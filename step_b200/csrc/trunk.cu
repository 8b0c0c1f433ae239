// Discrete-graph-learning "global feature" trunk, convolutional part, fused forward + backward (fp32):
//   x [N,1,L0] -> Conv1d(1,8,10) -> ReLU -> BatchNorm1d(8) -> Conv1d(8,16,10) -> ReLU -> BatchNorm1d(16) -> [N, 16*L2]
// Reference: step/step_arch/discrete_graph_learning.py:131-133 (conv1/bn1/conv2/bn2; the module is in train(), so
// both BatchNorms use batch statistics over the N nodes x L positions).
//
// The reference (cuDNN + ATen) materialises y1 = relu(conv1) [N,8,L1] and its normalised copy plus a dozen
// elementwise temporaries of the 318 MB conv2 output.  Here y1 is never stored: conv1 is 10 MACs per value, so
// it is recomputed inside every consumer (statistics pass, conv2 forward, conv2 backward, conv1 backward) from the
// 20 MB input series, and BatchNorm1 is an affine applied on the fly.  Stored tensors: y2 (pre-BN2, needed for the
// ReLU mask and x-hat in backward) and y2n (the Linear's input).
#include <math.h>
#include "common.cuh"
#include "trunk_tc.cuh"

namespace stepk {

constexpr int TK = 10;          // conv kernel width
constexpr int C1 = 8, C2 = 16;
constexpr int TL = 512;         // positions per tile
constexpr int HALO = TK - 1;

struct TrunkDims { int N, L0, L1, L2; };

__device__ __forceinline__ void block_reduce_add_double(float v, double *dst, float *red) {
  // red: >= 32 floats of smem; all threads call; adds the block sum of v to *dst
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float s = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.f;
    s = warp_sum(s);
    if (l == 0) atomicAdd(dst, (double)s);
  }
}

// ---------------------------------------------------------------------------
// F1: batch statistics of y1 = relu(conv1(x)) without storing it.  grid (ceil(L1/1024), N), 256 threads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trunk_conv1_stats_kernel(const float *__restrict__ x, TrunkDims d,
                                                                const float *__restrict__ w1, const float *__restrict__ b1,
                                                                double *__restrict__ sums /*[2][8]*/) {
  __shared__ float xs[1024 + HALO];
  __shared__ float ws[C1 * TK + C1];
  __shared__ float red[32];
  // grid (gx, N): each CTA loops over 1024-position tiles of its node and reduces ONCE at the end (one CTA per tile
  // put ~5000 CTAs x 16 double atomics on two cache lines: the atomics, not the math, set the run time)
  const int n = blockIdx.y, tid = threadIdx.x;
  const float *xr = x + (size_t)n * d.L0;
  if (tid < C1 * TK) ws[tid] = w1[tid];
  if (tid < C1) ws[C1 * TK + tid] = b1[tid];
  float s[C1], q[C1];
#pragma unroll
  for (int c = 0; c < C1; ++c) { s[c] = 0.f; q[c] = 0.f; }
  for (int t0 = blockIdx.x * 1024; t0 < d.L1; t0 += gridDim.x * 1024) {
  __syncthreads();
  for (int i = tid; i < 1024 + HALO; i += 256) xs[i] = (t0 + i < d.L0) ? xr[t0 + i] : 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = tid + 256 * j;
    if (t0 + p < d.L1) {
      float xv[TK];
#pragma unroll
      for (int k = 0; k < TK; ++k) xv[k] = xs[p + k];
#pragma unroll
      for (int c = 0; c < C1; ++c) {
        float a = ws[C1 * TK + c];
#pragma unroll
        for (int k = 0; k < TK; ++k) a = fmaf(ws[c * TK + k], xv[k], a);
        a = fmaxf(a, 0.f);
        s[c] += a;
        q[c] = fmaf(a, a, q[c]);
      }
    }
  }
  }  // tile loop
#pragma unroll
  for (int c = 0; c < C1; ++c) {
    block_reduce_add_double(s[c], sums + c, red);
    block_reduce_add_double(q[c], sums + C1 + c, red);
  }
}

// sums [2][C] (double) -> stats [4][C] floats: mean, biased var, scale = gamma*rstd, shift = beta - mean*scale
__global__ void trunk_bn_finalize_kernel(const double *sums, double count, int C, const float *gamma, const float *beta,
                                         float eps, float *stats) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  stats[c] = (float)mean;
  stats[C + c] = (float)var;
  stats[2 * C + c] = gamma[c] * rstd;
  stats[3 * C + c] = beta[c] - (float)mean * gamma[c] * rstd;
}

// ---------------------------------------------------------------------------
// shared tile helpers
// ---------------------------------------------------------------------------
// y1 (raw, post-ReLU) for `count` positions starting at y1-position p0 into dst[c][stride]; positions >= L1 give 0
__device__ __forceinline__ void compute_y1_tile(const float *xs /* x at [p0, p0+count+HALO) */, const float *w1s,
                                                const float *b1s, int p0, int count, int L1, float *dst, int stride,
                                                const float *scale, const float *shift /* null: raw */) {
  for (int idx = threadIdx.x; idx < C1 * count; idx += blockDim.x) {
    const int c = idx / count, j = idx - c * count;
    float a = 0.f;
    if (p0 + j < L1 && p0 + j >= 0) {
      a = b1s[c];
#pragma unroll
      for (int k = 0; k < TK; ++k) a = fmaf(w1s[c * TK + k], xs[j + k], a);
      a = fmaxf(a, 0.f);
      if (scale != nullptr) a = fmaf(a, scale[c], shift[c]);
    }
    dst[c * stride + j] = a;
  }
}

// ---------------------------------------------------------------------------
// F2: y2 = relu(conv2(BN1(relu(conv1 x)))) (pre-BN2) + BN2 batch sums.  grid (ceil(L2/512), N), 256 threads.
// ---------------------------------------------------------------------------
constexpr int Y1W = TL + HALO;          // 521 y1 positions per tile
constexpr int Y1S = Y1W + 3;            // smem row stride

__global__ void __launch_bounds__(256, 3) trunk_conv2_fwd_kernel(const float *__restrict__ x, TrunkDims d,
                                                              const float *__restrict__ w1, const float *__restrict__ b1,
                                                              const float *__restrict__ bn1 /*[4][8]*/,
                                                              const float *__restrict__ w2, const float *__restrict__ b2,
                                                              float *__restrict__ y2, double *__restrict__ sums2 /*[2][16] or null*/) {
  __shared__ float xs2[2][Y1W + HALO + 3];   // double-buffered series tile (cp.async prefetch of the next tile)
  __shared__ float y1s[C1 * Y1S];
  __shared__ __align__(16) float w2s[C1 * TK * C2];   // [ci][k][co]
  __shared__ float w1s[C1 * TK], b1s[C1], sc1[C1], sh1[C1], b2s[C2];
  __shared__ float red[32];
  const int n = blockIdx.y, tid = threadIdx.x;
  const float *xr = x + (size_t)n * d.L0;
  const int ntiles = (d.L2 + TL - 1) / TL;
  auto prefetch_x = [&](int tile, int buf) {
    const int t0 = tile * TL;
    for (int i = tid; i < Y1W + HALO; i += 256) {
      const bool ok = t0 + i < d.L0;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&xs2[buf][i]);
      const uint32_t nbytes = ok ? 4u : 0u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(xr + (ok ? t0 + i : 0)), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < ntiles) prefetch_x(blockIdx.x, 0);
  for (int i = tid; i < C2 * C1 * TK; i += 256) {
    const int co = i / (C1 * TK), r = i - co * (C1 * TK), ci = r / TK, k = r - ci * TK;
    w2s[(ci * TK + k) * C2 + co] = w2[i];
  }
  if (tid < C1 * TK) w1s[tid] = w1[tid];
  if (tid < C1) { b1s[tid] = b1[tid]; sc1[tid] = bn1[2 * C1 + tid]; sh1[tid] = bn1[3 * C1 + tid]; }
  if (tid < C2) b2s[tid] = b2[tid];
  float s[C2], q[C2];          // BN2 batch-sum partials, kept across this CTA's tiles
#pragma unroll
  for (int co = 0; co < C2; ++co) { s[co] = 0.f; q[co] = 0.f; }
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int t0 = tile * TL;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                       // series tile landed; previous tile's y1s reads are done
    if (tile + (int)gridDim.x < ntiles) prefetch_x(tile + gridDim.x, buf ^ 1);
    compute_y1_tile(xs2[buf], w1s, b1s, t0, Y1W, d.L1, y1s, Y1S, sc1, sh1);
    __syncthreads();

    float acc[2][C2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int co = 0; co < C2; ++co) acc[j][co] = b2s[co];
    const int p0 = tid, p1 = tid + 256;
#pragma unroll 1
    for (int ci = 0; ci < C1; ++ci) {
      const float *yr = y1s + ci * Y1S;
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float a0 = yr[p0 + k], a1 = yr[p1 + k];
        const float4 *w = reinterpret_cast<const float4 *>(w2s + (ci * TK + k) * C2);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 ww = w[c4];
          ffma2(acc[0][4 * c4], acc[0][4 * c4 + 1], ww.x, ww.y, a0, a0);
          ffma2(acc[0][4 * c4 + 2], acc[0][4 * c4 + 3], ww.z, ww.w, a0, a0);
          ffma2(acc[1][4 * c4], acc[1][4 * c4 + 1], ww.x, ww.y, a1, a1);
          ffma2(acc[1][4 * c4 + 2], acc[1][4 * c4 + 3], ww.z, ww.w, a1, a1);
        }
      }
    }
    const bool ok0 = t0 + p0 < d.L2, ok1 = t0 + p1 < d.L2;
    float *yo = y2 + (size_t)n * C2 * d.L2 + t0;
#pragma unroll
    for (int co = 0; co < C2; ++co) {
      const float v0 = fmaxf(acc[0][co], 0.f), v1 = fmaxf(acc[1][co], 0.f);
      if (ok0) { yo[(size_t)co * d.L2 + p0] = v0; s[co] += v0; q[co] = fmaf(v0, v0, q[co]); }
      if (ok1) { yo[(size_t)co * d.L2 + p1] = v1; s[co] += v1; q[co] = fmaf(v1, v1, q[co]); }
    }
  }
  if (sums2 != nullptr) {
#pragma unroll
    for (int co = 0; co < C2; ++co) {
      block_reduce_add_double(s[co], sums2 + co, red);
      block_reduce_add_double(q[co], sums2 + C2 + co, red);
    }
  }
}

// F3: y2n = y2 * scale[c] + shift[c]   ([N][C][L] layout).  grid (ceil(L/1024), N*C): one (node, channel) row per
// blockIdx.y, float4 per thread when the rows are 16-byte aligned (L % 4 == 0), scalar otherwise.
__global__ void __launch_bounds__(256) trunk_bn_apply_kernel(const float *__restrict__ y, int C, int L,
                                                             const float *__restrict__ stats, float *__restrict__ out) {
  const int row = blockIdx.y, c = row % C;
  const float sc = stats[2 * C + c], sh = stats[3 * C + c];
  const float *yr = y + (size_t)row * L;
  float *orow = out + (size_t)row * L;
  if ((L & 3) == 0) {
    const int i = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (i < L) {
      const float4 v = *reinterpret_cast<const float4 *>(yr + i);
      *reinterpret_cast<float4 *>(orow + i) = make_float4(fmaf(v.x, sc, sh), fmaf(v.y, sc, sh), fmaf(v.z, sc, sh), fmaf(v.w, sc, sh));
    }
  } else {
    for (int i = blockIdx.x * 1024 + threadIdx.x; i < min(L, (int)(blockIdx.x + 1) * 1024); i += 256) orow[i] = fmaf(yr[i], sc, sh);
  }
}

// ---------------------------------------------------------------------------
// B1: BatchNorm backward sums for one channel per blockIdx.x: S1 = sum dy, S2 = sum dy * xhat.
// grid (C, 64): blockIdx.y strides over the N rows.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trunk_bn_bwd_stats_kernel(const float *__restrict__ dy, const float *__restrict__ y, int N,
                                                                 int C, int L, const float *__restrict__ stats, float eps,
                                                                 double *__restrict__ sums /*[2][C]*/) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const float mean = stats[c], rstd = 1.0f / sqrtf(stats[C + c] + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const float *dr = dy + ((size_t)n * C + c) * L, *yr = y + ((size_t)n * C + c) * L;
    for (int l = threadIdx.x; l < L; l += 256) {
      const float g = dr[l];
      s1 += g;
      s2 = fmaf(g, (yr[l] - mean) * rstd, s2);
    }
  }
  block_reduce_add_double(s1, sums + c, red);
  block_reduce_add_double(s2, sums + C + c, red);
}

// coefficient table for BN backward: coef [5][C] = gamma*rstd, S1/M, S2/M, mean, rstd; also dgamma = S2, dbeta = S1
__global__ void trunk_bn_bwd_finalize_kernel(const double *sums, double count, int C, const float *gamma, const float *stats,
                                             float eps, float *coef, float *dgamma, float *dbeta) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const float rstd = 1.0f / sqrtf(stats[C + c] + eps);
  coef[c] = gamma[c] * rstd;
  coef[C + c] = (float)(sums[c] / count);
  coef[2 * C + c] = (float)(sums[C + c] / count);
  coef[3 * C + c] = stats[c];
  coef[4 * C + c] = rstd;
  dgamma[c] = (float)sums[C + c];
  dbeta[c] = (float)sums[c];
}

// ---------------------------------------------------------------------------
// B2: backward through BN2 -> ReLU -> conv2: d(y1n) [N,8,L1], dW2, db2, and the BN1 backward sums.
// grid (6, N): each CTA walks over tiles of 256 y1 positions of one node (one position per thread), keeps its
// dW2 / db2 / BN-sum partials in registers across tiles and flushes them with one round of atomics.
// 54 KB of dynamic smem per CTA -> 4 CTAs per SM.
// ---------------------------------------------------------------------------
constexpr int TLB = 256;
constexpr int DPW = TLB + HALO;    // dpre2 positions [t0-9, t0+256)
constexpr int DPS = DPW + 3;
constexpr int Y1WB = TLB + HALO;   // normalised y1 positions [t0, t0+265)
constexpr int Y1SB = Y1WB + 3;

__global__ void __launch_bounds__(256) trunk_conv2_bwd_kernel(const float *__restrict__ x, TrunkDims d,
                                                              const float *__restrict__ w1, const float *__restrict__ b1,
                                                              const float *__restrict__ bn1 /*[4][8]*/, float eps,
                                                              const float *__restrict__ w2, const float *__restrict__ dy2n,
                                                              const float *__restrict__ y2, const float *__restrict__ coef2 /*[5][16]*/,
                                                              float *__restrict__ dy1n, float *__restrict__ dw2,
                                                              float *__restrict__ db2, double *__restrict__ sums1 /*[2][8]*/) {
  extern __shared__ __align__(16) float sm[];
  float *dp_cl = sm;                         // [16][DPS]      dpre2 by channel (for the transposed conv)
  float *dp_lc = dp_cl + C2 * DPS;           // [DPW][16]      dpre2 by position (broadcast reads for dW2)
  float *y1s = dp_lc + DPW * C2;             // [8][Y1SB]      normalised y1 at [t0, t0+265)
  float *y1r = y1s + C1 * Y1SB;              // [8][TLB]       raw y1 at [t0, t0+256) for x-hat
  float *xs = y1r + C1 * TLB;                // [Y1WB + HALO + 2]
  float *w2t = xs + (Y1WB + HALO + 2 + 3) / 4 * 4;   // [co][k][ci]  (8 contiguous ci)
  float *w1s = w2t + C2 * TK * C1;           // 80
  float *b1s = w1s + C1 * TK;                // 8
  float *sc1 = b1s + C1, *sh1 = sc1 + C1;    // 8 + 8
  float *cf2 = sh1 + C1;                     // 5 * 16 BN2 backward coefficients
  float *red = cf2 + 5 * C2;                 // 32
  float *mean1 = red + 32, *rstd1 = mean1 + C1;   // BN1 mean / rstd for x-hat
  float *rawy = rstd1 + C1;                  // [16][DPS] y2 of the NEXT tile, landed by cp.async during this tile's math
  float *rawd = rawy + C2 * DPS;             // [16][DPS] dy2n, same
  const int n = blockIdx.y, tid = threadIdx.x;
  const float *xr = x + (size_t)n * d.L0;
  for (int i = tid; i < C2 * C1 * TK; i += 256) {
    const int co = i / (C1 * TK), r = i - co * (C1 * TK), ci = r / TK, k = r - ci * TK;
    w2t[(co * TK + k) * C1 + ci] = w2[i];
  }
  if (tid < C1) { mean1[tid] = bn1[tid]; rstd1[tid] = 1.0f / sqrtf(bn1[C1 + tid] + eps); }
  if (tid < C1 * TK) w1s[tid] = w1[tid];
  if (tid < C1) { b1s[tid] = b1[tid]; sc1[tid] = bn1[2 * C1 + tid]; sh1[tid] = bn1[3 * C1 + tid]; }
  if (tid < 5 * C2) cf2[tid] = coef2[tid];
  // warps 0-3 own d(y1n) (two adjacent positions per thread), warps 4-7 own dW2/db2 (lane = (ci, 4 output channels),
  // all 10 taps, one 64-position segment per warp): both halves do 2560 FMAs per thread and tile, and the register
  // tiling keeps shared-memory traffic at ~0.1 wavefronts per FFMA (it was the bound at 0.7).
  const bool is_dx = tid < 128;
  const int wl = tid & 31, wci = wl >> 2, wcog = wl & 3, wseg = (tid >> 5) - 4;
  float accw[TK][4], accb[4], s1acc[C1], s2acc[C1];
#pragma unroll
  for (int k = 0; k < TK; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) accw[k][c] = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) accb[c] = 0.f;
#pragma unroll
  for (int ci = 0; ci < C1; ++ci) { s1acc[ci] = 0.f; s2acc[ci] = 0.f; }
  const int ntiles = (d.L1 + TLB - 1) / TLB;
  // raw y2 / dy2n of one tile -> shared memory with 4-byte cp.async (rows start at t0 - 9: no 16-byte alignment for
  // bulk copies); positions outside [0, L2) are zero-filled, which makes dpre2 zero there.
  auto prefetch_tile = [&](int tile) {
    const int t0 = tile * TLB;
    for (int idx = tid; idx < C2 * DPW; idx += 256) {
      const int co = idx / DPW, j = idx - co * DPW, l = t0 - HALO + j;
      const bool ok = l >= 0 && l < d.L2;
      const size_t off = ((size_t)n * C2 + co) * d.L2 + (ok ? l : 0);
      const uint32_t nbytes = ok ? 4u : 0u;
      const uint32_t sy = (uint32_t)__cvta_generic_to_shared(rawy + co * DPS + j);
      const uint32_t sd = (uint32_t)__cvta_generic_to_shared(rawd + co * DPS + j);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sy), "l"(y2 + off), "r"(nbytes) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sd), "l"(dy2n + off), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < ntiles) prefetch_tile(blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int t0 = tile * TLB;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int i = tid; i < Y1WB + HALO; i += 256) xs[i] = (t0 + i < d.L0) ? xr[t0 + i] : 0.f;
    // dpre2 tile: position j <-> l = t0 - 9 + j (zero-filled raw values give v = 0 outside [0, L2))
    for (int idx = tid; idx < C2 * DPW; idx += 256) {
      const int co = idx / DPW, j = idx - co * DPW;
      const float yv = rawy[co * DPS + j];
      float v = 0.f;
      if (yv > 0.f) {
        const float xhat = (yv - cf2[3 * C2 + co]) * cf2[4 * C2 + co];
        v = cf2[co] * (rawd[co * DPS + j] - cf2[C2 + co] - xhat * cf2[2 * C2 + co]);
      }
      dp_cl[co * DPS + j] = v;
      dp_lc[j * C2 + co] = v;
    }
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) prefetch_tile(tile + gridDim.x);     // overlaps with everything below
    // y1 (raw for x-hat on the own range, normalised everywhere) with ONE conv1 evaluation per element
    for (int idx = tid; idx < C1 * Y1WB; idx += 256) {
      const int c = idx / Y1WB, j = idx - c * Y1WB;
      float a = 0.f, an = 0.f;
      if (t0 + j < d.L1) {
        a = b1s[c];
#pragma unroll
        for (int k = 0; k < TK; ++k) a = fmaf(w1s[c * TK + k], xs[j + k], a);
        a = fmaxf(a, 0.f);
        an = fmaf(a, sc1[c], sh1[c]);
      }
      y1s[c * Y1SB + j] = an;
      if (j < TLB) y1r[c * TLB + j] = a;
    }
    __syncthreads();

    if (is_dx) {
      // ---- d(y1n)[ci][l'] = sum_co sum_k w2[co][ci][k] dpre2[co][l'-k];  l' = t0 + p, p in {p0, p0 + 1} ----
      float acc0[C1], acc1[C1];
#pragma unroll
      for (int ci = 0; ci < C1; ++ci) { acc0[ci] = 0.f; acc1[ci] = 0.f; }
      const int p0 = 2 * tid;
#pragma unroll 1
      for (int co = 0; co < C2; ++co) {
        // v[i] = dpre2[co] at own-range position p0 - 9 + i (smem index p0 + i: HALO == 9 keeps it 8-byte aligned)
        const float *dr = dp_cl + co * DPS + p0;
        float v[TK + 2];
#pragma unroll
        for (int i = 0; i < TK; i += 2) {
          const float2 t = *reinterpret_cast<const float2 *>(dr + i);
          v[i] = t.x; v[i + 1] = t.y;
        }
        v[TK] = dr[TK];
#pragma unroll
        for (int k = 0; k < TK; ++k) {
          const float a0 = v[HALO - k], a1 = v[HALO + 1 - k];
          const float4 *w = reinterpret_cast<const float4 *>(w2t + (co * TK + k) * C1);
          const float4 wa = w[0], wb = w[1];
          ffma2(acc0[0], acc0[1], wa.x, wa.y, a0, a0); ffma2(acc0[2], acc0[3], wa.z, wa.w, a0, a0);
          ffma2(acc0[4], acc0[5], wb.x, wb.y, a0, a0); ffma2(acc0[6], acc0[7], wb.z, wb.w, a0, a0);
          ffma2(acc1[0], acc1[1], wa.x, wa.y, a1, a1); ffma2(acc1[2], acc1[3], wa.z, wa.w, a1, a1);
          ffma2(acc1[4], acc1[5], wb.x, wb.y, a1, a1); ffma2(acc1[6], acc1[7], wb.z, wb.w, a1, a1);
        }
      }
      float *o = dy1n + (size_t)n * C1 * d.L1 + t0;
      if (t0 + p0 < d.L1) {
#pragma unroll
        for (int ci = 0; ci < C1; ++ci) {
          o[(size_t)ci * d.L1 + p0] = acc0[ci];
          s1acc[ci] += acc0[ci];
          s2acc[ci] = fmaf(acc0[ci], (y1r[ci * TLB + p0] - mean1[ci]) * rstd1[ci], s2acc[ci]);
        }
      }
      if (t0 + p0 + 1 < d.L1) {
#pragma unroll
        for (int ci = 0; ci < C1; ++ci) {
          o[(size_t)ci * d.L1 + p0 + 1] = acc1[ci];
          s1acc[ci] += acc1[ci];
          s2acc[ci] = fmaf(acc1[ci], (y1r[ci * TLB + p0 + 1] - mean1[ci]) * rstd1[ci], s2acc[ci]);
        }
      }
    } else {
      // ---- dW2[co][ci][k] += sum_p dpre2[co][p] y1n[ci][p + k] over this warp's 64 positions; db2 on the ci == 0 lanes.
      // The 10-tap window of y1n slides through registers (rotating static indices, one new load per position).
      const int pbeg = wseg * 64;
      const float *yrow = y1s + wci * Y1SB + pbeg;
      const float *dprow = dp_lc + (size_t)(pbeg + HALO) * C2 + 4 * wcog;
      float y[TK];
#pragma unroll
      for (int k = 0; k < TK; ++k) y[k] = yrow[k];
      auto step10 = [&](int base, int count) {
#pragma unroll
        for (int i = 0; i < TK; ++i) {
          if (i < count) {
            const float4 g = *reinterpret_cast<const float4 *>(dprow + (size_t)(base + i) * C2);
#pragma unroll
            for (int k = 0; k < TK; ++k) {
              const float yv = y[(i + k) % TK];
              ffma2(accw[k][0], accw[k][1], g.x, g.y, yv, yv);
              ffma2(accw[k][2], accw[k][3], g.z, g.w, yv, yv);
            }
            if (wci == 0) { accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w; }
            y[i] = yrow[base + i + TK];
          }
        }
      };
#pragma unroll 1
      for (int base = 0; base < 60; base += TK) step10(base, TK);
      step10(60, 4);
    }
  }  // tile loop

  // flush: the four dW2 warps combine through shared memory (the tile buffers are free now), one atomic per entry and CTA
  __syncthreads();
  float *wred = sm;                          // [4 warps][C2*C1*TK] + [4][C2]
  if (!is_dx) {
    float *wr = wred + wseg * (C2 * C1 * TK);
#pragma unroll
    for (int k = 0; k < TK; ++k)
#pragma unroll
      for (int c = 0; c < 4; ++c) wr[((4 * wcog + c) * C1 + wci) * TK + k] = accw[k][c];
    if (wci == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) wred[4 * C2 * C1 * TK + wseg * C2 + 4 * wcog + c] = accb[c];
    }
  }
  __syncthreads();
  for (int i = tid; i < C2 * C1 * TK; i += 256)
    atomicAdd(dw2 + i, (wred[i] + wred[C2 * C1 * TK + i]) + (wred[2 * C2 * C1 * TK + i] + wred[3 * C2 * C1 * TK + i]));
  if (tid < C2) {
    const float *bb = wred + 4 * C2 * C1 * TK;
    atomicAdd(db2 + tid, (bb[tid] + bb[C2 + tid]) + (bb[2 * C2 + tid] + bb[3 * C2 + tid]));
  }
  __syncthreads();
#pragma unroll
  for (int ci = 0; ci < C1; ++ci) {
    block_reduce_add_double(s1acc[ci], sums1 + ci, red);
    block_reduce_add_double(s2acc[ci], sums1 + C1 + ci, red);
  }
}

// ---------------------------------------------------------------------------
// B3: backward through BN1 -> ReLU -> conv1: dW1 [8][10], db1 [8].  grid (ceil(L1/1024), N), 256 threads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) trunk_conv1_bwd_kernel(const float *__restrict__ x, TrunkDims d,
                                                              const float *__restrict__ w1, const float *__restrict__ b1,
                                                              const float *__restrict__ dy1n, const float *__restrict__ coef1 /*[5][8]*/,
                                                              float *__restrict__ dw1, float *__restrict__ db1) {
  __shared__ float xs[1024 + HALO];
  __shared__ float ws[C1 * TK + C1];
  __shared__ float accs[C1 * TK + C1];
  __shared__ float cf[5 * C1];           // BN1 backward coefficients
  // grid (gx, N): tile loop per CTA, shared-memory accumulators flushed once (see trunk_conv1_stats_kernel)
  const int n = blockIdx.y, tid = threadIdx.x;
  const float *xr = x + (size_t)n * d.L0;
  if (tid < C1 * TK) ws[tid] = w1[tid];
  if (tid < C1) ws[C1 * TK + tid] = b1[tid];
  if (tid < C1 * TK + C1) accs[tid] = 0.f;
  if (tid < 5 * C1) cf[tid] = coef1[tid];
  float gw[C1][TK], gb[C1];
#pragma unroll
  for (int c = 0; c < C1; ++c) {
    gb[c] = 0.f;
#pragma unroll
    for (int k = 0; k < TK; ++k) gw[c][k] = 0.f;
  }
  for (int t0 = blockIdx.x * 1024; t0 < d.L1; t0 += gridDim.x * 1024) {
  __syncthreads();
  for (int i = tid; i < 1024 + HALO; i += 256) xs[i] = (t0 + i < d.L0) ? xr[t0 + i] : 0.f;
  __syncthreads();
  // per-thread partials for all 8 channels stay in registers across this CTA's tiles (88 accumulators); the warp
  // reductions run ONCE per CTA - doing them per tile and channel made the kernel shuffle-throughput-bound
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = tid + 256 * j;
    if (t0 + p < d.L1) {
      float xv[TK];
#pragma unroll
      for (int k = 0; k < TK; ++k) xv[k] = xs[p + k];
      float dy[C1];
#pragma unroll
      for (int c = 0; c < C1; ++c) dy[c] = dy1n[((size_t)n * C1 + c) * d.L1 + t0 + p];     // 8 independent loads
#pragma unroll
      for (int c = 0; c < C1; ++c) {
        float a = ws[C1 * TK + c];
#pragma unroll
        for (int k = 0; k < TK; ++k) a = fmaf(ws[c * TK + k], xv[k], a);
        if (a > 0.f) {
          const float xhat = (a - cf[3 * C1 + c]) * cf[4 * C1 + c];
          const float g = cf[c] * (dy[c] - cf[C1 + c] - xhat * cf[2 * C1 + c]);
          gb[c] += g;
#pragma unroll
          for (int k = 0; k < TK; ++k) gw[c][k] = fmaf(g, xv[k], gw[c][k]);
        }
      }
    }
  }
  }  // tile loop
#pragma unroll
  for (int c = 0; c < C1; ++c) {
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float v = warp_sum(gw[c][k]);
      if ((tid & 31) == 0) atomicAdd(&accs[c * TK + k], v);
    }
    const float vb = warp_sum(gb[c]);
    if ((tid & 31) == 0) atomicAdd(&accs[C1 * TK + c], vb);
  }
  __syncthreads();
  if (tid < C1 * TK) atomicAdd(dw1 + tid, accs[tid]);
  else if (tid < C1 * TK + C1) atomicAdd(db1 + (tid - C1 * TK), accs[tid]);
}

// CTAs per node for the tile-loop kernels: fill whole waves of the resident CTA slots while keeping at least
// `min_tiles` tiles per CTA (so per-CTA prologue / final reductions are amortised) and the tile split even.
template <typename K>
static int wave_aware_ctas(K kernel, size_t smem, int ntiles, int N, int min_tiles) {
  int dev = 0, sms = 148, per_sm = 2;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem);
  if (per_sm < 1) per_sm = 1;
  const double slots = (double)sms * per_sm;
  int gmax = ntiles / min_tiles;
  if (gmax > 12) gmax = 12;
  if (gmax < 1) gmax = 1;
  int gx = gmax < 4 ? gmax : 4;
  double best = -1.0;
  for (int c = gx; c <= gmax; ++c) {
    const double waves = (double)N * c / slots;
    const double per_cta = (double)((ntiles + c - 1) / c) * c / ntiles;       // tile imbalance between CTAs
    const double eff = waves / ceil(waves) / per_cta;
    if (eff > best + 1e-9) { best = eff; gx = c; }
  }
  return gx;
}

static size_t conv2_bwd_smem() {
  size_t f = (size_t)C2 * DPS + (size_t)DPW * C2 + (size_t)C1 * Y1SB + (size_t)C1 * TLB + (Y1WB + HALO + 2 + 3) / 4 * 4 +
             C2 * TK * C1 + C1 * TK + 3 * C1 + 5 * C2 + 32 + 2 * C1 + 2 * (size_t)C2 * DPS;
  return f * sizeof(float);
}

}  // namespace stepk

using namespace stepk;

// scratch: 2*(2*8) + 2*(2*16) doubles then coefficient tables; the caller provides >= 4096 bytes, zeroed by us
extern "C" int step_dgl_conv_fwd(const float *x, int N, int L0, const float *w1, const float *b1, const float *g1,
                                 const float *be1, const float *w2, const float *b2, const float *g2, const float *be2,
                                 float eps, int training, float *bn1_stats /*[4][8]*/, float *bn2_stats /*[4][16]*/,
                                 float *y2, float *y2n, void *scratch, void *stream) {
  STEP_REQUIRE(x && w1 && b1 && g1 && be1 && w2 && b2 && g2 && be2 && bn1_stats && bn2_stats && y2 && y2n && scratch,
               "dgl_conv_fwd: null pointer");
  STEP_REQUIRE(N > 0 && N <= 65535 && L0 > 2 * HALO, "dgl_conv_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const TrunkDims d{N, L0, L0 - HALO, L0 - 2 * HALO};
  double *sums = reinterpret_cast<double *>(scratch);   // [0,16): bn1, [16,48): bn2
  if (training) {
    cudaError_t e = cudaMemsetAsync(sums, 0, 48 * sizeof(double), st);
    if (e != cudaSuccess) return fail_msg((int)e, cudaGetErrorString(e));
    trunk_conv1_stats_kernel<<<dim3(wave_aware_ctas(trunk_conv1_stats_kernel, 0, (d.L1 + 1023) / 1024, N, 6), N), 256, 0, st>>>(
        x, d, w1, b1, sums);
    STEP_LAUNCH_CHECK("trunk_conv1_stats_kernel");
    trunk_bn_finalize_kernel<<<1, 32, 0, st>>>(sums, (double)N * d.L1, C1, g1, be1, eps, bn1_stats);
    STEP_LAUNCH_CHECK("trunk_bn_finalize_kernel");
  }
  if (trunk_use_tc()) {
    // conv2 as an implicit GEMM on tcgen05 (trunk_tc.cuh): y1n planes rebuilt per tile, taps addressed in place
    TcConv2Args t{};
    t.x = x; t.w1 = w1; t.b1 = b1; t.bn1 = bn1_stats; t.w2 = w2; t.b2 = b2; t.y2 = y2;
    t.sums2 = training ? sums + 16 : nullptr; t.N = N; t.L0 = L0; t.L1 = d.L1; t.L2 = d.L2;
    int rc = trunk_conv2_tc_launch(t, st);
    if (rc) return rc;
  } else {
    trunk_conv2_fwd_kernel<<<dim3(wave_aware_ctas(trunk_conv2_fwd_kernel, 0, (d.L2 + TL - 1) / TL, N, 4), N), 256, 0, st>>>(
        x, d, w1, b1, bn1_stats, w2, b2, y2, training ? sums + 16 : nullptr);
    STEP_LAUNCH_CHECK("trunk_conv2_fwd_kernel");
  }
  if (training) {
    trunk_bn_finalize_kernel<<<1, 32, 0, st>>>(sums + 16, (double)N * d.L2, C2, g2, be2, eps, bn2_stats);
    STEP_LAUNCH_CHECK("trunk_bn_finalize_kernel");
  }
  const long long total = (long long)N * C2 * d.L2;
  trunk_bn_apply_kernel<<<dim3((d.L2 + 1023) / 1024, N * C2), 256, 0, st>>>(y2, C2, d.L2, bn2_stats, y2n);
  return check_launch("trunk_bn_apply_kernel");
}

extern "C" int step_dgl_conv_bwd(const float *dy2n, const float *x, int N, int L0, const float *w1, const float *b1,
                                 const float *g1, const float *w2, const float *g2, float eps, const float *bn1_stats,
                                 const float *bn2_stats, const float *y2, float *dy1n_scratch, float *dw1, float *db1,
                                 float *dg1, float *dbe1, float *dw2, float *db2, float *dg2, float *dbe2, void *scratch,
                                 void *stream) {
  STEP_REQUIRE(dy2n && x && w1 && b1 && g1 && w2 && g2 && bn1_stats && bn2_stats && y2 && dy1n_scratch && dw1 && db1 && dg1 &&
                   dbe1 && dw2 && db2 && dg2 && dbe2 && scratch,
               "dgl_conv_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const TrunkDims d{N, L0, L0 - HALO, L0 - 2 * HALO};
  double *sums = reinterpret_cast<double *>(scratch);          // [0,16): bn1 bwd, [16,48): bn2 bwd
  float *coef1 = reinterpret_cast<float *>(sums + 48);         // [5][8]
  float *coef2 = coef1 + 5 * C1;                               // [5][16]
  cudaError_t e = cudaMemsetAsync(sums, 0, 48 * sizeof(double), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dw1, 0, C1 * TK * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(db1, 0, C1 * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dw2, 0, C2 * C1 * TK * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(db2, 0, C2 * sizeof(float), st);
  if (e != cudaSuccess) return fail_msg((int)e, cudaGetErrorString(e));
  trunk_bn_bwd_stats_kernel<<<dim3(C2, 64), 256, 0, st>>>(dy2n, y2, N, C2, d.L2, bn2_stats, eps, sums + 16);
  STEP_LAUNCH_CHECK("trunk_bn_bwd_stats_kernel");
  trunk_bn_bwd_finalize_kernel<<<1, 32, 0, st>>>(sums + 16, (double)N * d.L2, C2, g2, bn2_stats, eps, coef2, dg2, dbe2);
  STEP_LAUNCH_CHECK("trunk_bn_bwd_finalize_kernel");
  int rc;
  if (trunk_use_tc()) {
    TcConv2BwdArgs t{};
    t.x = x; t.w1 = w1; t.b1 = b1; t.bn1 = bn1_stats; t.eps = eps; t.w2 = w2; t.dy2n = dy2n; t.y2 = y2; t.coef2 = coef2;
    t.dy1n = dy1n_scratch; t.dw2 = dw2; t.db2 = db2; t.sums1 = sums; t.N = N; t.L0 = L0; t.L1 = d.L1; t.L2 = d.L2;
    if ((rc = trunk_conv2_tc_bwd_launch(t, st))) return rc;
  } else {
    if ((rc = allow_smem(trunk_conv2_bwd_kernel, conv2_bwd_smem()))) return rc;
    const int gx = wave_aware_ctas(trunk_conv2_bwd_kernel, conv2_bwd_smem(), (d.L1 + TLB - 1) / TLB, N, 8);
    trunk_conv2_bwd_kernel<<<dim3(gx, N), 256, conv2_bwd_smem(), st>>>(x, d, w1, b1, bn1_stats, eps, w2, dy2n, y2,
                                                                      coef2, dy1n_scratch, dw2, db2, sums);
    STEP_LAUNCH_CHECK("trunk_conv2_bwd_kernel");
  }
  trunk_bn_bwd_finalize_kernel<<<1, 32, 0, st>>>(sums, (double)N * d.L1, C1, g1, bn1_stats, eps, coef1, dg1, dbe1);
  STEP_LAUNCH_CHECK("trunk_bn_bwd_finalize_kernel");
  trunk_conv1_bwd_kernel<<<dim3(wave_aware_ctas(trunk_conv1_bwd_kernel, 0, (d.L1 + 1023) / 1024, N, 6), N), 256, 0, st>>>(
      x, d, w1, b1, dy1n_scratch, coef1, dw1, db1);
  return check_launch("trunk_conv1_bwd_kernel");
}
