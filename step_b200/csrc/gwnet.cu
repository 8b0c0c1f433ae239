// Graph WaveNet layer stack: gated dilated temporal conv + skip conv + diffusion graph convolution
// (3 supports x 2 hops) + residual + BatchNorm, forward and backward, fp32.
//
// Reference semantics (file:line relative to the reference repo):
//   step/step_arch/graphwavenet/model.py:169-213   layer body
//   step/step_arch/graphwavenet/model.py:35-48     gcn (concat of 7 node-mixed copies -> 1x1 conv -> dropout)
//   step/step_arch/graphwavenet/model.py:10-16     nconv: out[w] = sum_v A[v,w] x[v]
//
// Design: activations are [B, T, N, 32] (channels innermost = one 128 B line per
// (b,t,n)); one CTA owns one (sample, time-step) column, i.e. all N nodes x 32 channels, so both
// diffusion hops, the 1x1 mixing, the residual and the BatchNorm partial sums happen in a single
// launch per layer.  The 224-channel concat of the reference is never formed: because node mixing
// and channel mixing commute, h = W0 u + sum_s P_s^T ( W_s1 u + P_s^T ( W_s2 u ) )  (Horner form),
// which needs the same 6 node-mixes per layer but only 32-channel operands.  BatchNorm of layer i is
// applied on load by layer i+1 (scale/shift from the batch statistics), the 256-channel skip conv is
// evaluated only at the last time step (the only column that survives `skip[..., -T:]`).
#include <stdlib.h>

#include "common.cuh"
#include "tc_mix.cuh"

namespace stepk {

constexpr int GC = 32;        // residual / dilation channels
constexpr int GSKIP = 256;    // skip channels
constexpr int GW_THREADS = 256;
constexpr int GW_MAX_LAYERS = 8;

struct GwPlan {
  int B, N, L;
  int Tin[GW_MAX_LAYERS], Tout[GW_MAX_LAYERS], dil[GW_MAX_LAYERS];
  size_t col;                       // floats per (b,t) column = N*32
  size_t off_sums_fwd, off_sums_bwd;  // doubles, [L][2][32] each (offsets in floats)
  size_t off_coef;                  // [L][4][32] floats: BN-backward coefficients
  size_t off_z[GW_MAX_LAYERS], off_f[GW_MAX_LAYERS], off_g[GW_MAX_LAYERS], off_q[GW_MAX_LAYERS][3];
  size_t off_a[GW_MAX_LAYERS][3];   // a_s = W_s2 u, kept for dP and the weight gradients (tensor-core mix path)
  size_t off_U, off_M, off_H;       // forward scratch, [B,12,N,32]
  size_t off_DH, off_DZC, off_DQ[3], off_A[3], off_DA, off_DU, off_DPF, off_DPG, off_DR[2];
  size_t off_M3[3], off_O3[3], off_DA3[3];   // tensor-core mix path: per-support mix outputs
  size_t off_img[3];                          // bf16 split images of P1, P2 (per sample) and P3
  size_t off_dus;                             // [L][B,N,32] skip-path gradient wrt u at each layer's last time step
  size_t total;
};

static GwPlan make_plan(int B, int N, int L) {
  GwPlan p{};
  p.B = B; p.N = N; p.L = L;
  p.col = (size_t)N * GC;
  int T = 13;
  for (int i = 0; i < L; ++i) {
    p.dil[i] = (i % 2 == 0) ? 1 : 2;
    p.Tin[i] = T;
    T -= p.dil[i];
    p.Tout[i] = T;
  }
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  p.off_sums_fwd = take((size_t)L * 2 * 32 * 2);
  p.off_sums_bwd = take((size_t)L * 2 * 32 * 2);
  p.off_coef = take((size_t)L * 4 * 32);
  for (int i = 0; i < L; ++i) {
    const size_t n = (size_t)B * p.Tout[i] * p.col;
    p.off_z[i] = take(n); p.off_f[i] = take(n); p.off_g[i] = take(n);
    for (int s = 0; s < 3; ++s) p.off_q[i][s] = take(n);
    for (int s = 0; s < 3; ++s) p.off_a[i][s] = take(n);
  }
  const size_t big = (size_t)B * 12 * p.col;
  p.off_U = take(big); p.off_M = take(big); p.off_H = take(big);
  p.off_DH = take(big); p.off_DZC = take(big);
  for (int s = 0; s < 3; ++s) { p.off_DQ[s] = take(big); p.off_A[s] = take(big); }
  p.off_DA = take(big); p.off_DU = take(big); p.off_DPF = take(big); p.off_DPG = take(big);
  p.off_DR[0] = take((size_t)B * 13 * p.col);
  p.off_DR[1] = take((size_t)B * 13 * p.col);
  for (int s = 0; s < 3; ++s) { p.off_M3[s] = take(big); p.off_O3[s] = take(big); p.off_DA3[s] = take(big); }
  const size_t img_floats = (mix_images_bytes(N) + 3) / 4;
  p.off_img[0] = take(img_floats * B);
  p.off_img[1] = take(img_floats * B);
  p.off_img[2] = take(img_floats);
  p.off_dus = take((size_t)L * B * p.col);
  p.total = o;
  return p;
}

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_row(const float *p, float (&r)[GC]) {
#pragma unroll
  for (int c = 0; c < GC; c += 4) {
    const float4 v = *reinterpret_cast<const float4 *>(p + c);
    r[c] = v.x; r[c + 1] = v.y; r[c + 2] = v.z; r[c + 3] = v.w;
  }
}
__device__ __forceinline__ void store_row(float *p, const float (&r)[GC]) {
#pragma unroll
  for (int c = 0; c < GC; c += 4) *reinterpret_cast<float4 *>(p + c) = make_float4(r[c], r[c + 1], r[c + 2], r[c + 3]);
}
// For cg = 0..7: f(cg, float4{ sum_ci W[4cg+j][ci] x[ci] }_j)   (W: smem [32][32] row-major [co][ci]).
// The output chunk loop is deliberately not unrolled: outputs go straight to memory through `f`, which
// keeps the code small and every register index static.
template <typename F>
__device__ __forceinline__ void matvec_chunks(const float *W, const float (&x)[GC], F f) {
#pragma unroll 1
  for (int cg = 0; cg < 8; ++cg) {
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 *w = reinterpret_cast<const float4 *>(W + (4 * cg + j) * GC);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // packed FFMA2: even / odd input channels, two chains
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 ww = w[c4];
        ffma2(a0, a1, ww.x, ww.y, x[4 * c4], x[4 * c4 + 1]);
        ffma2(a2, a3, ww.z, ww.w, x[4 * c4 + 2], x[4 * c4 + 3]);
      }
      o[j] = (a0 + a1) + (a2 + a3);
    }
    f(cg, make_float4(o[0], o[1], o[2], o[3]));
  }
}
// out[ci] += sum_co W[co][ci] xp[co]   (xp: a row in shared or global memory)
__device__ __forceinline__ void matvec_t_acc(const float *W, const float *xp, float (&out)[GC]) {
#pragma unroll 4
  for (int co = 0; co < GC; ++co) {
    const float4 *w = reinterpret_cast<const float4 *>(W + co * GC);
    const float xv = xp[co];
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 ww = w[c4];
      ffma2(out[4 * c4], out[4 * c4 + 1], ww.x, ww.y, xv, xv);
      ffma2(out[4 * c4 + 2], out[4 * c4 + 3], ww.z, ww.w, xv, xv);
    }
  }
}
// same with the input row already in registers (loaded with 8 independent 16-byte loads, so that one memory
// round trip is exposed per row instead of one per 4 elements); fully unrolled: every register index is static.
__device__ __forceinline__ void matvec_t_reg(const float *W, const float (&x)[GC], float (&out)[GC]) {
#pragma unroll
  for (int co = 0; co < GC; ++co) {
    const float4 *w = reinterpret_cast<const float4 *>(W + co * GC);
    const float xv = x[co];
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 ww = w[c4];
      ffma2(out[4 * c4], out[4 * c4 + 1], ww.x, ww.y, xv, xv);
      ffma2(out[4 * c4 + 2], out[4 * c4 + 3], ww.z, ww.w, xv, xv);
    }
  }
}
// Column sums of a [32 lanes][32] register tile: lane L returns sum over the warp's lanes of v[L] (recursive halving:
// 31 shuffles instead of 32 full butterflies).  v is destroyed.  All 32 lanes must call.
__device__ __forceinline__ float warp_colsum32(float (&v)[GC]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// out[w][c] = sum_k Pm[k][w] * Y[k][c]; Pm [N][N] in global (reduction index = row), Y [N][32] in smem.
// Thread <-> output node w (all 32 channels in registers); the Pm reads are coalesced across the warp,
// the Y reads are warp-wide broadcasts.
template <typename Sink>
__device__ __forceinline__ void mix_nodes(const float *__restrict__ Pm, const float *Y, int N, Sink sink) {
  for (int w0 = 0; w0 < N; w0 += GW_THREADS) {
    if (w0 + (int)(threadIdx.x & ~31u) >= N) continue;  // whole warp past the end
    const int w = w0 + threadIdx.x;
    const bool ok = w < N;
    const float *pc = Pm + (ok ? w : N - 1);
    float acc[GC];
#pragma unroll
    for (int c = 0; c < GC; ++c) acc[c] = 0.f;
    int k = 0;
    for (; k + 8 <= N; k += 8) {
      float p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] = pc[(size_t)(k + j) * N];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 *y = reinterpret_cast<const float4 *>(Y + (size_t)(k + j) * GC);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 yy = y[c4];
          acc[4 * c4] = fmaf(p[j], yy.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(p[j], yy.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(p[j], yy.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(p[j], yy.w, acc[4 * c4 + 3]);
        }
      }
    }
    for (; k < N; ++k) {
      const float pv = pc[(size_t)k * N];
      const float4 *y = reinterpret_cast<const float4 *>(Y + (size_t)k * GC);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 yy = y[c4];
        acc[4 * c4] = fmaf(pv, yy.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(pv, yy.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(pv, yy.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(pv, yy.w, acc[4 * c4 + 3]);
      }
    }
    if (ok) sink(w, acc);
  }
}

// global [N][32] -> smem Y (coalesced float4 copy)
__device__ __forceinline__ void copy_to_smem(float *Y, const float *src, int N) {
  const float4 *s = reinterpret_cast<const float4 *>(src);
  float4 *d = reinterpret_cast<float4 *>(Y);
  for (int i = threadIdx.x; i < N * (GC / 4); i += GW_THREADS) d[i] = s[i];
}

// per-channel sums over the N rows of the smem tile Y (and of Y*Y), added to two double accumulators.
// red: smem scratch of >= 512 floats.
__device__ __forceinline__ void column_sums_to(const float *Y, int N, float *red, double *sum_dst, double *sq_dst) {
  const int c = threadIdx.x & 31, grp = threadIdx.x >> 5;
  float s = 0.f, q = 0.f;
  for (int n = grp; n < N; n += GW_THREADS / 32) {
    const float v = Y[(size_t)n * GC + c];
    s += v;
    q = fmaf(v, v, q);
  }
  red[grp * 32 + c] = s;
  red[256 + grp * 32 + c] = q;
  __syncthreads();
  if (threadIdx.x < 32) {
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < GW_THREADS / 32; ++g2) { ts += red[g2 * 32 + c]; tq += red[256 + g2 * 32 + c]; }
    if (sum_dst) atomicAdd(sum_dst + c, (double)ts);
    if (sq_dst) atomicAdd(sq_dst + c, (double)tq);
  }
  __syncthreads();
}

// dst[co*ldd + ci] += sum_n X[n][co] * U[n][ci]   (X, U: [N][32] in global, CTA-local data)
// tiles: smem scratch of 2*64*33 floats.
__device__ __forceinline__ void outer_acc(const float *X, const float *U, int N, float *tiles, float *dst, int ldd) {
  float *T1 = tiles, *T2 = tiles + 64 * 32;          // [64 nodes][32] each, 16-byte aligned rows
  const int co = threadIdx.x >> 3, ci4 = (threadIdx.x & 7) * 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int n0 = 0; n0 < N; n0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * (GC / 4); i += GW_THREADS) {
      const int n = i >> 3, c4 = (i & 7) * 4;
      const bool ok = (n0 + n) < N;
      st4(T1 + n * GC + c4, ok ? ld4(X + (size_t)(n0 + n) * GC + c4) : make_float4(0, 0, 0, 0));
      st4(T2 + n * GC + c4, ok ? ld4(U + (size_t)(n0 + n) * GC + c4) : make_float4(0, 0, 0, 0));
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < 64; ++n) {
      const float x = T1[n * GC + co];                 // 4 distinct words per warp: broadcast
      const float4 u = ld4(T2 + n * GC + ci4);         // 8 distinct 16-byte words per warp: conflict-free
      ffma2(a0, a1, u.x, u.y, x, x); ffma2(a2, a3, u.z, u.w, x, x);
    }
  }
  float *d = dst + (size_t)co * ldd + ci4;
  atomicAdd(d, a0); atomicAdd(d + 1, a1); atomicAdd(d + 2, a2); atomicAdd(d + 3, a3);
  __syncthreads();
}

// dst[(blk)*32 + co*ldd + ci] += sum_n X_blk[n][co] * U[n][ci] for NB blocks sharing the U operand: one staging of
// each 64-node U tile serves all blocks (tiles: (NB + 1) * 64 * 32 floats).
template <int NB>
__device__ __forceinline__ void outer_acc_multi(const float *const (&X)[NB], const float *U, int N, float *tiles,
                                                float *const (&dst)[NB], int ldd) {
  float *TU = tiles;
  const int co = threadIdx.x >> 3, ci4 = (threadIdx.x & 7) * 4;
  float acc[NB][4];
#pragma unroll
  for (int k = 0; k < NB; ++k) { acc[k][0] = 0.f; acc[k][1] = 0.f; acc[k][2] = 0.f; acc[k][3] = 0.f; }
  for (int n0 = 0; n0 < N; n0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * (GC / 4); i += GW_THREADS) {
      const int n = i >> 3, c4 = (i & 7) * 4;
      const bool ok = (n0 + n) < N;
      st4(TU + n * GC + c4, ok ? ld4(U + (size_t)(n0 + n) * GC + c4) : make_float4(0, 0, 0, 0));
#pragma unroll
      for (int k = 0; k < NB; ++k)
        st4(tiles + (k + 1) * 64 * GC + n * GC + c4, ok ? ld4(X[k] + (size_t)(n0 + n) * GC + c4) : make_float4(0, 0, 0, 0));
    }
    __syncthreads();
#pragma unroll 4
    for (int n = 0; n < 64; ++n) {
      const float4 u = ld4(TU + n * GC + ci4);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const float x = tiles[(k + 1) * 64 * GC + n * GC + co];
        ffma2(acc[k][0], acc[k][1], u.x, u.y, x, x); ffma2(acc[k][2], acc[k][3], u.z, u.w, x, x);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    float *d = dst[k] + (size_t)co * ldd + ci4;
    atomicAdd(d, acc[k][0]); atomicAdd(d + 1, acc[k][1]); atomicAdd(d + 2, acc[k][2]); atomicAdd(d + 3, acc[k][3]);
  }
  __syncthreads();
}

// Register-tiled weight-gradient outer products: dst_k[((co)*ldd + ci)*es] += sum_n X_k[n][co] * U[n][ci] for NB blocks
// sharing U.  Thread = (node quarter q, 4 output rows co4.., 4 columns ci4..): per node one 16-byte load of U and one per
// block of X feed 8 packed FMAs per block (the one-row form spent 2 shared-memory loads per 2).  The four node quarters
// and the CTAs combine through the atomics.  tiles: (NB + 1) * 64 * 32 floats.  es: element stride of dst (2 for the
// [co][ci][tap] conv weights).
template <int NB>
__device__ __forceinline__ void outer_acc_tiled(const float *const (&X)[NB], const float *U, int N, float *tiles,
                                                float *const (&dst)[NB], int ldd, int es) {
  float *TU = tiles;
  const int q = threadIdx.x >> 6, co4 = ((threadIdx.x >> 3) & 7) * 4, ci4 = (threadIdx.x & 7) * 4;
  float acc[NB][4][4];
#pragma unroll
  for (int k = 0; k < NB; ++k)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[k][r][c] = 0.f;
  for (int n0 = 0; n0 < N; n0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * (GC / 4); i += GW_THREADS) {
      const int n = i >> 3, c4 = (i & 7) * 4;
      const bool ok = (n0 + n) < N;
      st4(TU + n * GC + c4, ok ? ld4(U + (size_t)(n0 + n) * GC + c4) : make_float4(0, 0, 0, 0));
#pragma unroll
      for (int k = 0; k < NB; ++k)
        st4(tiles + (k + 1) * 64 * GC + n * GC + c4, ok ? ld4(X[k] + (size_t)(n0 + n) * GC + c4) : make_float4(0, 0, 0, 0));
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      const int n = 4 * j + q;
      const float4 u = ld4(TU + n * GC + ci4);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const float4 x = ld4(tiles + (k + 1) * 64 * GC + n * GC + co4);
        ffma2(acc[k][0][0], acc[k][0][1], u.x, u.y, x.x, x.x); ffma2(acc[k][0][2], acc[k][0][3], u.z, u.w, x.x, x.x);
        ffma2(acc[k][1][0], acc[k][1][1], u.x, u.y, x.y, x.y); ffma2(acc[k][1][2], acc[k][1][3], u.z, u.w, x.y, x.y);
        ffma2(acc[k][2][0], acc[k][2][1], u.x, u.y, x.z, x.z); ffma2(acc[k][2][2], acc[k][2][3], u.z, u.w, x.z, x.z);
        ffma2(acc[k][3][0], acc[k][3][1], u.x, u.y, x.w, x.w); ffma2(acc[k][3][2], acc[k][3][3], u.z, u.w, x.w, x.w);
      }
    }
  }
  // combine the four node quarters through shared memory (the staging tiles are free now): one atomic per entry and CTA
  __syncthreads();
  float *R = tiles;                                    // [4 quarters][32 x 32]
#pragma unroll
  for (int k = 0; k < NB; ++k) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      st4(R + q * 1024 + (co4 + r) * GC + ci4, make_float4(acc[k][r][0], acc[k][r][1], acc[k][r][2], acc[k][r][3]));
    __syncthreads();
    const int o = threadIdx.x * 4;
    const float4 v = add4(add4(ld4(R + o), ld4(R + 1024 + o)), add4(ld4(R + 2048 + o), ld4(R + 3072 + o)));
    float *d = dst[k] + ((size_t)(o >> 5) * ldd + (o & 31)) * es;
    atomicAdd(d, v.x); atomicAdd(d + es, v.y); atomicAdd(d + 2 * es, v.z); atomicAdd(d + 3 * es, v.w);
    __syncthreads();
  }
}

__device__ __forceinline__ void dropout_row(float (&h)[GC], uint64_t elem0, uint32_t thr, float scale, uint64_t key) {
  // elem0: flat index of channel 0 of this row (multiple of 32)
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const uint4 r = philox4x32((elem0 >> 2) + c4, key);
    h[4 * c4] = (r.x >= thr) ? h[4 * c4] * scale : 0.f;
    h[4 * c4 + 1] = (r.y >= thr) ? h[4 * c4 + 1] * scale : 0.f;
    h[4 * c4 + 2] = (r.z >= thr) ? h[4 * c4 + 2] * scale : 0.f;
    h[4 * c4 + 3] = (r.w >= thr) ? h[4 * c4 + 3] * scale : 0.f;
  }
}

__device__ __forceinline__ float4 dropout4(float4 h, uint64_t elem0, int c4, uint32_t thr, float scale, uint64_t key) {
  const uint4 r = philox4x32((elem0 >> 2) + c4, key);
  return make_float4((r.x >= thr) ? h.x * scale : 0.f, (r.y >= thr) ? h.y * scale : 0.f,
                     (r.z >= thr) ? h.z * scale : 0.f, (r.w >= thr) ? h.w * scale : 0.f);
}

// The dropout site of the gcn output (graphwavenet/model.py:47) on a stand-alone [rows, 32] buffer: exactly the mask the
// layer kernels draw for the element range [0, rows*32) of layer `layer` (test hook for the distribution checks).
__global__ void gw_dropout_probe_kernel(const float *__restrict__ x, long long rows, uint32_t thr, float scale, uint64_t key,
                                        float *__restrict__ y) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float h[GC];
  load_row(x + r * GC, h);
  dropout_row(h, (uint64_t)(r * GC), thr, scale, key);
  store_row(y + r * GC, h);
}

// ---------------------------------------------------------------------------
// forward layer kernel: grid (T_out, B)
// ---------------------------------------------------------------------------
struct GwFwdArgs {
  const float *zin; float *zout;
  int Tin, Tout, dil, N, has_gcn, has_in_bn, collect_stats;
  const float *in_scale, *in_shift;      // [32] BN-on-load of the previous layer
  const float *P[3]; long long pstride[3];  // supports, per-sample stride (0 for the shared adaptive one)
  step_gw_layer_params w;
  float *f, *g, *q[3], *U, *M, *H;
  int stage;                              // 0: fused CUDA-core layer; 1: conv + skip + a_s; 2: q_s; 3: output (tensor-core mixes in between)
  float *Aout[3];                         // stage 1: a_s = W_s2 u
  const float *Min[3], *Oin[3];           // stage 2 / 3: tensor-core mix results
  double *sums;                           // [2][32]
  uint32_t drop_thr; float drop_scale; uint64_t key;
};

template <int STAGE>
__global__ void __launch_bounds__(GW_THREADS, 2) gw_layer_fwd_kernel(GwFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int N = a.N, tid = threadIdx.x, t = blockIdx.x, b = blockIdx.y;
  float *Y = smem;                          // [N][32]
  float *Wb = smem + (size_t)N * GC;        // 7 * 1024 floats
  const size_t col = (size_t)N * GC;
  const float *z0 = a.zin + ((size_t)b * a.Tin + t) * col;
  const float *z1 = a.zin + ((size_t)b * a.Tin + t + a.dil) * col;
  const size_t ocol = ((size_t)b * a.Tout + t) * col;
  float *U = a.U + ocol, *Mb = a.M + ocol, *Hb = a.H + ocol;

  if (STAGE <= 1) {
  for (int i = tid; i < 1024; i += GW_THREADS) {
    Wb[i] = a.w.filter_w[2 * i]; Wb[1024 + i] = a.w.filter_w[2 * i + 1];
    Wb[2048 + i] = a.w.gate_w[2 * i]; Wb[3072 + i] = a.w.gate_w[2 * i + 1];
  }
  __syncthreads();

  // ---- phase A: gated dilated conv, u = tanh(.) * sigmoid(.) ----
  for (int n = tid; n < N; n += GW_THREADS) {
    float r0[GC], r1[GC];
    load_row(z0 + (size_t)n * GC, r0);
    load_row(z1 + (size_t)n * GC, r1);
    if (a.has_in_bn) {
#pragma unroll
      for (int c = 0; c < GC; ++c) {
        const float sc = a.in_scale[c], sh = a.in_shift[c];
        r0[c] = fmaf(r0[c], sc, sh);
        r1[c] = fmaf(r1[c], sc, sh);
      }
    }
    float *fo = a.f + ocol + (size_t)n * GC, *go = a.g + ocol + (size_t)n * GC, *uo = U + (size_t)n * GC;
#pragma unroll 1
    for (int cg = 0; cg < GC; cg += 4) {
      float fv[4], gv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = cg + j;
        const float4 *wf0 = reinterpret_cast<const float4 *>(Wb + co * GC);
        const float4 *wf1 = reinterpret_cast<const float4 *>(Wb + 1024 + co * GC);
        const float4 *wg0 = reinterpret_cast<const float4 *>(Wb + 2048 + co * GC);
        const float4 *wg1 = reinterpret_cast<const float4 *>(Wb + 3072 + co * GC);
        float f0 = a.w.filter_b[co], f1 = 0.f, f2 = 0.f, f3 = 0.f, g0 = a.w.gate_b[co], g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 A0 = wf0[c4], A1 = wf1[c4], G0 = wg0[c4], G1 = wg1[c4];
          ffma2(f0, f1, A0.x, A0.y, r0[4 * c4], r0[4 * c4 + 1]); ffma2(f2, f3, A0.z, A0.w, r0[4 * c4 + 2], r0[4 * c4 + 3]);
          ffma2(f0, f1, A1.x, A1.y, r1[4 * c4], r1[4 * c4 + 1]); ffma2(f2, f3, A1.z, A1.w, r1[4 * c4 + 2], r1[4 * c4 + 3]);
          ffma2(g0, g1, G0.x, G0.y, r0[4 * c4], r0[4 * c4 + 1]); ffma2(g2, g3, G0.z, G0.w, r0[4 * c4 + 2], r0[4 * c4 + 3]);
          ffma2(g0, g1, G1.x, G1.y, r1[4 * c4], r1[4 * c4 + 1]); ffma2(g2, g3, G1.z, G1.w, r1[4 * c4 + 2], r1[4 * c4 + 3]);
        }
        const float af = (f0 + f1) + (f2 + f3), ag = (g0 + g1) + (g2 + g3);
        fv[j] = tanhf(af);
        gv[j] = 1.0f / (1.0f + expf(-ag));
      }
      *reinterpret_cast<float4 *>(fo + cg) = make_float4(fv[0], fv[1], fv[2], fv[3]);
      *reinterpret_cast<float4 *>(go + cg) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      *reinterpret_cast<float4 *>(uo + cg) = make_float4(fv[0] * gv[0], fv[1] * gv[1], fv[2] * gv[2], fv[3] * gv[3]);
    }
  }
  __syncthreads();

  }  // stage <= 1
  if (!a.has_gcn) return;

  // ---- stage the 7 [32x32] blocks of the gcn 1x1 conv: Wb[k][co][ci] = mlp_w[co][k*32+ci] ----
  // STAGE 0 keeps Wb[k][co][ci]; the split stages stage the TRANSPOSED blocks Wb[k][ci][co] so that every channel mix
  // is the register-row form out[co] += W^T[ci][co] * u[ci] (matvec_t_reg: accumulators and addend rows in registers)
  // Only the blocks a stage uses are staged (stage 1: W_s2, stage 2: W_s1, stage 3: W_0); global reads stay coalesced
  // (ci fastest), the transposition happens on the shared-memory store.
  if (STAGE == 0) {
    for (int i = tid; i < 7 * 1024; i += GW_THREADS) {
      const int k = i >> 10, co = (i >> 5) & 31, ci = i & 31;
      Wb[i] = a.w.mlp_w[(size_t)co * 224 + k * 32 + ci];
    }
  } else {
    constexpr int NK = (STAGE == 3) ? 1 : 3;
    for (int i = tid; i < NK * 1024; i += GW_THREADS) {
      const int j = i >> 10, co = (i >> 5) & 31, ci = i & 31;
      const int k = (STAGE == 3) ? 0 : (STAGE == 1 ? 2 + 2 * j : 1 + 2 * j);
      Wb[k * 1024 + ci * 32 + co] = a.w.mlp_w[(size_t)co * 224 + k * 32 + ci];
    }
  }
  __syncthreads();

  if (STAGE == 1) {
    // a_s = W_s2 u for the three supports -> global; the node mixes run on the tensor cores (tc_mix_kernel)
    for (int n = tid; n < N; n += GW_THREADS) {
      float u[GC];
      load_row(U + (size_t)n * GC, u);
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        float o[GC];
#pragma unroll
        for (int c = 0; c < GC; ++c) o[c] = 0.f;
        matvec_t_reg(Wb + (2 + 2 * s) * 1024, u, o);
        store_row(a.Aout[s] + ocol + (size_t)n * GC, o);
      }
    }
    return;
  }
  if (STAGE == 2) {
    // q_s = W_s1 u + (P_s^T a_s)
    for (int n = tid; n < N; n += GW_THREADS) {
      const size_t ro = ocol + (size_t)n * GC;
      float u[GC], q[GC], nx[GC];
      load_row(a.Min[0] + ro, q);
      load_row(U + (size_t)n * GC, u);
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        if (s < 2) load_row(a.Min[s + 1] + ro, nx);     // next support's addend row, in flight under the product
        matvec_t_reg(Wb + (1 + 2 * s) * 1024, u, q);
        store_row(a.q[s] + ro, q);
#pragma unroll
        for (int c = 0; c < GC; ++c) q[c] = nx[c];
      }
    }
    return;
  }
  // ---- phase M: diffusion, Horner form per support ----
  for (int s = 0; s < 3 && STAGE == 0; ++s) {
    const float *Ps = a.P[s] + (size_t)b * a.pstride[s];
    const float *W1 = Wb + (1 + 2 * s) * 1024, *W2 = Wb + (2 + 2 * s) * 1024;
    // a = W_s2 u  -> Y
    for (int n = tid; n < N; n += GW_THREADS) {
      float u[GC];
      load_row(U + (size_t)n * GC, u);
      float *yr = Y + (size_t)n * GC;
      matvec_chunks(W2, u, [&](int cg, float4 v) { st4(yr + 4 * cg, v); });
    }
    __syncthreads();
    // m = P^T a -> M (global scratch)
    mix_nodes(Ps, Y, N, [&](int w, float (&acc)[GC]) { store_row(Mb + (size_t)w * GC, acc); });
    __syncthreads();
    // q = W_s1 u + m -> Y and stash
    float *qs = a.q[s] + ocol;
    for (int n = tid; n < N; n += GW_THREADS) {
      float u[GC];
      load_row(U + (size_t)n * GC, u);
      float *yr = Y + (size_t)n * GC, *qr = qs + (size_t)n * GC;
      const float *mr = Mb + (size_t)n * GC;
      matvec_chunks(W1, u, [&](int cg, float4 v) {
        v = add4(v, ld4(mr + 4 * cg));
        st4(yr + 4 * cg, v);
        st4(qr + 4 * cg, v);
      });
    }
    __syncthreads();
    // o = P^T q, accumulated into H
    if (s == 0) {
      mix_nodes(Ps, Y, N, [&](int w, float (&acc)[GC]) { store_row(Hb + (size_t)w * GC, acc); });
    } else {
      mix_nodes(Ps, Y, N, [&](int w, float (&acc)[GC]) {
        float h[GC];
        load_row(Hb + (size_t)w * GC, h);
#pragma unroll
        for (int c = 0; c < GC; ++c) h[c] += acc[c];
        store_row(Hb + (size_t)w * GC, h);
      });
    }
    __syncthreads();
  }

  // ---- phase F: h = bm + W0 u + H; dropout; residual; pre-BN output + BN partial sums ----
  if (STAGE == 3) {
    for (int n = tid; n < N; n += GW_THREADS) {
      const size_t ro = ocol + (size_t)n * GC;
      float h[GC], u[GC], r[GC];
      load_row(a.Oin[0] + ro, h);
      load_row(a.Oin[1] + ro, u);
      load_row(a.Oin[2] + ro, r);
#pragma unroll
      for (int c = 0; c < GC; ++c) h[c] = (h[c] + u[c]) + (r[c] + a.w.mlp_b[c]);
      load_row(U + (size_t)n * GC, u);
      load_row(z1 + (size_t)n * GC, r);                 // residual row, in flight under the product
      matvec_t_reg(Wb, u, h);
      if (a.drop_thr) dropout_row(h, (uint64_t)ro, a.drop_thr, a.drop_scale, a.key);
#pragma unroll
      for (int c = 0; c < GC; ++c) {
        const float rv = a.has_in_bn ? fmaf(r[c], a.in_scale[c], a.in_shift[c]) : r[c];
        h[c] += rv;
      }
      store_row(a.zout + ro, h);
      store_row(Y + (size_t)n * GC, h);
    }
  } else
  for (int n = tid; n < N; n += GW_THREADS) {
    float u[GC];
    load_row(U + (size_t)n * GC, u);
    const float *hr = Hb + (size_t)n * GC, *zr = z1 + (size_t)n * GC;
    const float *o0 = a.Oin[0] + ocol + (size_t)n * GC, *o1 = a.Oin[1] + ocol + (size_t)n * GC, *o2 = a.Oin[2] + ocol + (size_t)n * GC;
    float *yr = Y + (size_t)n * GC, *zo = a.zout + ocol + (size_t)n * GC;
    const uint64_t elem0 = (uint64_t)(ocol + (size_t)n * GC);
    matvec_chunks(Wb, u, [&](int cg, float4 v) {
      const float4 hsum = (STAGE == 0) ? ld4(hr + 4 * cg) : add4(add4(ld4(o0 + 4 * cg), ld4(o1 + 4 * cg)), ld4(o2 + 4 * cg));
      v = add4(add4(v, hsum), ld4(a.w.mlp_b + 4 * cg));
      if (a.drop_thr) v = dropout4(v, elem0, cg, a.drop_thr, a.drop_scale, a.key);
      float4 r = ld4(zr + 4 * cg);
      if (a.has_in_bn) {
        const float4 sc = ld4(a.in_scale + 4 * cg), sh = ld4(a.in_shift + 4 * cg);
        r = make_float4(fmaf(r.x, sc.x, sh.x), fmaf(r.y, sc.y, sh.y), fmaf(r.z, sc.z, sh.z), fmaf(r.w, sc.w, sh.w));
      }
      v = add4(v, r);
      st4(zo + 4 * cg, v);
      st4(yr + 4 * cg, v);
    });
  }
  __syncthreads();
  if (a.collect_stats) column_sums_to(Y, N, Wb, a.sums, a.sums + 32);
}

// ---------------------------------------------------------------------------
// ONE launch per Graph WaveNet layer (forward): the whole layer body of graphwavenet/model.py:169-213 for a
// (sample, pair of time steps) column block - gated dilated conv, the three supports' two diffusion hops and the 1x1
// mixing (Horner form, see the header), dropout, residual, BatchNorm partial sums - with every product on tcgen05 INSIDE
// the kernel (template TCM = true): nothing but the layer input, the stash rows the backward needs and the layer output
// touch HBM, and the CUDA cores only do the elementwise work (BN on load, tanh / sigmoid, hi/lo splits, dropout, residual).
//   workers (8 warps, thread = node):  z_t, z_{t+dil} (BN applied) -> bf16 hi/lo A images (in the idle support-slice ring)
//   MMA lane:                          [f | g] pre-activations = sum_tap Z_tap W_tap            (K = 64, N = 64)
//   workers:                           u = tanh(.) * sigmoid(.) -> f, g stash; u as bf16 hi/lo A images
//   MMA lane:                          a_s = U W_s2^T (scratch tile)                              (K = 32, N = 32 per time step)
//   workers:                           a_s -> stash + bf16 hi/lo B image
//   MMA lane:                          q_s = P_s^T a_s + U W_s1^T  (A = P_s^T images streamed by the TMA lane, K slices of 32 nodes)
//   workers:                           q_s (TMEM -> registers) -> stash + B image (overwrites a_s)
//   MMA lane:                          H  += P_s^T q_s (accumulated over the three supports), finally H += U W_0^T
//   workers:                           h = H + b -> dropout -> + BN(residual) -> z, BN partial sums
// All MMAs are split-bf16 (hi*hi + hi*lo + lo*hi).  TCM = false keeps the conv and the channel mixes on CUDA cores
// (numerically identical to the split path gw_layer_fwd_kernel<1,2,3> + 2 x tc_mix_kernel; STEP_B200_GW_FUSED=1).
// N <= 256 (two 128-row tiles).  grid = (ceil(T_out / 2), B), 1 CTA/SM.
// ---------------------------------------------------------------------------
constexpr int GWF_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 workers
constexpr int GWF_NT = 2;             // time steps per CTA

struct GwFusedArgs {
  const float *zin; float *zout;
  int Tin, Tout, dil, N, has_in_bn, collect_stats;
  const float *in_scale, *in_shift;
  step_gw_layer_params w;
  float *f, *g, *q[3], *a[3];
  const uint8_t *img[3];
  long long img_bstride[3];
  MixGeom geom;
  double *sums;
  uint32_t drop_thr; float drop_scale; uint64_t key;
};

// TCM: the seven [32 x 32] channel mixes of the gcn on tcgen05 as well (u as bf16 hi/lo A-operand images, the weight blocks
// as B images; W_s1 u accumulates into hop A's tile and W_0 u into hop B's, a_s = W_s2 u gets a scratch tile of its own)
static size_t gwf_smem_bytes(const MixGeom &g, bool tcm) {
  const size_t u_bytes = tcm ? (size_t)GWF_NT * 2 * 4 * 128 * g.MT * 16 : (size_t)GWF_NT * g.N * GC * 4;
  return 2 * (size_t)(2 * 4 * 128 * g.MT * 16) + 2 * (size_t)GWF_NT * 4 * g.Kpad * 16 + u_bytes + 7 * 1024 * 4 +
         2 * 8 * 64 * 4 + 16 * 8 + 16;
}

template <bool TCM>
__global__ void __launch_bounds__(GWF_THREADS, 1) gw_fused_fwd_kernel(GwFusedArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t gwf_smem[];
  const MixGeom g = a.geom;
  const int N = a.N, b = blockIdx.y, t0 = blockIdx.x * GWF_NT, nt = min(GWF_NT, a.Tout - t0);
  const uint32_t rows = 128u * g.MT;
  const uint32_t half_bytes = 4 * rows * 16, stage_bytes = 2 * half_bytes;
  const uint32_t bimg = (uint32_t)GWF_NT * 4 * g.Kpad * 16;
  uint8_t *sA = gwf_smem;
  uint8_t *sBh = sA + 2 * stage_bytes, *sBl = sBh + bimg;
  float *sU = reinterpret_cast<float *>(sBl + bimg);                 // !TCM: [nt][N][32] fp32
  uint8_t *sUimg = sBl + bimg;                                        // TCM: [t][hi, lo][4 chunks][rows][16 B]
  const uint32_t uimg_bytes = 4 * rows * 16;                          // one (t, hi|lo) image
  float *sW = reinterpret_cast<float *>(sBl + bimg + (TCM ? (size_t)GWF_NT * 2 * uimg_bytes : (size_t)GWF_NT * N * GC * 4));   // 7 x 1024
  uint8_t *sWimg = reinterpret_cast<uint8_t *>(sW);                   // TCM (after the conv): [7 blocks][hi, lo][4 chunks][32 rows][16 B]
  float *sRed = sW + 7 * 1024;                                        // [2][8 warps][64]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sRed + 2 * 8 * 64);
  uint64_t *full = bars, *empty = bars + 2, *b_ready = bars + 4, *acc1_full = bars + 5, *hopb_done = bars + 6;
  uint64_t *u_ready = bars + 7, *a_full = bars + 8, *a_empty = bars + 9, *z_ready = bars + 10, *conv_full = bars + 11;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t col = (size_t)N * GC;
  const int nslices = g.Kpad / 32;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(b_ready, 8); mbar_init(acc1_full, 1); mbar_init(hopb_done, 1);
    mbar_init(u_ready, 8); mbar_init(a_full, 1); mbar_init(a_empty, 8); mbar_init(z_ready, 8); mbar_init(conv_full, 1);
    fence_barrier_init();
  }
  // B-image rows of the padding nodes [N, Kpad) stay zero for the whole kernel
  for (uint32_t i = threadIdx.x; i < 2 * bimg / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sBh)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, TCM ? 512 : 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA lane: P_s^T image slices, 6 passes (3 supports x 2 hops) ------------------------------
    if (lane == 0) {
      uint32_t n = 0;
      if (TCM) {                                                // the ring holds the conv's input images until then
        for (int t = 0; t < nt; ++t) mbar_wait(conv_full, t & 1);
      }
      for (int s = 0; s < 3; ++s) {
        const uint8_t *img_hi = a.img[s] + (size_t)b * a.img_bstride[s], *img_lo = img_hi + g.img_bytes;
        for (int hop = 0; hop < 2; ++hop) {
          for (int i = 0; i < nslices; ++i, ++n) {
            const uint32_t st = n & 1;
            mbar_wait(&empty[st], ((n >> 1) & 1) ^ 1);
            mbar_expect_tx(&full[st], stage_bytes);
            for (int c = 0; c < 4; ++c) {
              const size_t off = ((size_t)(i * 4 + c) * g.Mpad) * 16;
              tma_bulk_g2s(sA + st * stage_bytes + c * rows * 16, img_hi + off, rows * 16, &full[st]);
              tma_bulk_g2s(sA + st * stage_bytes + half_bytes + c * rows * 16, img_lo + off, rows * 16, &full[st]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA lane ------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, nt * 32, 0, 1);
      const uint32_t bh = smem_u32(sBh), bl = smem_u32(sBl);
      const uint32_t b_sbo = (uint32_t)g.Kpad * 16, a_lbo = rows * 16;
      uint32_t n = 0, ph = 0;
      // channel mix of block k on the tensor core: D[128 nodes of tile m, 32] (+)= U_t[128, 32] W_k^T for every (t, m)
      const uint32_t idesc_mix = umma_idesc_bf16(128, 32, 0, 0);
      const uint32_t uimg = smem_u32(sUimg), wimg = smem_u32(sWimg);
      auto mix = [&](int k, uint32_t dcol, uint32_t accumulate) {
        for (int t = 0; t < nt; ++t)
          for (int m = 0; m < g.MT; ++m) {
            const uint32_t uh = uimg + (uint32_t)(t * 2) * uimg_bytes + m * 2048, ul = uh + uimg_bytes;
            const uint32_t whi = wimg + (uint32_t)(k * 2) * 2048, wlo = whi + 2048;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t dah = umma_desc(uh + kk * 2 * a_lbo, a_lbo, 128), dal = umma_desc(ul + kk * 2 * a_lbo, a_lbo, 128);
              const uint64_t dbh = umma_desc(whi + kk * 2 * 512, 512, 128), dbl = umma_desc(wlo + kk * 2 * 512, 512, 128);
              const uint32_t d = tmem + dcol + m * 64 + t * 32;
              umma_bf16(d, dah, dbh, idesc_mix, (accumulate | kk) != 0 ? 1u : 0u);
              umma_bf16(d, dah, dbl, idesc_mix, 1u);
              umma_bf16(d, dal, dbh, idesc_mix, 1u);
            }
          }
      };
      if (TCM) {
        // gated conv: per time step 4 K-steps (tap 0: channels 0-15, 16-31; tap 1: ...) x 3 split products per row tile
        const uint32_t idesc_conv = umma_idesc_bf16(128, 64, 0, 0);
        const uint32_t zimg = smem_u32(sA);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(z_ready, t & 1);
          tc_fence_after();
          for (int m = 0; m < g.MT; ++m) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t zh = zimg + (uint32_t)((kk >> 1) * 2) * uimg_bytes + (uint32_t)(2 * (kk & 1)) * a_lbo + m * 2048, zl = zh + uimg_bytes;
              const uint64_t dah = umma_desc(zh, a_lbo, 128), dal = umma_desc(zl, a_lbo, 128);
              const uint64_t dbh = umma_desc(wimg + kk * 2 * 1024, 1024, 128), dbl = umma_desc(wimg + 8192 + kk * 2 * 1024, 1024, 128);
              const uint32_t d = tmem + m * 64;
              umma_bf16(d, dah, dbh, idesc_conv, kk != 0 ? 1u : 0u);
              umma_bf16(d, dah, dbl, idesc_conv, 1u);
              umma_bf16(d, dal, dbh, idesc_conv, 1u);
            }
          }
          umma_commit(conv_full);
        }
        mbar_wait(u_ready, 0);
        tc_fence_after();
      }
      for (int s = 0; s < 3; ++s) {
        if (TCM) {
          // a_s = W_s2 u into the scratch tile (columns 256..), once the workers have read a_{s-1} out of it
          mbar_wait(a_empty, (s & 1) ^ 1);
          tc_fence_after();
          mix(2 + 2 * s, 256, 0u);
          umma_commit(a_full);
        }
        for (int hop = 0; hop < 2; ++hop, ++ph) {
          mbar_wait(b_ready, ph & 1);
          tc_fence_after();
          for (int i = 0; i < nslices; ++i, ++n) {
            const uint32_t st = n & 1;
            mbar_wait(&full[st], (n >> 1) & 1);
            tc_fence_after();
            const uint32_t ah = smem_u32(sA + st * stage_bytes), al = ah + half_bytes;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t koff = (uint32_t)(i * 2 + kk) * 256;
              const uint64_t dbh = umma_desc(bh + koff, 128, b_sbo), dbl = umma_desc(bl + koff, 128, b_sbo);
              for (int m = 0; m < g.MT; ++m) {
                const uint64_t dah = umma_desc(ah + kk * 2 * a_lbo + m * 2048, a_lbo, 128);
                const uint64_t dal = umma_desc(al + kk * 2 * a_lbo + m * 2048, a_lbo, 128);
                const uint32_t d = tmem + (hop ? 128 : 0) + m * 64;
                const uint32_t acc = hop ? ((s | i | kk) != 0 ? 1u : 0u) : ((i | kk) != 0 ? 1u : 0u);
                umma_bf16(d, dah, dbh, idesc, acc);
                umma_bf16(d, dah, dbl, idesc, 1u);
                umma_bf16(d, dal, dbh, idesc, 1u);
              }
            }
            umma_commit(&empty[st]);
          }
          if (TCM && hop == 0) mix(1 + 2 * s, 0, 1u);              // q_s = P_s^T a_s + W_s1 u in the same accumulator
          if (TCM && hop == 1 && s == 2) mix(0, 128, 1u);          // H += W_0 u
          umma_commit(hop ? hopb_done : acc1_full);
        }
      }
    }
  } else {
    // ------------------------------ workers: thread = node ------------------------------
    const int q = warp & 3, m = (warp - 2) >> 2;
    const int node = m * 128 + q * 32 + lane;
    const bool valid = node < N;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int wt = threadIdx.x - 64;
    auto worker_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    // store one [32]-channel row of (time step t) as four 16-byte hi/lo units of the B image
    auto put_b_row = [&](int t, const float (&r)[GC]) {
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = __bfloat162float(__float2bfloat16_rn(r[cg * 8 + j]));
          hi[j] = h;
          lo[j] = r[cg * 8 + j] - h;
        }
        const uint32_t u = (uint32_t)(t * 4 + cg) * g.Kpad + node;
        reinterpret_cast<uint4 *>(sBh)[u] = pack8_bf16(hi);
        reinterpret_cast<uint4 *>(sBl)[u] = pack8_bf16(lo);
      }
    };

    if (!TCM) {
      // ---- P1: gated dilated conv (filter / gate taps staged in sW) ----
      for (int i = wt; i < 1024; i += 256) {
        sW[i] = a.w.filter_w[2 * i]; sW[1024 + i] = a.w.filter_w[2 * i + 1];
        sW[2048 + i] = a.w.gate_w[2 * i]; sW[3072 + i] = a.w.gate_w[2 * i + 1];
      }
      worker_sync();
      if (valid) {
  #pragma unroll
        for (int t = 0; t < GWF_NT; ++t) {
          if (t >= nt) continue;
          const float *z0 = a.zin + ((size_t)b * a.Tin + t0 + t) * col + (size_t)node * GC;
          const float *z1 = a.zin + ((size_t)b * a.Tin + t0 + t + a.dil) * col + (size_t)node * GC;
          const size_t ro = ((size_t)b * a.Tout + t0 + t) * col + (size_t)node * GC;
          float r0[GC], r1[GC];
          load_row(z0, r0);
          load_row(z1, r1);
          if (a.has_in_bn) {
  #pragma unroll
            for (int c = 0; c < GC; ++c) {
              const float sc = a.in_scale[c], sh = a.in_shift[c];
              r0[c] = fmaf(r0[c], sc, sh);
              r1[c] = fmaf(r1[c], sc, sh);
            }
          }
          float *fo = a.f + ro, *go = a.g + ro, *uo = sU + ((size_t)t * N + node) * GC;
  #pragma unroll 1
          for (int cg = 0; cg < GC; cg += 4) {
            float fv[4], gv[4];
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int co = cg + j;
              const float4 *wf0 = reinterpret_cast<const float4 *>(sW + co * GC);
              const float4 *wf1 = reinterpret_cast<const float4 *>(sW + 1024 + co * GC);
              const float4 *wg0 = reinterpret_cast<const float4 *>(sW + 2048 + co * GC);
              const float4 *wg1 = reinterpret_cast<const float4 *>(sW + 3072 + co * GC);
              float f0 = a.w.filter_b[co], f1 = 0.f, f2 = 0.f, f3 = 0.f, g0 = a.w.gate_b[co], g1 = 0.f, g2 = 0.f, g3 = 0.f;
  #pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 A0 = wf0[c4], A1 = wf1[c4], G0 = wg0[c4], G1 = wg1[c4];
                ffma2(f0, f1, A0.x, A0.y, r0[4 * c4], r0[4 * c4 + 1]); ffma2(f2, f3, A0.z, A0.w, r0[4 * c4 + 2], r0[4 * c4 + 3]);
                ffma2(f0, f1, A1.x, A1.y, r1[4 * c4], r1[4 * c4 + 1]); ffma2(f2, f3, A1.z, A1.w, r1[4 * c4 + 2], r1[4 * c4 + 3]);
                ffma2(g0, g1, G0.x, G0.y, r0[4 * c4], r0[4 * c4 + 1]); ffma2(g2, g3, G0.z, G0.w, r0[4 * c4 + 2], r0[4 * c4 + 3]);
                ffma2(g0, g1, G1.x, G1.y, r1[4 * c4], r1[4 * c4 + 1]); ffma2(g2, g3, G1.z, G1.w, r1[4 * c4 + 2], r1[4 * c4 + 3]);
              }
              const float af = (f0 + f1) + (f2 + f3), ag = (g0 + g1) + (g2 + g3);
              fv[j] = tanhf(af);
              gv[j] = 1.0f / (1.0f + expf(-ag));
            }
            *reinterpret_cast<float4 *>(fo + cg) = make_float4(fv[0], fv[1], fv[2], fv[3]);
            *reinterpret_cast<float4 *>(go + cg) = make_float4(gv[0], gv[1], gv[2], gv[3]);
            *reinterpret_cast<float4 *>(uo + cg) = make_float4(fv[0] * gv[0], fv[1] * gv[1], fv[2] * gv[2], fv[3] * gv[3]);
          }
        }
      }
      worker_sync();
    } else {
      // ---- P1 on tcgen05: the gated dilated conv as D[128 nodes, 64 = filter | gate] = sum_tap Z_tap[128, 32] W_tap[32, 64] ----
      // One time step at a time: its two input rows (taps t and t + dil, BatchNorm applied on load) become bf16 hi/lo
      // A-operand images in the (still idle) support-slice ring, the issuer runs 4 K-steps x 3 split products per row
      // tile into TMEM columns [64 m, 64 m + 64), the workers turn the accumulators into f, g (stash) and the u images.
      {
        // conv weights as B images [hi, lo][8 chunks = (tap, ci / 8)][64 rows = filter co | gate co][8 ci]
        for (int i = wt; i < 8 * 64; i += 256) {
          const int chunk = i >> 6, r = i & 63, tap = chunk >> 2, c = chunk & 3;
          const float *wsrc = (r < 32 ? a.w.filter_w : a.w.gate_w) + ((size_t)(r & 31) * 32 + c * 8) * 2 + tap;
          float hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float w = wsrc[2 * j];
            const float h = __bfloat162float(__float2bfloat16_rn(w));
            hi[j] = h; lo[j] = w - h;
          }
          reinterpret_cast<uint4 *>(sWimg)[i] = pack8_bf16(hi);
          reinterpret_cast<uint4 *>(sWimg + 8192)[i] = pack8_bf16(lo);
        }
      }
      for (int t = 0; t < nt; ++t) {
        if (valid) {
#pragma unroll
          for (int tap = 0; tap < 2; ++tap) {
            const float *zr = a.zin + ((size_t)b * a.Tin + t0 + t + tap * a.dil) * col + (size_t)node * GC;
            float r[GC];
            load_row(zr, r);
            if (a.has_in_bn) {
#pragma unroll
              for (int c = 0; c < GC; ++c) r[c] = fmaf(r[c], a.in_scale[c], a.in_shift[c]);
            }
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              float hi[8], lo[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float h = __bfloat162float(__float2bfloat16_rn(r[c8 * 8 + j]));
                hi[j] = h; lo[j] = r[c8 * 8 + j] - h;
              }
              const uint32_t unit = (uint32_t)c8 * rows + (uint32_t)node;
              reinterpret_cast<uint4 *>(sA + (size_t)(tap * 2) * uimg_bytes)[unit] = pack8_bf16(hi);
              reinterpret_cast<uint4 *>(sA + (size_t)(tap * 2 + 1) * uimg_bytes)[unit] = pack8_bf16(lo);
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(z_ready);
        mbar_wait(conv_full, t & 1);
        tc_fence_after();
        float acc[64];
        if (m < g.MT) {
          float h32[32];
          tmem_ld32(tmem + lane_base + m * 64, h32);
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[c] = h32[c];
          tmem_ld32(tmem + lane_base + m * 64 + 32, h32);
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[32 + c] = h32[c];
        }
        tc_fence_before();
        if (valid) {
          const size_t ro = ((size_t)b * a.Tout + t0 + t) * col + (size_t)node * GC;
          float *fo = a.f + ro, *go = a.g + ro;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            float fv[8], gv[8], hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int co = c8 * 8 + j;
              fv[j] = tanhf(acc[co] + a.w.filter_b[co]);
              gv[j] = 1.0f / (1.0f + expf(-(acc[32 + co] + a.w.gate_b[co])));
              const float x = fv[j] * gv[j];
              const float h = __bfloat162float(__float2bfloat16_rn(x));
              hi[j] = h; lo[j] = x - h;
            }
            *reinterpret_cast<float4 *>(fo + c8 * 8) = make_float4(fv[0], fv[1], fv[2], fv[3]);
            *reinterpret_cast<float4 *>(fo + c8 * 8 + 4) = make_float4(fv[4], fv[5], fv[6], fv[7]);
            *reinterpret_cast<float4 *>(go + c8 * 8) = make_float4(gv[0], gv[1], gv[2], gv[3]);
            *reinterpret_cast<float4 *>(go + c8 * 8 + 4) = make_float4(gv[4], gv[5], gv[6], gv[7]);
            const uint32_t unit = (uint32_t)c8 * rows + (uint32_t)node;
            reinterpret_cast<uint4 *>(sUimg + (size_t)(t * 2) * uimg_bytes)[unit] = pack8_bf16(hi);
            reinterpret_cast<uint4 *>(sUimg + (size_t)(t * 2 + 1) * uimg_bytes)[unit] = pack8_bf16(lo);
          }
        }
      }
      worker_sync();
    }
    if (!TCM) {
      // the 7 transposed [32x32] blocks of the gcn 1x1 conv: sW[k][ci][co] = mlp_w[co][k*32 + ci]
      for (int i = wt; i < 7 * 1024; i += 256) {
        const int k = i >> 10, co = (i >> 5) & 31, ci = i & 31;
        sW[k * 1024 + ci * 32 + co] = a.w.mlp_w[(size_t)co * 224 + k * 32 + ci];
      }
      worker_sync();
    } else {
      // the 7 blocks as K-major B images (K = input channel): [k][hi, lo][4 chunks][32 rows = co][8 ci]
      for (int i = wt; i < 7 * 4 * 32; i += 256) {
        const int k = i >> 7, c = (i >> 5) & 3, co = i & 31;
        const float *src = a.w.mlp_w + (size_t)co * 224 + k * 32 + c * 8;
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = __bfloat162float(__float2bfloat16_rn(src[j]));
          hi[j] = h; lo[j] = src[j] - h;
        }
        reinterpret_cast<uint4 *>(sWimg + (size_t)(k * 2) * 2048)[c * 32 + co] = pack8_bf16(hi);
        reinterpret_cast<uint4 *>(sWimg + (size_t)(k * 2 + 1) * 2048)[c * 32 + co] = pack8_bf16(lo);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(u_ready);
    }

    // ---- P2: per support: a_s -> hop A -> q_s -> hop B ----
    for (int s = 0; s < 3; ++s) {
      float arow[GWF_NT][GC];
      if (TCM) {
        mbar_wait(a_full, s & 1);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < GWF_NT; ++t) {
          if (t >= nt) continue;
          if (m < g.MT) tmem_ld32(tmem + lane_base + 256 + m * 64 + t * 32, arow[t]);
          if (valid) store_row(a.a[s] + ((size_t)b * a.Tout + t0 + t) * col + (size_t)node * GC, arow[t]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_empty);
      } else if (valid) {
#pragma unroll
        for (int t = 0; t < GWF_NT; ++t) {
          if (t >= nt) continue;
          float u[GC];
          load_row(sU + ((size_t)t * N + node) * GC, u);
#pragma unroll
          for (int c = 0; c < GC; ++c) arow[t][c] = 0.f;
          matvec_t_reg(sW + (2 + 2 * s) * 1024, u, arow[t]);
          store_row(a.a[s] + ((size_t)b * a.Tout + t0 + t) * col + (size_t)node * GC, arow[t]);
        }
      }
      if (s > 0) mbar_wait(hopb_done, (s - 1) & 1);          // hop B of the previous support has read the B image
      if (valid) {
#pragma unroll
        for (int t = 0; t < GWF_NT; ++t)
          if (t < nt) put_b_row(t, arow[t]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready);
      // q_s = W_s1 u + m_s
      mbar_wait(acc1_full, s & 1);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < GWF_NT; ++t) {
        if (t >= nt) continue;
        float qrow[GC];
        if (m < g.MT) tmem_ld32(tmem + lane_base + m * 64 + t * 32, qrow);
        if (valid) {
          if (!TCM) {
            float u[GC];
            load_row(sU + ((size_t)t * N + node) * GC, u);
            matvec_t_reg(sW + (1 + 2 * s) * 1024, u, qrow);
          }
          store_row(a.q[s] + ((size_t)b * a.Tout + t0 + t) * col + (size_t)node * GC, qrow);
          put_b_row(t, qrow);                                 // hop A has completed (acc1_full): the image is free
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready);
    }

    // ---- P3: h = H + W_0 u + b; dropout; residual; z; BatchNorm partial sums ----
    mbar_wait(hopb_done, 0);                                  // third completion of the barrier (phases 0, 1, 0)
    tc_fence_after();
    float csum[GC], csq[GC];
#pragma unroll
    for (int c = 0; c < GC; ++c) { csum[c] = 0.f; csq[c] = 0.f; }
#pragma unroll
    for (int t = 0; t < GWF_NT; ++t) {
      if (t >= nt) continue;
      float h[GC];
      if (m < g.MT) tmem_ld32(tmem + lane_base + 128 + m * 64 + t * 32, h);
      if (valid) {
        const size_t ro = ((size_t)b * a.Tout + t0 + t) * col + (size_t)node * GC;
        float r[GC];
        load_row(a.zin + ((size_t)b * a.Tin + t0 + t + a.dil) * col + (size_t)node * GC, r);
#pragma unroll
        for (int c = 0; c < GC; ++c) h[c] += a.w.mlp_b[c];
        if (!TCM) {
          float u[GC];
          load_row(sU + ((size_t)t * N + node) * GC, u);
          matvec_t_reg(sW, u, h);
        }
        if (a.drop_thr) dropout_row(h, (uint64_t)ro, a.drop_thr, a.drop_scale, a.key);
#pragma unroll
        for (int c = 0; c < GC; ++c) {
          const float rv = a.has_in_bn ? fmaf(r[c], a.in_scale[c], a.in_shift[c]) : r[c];
          h[c] += rv;
          csum[c] += h[c];
          csq[c] = fmaf(h[c], h[c], csq[c]);
        }
        store_row(a.zout + ro, h);
      }
    }
    tc_fence_before();
    if (a.collect_stats) {
      const float s1 = warp_colsum32(csum), s2 = warp_colsum32(csq);     // lane c: channel c over the warp's 32 nodes
      sRed[(warp - 2) * 64 + lane] = s1;
      sRed[(warp - 2) * 64 + 32 + lane] = s2;
      worker_sync();
      if (wt < 64) {
        float tsum = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) tsum += sRed[w8 * 64 + wt];
        atomicAdd(a.sums + wt, (double)tsum);                            // [0,32): sums, [32,64): sums of squares
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCM ? 512 : 256);
}

// ---------------------------------------------------------------------------
// skip convolutions of ALL layers, hoisted out of the layer kernels (only the last time step of each layer's
// u = f*g reaches the output because skip[..., -T:] truncation ends at T = 1; model.py:189-197):
//   skip[b,n,:] = sum_l ( Wk_l u_l[b, Tout_l - 1, n, :] + bk_l )
// grid (ceil(N/64), B), 256 threads = skip channels; a 64-node tile of u_l is staged in shared memory per layer.
// ---------------------------------------------------------------------------
struct GwSkipArgs {
  int L, N, B;
  int Tout[GW_MAX_LAYERS];
  const float *f[GW_MAX_LAYERS], *g[GW_MAX_LAYERS];      // [B,Tout_l,N,32]
  const float *skip_w[GW_MAX_LAYERS], *skip_b[GW_MAX_LAYERS];
  float *dskip_w[GW_MAX_LAYERS], *dskip_b[GW_MAX_LAYERS];
  float *skip;              // forward out [B,N,256]
  const float *dskip;       // backward in [B,N,256]
  float *dus;               // backward out [L][B,N,32]
};

constexpr int SKIP_NT = 64;

__global__ void __launch_bounds__(GSKIP) gw_skip_fwd_kernel(GwSkipArgs a) {
  __shared__ __align__(16) float Us[SKIP_NT][GC];
  const int k = threadIdx.x, b = blockIdx.y, n0 = blockIdx.x * SKIP_NT;
  const int nn = min(SKIP_NT, a.N - n0);
  float acc[SKIP_NT];
  float bsum = 0.f;
#pragma unroll
  for (int n = 0; n < SKIP_NT; ++n) acc[n] = 0.f;
  for (int l = 0; l < a.L; ++l) {
    const size_t base = (((size_t)b * a.Tout[l] + a.Tout[l] - 1) * a.N + n0) * GC;
    __syncthreads();
    for (int i = k; i < SKIP_NT * GC / 4; i += GSKIP) {
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i * 4 < nn * GC) {
        const float4 fv = *reinterpret_cast<const float4 *>(a.f[l] + base + (size_t)i * 4);
        const float4 gv = *reinterpret_cast<const float4 *>(a.g[l] + base + (size_t)i * 4);
        u = make_float4(fv.x * gv.x, fv.y * gv.y, fv.z * gv.z, fv.w * gv.w);
      }
      reinterpret_cast<float4 *>(&Us[0][0])[i] = u;
    }
    __syncthreads();
    float wk[GC];
    load_row(a.skip_w[l] + (size_t)k * GC, wk);
    bsum += a.skip_b[l][k];
#pragma unroll
    for (int n = 0; n < SKIP_NT; ++n) {
      const float4 *y = reinterpret_cast<const float4 *>(&Us[n][0]);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 8; c4 += 2) {
        const float4 y0 = y[c4], y1 = y[c4 + 1];
        s0 = fmaf(wk[4 * c4], y0.x, s0); s0 = fmaf(wk[4 * c4 + 1], y0.y, s0);
        s0 = fmaf(wk[4 * c4 + 2], y0.z, s0); s0 = fmaf(wk[4 * c4 + 3], y0.w, s0);
        s1 = fmaf(wk[4 * c4 + 4], y1.x, s1); s1 = fmaf(wk[4 * c4 + 5], y1.y, s1);
        s1 = fmaf(wk[4 * c4 + 6], y1.z, s1); s1 = fmaf(wk[4 * c4 + 7], y1.w, s1);
      }
      acc[n] += s0 + s1;
    }
  }
  float *sk = a.skip + ((size_t)b * a.N + n0) * GSKIP + k;
#pragma unroll
  for (int n = 0; n < SKIP_NT; ++n)
    if (n < nn) sk[(size_t)n * GSKIP] = acc[n] + bsum;
}

// backward of the same: grid (L, B), 256 threads.
//   dus_l[b,n,:] = Wk_l^T dskip[b,n,:]  (thread = node),  dWk_l += dskip[b]^T u_l,  dbk_l += column sums (thread = channel)
__global__ void __launch_bounds__(GSKIP) gw_skip_bwd_kernel(GwSkipArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, l = blockIdx.x, b = blockIdx.y, N = a.N;
  float *Wk = smem;                    // [256][32]
  float *Us = smem + GSKIP * GC;       // [N][32]
  const float *ds = a.dskip + (size_t)b * N * GSKIP;
  const size_t base = (((size_t)b * a.Tout[l] + a.Tout[l] - 1) * N) * GC;
  for (int i = tid; i < GSKIP * GC / 4; i += GSKIP)
    reinterpret_cast<float4 *>(Wk)[i] = reinterpret_cast<const float4 *>(a.skip_w[l])[i];
  for (int i = tid; i < N * GC / 4; i += GSKIP) {
    const float4 fv = *reinterpret_cast<const float4 *>(a.f[l] + base + (size_t)i * 4);
    const float4 gv = *reinterpret_cast<const float4 *>(a.g[l] + base + (size_t)i * 4);
    reinterpret_cast<float4 *>(Us)[i] = make_float4(fv.x * gv.x, fv.y * gv.y, fv.z * gv.z, fv.w * gv.w);
  }
  __syncthreads();
  for (int n = tid; n < N; n += GSKIP) {
    float du[GC];
#pragma unroll
    for (int c = 0; c < GC; ++c) du[c] = 0.f;
    const float4 *dsn = reinterpret_cast<const float4 *>(ds + (size_t)n * GSKIP);
    for (int k4 = 0; k4 < GSKIP / 4; ++k4) {
      const float4 d4 = dsn[k4];
      const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 *wk = reinterpret_cast<const float4 *>(Wk + (size_t)(4 * k4 + j) * GC);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 ww = wk[c4];
          du[4 * c4] = fmaf(ww.x, dd[j], du[4 * c4]); du[4 * c4 + 1] = fmaf(ww.y, dd[j], du[4 * c4 + 1]);
          du[4 * c4 + 2] = fmaf(ww.z, dd[j], du[4 * c4 + 2]); du[4 * c4 + 3] = fmaf(ww.w, dd[j], du[4 * c4 + 3]);
        }
      }
    }
    store_row(a.dus + (((size_t)l * a.B + b) * N + n) * GC, du);
  }
  {
    const int k = tid;
    float acc[GC];
#pragma unroll
    for (int c = 0; c < GC; ++c) acc[c] = 0.f;
    float bsum = 0.f;
    for (int n = 0; n < N; ++n) {
      const float d = ds[(size_t)n * GSKIP + k];
      bsum += d;
      const float4 *y = reinterpret_cast<const float4 *>(Us + (size_t)n * GC);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 yy = y[c4];
        acc[4 * c4] = fmaf(d, yy.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(d, yy.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(d, yy.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(d, yy.w, acc[4 * c4 + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < GC; ++c) atomicAdd(a.dskip_w[l] + (size_t)k * GC + c, acc[c]);
    atomicAdd(a.dskip_b[l] + k, bsum);
  }
}

// sums (double [2][32]) -> stats [4][32] floats: mean, biased var, scale, shift
__global__ void bn_finalize_kernel(const double *sums, double count, const float *gamma, const float *beta, float *stats) {
  const int c = threadIdx.x;
  const double mean = sums[c] / count;
  double var = sums[32 + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + 1e-5));
  stats[c] = (float)mean;
  stats[32 + c] = (float)var;
  stats[64 + c] = gamma[c] * rstd;
  stats[96 + c] = beta[c] - (float)mean * gamma[c] * rstd;
}

// ---------------------------------------------------------------------------
// backward, phase 1: grid (T_out, B).  From d(BN output) of this layer (or nothing for the dead last
// gcn) and dskip, through the gcn and the gate non-linearities, to d(pre-activations) of both convs.
// ---------------------------------------------------------------------------
struct GwBwdArgs {
  int Tin, Tout, dil, N, has_gcn;
  const float *drnext;      // [B,Tout,N,32] grad wrt BN(z) of this layer
  const float *z;           // [B,Tout,N,32] pre-BN output of this layer
  const float *stats;       // this layer's BN [4][32]
  const float *coef;        // BN backward: [0]=gamma*rstd, [1]=S1/M, [2]=S2/M   ([4][32])
  const float *Pt[3]; long long pstride[3];
  step_gw_layer_params w;
  step_gw_layer_grads gr;
  const float *f, *g;
  const float *dus;         // [B,N,32] = Wk^T dskip of this layer (gw_skip_bwd_kernel), added at the last time step
  float *U, *DH, *DZC, *DQ[3], *A[3], *DA, *DU, *DPF, *DPG;
  int stage;                // 0: fused CUDA-core layer; 1: up to dh/du0 (before the tensor-core mixes); 2: after them
  const float *DA3[3];      // stage 2: da_s = P_s dq_s from the tensor-core mix
  uint32_t drop_thr; float drop_scale; uint64_t key;
};

template <int STAGE>
__global__ void __launch_bounds__(GW_THREADS) gw_layer_bwd_kernel(GwBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int N = a.N, tid = threadIdx.x, t = blockIdx.x, b = blockIdx.y;
  float *Y = smem;
  float *Wb = smem + (size_t)N * GC;          // 7*1024
  float *tiles = Wb + 7 * 1024;               // 2*64*33
  const size_t col = (size_t)N * GC;
  const size_t ocol = ((size_t)b * a.Tout + t) * col;
  float *U = a.U + ocol, *DH = a.DH + ocol, *DZC = a.DZC + ocol, *DA = a.DA + ocol, *DU = a.DU + ocol;
  const float *fb = a.f + ocol, *gb = a.g + ocol;

  if (a.has_gcn) {
    // stage 1 only needs W_0, stage 2 the six support blocks
    const int i0 = (STAGE == 2) ? 1024 : 0, i1 = (STAGE == 1) ? 1024 : 7 * 1024;
    for (int i = i0 + tid; i < i1; i += GW_THREADS) {
      const int k = i >> 10, co = (i >> 5) & 31, ci = i & 31;
      Wb[i] = a.w.mlp_w[(size_t)co * 224 + k * 32 + ci];
    }
  }
  __syncthreads();

  // ---- phase 0: u, dz (through BatchNorm), dh (through dropout), du = W0^T dh ----
  if (STAGE != 2) {
  for (int n = tid; n < N; n += GW_THREADS) {
    {
      float fv[GC], gv[GC];
      load_row(fb + (size_t)n * GC, fv);
      load_row(gb + (size_t)n * GC, gv);
#pragma unroll
      for (int c = 0; c < GC; ++c) fv[c] *= gv[c];
      store_row(U + (size_t)n * GC, fv);
    }
    float du[GC];
    if (t == a.Tout - 1) {          // skip path: Wk^T dskip of this layer, from gw_skip_bwd_kernel
      load_row(a.dus + ((size_t)b * N + n) * GC, du);
    } else {
#pragma unroll
      for (int c = 0; c < GC; ++c) du[c] = 0.f;
    }
    if (a.has_gcn) {
      float dh[GC];
      {
        float zz[GC];
        load_row(a.drnext + ocol + (size_t)n * GC, dh);
        load_row(a.z + ocol + (size_t)n * GC, zz);
#pragma unroll
        for (int c = 0; c < GC; ++c) {
          const float xhat = (zz[c] - a.stats[c]) * a.coef[96 + c];
          dh[c] = a.coef[c] * (dh[c] - a.coef[32 + c] - xhat * a.coef[64 + c]);
        }
      }
      store_row(DZC + (size_t)n * GC, dh);
      if (a.drop_thr) dropout_row(dh, (uint64_t)(ocol + (size_t)n * GC), a.drop_thr, a.drop_scale, a.key);
      store_row(DH + (size_t)n * GC, dh);
      store_row(Y + (size_t)n * GC, dh);
      matvec_t_reg(Wb, dh, du);
    }
    store_row(DU + (size_t)n * GC, du);
  }
  __syncthreads();
  }  // stage != 2

  if (a.has_gcn && STAGE != 2) {
    // d mlp bias = column sums of dh
    {
      const int c = tid & 31, grp = tid >> 5;
      float s = 0.f;
      for (int n = grp; n < N; n += GW_THREADS / 32) s += Y[(size_t)n * GC + c];
      tiles[grp * 32 + c] = s;
      __syncthreads();
      if (tid < 32) {
        float ts = 0.f;
#pragma unroll
        for (int g2 = 0; g2 < GW_THREADS / 32; ++g2) ts += tiles[g2 * 32 + c];
        atomicAdd(a.gr.mlp_b + c, ts);
      }
      __syncthreads();
    }
    {
      const float *const Xs[1] = {DH};
      float *const Ds[1] = {a.gr.mlp_w};
      outer_acc_tiled<1>(Xs, U, N, tiles, Ds, 224, 1);   // block 0: dW0[co][ci] += dh[n][co] u[n][ci]
    }
  }
  if (STAGE == 1) return;
  if (a.has_gcn && STAGE == 2) {
    // dq_s = P_s dh and da_s = P_s dq_s were produced by tc_mix_kernel: du += sum_s W_s1^T dq_s + W_s2^T da_s
    // (a_s = W_s2 u, needed by dP, comes from the forward stash)
    for (int n = tid; n < N; n += GW_THREADS) {
      const size_t ro = ocol + (size_t)n * GC;
      float du[GC], xa[GC], xb[GC];
      load_row(a.DQ[0] + ro, xa);
      load_row(DU + (size_t)n * GC, du);
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        load_row(a.DA3[s] + ro, xb);                        // in flight under the first product
        matvec_t_reg(Wb + (1 + 2 * s) * 1024, xa, du);
        if (s < 2) load_row(a.DQ[s + 1] + ro, xa);          // in flight under the second product
        matvec_t_reg(Wb + (2 + 2 * s) * 1024, xb, du);
      }
      store_row(DU + (size_t)n * GC, du);
    }
    {
      const float *const Xa[3] = {a.DQ[0] + ocol, a.DA3[0] + ocol, a.DQ[1] + ocol};
      float *const Da[3] = {a.gr.mlp_w + 1 * 32, a.gr.mlp_w + 2 * 32, a.gr.mlp_w + 3 * 32};
      outer_acc_tiled<3>(Xa, U, N, tiles, Da, 224, 1);
      const float *const Xb[3] = {a.DA3[1] + ocol, a.DQ[2] + ocol, a.DA3[2] + ocol};
      float *const Db[3] = {a.gr.mlp_w + 4 * 32, a.gr.mlp_w + 5 * 32, a.gr.mlp_w + 6 * 32};
      outer_acc_tiled<3>(Xb, U, N, tiles, Db, 224, 1);
    }
  }
  if (a.has_gcn && STAGE == 0) {
    for (int s = 0; s < 3; ++s) {
      const float *Pts = a.Pt[s] + (size_t)b * a.pstride[s];
      const float *W1 = Wb + (1 + 2 * s) * 1024, *W2 = Wb + (2 + 2 * s) * 1024;
      float *DQ = a.DQ[s] + ocol, *As = a.A[s] + ocol;
      // Y holds dh.  dq = P dh
      mix_nodes(Pts, Y, N, [&](int v, float (&acc)[GC]) { store_row(DQ + (size_t)v * GC, acc); });
      __syncthreads();
      copy_to_smem(Y, DQ, N);
      __syncthreads();
      // da = P dq
      mix_nodes(Pts, Y, N, [&](int v, float (&acc)[GC]) { store_row(DA + (size_t)v * GC, acc); });
      __syncthreads();
      // per node: a = W_s2 u (stash for dP), du += W_s1^T dq + W_s2^T da
      for (int n = tid; n < N; n += GW_THREADS) {
        {
          float u[GC];
          load_row(U + (size_t)n * GC, u);
          float *ar = As + (size_t)n * GC;
          matvec_chunks(W2, u, [&](int cg, float4 v) { st4(ar + 4 * cg, v); });
        }
        float du[GC];
        load_row(DU + (size_t)n * GC, du);
        matvec_t_acc(W1, Y + (size_t)n * GC, du);   // dq
        matvec_t_acc(W2, DA + (size_t)n * GC, du);
        store_row(DU + (size_t)n * GC, du);
      }
      outer_acc(DQ, U, N, tiles, a.gr.mlp_w + (1 + 2 * s) * 32, 224);
      outer_acc(DA, U, N, tiles, a.gr.mlp_w + (2 + 2 * s) * 32, 224);
      copy_to_smem(Y, DH, N);
      __syncthreads();
    }
  }

  // ---- through the gate: d(pre-tanh), d(pre-sigmoid) ----
  for (int n = tid; n < N; n += GW_THREADS) {
    float fv[GC], gv[GC], du[GC], pf[GC], pg[GC];
    load_row(fb + (size_t)n * GC, fv);
    load_row(gb + (size_t)n * GC, gv);
    load_row(DU + (size_t)n * GC, du);
#pragma unroll
    for (int c = 0; c < GC; ++c) {
      pf[c] = du[c] * gv[c] * (1.f - fv[c] * fv[c]);
      pg[c] = du[c] * fv[c] * gv[c] * (1.f - gv[c]);
    }
    store_row(a.DPF + ocol + (size_t)n * GC, pf);
    store_row(a.DPG + ocol + (size_t)n * GC, pg);
  }
}

// ---------------------------------------------------------------------------
// dP_s[b][v][w] += sum_t sum_c ( Q_s[b,t,v,c] DH[b,t,w,c] + A_s[b,t,v,c] DQ_s[b,t,w,c] )
// grid (ceil(N/64), ceil(N/64), 3*B); 64x64 tile, 4x4 per thread.
// ---------------------------------------------------------------------------
struct GwDpArgs {
  int N, T, B;
  const float *Q[3], *A[3], *DQ[3], *DH;
  float *dP[3]; long long pstride[3];
};

__global__ void __launch_bounds__(256) gw_dP_kernel(GwDpArgs a) {
  __shared__ __align__(16) float Xs[GC][68];
  __shared__ __align__(16) float Ws[GC][68];
  const int N = a.N, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int s = blockIdx.z / a.B, b = blockIdx.z % a.B;
  const int v0 = blockIdx.y * 64, w0 = blockIdx.x * 64;
  const size_t col = (size_t)N * GC;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int it = 0; it < 2 * a.T; ++it) {
    const int t = it >> 1, pair = it & 1;
    const float *X = (pair ? a.A[s] : a.Q[s]) + ((size_t)b * a.T + t) * col;
    const float *Wm = (pair ? a.DQ[s] : a.DH) + ((size_t)b * a.T + t) * col;
    __syncthreads();
    for (int i = tid; i < 64 * 8; i += 256) {
      const int n = i >> 3, c4 = (i & 7) * 4;
      float4 xv = make_float4(0, 0, 0, 0), wv = xv;
      if (v0 + n < N) xv = *reinterpret_cast<const float4 *>(X + (size_t)(v0 + n) * GC + c4);
      if (w0 + n < N) wv = *reinterpret_cast<const float4 *>(Wm + (size_t)(w0 + n) * GC + c4);
      Xs[c4][n] = xv.x; Xs[c4 + 1][n] = xv.y; Xs[c4 + 2][n] = xv.z; Xs[c4 + 3][n] = xv.w;
      Ws[c4][n] = wv.x; Ws[c4 + 1][n] = wv.y; Ws[c4 + 2][n] = wv.z; Ws[c4 + 3][n] = wv.w;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < GC; ++c) {
      const float4 xv = *reinterpret_cast<const float4 *>(&Xs[c][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4 *>(&Ws[c][tx * 4]);
      const float x[4] = {xv.x, xv.y, xv.z, xv.w}, w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
  }
  float *dP = a.dP[s] + (size_t)b * a.pstride[s];
  const bool shared_across_batch = (a.pstride[s] == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = v0 + ty * 4 + i;
    if (v >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w = w0 + tx * 4 + j;
      if (w >= N) continue;
      if (shared_across_batch) atomicAdd(dP + (size_t)v * N + w, acc[i][j]);
      else dP[(size_t)v * N + w] += acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------
// backward, phase 2: grid (T_in, B).  Transposed temporal conv + residual path -> d(normalised
// input); conv weight/bias grads; partial sums for the previous layer's BatchNorm backward.
// ---------------------------------------------------------------------------
struct GwBwdInArgs {
  int Tin, Tout, dil, N, has_gcn, has_in_bn;
  const float *zin;          // [B,Tin,N,32] pre-BN output of the previous layer (or x0)
  const float *in_stats;     // previous layer's BN [4][32] (mean,var,scale,shift)
  const float *DPF, *DPG, *DZC;
  step_gw_layer_params w;
  step_gw_layer_grads gr;
  float *drin;               // [B,Tin,N,32]
  double *sums;              // previous layer's backward sums [2][32] (S1, S2) or null
};

__global__ void __launch_bounds__(GW_THREADS, 2) gw_layer_bwd_in_kernel(GwBwdInArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int N = a.N, tid = threadIdx.x, tau = blockIdx.x, b = blockIdx.y;
  float *Y = smem;                       // r (normalised input) [N][32]
  float *Wb = smem + (size_t)N * GC;     // 4*1024 conv weights
  float *tiles = Wb + 4 * 1024;          // 3*64*32 (outer_acc_tiled<2>)
  float *red = tiles + 3 * 64 * 32;      // 128 floats
  const size_t col = (size_t)N * GC;
  const bool has0 = tau < a.Tout, has1 = (tau - a.dil) >= 0;
  const size_t c0 = ((size_t)b * a.Tout + tau) * col, c1 = ((size_t)b * a.Tout + (tau - a.dil)) * col;
  const size_t icol = ((size_t)b * a.Tin + tau) * col;

  for (int i = tid; i < 1024; i += GW_THREADS) {
    Wb[i] = a.w.filter_w[2 * i]; Wb[1024 + i] = a.w.filter_w[2 * i + 1];
    Wb[2048 + i] = a.w.gate_w[2 * i]; Wb[3072 + i] = a.w.gate_w[2 * i + 1];
  }
  if (tid < 128) red[tid] = 0.f;
  __syncthreads();

  // every warp walks its node slots with all 32 lanes (inactive lanes carry zero rows) so that the per-channel sums
  // S1 = sum d, S2 = sum d * xhat, d filter_b, d gate_b can be reduced with register-tile column sums
  for (int n0 = 0; n0 < N; n0 += GW_THREADS) {
    if (n0 + (int)(tid & ~31u) >= N) continue;          // whole warp past the end
    const int n = n0 + tid;
    const bool active = n < N;
    const size_t ro = (size_t)(active ? n : 0) * GC;
    float acc[GC], x[GC], xb[GC];
    float k1 = 0.f, k2 = 0.f, kf = 0.f, kg = 0.f;
#pragma unroll
    for (int c = 0; c < GC; ++c) acc[c] = 0.f;
    auto load_or_zero = [&](const float *p, float (&r)[GC]) {
      if (active) load_row(p, r);
      else {
#pragma unroll
        for (int c = 0; c < GC; ++c) r[c] = 0.f;
      }
    };
    if (has0) {
      load_or_zero(a.DPF + c0 + ro, x);
      load_or_zero(a.DPG + c0 + ro, xb);                  // in flight under the first product
      matvec_t_reg(Wb, x, acc);
      kf = warp_colsum32(x);
      if (has1) load_or_zero(a.DPF + c1 + ro, x);
      matvec_t_reg(Wb + 2048, xb, acc);
      kg = warp_colsum32(xb);
    } else if (has1) {
      load_or_zero(a.DPF + c1 + ro, x);
    }
    if (has1) {
      load_or_zero(a.DPG + c1 + ro, xb);
      matvec_t_reg(Wb + 1024, x, acc);
      if (a.has_gcn) load_or_zero(a.DZC + c1 + ro, x);
      matvec_t_reg(Wb + 3072, xb, acc);
      if (a.has_gcn) {
#pragma unroll
        for (int c = 0; c < GC; ++c) acc[c] += x[c];
      }
    }
    if (active) store_row(a.drin + icol + ro, acc);
    load_or_zero(a.zin + icol + ro, x);
    if (a.has_in_bn) {
#pragma unroll
      for (int c = 0; c < GC; ++c) {
        const float mean = a.in_stats[c], var = a.in_stats[32 + c];
        const float xhat = (x[c] - mean) * (1.0f / sqrtf(var + 1e-5f));
        xb[c] = acc[c] * xhat;
        x[c] = fmaf(x[c], a.in_stats[64 + c], a.in_stats[96 + c]);
      }
    }
    if (active) store_row(Y + ro, x);
    if (a.has_in_bn) {
      k1 = warp_colsum32(acc);
      k2 = warp_colsum32(xb);
    }
    const int lane = tid & 31;
    if (a.has_in_bn) { atomicAdd(&red[lane], k1); atomicAdd(&red[32 + lane], k2); }
    if (has0) { atomicAdd(&red[64 + lane], kf); atomicAdd(&red[96 + lane], kg); }
  }
  __syncthreads();
  if (tid < 32) {
    if (a.sums != nullptr) {
      atomicAdd(a.sums + tid, (double)red[tid]);
      atomicAdd(a.sums + 32 + tid, (double)red[32 + tid]);
    }
    if (has0) {
      atomicAdd(a.gr.filter_b + tid, red[64 + tid]);
      atomicAdd(a.gr.gate_b + tid, red[96 + tid]);
    }
  }
  // conv weight grads: dW[co][ci][tap] += sum_n dpre[tau - tap*dil][n][co] * r[tau][n][ci]  (r rows live in Y)
  __syncthreads();
  if (has0) {
    const float *const Xs[2] = {a.DPF + c0, a.DPG + c0};
    float *const Ds[2] = {a.gr.filter_w, a.gr.gate_w};
    outer_acc_tiled<2>(Xs, Y, N, tiles, Ds, GC, 2);
  }
  if (has1) {
    const float *const Xs[2] = {a.DPF + c1, a.DPG + c1};
    float *const Ds[2] = {a.gr.filter_w + 1, a.gr.gate_w + 1};
    outer_acc_tiled<2>(Xs, Y, N, tiles, Ds, GC, 2);
  }
}

// backward sums (double [2][32]: S1 = sum d, S2 = sum d*xhat) -> coef [4][32] + BN param grads
__global__ void bn_bwd_finalize_kernel(const double *sums, double count, const float *gamma, const float *stats,
                                       float *coef, float *dgamma, float *dbeta) {
  const int c = threadIdx.x;
  const float rstd = 1.0f / sqrtf(stats[32 + c] + 1e-5f);
  coef[c] = gamma[c] * rstd;
  coef[32 + c] = (float)(sums[c] / count);
  coef[64 + c] = (float)(sums[32 + c] / count);
  coef[96 + c] = rstd;
  dgamma[c] = (float)sums[32 + c];
  dbeta[c] = (float)sums[c];
}

}  // namespace stepk

using namespace stepk;

extern "C" size_t step_gwnet_stash_floats(int B, int N, int n_layers) {
  if (B <= 0 || N <= 0 || n_layers <= 0 || n_layers > GW_MAX_LAYERS) return 0;
  return make_plan(B, N, n_layers).total;
}

// Node mixes on the tensor cores (tc_mix_kernel, split-bf16 = fp32-level accuracy) unless STEP_B200_GW_MIX=simt
static bool gw_use_tc(int N) {
  const char *e = getenv("STEP_B200_GW_MIX");
  if (e != nullptr && strcmp(e, "simt") == 0) return false;
  return mix_smem_bytes(mix_geom(N)) <= 227 * 1024;
}

// One fused launch per layer (gw_fused_fwd_kernel) when the graph fits two 128-row tiles and the shared-memory plan;
// STEP_B200_GW_FUSED=0 selects the five-launch split path (A/B comparison, tests)
static bool gw_use_fused(const MixGeom &g) {
  const char *e = getenv("STEP_B200_GW_FUSED");
  if (e != nullptr && strcmp(e, "0") == 0) return false;
  return g.MT <= 2 && gwf_smem_bytes(g, false) <= 227 * 1024;
}
// STEP_B200_GW_FUSED=1 keeps the CUDA-core channel mixes inside the fused kernel (default: on tcgen05 when the plan fits)
static bool gw_fused_tc_mix(const MixGeom &g) {
  const char *e = getenv("STEP_B200_GW_FUSED");
  if (e != nullptr && strcmp(e, "1") == 0) return false;
  return gwf_smem_bytes(g, true) <= 227 * 1024;
}

static size_t fwd_smem_bytes(int N) { return ((size_t)N * GC + 7 * 1024) * sizeof(float); }
static size_t bwd_smem_bytes(int N) { return ((size_t)N * GC + 7 * 1024 + 7 * 64 * 32) * sizeof(float); }
static size_t bwd_in_smem_bytes(int N) { return ((size_t)N * GC + 4 * 1024 + 3 * 64 * 32 + 128) * sizeof(float); }

static int gw_prepare(int N) {
  if (bwd_smem_bytes(N) > 227 * 1024)
    return fail(STEP_EUNSUPPORTED, "gwnet: N=%lld needs more shared memory than one SM has", N);
  int rc;
  if ((rc = allow_smem(gw_layer_fwd_kernel<0>, 227 * 1024))) return rc;
  if ((rc = allow_smem(gw_layer_fwd_kernel<1>, 227 * 1024))) return rc;
  if ((rc = allow_smem(gw_layer_fwd_kernel<2>, 227 * 1024))) return rc;
  if ((rc = allow_smem(gw_layer_fwd_kernel<3>, 227 * 1024))) return rc;
  if ((rc = allow_smem(gw_layer_bwd_kernel<0>, 227 * 1024))) return rc;
  if ((rc = allow_smem(gw_layer_bwd_kernel<1>, 227 * 1024))) return rc;
  if ((rc = allow_smem(gw_layer_bwd_kernel<2>, 227 * 1024))) return rc;
  if ((rc = allow_smem(gw_layer_bwd_in_kernel, 227 * 1024))) return rc;
  return STEP_OK;
}

extern "C" int step_gwnet_stack_fwd(const float *x0, const float *P1, const float *P2, const float *P3,
                                    const step_gw_layer_params *Lp, int n_layers, int B, int N, int training,
                                    float drop_p, unsigned long long seed, float *skip_out, float *bn_stats, float *stash,
                                    void *stream) {
  STEP_REQUIRE(x0 && P1 && P2 && P3 && Lp && skip_out && bn_stats && stash, "gwnet_fwd: null pointer");
  STEP_REQUIRE(B > 0 && B <= 65535 && N > 0 && n_layers >= 1 && n_layers <= GW_MAX_LAYERS, "gwnet_fwd: bad shape");
  int rc = gw_prepare(N);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const GwPlan p = make_plan(B, N, n_layers);
  cudaError_t e = cudaMemsetAsync(stash + p.off_sums_fwd, 0, (size_t)n_layers * 2 * 32 * sizeof(double), st);
  if (e != cudaSuccess) return fail_msg((int)e, cudaGetErrorString(e));
  const bool use_tc = gw_use_tc(N);
  const MixGeom geom = mix_geom(N);
  const size_t img_floats = (mix_images_bytes(N) + 3) / 4;
  if (use_tc) {
    // bf16 hi/lo operand images of P and P^T for the three supports (kept in the stash for the backward pass)
    if ((rc = tc_support_images_launch(P1, B, geom, reinterpret_cast<uint8_t *>(stash + p.off_img[0]), (long long)img_floats * 4, st))) return rc;
    if ((rc = tc_support_images_launch(P2, B, geom, reinterpret_cast<uint8_t *>(stash + p.off_img[1]), (long long)img_floats * 4, st))) return rc;
    if ((rc = tc_support_images_launch(P3, 1, geom, reinterpret_cast<uint8_t *>(stash + p.off_img[2]), 0, st))) return rc;
  }
  for (int i = 0; i < n_layers; ++i) {
    GwFwdArgs a{};
    a.zin = (i == 0) ? x0 : stash + p.off_z[i - 1];
    a.zout = stash + p.off_z[i];
    a.Tin = p.Tin[i]; a.Tout = p.Tout[i]; a.dil = p.dil[i]; a.N = N;
    a.has_gcn = (Lp[i].mlp_w != nullptr);
    a.has_in_bn = (i > 0);
    a.collect_stats = training ? 1 : 0;
    if (i > 0) { a.in_scale = bn_stats + (size_t)(i - 1) * 128 + 64; a.in_shift = bn_stats + (size_t)(i - 1) * 128 + 96; }
    a.P[0] = P1; a.P[1] = P2; a.P[2] = P3;
    a.pstride[0] = (long long)N * N; a.pstride[1] = (long long)N * N; a.pstride[2] = 0;
    a.w = Lp[i];
    a.f = stash + p.off_f[i]; a.g = stash + p.off_g[i];
    for (int s = 0; s < 3; ++s) a.q[s] = stash + p.off_q[i][s];
    a.U = stash + p.off_U; a.M = stash + p.off_M; a.H = stash + p.off_H;
    a.sums = reinterpret_cast<double *>(stash + p.off_sums_fwd) + (size_t)i * 64;
    if (training && drop_p > 0.f) { a.drop_thr = drop_threshold(drop_p); a.drop_scale = 1.f / (1.f - drop_p); }
    else { a.drop_thr = 0; a.drop_scale = 1.f; }
    a.key = rng_key(seed, 0x100u + i);
    if (!use_tc || !a.has_gcn) {
      if (a.has_gcn) gw_layer_fwd_kernel<0><<<dim3(a.Tout, B), GW_THREADS, fwd_smem_bytes(N), st>>>(a);
      else gw_layer_fwd_kernel<1><<<dim3(a.Tout, B), GW_THREADS, fwd_smem_bytes(N), st>>>(a);
      STEP_LAUNCH_CHECK("gw_layer_fwd_kernel");
    } else {
      TcMixArgs m{};
      m.g = geom; m.B = B; m.T = a.Tout; m.nsup = 3; m.transposed_type = 1;
      for (int s = 0; s < 3; ++s) {
        m.img[s] = reinterpret_cast<const uint8_t *>(stash + p.off_img[s]);
        m.img_bstride[s] = (s < 2) ? (long long)img_floats * 4 : 0;
        a.Aout[s] = stash + p.off_a[i][s];
        a.Min[s] = stash + p.off_M3[s];
        a.Oin[s] = stash + p.off_O3[s];
      }
      if (gw_use_fused(geom)) {
        // one launch per layer: neighbour aggregation (tcgen05) and gated conv / channel mixing fused in shared memory
        GwFusedArgs fa{};
        fa.zin = a.zin; fa.zout = a.zout; fa.Tin = a.Tin; fa.Tout = a.Tout; fa.dil = a.dil; fa.N = N;
        fa.has_in_bn = a.has_in_bn; fa.collect_stats = a.collect_stats; fa.in_scale = a.in_scale; fa.in_shift = a.in_shift;
        fa.w = a.w; fa.f = a.f; fa.g = a.g; fa.geom = geom; fa.sums = a.sums;
        fa.drop_thr = a.drop_thr; fa.drop_scale = a.drop_scale; fa.key = a.key;
        for (int s = 0; s < 3; ++s) {
          fa.q[s] = a.q[s]; fa.a[s] = stash + p.off_a[i][s];
          fa.img[s] = m.img[s]; fa.img_bstride[s] = m.img_bstride[s];
        }
        if (gw_fused_tc_mix(geom)) {
          if ((rc = allow_smem(gw_fused_fwd_kernel<true>, 227 * 1024))) return rc;
          gw_fused_fwd_kernel<true><<<dim3((a.Tout + GWF_NT - 1) / GWF_NT, B), GWF_THREADS, gwf_smem_bytes(geom, true), st>>>(fa);
        } else {
          if ((rc = allow_smem(gw_fused_fwd_kernel<false>, 227 * 1024))) return rc;
          gw_fused_fwd_kernel<false><<<dim3((a.Tout + GWF_NT - 1) / GWF_NT, B), GWF_THREADS, gwf_smem_bytes(geom, false), st>>>(fa);
        }
        STEP_LAUNCH_CHECK("gw_fused_fwd_kernel");
      } else {
      gw_layer_fwd_kernel<1><<<dim3(a.Tout, B), GW_THREADS, fwd_smem_bytes(N), st>>>(a);
      STEP_LAUNCH_CHECK("gw_layer_fwd_kernel[conv]");
      for (int s = 0; s < 3; ++s) { m.Y[s] = stash + p.off_a[i][s]; m.out[s] = stash + p.off_M3[s]; }
      if ((rc = tc_mix_launch(m, st))) return rc;
      gw_layer_fwd_kernel<2><<<dim3(a.Tout, B), GW_THREADS, fwd_smem_bytes(N), st>>>(a);
      STEP_LAUNCH_CHECK("gw_layer_fwd_kernel[q]");
      for (int s = 0; s < 3; ++s) { m.Y[s] = a.q[s]; m.out[s] = stash + p.off_O3[s]; }
      if ((rc = tc_mix_launch(m, st))) return rc;
      gw_layer_fwd_kernel<3><<<dim3(a.Tout, B), GW_THREADS, fwd_smem_bytes(N), st>>>(a);
      STEP_LAUNCH_CHECK("gw_layer_fwd_kernel[out]");
      }
    }
    if (a.has_gcn && training) {
      bn_finalize_kernel<<<1, 32, 0, st>>>(a.sums, (double)B * a.Tout * N, Lp[i].bn_w, Lp[i].bn_b, bn_stats + (size_t)i * 128);
      STEP_LAUNCH_CHECK("bn_finalize_kernel");
    }
  }
  {
    GwSkipArgs k{};
    k.L = n_layers; k.N = N; k.B = B; k.skip = skip_out;
    for (int i = 0; i < n_layers; ++i) {
      k.Tout[i] = p.Tout[i]; k.f[i] = stash + p.off_f[i]; k.g[i] = stash + p.off_g[i];
      k.skip_w[i] = Lp[i].skip_w; k.skip_b[i] = Lp[i].skip_b;
    }
    gw_skip_fwd_kernel<<<dim3((N + SKIP_NT - 1) / SKIP_NT, B), GSKIP, 0, st>>>(k);
    STEP_LAUNCH_CHECK("gw_skip_fwd_kernel");
  }
  return STEP_OK;
}

extern "C" int step_gwnet_stack_bwd(const float *dskip, const float *x0, const float *P1, const float *P2, const float *P3,
                                    const float *P1t, const float *P2t, const float *P3t,
                                    const step_gw_layer_params *Lp, const step_gw_layer_grads *Lg, int n_layers, int B,
                                    int N, float drop_p, unsigned long long seed, const float *bn_stats, float *stash,
                                    float *dx0, float *dP1, float *dP2, float *dP3, void *stream) {
  STEP_REQUIRE(dskip && x0 && P1 && P2 && P3 && P1t && P2t && P3t && Lp && Lg && bn_stats && stash && dx0 && dP1 && dP2 && dP3,
               "gwnet_bwd: null pointer");
  STEP_REQUIRE(B > 0 && B <= 21845 && N > 0 && n_layers >= 1 && n_layers <= GW_MAX_LAYERS, "gwnet_bwd: bad shape");
  int rc = gw_prepare(N);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const GwPlan p = make_plan(B, N, n_layers);
  const size_t nn = (size_t)N * N;
  cudaError_t e = cudaMemsetAsync(dP1, 0, (size_t)B * nn * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dP2, 0, (size_t)B * nn * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dP3, 0, nn * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(stash + p.off_sums_bwd, 0, (size_t)n_layers * 2 * 32 * sizeof(double), st);
  for (int i = 0; i < n_layers && e == cudaSuccess; ++i) {
    const step_gw_layer_grads &g = Lg[i];
    STEP_REQUIRE(g.filter_w && g.filter_b && g.gate_w && g.gate_b && g.skip_w && g.skip_b, "gwnet_bwd: null grad pointer");
    e = cudaMemsetAsync(g.filter_w, 0, 2048 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g.gate_w, 0, 2048 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g.filter_b, 0, 32 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g.gate_b, 0, 32 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g.skip_w, 0, GSKIP * 32 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g.skip_b, 0, GSKIP * sizeof(float), st);
    if (Lp[i].mlp_w != nullptr) {
      STEP_REQUIRE(g.mlp_w && g.mlp_b && g.bn_w && g.bn_b, "gwnet_bwd: null grad pointer");
      if (e == cudaSuccess) e = cudaMemsetAsync(g.mlp_w, 0, 32 * 224 * sizeof(float), st);
      if (e == cudaSuccess) e = cudaMemsetAsync(g.mlp_b, 0, 32 * sizeof(float), st);
    }
  }
  if (e != cudaSuccess) return fail_msg((int)e, cudaGetErrorString(e));

  uint32_t thr = 0; float dscale = 1.f;
  if (drop_p > 0.f) { thr = drop_threshold(drop_p); dscale = 1.f / (1.f - drop_p); }
  float *coef = stash + p.off_coef;
  double *bsums = reinterpret_cast<double *>(stash + p.off_sums_bwd);
  const bool use_tc = gw_use_tc(N);
  const MixGeom geom = mix_geom(N);
  const size_t img_floats = (mix_images_bytes(N) + 3) / 4;

  {
    // skip path of every layer in one launch: dus_l = Wk_l^T dskip, dWk_l, dbk_l
    GwSkipArgs k{};
    k.L = n_layers; k.N = N; k.B = B; k.dskip = dskip; k.dus = stash + p.off_dus;
    for (int i = 0; i < n_layers; ++i) {
      k.Tout[i] = p.Tout[i]; k.f[i] = stash + p.off_f[i]; k.g[i] = stash + p.off_g[i];
      k.skip_w[i] = Lp[i].skip_w; k.dskip_w[i] = Lg[i].skip_w; k.dskip_b[i] = Lg[i].skip_b;
    }
    const size_t smem_skip = ((size_t)GSKIP * GC + (size_t)N * GC) * sizeof(float);
    if ((rc = allow_smem(gw_skip_bwd_kernel, 227 * 1024))) return rc;
    gw_skip_bwd_kernel<<<dim3(n_layers, B), GSKIP, smem_skip, st>>>(k);
    STEP_LAUNCH_CHECK("gw_skip_bwd_kernel");
  }

  for (int i = n_layers - 1; i >= 0; --i) {
    const bool has_gcn = (Lp[i].mlp_w != nullptr);
    GwBwdArgs a{};
    a.Tin = p.Tin[i]; a.Tout = p.Tout[i]; a.dil = p.dil[i]; a.N = N; a.has_gcn = has_gcn;
    a.drnext = stash + p.off_DR[(i + 1) & 1];
    a.z = stash + p.off_z[i];
    a.stats = bn_stats + (size_t)i * 128;
    a.coef = coef + (size_t)i * 128;
    a.Pt[0] = P1t; a.Pt[1] = P2t; a.Pt[2] = P3t;
    a.pstride[0] = (long long)nn; a.pstride[1] = (long long)nn; a.pstride[2] = 0;
    a.w = Lp[i]; a.gr = Lg[i];
    a.f = stash + p.off_f[i]; a.g = stash + p.off_g[i];
    a.dus = stash + p.off_dus + (size_t)i * B * p.col;
    a.U = stash + p.off_U; a.DH = stash + p.off_DH; a.DZC = stash + p.off_DZC;
    for (int s = 0; s < 3; ++s) { a.DQ[s] = stash + p.off_DQ[s]; a.A[s] = stash + p.off_A[s]; }
    a.DA = stash + p.off_DA; a.DU = stash + p.off_DU; a.DPF = stash + p.off_DPF; a.DPG = stash + p.off_DPG;
    a.drop_thr = thr; a.drop_scale = dscale; a.key = rng_key(seed, 0x100u + i);
    if (!use_tc || !has_gcn) {
      gw_layer_bwd_kernel<0><<<dim3(a.Tout, B), GW_THREADS, bwd_smem_bytes(N), st>>>(a);
      STEP_LAUNCH_CHECK("gw_layer_bwd_kernel");
    } else {
      TcMixArgs m{};
      m.g = geom; m.B = B; m.T = a.Tout; m.nsup = 3; m.transposed_type = 0;
      for (int s = 0; s < 3; ++s) {
        m.img[s] = reinterpret_cast<const uint8_t *>(stash + p.off_img[s]);
        m.img_bstride[s] = (s < 2) ? (long long)img_floats * 4 : 0;
        a.DA3[s] = stash + p.off_DA3[s];
      }
      gw_layer_bwd_kernel<1><<<dim3(a.Tout, B), GW_THREADS, bwd_smem_bytes(N), st>>>(a);
      STEP_LAUNCH_CHECK("gw_layer_bwd_kernel[pre]");
      for (int s = 0; s < 3; ++s) { m.Y[s] = stash + p.off_DH; m.out[s] = stash + p.off_DQ[s]; }
      if ((rc = tc_mix_launch(m, st))) return rc;
      for (int s = 0; s < 3; ++s) { m.Y[s] = stash + p.off_DQ[s]; m.out[s] = stash + p.off_DA3[s]; }
      if ((rc = tc_mix_launch(m, st))) return rc;
      gw_layer_bwd_kernel<2><<<dim3(a.Tout, B), GW_THREADS, bwd_smem_bytes(N), st>>>(a);
      STEP_LAUNCH_CHECK("gw_layer_bwd_kernel[post]");
    }

    if (has_gcn) {
      GwDpArgs d{};
      d.N = N; d.T = p.Tout[i]; d.B = B;
      // a_s: stashed by the forward on the tensor-core path, recomputed into scratch by the fused CUDA-core backward
      for (int s = 0; s < 3; ++s) {
        d.Q[s] = stash + p.off_q[i][s]; d.A[s] = stash + (use_tc ? p.off_a[i][s] : p.off_A[s]); d.DQ[s] = stash + p.off_DQ[s];
      }
      d.DH = stash + p.off_DH;
      d.dP[0] = dP1; d.dP[1] = dP2; d.dP[2] = dP3;
      d.pstride[0] = (long long)nn; d.pstride[1] = (long long)nn; d.pstride[2] = 0;
      if (use_tc && tc_dP_supported(N)) {
        TcDpArgs td{};
        for (int s = 0; s < 3; ++s) {
          td.Q[s] = d.Q[s]; td.A[s] = d.A[s]; td.DQ[s] = d.DQ[s]; td.dP[s] = d.dP[s]; td.pstride[s] = d.pstride[s];
        }
        td.DH = d.DH; td.B = B; td.T = d.T; td.N = N;
        if ((rc = tc_dP_launch(td, st))) return rc;
      } else {
        gw_dP_kernel<<<dim3((N + 63) / 64, (N + 63) / 64, 3 * B), 256, 0, st>>>(d);
        STEP_LAUNCH_CHECK("gw_dP_kernel");
      }
    }

    GwBwdInArgs c{};
    c.Tin = p.Tin[i]; c.Tout = p.Tout[i]; c.dil = p.dil[i]; c.N = N; c.has_gcn = has_gcn; c.has_in_bn = (i > 0);
    c.zin = (i == 0) ? x0 : stash + p.off_z[i - 1];
    c.in_stats = (i > 0) ? bn_stats + (size_t)(i - 1) * 128 : nullptr;
    c.DPF = stash + p.off_DPF; c.DPG = stash + p.off_DPG; c.DZC = stash + p.off_DZC;
    c.w = Lp[i]; c.gr = Lg[i];
    c.drin = (i == 0) ? dx0 : stash + p.off_DR[i & 1];
    c.sums = (i > 0) ? bsums + (size_t)(i - 1) * 64 : nullptr;
    gw_layer_bwd_in_kernel<<<dim3(c.Tin, B), GW_THREADS, bwd_in_smem_bytes(N), st>>>(c);
    STEP_LAUNCH_CHECK("gw_layer_bwd_in_kernel");

    if (i > 0) {
      bn_bwd_finalize_kernel<<<1, 32, 0, st>>>(bsums + (size_t)(i - 1) * 64, (double)B * p.Tout[i - 1] * N, Lp[i - 1].bn_w,
                                              bn_stats + (size_t)(i - 1) * 128, coef + (size_t)(i - 1) * 128,
                                              Lg[i - 1].bn_w, Lg[i - 1].bn_b);
      STEP_LAUNCH_CHECK("bn_bwd_finalize_kernel");
    }
  }
  return STEP_OK;
}

extern "C" int step_gwnet_dropout_probe(const float *x, long long rows, float drop_p, unsigned long long seed, int layer, float *y,
                                        void *stream) {
  STEP_REQUIRE(x && y && rows > 0 && drop_p > 0.f && drop_p < 1.f && layer >= 0, "gwnet_dropout_probe: bad argument");
  gw_dropout_probe_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      x, rows, drop_threshold(drop_p), 1.f / (1.f - drop_p), rng_key(seed, 0x100u + (unsigned)layer), y);
  return check_launch("gw_dropout_probe_kernel");
}
