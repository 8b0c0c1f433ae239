// Conv1d(8 -> 16, k = 10) of the discrete-graph-learning trunk on tcgen05 (forward), included by trunk.cu.
// Reference: step/step_arch/discrete_graph_learning.py:131-133 (conv2 applied to bn1(relu(conv1(x))) of the whole training
// series, [N, 8, L1] -> [N, 16, L2], L2 = L1 - 9 ~ 24 k positions per node at METR-LA).
//
// Implicit GEMM without an im2col copy.  y1n = BN1(relu(conv1 x)) is rebuilt per tile (80 MACs per position, as in the
// CUDA-core kernel it replaces) and stored position-major: plane[pos][8 channels] bf16 = one 16-byte row per position, in
// a hi and a lo plane (split-bf16: hi*hi + lo*hi + hi*lo, ~2^-17 relative).  With the K index ordered (tap, channel), the
// K-chunk of tap j for output position l IS the plane row l + j: an UMMA K-major descriptor with
//     start = plane + (l0 + 2 kk) * 16,   SBO = 128 (8 consecutive positions),   LBO = 16 (next tap = next row)
// addresses the [128 positions x 16 K] operand of K-step kk (taps 2kk, 2kk+1) in place - consecutive K-chunks overlap in
// memory by design.  B = W2 as [10 taps][16 out channels][8 in channels] bf16 images (hi / lo, 2.5 KB each).
//   D[128 positions, 16 channels] += A[128, 80] B[80, 16]:  5 K-steps x 3 split products = 15 MMAs per 128 positions,
//   each bounded by its 4 KB operand read from shared memory (32 cycles) -> ~1024 MAC / clk / SM.
// CTA (persistent, 2 per SM): tiles of 1024 output positions of one node; warp 1 issues the MMAs, warps 2-9 rebuild the
// planes of tile t+1 and run the epilogue of tile t-1 (bias, ReLU, y2 in [N][16][L2] for the Linear, BN2 batch sums)
// while tile t is on the tensor core; two plane stages, two 128-column accumulator stages.
#pragma once
#include "tc_common.cuh"

namespace stepk {

constexpr int TCV_TP = 1024;                     // output positions per tile (8 MMA row tiles)
constexpr int TCV_ROWS = TCV_TP + 16;            // plane rows per stage (1024 + 9 halo, rounded to 8)
constexpr int TCV_THREADS = 320;
constexpr uint32_t TCV_PLANE = TCV_ROWS * 16;    // bytes of one plane
constexpr uint32_t TCV_WIMG = 10 * 16 * 16;      // bytes of one weight image

struct TcConv2Args {
  const float *x;            // [N][L0]
  const float *w1, *b1;      // conv1 [8][10], [8]
  const float *bn1;          // [4][8]: mean, var, scale, shift
  const float *w2, *b2;      // conv2 [16][8][10], [16]
  float *y2;                 // [N][16][L2] = relu(conv2(...)), pre-BN2
  double *sums2;             // [2][16] batch sums (null in eval mode)
  int N, L0, L1, L2, tiles_per_node;
};

static size_t tcv_smem_bytes() { return 2 * 2 * (size_t)TCV_PLANE + 2 * (size_t)TCV_WIMG + 16 * 8 + (80 + 8 + 8 + 8 + 16) * 4 + 64; }

__global__ void __launch_bounds__(TCV_THREADS, 2) trunk_conv2_tc_fwd_kernel(TcConv2Args a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t tcv_smem[];
  uint8_t *sPl = tcv_smem;                                  // [stage][hi, lo][TCV_ROWS][16 B]
  uint8_t *sWh = sPl + 4 * TCV_PLANE, *sWl = sWh + TCV_WIMG;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sWl + TCV_WIMG);
  uint64_t *built = bars, *consumed = bars + 2, *acc_full = bars + 4, *acc_empty = bars + 6;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
  float *sw1 = reinterpret_cast<float *>(bars + 10);        // w1[80], b1[8], scale[8], shift[8], b2[16]
  float *sb1 = sw1 + 80, *ssc = sb1 + 8, *ssh = ssc + 8, *sb2 = ssh + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // one-time: conv1 / BN1 / bias tables and the two conv2 weight images [tap][co][ci]
  for (int i = threadIdx.x; i < 80; i += blockDim.x) sw1[i] = a.w1[i];
  if (threadIdx.x < 8) { sb1[threadIdx.x] = a.b1[threadIdx.x]; ssc[threadIdx.x] = a.bn1[16 + threadIdx.x]; ssh[threadIdx.x] = a.bn1[24 + threadIdx.x]; }
  if (threadIdx.x < 16) sb2[threadIdx.x] = a.b2[threadIdx.x];
  for (int i = threadIdx.x; i < 10 * 16; i += blockDim.x) {
    const int tap = i / 16, co = i - tap * 16;
    float hi[8], lo[8];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const float w = a.w2[(co * 8 + ci) * 10 + tap];
      const float h = __bfloat162float(__float2bfloat16_rn(w));
      hi[ci] = h; lo[ci] = w - h;
    }
    reinterpret_cast<uint4 *>(sWh)[i] = pack8_bf16(hi);
    reinterpret_cast<uint4 *>(sWl)[i] = pack8_bf16(lo);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&built[i], 8); mbar_init(&consumed[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int total = a.N * a.tiles_per_node;
  const int ntiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 16, 0, 0);
      const uint32_t wh = smem_u32(sWh), wl = smem_u32(sWl);
      for (int it = 0; it < ntiles; ++it) {
        const int s = it & 1, use = it >> 1;
        mbar_wait(&built[s], use & 1);
        mbar_wait(&acc_empty[s], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t ph = smem_u32(sPl + (size_t)s * 2 * TCV_PLANE), pl = ph + TCV_PLANE;
        for (int rt = 0; rt < TCV_TP / 128; ++rt) {
          const uint32_t d = tmem + s * 128 + rt * 16;
#pragma unroll
          for (int kk = 0; kk < 5; ++kk) {
            const uint32_t ao = (uint32_t)(rt * 128 + 2 * kk) * 16u, bo = (uint32_t)(2 * kk) * 256u;
            const uint64_t ah = umma_desc(ph + ao, 16, 128), al = umma_desc(pl + ao, 16, 128);
            const uint64_t bh = umma_desc(wh + bo, 256, 128), bl = umma_desc(wl + bo, 256, 128);
            umma_bf16(d, ah, bh, idesc, kk != 0 ? 1u : 0u);
            umma_bf16(d, al, bh, idesc, 1u);
            umma_bf16(d, ah, bl, idesc, 1u);
          }
        }
        umma_commit(&consumed[s]);
        umma_commit(&acc_full[s]);
      }
    }
  } else if (warp >= 2) {
    const int wt = threadIdx.x - 64;                 // 0..255
    const int q = warp & 3, grp = (warp - 2) >> 2;   // TMEM lane quadrant; epilogue group (row tiles grp, grp+2, ...)
    float sacc[16], qacc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { sacc[c] = 0.f; qacc[c] = 0.f; }
    auto tile_of = [&](int it, int &n, int &t0) {
      const int g = blockIdx.x + it * gridDim.x;
      n = g / a.tiles_per_node;
      t0 = (g - n * a.tiles_per_node) * TCV_TP;
    };
    auto build = [&](int it) {
      const int s = it & 1, use = it >> 1;
      int n, t0;
      tile_of(it, n, t0);
      mbar_wait(&consumed[s], (use & 1) ^ 1);
      const float *xr = a.x + (size_t)n * a.L0;
      uint4 *ph = reinterpret_cast<uint4 *>(sPl + (size_t)s * 2 * TCV_PLANE), *pl = ph + TCV_ROWS;
      for (int j = wt; j < TCV_ROWS; j += 256) {
        const int p = t0 + j;                        // y1 position
        float hi[8], lo[8];
        if (p < a.L1) {
          float xv[10];
#pragma unroll
          for (int k = 0; k < 10; ++k) xv[k] = __ldg(xr + p + k);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float v = sb1[c];
#pragma unroll
            for (int k = 0; k < 10; ++k) v = fmaf(sw1[c * 10 + k], xv[k], v);
            v = fmaf(fmaxf(v, 0.f), ssc[c], ssh[c]);
            const float h = __bfloat162float(__float2bfloat16_rn(v));
            hi[c] = h; lo[c] = v - h;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) { hi[c] = 0.f; lo[c] = 0.f; }
        }
        ph[j] = pack8_bf16(hi);
        pl[j] = pack8_bf16(lo);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&built[s]);
    };
    auto epilogue = [&](int it) {
      const int s = it & 1, use = it >> 1;
      int n, t0;
      tile_of(it, n, t0);
      mbar_wait(&acc_full[s], use & 1);
      tc_fence_after();
      float *yo = a.y2 + (size_t)n * 16 * a.L2;
      for (int rt = grp; rt < TCV_TP / 128; rt += 2) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + s * 128 + rt * 16, v);
        const int p = t0 + rt * 128 + q * 32 + lane;
        if (p < a.L2) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float o = fmaxf(v[c] + sb2[c], 0.f);
            yo[(size_t)c * a.L2 + p] = o;
            sacc[c] += o;
            qacc[c] = fmaf(o, o, qacc[c]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);
    };
    if (ntiles > 0) build(0);
    for (int it = 0; it < ntiles; ++it) {
      if (it + 1 < ntiles) build(it + 1);
      epilogue(it);
    }
    if (a.sums2 != nullptr) {
      // CTA-level reduction through the (now dead) plane memory, then 32 double atomics per CTA
      float *red = reinterpret_cast<float *>(sPl);             // [8 warps][32]
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float s1 = warp_sum(sacc[c]), s2 = warp_sum(qacc[c]);
        if (lane == 0) { red[(warp - 2) * 32 + c] = s1; red[(warp - 2) * 32 + 16 + c] = s2; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w * 32 + lane];
        atomicAdd(a.sums2 + lane, (double)t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

inline int trunk_conv2_tc_launch(const TcConv2Args &a0, cudaStream_t st) {
  TcConv2Args a = a0;
  a.tiles_per_node = (a.L2 + TCV_TP - 1) / TCV_TP;
  int dev = 0, sms = 148, rc;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = tcv_smem_bytes();
  if ((rc = allow_smem(trunk_conv2_tc_fwd_kernel, smem))) return rc;
  const long long total = (long long)a.N * a.tiles_per_node;
  const int grid = (int)(total < 2LL * sms ? total : 2LL * sms);
  trunk_conv2_tc_fwd_kernel<<<grid, TCV_THREADS, smem, st>>>(a);
  return check_launch("trunk_conv2_tc_fwd_kernel");
}

// ===========================================================================
// Backward through BN2 -> ReLU -> conv2 on tcgen05 (replaces trunk_conv2_bwd_kernel): d(y1n) [N,8,L1], dW2, db2 and the
// BatchNorm-1 backward sums, per tile of 1024 y1 positions p of one node.
//   dpre2[co][l] = [y2 > 0] g2 rstd2 (dy2n - S1/M - xhat2 S2/M)   rebuilt per tile for l in [t0 - 16, t0 + 1024) and
//   stored position-major as two half-planes (channels 0-7 / 8-15), each in a hi and a lo bf16 plane (16 B per position).
//   d(y1n)[ci][p] = sum_{j, co} dpre2[co][p - 9 + j] w2[co][ci][9 - j]:   D[128 p, 16 (8 ci + 8 zero)] with K = (tap j, co):
//       A K-step j = chunks (j, half 0), (j, half 1) = rows l = p - 9 + j of the two half-planes (LBO = plane stride, SBO = 128),
//       B = w2 re-ordered to [20 chunks][16 rows][8 co]: 10 K-steps x 3 split products per 128 positions.
//   dW2[co][ci][tap] = sum_l dpre2[co][l] y1n[ci][l + tap]:   D[128 (80 valid) m = (tap, ci), 16 co], K = positions l:
//       A = the y1n plane read MN-major (element (m, l) = plane row l + tap, channel ci: M-chunk stride 16 B = next tap,
//       K-group stride 128 B), B = the dpre2 half-planes read MN-major (N-chunk stride = plane stride); one accumulator
//       for the whole CTA, 64 K-steps x 3 split products per tile, read once at the end.
// Same pipeline as the forward: warp 1 issues, warps 2-9 build the planes of tile t+1 and run the d(y1n) epilogue of tile t-1.
// ===========================================================================
constexpr uint32_t TCV_WBIMG = 20 * 16 * 16;     // bytes of one re-ordered w2 image of the backward
constexpr int TCVB_WORKERS = 512;                // 16 plane-builder / epilogue warps (the kernel is bound by their load latency)
constexpr int TCVB_THREADS = TCVB_WORKERS + 64;

struct TcConv2BwdArgs {
  const float *x, *w1, *b1, *bn1;   // bn1 [4][8]: mean, var, scale, shift
  float eps;
  const float *w2;                  // [16][8][10]
  const float *dy2n, *y2;           // [N][16][L2]
  const float *coef2;               // [5][16]: g2 rstd2, S1/M, S2/M, mean2, rstd2
  float *dy1n;                      // [N][8][L1]
  float *dw2, *db2;                 // accumulated with atomics (zeroed by the caller)
  double *sums1;                    // [2][8]: sum d(y1n), sum d(y1n) xhat1
  int N, L0, L1, L2, tiles_per_node;
};

static size_t tcvb_smem_bytes() { return 2 * 6 * (size_t)TCV_PLANE + 2 * (size_t)TCV_WBIMG + 16 * 8 + (80 + 8 * 5 + 16 * 5) * 4 + 64; }

__global__ void __launch_bounds__(TCVB_THREADS, 1) trunk_conv2_tc_bwd_kernel(TcConv2BwdArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t tcvb_smem[];
  // stage layout: [dp hi h0][dp hi h1][dp lo h0][dp lo h1][y1n hi][y1n lo], TCV_PLANE bytes each
  uint8_t *sPl = tcvb_smem;
  uint8_t *sWh = sPl + 12 * TCV_PLANE, *sWl = sWh + TCV_WBIMG;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sWl + TCV_WBIMG);
  uint64_t *built = bars, *consumed = bars + 2, *acc_full = bars + 4, *acc_empty = bars + 6, *dw_full = bars + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
  float *sw1 = reinterpret_cast<float *>(bars + 10);        // w1[80], b1[8], scale1[8], shift1[8], mean1[8], rstd1[8], coef2[80]
  float *sb1 = sw1 + 80, *ssc = sb1 + 8, *ssh = ssc + 8, *smean = ssh + 8, *srstd = smean + 8, *scf = srstd + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 80; i += blockDim.x) { sw1[i] = a.w1[i]; scf[i] = a.coef2[i]; }
  if (threadIdx.x < 8) {
    const int c = threadIdx.x;
    sb1[c] = a.b1[c]; ssc[c] = a.bn1[16 + c]; ssh[c] = a.bn1[24 + c];
    smean[c] = a.bn1[c]; srstd[c] = 1.0f / sqrtf(a.bn1[8 + c] + a.eps);
  }
  // w2 re-ordered for d(y1n): chunk (j, half) row n = ci (rows 8..15 zero), elements e = co - 8 half: w2[co][ci][9 - j]
  for (int i = threadIdx.x; i < 20 * 16; i += blockDim.x) {
    const int chunk = i / 16, n = i - chunk * 16, j = chunk >> 1, half = chunk & 1;
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float w = n < 8 ? a.w2[((8 * half + e) * 8 + n) * 10 + (9 - j)] : 0.f;
      const float h = __bfloat162float(__float2bfloat16_rn(w));
      hi[e] = h; lo[e] = w - h;
    }
    reinterpret_cast<uint4 *>(sWh)[i] = pack8_bf16(hi);
    reinterpret_cast<uint4 *>(sWl)[i] = pack8_bf16(lo);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&built[i], TCVB_WORKERS / 32); mbar_init(&consumed[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], TCVB_WORKERS / 32); }
    mbar_init(dw_full, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int total = a.N * a.tiles_per_node;
  const int ntiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_dx = umma_idesc_bf16(128, 16, 0, 0), idesc_dw = umma_idesc_bf16(128, 16, 1, 1);
      const uint32_t wh = smem_u32(sWh), wl = smem_u32(sWl);
      for (int it = 0; it < ntiles; ++it) {
        const int s = it & 1, use = it >> 1;
        mbar_wait(&built[s], use & 1);
        mbar_wait(&acc_empty[s], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t base = smem_u32(sPl + (size_t)s * 6 * TCV_PLANE);
        const uint32_t dph = base, dpl = base + 2 * TCV_PLANE, yh = base + 4 * TCV_PLANE, yl = base + 5 * TCV_PLANE;
        // d(y1n): 8 row tiles x 10 taps
        for (int rt = 0; rt < TCV_TP / 128; ++rt) {
          const uint32_t d = tmem + s * 128 + rt * 16;
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const uint32_t ao = (uint32_t)(rt * 128 + j + 7) * 16u, bo = (uint32_t)(2 * j) * 256u;   // row of l = p - 9 + j
            const uint64_t ah = umma_desc(dph + ao, TCV_PLANE, 128), al = umma_desc(dpl + ao, TCV_PLANE, 128);
            const uint64_t bh = umma_desc(wh + bo, 256, 128), bl = umma_desc(wl + bo, 256, 128);
            umma_bf16(d, ah, bh, idesc_dx, j != 0 ? 1u : 0u);
            umma_bf16(d, al, bh, idesc_dx, 1u);
            umma_bf16(d, ah, bl, idesc_dx, 1u);
          }
        }
        // dW2: K = the tile's 1024 output positions l = t0 + 16 ks + (0..15); dpre2 plane row of l is 16 + l - t0 (the
        // planes start at l = t0 - 16 so that these MN-major operand starts stay 128-byte aligned)
        for (int ks = 0; ks < TCV_TP / 16; ++ks) {
          const uint32_t ao = (uint32_t)(ks * 16) * 16u, bo = (uint32_t)(16 + ks * 16) * 16u;
          const uint64_t ah = umma_desc(yh + ao, 128, 16), al = umma_desc(yl + ao, 128, 16);
          const uint64_t bh = umma_desc(dph + bo, 128, TCV_PLANE), bl = umma_desc(dpl + bo, 128, TCV_PLANE);
          umma_bf16(tmem + 256, ah, bh, idesc_dw, (it | ks) != 0 ? 1u : 0u);
          umma_bf16(tmem + 256, al, bh, idesc_dw, 1u);
          umma_bf16(tmem + 256, ah, bl, idesc_dw, 1u);
        }
        umma_commit(&consumed[s]);
        umma_commit(&acc_full[s]);
      }
      umma_commit(dw_full);
    }
  } else if (warp >= 2) {
    const int wt = threadIdx.x - 64;                 // 0..TCVB_WORKERS-1
    const int q = warp & 3, grp = (warp - 2) >> 2;
    float s1acc[8], s2acc[8], dbacc[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) { s1acc[c] = 0.f; s2acc[c] = 0.f; }
#pragma unroll
    for (int c = 0; c < 16; ++c) dbacc[c] = 0.f;
    auto tile_of = [&](int it, int &n, int &t0) {
      const int g = blockIdx.x + it * gridDim.x;
      n = g / a.tiles_per_node;
      t0 = (g - n * a.tiles_per_node) * TCV_TP;
    };
    auto y1_raw = [&](const float *xr, int p, float *y /*[8]*/) {      // relu(conv1 x) at y1 position p < L1
      float xv[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) xv[k] = __ldg(xr + p + k);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = sb1[c];
#pragma unroll
        for (int k = 0; k < 10; ++k) v = fmaf(sw1[c * 10 + k], xv[k], v);
        y[c] = fmaxf(v, 0.f);
      }
    };
    auto build = [&](int it) {
      const int s = it & 1, use = it >> 1;
      int n, t0;
      tile_of(it, n, t0);
      mbar_wait(&consumed[s], (use & 1) ^ 1);
      uint4 *pl = reinterpret_cast<uint4 *>(sPl + (size_t)s * 6 * TCV_PLANE);
      const float *xr = a.x + (size_t)n * a.L0;
      const size_t nb = (size_t)n * 16 * a.L2;
      for (int r = wt; r < TCV_ROWS; r += TCVB_WORKERS) {
        // ---- dpre2 at l = t0 - 16 + r ----
        const int l = t0 - 16 + r;
        float hi[16], lo[16];
        if (l >= 0 && l < a.L2) {
#pragma unroll
          for (int co = 0; co < 16; ++co) {
            const float yv = __ldg(a.y2 + nb + (size_t)co * a.L2 + l);
            float v = 0.f;
            if (yv > 0.f) {
              const float xhat = (yv - scf[48 + co]) * scf[64 + co];
              v = scf[co] * (__ldg(a.dy2n + nb + (size_t)co * a.L2 + l) - scf[16 + co] - xhat * scf[32 + co]);
            }
            if (r >= 16) dbacc[co] += v;                   // every l is owned by exactly one tile: l in [t0, t0 + 1024)
            const float h = __bfloat162float(__float2bfloat16_rn(v));
            hi[co] = h; lo[co] = v - h;
          }
        } else {
#pragma unroll
          for (int co = 0; co < 16; ++co) { hi[co] = 0.f; lo[co] = 0.f; }
        }
        pl[r] = pack8_bf16(hi);
        pl[TCV_ROWS + r] = pack8_bf16(hi + 8);
        pl[2 * TCV_ROWS + r] = pack8_bf16(lo);
        pl[3 * TCV_ROWS + r] = pack8_bf16(lo + 8);
        // ---- y1n at position t0 + r ----
        const int p = t0 + r;
        float yh[8], yl[8];
        if (p < a.L1) {
          float y[8];
          y1_raw(xr, p, y);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float v = fmaf(y[c], ssc[c], ssh[c]);
            const float h = __bfloat162float(__float2bfloat16_rn(v));
            yh[c] = h; yl[c] = v - h;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) { yh[c] = 0.f; yl[c] = 0.f; }
        }
        pl[4 * TCV_ROWS + r] = pack8_bf16(yh);
        pl[5 * TCV_ROWS + r] = pack8_bf16(yl);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&built[s]);
    };
    auto epilogue = [&](int it) {
      const int s = it & 1, use = it >> 1;
      int n, t0;
      tile_of(it, n, t0);
      mbar_wait(&acc_full[s], use & 1);
      tc_fence_after();
      const float *xr = a.x + (size_t)n * a.L0;
      float *o = a.dy1n + (size_t)n * 8 * a.L1;
      for (int rt = grp; rt < TCV_TP / 128; rt += TCVB_WORKERS / 128) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + s * 128 + rt * 16, v);
        const int p = t0 + rt * 128 + q * 32 + lane;
        if (p < a.L1) {
          float y[8];
          y1_raw(xr, p, y);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            o[(size_t)c * a.L1 + p] = v[c];
            s1acc[c] += v[c];
            s2acc[c] = fmaf(v[c], (y[c] - smean[c]) * srstd[c], s2acc[c]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);
    };
    if (ntiles > 0) build(0);
    for (int it = 0; it < ntiles; ++it) {
      if (it + 1 < ntiles) build(it + 1);
      epilogue(it);
    }
    // ---- dW2 accumulator -> atomics (rows m = 8 tap + ci < 80), db2 and the BN1 sums through shared memory ----
    mbar_wait(dw_full, 0);
    tc_fence_after();
    if (grp == 0 && ntiles > 0) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + 256, v);
      const int m = q * 32 + lane;
      if (m < 80) {
        const int tap = m >> 3, ci = m & 7;
#pragma unroll
        for (int co = 0; co < 16; ++co) atomicAdd(a.dw2 + (co * 8 + ci) * 10 + tap, v[co]);
      }
    }
    float *red = reinterpret_cast<float *>(sPl);               // [worker warps][32]: 16 db2 + 8 s1 + 8 s2
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float t = warp_sum(dbacc[c]);
      if (lane == 0) red[(warp - 2) * 32 + c] = t;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float t1 = warp_sum(s1acc[c]), t2 = warp_sum(s2acc[c]);
      if (lane == 0) { red[(warp - 2) * 32 + 16 + c] = t1; red[(warp - 2) * 32 + 24 + c] = t2; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(TCVB_WORKERS) : "memory");
    if (warp == 2) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < TCVB_WORKERS / 32; ++w) t += red[w * 32 + lane];
      if (lane < 16) atomicAdd(a.db2 + lane, t);
      else atomicAdd(a.sums1 + (lane - 16), (double)t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

inline int trunk_conv2_tc_bwd_launch(const TcConv2BwdArgs &a0, cudaStream_t st) {
  TcConv2BwdArgs a = a0;
  a.tiles_per_node = (a.L1 + TCV_TP - 1) / TCV_TP;
  int dev = 0, sms = 148, rc;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = tcvb_smem_bytes();
  if ((rc = allow_smem(trunk_conv2_tc_bwd_kernel, smem))) return rc;
  const long long total = (long long)a.N * a.tiles_per_node;
  const int grid = (int)(total < (long long)sms ? total : (long long)sms);
  trunk_conv2_tc_bwd_kernel<<<grid, TCVB_THREADS, smem, st>>>(a);
  return check_launch("trunk_conv2_tc_bwd_kernel");
}

// STEP_B200_TRUNK_TC=0 keeps the CUDA-core conv2 forward and backward
inline bool trunk_use_tc() {
  const char *e = getenv("STEP_B200_TRUNK_TC");
  return !(e && e[0] == '0');
}

}  // namespace stepk
