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

// STEP_B200_TRUNK_TC=0 keeps the CUDA-core conv2 forward
inline bool trunk_use_tc() {
  const char *e = getenv("STEP_B200_TRUNK_TC");
  return !(e && e[0] == '0');
}

}  // namespace stepk
