// General fp32-in / fp32-out GEMM on tcgen05 with split-bf16 operands (fp32-class accuracy), for the mid-size
// dense products of the STEP path that the reference leaves to cuBLAS through torch.nn.Linear / Conv2d(1x1) / matmul:
//   Graph WaveNet epilogue  fc_his (96->512->256), end_conv_1 (256->512), end_conv_2 (512->12) and their backward
//                           (step/step_arch/graphwavenet/model.py:215-220),
//   discrete graph learning  the two [N,100] x [100,100] halves of fc_out (discrete_graph_learning.py:148-151),
//   TSFormer pre-training    backward GEMMs of the masked auto-encoder (tsformer.py:71-160).
//
//   C[M,N] (+)= alpha * sum_k opA(m,k) opB(n,k)  [+ bias[n]]  [epilogue]
//   opA: transA == 0 -> A is row-major [M][K] (K contiguous);  transA == 1 -> A is row-major [K][M]
//   opB: transB == 0 -> B is row-major [N][K] (a Linear weight); transB == 1 -> B is row-major [K][N]
//
// Each fp32 operand tile is read by the CTA's worker warps (full-line coalesced), split into bf16 hi + lo and stored
// as UMMA no-swizzle canonical images ([8-element group][row][8]; K-major when the contraction runs along the
// contiguous dimension, MN-major otherwise - the same bytes, only the descriptor changes); one thread issues
// hi*hi + hi*lo + lo*hi per k-step into a [128 x 128] fp32 TMEM accumulator.  3-stage mbarrier ring, 2 CTAs per SM.
// Split-K (grid.z) accumulates with fp32 atomics into a zero-initialised / pre-existing C.
#include "common.cuh"
#include "tc_common.cuh"

namespace stepk {
using namespace tc;

enum { GE_NONE = 0, GE_RELU = 1, GE_MASK = 2, GE_RELU_ADD_RELU = 3 };

struct GemmArgs {
  const float *A, *B;
  long long lda, ldb;
  int transA, transB;
  int M, N, K;
  float *C;
  long long ldc;
  const float *bias;       // [N] or null
  const float *aux;        // GE_MASK: C = acc * (aux > 0);  GE_RELU_ADD_RELU: C = relu(relu(acc + bias) + aux)
  long long ldaux;
  float *aux_out;          // GE_RELU_ADD_RELU: relu(acc + bias) (same leading dimension as C), may be null
  int epi, accumulate;
  float alpha;
};

constexpr int GM_THREADS = 320;       // warp 0 idle, warp 1 MMA issuer, warps 2-9 operand builders (2-5 also epilogue)
constexpr int GM_STAGES = 3;
constexpr int GM_KC = 32;
constexpr uint32_t GM_IMG = 512 * 16;            // one (hi or lo) image of one operand: 512 units
constexpr uint32_t GM_STAGE = 4 * GM_IMG;        // A hi | A lo | B hi | B lo

__device__ __forceinline__ void gm_split8(const float *x, uint4 &hi, uint4 &lo) {
  float h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
    l[j] = x[j] - h[j];
  }
  hi = pack8_bf16(h);
  lo = pack8_bf16(l);
}

// 8 consecutive elements of row `row` starting at column `col` of a row-major matrix with `nrows` x `ncols` valid
__device__ __forceinline__ void gm_load_unit(const float *__restrict__ base, long long ld, long long row, long long col,
                                             long long nrows, long long ncols, bool vec_ok, float *x) {
  if (row >= nrows || col >= ncols) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    return;
  }
  const float *p = base + row * ld + col;
  if (vec_ok && col + 8 <= ncols) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (col + j < ncols) ? __ldg(p + j) : 0.f;
  }
}

__global__ void __launch_bounds__(GM_THREADS, 2) tc_gemm_kernel(GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t gm_smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(gm_smem + GM_STAGES * GM_STAGE);
  uint64_t *built = bars, *consumed = bars + GM_STAGES, *d_full = bars + 2 * GM_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128;
  const int ncols = min(128, a.N - n0);
  const int nmma = (ncols + 15) / 16 * 16;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GM_STAGES; ++i) { mbar_init(&built[i], 8); mbar_init(&consumed[i], 1); }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int total = (a.K + GM_KC - 1) / GM_KC;
  const int s0 = (int)((long long)total * blockIdx.z / gridDim.z), s1 = (int)((long long)total * (blockIdx.z + 1) / gridDim.z);
  const int nst = s1 - s0;

  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, nmma, a.transA, a.transB);
      for (int i = 0; i < nst; ++i) {
        const int st = i % GM_STAGES;
        mbar_wait(&built[st], (i / GM_STAGES) & 1);
        tc_fence_after();
        const uint32_t ah = smem_u32(gm_smem + st * GM_STAGE), al = ah + GM_IMG, bh = al + GM_IMG, bl = bh + GM_IMG;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          // K-major image: [4 k-groups][128 rows][16 B]; MN-major image: [16 mn-groups][32 k rows][16 B]
          const uint64_t dah = a.transA ? umma_desc(ah + kk * 256, 128, 512) : umma_desc(ah + kk * 4096, 2048, 128);
          const uint64_t dal = a.transA ? umma_desc(al + kk * 256, 128, 512) : umma_desc(al + kk * 4096, 2048, 128);
          const uint64_t dbh = a.transB ? umma_desc(bh + kk * 256, 128, 512) : umma_desc(bh + kk * 4096, 2048, 128);
          const uint64_t dbl = a.transB ? umma_desc(bl + kk * 256, 128, 512) : umma_desc(bl + kk * 4096, 2048, 128);
          umma_bf16(tmem, dah, dbh, idesc, (i | kk) != 0 ? 1u : 0u);
          umma_bf16(tmem, dah, dbl, idesc, 1u);
          umma_bf16(tmem, dal, dbh, idesc, 1u);
        }
        umma_commit(&consumed[st]);
      }
      umma_commit(d_full);
    }
  } else if (warp >= 2) {
    const int ww = warp - 2;
    const bool vecA = (a.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    const bool vecB = (a.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.B) & 15) == 0);
    for (int i = 0; i < nst; ++i) {
      const int st = i % GM_STAGES;
      const long long k0 = (long long)(s0 + i) * GM_KC;
      mbar_wait(&consumed[st], ((i / GM_STAGES) & 1) ^ 1);
      uint8_t *base = gm_smem + st * GM_STAGE;
      float x[4][8];
      // four warp-blocks per worker warp: blocks 0-15 -> A, 16-31 -> B; all loads are issued before the first split
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int blk = ww + 8 * (p & 1), isB = p >> 1;
        const int trans = isB ? a.transB : a.transA;
        const float *src = isB ? a.B : a.A;
        const long long ld = isB ? a.ldb : a.lda;
        const long long mn0 = isB ? n0 : m0, mnN = isB ? a.N : a.M;
        if (!trans) {   // rows = mn (16 blocks of 8), groups = k (4)
          const int r = blk * 8 + (lane & 7), g = lane >> 3;
          gm_load_unit(src, ld, mn0 + r, k0 + g * 8, mnN, a.K, isB ? vecB : vecA, x[p]);
        } else {        // rows = k (4 blocks of 8), groups = mn (4 quads of 4)
          const int r = (blk & 3) * 8 + (lane & 7), g = (blk >> 2) * 4 + (lane >> 3);
          gm_load_unit(src, ld, k0 + r, mn0 + g * 8, a.K, mnN, isB ? vecB : vecA, x[p]);
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int blk = ww + 8 * (p & 1), isB = p >> 1;
        const int trans = isB ? a.transB : a.transA;
        uint32_t idx;
        if (!trans) idx = (uint32_t)(lane >> 3) * 128 + blk * 8 + (lane & 7);
        else idx = (uint32_t)((blk >> 2) * 4 + (lane >> 3)) * 32 + (blk & 3) * 8 + (lane & 7);
        uint4 hi, lo;
        gm_split8(x[p], hi, lo);
        uint8_t *img = base + (isB ? 2 * GM_IMG : 0);
        reinterpret_cast<uint4 *>(img)[idx] = hi;
        reinterpret_cast<uint4 *>(img + GM_IMG)[idx] = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&built[st]);
    }
    if (warp < 6) {
      // ---- epilogue: thread = output row ----
      mbar_wait(d_full, 0);
      tc_fence_after();
      const int q = warp & 3, m = m0 + q * 32 + lane;
      const bool row_ok = m < a.M;
      const bool atomic = gridDim.z > 1;
      float *crow = a.C + (size_t)(row_ok ? m : 0) * a.ldc + n0;
      const float *auxrow = a.aux ? a.aux + (size_t)(row_ok ? m : 0) * a.ldaux + n0 : nullptr;
      float *aorow = a.aux_out ? a.aux_out + (size_t)(row_ok ? m : 0) * a.ldc + n0 : nullptr;
      for (int c0 = 0; c0 < nmma; c0 += 32) {
        float t[32];
        if (nst == 0) {
#pragma unroll
          for (int xx = 0; xx < 32; ++xx) t[xx] = 0.f;
        } else if (c0 + 32 <= nmma) {
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, t);
        } else {
          float t16[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, t16);
#pragma unroll
          for (int xx = 0; xx < 16; ++xx) { t[xx] = t16[xx]; t[16 + xx] = 0.f; }
        }
        if (!row_ok) continue;
#pragma unroll
        for (int xx = 0; xx < 32; ++xx) {
          const int c = c0 + xx;
          if (c < ncols) {
            float v = t[xx] * a.alpha;
            if (a.bias && blockIdx.z == 0) v += a.bias[n0 + c];
            if (a.epi == GE_RELU) v = fmaxf(v, 0.f);
            else if (a.epi == GE_MASK) v = auxrow[c] > 0.f ? v : 0.f;
            else if (a.epi == GE_RELU_ADD_RELU) {
              v = fmaxf(v, 0.f);
              if (aorow) aorow[c] = v;
              v = fmaxf(v + auxrow[c], 0.f);
            }
            if (atomic) atomicAdd(crow + c, v);
            else if (a.accumulate) crow[c] += v;
            else crow[c] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

// column sums (bias gradients): out[n] = sum_m x[m][n]
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ x, long long M, int N, long long ld,
                                                     float *__restrict__ out) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  float s = 0.f;
  if (n < N)
    for (long long m = (long long)blockIdx.y * 8 + ty; m < M; m += 8LL * gridDim.y) s += x[m * ld + n];
  red[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    atomicAdd(out + n, t);
  }
}

// dz = dy * (y > 0)   (ReLU backward where the mask cannot ride on a GEMM epilogue)
__global__ void relu_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ y, long long n, float *__restrict__ dz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dz[i] = y[i] > 0.f ? dy[i] : 0.f;
}

}  // namespace stepk

using namespace stepk;

extern "C" int step_gemm_f32(const float *A, long long lda, int transA, const float *B, long long ldb, int transB, int M, int N,
                             int K, float alpha, const float *bias, int epilogue, const float *aux, long long ldaux,
                             float *aux_out, int accumulate, int ksplit, float *C, long long ldc, void *stream) {
  STEP_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gemm_f32: bad argument");
  STEP_REQUIRE(epilogue >= GE_NONE && epilogue <= GE_RELU_ADD_RELU, "gemm_f32: bad epilogue");
  STEP_REQUIRE((epilogue != GE_MASK && epilogue != GE_RELU_ADD_RELU) || aux, "gemm_f32: this epilogue needs aux");
  STEP_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? N : K) && ldc >= N, "gemm_f32: leading dimension too small");
  if (ksplit < 1) ksplit = 1;
  const int total = (K + GM_KC - 1) / GM_KC;
  if (ksplit > total) ksplit = total;
  STEP_REQUIRE(ksplit == 1 || (epilogue == GE_NONE), "gemm_f32: split-K supports no epilogue (fp32 atomics into C)");
  GemmArgs a{};
  a.A = A; a.B = B; a.lda = lda; a.ldb = ldb; a.transA = transA ? 1 : 0; a.transB = transB ? 1 : 0;
  a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.bias = bias; a.aux = aux; a.ldaux = ldaux; a.aux_out = aux_out;
  a.epi = epilogue; a.accumulate = accumulate; a.alpha = alpha;
  const size_t smem = GM_STAGES * (size_t)GM_STAGE + 8 * 8 + 16;
  int rc = allow_smem(tc_gemm_kernel, smem);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (ksplit > 1 && !accumulate) {
    if (ldc == N) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st);
    else cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st);
  }
  tc_gemm_kernel<<<dim3((N + 127) / 128, (M + 127) / 128, ksplit), GM_THREADS, smem, st>>>(a);
  return check_launch("tc_gemm_kernel");
}

extern "C" int step_colsum_f32(const float *x, long long M, int N, long long ld, float *out, void *stream) {
  STEP_REQUIRE(x && out && M > 0 && N > 0 && ld >= N, "colsum_f32: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st);
  long long gy = (M + 255) / 256;
  if (gy > 64) gy = 64;
  if (gy < 1) gy = 1;
  colsum_kernel<<<dim3((N + 31) / 32, (unsigned)gy), 256, 0, st>>>(x, M, N, ld, out);
  return check_launch("colsum_kernel");
}

extern "C" int step_relu_bwd_f32(const float *dy, const float *y, long long n, float *dz, void *stream) {
  STEP_REQUIRE(dy && y && dz && n > 0, "relu_bwd_f32: bad argument");
  relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, y, n, dz);
  return check_launch("relu_bwd_kernel");
}
