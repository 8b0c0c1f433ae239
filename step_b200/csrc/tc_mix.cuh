// Node mixing (the diffusion step of the graph convolution) on tcgen05 with fp32-level accuracy.
//
//   out[b,t,w,c] = sum_v Amat[b,w,v] * Y[b,t,v,c]         (Amat = P^T for the forward hop, P for the backward hop)
//
// Both operands are split into bf16 high + low parts (x = hi + lo, |lo| <= 2^-9 |x|) and the product is
// accumulated in fp32 in TMEM as hi*hi + hi*lo + lo*hi, which leaves a relative error of ~2^-17 per term:
// the result is indistinguishable from the fp32 CUDA-core mix at the 1e-4 parity bar, so this path serves
// both precisions (graphwavenet/model.py:10-16 `nconv`, reference fp32 einsum).
//
//   A operand: per support and sample, bf16 K-major images [hi|lo][KC chunks][Mpad rows][8] built once per step
//              (tc_support_images_kernel) for P and P^T; streamed through a TMA ring in K slices of 32.
//   B operand: Y is read as fp32, split and written to smem by the CTA itself as MN-major images
//              [NT*4 channel groups][Kpad rows][8] (hi and lo) - NT time steps share one pass over A.
//   One CTA = (time group, support x sample, 128-row tile of output nodes).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace stepk {

constexpr int MIX_STAGES = 3;
constexpr int MIX_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 operand builders + epilogue

struct MixGeom {
  int N, Kpad, KC, Mpad, MT, NT;   // nodes, padded K (multiple of 32), chunks, padded rows, row tiles, time steps per CTA
  int MPC;                         // row tiles handled by one CTA (2 when they fit: the B images are then built once)
  size_t img_bytes;                // one (hi or lo) image
};

inline MixGeom mix_geom(int N) {
  MixGeom g{};
  g.N = N;
  g.Kpad = (N + 31) / 32 * 32;
  g.KC = g.Kpad / 8;
  g.Mpad = (N + 127) / 128 * 128;
  g.MT = g.Mpad / 128;
  g.NT = g.Kpad <= 256 ? 4 : (g.Kpad <= 480 ? 2 : 1);
  g.img_bytes = (size_t)g.KC * g.Mpad * 16;
  g.MPC = (g.MT % 2 == 0 && g.NT * 32 * 2 <= 256) ? 2 : 1;
  return g;
}
// images of one matrix: [P^T-type hi][P^T-type lo][P-type hi][P-type lo]
inline size_t mix_images_bytes(int N) { return 4 * mix_geom(N).img_bytes; }
inline size_t mix_smem_bytes(const MixGeom &g) {
  return (size_t)MIX_STAGES * 2 * 4 * 2048 * g.MPC + 2 * (size_t)g.NT * 4 * g.Kpad * 16 + 16 * 8 + 16;
}

struct TcMixArgs {
  const uint8_t *img[3];      // per support: image set base (see mix_images_bytes)
  long long img_bstride[3];   // bytes between samples (0: shared across the batch)
  int transposed_type;        // 1: use the P^T-type images (forward hop), 0: the P-type images (backward hop)
  const float *Y[3];          // [B,T,N,32]
  float *out[3];              // [B,T,N,32]
  int B, T, nsup;
  MixGeom g;
};

// P [nb][N][N] fp32 -> four bf16 images per matrix (zero padded): type 0 rows = w, K = v holds P[v][w] (P^T);
// type 1 rows = v, K = w holds P[v][w].
__global__ void tc_support_images_kernel(const float *__restrict__ P, int nb, MixGeom g, uint8_t *__restrict__ out,
                                         long long out_bstride) {
  const long long unit = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte unit of one type
  const long long per = (long long)g.KC * g.Mpad;
  if (unit >= per * 2 * nb) return;
  const int b = (int)(unit / (per * 2));
  const long long r0 = unit - (long long)b * per * 2;
  const int type = (int)(r0 / per);
  const long long u = r0 - (long long)type * per;
  const int chunk = (int)(u / g.Mpad), row = (int)(u - (long long)chunk * g.Mpad);
  const float *Pb = P + (size_t)b * g.N * g.N;
  float hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = chunk * 8 + j;
    float v = 0.f;
    if (row < g.N && k < g.N) v = (type == 0) ? Pb[(size_t)k * g.N + row] : Pb[(size_t)row * g.N + k];
    const float h = __bfloat162float(__float2bfloat16_rn(v));
    hi[j] = h;
    lo[j] = v - h;
  }
  uint8_t *base = out + (size_t)b * out_bstride + (size_t)type * 2 * g.img_bytes;
  reinterpret_cast<uint4 *>(base)[u] = tc::pack8_bf16(hi);
  reinterpret_cast<uint4 *>(base + g.img_bytes)[u] = tc::pack8_bf16(lo);
}

__global__ void __launch_bounds__(MIX_THREADS, 1) tc_mix_kernel(TcMixArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t mix_smem[];
  uint8_t *smem = mix_smem;
  const MixGeom g = a.g;
  const int tg = blockIdx.x, sb = blockIdx.y, mt0 = blockIdx.z * g.MPC;
  const int s = sb / a.B, b = sb - s * a.B;
  const int t0 = tg * g.NT, nt = min(g.NT, a.T - t0);
  const int Ncols = nt * 32;
  const uint32_t rows = 128u * g.MPC;                   // A rows staged per chunk
  const uint32_t half_bytes = 4 * rows * 16;            // one of (hi, lo): 4 chunks
  const uint32_t stage_bytes = 2 * half_bytes;
  const uint32_t bimg = (uint32_t)g.NT * 4 * g.Kpad * 16;
  uint8_t *sA = smem;
  uint8_t *sBh = smem + MIX_STAGES * stage_bytes, *sBl = sBh + bimg;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sBl + bimg);
  uint64_t *full = bars, *empty = bars + MIX_STAGES, *b_ready = bars + 2 * MIX_STAGES, *d_full = b_ready + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MIX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(b_ready, 8);
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nslices = g.Kpad / 32;
  const uint8_t *img_hi = a.img[s] + (size_t)b * a.img_bstride[s] + (a.transposed_type ? 0 : 2 * g.img_bytes);
  const uint8_t *img_lo = img_hi + g.img_bytes;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nslices; ++i) {
        const int st = i % MIX_STAGES, ph = (i / MIX_STAGES) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], stage_bytes);
        for (int c = 0; c < 4; ++c) {
          const size_t off = ((size_t)(i * 4 + c) * g.Mpad + (size_t)mt0 * 128) * 16;
          tma_bulk_g2s(sA + st * stage_bytes + c * rows * 16, img_hi + off, rows * 16, &full[st]);
          tma_bulk_g2s(sA + st * stage_bytes + half_bytes + c * rows * 16, img_lo + off, rows * 16, &full[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, Ncols, 0, 1);
      mbar_wait(b_ready, 0);
      tc_fence_after();
      const uint32_t bh = smem_u32(sBh), bl = smem_u32(sBl);
      const uint32_t b_sbo = (uint32_t)g.Kpad * 16;      // distance between 8-column groups
      const uint32_t a_lbo = rows * 16;                  // distance between K chunks
      for (int i = 0; i < nslices; ++i) {
        const int st = i % MIX_STAGES, ph = (i / MIX_STAGES) & 1;
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t ah = smem_u32(sA + st * stage_bytes), al = ah + half_bytes;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint32_t koff = (uint32_t)(i * 2 + kk) * 256;   // 16 K rows x 16 B
          const uint64_t dbh = umma_desc(bh + koff, 128, b_sbo), dbl = umma_desc(bl + koff, 128, b_sbo);
          for (int m = 0; m < g.MPC; ++m) {
            const uint64_t dah = umma_desc(ah + kk * 2 * a_lbo + m * 2048, a_lbo, 128);
            const uint64_t dal = umma_desc(al + kk * 2 * a_lbo + m * 2048, a_lbo, 128);
            const uint32_t d = tmem + m * 128;
            umma_bf16(d, dah, dbh, idesc, (i | kk) != 0 ? 1u : 0u);
            umma_bf16(d, dah, dbl, idesc, 1u);
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
        }
        umma_commit(&empty[st]);
      }
      umma_commit(d_full);
    }
  } else {
    // ---- build the split B images of Y[b, t0..t0+nt) : element (t, v, c) -> group t*4 + c/8, row v ----
    const int wt = threadIdx.x - 64;                      // 0..255
    const float *Yb = a.Y[s] + ((size_t)b * a.T + t0) * g.N * 32;
    const int units = nt * 4 * g.Kpad;                    // 16-byte units per image
    for (int u = wt; u < units; u += 256) {
      const int grp = u / g.Kpad, v = u - grp * g.Kpad;   // consecutive threads -> consecutive rows: conflict-free 16 B stores
      const int t = grp >> 2, cg = grp & 3;
      float hi[8], lo[8];
      if (v < g.N) {
        const float4 *src = reinterpret_cast<const float4 *>(Yb + ((size_t)t * g.N + v) * 32 + cg * 8);
        const float4 x0 = src[0], x1 = src[1];
        const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = __bfloat162float(__float2bfloat16_rn(x[j]));
          hi[j] = h;
          lo[j] = x[j] - h;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { hi[j] = 0.f; lo[j] = 0.f; }
      }
      reinterpret_cast<uint4 *>(sBh)[u] = pack8_bf16(hi);
      reinterpret_cast<uint4 *>(sBl)[u] = pack8_bf16(lo);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(b_ready);
    // ---- epilogue: TMEM -> out[b, t, w, :]; warps 2-5 take row tile 0, warps 6-9 row tile 1 ----
    const int q = warp & 3, m = (warp - 2) >> 2;
    mbar_wait(d_full, 0);
    tc_fence_after();
    if (m < g.MPC) {
      const int w = (mt0 + m) * 128 + q * 32 + lane;
      float *ob = a.out[s] + ((size_t)b * a.T + t0) * g.N * 32;
      for (int t = 0; t < nt; ++t) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + m * 128 + t * 32, v);
        if (w < g.N) {
          float *o = ob + ((size_t)t * g.N + w) * 32;
#pragma unroll
          for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4 *>(o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

inline int tc_mix_launch(const TcMixArgs &a, cudaStream_t st) {
  int rc = allow_smem(tc_mix_kernel, 227 * 1024);
  if (rc) return rc;
  const size_t smem = mix_smem_bytes(a.g);
  if (smem > 227 * 1024) return fail(STEP_EUNSUPPORTED, "tc_mix: N=%lld needs too much shared memory", a.g.N);
  dim3 grid((a.T + a.g.NT - 1) / a.g.NT, a.nsup * a.B, a.g.MT / a.g.MPC);
  tc_mix_kernel<<<grid, MIX_THREADS, smem, st>>>(a);
  return check_launch("tc_mix_kernel");
}

inline int tc_support_images_launch(const float *P, int nb, const MixGeom &g, uint8_t *out, long long bstride, cudaStream_t st) {
  const long long units = (long long)g.KC * g.Mpad * 2 * nb;
  tc_support_images_kernel<<<(unsigned)((units + 255) / 256), 256, 0, st>>>(P, nb, g, out, bstride);
  return check_launch("tc_support_images_kernel");
}


// ===========================================================================
// Dense support gradient on tcgen05 (the straight-through estimator needs dL/dP for EVERY edge):
//   dP_s[b][v][w] += sum_t sum_c ( q_s[b,t,v,c] dh[b,t,w,c] + a_s[b,t,v,c] dq_s[b,t,w,c] )
// an "NT" GEMM with K = 2*T*32 per (sample, support).  Operands are fp32 activations in [B,T,N,32]; the CTA
// splits them into bf16 hi/lo K-major images in shared memory itself (double buffered over t) and issues
// hi*hi + hi*lo + lo*hi.  One CTA = (128-row tile of v, support x sample).
// ===========================================================================
struct TcDpArgs {
  const float *Q[3], *A[3], *DQ[3], *DH;
  float *dP[3];
  long long pstride[3];       // 0: shared across the batch (adaptive adjacency) -> atomic accumulation
  int B, T, N;
};
constexpr int DP_THREADS = 576;          // warp 0 idle, warp 1 MMA issuer, 16 operand-builder / epilogue warps
constexpr int DP_BUILDERS = DP_THREADS - 64;

__global__ void __launch_bounds__(DP_THREADS, 1) tc_dP_kernel(TcDpArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t dp_smem[];
  const int mt = blockIdx.x, sb = blockIdx.y, s = sb / a.B, b = sb - s * a.B;
  // column tile of up to 256 target nodes (MMA N <= 256, 256 TMEM columns): grid.z tiles cover N > 256 (PEMS04/BAY/03/07)
  const int N = a.N, col0 = blockIdx.z * 256;
  const int Nb = min(256, (N - col0 + 15) / 16 * 16);
  const uint32_t a_img = 8 * 128 * 16, b_img = 8u * Nb * 16;      // 8 chunks (2 pairs x 32 channels) per time step
  const uint32_t stage_bytes = 2 * a_img + 2 * b_img;
  uint64_t *bars = reinterpret_cast<uint64_t *>(dp_smem + 2 * stage_bytes);
  uint64_t *built = bars, *consumed = bars + 2, *d_full = bars + 4;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&built[i], DP_BUILDERS / 32); mbar_init(&consumed[i], 1); }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const size_t col = (size_t)N * 32;

  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, Nb, 0, 0);
      for (int t = 0; t < a.T; ++t) {
        const int st = t & 1;
        mbar_wait(&built[st], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t ah = smem_u32(dp_smem + st * stage_bytes), al = ah + a_img, bh = al + a_img, bl = bh + b_img;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t dah = umma_desc(ah + kk * 2 * 2048, 2048, 128), dal = umma_desc(al + kk * 2 * 2048, 2048, 128);
          const uint64_t dbh = umma_desc(bh + kk * 2 * Nb * 16, Nb * 16, 128), dbl = umma_desc(bl + kk * 2 * Nb * 16, Nb * 16, 128);
          umma_bf16(tmem, dah, dbh, idesc, (t | kk) != 0 ? 1u : 0u);
          umma_bf16(tmem, dah, dbl, idesc, 1u);
          umma_bf16(tmem, dal, dbh, idesc, 1u);
        }
        umma_commit(&consumed[st]);
      }
      umma_commit(d_full);
    }
  } else if (warp >= 2) {
    const int wt = threadIdx.x - 64;       // 0..DP_BUILDERS-1
    for (int t = 0; t < a.T; ++t) {
      const int st = t & 1;
      mbar_wait(&consumed[st], ((t >> 1) & 1) ^ 1);
      uint8_t *base = dp_smem + st * stage_bytes;
      uint4 *Ah = reinterpret_cast<uint4 *>(base), *Al = reinterpret_cast<uint4 *>(base + a_img);
      uint4 *Bh = reinterpret_cast<uint4 *>(base + 2 * a_img), *Bl = reinterpret_cast<uint4 *>(base + 2 * a_img + b_img);
      const size_t toff = ((size_t)b * a.T + t) * col;
      // A: rows v = mt*128 + r; chunk = pair*4 + cg
      for (int u = wt; u < 8 * 128; u += DP_BUILDERS) {
        const int chunk = u >> 7, r = u & 127, pair = chunk >> 2, cg = chunk & 3, v = mt * 128 + r;
        float hi[8], lo[8];
        if (v < N) {
          const float *src = (pair ? a.A[s] : a.Q[s]) + toff + (size_t)v * 32 + cg * 8;
          const float4 x0 = reinterpret_cast<const float4 *>(src)[0], x1 = reinterpret_cast<const float4 *>(src)[1];
          const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float h = __bfloat162float(__float2bfloat16_rn(x[j])); hi[j] = h; lo[j] = x[j] - h; }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { hi[j] = 0.f; lo[j] = 0.f; }
        }
        Ah[u] = pack8_bf16(hi);
        Al[u] = pack8_bf16(lo);
      }
      // B: rows w; chunk = pair*4 + cg
      for (int u = wt; u < 8 * Nb; u += DP_BUILDERS) {
        const int chunk = u / Nb, wl = u - chunk * Nb, w = col0 + wl, pair = chunk >> 2, cg = chunk & 3;
        float hi[8], lo[8];
        if (w < N) {
          const float *src = (pair ? a.DQ[s] : a.DH) + toff + (size_t)w * 32 + cg * 8;
          const float4 x0 = reinterpret_cast<const float4 *>(src)[0], x1 = reinterpret_cast<const float4 *>(src)[1];
          const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float h = __bfloat162float(__float2bfloat16_rn(x[j])); hi[j] = h; lo[j] = x[j] - h; }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { hi[j] = 0.f; lo[j] = 0.f; }
        }
        Bh[u] = pack8_bf16(hi);
        Bl[u] = pack8_bf16(lo);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&built[st]);
    }
    // epilogue: TMEM -> smem tile [128][Nb+1] (warps 2-5, lane = row) -> coalesced read-modify-write of dP by all
    // eight warps (one warp per row, consecutive lanes = consecutive columns)
    float *tile = reinterpret_cast<float *>(dp_smem);       // the operand stages are dead once d_full has fired
    const int ldt = Nb + 1;
    mbar_wait(d_full, 0);
    tc_fence_after();
    if (warp < 6) {
      const int q = warp & 3, r = q * 32 + lane;
      for (int c0 = 0; c0 < Nb; c0 += 32) {
        float tv[32];
        if (c0 + 32 <= Nb) {
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, tv);
        } else {
          float t16[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, t16);
#pragma unroll
          for (int c = 0; c < 16; ++c) tv[c] = t16[c];
#pragma unroll
          for (int c = 16; c < 32; ++c) tv[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c0 + c < Nb) tile[r * ldt + c0 + c] = tv[c];
      }
    }
    // named barrier over the builder/epilogue warps; warps 0/1 do not take part
    asm volatile("bar.sync 1, %0;" ::"n"(DP_BUILDERS) : "memory");
    {
      const bool shared = a.pstride[s] == 0;
      float *dstb = a.dP[s] + (size_t)b * a.pstride[s];
      for (int r = warp - 2; r < 128; r += DP_BUILDERS / 32) {
        const int v = mt * 128 + r;
        if (v >= N) break;
        float *dst = dstb + (size_t)v * N + col0;
        const int ncols = min(Nb, N - col0);
        for (int c = lane; c < ncols; c += 32) {
          const float x = tile[r * ldt + c];
          if (shared) atomicAdd(dst + c, x);
          else dst[c] += x;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

inline bool tc_dP_supported(int N) { return N > 0; }

inline int tc_dP_launch(const TcDpArgs &a, cudaStream_t st) {
  int rc = allow_smem(tc_dP_kernel, 227 * 1024);
  if (rc) return rc;
  const int Nb = a.N >= 256 ? 256 : (a.N + 15) / 16 * 16;
  const size_t smem = 2 * (size_t)(2 * 8 * 128 * 16 + 2 * 8 * Nb * 16) + 8 * 8 + 16;
  tc_dP_kernel<<<dim3((a.N + 127) / 128, 3 * a.B, (a.N + 255) / 256), DP_THREADS, smem, st>>>(a);
  return check_launch("tc_dP_kernel");
}

}  // namespace stepk
