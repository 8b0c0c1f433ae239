// TSFormer encoder, forecasting mode - fp32 CUDA-core path ("parity precision").
//
// Reference semantics (file:line relative to the reference repo):
//   step/step_arch/tsformer/patch.py:31-42            patch embedding
//   step/step_arch/tsformer/positional_encoding.py:24-35
//   step/step_arch/tsformer/transformer_layers.py:13-20  4 x post-norm encoder layer
//   step/step_arch/tsformer/tsformer.py:86-105         encoding(mask=False) + encoder_norm
//
// Token layout: x[(b*N + n)*P + p][96]; the reference's [P, B*N, 96] sequence-first view is the
// same math (every sequence is independent).
#include "common.cuh"

namespace stepk {

thread_local char g_last_error[512] = {0};
unsigned long long g_launch_count = 0;

constexpr int D_MODEL = 96;
constexpr int PATCH = 12;
constexpr int HEADS = 4;
constexpr int HEAD_DIM = 24;

// ===========================================================================
// patch embedding + positional embedding + sqrt(d)
// ===========================================================================
// grid: (ceil(N/32), P, B); block: 32 nodes x 8 feature groups (12 features each)
__global__ void __launch_bounds__(256) ts_embed_kernel(const float *__restrict__ series, long long sB, long long sT,
                                                       long long sN, int N, int P, const float *__restrict__ w,
                                                       const float *__restrict__ bias, const float *__restrict__ pos,
                                                       float *__restrict__ x, uint32_t drop_thr, float drop_scale,
                                                       uint64_t key) {
  __shared__ float sw[D_MODEL * PATCH];
  __shared__ float sb[D_MODEL];
  __shared__ float sv[32][PATCH + 1];
  const int b = blockIdx.z, p = blockIdx.y, n0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  for (int i = tid; i < D_MODEL * PATCH; i += 256) sw[i] = w[i];
  if (tid < D_MODEL) sb[tid] = bias[tid] + pos[(size_t)p * D_MODEL + tid];
  // 32 nodes x 12 time steps; consecutive threads read consecutive nodes (stride sN)
  for (int i = tid; i < 32 * PATCH; i += 256) {
    int t = i / 32, nn = i % 32;
    int n = n0 + nn;
    sv[nn][t] = (n < N) ? series[b * sB + (long long)(p * PATCH + t) * sT + n * sN] : 0.f;
  }
  __syncthreads();
  const int nn = tid / 8, fg = tid % 8;
  const int n = n0 + nn;
  if (n >= N) return;
  float v[PATCH];
#pragma unroll
  for (int t = 0; t < PATCH; ++t) v[t] = sv[nn][t];
  const size_t row = ((size_t)(b * N + n) * P + p);
  float *dst = x + row * D_MODEL;
  const float scale = sqrtf((float)D_MODEL);
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    int f = fg + 8 * j;  // consecutive lanes -> consecutive features: coalesced 32 B pieces
    float acc = sb[f];
#pragma unroll
    for (int t = 0; t < PATCH; ++t) acc = fmaf(sw[f * PATCH + t], v[t], acc);
    if (drop_thr) {
      uint4 r = philox4x32(row * D_MODEL + f, key);
      acc = (r.x >= drop_thr) ? acc * drop_scale : 0.f;
    }
    dst[f] = acc * scale;
  }
}

// ===========================================================================
// C = A W^T + b  (+ epilogue)     A [M,K] row-major, W [Nout,K] row-major
// 128 x 96 block tile, BK = 16, 256 threads, 8 x 6 register tile per thread.
// ===========================================================================
constexpr int GBM = 128, GBN = 96, GBK = 16, GAS = GBM + 4;

enum { EPI_NONE = 0, EPI_RELU = 1, EPI_RES_LN = 2, EPI_COSINE = 3 };

struct GemmArgs {
  const float *A, *W, *bias;
  float *C;
  long long M;
  int K, Nout;
  long long strideA, strideW, strideC;  // per blockIdx.z (batched Gram)
  const float *residual, *ln_w, *ln_b, *ln2_w, *ln2_b;
  const float *norm_a, *norm_b;  // EPI_COSINE: row norms of A rows / W rows, per batch stride Nn
  long long strideNorm;
  uint32_t drop_thr;
  float drop_scale;
  uint64_t key;
};

template <int EPI>
__global__ void __launch_bounds__(256) gemm_tn_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[GBK][GAS];
  __shared__ __align__(16) float Ws[GBK][GBN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  const float *A = g.A + blockIdx.z * g.strideA;
  const float *W = g.W + blockIdx.z * g.strideW;
  float *C = g.C + blockIdx.z * g.strideC;
  const int K = g.K;

  float acc[8][6];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;

  // global -> register staging
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;  // rows a_row, a_row + 64
  float4 ra[2], rw[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      long long r = m0 + a_row + 64 * i;
      ra[i] = (r < g.M) ? *reinterpret_cast<const float4 *>(A + r * K + k0 + a_kq) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + 256 * i;
      if (idx < GBN * 4) {
        int n = n0 + (idx >> 2);
        rw[i] = (n < g.Nout) ? *reinterpret_cast<const float4 *>(W + (long long)n * K + k0 + (idx & 3) * 4)
                             : make_float4(0, 0, 0, 0);
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = a_row + 64 * i;
      As[a_kq + 0][r] = ra[i].x; As[a_kq + 1][r] = ra[i].y; As[a_kq + 2][r] = ra[i].z; As[a_kq + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + 256 * i;
      if (idx < GBN * 4) {
        int n = idx >> 2, kq = (idx & 3) * 4;
        Ws[kq + 0][n] = rw[i].x; Ws[kq + 1][n] = rw[i].y; Ws[kq + 2][n] = rw[i].z; Ws[kq + 3][n] = rw[i].w;
      }
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += GBK) {
    store_tiles();
    __syncthreads();
    if (k0 + GBK < K) load_tiles(k0 + GBK);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[k][ty * 8 + 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float w[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) w[j] = Ws[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long row = m0 + ty * 8 + i;
    const bool row_ok = row < g.M;
    float v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int col = n0 + tx + 16 * j;
      const bool ok = row_ok && col < g.Nout;
      float t = acc[i][j];
      if (EPI == EPI_COSINE) {
        if (ok) {
          const float na = g.norm_a[blockIdx.z * g.strideNorm + row] + 1e-7f;
          const float nb = g.norm_b[blockIdx.z * g.strideNorm + col] + 1e-7f;
          t = t / (na * nb);
        }
      } else {
        if (g.bias != nullptr && col < g.Nout) t += g.bias[col];
        if (EPI == EPI_RELU) t = fmaxf(t, 0.f);
        if (g.drop_thr && ok) {
          uint4 r = philox4x32((uint64_t)row * g.Nout + col, g.key);
          t = (r.x >= g.drop_thr) ? t * g.drop_scale : 0.f;
        }
        if (EPI == EPI_RES_LN && ok) t += g.residual[row * g.Nout + col];
      }
      v[j] = t;
    }
    if (EPI == EPI_RES_LN) {
      // the 96 columns of a row live in the 16 lanes that share ty (lane bits 0..3)
      auto ln = [&](const float *lw, const float *lb) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) s += v[j];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.f / 96.f);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) { float d = v[j] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = 1.0f / sqrtf(q * (1.f / 96.f) + 1e-5f);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int col = tx + 16 * j;
          v[j] = (v[j] - mean) * rstd * lw[col] + lb[col];
        }
      };
      ln(g.ln_w, g.ln_b);
      if (g.ln2_w != nullptr) ln(g.ln2_w, g.ln2_b);
    }
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int col = n0 + tx + 16 * j;
        if (col < g.Nout) C[row * g.Nout + col] = v[j];
      }
    }
  }
}

static int launch_gemm(GemmArgs &g, int epilogue, int batch, cudaStream_t st) {
  if (g.K % GBK != 0) return fail(STEP_EUNSUPPORTED, "linear: K=%lld must be a multiple of 16", g.K);
  if (g.M <= 0 || g.Nout <= 0) return fail(STEP_EINVAL, "linear: empty problem");
  dim3 grid((unsigned)((g.M + GBM - 1) / GBM), (unsigned)((g.Nout + GBN - 1) / GBN), (unsigned)batch);
  switch (epilogue) {
    case EPI_NONE: gemm_tn_kernel<EPI_NONE><<<grid, 256, 0, st>>>(g); break;
    case EPI_RELU: gemm_tn_kernel<EPI_RELU><<<grid, 256, 0, st>>>(g); break;
    case EPI_RES_LN:
      if (g.Nout != 96) return fail(STEP_EUNSUPPORTED, "linear: residual+LayerNorm epilogue needs Nout == 96");
      gemm_tn_kernel<EPI_RES_LN><<<grid, 256, 0, st>>>(g);
      break;
    case EPI_COSINE: gemm_tn_kernel<EPI_COSINE><<<grid, 256, 0, st>>>(g); break;
    default: return fail(STEP_EINVAL, "linear: unknown epilogue %lld", epilogue);
  }
  return check_launch("gemm_tn_kernel");
}

// ===========================================================================
// attention: one block per (sequence, head); two query rows per thread, K/V in smem,
// online softmax over chunks of 8 keys.
// ===========================================================================
constexpr int ATT_CHUNK = 8;

__global__ void attn_fwd_kernel(const float *__restrict__ qkv, float *__restrict__ out, int P, int Ppad,
                                uint32_t drop_thr, float drop_scale, uint64_t key) {
  extern __shared__ __align__(16) float smem[];
  float *Ks = smem;                         // [Ppad][24]
  float *Vs = smem + (size_t)Ppad * HEAD_DIM;  // [Ppad][24]
  const int s = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const float *base = qkv + (size_t)s * P * (3 * D_MODEL);
  for (int idx = tid; idx < Ppad * 6; idx += nt) {
    int r = idx / 6, part = idx % 6;
    float4 kv = make_float4(0, 0, 0, 0), vv = kv;
    if (r < P) {
      const float *rowp = base + (size_t)r * (3 * D_MODEL) + h * HEAD_DIM + part * 4;
      kv = *reinterpret_cast<const float4 *>(rowp + D_MODEL);
      vv = *reinterpret_cast<const float4 *>(rowp + 2 * D_MODEL);
    }
    *reinterpret_cast<float4 *>(Ks + r * HEAD_DIM + part * 4) = kv;
    *reinterpret_cast<float4 *>(Vs + r * HEAD_DIM + part * 4) = vv;
  }
  const int q0 = tid, q1 = tid + nt;
  const bool ok0 = q0 < P, ok1 = q1 < P;
  // scores are kept in the log2 domain: s' = (q . k) / sqrt(24) * log2(e)
  const float qscale = 0.20412414523193154f * 1.4426950408889634f;
  float qa[HEAD_DIM], qb[HEAD_DIM], oa[HEAD_DIM], ob[HEAD_DIM];
#pragma unroll
  for (int c = 0; c < HEAD_DIM; c += 4) {
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if (ok0) a = *reinterpret_cast<const float4 *>(base + (size_t)q0 * (3 * D_MODEL) + h * HEAD_DIM + c);
    if (ok1) b = *reinterpret_cast<const float4 *>(base + (size_t)q1 * (3 * D_MODEL) + h * HEAD_DIM + c);
    qa[c] = a.x * qscale; qa[c + 1] = a.y * qscale; qa[c + 2] = a.z * qscale; qa[c + 3] = a.w * qscale;
    qb[c] = b.x * qscale; qb[c + 1] = b.y * qscale; qb[c + 2] = b.z * qscale; qb[c + 3] = b.w * qscale;
    oa[c] = oa[c + 1] = oa[c + 2] = oa[c + 3] = 0.f;
    ob[c] = ob[c + 1] = ob[c + 2] = ob[c + 3] = 0.f;
  }
  float ma = -INFINITY, mb = -INFINITY, la = 0.f, lb = 0.f;
  __syncthreads();

  for (int j0 = 0; j0 < Ppad; j0 += ATT_CHUNK) {
    float sa[ATT_CHUNK], sb[ATT_CHUNK];
#pragma unroll
    for (int jj = 0; jj < ATT_CHUNK; ++jj) {
      const float4 *kr = reinterpret_cast<const float4 *>(Ks + (j0 + jj) * HEAD_DIM);
      float da = 0.f, db = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 6; ++c4) {
        float4 kk = kr[c4];
        da = fmaf(qa[4 * c4], kk.x, da); da = fmaf(qa[4 * c4 + 1], kk.y, da);
        da = fmaf(qa[4 * c4 + 2], kk.z, da); da = fmaf(qa[4 * c4 + 3], kk.w, da);
        db = fmaf(qb[4 * c4], kk.x, db); db = fmaf(qb[4 * c4 + 1], kk.y, db);
        db = fmaf(qb[4 * c4 + 2], kk.z, db); db = fmaf(qb[4 * c4 + 3], kk.w, db);
      }
      const bool valid = (j0 + jj) < P;
      sa[jj] = valid ? da : -INFINITY;
      sb[jj] = valid ? db : -INFINITY;
    }
    float mxa = ma, mxb = mb;
#pragma unroll
    for (int jj = 0; jj < ATT_CHUNK; ++jj) { mxa = fmaxf(mxa, sa[jj]); mxb = fmaxf(mxb, sb[jj]); }
    const float ca = exp2f(ma - mxa), cb = exp2f(mb - mxb);  // first chunk: exp2(-inf) = 0
    ma = mxa; mb = mxb;
    la *= ca; lb *= cb;
#pragma unroll
    for (int c = 0; c < HEAD_DIM; ++c) { oa[c] *= ca; ob[c] *= cb; }
    uint4 ra0, ra1, rb0, rb1;
    if (drop_thr) {
      const uint64_t ba = ((uint64_t)(s * HEADS + h) * P + q0) * Ppad + j0;
      const uint64_t bb = ((uint64_t)(s * HEADS + h) * P + q1) * Ppad + j0;
      ra0 = philox4x32(ba, key); ra1 = philox4x32(ba + 4, key);
      rb0 = philox4x32(bb, key); rb1 = philox4x32(bb + 4, key);
    }
#pragma unroll
    for (int jj = 0; jj < ATT_CHUNK; ++jj) {
      float pa = exp2f(sa[jj] - ma), pb = exp2f(sb[jj] - mb);
      la += pa; lb += pb;
      if (drop_thr) {
        const uint32_t wa = jj < 4 ? (&ra0.x)[jj] : (&ra1.x)[jj - 4];
        const uint32_t wb = jj < 4 ? (&rb0.x)[jj] : (&rb1.x)[jj - 4];
        pa = (wa >= drop_thr) ? pa : 0.f;
        pb = (wb >= drop_thr) ? pb : 0.f;
      }
      const float4 *vr = reinterpret_cast<const float4 *>(Vs + (j0 + jj) * HEAD_DIM);
#pragma unroll
      for (int c4 = 0; c4 < 6; ++c4) {
        float4 vv = vr[c4];
        oa[4 * c4] = fmaf(pa, vv.x, oa[4 * c4]); oa[4 * c4 + 1] = fmaf(pa, vv.y, oa[4 * c4 + 1]);
        oa[4 * c4 + 2] = fmaf(pa, vv.z, oa[4 * c4 + 2]); oa[4 * c4 + 3] = fmaf(pa, vv.w, oa[4 * c4 + 3]);
        ob[4 * c4] = fmaf(pb, vv.x, ob[4 * c4]); ob[4 * c4 + 1] = fmaf(pb, vv.y, ob[4 * c4 + 1]);
        ob[4 * c4 + 2] = fmaf(pb, vv.z, ob[4 * c4 + 2]); ob[4 * c4 + 3] = fmaf(pb, vv.w, ob[4 * c4 + 3]);
      }
    }
  }
  const float ia = drop_scale / la, ib = drop_scale / lb;
  if (ok0) {
    float *o = out + ((size_t)s * P + q0) * D_MODEL + h * HEAD_DIM;
#pragma unroll
    for (int c = 0; c < HEAD_DIM; c += 4)
      *reinterpret_cast<float4 *>(o + c) = make_float4(oa[c] * ia, oa[c + 1] * ia, oa[c + 2] * ia, oa[c + 3] * ia);
  }
  if (ok1) {
    float *o = out + ((size_t)s * P + q1) * D_MODEL + h * HEAD_DIM;
#pragma unroll
    for (int c = 0; c < HEAD_DIM; c += 4)
      *reinterpret_cast<float4 *>(o + c) = make_float4(ob[c] * ib, ob[c + 1] * ib, ob[c + 2] * ib, ob[c + 3] * ib);
  }
}

// ===========================================================================
// LayerNorm over 96 features, one warp per row
// ===========================================================================
__global__ void layernorm96_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                   const float *__restrict__ b, float *__restrict__ y, long long M) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float *xr = x + row * D_MODEL;
  float v0 = xr[lane], v1 = xr[lane + 32], v2 = xr[lane + 64];
  const float mean = warp_sum(v0 + v1 + v2) * (1.f / 96.f);
  const float d0 = v0 - mean, d1 = v1 - mean, d2 = v2 - mean;
  const float rstd = 1.0f / sqrtf(warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.f / 96.f) + 1e-5f);
  float *yr = y + row * D_MODEL;
  yr[lane] = d0 * rstd * w[lane] + b[lane];
  yr[lane + 32] = d1 * rstd * w[lane + 32] + b[lane + 32];
  yr[lane + 64] = d2 * rstd * w[lane + 64] + b[lane + 64];
}

static void drop_consts(float p, uint32_t &thr, float &scale) {
  if (p > 0.f) { thr = drop_threshold(p); scale = 1.0f / (1.0f - p); }
  else { thr = 0; scale = 1.0f; }
}

static int attn_launch(const float *qkv, float *out, int S, int P, float drop_p, uint64_t seed, uint32_t site,
                       cudaStream_t st) {
  const int Ppad = (P + ATT_CHUNK - 1) / ATT_CHUNK * ATT_CHUNK;
  const int threads = (((P + 1) / 2) + 31) / 32 * 32;
  if (threads > 1024) return fail(STEP_EUNSUPPORTED, "attention: P=%lld too long for one block", P);
  const size_t smem = (size_t)Ppad * HEAD_DIM * 2 * sizeof(float);
  if (smem > 220 * 1024) return fail(STEP_EUNSUPPORTED, "attention: P=%lld does not fit shared memory", P);
  if (smem > 48 * 1024) {
    int rc = allow_smem(attn_fwd_kernel, 220 * 1024);
    if (rc) return rc;
  }
  uint32_t thr; float scale;
  drop_consts(drop_p, thr, scale);
  attn_fwd_kernel<<<dim3(S, HEADS), threads, smem, st>>>(qkv, out, P, Ppad, thr, scale, rng_key(seed, site));
  return check_launch("attn_fwd_kernel");
}

}  // namespace stepk

using namespace stepk;

extern "C" int step_abi_version(void) { return STEP_B200_ABI_VERSION; }
extern "C" const char *step_last_error_string(void) { return g_last_error; }
extern "C" unsigned long long step_launch_count(void) { return g_launch_count; }
extern "C" int step_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail_msg((int)e, cudaGetErrorString(e));
  return STEP_OK;
}

extern "C" int step_ts_embed_fwd(const float *series, long long sB, long long sT, long long sN, int B, int N, int P,
                                 const float *patch_w, const float *patch_b, const float *pos, float *x, float drop_p,
                                 unsigned long long seed, void *stream) {
  STEP_REQUIRE(series && patch_w && patch_b && pos && x, "ts_embed: null pointer");
  STEP_REQUIRE(B > 0 && N > 0 && P > 0 && P <= 65535 && B <= 65535, "ts_embed: bad shape");
  uint32_t thr; float scale;
  drop_consts(drop_p, thr, scale);
  ts_embed_kernel<<<dim3((N + 31) / 32, P, B), 256, 0, (cudaStream_t)stream>>>(series, sB, sT, sN, N, P, patch_w, patch_b,
                                                                               pos, x, thr, scale, rng_key(seed, 1));
  return check_launch("ts_embed_kernel");
}

extern "C" int step_linear_f32(const float *A, const float *W, const float *bias, float *C, long long M, int K, int Nout,
                               int epilogue, const float *residual, const float *ln_w, const float *ln_b, float drop_p,
                               unsigned long long seed, unsigned drop_site, void *stream) {
  STEP_REQUIRE(A && W && C, "linear: null pointer");
  STEP_REQUIRE(epilogue == EPI_NONE || epilogue == EPI_RELU || epilogue == EPI_RES_LN, "linear: bad epilogue");
  if (epilogue == EPI_RES_LN) STEP_REQUIRE(residual && ln_w && ln_b, "linear: residual+LN epilogue needs its operands");
  GemmArgs g{};
  g.A = A; g.W = W; g.bias = bias; g.C = C; g.M = M; g.K = K; g.Nout = Nout;
  g.residual = residual; g.ln_w = ln_w; g.ln_b = ln_b;
  drop_consts(drop_p, g.drop_thr, g.drop_scale);
  g.key = rng_key(seed, drop_site);
  return launch_gemm(g, epilogue, 1, (cudaStream_t)stream);
}

extern "C" int step_attn_fwd_f32(const float *qkv, float *out, int S, int P, float drop_p, unsigned long long seed,
                                 unsigned drop_site, void *stream) {
  STEP_REQUIRE(qkv && out && S > 0 && P > 0, "attention: bad argument");
  return attn_launch(qkv, out, S, P, drop_p, seed, drop_site, (cudaStream_t)stream);
}

extern "C" int step_layernorm96_f32(const float *x, const float *w, const float *b, float *y, long long M, void *stream) {
  STEP_REQUIRE(x && w && b && y && M > 0, "layernorm: bad argument");
  layernorm96_kernel<<<(unsigned)((M + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, w, b, y, M);
  return check_launch("layernorm96_kernel");
}

// workspace per chunk of T tokens: X [T,96], X1 [T,96], O [T,96], QKV [T,288] (reused as H [T,384])
static size_t enc_ws_floats(size_t T) { return T * (96 * 3 + 384); }

extern "C" size_t step_ts_encoder_workspace_bytes(int chunk_seqs, int P) {
  return enc_ws_floats((size_t)chunk_seqs * P) * sizeof(float) + 256;
}

// the encoder layer stack on T = Sc*P tokens held in X [T,96] (already scaled by sqrt(d)); the result replaces X.
// fnw/fnb: optional final LayerNorm (encoder_norm / decoder_norm) fused into the last layer's epilogue.
static int run_layer_stack(float *X, int Sc, int P, const step_ts_layer_weights *L, int n_layers, const float *fnw,
                           const float *fnb, float *ws, size_t Tc_max, float drop_p, uint64_t cseed, cudaStream_t st) {
  float *X1 = ws, *O = X1 + Tc_max * 96, *X2 = O + Tc_max * 96, *QKV = X2 + Tc_max * 96;  // QKV/H share
  uint32_t thr; float dscale;
  drop_consts(drop_p, thr, dscale);
  const long long T = (long long)Sc * P;
  const float *cur = X;
  int rc;
  for (int l = 0; l < n_layers; ++l) {
    const step_ts_layer_weights &w = L[l];
    const uint32_t site = 16u * (l + 1);
    GemmArgs g{};
    // QKV = cur W_in^T + b_in
    g.A = cur; g.W = w.in_proj_w; g.bias = w.in_proj_b; g.C = QKV; g.M = T; g.K = 96; g.Nout = 288;
    g.drop_thr = 0; g.drop_scale = 1.f;
    if ((rc = launch_gemm(g, EPI_NONE, 1, st))) return rc;
    if ((rc = attn_launch(QKV, O, Sc, P, drop_p, cseed, site + 1, st))) return rc;
    // X1 = LN1(cur + drop(O W_o^T + b_o))
    g = GemmArgs{};
    g.A = O; g.W = w.out_proj_w; g.bias = w.out_proj_b; g.C = X1; g.M = T; g.K = 96; g.Nout = 96;
    g.residual = cur; g.ln_w = w.norm1_w; g.ln_b = w.norm1_b;
    g.drop_thr = thr; g.drop_scale = dscale; g.key = rng_key(cseed, site + 2);
    if ((rc = launch_gemm(g, EPI_RES_LN, 1, st))) return rc;
    // H = drop(relu(X1 W_1^T + b_1))
    float *H = QKV;
    g = GemmArgs{};
    g.A = X1; g.W = w.lin1_w; g.bias = w.lin1_b; g.C = H; g.M = T; g.K = 96; g.Nout = 384;
    g.drop_thr = thr; g.drop_scale = dscale; g.key = rng_key(cseed, site + 3);
    if ((rc = launch_gemm(g, EPI_RELU, 1, st))) return rc;
    // X2 = LN2(X1 + drop(H W_2^T + b_2)); the last layer also applies the final norm and lands in X
    const bool last = (l == n_layers - 1);
    float *dst = last ? X : X2;
    g = GemmArgs{};
    g.A = H; g.W = w.lin2_w; g.bias = w.lin2_b; g.C = dst; g.M = T; g.K = 384; g.Nout = 96;
    g.residual = X1; g.ln_w = w.norm2_w; g.ln_b = w.norm2_b;
    if (last && fnw != nullptr) { g.ln2_w = fnw; g.ln2_b = fnb; }
    g.drop_thr = thr; g.drop_scale = dscale; g.key = rng_key(cseed, site + 4);
    if ((rc = launch_gemm(g, EPI_RES_LN, 1, st))) return rc;
    cur = dst;
  }
  return STEP_OK;
}

extern "C" int step_ts_encoder_fwd(const float *series, long long sB, long long sT, long long sN, int B, int N, int P,
                                   const float *patch_w, const float *patch_b, const float *pos,
                                   const step_ts_layer_weights *L, int n_layers, const float *fnw, const float *fnb,
                                   float *hidden, void *workspace, size_t workspace_bytes, int chunk_seqs, float drop_p,
                                   unsigned long long seed, void *stream) {
  STEP_REQUIRE(series && patch_w && patch_b && pos && L && fnw && fnb && hidden && workspace, "ts_encoder: null pointer");
  STEP_REQUIRE(n_layers >= 1, "ts_encoder: needs at least one layer");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = B * N;
  if (chunk_seqs <= 0 || chunk_seqs > S) chunk_seqs = S;
  if (workspace_bytes < step_ts_encoder_workspace_bytes(chunk_seqs, P))
    return fail(STEP_EWORKSPACE, "ts_encoder: workspace too small (%lld bytes needed)",
                (long long)step_ts_encoder_workspace_bytes(chunk_seqs, P));
  // the embedding writes straight into `hidden` (it is [B,N,P,96] = the token layout) and every
  // chunk is then transformed in place through the workspace.
  int rc = step_ts_embed_fwd(series, sB, sT, sN, B, N, P, patch_w, patch_b, pos, hidden, drop_p, seed, stream);
  if (rc) return rc;
  const size_t Tc_max = (size_t)chunk_seqs * P;
  float *ws = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  for (int s0 = 0; s0 < S; s0 += chunk_seqs) {
    const int Sc = (S - s0 < chunk_seqs) ? (S - s0) : chunk_seqs;
    const uint64_t cseed = seed + 0x51ED27ull * (uint64_t)s0;  // decorrelate chunks
    if ((rc = run_layer_stack(hidden + (size_t)s0 * P * 96, Sc, P, L, n_layers, fnw, fnb, ws, Tc_max, drop_p, cseed, st)))
      return rc;
  }
  return STEP_OK;
}

// Transformer layer stack on caller-provided tokens (TSFormer pre-training: the encoder over the unmasked tokens
// and the decoder over [unmasked | mask tokens]; reference tsformer.py:86-136, transformer_layers.py:13-20).
// x: [S*P, 96] tokens already multiplied by sqrt(96), replaced by the (optionally final-normed) output.
extern "C" int step_ts_layers_fwd(float *x, int S, int P, const step_ts_layer_weights *L, int n_layers, const float *fnw,
                                  const float *fnb, void *workspace, size_t workspace_bytes, float drop_p,
                                  unsigned long long seed, void *stream) {
  STEP_REQUIRE(x && L && workspace && S > 0 && P > 0 && n_layers >= 1, "ts_layers: bad argument");
  STEP_REQUIRE((fnw == nullptr) == (fnb == nullptr), "ts_layers: final norm needs both weight and bias");
  if (workspace_bytes < step_ts_encoder_workspace_bytes(S, P))
    return fail(STEP_EWORKSPACE, "ts_layers: workspace too small (%lld bytes needed)",
                (long long)step_ts_encoder_workspace_bytes(S, P));
  float *ws = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  return run_layer_stack(x, S, P, L, n_layers, fnw, fnb, ws, (size_t)S * P, drop_p, seed, (cudaStream_t)stream);
}

// ===========================================================================
// cosine-similarity Gram matrix (similarity.py:6-16): row norms + batched X X^T with the
// normalisation folded into the GEMM epilogue.
// ===========================================================================
namespace stepk {
__global__ void row_norm_kernel(const float *__restrict__ x, long long D, float *__restrict__ norms) {
  const long long row = blockIdx.x;
  const float4 *p = reinterpret_cast<const float4 *>(x + row * D);
  float s = 0.f;
  for (long long i = threadIdx.x; i < D / 4; i += blockDim.x) {
    float4 v = p[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  __shared__ float red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) norms[row] = sqrtf(s);
  }
}
}  // namespace stepk

extern "C" int step_cosine_gram_f32(const float *x, int B, int N, long long D, float *norms, float *sim, void *stream) {
  STEP_REQUIRE(x && norms && sim && B > 0 && N > 0 && D > 0, "cosine_gram: bad argument");
  if (D % 16 != 0) return fail(STEP_EUNSUPPORTED, "cosine_gram: D=%lld must be a multiple of 16", D);
  cudaStream_t st = (cudaStream_t)stream;
  row_norm_kernel<<<(unsigned)(B * N), 256, 0, st>>>(x, D, norms);
  STEP_LAUNCH_CHECK("row_norm_kernel");
  GemmArgs g{};
  g.A = x; g.W = x; g.C = sim; g.M = N; g.K = (int)D; g.Nout = N;
  g.strideA = (long long)N * D; g.strideW = (long long)N * D; g.strideC = (long long)N * N;
  g.norm_a = norms; g.norm_b = norms; g.strideNorm = N;
  g.drop_scale = 1.f;
  return launch_gemm(g, EPI_COSINE, B, st);
}
