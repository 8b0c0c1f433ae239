// Discrete graph learning kernels: global top-k select for the kNN prior, the batch-invariant
// edge-logit MLP (fwd/bwd) and the hard Gumbel-softmax sample (fwd/bwd).
//
// Reference semantics (file:line relative to the reference repo):
//   step/step_arch/discrete_graph_learning.py:91-111   get_k_nn_neighbor (topk / scatter / where)
//   step/step_arch/discrete_graph_learning.py:148-153  edge MLP on [senders, receivers]
//   step/step_arch/discrete_graph_learning.py:11-45,157-161  gumbel_softmax(hard=True), diagonal removal
#include "common.cuh"

namespace stepk {

// ===========================================================================
// top-k mask: one 1024-thread block per sample, 4-pass MSB radix select on order-preserving
// integer keys, ties at the threshold resolved lowest-flat-index-first.
// ===========================================================================
__device__ __forceinline__ uint32_t order_key(float x) {
  uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(1024) topk_mask_kernel(const float *__restrict__ sim, int N, int k,
                                                         float *__restrict__ adj) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_k;
  __shared__ unsigned s_scan[1024];
  const int tid = threadIdx.x, lane = tid & 31;
  const long long total = (long long)N * N;
  const float *x = sim + (long long)blockIdx.x * total;
  float *out = adj + (long long)blockIdx.x * total;

  if (tid == 0) { s_prefix = 0; s_k = (unsigned)k; }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    for (long long i0 = 0; i0 < total; i0 += 1024) {
      const long long i = i0 + tid;
      bool active = i < total;
      unsigned key = active ? order_key(x[i]) : 0u;
      if (active && pass > 0) active = (key >> (shift + 8)) == prefix;
      const unsigned digit = (key >> shift) & 0xFFu;
      // warp-aggregated shared atomics: one add per distinct digit in the warp
      const unsigned amask = __ballot_sync(0xffffffffu, active);
      if (active) {
        const unsigned peers = __match_any_sync(amask, digit);
        if ((__ffs(peers) - 1) == lane) atomicAdd(&hist[digit], __popc(peers));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned need = s_k, cum = 0;
      int bin = 255;
      for (; bin > 0; --bin) {
        if (cum + hist[bin] >= need) break;
        cum += hist[bin];
      }
      s_k = need - cum;                       // how many to take inside the chosen bin
      s_prefix = (prefix << 8) | (unsigned)bin;
    }
    __syncthreads();
  }
  const unsigned thr_key = s_prefix;   // exact key of the k-th largest element
  const unsigned take_eq = s_k;        // number of elements equal to it that belong to the top-k

  // rank the threshold ties in flat-index order: contiguous chunk per thread + block scan
  const long long chunk = (total + 1023) / 1024;
  const long long c0 = (long long)tid * chunk, c1 = (c0 + chunk < total) ? (c0 + chunk) : total;
  unsigned eq = 0;
  for (long long i = c0; i < c1; ++i) eq += (order_key(x[i]) == thr_key);
  s_scan[tid] = eq;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 entries
  for (int off = 1; off < 1024; off <<= 1) {
    unsigned v = (tid >= off) ? s_scan[tid - off] : 0u;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  unsigned rank = s_scan[tid] - eq;  // exclusive prefix
  for (long long i = c0; i < c1; ++i) {
    const float v = x[i];
    const unsigned key = order_key(v);
    bool sel = key > thr_key;
    if (key == thr_key) { sel = rank < take_eq; ++rank; }
    const int r = (int)(i / N), c = (int)(i - (long long)r * N);
    out[i] = (sel && v != 0.0f && r != c) ? 1.0f : 0.0f;
  }
}

// ===========================================================================
// edge logits
// ===========================================================================
// grid (ceil(N/128), N): block = receiver i, 128 senders j
__global__ void __launch_bounds__(128) edge_logits_fwd_kernel(const float *__restrict__ ut, const float *__restrict__ v,
                                                              const float *__restrict__ cat_w,
                                                              const float *__restrict__ cat_b, int N, int F,
                                                              float *__restrict__ logits, float *__restrict__ theta) {
  extern __shared__ float sm[];
  float *sv = sm, *sw0 = sm + F, *sw1 = sm + 2 * F;
  const int i = blockIdx.y, j = blockIdx.x * 128 + threadIdx.x;
  for (int f = threadIdx.x; f < F; f += 128) {
    sv[f] = v[(size_t)i * F + f];
    sw0[f] = cat_w[f];
    sw1[f] = cat_w[F + f];
  }
  __syncthreads();
  if (j >= N) return;
  float l0 = 0.f, l1 = 0.f;
  for (int f = 0; f < F; ++f) {
    const float h = fmaxf(ut[(size_t)f * N + j] + sv[f], 0.f);
    l0 = fmaf(sw0[f], h, l0);
    l1 = fmaf(sw1[f], h, l1);
  }
  l0 += cat_b[0];
  l1 += cat_b[1];
  const size_t e = (size_t)i * N + j;
  reinterpret_cast<float2 *>(logits)[e] = make_float2(l0, l1);
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  theta[e] = e0 / (e0 + e1);
}

// rows: block = receiver i, thread = feature f.  dv[i,f], partial dcat_w / dcat_b (atomics over i).
__global__ void edge_logits_bwd_rows_kernel(const float *__restrict__ dl, const float *__restrict__ ut,
                                            const float *__restrict__ v, const float *__restrict__ cat_w, int N, int F,
                                            float *__restrict__ dv, float *__restrict__ dcat_w,
                                            float *__restrict__ dcat_b) {
  const int i = blockIdx.x, f = threadIdx.x;
  const float2 *dli = reinterpret_cast<const float2 *>(dl) + (size_t)i * N;
  if (f < F) {
    const float vi = v[(size_t)i * F + f], w0 = cat_w[f], w1 = cat_w[F + f];
    const float *utf = ut + (size_t)f * N;
    float a_dv = 0.f, a_w0 = 0.f, a_w1 = 0.f;
    for (int j = 0; j < N; ++j) {
      const float2 g = dli[j];
      const float pre = utf[j] + vi;
      if (pre > 0.f) {
        a_dv += w0 * g.x + w1 * g.y;
        a_w0 = fmaf(g.x, pre, a_w0);
        a_w1 = fmaf(g.y, pre, a_w1);
      }
    }
    dv[(size_t)i * F + f] = a_dv;
    atomicAdd(&dcat_w[f], a_w0);
    atomicAdd(&dcat_w[F + f], a_w1);
  }
  if (f >= blockDim.x - 2) {  // two spare lanes sum the bias gradient of this row
    const int c = f - (blockDim.x - 2);
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += c ? dli[j].y : dli[j].x;
    atomicAdd(&dcat_b[c], s);
  }
}

// cols: block = sender j, thread = feature f.  dut[f,j]
__global__ void edge_logits_bwd_cols_kernel(const float *__restrict__ dl, const float *__restrict__ ut,
                                            const float *__restrict__ v, const float *__restrict__ cat_w, int N, int F,
                                            float *__restrict__ dut) {
  const int j = blockIdx.x, f = threadIdx.x;
  if (f >= F) return;
  const float uj = ut[(size_t)f * N + j], w0 = cat_w[f], w1 = cat_w[F + f];
  const float2 *dlp = reinterpret_cast<const float2 *>(dl);
  float a = 0.f;
  for (int i = 0; i < N; ++i) {
    const float2 g = dlp[(size_t)i * N + j];
    const float pre = uj + v[(size_t)i * F + f];
    if (pre > 0.f) a += w0 * g.x + w1 * g.y;
  }
  dut[(size_t)f * N + j] = a;
}

// ===========================================================================
// hard Gumbel-softmax sample
// ===========================================================================
__global__ void gumbel_fwd_kernel(const float *__restrict__ logits, const float *__restrict__ uniform, int N,
                                  long long per_sample, float inv_tau, uint64_t key, float *__restrict__ sampled,
                                  float *__restrict__ y0) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (e >= per_sample) return;
  const float2 l = reinterpret_cast<const float2 *>(logits)[e];
  const long long idx = (long long)b * per_sample + e;
  float u0, u1;
  if (uniform != nullptr) {
    const float2 u = reinterpret_cast<const float2 *>(uniform)[idx];
    u0 = u.x; u1 = u.y;
  } else {
    const uint4 r = philox4x32((uint64_t)idx, key);
    u0 = u01(r.x); u1 = u01(r.y);
  }
  const float eps = 1e-10f;
  const float g0 = -logf(-logf(u0 + eps) + eps), g1 = -logf(-logf(u1 + eps) + eps);
  const float a = (l.x + g0) * inv_tau, c = (l.y + g1) * inv_tau;
  const float m = fmaxf(a, c);
  const float e0 = expf(a - m), e1 = expf(c - m);
  const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
  const int r = (int)(e / N), col = (int)(e - (long long)r * N);
  y0[idx] = p0;
  sampled[idx] = (p0 >= p1 && r != col) ? 1.0f : 0.0f;   // argmax picks class 0 on ties (first index)
}

__global__ void gumbel_bwd_kernel(const float *__restrict__ dsampled, const float *__restrict__ y0, int B, int N,
                                  long long per_sample, float inv_tau, int accumulate, float *__restrict__ dlogits) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per_sample) return;
  const int r = (int)(e / N), col = (int)(e - (long long)r * N);
  float acc = 0.f;
  if (r != col) {
    for (int b = 0; b < B; ++b) {
      const float p = y0[(long long)b * per_sample + e];
      acc = fmaf(dsampled[(long long)b * per_sample + e], p * (1.f - p), acc);
    }
    acc *= inv_tau;
  }
  float2 *d = reinterpret_cast<float2 *>(dlogits) + e;
  if (accumulate) { float2 o = *d; *d = make_float2(o.x + acc, o.y - acc); }
  else *d = make_float2(acc, -acc);
}

}  // namespace stepk

using namespace stepk;

extern "C" int step_topk_mask_f32(const float *sim, int B, int N, int k, float *adj, void *stream) {
  STEP_REQUIRE(sim && adj && B > 0 && N > 0, "topk_mask: bad argument");
  STEP_REQUIRE(k >= 1 && (long long)k <= (long long)N * N, "topk_mask: k out of range");
  topk_mask_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(sim, N, k, adj);
  return check_launch("topk_mask_kernel");
}

extern "C" int step_edge_logits_fwd(const float *ut, const float *v, const float *cat_w, const float *cat_b, int N, int F,
                                    float *logits, float *theta, void *stream) {
  STEP_REQUIRE(ut && v && cat_w && cat_b && logits && theta, "edge_logits_fwd: null pointer");
  STEP_REQUIRE(N > 0 && N <= 65535 && F > 0 && F <= 1024, "edge_logits_fwd: bad shape");
  edge_logits_fwd_kernel<<<dim3((N + 127) / 128, N), 128, 3 * F * sizeof(float), (cudaStream_t)stream>>>(
      ut, v, cat_w, cat_b, N, F, logits, theta);
  return check_launch("edge_logits_fwd_kernel");
}

extern "C" int step_edge_logits_bwd(const float *dlogits, const float *ut, const float *v, const float *cat_w, int N, int F,
                                    float *dut, float *dv, float *dcat_w, float *dcat_b, void *stream) {
  STEP_REQUIRE(dlogits && ut && v && cat_w && dut && dv && dcat_w && dcat_b, "edge_logits_bwd: null pointer");
  STEP_REQUIRE(N > 0 && F > 0 && F <= 1022, "edge_logits_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dcat_w, 0, 2 * F * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dcat_b, 0, 2 * sizeof(float), st);
  if (e != cudaSuccess) return fail_msg((int)e, cudaGetErrorString(e));
  const int threads = ((F + 2) + 31) / 32 * 32;
  edge_logits_bwd_rows_kernel<<<N, threads, 0, st>>>(dlogits, ut, v, cat_w, N, F, dv, dcat_w, dcat_b);
  STEP_LAUNCH_CHECK("edge_logits_bwd_rows_kernel");
  edge_logits_bwd_cols_kernel<<<N, (F + 31) / 32 * 32, 0, st>>>(dlogits, ut, v, cat_w, N, F, dut);
  return check_launch("edge_logits_bwd_cols_kernel");
}

extern "C" int step_gumbel_sample_fwd(const float *logits, const float *uniform, int B, int N, float tau,
                                      unsigned long long seed, float *sampled, float *y0, void *stream) {
  STEP_REQUIRE(logits && sampled && y0 && B > 0 && B <= 65535 && N > 0 && tau > 0.f, "gumbel_sample_fwd: bad argument");
  const long long per = (long long)N * N;
  gumbel_fwd_kernel<<<dim3((unsigned)((per + 255) / 256), B), 256, 0, (cudaStream_t)stream>>>(
      logits, uniform, N, per, 1.0f / tau, rng_key(seed, 0x6B), sampled, y0);
  return check_launch("gumbel_fwd_kernel");
}

extern "C" int step_gumbel_sample_bwd(const float *dsampled, const float *y0, int B, int N, float tau, int accumulate,
                                      float *dlogits, void *stream) {
  STEP_REQUIRE(dsampled && y0 && dlogits && B > 0 && N > 0 && tau > 0.f, "gumbel_sample_bwd: bad argument");
  const long long per = (long long)N * N;
  gumbel_bwd_kernel<<<(unsigned)((per + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dsampled, y0, B, N, per, 1.0f / tau,
                                                                                    accumulate, dlogits);
  return check_launch("gumbel_bwd_kernel");
}

// ===========================================================================
// STEP loss, forward + backward in two launches:
//   loss = masked_mae(pred, real, null_val) + coeff * BCE(theta, adj_knn)
//   step/step_loss/step_loss.py:5-16, basicts/metrics/mae.py:5-28 (mask = |real - null| > 5e-5, weights mask/mean(mask)),
//   nn.BCELoss (log clamped at -100).  theta is the batch-invariant [N,N] probability matrix (the reference's
//   [B,N,N] tensor has B identical slices), adj_knn [B,N,N].
// ===========================================================================
namespace stepk {

// sums[0] = sum |p - y| * m, sums[1] = sum m, sums[2] = sum_ij (c1 log(th) + (B - c1) log(1 - th))
__global__ void __launch_bounds__(256) step_loss_reduce_kernel(const float *__restrict__ pred, const float *__restrict__ real,
                                                               long long n_pred, float mean, float stdv, float null_val,
                                                               int use_nan_mask, const float *__restrict__ theta,
                                                               const float *__restrict__ knn, int B, long long nn,
                                                               double *__restrict__ sums) {
  float s_abs = 0.f, s_m = 0.f, s_bce = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pred; i += stride) {
    const float y = fmaf(real[i], stdv, mean), p = fmaf(pred[i], stdv, mean);
    const bool m = use_nan_mask ? !isnan(y) : (fabsf(y - null_val) > 5e-5f);
    if (m) { s_abs += fabsf(p - y); s_m += 1.f; }
  }
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += stride) {
    float c1 = 0.f;
    for (int b = 0; b < B; ++b) c1 += knn[(size_t)b * nn + e];
    const float th = theta[e];
    s_bce += c1 * fmaxf(logf(th), -100.f) + ((float)B - c1) * fmaxf(logf(1.f - th), -100.f);
  }
  __shared__ float red[3][8];
  s_abs = warp_sum(s_abs); s_m = warp_sum(s_m); s_bce = warp_sum(s_bce);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = s_abs; red[1][w] = s_m; red[2][w] = s_bce; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, (double)t);
  }
}

// loss value + gradients (d loss / d pred in the *unscaled* prediction space, d loss / d theta)
__global__ void __launch_bounds__(256) step_loss_finish_kernel(const float *__restrict__ pred, const float *__restrict__ real,
                                                               long long n_pred, float mean, float stdv, float null_val,
                                                               int use_nan_mask, const float *__restrict__ theta,
                                                               const float *__restrict__ knn, int B, long long nn, float coeff,
                                                               const double *__restrict__ sums, float *__restrict__ loss,
                                                               float *__restrict__ dpred, float *__restrict__ dtheta) {
  const double sm = sums[1];
  const float inv_m = sm > 0.0 ? (float)(1.0 / sm) : 0.f;
  const float bce_scale = coeff / ((float)B * (float)nn);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    loss[0] = (sm > 0.0 ? (float)(sums[0] / sm) : 0.f) - bce_scale * (float)sums[2];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pred; i += stride) {
    const float y = fmaf(real[i], stdv, mean), p = fmaf(pred[i], stdv, mean);
    const bool m = use_nan_mask ? !isnan(y) : (fabsf(y - null_val) > 5e-5f);
    const float d = p - y;
    dpred[i] = m ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_m * stdv : 0.f;
  }
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += stride) {
    float c1 = 0.f;
    for (int b = 0; b < B; ++b) c1 += knn[(size_t)b * nn + e];
    const float th = theta[e];
    // d/dth of -(c1 log th + (B-c1) log(1-th)); the -100 clamp has zero slope where it is active
    const float g1 = (logf(th) > -100.f) ? c1 / th : 0.f;
    const float g0 = (logf(1.f - th) > -100.f) ? ((float)B - c1) / (1.f - th) : 0.f;
    dtheta[e] = -bce_scale * (g1 - g0);
  }
}

}  // namespace stepk

extern "C" int step_loss_fwd_bwd(const float *pred, const float *real, long long n_pred, float mean, float stdv, float null_val,
                                 int use_nan_mask, const float *theta, const float *adj_knn, int B, int N, float coeff,
                                 float *loss, float *dpred, float *dtheta, void *scratch, void *stream) {
  STEP_REQUIRE(pred && real && theta && adj_knn && loss && dpred && dtheta && scratch && n_pred > 0 && B > 0 && N > 0,
               "step_loss_fwd_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  double *sums = reinterpret_cast<double *>(scratch);
  cudaError_t e = cudaMemsetAsync(sums, 0, 3 * sizeof(double), st);
  if (e != cudaSuccess) return fail_msg((int)e, cudaGetErrorString(e));
  const long long nn = (long long)N * N;
  step_loss_reduce_kernel<<<296, 256, 0, st>>>(pred, real, n_pred, mean, stdv, null_val, use_nan_mask, theta, adj_knn, B, nn, sums);
  STEP_LAUNCH_CHECK("step_loss_reduce_kernel");
  step_loss_finish_kernel<<<296, 256, 0, st>>>(pred, real, n_pred, mean, stdv, null_val, use_nan_mask, theta, adj_knn, B, nn, coeff,
                                               sums, loss, dpred, dtheta);
  return check_launch("step_loss_finish_kernel");
}
