// Optimiser step and metric accumulation for the STEP training loop:
//   global-norm gradient clipping (torch.nn.utils.clip_grad_norm_, max_norm 3.0 / 5.0: reference step/STEP_METR-LA.py:105-107)
//   fused with torch.optim.Adam's update (lr, betas, eps, L2 weight decay: STEP_METR-LA.py:88-96) in TWO launches over all
//   ~130 parameter tensors (the reference path is clip_grad_norm_'s ~2 launches per tensor + Adam's multi-tensor loops), and
//   masked MAE / RMSE / MAPE accumulated on the device without the three `.item()` host syncs per step of
//   basicts/runners/base_tsf_runner.py:252-254.
// Tensors are addressed through a chunk table (tensor id, offset) so that parameters and gradients stay ordinary,
// separately allocated torch tensors; Adam's moments live in two flat fp32 buffers.
#include <math.h>
#include "common.cuh"

namespace stepk {

constexpr int OPT_CHUNK = 16384;      // elements per block

struct AdamTables {
  const long long *p_ptr, *g_ptr;     // per tensor: device addresses (g_ptr 0 = no gradient this step: skipped, as torch does)
  const long long *numel, *state_off; // per tensor
  const int *steps_in;                // per tensor: Adam steps taken so far (torch keeps `step` per parameter: a tensor
  int *steps_out;                     //   without gradient is skipped and does not advance)
  const int *chunk_tensor;            // per chunk
  const long long *chunk_off;         // per chunk: element offset inside the tensor
};

__global__ void __launch_bounds__(256) grad_sumsq_kernel(AdamTables t, double *__restrict__ sumsq) {
  __shared__ double red[8];
  const int ti = t.chunk_tensor[blockIdx.x];
  const float *g = reinterpret_cast<const float *>(t.g_ptr[ti]);
  double s = 0.0;
  if (g != nullptr) {
    const long long off = t.chunk_off[blockIdx.x], n = t.numel[ti];
    const long long end = off + OPT_CHUNK < n ? off + OPT_CHUNK : n;
    float a0 = 0.f, a1 = 0.f;
    long long i = off + threadIdx.x;
    for (; i + 256 < end; i += 512) { a0 = fmaf(g[i], g[i], a0); a1 = fmaf(g[i + 256], g[i + 256], a1); }
    if (i < end) a0 = fmaf(g[i], g[i], a0);
    s = (double)a0 + (double)a1;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += red[w];
    if (a != 0.0) atomicAdd(sumsq, a);
  }
}

// clip_coef = min(1, max_norm / (||g|| + 1e-6)); g' = clip_coef g + wd p; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2;
// p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)        (torch.optim.Adam, amsgrad off, maximize off)
__global__ void __launch_bounds__(256) adam_update_kernel(AdamTables t, const double *__restrict__ sumsq, float max_norm,
                                                          float lr, float b1, float b2, float eps, float wd,
                                                          float *__restrict__ m, float *__restrict__ v,
                                                          float *__restrict__ norm_out) {
  const int ti = t.chunk_tensor[blockIdx.x];
  const float *g = reinterpret_cast<const float *>(t.g_ptr[ti]);
  const float total = (float)sqrt(*sumsq);
  if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = total;
  const int step_prev = t.steps_in[ti];
  if (t.chunk_off[blockIdx.x] == 0 && threadIdx.x == 0) t.steps_out[ti] = step_prev + (g != nullptr ? 1 : 0);
  if (g == nullptr) return;
  const float tstep = (float)(step_prev + 1);
  const float bc1 = 1.f - powf(b1, tstep), bc2_sqrt = sqrtf(1.f - powf(b2, tstep));
  float coef = 1.f;
  if (max_norm > 0.f) { coef = max_norm / (total + 1e-6f); coef = coef > 1.f ? 1.f : coef; }
  float *p = reinterpret_cast<float *>(t.p_ptr[ti]);
  const long long off = t.chunk_off[blockIdx.x], n = t.numel[ti], so = t.state_off[ti];
  const long long end = off + OPT_CHUNK < n ? off + OPT_CHUNK : n;
  const float step = lr / bc1;
  for (long long i = off + threadIdx.x; i < end; i += 256) {
    const float pv = p[i];
    const float gv = fmaf(wd, pv, g[i] * coef);
    const float mv = fmaf(b1, m[so + i], (1.f - b1) * gv);
    const float vv = fmaf(b2, v[so + i], (1.f - b2) * gv * gv);
    m[so + i] = mv; v[so + i] = vv;
    p[i] = pv - step * mv / (sqrtf(vv) / bc2_sqrt + eps);
  }
}

// sums[0..4] += sum |e| w, sum e^2 w, sum |e / y| w2, sum w, sum w2   with w = [y is valid] and, for MAPE, w2 = [|y| >= 1e-4]
// (basicts/metrics/mape.py:21 zeroes labels below 1e-4 and always masks zeros)
__global__ void __launch_bounds__(256) metrics_reduce_kernel(const float *__restrict__ pred, const float *__restrict__ real,
                                                             long long n, float mean, float stdv, float null_val, int nan_mask,
                                                             double *__restrict__ sums) {
  __shared__ double red[5][8];
  double a = 0, b = 0, c = 0, w = 0, w2 = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += 256LL * gridDim.x) {
    const float y = fmaf(real[i], stdv, mean), p = fmaf(pred[i], stdv, mean);
    const bool ok = nan_mask ? !isnan(y) : fabsf(y - null_val) > 5e-5f;
    const float e = p - y;
    if (ok) { a += fabsf(e); b += (double)e * e; w += 1.0; }
    if (!isnan(y) && fabsf(y) >= 1e-4f) { c += fabsf(e / y); w2 += 1.0; }
  }
  a = warp_sum_d(a); b = warp_sum_d(b); c = warp_sum_d(c); w = warp_sum_d(w); w2 = warp_sum_d(w2);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; red[3][threadIdx.x >> 5] = w;
    red[4][threadIdx.x >> 5] = w2;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, s);
  }
}
// masked_mae / masked_rmse / masked_mape of THIS batch (mean(|e| mask / mean(mask)) = sum|e|w / sum w) added to the epoch
// accumulators acc[0..2], acc[3] += 1 batch; the per-batch sums are cleared for the next call
__global__ void metrics_finish_kernel(double *__restrict__ sums, double *__restrict__ acc) {
  const double w = sums[3], w2 = sums[4];
  if (w > 0) { acc[0] += sums[0] / w; acc[1] += sqrt(sums[1] / w); }
  if (w2 > 0) acc[2] += sums[2] / w2;
  acc[3] += 1.0;
  sums[0] = sums[1] = sums[2] = sums[3] = sums[4] = 0.0;
}

}  // namespace stepk

using namespace stepk;

extern "C" int step_opt_chunk_elems(void) { return OPT_CHUNK; }

extern "C" int step_clip_adam_step(const long long *p_ptr, const long long *g_ptr, const long long *numel,
                                   const long long *state_off, const int *chunk_tensor, const long long *chunk_off, int n_chunks,
                                   const int *steps_in, int *steps_out, float *m, float *v, double *sumsq, float max_norm, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float *norm_out, void *stream) {
  STEP_REQUIRE(p_ptr && g_ptr && numel && state_off && chunk_tensor && chunk_off && steps_in && steps_out && steps_in != steps_out &&
                   m && v && sumsq && n_chunks > 0, "clip_adam_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  AdamTables t{p_ptr, g_ptr, numel, state_off, steps_in, steps_out, chunk_tensor, chunk_off};
  cudaMemsetAsync(sumsq, 0, sizeof(double), st);
  grad_sumsq_kernel<<<n_chunks, 256, 0, st>>>(t, sumsq);
  STEP_LAUNCH_CHECK("grad_sumsq_kernel");
  adam_update_kernel<<<n_chunks, 256, 0, st>>>(t, sumsq, max_norm, lr, beta1, beta2, eps, weight_decay, m, v, norm_out);
  return check_launch("adam_update_kernel");
}

extern "C" int step_metrics_accumulate(const float *pred, const float *real, long long n, float mean, float stdv, float null_val,
                                       int use_nan_mask, double *sums /*[5] zero-initialised once*/, double *acc /*[4]*/,
                                       void *stream) {
  STEP_REQUIRE(pred && real && sums && acc && n > 0, "metrics_accumulate: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  long long grid = (n + 255) / 256;
  if (grid > 592) grid = 592;
  metrics_reduce_kernel<<<(unsigned)grid, 256, 0, st>>>(pred, real, n, mean, stdv, null_val, use_nan_mask, sums);
  STEP_LAUNCH_CHECK("metrics_reduce_kernel");
  metrics_finish_kernel<<<1, 1, 0, st>>>(sums, acc);
  return check_launch("metrics_finish_kernel");
}
