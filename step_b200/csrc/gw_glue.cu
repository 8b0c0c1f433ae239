// Graph WaveNet prologue (step/step_arch/graphwavenet/model.py:144-166), forward and backward, as small fused kernels:
//   start conv      x0[b,t,n,:] = W_s xin[b,t,n,0:2] + b_s  on the left-padded history (t = 0 is the zero pad), written
//                   directly in the layer stack's [B,13,N,32] layout                                   (:145-155)
//   supports        P1 = D^-1 (A + I),  P2 = D'^-1 (A^T + I)   (_calculate_random_walk_matrix, :121-130, :160)
//   adaptive graph  P3 = softmax_row(relu(E1 E2))                                                       (:165)
// The reference runs these as ~15 library kernels forward and ~40 in autograd's backward; they are HBM/latency
// trivial, so each direction is 1-3 launches here.
#include "common.cuh"

namespace stepk {

// ---------------------------------------------------------------------------
// start conv
// ---------------------------------------------------------------------------
__global__ void gw_start_fwd_kernel(const float *__restrict__ hist, int B, int T, int N, int C, const float *__restrict__ w,
                                    const float *__restrict__ bias, float4 *__restrict__ x0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (row = (b,t,n), c4)
  const long long rows = (long long)B * (T + 1) * N;
  if (idx >= rows * 8) return;
  const int c4 = (int)(idx & 7);
  const long long row = idx >> 3;
  const int n = (int)(row % N);
  const int t = (int)((row / N) % (T + 1));
  const long long b = row / ((long long)N * (T + 1));
  float i0 = 0.f, i1 = 0.f;
  if (t > 0) {
    const float *h = hist + (((size_t)b * T + (t - 1)) * N + n) * C;
    i0 = h[0]; i1 = h[1];
  }
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c4 * 4 + j;
    o[j] = fmaf(w[c * 2], i0, fmaf(w[c * 2 + 1], i1, bias[c]));
  }
  x0[idx] = make_float4(o[0], o[1], o[2], o[3]);
}

// dW[c][i] = sum dx0[row][c] xin[row][i], db[c] = sum dx0[row][c]; out = [32*2 dW | 32 db], zero-initialised by the host
__global__ void __launch_bounds__(256) gw_start_bwd_kernel(const float *__restrict__ hist, int B, int T, int N, int C,
                                                           const float *__restrict__ dx0, float *__restrict__ out) {
  __shared__ float red[3][8][32];
  const int c = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const long long rows = (long long)B * (T + 1) * N;
  float a0 = 0.f, a1 = 0.f, ab = 0.f;
  for (long long row = (long long)blockIdx.x * 8 + slot; row < rows; row += 8LL * gridDim.x) {
    const int n = (int)(row % N);
    const int t = (int)((row / N) % (T + 1));
    const long long b = row / ((long long)N * (T + 1));
    const float d = dx0[row * 32 + c];
    ab += d;
    if (t > 0) {
      const float *h = hist + (((size_t)b * T + (t - 1)) * N + n) * C;
      a0 = fmaf(d, h[0], a0);
      a1 = fmaf(d, h[1], a1);
    }
  }
  red[0][slot][c] = a0; red[1][slot][c] = a1; red[2][slot][c] = ab;
  __syncthreads();
  if (slot < 3) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[slot][i][c];
    atomicAdd(out + (slot < 2 ? c * 2 + slot : 64 + c), s);
  }
}

// ---------------------------------------------------------------------------
// random-walk supports
// ---------------------------------------------------------------------------
// deg[0][b][v] = 1 + sum_w A[b,v,w] (row degree), deg[1][b][v] = 1 + sum_u A[b,u,v] (column degree)
__global__ void __launch_bounds__(256) gw_degree_kernel(const float *__restrict__ A, int B, int N, float *__restrict__ deg) {
  const int b = blockIdx.y;
  const float *Ab = A + (size_t)b * N * N;
  if (blockIdx.x < (unsigned)N) {                    // one block per row: row degree
    __shared__ float red[8];
    const int v = blockIdx.x;
    float s = 0.f;
    for (int w = threadIdx.x; w < N; w += 256) s += Ab[(size_t)v * N + w];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 1.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      deg[(size_t)b * N + v] = t;
    }
  } else {                                           // column degrees: thread = column (coalesced over rows)
    const int w = (blockIdx.x - N) * 256 + threadIdx.x;
    if (w >= N) return;
    float s = 1.f;
    for (int u = 0; u < N; ++u) s += Ab[(size_t)u * N + w];
    deg[((size_t)B + b) * N + w] = s;
  }
}

// P1[b,v,w] = (A[b,v,w] + [v==w]) / deg0[b,v];  P2[b,v,w] = (A[b,w,v] + [v==w]) / deg1[b,v]   (32x32 tiles, smem transpose)
__global__ void __launch_bounds__(256) gw_supports_fwd_kernel(const float *__restrict__ A, const float *__restrict__ deg, int B,
                                                              int N, float *__restrict__ P1, float *__restrict__ P2) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, tv = blockIdx.y * 32, tw = blockIdx.x * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const float *Ab = A + (size_t)b * N * N;
  const float *d0 = deg + (size_t)b * N, *d1 = deg + ((size_t)B + b) * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = tv + ly + 8 * i, w = tw + lx;
    float x = 0.f;
    if (v < N && w < N) {
      x = Ab[(size_t)v * N + w] + (v == w ? 1.f : 0.f);
      P1[((size_t)b * N + v) * N + w] = x / d0[v];
    }
    tile[ly + 8 * i][lx] = x;
  }
  __syncthreads();
  // transposed tile: P2 rows are the columns w of A
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tw + ly + 8 * i, cc = tv + lx;       // P2[b, r, cc] = (A[b, cc, r] + [r==cc]) / deg1[r]
    if (r < N && cc < N) P2[((size_t)b * N + r) * N + cc] = tile[lx][ly + 8 * i] / d1[r];
  }
}

// t[0][b][v] = sum_w dP1[b,v,w] P1[b,v,w];  t[1][b][v] = sum_w dP2[b,v,w] P2[b,v,w]
__global__ void __launch_bounds__(256) gw_supports_dot_kernel(const float *__restrict__ dP1, const float *__restrict__ P1,
                                                              const float *__restrict__ dP2, const float *__restrict__ P2, int B, int N,
                                                              float *__restrict__ t) {
  __shared__ float red[2][8];
  const int b = blockIdx.y, v = blockIdx.x;
  const size_t off = ((size_t)b * N + v) * N;
  float s1 = 0.f, s2 = 0.f;
  for (int w = threadIdx.x; w < N; w += 256) {
    s1 = fmaf(dP1[off + w], P1[off + w], s1);
    s2 = fmaf(dP2[off + w], P2[off + w], s2);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; c += red[1][i]; }
    t[(size_t)b * N + v] = a;
    t[((size_t)B + b) * N + v] = c;
  }
}

// dA[b,v,w] = (dP1[b,v,w] - t0[b,v]) / deg0[b,v] + (dP2[b,w,v] - t1[b,w]) / deg1[b,w]
__global__ void __launch_bounds__(256) gw_supports_bwd_kernel(const float *__restrict__ dP1, const float *__restrict__ dP2,
                                                              const float *__restrict__ deg, const float *__restrict__ t, int B, int N,
                                                              float *__restrict__ dA) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, tv = blockIdx.y * 32, tw = blockIdx.x * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const float *d0 = deg + (size_t)b * N, *d1 = deg + ((size_t)B + b) * N;
  const float *t0 = t + (size_t)b * N, *t1 = t + ((size_t)B + b) * N;
  // stage the dP2 tile (rows w, columns v) so that it can be read transposed
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tw + ly + 8 * i, cc = tv + lx;
    tile[ly + 8 * i][lx] = (r < N && cc < N) ? (dP2[((size_t)b * N + r) * N + cc] - t1[r]) / d1[r] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = tv + ly + 8 * i, w = tw + lx;
    if (v < N && w < N)
      dA[((size_t)b * N + v) * N + w] = (dP1[((size_t)b * N + v) * N + w] - t0[v]) / d0[v] + tile[lx][ly + 8 * i];
  }
}

// ---------------------------------------------------------------------------
// adaptive adjacency
// ---------------------------------------------------------------------------
constexpr int ADP_R = 10;
__global__ void __launch_bounds__(256) gw_adp_fwd_kernel(const float *__restrict__ E1, const float *__restrict__ E2, int N,
                                                         float *__restrict__ P3) {
  __shared__ float red[8];
  __shared__ float bc;
  const int v = blockIdx.x;
  float e[ADP_R];
#pragma unroll
  for (int k = 0; k < ADP_R; ++k) e[k] = E1[(size_t)v * ADP_R + k];
  float *row = P3 + (size_t)v * N;
  float m = 0.f;                                     // relu output is >= 0
  for (int w = threadIdx.x; w < N; w += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ADP_R; ++k) s = fmaf(e[k], E2[(size_t)k * N + w], s);
    s = fmaxf(s, 0.f);
    row[w] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { float t = red[0]; for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]); bc = t; }
  __syncthreads();
  m = bc;
  float z = 0.f;
  for (int w = threadIdx.x; w < N; w += 256) { const float p = __expf(row[w] - m); row[w] = p; z += p; }
  z = warp_sum(z);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = z;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bc = 1.f / t; }
  __syncthreads();
  const float inv = bc;
  for (int w = threadIdx.x; w < N; w += 256) row[w] *= inv;
}

// per row v: dS = P3 (dP3 - <dP3, P3>), dR = dS [E1 E2 > 0] -> dR[v,:] (scratch), dE1[v,k] = sum_w dR[v,w] E2[k,w]
__global__ void __launch_bounds__(256) gw_adp_bwd_kernel(const float *__restrict__ E1, const float *__restrict__ E2,
                                                         const float *__restrict__ P3, const float *__restrict__ dP3, int N,
                                                         float *__restrict__ dR, float *__restrict__ dE1) {
  __shared__ float red[ADP_R + 1][8];
  __shared__ float bc;
  const int v = blockIdx.x;
  float e[ADP_R];
#pragma unroll
  for (int k = 0; k < ADP_R; ++k) e[k] = E1[(size_t)v * ADP_R + k];
  const float *p = P3 + (size_t)v * N, *dp = dP3 + (size_t)v * N;
  float t = 0.f;
  for (int w = threadIdx.x; w < N; w += 256) t = fmaf(p[w], dp[w], t);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[ADP_R][threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) { float s = 0.f; for (int i = 0; i < 8; ++i) s += red[ADP_R][i]; bc = s; }
  __syncthreads();
  t = bc;
  float acc[ADP_R];
#pragma unroll
  for (int k = 0; k < ADP_R; ++k) acc[k] = 0.f;
  for (int w = threadIdx.x; w < N; w += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ADP_R; ++k) s = fmaf(e[k], E2[(size_t)k * N + w], s);
    const float d = s > 0.f ? p[w] * (dp[w] - t) : 0.f;
    dR[(size_t)v * N + w] = d;
#pragma unroll
    for (int k = 0; k < ADP_R; ++k) acc[k] = fmaf(d, E2[(size_t)k * N + w], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < ADP_R; ++k) {
    const float s = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < ADP_R) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    dE1[(size_t)v * ADP_R + threadIdx.x] = s;
  }
}

// dE2[k,w] += sum_{v in slab} E1[v,k] dR[v,w]   (grid.y slabs of rows; dE2 zero-initialised by the host)
__global__ void gw_adp_bwd_e2_kernel(const float *__restrict__ E1, const float *__restrict__ dR, int N, float *__restrict__ dE2) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  const int per = (N + gridDim.y - 1) / gridDim.y, v0 = blockIdx.y * per, v1 = min(N, v0 + per);
  float acc[ADP_R];
#pragma unroll
  for (int k = 0; k < ADP_R; ++k) acc[k] = 0.f;
  for (int v = v0; v < v1; ++v) {
    const float d = dR[(size_t)v * N + w];
#pragma unroll
    for (int k = 0; k < ADP_R; ++k) acc[k] = fmaf(E1[(size_t)v * ADP_R + k], d, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < ADP_R; ++k) atomicAdd(dE2 + (size_t)k * N + w, acc[k]);
}

}  // namespace stepk

using namespace stepk;

extern "C" int step_gw_start_fwd(const float *history, int B, int T, int N, int C, const float *w, const float *bias, float *x0,
                                 void *stream) {
  STEP_REQUIRE(history && w && bias && x0 && B > 0 && T > 0 && N > 0 && C >= 2, "gw_start_fwd: bad argument");
  const long long n = (long long)B * (T + 1) * N * 8;
  gw_start_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(history, B, T, N, C, w, bias,
                                                                                     reinterpret_cast<float4 *>(x0));
  return check_launch("gw_start_fwd_kernel");
}

extern "C" int step_gw_start_bwd(const float *history, int B, int T, int N, int C, const float *dx0, float *dw_db, void *stream) {
  STEP_REQUIRE(history && dx0 && dw_db && B > 0 && T > 0 && N > 0 && C >= 2, "gw_start_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dw_db, 0, 96 * sizeof(float), st);
  gw_start_bwd_kernel<<<296, 256, 0, st>>>(history, B, T, N, C, dx0, dw_db);
  return check_launch("gw_start_bwd_kernel");
}

extern "C" int step_gw_supports_fwd(const float *adj, int B, int N, float *deg, float *P1, float *P2, void *stream) {
  STEP_REQUIRE(adj && deg && P1 && P2 && B > 0 && N > 0, "gw_supports_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  gw_degree_kernel<<<dim3(N + (N + 255) / 256, B), 256, 0, st>>>(adj, B, N, deg);
  STEP_LAUNCH_CHECK("gw_degree_kernel");
  gw_supports_fwd_kernel<<<dim3((N + 31) / 32, (N + 31) / 32, B), 256, 0, st>>>(adj, deg, B, N, P1, P2);
  return check_launch("gw_supports_fwd_kernel");
}

extern "C" int step_gw_supports_bwd(const float *dP1, const float *dP2, const float *P1, const float *P2, const float *deg, int B,
                                    int N, float *dots, float *dadj, void *stream) {
  STEP_REQUIRE(dP1 && dP2 && P1 && P2 && deg && dots && dadj && B > 0 && N > 0, "gw_supports_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  gw_supports_dot_kernel<<<dim3(N, B), 256, 0, st>>>(dP1, P1, dP2, P2, B, N, dots);
  STEP_LAUNCH_CHECK("gw_supports_dot_kernel");
  gw_supports_bwd_kernel<<<dim3((N + 31) / 32, (N + 31) / 32, B), 256, 0, st>>>(dP1, dP2, deg, dots, B, N, dadj);
  return check_launch("gw_supports_bwd_kernel");
}

extern "C" int step_gw_adp_fwd(const float *E1, const float *E2, int N, int R, float *P3, void *stream) {
  STEP_REQUIRE(E1 && E2 && P3 && N > 0, "gw_adp_fwd: bad argument");
  if (R != ADP_R) return fail(STEP_EUNSUPPORTED, "gw_adp_fwd: node embedding rank %lld (STEP uses 10)", R);
  gw_adp_fwd_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(E1, E2, N, P3);
  return check_launch("gw_adp_fwd_kernel");
}

extern "C" int step_gw_adp_bwd(const float *E1, const float *E2, const float *P3, const float *dP3, int N, int R, float *scratch,
                               float *dE1, float *dE2, void *stream) {
  STEP_REQUIRE(E1 && E2 && P3 && dP3 && scratch && dE1 && dE2 && N > 0, "gw_adp_bwd: bad argument");
  if (R != ADP_R) return fail(STEP_EUNSUPPORTED, "gw_adp_bwd: node embedding rank %lld (STEP uses 10)", R);
  cudaStream_t st = (cudaStream_t)stream;
  gw_adp_bwd_kernel<<<N, 256, 0, st>>>(E1, E2, P3, dP3, N, scratch, dE1);
  STEP_LAUNCH_CHECK("gw_adp_bwd_kernel");
  cudaMemsetAsync(dE2, 0, (size_t)ADP_R * N * sizeof(float), st);
  gw_adp_bwd_e2_kernel<<<dim3((N + 127) / 128, 32), 128, 0, st>>>(E1, scratch, N, dE2);
  return check_launch("gw_adp_bwd_e2_kernel");
}
