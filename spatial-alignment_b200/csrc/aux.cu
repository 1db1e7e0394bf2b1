// The two pieces of a training run that sit right next to the ELBO iteration:
//   * the optimiser step -- Adam exactly as torch.optim.Adam (no weight decay, no amsgrad) as ONE launch over every
//     parameter tensor (reference examples/grid_example.py:59,76 `optimizer.step()`), step count on the device so the
//     launch can live inside a CUDA graph;
//   * initialisation of the inducing locations -- Lloyd's k-means on the spot coordinates (reference
//     gpsa/models/vgpsa.py:61-92 runs sklearn.cluster.KMeans on the host: 1.5 s at C3, minutes at C5).
#include "common.cuh"
#include "gpsa_b200.h"

#include <math.h>

namespace {

// ------------------------------------------------------------------------------------------------
// Adam
// ------------------------------------------------------------------------------------------------
// per-tensor step counts (torch.optim.Adam keeps `step` per parameter: a tensor without a gradient does not advance)
__global__ void adam_tick_kernel(const gpsa_adam_args a) {
  const int k = threadIdx.x;
  if (k < a.count && a.g[k] != nullptr) a.step[k] += 1.0f;
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const gpsa_adam_args a) {
  const float t = a.step[blockIdx.y];  // already incremented by adam_tick_kernel on the same stream
  // bias corrections in double like torch's host-side scalars (1 - 0.999^t loses 5 digits in float at small t)
  const double bc1d = 1.0 - pow(a.beta1, (double)t);
  const float bc2s = (float)sqrt(1.0 - pow(a.beta2, (double)t));
  const float step_size = (float)(a.lr / bc1d);
  const float b2 = (float)a.beta2, omb1 = (float)(1.0 - a.beta1), omb2 = (float)(1.0 - a.beta2);
  const float eps = (float)a.eps;
  // blockIdx.y = tensor, blockIdx.x strides over its elements
  const int k = blockIdx.y;
  float* __restrict__ p = a.p[k];
  const float* __restrict__ g = a.g[k];
  float* __restrict__ m = a.m[k];
  float* __restrict__ v = a.v[k];
  const long n = a.n[k];
  if (g == nullptr) return;  // parameter without a gradient this step: untouched, like torch
  const long stride = (long)gridDim.x * blockDim.x;
  const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long n4 = al ? n / 4 : 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
#define ADAM1(c)                                                  \
  mm.c = mm.c + omb1 * (gg.c - mm.c);       /* exp_avg.lerp_(grad, 1 - beta1) */ \
  vv.c = b2 * vv.c + omb2 * gg.c * gg.c;    /* exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2) */ \
  pp.c -= step_size * mm.c / (sqrtf(vv.c) / bc2s + eps);
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long i = 4 * n4 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = m[i] + omb1 * (gi - m[i]);
    const float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2s + eps);
  }
}

// ------------------------------------------------------------------------------------------------
// Lloyd's k-means, D <= 3, K centres in shared memory
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) kmeans_assign_kernel(long N, int K, const float* __restrict__ X,
                                                            const float* __restrict__ Cn, int* __restrict__ assign,
                                                            double* __restrict__ sums, double* inertia) {
  extern __shared__ float sc[];  // [K*D]
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) sc[i] = Cn[i];
  __syncthreads();
  double cost = 0.0;
  for (long n = (long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long)gridDim.x * blockDim.x) {
    float x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = X[n * D + d];
    float best = INFINITY;
    int bk = 0;
    for (int k = 0; k < K; ++k) {
      float r2 = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float t = x[d] - sc[k * D + d];
        r2 = fmaf(t, t, r2);
      }
      if (r2 < best) { best = r2; bk = k; }
    }
    assign[n] = bk;
    cost += (double)best;
#pragma unroll
    for (int d = 0; d < D; ++d) atomicAdd(&sums[bk * (D + 1) + d], (double)x[d]);
    atomicAdd(&sums[bk * (D + 1) + D], 1.0);
  }
  __shared__ double red[32];
  cost = block_sum<double>(cost, red);
  if (threadIdx.x == 0) atomicAdd(inertia, cost);
}

template <int D>
__global__ void kmeans_update_kernel(int K, const double* __restrict__ sums, float* __restrict__ Cn) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double cnt = sums[k * (D + 1) + D];
  if (cnt > 0.5) {  // an empty cluster keeps its centre
#pragma unroll
    for (int d = 0; d < D; ++d) Cn[k * D + d] = (float)(sums[k * (D + 1) + d] / cnt);
  }
}

}  // namespace

extern "C" int gpsa_adam_step(const gpsa_adam_args* a, cudaStream_t st) {
  if (!a || a->count < 0 || a->count > GPSA_ADAM_MAX_TENSORS || !a->step) return GPSA_ERR_ARG;
  if (a->count == 0) return GPSA_OK;
  long nmax = 0;
  for (int k = 0; k < a->count; ++k) nmax = a->n[k] > nmax ? a->n[k] : nmax;
  adam_tick_kernel<<<1, GPSA_ADAM_MAX_TENSORS, 0, st>>>(*a);
  GPSA_LAUNCH_CHECK();
  long bx = (nmax / 4 + 255) / 256;
  if (bx > 148 * 4) bx = 148 * 4;
  if (bx < 1) bx = 1;
  adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)a->count), 256, 0, st>>>(*a);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_kmeans_lloyd(long N, int D, int K, const float* X, float* centres, int iters, int* assign,
                                 double* sums, double* inertia, cudaStream_t st) {
  if (N <= 0 || K <= 0 || D < 1 || D > 3 || iters < 0) return GPSA_ERR_ARG;
  if ((size_t)K * D * sizeof(float) > 48 * 1024) return GPSA_ERR_UNSUPPORTED;
  long b = (N + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  const size_t smem = (size_t)K * D * sizeof(float);
  for (int it = 0; it <= iters; ++it) {  // the last pass only assigns (inertia of the final centres)
    if (cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)K * (D + 1), st) != cudaSuccess ||
        cudaMemsetAsync(inertia, 0, sizeof(double), st) != cudaSuccess)
      return GPSA_ERR_CUDA;
    if (D == 1) kmeans_assign_kernel<1><<<(int)b, 256, smem, st>>>(N, K, X, centres, assign, sums, inertia);
    else if (D == 2) kmeans_assign_kernel<2><<<(int)b, 256, smem, st>>>(N, K, X, centres, assign, sums, inertia);
    else kmeans_assign_kernel<3><<<(int)b, 256, smem, st>>>(N, K, X, centres, assign, sums, inertia);
    GPSA_LAUNCH_CHECK();
    if (it == iters) break;
    if (D == 1) kmeans_update_kernel<1><<<gpsa_cdiv(K, 128), 128, 0, st>>>(K, sums, centres);
    else if (D == 2) kmeans_update_kernel<2><<<gpsa_cdiv(K, 128), 128, 0, st>>>(K, sums, centres);
    else kmeans_update_kernel<3><<<gpsa_cdiv(K, 128), 128, 0, st>>>(K, sums, centres);
    GPSA_LAUNCH_CHECK();
  }
  return GPSA_OK;
}
