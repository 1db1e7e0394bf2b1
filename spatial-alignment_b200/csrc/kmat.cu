// Fused covariance-function evaluation and its gradients.
//
// Replaces the chains of elementwise ATen kernels behind
//   rbf_kernel       reference gpsa/util/util.py:8-23
//   matern12_kernel  reference gpsa/util/util.py:33-47
//   matern32_kernel  reference gpsa/util/util.py:50-66
// as called from gpsa/models/vgpsa.py:314-318 (warp K_uu, K_uf) and :390,:409 (data K_uu, K_uf).
// Coordinates are D <= 3 floats per point and live in registers; no [M,R,D] difference tensor
// and no distance matrix is ever materialised.  Parameters are log-scale device scalars, so
// nothing here synchronises with the host.
//
// Layout: x1 [M,D], x2 [R,D] row-major, K [M,R] row-major (R contiguous).
#include "common.cuh"
#include "gpsa_b200.h"

namespace {

template <int D>
struct Pt {
  float v[D];
};

template <int D>
__device__ __forceinline__ Pt<D> load_pt(const float* p, long i) {
  Pt<D> r;
#pragma unroll
  for (int d = 0; d < D; ++d) r.v[d] = p[i * D + d];
  return r;
}

// k(x,z) from the squared distance.  RBF: var*exp(-0.5*r2/ls^2).  Matern-1/2 as the reference
// defines it: var*exp(-0.5*sqrt(r2+1e-10)/ls)  (eps inside the sqrt, extra factor 0.5).
template <int KIND>
__device__ __forceinline__ float kval(float r2, float inv_ls, float var) {
  if (KIND == GPSA_KIND_RBF) return var * expf(-0.5f * r2 * inv_ls * inv_ls);
  if (KIND == GPSA_KIND_MATERN32) {  // var (1 + t) exp(-t), t = sqrt(3) sqrt(r2 + 1e-10) / ls
    const float t = 1.7320508075688772f * sqrtf(r2 + 1e-10f) * inv_ls;
    return var * (1.f + t) * expf(-t);
  }
  return var * expf(-0.5f * sqrtf(r2 + 1e-10f) * inv_ls);
}

constexpr int FWD_MT = 16;  // rows of K per thread

template <int D, int KIND>
__global__ void __launch_bounds__(256) kmat_fwd_kernel(int M, long R, const float* __restrict__ x1,
                                                       const float* __restrict__ x2,
                                                       const float* __restrict__ log_ls,
                                                       const float* __restrict__ log_var, float* __restrict__ K) {
  __shared__ float z[FWD_MT * D];
  const int m0 = blockIdx.y * FWD_MT;
  if (threadIdx.x < FWD_MT * D) {
    const int mm = threadIdx.x / D, d = threadIdx.x % D;
    z[threadIdx.x] = (m0 + mm < M) ? x1[(long)(m0 + mm) * D + d] : 0.f;
  }
  __syncthreads();
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float inv_ls = expf(-log_ls[0]), var = expf(log_var[0]);
  const Pt<D> x = load_pt<D>(x2, r);
#pragma unroll
  for (int mm = 0; mm < FWD_MT; ++mm) {
    if (m0 + mm >= M) break;
    float r2 = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float t = z[mm * D + d] - x.v[d];
      r2 = fmaf(t, t, r2);
    }
    K[(long)(m0 + mm) * R + r] = kval<KIND>(r2, inv_ls, var);
  }
}

// 128-bit variant (R % 4 == 0, 16-byte aligned x2 and K): one thread = four consecutive columns r, their 4*D
// coordinates arrive as D float4 loads, every row of K leaves as one float4 store.
template <int D, int KIND>
__global__ void __launch_bounds__(256) kmat_fwd4_kernel(int M, long R4, const float* __restrict__ x1,
                                                        const float4* __restrict__ x2, const float* __restrict__ log_ls,
                                                        const float* __restrict__ log_var, float4* __restrict__ K) {
  __shared__ float z[FWD_MT * D];
  const int m0 = blockIdx.y * FWD_MT;
  if (threadIdx.x < FWD_MT * D) {
    const int mm = threadIdx.x / D, d = threadIdx.x % D;
    z[threadIdx.x] = (m0 + mm < M) ? x1[(long)(m0 + mm) * D + d] : 0.f;
  }
  __syncthreads();
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;  // columns 4q .. 4q+3
  if (q >= R4) return;
  const float inv_ls = expf(-log_ls[0]), var = expf(log_var[0]);
  float c[4 * D];  // c[j*D + d] = x2[4q + j][d]
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float4 v = x2[q * D + i];
    c[4 * i] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w;
  }
#pragma unroll
  for (int mm = 0; mm < FWD_MT; ++mm) {
    if (m0 + mm >= M) break;
    float k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r2 = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float t = z[mm * D + d] - c[j * D + d];
        r2 = fmaf(t, t, r2);
      }
      k[j] = kval<KIND>(r2, inv_ls, var);
    }
    K[(long)(m0 + mm) * R4 + q] = make_float4(k[0], k[1], k[2], k[3]);
  }
}

// dK/d(stuff) pieces shared by both backward phases.  Returns K and writes
//   coef : dK/dx1_d = -coef * (x1_d - x2_d),  dK/dx2_d = +coef * (x1_d - x2_d)
//   dls  : dK/dlog_ls
template <int KIND>
__device__ __forceinline__ float kgrad(float r2, float inv_ls, float var, float& coef, float& dls) {
  if (KIND == GPSA_KIND_RBF) {
    const float k = var * expf(-0.5f * r2 * inv_ls * inv_ls);
    coef = k * inv_ls * inv_ls;
    dls = coef * r2;
    return k;
  }
  if (KIND == GPSA_KIND_MATERN32) {
    // dk/dr = -var (sqrt3/ls) t e^-t  ->  dk/dx1 = -(3 var e^-t / ls^2) (x1 - x2);  dk/dlog_ls = var t^2 e^-t
    const float t = 1.7320508075688772f * sqrtf(r2 + 1e-10f) * inv_ls;
    const float e = var * expf(-t);
    coef = 3.f * e * inv_ls * inv_ls;
    dls = e * t * t;
    return e * (1.f + t);
  }
  const float t = sqrtf(r2 + 1e-10f);
  const float k = var * expf(-0.5f * t * inv_ls);
  coef = 0.5f * k * inv_ls / t;
  dls = 0.5f * k * t * inv_ls;
  return k;
}

// Phase A: one thread per column r, loops over all rows m -> x2bar[r,:].
// If acc_x2 != nullptr the result is atomically added there (double, K(Z,Z) case where x2 is the
// same parameter as x1) instead of being written to x2bar.
template <int D, int KIND>
__global__ void __launch_bounds__(256) kmat_bwd_cols_kernel(int M, long R, const float* __restrict__ x1,
                                                            const float* __restrict__ x2,
                                                            const float* __restrict__ log_ls,
                                                            const float* __restrict__ log_var,
                                                            const float* __restrict__ Kbar,
                                                            float* __restrict__ x2bar, double* acc_x2) {
  extern __shared__ float z[];  // [M*D]
  for (int i = threadIdx.x; i < M * D; i += blockDim.x) z[i] = x1[i];
  __syncthreads();
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float inv_ls = expf(-log_ls[0]), var = expf(log_var[0]);
  const Pt<D> x = load_pt<D>(x2, r);
  float g[D];
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = 0.f;
  for (int m = 0; m < M; ++m) {
    float dd[D], r2 = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      dd[d] = z[m * D + d] - x.v[d];
      r2 = fmaf(dd[d], dd[d], r2);
    }
    float coef, dls;
    kgrad<KIND>(r2, inv_ls, var, coef, dls);
    const float w = Kbar[(long)m * R + r] * coef;
#pragma unroll
    for (int d = 0; d < D; ++d) g[d] = fmaf(w, dd[d], g[d]);
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (acc_x2) atomicAdd(&acc_x2[r * D + d], (double)g[d]);
    else x2bar[r * D + d] = g[d];
  }
}

// Phase B: one warp per row m, lanes stride over a chunk of columns -> x1bar[m,:] and the two
// hyper-parameter gradients; partial sums go to double accumulators.
template <int D, int KIND>
__global__ void __launch_bounds__(256) kmat_bwd_rows_kernel(int M, long R, long chunk, const float* __restrict__ x1,
                                                            const float* __restrict__ x2,
                                                            const float* __restrict__ log_ls,
                                                            const float* __restrict__ log_var,
                                                            const float* __restrict__ Kbar, double* acc_x1,
                                                            double* acc_hyp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + warp;
  const long r_begin = (long)blockIdx.y * chunk;
  const long r_end = (r_begin + chunk < R) ? r_begin + chunk : R;
  float g[D], gls = 0.f, gvar = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = 0.f;
  if (m < M) {
    const float inv_ls = expf(-log_ls[0]), var = expf(log_var[0]);
    const Pt<D> zc = load_pt<D>(x1, m);
    for (long r = r_begin + lane; r < r_end; r += 32) {
      float dd[D], r2 = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        dd[d] = zc.v[d] - x2[r * D + d];
        r2 = fmaf(dd[d], dd[d], r2);
      }
      float coef, dls;
      const float k = kgrad<KIND>(r2, inv_ls, var, coef, dls);
      const float kb = Kbar[(long)m * R + r];
      const float w = -kb * coef;
#pragma unroll
      for (int d = 0; d < D; ++d) g[d] = fmaf(w, dd[d], g[d]);
      gls = fmaf(kb, dls, gls);
      gvar = fmaf(kb, k, gvar);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = warp_sum(g[d]);
  gls = warp_sum(gls);
  gvar = warp_sum(gvar);
  if (lane == 0 && m < M) {
#pragma unroll
    for (int d = 0; d < D; ++d) atomicAdd(&acc_x1[m * D + d], (double)g[d]);
  }
  // one atomic per CTA for the two scalars
  __shared__ float s_ls[8], s_var[8];
  if (lane == 0) { s_ls[warp] = gls; s_var[warp] = gvar; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_ls[w]; b += s_var[w]; }
    atomicAdd(&acc_hyp[0], a);
    atomicAdd(&acc_hyp[1], b);
  }
}

template <int D, int KIND>
int launch_fwd(int M, long R, const float* x1, const float* x2, const float* ls, const float* var, float* K,
               cudaStream_t st) {
  if ((R & 3) == 0 && R >= 1024 && ((reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(K)) & 15) == 0) {
    dim3 grid(gpsa_cdiv(R / 4, 256), gpsa_cdiv(M, FWD_MT));
    kmat_fwd4_kernel<D, KIND><<<grid, 256, 0, st>>>(M, R / 4, x1, reinterpret_cast<const float4*>(x2), ls, var,
                                                    reinterpret_cast<float4*>(K));
  } else {
    dim3 grid(gpsa_cdiv(R, 256), gpsa_cdiv(M, FWD_MT));
    kmat_fwd_kernel<D, KIND><<<grid, 256, 0, st>>>(M, R, x1, x2, ls, var, K);
  }
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

template <int D, int KIND>
int launch_bwd(int M, long R, const float* x1, const float* x2, const float* ls, const float* var, const float* Kbar,
               double* acc_x1, float* x2bar, double* acc_x2, double* acc_hyp, cudaStream_t st) {
  if (x2bar || acc_x2) {
    const size_t smem = (size_t)M * D * sizeof(float);
    kmat_bwd_cols_kernel<D, KIND><<<gpsa_cdiv(R, 256), 256, smem, st>>>(M, R, x1, x2, ls, var, Kbar, x2bar, acc_x2);
    GPSA_LAUNCH_CHECK();
  }
  // enough column chunks to fill the machine: ~148*8 warps-of-work CTAs
  const int row_ctas = gpsa_cdiv(M, 8);
  long nchunk = (148 * 4 + row_ctas - 1) / row_ctas;
  const long max_chunks = (R + 1023) / 1024;
  if (nchunk > max_chunks) nchunk = max_chunks;
  if (nchunk < 1) nchunk = 1;
  const long chunk = ((R + nchunk - 1) / nchunk + 31) / 32 * 32;
  dim3 grid(row_ctas, gpsa_cdiv(R, chunk));
  kmat_bwd_rows_kernel<D, KIND><<<grid, 256, 0, st>>>(M, R, chunk, x1, x2, ls, var, Kbar, acc_x1, acc_hyp);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

}  // namespace

#define DISPATCH_DK(D, kind, CALL)                                                   \
  do {                                                                               \
    if (kind == GPSA_KIND_RBF) {                                                     \
      if (D == 1) return CALL(1, GPSA_KIND_RBF);                                     \
      if (D == 2) return CALL(2, GPSA_KIND_RBF);                                     \
      if (D == 3) return CALL(3, GPSA_KIND_RBF);                                     \
    } else if (kind == GPSA_KIND_MATERN12) {                                         \
      if (D == 1) return CALL(1, GPSA_KIND_MATERN12);                                \
      if (D == 2) return CALL(2, GPSA_KIND_MATERN12);                                \
      if (D == 3) return CALL(3, GPSA_KIND_MATERN12);                                \
    } else if (kind == GPSA_KIND_MATERN32) {                                         \
      if (D == 1) return CALL(1, GPSA_KIND_MATERN32);                                \
      if (D == 2) return CALL(2, GPSA_KIND_MATERN32);                                \
      if (D == 3) return CALL(3, GPSA_KIND_MATERN32);                                \
    }                                                                                \
    return GPSA_ERR_UNSUPPORTED;                                                     \
  } while (0)

extern "C" int gpsa_kernel_matrix_fwd(int kind, int D, int M, long R, const float* x1, const float* x2,
                                      const float* log_ls, const float* log_var, float* K, cudaStream_t st) {
  if (M <= 0 || R <= 0) return GPSA_OK;
#define CALL(DD, KK) launch_fwd<DD, KK>(M, R, x1, x2, log_ls, log_var, K, st)
  DISPATCH_DK(D, kind, CALL);
#undef CALL
}

extern "C" int gpsa_kernel_matrix_bwd(int kind, int D, int M, long R, const float* x1, const float* x2,
                                      const float* log_ls, const float* log_var, const float* Kbar, double* acc_x1,
                                      float* x2bar, double* acc_x2, double* acc_hyp, cudaStream_t st) {
  if (M <= 0 || R <= 0) return GPSA_OK;
  if ((size_t)M * D * sizeof(float) > 48 * 1024) return GPSA_ERR_UNSUPPORTED;
#define CALL(DD, KK) launch_bwd<DD, KK>(M, R, x1, x2, log_ls, log_var, Kbar, acc_x1, x2bar, acc_x2, acc_hyp, st)
  DISPATCH_DK(D, kind, CALL);
#undef CALL
}
