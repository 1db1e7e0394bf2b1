// Layer-level entry points of the ELBO hot path: prior / variational covariance preparation, the
// warp layer (one view), the data layer (one modality) and the Gaussian log-likelihood, each as
// an explicit forward and an explicit analytic backward (no autograd tape below this boundary).
// Every function only enqueues kernels on the caller's stream.
//
// Precision policy.  Data, parameters and every R- or R x L-sized tensor are fp32 like the
// reference.  Everything M x M (factorisations, inverses, the K-bar assembly) and every product
// that applies K^-1 accumulates in fp64: with the reference's default RBF initialisation
// cond(K_uu) ~ 1e7 ~ 1/eps_fp32, and an explicit fp32 inverse would not be backward stable the way
// the reference's triangular solves are.  That work is O(M^3 + M^2 R), < 2 % of the iteration.
#include "gemm.cuh"
#include "gpsa_b200.h"

#include <math.h>

#define TRY(x)                        \
  do {                                \
    int rc__ = (x);                   \
    if (rc__ != GPSA_OK) return rc__; \
  } while (0)

namespace {

// ------------------------------------------------------------------------------------------------
// GEMM conveniences (row-major)
// ------------------------------------------------------------------------------------------------
int pick_split(int M, int N, long K, int batch, int tile) {
  const long tiles = (long)gpsa_cdiv(M, tile) * gpsa_cdiv(N, tile) * batch;
  if (tiles >= 296 || K < 1024) return 1;
  long s = (592 + tiles - 1) / tiles;
  const long smax = K / 512;
  if (s > smax) s = smax;
  return s < 1 ? 1 : (int)s;
}
template <typename T>
int tile_of(int M, int N) { return (sizeof(T) == 4 && M >= 96 && N >= 96) ? 128 : 64; }

// C = alpha*A*B + beta*C with A [M,K] (lda), B [K,N] (ldb).  T = accumulation type.
template <typename T = float, typename TA = float, typename TB = float, typename TC = float>
int gemm_nn(cudaStream_t st, int M, int N, long K, double alpha, const TA* A, long lda, const TB* B, long ldb,
            double beta, TC* C, long ldc, const float* alpha_dev = nullptr) {
  return gemm_strided<T, TA, TB, TC>(st, M, N, K, alpha, A, lda, 1, 0, B, ldb, 1, 0, beta, C, ldc, 0, 1, 1, 0.0, 0,
                                     alpha_dev, 0);
}
// C = alpha*A*B^T + beta*C with A [M,K] (lda), B [N,K] (ldb).  beta must be 0 or 1; a long K is split.
template <typename T = float, typename TA = float, typename TB = float, typename TC = float>
int gemm_nt(cudaStream_t st, int M, int N, long K, double alpha, const TA* A, long lda, const TB* B, long ldb,
            double beta, TC* C, long ldc, const float* alpha_dev = nullptr) {
  const int split = pick_split(M, N, K, 1, tile_of<T>(M, N));
  if (split > 1 && beta == 0.0) {
    if (ldc != N) return GPSA_ERR_ARG;
    if (cudaMemsetAsync(C, 0, sizeof(TC) * (size_t)M * N, st) != cudaSuccess) return GPSA_ERR_CUDA;
  }
  return gemm_strided<T, TA, TB, TC>(st, M, N, K, alpha, A, lda, 1, 0, B, 1, ldb, 0, beta, C, ldc, 0, 1, split, 0.0,
                                     0, alpha_dev, 0);
}

int grid_for(long n, int per_block = 256, int cap = 148 * 16) {
  const long b = (n + per_block - 1) / per_block;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// ------------------------------------------------------------------------------------------------
// prior covariance in fp64
// ------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double kval64(double r2, double inv_ls, double var) {
  if (KIND == GPSA_KIND_RBF) return var * exp(-0.5 * r2 * inv_ls * inv_ls);
  if (KIND == GPSA_KIND_MATERN32) {
    const double t = 1.7320508075688772 * sqrt(r2 + 1e-10) * inv_ls;
    return var * (1.0 + t) * exp(-t);
  }
  return var * exp(-0.5 * sqrt(r2 + 1e-10) * inv_ls);
}

// K and the pieces of its gradient in fp64 (same conventions as kgrad in kmat.cu):
//   dK/dz_d = -coef (z_d - x_d),  dK/dlog_ls = dls
template <int KIND>
__device__ __forceinline__ double kgrad64(double r2, double inv_ls, double var, double& coef, double& dls) {
  if (KIND == GPSA_KIND_RBF) {
    const double k = var * exp(-0.5 * r2 * inv_ls * inv_ls);
    coef = k * inv_ls * inv_ls;
    dls = coef * r2;
    return k;
  }
  if (KIND == GPSA_KIND_MATERN32) {
    const double t = 1.7320508075688772 * sqrt(r2 + 1e-10) * inv_ls;
    const double e = var * exp(-t);
    coef = 3.0 * e * inv_ls * inv_ls;
    dls = e * t * t;
    return e * (1.0 + t);
  }
  const double t = sqrt(r2 + 1e-10);
  const double k = var * exp(-0.5 * t * inv_ls);
  coef = 0.5 * k * inv_ls / t;
  dls = 0.5 * k * t * inv_ls;
  return k;
}

template <int D, int KIND>
__global__ void prior_kuu_kernel(int M, const float* __restrict__ Z, const float* __restrict__ log_ls,
                                 const float* __restrict__ log_var, double* __restrict__ K) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * M) return;
  const int i = idx / M, j = idx % M;
  const double inv_ls = exp(-(double)log_ls[0]), var = exp((double)log_var[0]);
  double r2 = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double t = (double)Z[i * D + d] - (double)Z[j * D + d];
    r2 += t * t;
  }
  double k = kval64<KIND>(r2, inv_ls, var);
  if (i == j) k += 1e-5;
  K[idx] = k;
}

// Gradient of sum(Kbar o k(Z,Z)) with Kbar in fp64 (both argument roles of Z): one warp per row i.
//   Zbar_i = -sum_j (Kbar_ij + Kbar_ji) coef_ij (z_i - z_j);  hyp += sum_ij Kbar_ij (dK/dlog_ls, K)
template <int D, int KIND>
__global__ void __launch_bounds__(256) prior_bwd_kernel(int M, const float* __restrict__ Z,
                                                        const float* __restrict__ log_ls,
                                                        const float* __restrict__ log_var,
                                                        const double* __restrict__ Kbar, double* acc_Z,
                                                        double* acc_hyp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  const double inv_ls = exp(-(double)log_ls[0]), var = exp((double)log_var[0]);
  double g[D], gls = 0, gvar = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = 0;
  if (i < M) {
    double zi[D];
#pragma unroll
    for (int d = 0; d < D; ++d) zi[d] = (double)Z[i * D + d];
    for (int j = lane; j < M; j += 32) {
      double dd[D], r2 = 0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        dd[d] = zi[d] - (double)Z[j * D + d];
        r2 += dd[d] * dd[d];
      }
      double coef, dls;
      const double k = kgrad64<KIND>(r2, inv_ls, var, coef, dls);
      const double kij = Kbar[(long)i * M + j], kji = Kbar[(long)j * M + i];
      const double w = -(kij + kji) * coef;
#pragma unroll
      for (int d = 0; d < D; ++d) g[d] += w * dd[d];
      gls += kij * dls;
      gvar += kij * k;
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = warp_sum(g[d]);
  gls = warp_sum(gls);
  gvar = warp_sum(gvar);
  if (lane == 0 && i < M) {
#pragma unroll
    for (int d = 0; d < D; ++d) atomicAdd(&acc_Z[i * D + d], g[d]);
    atomicAdd(&acc_hyp[0], gls);
    atomicAdd(&acc_hyp[1], gvar);
  }
}

// K_uf in fp64 for the warp layer: B[m,r] = k(Z[m], X[r]).  The warp GP's interpolation weights
// A = K^-1 B reach 1e3 with the reference's default lengthscale, so an fp32-rounded B would already
// cost 1e-3 in the marginal variance; the layer is O(V M^2 n) and is kept in fp64 end to end.
template <int D, int KIND>
__global__ void __launch_bounds__(256) kuf64_kernel(int M, long n, const float* __restrict__ Z,
                                                    const float* __restrict__ X, const float* __restrict__ log_ls,
                                                    const float* __restrict__ log_var, double* __restrict__ B) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double inv_ls = exp(-(double)log_ls[0]), var = exp((double)log_var[0]);
  double x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = (double)X[r * D + d];
  for (int m = blockIdx.y; m < M; m += gridDim.y) {
    double r2 = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double t = (double)Z[m * D + d] - x[d];
      r2 += t * t;
    }
    B[(long)m * n + r] = kval64<KIND>(r2, inv_ls, var);
  }
}

// Gradient of sum(Bbar o k(Z, X)) w.r.t. Z and the two hyper-parameters, fp64: one warp per row m.
template <int D, int KIND>
__global__ void __launch_bounds__(256) kuf64_bwd_kernel(int M, long n, long chunk, const float* __restrict__ Z,
                                                        const float* __restrict__ X,
                                                        const float* __restrict__ log_ls,
                                                        const float* __restrict__ log_var,
                                                        const double* __restrict__ Bbar, double* acc_Z,
                                                        double* acc_hyp) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const long r0 = (long)blockIdx.y * chunk, r1 = (r0 + chunk < n) ? r0 + chunk : n;
  const double inv_ls = exp(-(double)log_ls[0]), var = exp((double)log_var[0]);
  double z[D], g[D], gls = 0, gvar = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) { z[d] = (double)Z[m * D + d]; g[d] = 0; }
  for (long r = r0 + lane; r < r1; r += 32) {
    double dd[D], r2 = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      dd[d] = z[d] - (double)X[r * D + d];
      r2 += dd[d] * dd[d];
    }
    double coef, dls;
    const double k = kgrad64<KIND>(r2, inv_ls, var, coef, dls);
    const double kb = Bbar[(long)m * n + r];
    const double w = -kb * coef;
#pragma unroll
    for (int d = 0; d < D; ++d) g[d] += w * dd[d];
    gls += kb * dls;
    gvar += kb * k;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = warp_sum(g[d]);
  gls = warp_sum(gls);
  gvar = warp_sum(gvar);
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) atomicAdd(&acc_Z[m * D + d], g[d]);
    atomicAdd(&acc_hyp[0], gls);
    atomicAdd(&acc_hyp[1], gvar);
  }
}

// dst[b,i,j] = (float) src[b, max(i,j), min(i,j)]: fp32 copy of a symmetric matrix whose lower triangle is valid
__global__ void cvt_sym_kernel(long n, int M, const double* __restrict__ src, float* __restrict__ dst) {
  const long MM = (long)M * M;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const long b = idx / MM;
    const int e = (int)(idx - b * MM);
    const int i = e / M, j = e - i * M;
    dst[idx] = (float)src[b * MM + (i >= j ? (long)i * M + j : (long)j * M + i)];
  }
}

template <typename TS, typename TD>
__global__ void cvt_kernel(long n, const TS* __restrict__ src, TD* __restrict__ dst) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dst[i] = (TD)src[i];
}

// ------------------------------------------------------------------------------------------------
// small elementwise / reduction kernels
// ------------------------------------------------------------------------------------------------
// y += alpha * (*alpha_dev) * x
template <typename TX, typename TY>
__global__ void axpy_dev_kernel(long n, double alpha, const float* __restrict__ alpha_dev, const TX* __restrict__ x,
                                TY* __restrict__ y) {
  const double a = alpha * (alpha_dev ? (double)alpha_dev[0] : 1.0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = (TY)((double)y[i] + a * (double)x[i]);
}
template <typename TX, typename TY>
int axpy_dev(cudaStream_t st, long n, double alpha, const float* alpha_dev, const TX* x, TY* y) {
  axpy_dev_kernel<TX, TY><<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(n, alpha, alpha_dev, x, y);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

// S[e] = sum_{b<count} X[b*stride + e]   (fp64 accumulate)
__global__ void sum_batch_kernel(long n, int count, long stride, const float* __restrict__ X, double* __restrict__ S) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double s = 0.0;
  for (int b = 0; b < count; ++b) s += (double)X[(long)b * stride + e];
  S[e] = s;
}

// kq[r] = sigma2 - sum_m A[m,r] B[m,r]   (K_ff - a^T K a, fp64 accumulate, rounded once)
__global__ void kq_kernel(int M, long R, const float* __restrict__ A, const float* __restrict__ B,
                          const float* __restrict__ log_var, float* __restrict__ kq) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double s = 0.0;
  for (int m = 0; m < M; ++m) s += (double)A[(long)m * R + r] * (double)B[(long)m * R + r];
  kq[r] = (float)(exp((double)log_var[0]) - s);
}

// out[m,r] = q[r] * X[m,r]   (accumulate = 0)   or   out[m,r] += q[r] * X[m,r]  (accumulate = 1)
template <typename T>
__global__ void colscale_kernel(long total, long R, const float* __restrict__ q, const T* __restrict__ X,
                                T* __restrict__ out, int accumulate) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const T v = (T)q[i % R] * X[i];
    out[i] = accumulate ? out[i] + v : v;
  }
}
template <typename T>
int colscale(cudaStream_t st, int M, long R, const float* q, const T* X, T* out, int accumulate) {
  const long total = (long)M * R;
  colscale_kernel<T><<<grid_for(total), 256, 0, st>>>(total, R, q, X, out, accumulate);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

// ------------------------------------------------------------------------------------------------
// data layer: sampling and its backward
// ------------------------------------------------------------------------------------------------
// in:  F = predictive mean, var = q2.   out: var = (sigma2 - q1) + q2 + 2 off, F = mean + sqrt(var) eps
__global__ void sample_fwd_kernel(long total, int L, const float* __restrict__ kq, const float* __restrict__ eps,
                                  float* __restrict__ F, float* __restrict__ var) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / L;
    const float v = (kq[r] + var[i] + GPSA_OFF) + GPSA_OFF;  // jitter twice: vgpsa.py:201,:204
    var[i] = v;
    F[i] = fmaf(sqrtf(v), eps[i], F[i]);
  }
}
// same, one CTA per row r at a time (L >= 128): no 64-bit division per element
__global__ void __launch_bounds__(256) sample_fwd_rows_kernel(long R, int L, const float* __restrict__ kq,
                                                              const float* __restrict__ eps, float* __restrict__ F,
                                                              float* __restrict__ var) {
  for (long r = blockIdx.x; r < R; r += gridDim.x) {
    const float k = kq[r];
    const long base = r * L;
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
      const float v = (k + var[base + p] + GPSA_OFF) + GPSA_OFF;
      var[base + p] = v;
      F[base + p] = fmaf(sqrtf(v), eps[base + p], F[base + p]);
    }
  }
}

// 128-bit variant (L % 4 == 0, 16-byte aligned rows): four genes per thread and access
__global__ void __launch_bounds__(256) sample_fwd_rows4_kernel(long R, int L4, const float* __restrict__ kq,
                                                               const float4* __restrict__ eps, float4* __restrict__ F,
                                                               float4* __restrict__ var) {
  for (long r = blockIdx.x; r < R; r += gridDim.x) {
    const float k = kq[r];
    const long base = r * L4;
    for (int p = threadIdx.x; p < L4; p += blockDim.x) {
      float4 v = var[base + p];
      const float4 e = eps[base + p];
      float4 f = F[base + p];
      v.x = (k + v.x + GPSA_OFF) + GPSA_OFF; v.y = (k + v.y + GPSA_OFF) + GPSA_OFF;
      v.z = (k + v.z + GPSA_OFF) + GPSA_OFF; v.w = (k + v.w + GPSA_OFF) + GPSA_OFF;
      f.x = fmaf(sqrtf(v.x), e.x, f.x); f.y = fmaf(sqrtf(v.y), e.y, f.y);
      f.z = fmaf(sqrtf(v.z), e.z, f.z); f.w = fmaf(sqrtf(v.w), e.w, f.w);
      var[base + p] = v;
      F[base + p] = f;
    }
  }
}

// one warp per row r: Gm[r,p] = Fbar*eps/(2 sqrt(var)) = dLoss/dq2; kq_bar[r] = sum_p Gm[r,p] = dLoss/dkq
__global__ void __launch_bounds__(256) sample_bwd_kernel(long R, int L, const float* __restrict__ Fbar,
                                                         const float* __restrict__ eps, const float* __restrict__ var,
                                                         float* __restrict__ Gm, float* __restrict__ kq_bar) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  for (long r = (long)blockIdx.x * (blockDim.x >> 5) + warp; r < R; r += nwarps) {
    float s = 0.f;
    for (int p = lane; p < L; p += 32) {
      const long i = r * L + p;
      const float g = 0.5f * Fbar[i] * eps[i] / sqrtf(var[i]);
      Gm[i] = g;
      s += g;
    }
    s = warp_sum(s);
    if (lane == 0) kq_bar[r] = s;
  }
}

// 128-bit variant of sample_bwd_kernel (L % 4 == 0)
__global__ void __launch_bounds__(256) sample_bwd4_kernel(long R, int L4, const float4* __restrict__ Fbar,
                                                          const float4* __restrict__ eps, const float4* __restrict__ var,
                                                          float4* __restrict__ Gm, float* __restrict__ kq_bar) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  for (long r = (long)blockIdx.x * (blockDim.x >> 5) + warp; r < R; r += nwarps) {
    float s = 0.f;
    for (int p = lane; p < L4; p += 32) {
      const long i = r * L4 + p;
      const float4 fb = Fbar[i], e = eps[i], v = var[i];
      float4 g;
      g.x = 0.5f * fb.x * e.x / sqrtf(v.x); g.y = 0.5f * fb.y * e.y / sqrtf(v.y);
      g.z = 0.5f * fb.z * e.z / sqrtf(v.z); g.w = 0.5f * fb.w * e.w / sqrtf(v.w);
      Gm[i] = g;
      s += (g.x + g.y) + (g.z + g.w);
    }
    s = warp_sum(s);
    if (lane == 0) kq_bar[r] = s;
  }
}

// kq[r] = sigma2 - q1[r]:  q1bar[r] = -kq_bar[r];  acc_hyp[1] += sigma2 * sum_r kq_bar[r]
__global__ void __launch_bounds__(256) kq_bwd_kernel(long R, const float* __restrict__ kq_bar,
                                                     const float* __restrict__ log_var, float* __restrict__ q1bar,
                                                     double* acc_hyp) {
  double tot = 0.0;
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (long)gridDim.x * blockDim.x) {
    const float k = kq_bar[r];
    q1bar[r] = -k;
    tot += (double)k;
  }
  __shared__ double red[32];
  tot = block_sum<double>(tot, red);
  if (threadIdx.x == 0) atomicAdd(&acc_hyp[1], tot * exp((double)log_var[0]));
}

// KL(q(u_p) || p(u)) summed over genes: one CTA per gene.
//   kl_p = hldK - hldOm[p] + 0.5 (tr(K^-1 Omega_p) + delta_p^T K^-1 delta_p - M)
__global__ void __launch_bounds__(256) kl_F_kernel(int M, int L, const double* __restrict__ Kinv,
                                                   const float* __restrict__ Omega, const double* __restrict__ hldOm,
                                                   const float* __restrict__ dlt, const double* __restrict__ KD,
                                                   const double* __restrict__ hldK, double* kl_acc) {
  const int p = blockIdx.x;
  const float* Om = Omega + (long)p * M * M;
  double s = 0.0;
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) s += Kinv[i] * (double)Om[i];
  for (int m = threadIdx.x; m < M; m += blockDim.x) s += (double)dlt[(long)m * L + p] * KD[(long)m * L + p];
  __shared__ double red[32];
  s = block_sum<double>(s, red);
  if (threadIdx.x == 0) atomicAdd(kl_acc, hldK[0] - hldOm[p] + 0.5 * (s - (double)M));
}

// ------------------------------------------------------------------------------------------------
// warp layer kernels
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) warp_predict_kernel(int M, long n, int S, const float* __restrict__ Z,
                                                           const float* __restrict__ dlt, const double* __restrict__ A,
                                                           const double* __restrict__ B, const double* __restrict__ T,
                                                           const float* __restrict__ X, const float* __restrict__ eps,
                                                           const float* __restrict__ log_var, float* __restrict__ var,
                                                           float* __restrict__ Gmean, float* __restrict__ Gs,
                                                           long gs_stride) {
  extern __shared__ double dmz[];  // [M*D]  delta - mu_z  (mean function = identity: mu_z = Z)
  for (int i = threadIdx.x; i < M * D; i += blockDim.x) dmz[i] = (double)dlt[i] - (double)Z[i];
  __syncthreads();
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double q1 = 0.0, mu[D], q2[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { mu[d] = (double)X[r * D + d]; q2[d] = 0.0; }
  for (int m = 0; m < M; ++m) {
    const double a = A[(long)m * n + r];
    q1 += a * B[(long)m * n + r];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      mu[d] += a * dmz[m * D + d];
      q2[d] += a * T[((long)d * M + m) * n + r];
    }
  }
  const double kq = exp((double)log_var[0]) - q1;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double v = kq + q2[d] + 2e-5;  // jitter added twice (vgpsa.py:191,:204)
    var[r * D + d] = (float)v;
    Gmean[r * D + d] = (float)mu[d];
    for (int s = 0; s < S; ++s)  // the reference uses the VARIANCE as the Normal scale (vgpsa.py:334-340)
      Gs[(long)s * gs_stride + r * D + d] = (float)(mu[d] + v * (double)eps[((long)s * n + r) * D + d]);
  }
}

// KL terms of view v: one CTA per spatial dim j.  Stores Ke[j,:] = K^-1 (Z_j - delta_j) in fp64.
__global__ void __launch_bounds__(256) kl_G_kernel(int M, int D, int V, int v, const double* __restrict__ Kinv,
                                                   const float* __restrict__ Omega_G,
                                                   const double* __restrict__ hldOm, const float* __restrict__ Z,
                                                   const float* __restrict__ dlt, const double* __restrict__ hldK,
                                                   double* __restrict__ Ke, double* kl_acc) {
  const int j = blockIdx.x;
  const int slot = j * V + v;  // the KL uses slice j*V+v (vgpsa.py:508)
  const float* Om = Omega_G + (long)slot * M * M;
  extern __shared__ double e[];  // [M]
  for (int m = threadIdx.x; m < M; m += blockDim.x) e[m] = (double)Z[m * D + j] - (double)dlt[m * D + j];
  __syncthreads();
  double s = 0.0;
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) s += Kinv[i] * (double)Om[i];
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    double t = 0.0;
    for (int k = 0; k < M; ++k) t += Kinv[(long)k * M + m] * e[k];  // K^-1 = X^T X is bitwise symmetric: coalesced over m
    Ke[(long)j * M + m] = t;
    s += t * e[m];
  }
  __shared__ double red[32];
  s = block_sum<double>(s, red);
  if (threadIdx.x == 0 && kl_acc) atomicAdd(kl_acc, hldK[0] - hldOm[slot] + 0.5 * (s - (double)M));
}

template <int D>
__global__ void __launch_bounds__(256) warp_bwd_prep_kernel(long n, int S, const float* __restrict__ Gs_bar,
                                                            long gs_stride, const float* __restrict__ Gm_bar,
                                                            const float* __restrict__ eps,
                                                            const float* __restrict__ log_var,
                                                            float* __restrict__ mubar, float* __restrict__ varbar,
                                                            float* __restrict__ q1bar, double* acc_hyp) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double tot = 0.0;
  if (r < n) {
    float q = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float mb = Gm_bar ? Gm_bar[r * D + d] : 0.f, vb = 0.f;
      if (Gs_bar) {
        for (int s = 0; s < S; ++s) {
          const float g = Gs_bar[(long)s * gs_stride + r * D + d];
          mb += g;
          vb = fmaf(g, eps[((long)s * n + r) * D + d], vb);
        }
      }
      mubar[r * D + d] = mb;
      varbar[r * D + d] = vb;
      q += vb;
    }
    q1bar[r] = -q;
    tot = (double)q;
  }
  __shared__ double red[32];
  tot = block_sum<double>(tot, red);
  if (threadIdx.x == 0) atomicAdd(&acc_hyp[1], tot * exp((double)log_var[0]));
}

// Abar = (delta - Z) mubar^T + q1bar o B + 2 sum_j varbar_j o T_j ;  AS_j = A o varbar_j   (fp64)
template <int D>
__global__ void warp_abar_kernel(int M, long n, const float* __restrict__ Z, const float* __restrict__ dlt,
                                 const double* __restrict__ A, const double* __restrict__ B,
                                 const double* __restrict__ T, const float* __restrict__ mubar,
                                 const float* __restrict__ varbar, const float* __restrict__ q1bar,
                                 double* __restrict__ Abar, double* __restrict__ AS) {
  const long total = (long)M * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int m = i / n;
    const long r = i % n;
    const double a = A[i];
    double s = (double)q1bar[r] * B[i];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double vb = (double)varbar[r * D + d];
      s += ((double)dlt[m * D + d] - (double)Z[m * D + d]) * (double)mubar[r * D + d];
      s += 2.0 * vb * T[(long)d * total + i];
      AS[(long)d * total + i] = a * vb;
    }
    Abar[i] = s;
  }
}

// d(delta - Z)[m,j] = sum_r A[m,r] mubar[r,j]:  +acc_dlt, -acc_Z.  One warp per m, column chunks on y.
template <int D>
__global__ void __launch_bounds__(256) warp_dmz_kernel(int M, long n, long chunk, const double* __restrict__ A,
                                                       const float* __restrict__ mubar, double* acc_dlt,
                                                       double* acc_Z) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const long r0 = (long)blockIdx.y * chunk, r1 = (r0 + chunk < n) ? r0 + chunk : n;
  double g[D];
#pragma unroll
  for (int d = 0; d < D; ++d) g[d] = 0.0;
  for (long r = r0 + lane; r < r1; r += 32) {
    const double a = A[(long)m * n + r];
#pragma unroll
    for (int d = 0; d < D; ++d) g[d] += a * (double)mubar[r * D + d];
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double t = warp_sum(g[d]);
    if (lane == 0) {
      atomicAdd(&acc_dlt[m * D + d], t);
      atomicAdd(&acc_Z[m * D + d], -t);
    }
  }
}

// e-bar_j = kl_bar * Ke_j :  acc_Z[:,j] += , acc_dlt[:,j] -=
__global__ void kl_e_bwd_kernel(int M, int D, const double* __restrict__ Ke, const float* __restrict__ kl_bar,
                                double* acc_Z, double* acc_dlt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const int m = i / D, j = i % D;
  const double g = (double)kl_bar[0] * Ke[(long)j * M + m];
  acc_Z[i] += g;
  acc_dlt[i] -= g;
}

// ------------------------------------------------------------------------------------------------
// Gaussian log-likelihood
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ll_fwd_kernel(long N, int P, int S, const float* __restrict__ F,
                                                     const float* __restrict__ Y, const float* __restrict__ log_noise,
                                                     double* ll_acc) {
  const float sigma = expf(log_noise[0]) + GPSA_OFF;  // vgpsa.py:217
  const float inv = 1.f / sigma;
  const long NP = N * P, total = NP * S;
  double acc = 0.0;
  // one thread = one (spot, gene) for all S samples: Y is read once, no 64-bit modulo per element
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < NP; i += (long)gridDim.x * blockDim.x) {
    const float y = Y[i];
    float part = 0.f;
    for (int s = 0; s < S; ++s) {
      const float z = (y - F[(long)s * NP + i]) * inv;
      part = fmaf(z, z, part);
    }
    acc += (double)part;
  }
  __shared__ double red[32];
  acc = block_sum<double>(acc, red);
  if (threadIdx.x == 0) {
    double v = -0.5 * acc;
    if (blockIdx.x == 0) v -= (double)total * (log((double)sigma) + 0.91893853320467274178);
    atomicAdd(ll_acc, v / (double)S);
  }
}

__global__ void __launch_bounds__(256) ll_bwd_kernel(long N, int P, int S, const float* __restrict__ F,
                                                     const float* __restrict__ Y, const float* __restrict__ log_noise,
                                                     const float* __restrict__ ll_bar, float* __restrict__ F_bar,
                                                     double* acc_noise) {
  const float en = expf(log_noise[0]);
  const float sigma = en + GPSA_OFF;
  const float inv = 1.f / sigma;
  const float c = ll_bar[0] * inv * inv / (float)S;
  const long NP = N * P, total = NP * S;
  double acc = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < NP; i += (long)gridDim.x * blockDim.x) {
    const float y = Y[i];
    float part = 0.f;
    for (int s = 0; s < S; ++s) {
      const float d = y - F[(long)s * NP + i];
      F_bar[(long)s * NP + i] = c * d;
      const float z = d * inv;
      part = fmaf(z, z, part);
    }
    acc += (double)part;
  }
  __shared__ double red[32];
  acc = block_sum<double>(acc, red);
  if (threadIdx.x == 0) {
    // dLL/dsigma = sum((y-f)^2/sigma^3 - 1/sigma)/S ; dsigma/dlog_noise = exp(log_noise)
    double v = acc * (double)inv;
    if (blockIdx.x == 0) v -= (double)total * (double)inv;
    atomicAdd(acc_noise, (double)ll_bar[0] * v * (double)en / (double)S);
  }
}

// 128-bit variants (N*P % 4 == 0, 16-byte aligned): one thread = four (spot, gene) entries over all S samples
__global__ void __launch_bounds__(256) ll_fwd4_kernel(long NP4, int S, const float4* __restrict__ F,
                                                      const float4* __restrict__ Y, const float* __restrict__ log_noise,
                                                      double* ll_acc) {
  const float sigma = expf(log_noise[0]) + GPSA_OFF;
  const float inv = 1.f / sigma;
  double acc = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < NP4; i += (long)gridDim.x * blockDim.x) {
    const float4 y = Y[i];
    float part = 0.f;
    for (int s = 0; s < S; ++s) {
      const float4 f = F[(long)s * NP4 + i];
      const float zx = (y.x - f.x) * inv, zy = (y.y - f.y) * inv, zz = (y.z - f.z) * inv, zw = (y.w - f.w) * inv;
      part = fmaf(zx, zx, part); part = fmaf(zy, zy, part); part = fmaf(zz, zz, part); part = fmaf(zw, zw, part);
    }
    acc += (double)part;
  }
  __shared__ double red[32];
  acc = block_sum<double>(acc, red);
  if (threadIdx.x == 0) {
    double v = -0.5 * acc;
    if (blockIdx.x == 0) v -= (double)(NP4 * 4 * S) * (log((double)sigma) + 0.91893853320467274178);
    atomicAdd(ll_acc, v / (double)S);
  }
}

__global__ void __launch_bounds__(256) ll_bwd4_kernel(long NP4, int S, const float4* __restrict__ F,
                                                      const float4* __restrict__ Y, const float* __restrict__ log_noise,
                                                      const float* __restrict__ ll_bar, float4* __restrict__ F_bar,
                                                      double* acc_noise) {
  const float en = expf(log_noise[0]);
  const float sigma = en + GPSA_OFF;
  const float inv = 1.f / sigma;
  const float c = ll_bar[0] * inv * inv / (float)S;
  double acc = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < NP4; i += (long)gridDim.x * blockDim.x) {
    const float4 y = Y[i];
    float part = 0.f;
    for (int s = 0; s < S; ++s) {
      const float4 f = F[(long)s * NP4 + i];
      const float dx = y.x - f.x, dy = y.y - f.y, dz = y.z - f.z, dw = y.w - f.w;
      F_bar[(long)s * NP4 + i] = make_float4(c * dx, c * dy, c * dz, c * dw);
      const float zx = dx * inv, zy = dy * inv, zz = dz * inv, zw = dw * inv;
      part = fmaf(zx, zx, part); part = fmaf(zy, zy, part); part = fmaf(zz, zz, part); part = fmaf(zw, zw, part);
    }
    acc += (double)part;
  }
  __shared__ double red[32];
  acc = block_sum<double>(acc, red);
  if (threadIdx.x == 0) {
    double v = acc * (double)inv;
    if (blockIdx.x == 0) v -= (double)(NP4 * 4 * S) * (double)inv;
    atomicAdd(acc_noise, (double)ll_bar[0] * v * (double)en / (double)S);
  }
}

// prior_bwd launcher
int prior_bwd(int kind, int D, int M, const float* Z, const float* ls, const float* var, const double* Kbar,
              double* acc_Z, double* acc_hyp, cudaStream_t st) {
  const int blocks = gpsa_cdiv(M, 8);
#define PB(DD, KK) prior_bwd_kernel<DD, KK><<<blocks, 256, 0, st>>>(M, Z, ls, var, Kbar, acc_Z, acc_hyp)
  if (kind == GPSA_KIND_RBF) {
    if (D == 1) PB(1, GPSA_KIND_RBF); else if (D == 2) PB(2, GPSA_KIND_RBF); else PB(3, GPSA_KIND_RBF);
  } else if (kind == GPSA_KIND_MATERN12) {
    if (D == 1) PB(1, GPSA_KIND_MATERN12); else if (D == 2) PB(2, GPSA_KIND_MATERN12); else PB(3, GPSA_KIND_MATERN12);
  } else if (kind == GPSA_KIND_MATERN32) {
    if (D == 1) PB(1, GPSA_KIND_MATERN32); else if (D == 2) PB(2, GPSA_KIND_MATERN32); else PB(3, GPSA_KIND_MATERN32);
  } else {
    return GPSA_ERR_UNSUPPORTED;
  }
#undef PB
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

// K-bar (fp64) of a sparse-GP layer:  Kbar = -K^-1 (Abar A^T)  [+ kl_bar * KL terms]
//   P = Abar A^T (fp64 accumulate over R), Kbar = -Kinv P.
template <typename TV>
int kbar_from_solve(cudaStream_t st, int M, long R, const double* Kinv64, const TV* Abar, const TV* A, double* P,
                    double* Kbar) {
  TRY((gemm_nt<double, TV, TV, double>(st, M, M, R, 1.0, Abar, R, A, R, 0.0, P, M)));
  return gemm_nn<double, double, double, double>(st, M, M, M, -1.0, Kinv64, M, P, M, 0.0, Kbar, M);
}

// chunking of a warp-per-row reduction over n columns so that the grid fills the machine
void row_chunks(int M, long n, dim3& grid, long& chunk) {
  const int row_ctas = gpsa_cdiv(M, 8);
  long nchunk = (148 * 4 + row_ctas - 1) / row_ctas;
  const long maxc = (n + 1023) / 1024;
  if (nchunk > maxc) nchunk = maxc;
  if (nchunk < 1) nchunk = 1;
  chunk = ((n + nchunk - 1) / nchunk + 31) / 32 * 32;
  grid = dim3(row_ctas, gpsa_cdiv(n, chunk));
}

}  // namespace

// ================================================================================================
// launch counter and hot-kernel event profiler
// ================================================================================================
long g_gpsa_launches = 0;

namespace {
constexpr int PROF_RING = 1024;
struct ProfSlot {
  cudaEvent_t beg[PROF_RING], end[PROF_RING];
  int n = 0;
  bool made = false;
};
ProfSlot g_prof[GPSA_PROF_SLOTS];
int g_prof_on = 0;
}  // namespace

void gpsa_prof_begin(int slot, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfSlot& p = g_prof[slot];
  if (!p.made) {
    for (int i = 0; i < PROF_RING; ++i) { cudaEventCreate(&p.beg[i]); cudaEventCreate(&p.end[i]); }
    p.made = true;
  }
  if (p.n < PROF_RING) cudaEventRecord(p.beg[p.n], st);
}
void gpsa_prof_end(int slot, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfSlot& p = g_prof[slot];
  if (p.n < PROF_RING) { cudaEventRecord(p.end[p.n], st); ++p.n; }
}

namespace { int g_no_vec4 = 0; }
// testing aid: 1 = run the scalar variants of the sampling / log-likelihood kernels even where the 128-bit ones apply
extern "C" void gpsa_debug_disable_vec4(int off) { g_no_vec4 = off; }
extern "C" long gpsa_launch_count(void) { return g_gpsa_launches; }
extern "C" void gpsa_prof_enable(int on) {
  g_prof_on = on;
  for (int s = 0; s < GPSA_PROF_SLOTS; ++s) g_prof[s].n = 0;
}
// Synchronises the device; returns per slot the number of timed launches and their total milliseconds.
extern "C" int gpsa_prof_read(int* counts, double* total_ms) {
  if (cudaDeviceSynchronize() != cudaSuccess) return GPSA_ERR_CUDA;
  for (int s = 0; s < GPSA_PROF_SLOTS; ++s) {
    double t = 0;
    for (int i = 0; i < g_prof[s].n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_prof[s].beg[i], g_prof[s].end[i]);
      t += ms;
    }
    counts[s] = g_prof[s].n;
    total_ms[s] = t;
    g_prof[s].n = 0;
  }
  return GPSA_OK;
}

// ================================================================================================
// exported entry points
// ================================================================================================
extern "C" int gpsa_version(void) { return 104; }

extern "C" int gpsa_gemm_f32(int M, int N, long K, float alpha, const float* A, long ars, long acs, long sA,
                             const float* B, long brs, long bcs, long sB, float beta, float* C, long ldc, long sC,
                             int batch, cudaStream_t st) {
  return gemm_strided<float, float, float, float>(st, M, N, K, alpha, A, ars, acs, sA, B, brs, bcs, sB, beta, C, ldc,
                                                  sC, batch);
}

namespace {
// Kd (fp64, M x M, jitter already on its diagonal) -> factor, inverse, fp32 copies.  Xd: M*M doubles of scratch.
int prior_factorise(int M, double* Kd, double* Xd, float* Lk, float* Kinv, double* Kinv64, double* half_logdet, int* info,
                    cudaStream_t st) {
  const long MM = (long)M * M;
  TRY(gpsa_potrf_batched_f64(M, 1, Kd, half_logdet, info, st));
  TRY(gpsa_trtri_batched_f64(M, 1, Kd, Xd, st));
  // K^-1 = X^T X
  TRY((gemm_strided<double, double, double, double>(st, M, M, M, 1.0, Xd, 1, M, 0, Xd, M, 1, 0, 0.0, Kinv64, M, 0, 1)));
  cvt_kernel<double, float><<<grid_for(MM), 256, 0, st>>>(MM, Kd, Lk);
  GPSA_LAUNCH_CHECK();
  if (Kinv) {
    cvt_kernel<double, float><<<grid_for(MM), 256, 0, st>>>(MM, Kinv64, Kinv);
    GPSA_LAUNCH_CHECK();
  }
  return GPSA_OK;
}

// Kd = (double) K + 1e-5 I
__global__ void ext_kuu_kernel(int M, const float* __restrict__ K, double* __restrict__ Kd) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * M) return;
  Kd[idx] = (double)K[idx] + ((idx / M == idx % M) ? 1e-5 : 0.0);
}
}  // namespace

extern "C" int gpsa_prior_prepare(int kind, int D, int M, const float* Z, const float* log_ls, const float* log_var,
                                  float* Lk, float* Kinv, double* Kinv64, double* half_logdet, int* info, double* ws64,
                                  cudaStream_t st) {
  if (M <= 0 || D < 1 || D > 3) return GPSA_ERR_ARG;
  const long MM = (long)M * M;
  double* Kd = ws64;
  double* Xd = ws64 + MM;
  const int blocks = gpsa_cdiv(MM, 256);
#define PK(DD, KK) prior_kuu_kernel<DD, KK><<<blocks, 256, 0, st>>>(M, Z, log_ls, log_var, Kd)
  if (kind == GPSA_KIND_RBF) {
    if (D == 1) PK(1, GPSA_KIND_RBF); else if (D == 2) PK(2, GPSA_KIND_RBF); else PK(3, GPSA_KIND_RBF);
  } else if (kind == GPSA_KIND_MATERN12) {
    if (D == 1) PK(1, GPSA_KIND_MATERN12); else if (D == 2) PK(2, GPSA_KIND_MATERN12); else PK(3, GPSA_KIND_MATERN12);
  } else if (kind == GPSA_KIND_MATERN32) {
    if (D == 1) PK(1, GPSA_KIND_MATERN32); else if (D == 2) PK(2, GPSA_KIND_MATERN32); else PK(3, GPSA_KIND_MATERN32);
  } else {
    return GPSA_ERR_UNSUPPORTED;
  }
#undef PK
  GPSA_LAUNCH_CHECK();
  return prior_factorise(M, Kd, Xd, Lk, Kinv, Kinv64, half_logdet, info, st);
}

// Same with K_uu [M,M] (fp32, WITHOUT the jitter) evaluated by the caller: user-supplied covariance functions.
extern "C" int gpsa_prior_prepare_ext(int M, const float* Kuu, float* Lk, float* Kinv, double* Kinv64, double* half_logdet,
                                      int* info, double* ws64, cudaStream_t st) {
  if (M <= 0 || !Kuu) return GPSA_ERR_ARG;
  const long MM = (long)M * M;
  ext_kuu_kernel<<<gpsa_cdiv(MM, 256), 256, 0, st>>>(M, Kuu, ws64);
  GPSA_LAUNCH_CHECK();
  return prior_factorise(M, ws64, ws64 + MM, Lk, Kinv, Kinv64, half_logdet, info, st);
}

namespace {
// in place: upper triangle <- lower triangle, per matrix.  One CTA (32 x 8) per pair of 32 x 32 tiles (ti >= tj) of one
// matrix: the lower tile is read along its rows and its transpose written along the rows of the upper tile (a thread
// per upper element reading A[j][i] directly walks a column: one 32-byte sector per value).
__global__ void __launch_bounds__(256) mirror_lower_kernel(int M, int nt, float* __restrict__ A) {
  __shared__ float t[32][33];
  int ti = 0, rem = blockIdx.x;  // blockIdx.x enumerates the nt (nt + 1) / 2 tile pairs, row by row
  while (rem > ti) { rem -= ti + 1; ++ti; }
  const int tj = rem;
  float* Ab = A + (long)blockIdx.y * M * M;
  const int i0 = ti * 32, j0 = tj * 32;
  for (int yy = threadIdx.y; yy < 32; yy += 8) {
    const int i = i0 + yy, j = j0 + threadIdx.x;
    t[yy][threadIdx.x] = (i < M && j < M) ? Ab[(long)i * M + j] : 0.f;
  }
  __syncthreads();
  for (int yy = threadIdx.y; yy < 32; yy += 8) {
    const int r = j0 + yy, c = i0 + threadIdx.x;  // element (r, c) of the upper tile (tj, ti) = element (c, r) below
    if (r < M && c < M && c > r) Ab[(long)r * M + c] = t[threadIdx.x][yy];
  }
}
}  // namespace

extern "C" int gpsa_omega_prepare(int M, int B, const float* Osq, float* Omega, float* Ltril, double* L64,
                                  double* half_logdet, int* info, cudaStream_t st) {
  if (M <= 0 || B <= 0) return GPSA_OK;
  const long MM = (long)M * M;
  if (!L64) {
    // fp32 factorisation (the gene-batched Omega_F): the product Osq Osq^T is still ACCUMULATED in fp64 and rounded
    // once -- forming Omega in fp32 is what dominates the reference's own error here (cond(Omega) ~ 1e6 at its
    // initialisation: 3-8x further from float64 than this path, tools/ measurements in DESIGN.md) -- then one fp32
    // Cholesky per matrix, two CTAs per SM, log-determinant summed in fp64.
    TRY((gemm_strided<double, float, float, float>(st, M, M, M, 1.0, Osq, M, 1, MM, Osq, 1, M, MM, 0.0, Omega, M, MM, B, 1,
                                                   (double)GPSA_OFF, 1)));
    {
      const int nt = gpsa_cdiv(M, 32);
      for (int b0 = 0; b0 < B; b0 += 65535) {  // grid.y limit
        const int nb_ = B - b0 < 65535 ? B - b0 : 65535;
        mirror_lower_kernel<<<dim3(nt * (nt + 1) / 2, nb_), dim3(32, 8), 0, st>>>(M, nt, Omega + (long)b0 * MM);
        GPSA_LAUNCH_CHECK();
      }
    }
    return gpsa_potrf_batched_f32_ld64(M, B, Omega, Ltril, half_logdet, info, st);
  }
  // Omega = Osq Osq^T + 1e-5 I with fp64 accumulation, kept in fp64 for the factorisation
  // (symmetric: only the tiles that touch the lower triangle are computed; the fp32 copy mirrors them)
  TRY((gemm_strided<double, float, float, double>(st, M, M, M, 1.0, Osq, M, 1, MM, Osq, 1, M, MM, 0.0, L64, M, MM, B, 1,
                                                  (double)GPSA_OFF, 1)));
  cvt_sym_kernel<<<grid_for(MM * B), 256, 0, st>>>(MM * B, M, L64, Omega);
  GPSA_LAUNCH_CHECK();
  TRY(gpsa_potrf_batched_f64(M, B, L64, half_logdet, info, st));
  cvt_kernel<double, float><<<grid_for(MM * B), 256, 0, st>>>(MM * B, L64, Ltril);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_omega_grad(int M, int B, const float* Osq, const double* L64, const float* Obar, const float* coef,
                               double* Linv64, double* Y64, float* Osq_bar, cudaStream_t st) {
  return gpsa_omega_grad_tc(M, B, Osq, L64, Obar, coef, Linv64, Y64, Osq_bar, nullptr, 0, st);
}

extern "C" int gpsa_omega_grad_tc(int M, int B, const float* Osq, const double* L64, const float* Obar, const float* coef,
                                  double* Linv64, double* Y64, float* Osq_bar, void* tc_ws, size_t tc_ws_bytes,
                                  cudaStream_t st) {
  if (M <= 0 || B <= 0) return GPSA_OK;
  const long MM = (long)M * M;
  // Osq_bar = (Obar + Obar^T) Osq = 2 Obar Osq  (Obar symmetric) ...
  if (tc_ws && M >= 32) {
    TRY(gpsa_gemm_tc(M, M, M, B, Obar, M, MM, 1, Osq, M, MM, 0, Osq_bar, M, MM, 2.f, 0, 1, tc_ws, tc_ws_bytes, st));
  } else {
    TRY((gemm_strided<float, float, float, float>(st, M, M, M, 2.0, Obar, M, 1, MM, Osq, M, 1, MM, 0.0, Osq_bar, M, MM,
                                                  B)));
  }
  if (coef) {
    // ... + 2 coef[b] Omega^-1 Osq, Omega^-1 = Linv^T Linv, in fp64
    TRY(gpsa_trtri_batched_f64(M, B, L64, Linv64, st));
    // (Linv is lower triangular: the K range of each tile is trimmed to its non-zero part in both products)
    TRY((gemm_strided<double, double, float, double>(st, M, M, M, 1.0, Linv64, M, 1, MM, Osq, M, 1, MM, 0.0, Y64, M, MM,
                                                     B, 1, 0.0, 0, nullptr, 0, 1)));
    TRY((gemm_strided<double, double, double, float>(st, M, M, M, 2.0, Linv64, 1, M, MM, Y64, M, 1, MM, 1.0, Osq_bar, M,
                                                     MM, B, 1, 0.0, 0, coef, 1, 2)));
  }
  return GPSA_OK;
}

// fp32 form of the same backward for factors that came from the fp32 branch of gpsa_omega_prepare:
// Linv32, Y32: [B,M,M] fp32 scratch.
extern "C" int gpsa_omega_grad_f32(int M, int B, const float* Osq, const float* Ltril, const float* Obar, const float* coef,
                                   float* Linv32, float* Y32, float* Osq_bar, void* tc_ws, size_t tc_ws_bytes,
                                   cudaStream_t st) {
  if (M <= 0 || B <= 0) return GPSA_OK;
  const long MM = (long)M * M;
  if (tc_ws && M >= 32) {
    TRY(gpsa_gemm_tc(M, M, M, B, Obar, M, MM, 1, Osq, M, MM, 0, Osq_bar, M, MM, 2.f, 0, 1, tc_ws, tc_ws_bytes, st));
  } else {
    TRY((gemm_strided<float, float, float, float>(st, M, M, M, 2.0, Obar, M, 1, MM, Osq, M, 1, MM, 0.0, Osq_bar, M, MM,
                                                  B)));
  }
  if (coef) {
    // ... + 2 coef[b] Omega^-1 Osq, Omega^-1 = Linv^T Linv (triangular K ranges trimmed in both products)
    TRY(gpsa_trtri_batched_f32(M, B, Ltril, Linv32, st));
    TRY((gemm_strided<float, float, float, float>(st, M, M, M, 1.0, Linv32, M, 1, MM, Osq, M, 1, MM, 0.0, Y32, M, MM, B, 1,
                                                  0.0, 0, nullptr, 0, 1)));
    TRY((gemm_strided<float, float, float, float>(st, M, M, M, 2.0, Linv32, 1, M, MM, Y32, M, 1, MM, 1.0, Osq_bar, M, MM, B,
                                                  1, 0.0, 0, coef, 1, 2)));
  }
  return GPSA_OK;
}

// ------------------------------------------------------------------------------------------------
#define DISPATCH_WARP(D, kind, CALL)                                                                        \
  do {                                                                                                      \
    if (kind == GPSA_KIND_RBF) {                                                                            \
      if (D == 1) CALL(1, GPSA_KIND_RBF); else if (D == 2) CALL(2, GPSA_KIND_RBF); else CALL(3, GPSA_KIND_RBF); \
    } else if (kind == GPSA_KIND_MATERN12) {                                                                \
      if (D == 1) CALL(1, GPSA_KIND_MATERN12); else if (D == 2) CALL(2, GPSA_KIND_MATERN12);                \
      else CALL(3, GPSA_KIND_MATERN12);                                                                     \
    } else if (kind == GPSA_KIND_MATERN32) {                                                                \
      if (D == 1) CALL(1, GPSA_KIND_MATERN32); else if (D == 2) CALL(2, GPSA_KIND_MATERN32);                \
      else CALL(3, GPSA_KIND_MATERN32);                                                                     \
    } else {                                                                                                \
      return GPSA_ERR_UNSUPPORTED;                                                                          \
    }                                                                                                       \
  } while (0)

extern "C" int gpsa_warp_view_fwd(const gpsa_warp_fwd_args* a, cudaStream_t st) {
  const int M = a->M, D = a->D;
  const long n = a->n, MM = (long)M * M;
  if (n <= 0) return GPSA_OK;
  if (D < 1 || D > 3) return GPSA_ERR_ARG;
  if (a->kind == GPSA_KIND_EXTERNAL) {
    if (!a->Kuu_ext || !a->Kuf_ext) return GPSA_ERR_ARG;
    TRY(gpsa_prior_prepare_ext(M, a->Kuu_ext, a->Lk, a->Kinv, a->Kinv64, a->hld_K, a->info, a->ws64, st));
    cvt_kernel<float, double><<<grid_for((long)M * n), 256, 0, st>>>((long)M * n, a->Kuf_ext, a->B);
    GPSA_LAUNCH_CHECK();
  } else {
    TRY(gpsa_prior_prepare(a->kind, D, M, a->Z, a->log_ls, a->log_var, a->Lk, a->Kinv, a->Kinv64, a->hld_K, a->info,
                           a->ws64, st));
    dim3 grid(gpsa_cdiv(n, 256), M < 32 ? M : 32);
#define KF(DD, KK) kuf64_kernel<DD, KK><<<grid, 256, 0, st>>>(M, n, a->Z, a->X, a->log_ls, a->log_var, a->B)
    DISPATCH_WARP(D, a->kind, KF);
#undef KF
    GPSA_LAUNCH_CHECK();
  }
  // A = K^-1 K_uf  (replaces torch.cholesky_solve, vgpsa.py:177)
  TRY((gemm_nn<double, double, double, double>(st, M, (int)n, M, 1.0, a->Kinv64, M, a->B, n, 0.0, a->A, n)));
  // T_j = Omega_{v*D+j} A  -- the marginal variance uses slice v*D+j (vgpsa.py:336-339)
  TRY((gemm_strided<double, float, double, double>(st, M, (int)n, M, 1.0, a->Omega_G + (long)a->v * D * MM, M, 1, MM,
                                                   a->A, n, 1, 0, 0.0, a->T, n, (long)M * n, D)));
  const size_t smem = (size_t)M * D * sizeof(double);
  const int blocks = gpsa_cdiv(n, 256);
#define WP(DD)                                                                                                 \
  warp_predict_kernel<DD><<<blocks, 256, smem, st>>>(M, n, a->S, a->Z, a->dlt, a->A, a->B, a->T, a->X, a->eps, \
                                                     a->log_var, a->var, a->Gmean, a->Gs, a->gs_stride)
  if (D == 1) WP(1); else if (D == 2) WP(2); else WP(3);
#undef WP
  GPSA_LAUNCH_CHECK();
  kl_G_kernel<<<D, 256, M * sizeof(double), st>>>(M, D, a->V, a->v, a->Kinv64, a->Omega_G, a->hld_Omega, a->Z, a->dlt,
                                                  a->hld_K, a->Ke, a->kl_acc);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_warp_view_bwd(const gpsa_warp_bwd_args* a, cudaStream_t st) {
  const int M = a->M, D = a->D, V = a->V, v = a->v;
  const long n = a->n, MM = (long)M * M;
  if (n <= 0) return GPSA_OK;
  if (D < 1 || D > 3) return GPSA_ERR_ARG;
  double* Kbar = a->ws64;
  double* P = a->ws64 + MM;
  double* T1 = a->ws64 + 2 * MM;
  const int blocks = gpsa_cdiv(n, 256);
#define WB(DD)                                                                                          \
  warp_bwd_prep_kernel<DD><<<blocks, 256, 0, st>>>(n, a->S, a->Gs_bar, a->gs_stride, a->Gm_bar, a->eps, \
                                                   a->log_var, a->mubar, a->varbar, a->q1bar, a->acc_hyp)
  if (D == 1) WB(1); else if (D == 2) WB(2); else WB(3);
#undef WB
  GPSA_LAUNCH_CHECK();
  const int eb = grid_for((long)M * n);
#define WA(DD)                                                                                                  \
  warp_abar_kernel<DD><<<eb, 256, 0, st>>>(M, n, a->Z, a->dlt, a->A, a->B, a->T, a->mubar, a->varbar, a->q1bar, \
                                           a->Abar, a->AS)
  if (D == 1) WA(1); else if (D == 2) WA(2); else WA(3);
#undef WA
  GPSA_LAUNCH_CHECK();
  dim3 rgrid;
  long chunk;
  row_chunks(M, n, rgrid, chunk);
#define WD(DD) warp_dmz_kernel<DD><<<rgrid, 256, 0, st>>>(M, n, chunk, a->A, a->mubar, a->acc_dlt, a->acc_Z)
  if (D == 1) WD(1); else if (D == 2) WD(2); else WD(3);
#undef WD
  GPSA_LAUNCH_CHECK();
  // Omega-bar_{v*D+j} += (A o varbar_j) A^T
  {
    const int split = pick_split(M, M, n, D, 64);
    TRY((gemm_strided<double, double, double, float>(st, M, M, n, 1.0, a->AS, n, 1, (long)M * n, a->A, 1, n, 0, 1.0,
                                                     a->Obar_G + (long)v * D * MM, M, MM, D, split)));
  }
  // C = K^-1 Abar ; Kbar = -K^-1 (Abar A^T) ; Bbar = C + q1bar o A
  TRY((gemm_nn<double, double, double, double>(st, M, (int)n, M, 1.0, a->Kinv64, M, a->Abar, n, 0.0, a->C, n)));
  TRY(kbar_from_solve<double>(st, M, n, a->Kinv64, a->Abar, a->A, P, Kbar));
  TRY(colscale<double>(st, M, n, a->q1bar, a->A, a->C, 1));
  if (a->kl_bar) {
    // KL_v = sum_j [hldK - hldOm_j' + 0.5 (tr(K^-1 Om_j') + e_j^T K^-1 e_j - M)],  j' = j*V+v
    sum_batch_kernel<<<gpsa_cdiv(MM, 256), 256, 0, st>>>(MM, D, (long)V * MM, a->Omega_G + (long)v * MM, P);
    GPSA_LAUNCH_CHECK();
    TRY((gemm_nn<double, double, double, double>(st, M, M, M, 1.0, a->Kinv64, M, P, M, 0.0, T1, M)));
    TRY((gemm_nn<double, double, double, double>(st, M, M, M, -0.5, T1, M, a->Kinv64, M, 1.0, Kbar, M, a->kl_bar)));
    TRY((gemm_strided<double, double, double, double>(st, M, M, D, -0.5, a->Ke, 1, M, 0, a->Ke, M, 1, 0, 1.0, Kbar, M, 0,
                                                      1, 1, 0.0, 0, a->kl_bar, 0)));
    TRY((axpy_dev<double, double>(st, MM, 0.5 * D, a->kl_bar, a->Kinv64, Kbar)));
    for (int j = 0; j < D; ++j)
      TRY((axpy_dev<double, float>(st, MM, 0.5, a->kl_bar, a->Kinv64, a->Obar_G + (long)(j * V + v) * MM)));
    kl_e_bwd_kernel<<<gpsa_cdiv(M * D, 256), 256, 0, st>>>(M, D, a->Ke, a->kl_bar, a->acc_Z, a->acc_dlt);
    GPSA_LAUNCH_CHECK();
  }
  if (a->kind == GPSA_KIND_EXTERNAL) {
    // the caller differentiates its own covariance function: hand back dLoss/dK_uf and dLoss/dK_uu
    if (!a->Kuu_bar || !a->Kuf_bar) return GPSA_ERR_ARG;
    cvt_kernel<double, float><<<grid_for((long)M * n), 256, 0, st>>>((long)M * n, a->C, a->Kuf_bar);
    GPSA_LAUNCH_CHECK();
    cvt_kernel<double, float><<<grid_for(MM), 256, 0, st>>>(MM, Kbar, a->Kuu_bar);
    GPSA_LAUNCH_CHECK();
    return GPSA_OK;
  }
#define KB(DD, KK)                                                                                            \
  kuf64_bwd_kernel<DD, KK><<<rgrid, 256, 0, st>>>(M, n, chunk, a->Z, a->X, a->log_ls, a->log_var, a->C, a->acc_Z, \
                                                  a->acc_hyp)
  DISPATCH_WARP(D, a->kind, KB);
#undef KB
  GPSA_LAUNCH_CHECK();
  return prior_bwd(a->kind, D, M, a->Z, a->log_ls, a->log_var, Kbar, a->acc_Z, a->acc_hyp, st);
}

// ------------------------------------------------------------------------------------------------
extern "C" int gpsa_data_layer_fwd(const gpsa_data_fwd_args* a, cudaStream_t st) {
  const int M = a->M, D = a->D, L = a->L;
  const long R = a->R;
  if (R <= 0 || L <= 0) return GPSA_OK;
  if (D < 1 || D > 3) return GPSA_ERR_ARG;
  if (a->engine < 0 || a->engine > 1) return GPSA_ERR_UNSUPPORTED;
  if (a->engine == 1 && (!gpsa_tc_supported(M) || !a->tc_ws)) return GPSA_ERR_UNSUPPORTED;
  if (a->kind == GPSA_KIND_EXTERNAL) {
    // user-supplied covariance function: K_uu comes in, and B already holds K_uf [M,R]
    if (!a->Kuu_ext && !a->prior_ready) return GPSA_ERR_ARG;
    if (!a->prior_ready)
      TRY(gpsa_prior_prepare_ext(M, a->Kuu_ext, a->Lk, a->Kinv, a->Kinv64, a->hld_K, a->info, a->ws64, st));
  } else {
    if (!a->prior_ready)
      TRY(gpsa_prior_prepare(a->kind, D, M, a->Gt, a->log_ls, a->log_var, a->Lk, a->Kinv, a->Kinv64, a->hld_K, a->info,
                             a->ws64, st));
    TRY(gpsa_kernel_matrix_fwd(a->kind, D, M, R, a->Gt, a->G, a->log_ls, a->log_var, a->B, st));
  }
  TRY((gemm_nn<double, double, float, float>(st, M, (int)R, M, 1.0, a->Kinv64, M, a->B, R, 0.0, a->A, R)));
  kq_kernel<<<gpsa_cdiv(R, 256), 256, 0, st>>>(M, R, a->A, a->B, a->log_var, a->kq);
  GPSA_LAUNCH_CHECK();
  // predictive mean  F[r,p] = sum_m A[m,r] delta[m,p]   (vgpsa.py:182-184 with mu_x = mu_z = 0)
  if (a->engine >= 1) {
    TRY(gpsa_gemm_tc(R, L, M, 1, a->A, R, 0, 0, a->dlt, L, 0, 0, a->mean, L, 0, 1.f, 0, 1, a->tc_ws, a->tc_ws_bytes, st));
  } else {
    TRY((gemm_strided<float, float, float, float>(st, (int)R, L, M, 1.0, a->A, 1, R, 0, a->dlt, L, 1, 0, 0.0, a->mean, L, 0,
                                                  1)));
  }
  if (a->engine == 0) {
    TRY(gpsa_feat_pack(M, L, a->Omega, a->W, st));
    gpsa_prof_begin(0, st);
    TRY(gpsa_quadform_fwd_f32(M, R, L, a->A, a->W, a->q2, st));
    gpsa_prof_end(0, st);
  } else {
    TRY(gpsa_quadform_fwd_feat_tc(M, R, L, a->A, a->Omega, a->q2, a->tc_ws, a->tc_ws_bytes, st));  // times its own kernel
  }
  // KD = K^-1 delta (fp64)
  TRY((gemm_nn<double, double, float, double>(st, M, L, M, 1.0, a->Kinv64, M, a->dlt, L, 0.0, a->KD, L)));
  if (a->kl_acc) {
    kl_F_kernel<<<L, 256, 0, st>>>(M, L, a->Kinv64, a->Omega, a->hld_Omega, a->dlt, a->KD, a->hld_K, a->kl_acc);
    GPSA_LAUNCH_CHECK();
  }
  return GPSA_OK;
}

extern "C" int gpsa_data_layer_bwd(const gpsa_data_bwd_args* a, cudaStream_t st) {
  const int M = a->M, D = a->D, L = a->L;
  const long R = a->R, MM = (long)M * M;
  if (R <= 0 || L <= 0) return GPSA_OK;
  if (D < 1 || D > 3) return GPSA_ERR_ARG;
  if (a->engine < 0 || a->engine > 1 || (a->engine == 1 && !a->tc_ws)) return GPSA_ERR_UNSUPPORTED;
  double* Kbar = a->ws64;
  double* P = a->ws64 + MM;
  double* T1 = a->ws64 + 2 * MM;
  // var = kq + q2 + 2 off with kq = sigma2 - q1:  q1bar = -kq_bar, d/dlog_var += sigma2 sum kq_bar
  kq_bwd_kernel<<<grid_for(R, 256, 148 * 4), 256, 0, st>>>(R, a->kq_bar, a->log_var, a->q1bar, a->acc_hyp);
  GPSA_LAUNCH_CHECK();
  // delta-bar = A Fbar (+ kl_bar K^-1 delta)
  if (a->engine >= 1 && R <= 2000000000L) {
    TRY(gpsa_gemm_tc(M, L, (int)R, 1, a->A, R, 0, 1, a->mean_bar, L, 0, 0, a->dlt_bar, L, 0, 1.f, 0, 0, a->tc_ws,
                     a->tc_ws_bytes, st));
  } else {
    const int split = pick_split(M, L, R, 1, tile_of<float>(M, L));
    if (cudaMemsetAsync(a->dlt_bar, 0, sizeof(float) * (size_t)M * L, st) != cudaSuccess) return GPSA_ERR_CUDA;
    TRY((gemm_strided<float, float, float, float>(st, M, L, R, 1.0, a->A, R, 1, 0, a->mean_bar, L, 1, 0, 1.0, a->dlt_bar,
                                                  L, 0, 1, split)));
  }
  if (a->kl_bar) TRY((axpy_dev<double, float>(st, (long)M * L, 1.0, a->kl_bar, a->KD, a->dlt_bar)));
  // Abar = q1bar o B + delta Fbar^T + 2 (sum_p Gm Omega_p) a
  TRY(colscale<float>(st, M, R, a->q1bar, a->B, a->Abar, 0));
  if (a->engine >= 1) {
    TRY(gpsa_gemm_tc(R, M, L, 1, a->mean_bar, L, 0, 1, a->dlt, L, 0, 1, a->Abar, R, 0, 1.f, 1, 1, a->tc_ws, a->tc_ws_bytes, st));
  } else {
    TRY((gemm_strided<float, float, float, float>(st, M, (int)R, L, 1.0, a->dlt, L, 1, 0, a->mean_bar, 1, L, 0, 1.0, a->Abar,
                                                  R, 0, 1)));
  }
  // (the tcgen05 entry points time their own kernel launch -- not the operand packs -- in profiler slots 1 and 2)
  if (a->engine == 0) {
    gpsa_prof_begin(1, st);
    TRY(gpsa_quadform_bwd_alpha_f32(M, R, L, a->A, a->q2_bar, a->W, a->Abar, st));
    gpsa_prof_end(1, st);
  } else {
    TRY(gpsa_quadform_bwd_alpha_tc(M, R, L, a->A, a->q2_bar, a->Omega, a->Abar, a->tc_ws, a->tc_ws_bytes, st));
  }
  // Omega-bar = sum_r Gm a a^T (+ 0.5 kl_bar K^-1)
  if (a->engine == 0) {
    gpsa_prof_begin(2, st);
    TRY(gpsa_quadform_bwd_omega_f32(M, R, L, a->A, a->q2_bar, a->H, st));
    gpsa_prof_end(2, st);
  } else {
    TRY(gpsa_quadform_bwd_omega_tc(M, R, L, a->A, a->q2_bar, a->H, a->tc_ws, a->tc_ws_bytes, st));
  }
  TRY(gpsa_feat_unpack(M, L, a->H, a->kl_bar ? a->Kinv : nullptr, 0.5f, a->kl_bar, a->Obar, st));
  // C = K^-1 Abar ; Kbar = -K^-1 (Abar A^T) ; Bbar = C + q1bar o A
  TRY((gemm_nn<double, double, float, float>(st, M, (int)R, M, 1.0, a->Kinv64, M, a->Abar, R, 0.0, a->C, R)));
  TRY(kbar_from_solve<float>(st, M, R, a->Kinv64, a->Abar, a->A, P, Kbar));
  TRY(colscale<float>(st, M, R, a->q1bar, a->A, a->C, 1));
  if (a->kl_bar) {
    // KL_F = sum_p [hldK - hldOm_p + 0.5 (tr(K^-1 Om_p) + d_p^T K^-1 d_p - M)]
    sum_batch_kernel<<<gpsa_cdiv(MM, 256), 256, 0, st>>>(MM, L, MM, a->Omega, P);
    GPSA_LAUNCH_CHECK();
    TRY((gemm_nn<double, double, double, double>(st, M, M, M, 1.0, a->Kinv64, M, P, M, 0.0, T1, M)));
    TRY((gemm_nn<double, double, double, double>(st, M, M, M, -0.5, T1, M, a->Kinv64, M, 1.0, Kbar, M, a->kl_bar)));
    TRY((gemm_nt<double, double, double, double>(st, M, M, L, -0.5, a->KD, L, a->KD, L, 1.0, Kbar, M, a->kl_bar)));
    TRY((axpy_dev<double, double>(st, MM, 0.5 * L, a->kl_bar, a->Kinv64, Kbar)));
  }
  if (a->kind == GPSA_KIND_EXTERNAL) {
    // dLoss/dK_uf is C [M,R] as it stands; dLoss/dK_uu goes out in fp32
    if (!a->Kuu_bar) return GPSA_ERR_ARG;
    cvt_kernel<double, float><<<grid_for(MM), 256, 0, st>>>(MM, Kbar, a->Kuu_bar);
    GPSA_LAUNCH_CHECK();
    return GPSA_OK;
  }
  TRY(gpsa_kernel_matrix_bwd(a->kind, D, M, R, a->Gt, a->G, a->log_ls, a->log_var, a->C, a->acc_Gt, a->G_bar, nullptr,
                             a->acc_hyp, st));
  return prior_bwd(a->kind, D, M, a->Gt, a->log_ls, a->log_var, Kbar, a->acc_Gt, a->acc_hyp, st);
}

// ------------------------------------------------------------------------------------------------
// linear model of coregionalisation, materialised form: F_obs [R,P] = F_lat [R,L] W [L,P]  (vgpsa.py:428-432)
extern "C" int gpsa_lmc_fwd(long R, int L, int P, const float* F_lat, const float* W, float* F_obs, cudaStream_t st) {
  if (R <= 0 || L <= 0 || P <= 0) return GPSA_OK;
  if (R > 2000000000L) return GPSA_ERR_UNSUPPORTED;
  return gemm_strided<float, float, float, float>(st, (int)R, P, L, 1.0, F_lat, L, 1, 0, W, P, 1, 0, 0.0, F_obs, P, 0, 1);
}
// F_lat_bar = F_obs_bar W^T,  W_bar = F_lat^T F_obs_bar (split over the R rows)
extern "C" int gpsa_lmc_bwd(long R, int L, int P, const float* F_lat, const float* W, const float* F_obs_bar,
                            float* F_lat_bar, float* W_bar, cudaStream_t st) {
  if (R <= 0 || L <= 0 || P <= 0) return GPSA_OK;
  if (R > 2000000000L) return GPSA_ERR_UNSUPPORTED;
  TRY((gemm_strided<float, float, float, float>(st, (int)R, L, P, 1.0, F_obs_bar, P, 1, 0, W, 1, P, 0, 0.0, F_lat_bar, L, 0, 1)));
  const int split = pick_split(L, P, R, 1, tile_of<float>(L, P));
  if (cudaMemsetAsync(W_bar, 0, sizeof(float) * (size_t)L * P, st) != cudaSuccess) return GPSA_ERR_CUDA;
  return gemm_strided<float, float, float, float>(st, L, P, R, 1.0, F_lat, 1, L, 0, F_obs_bar, P, 1, 0, 1.0, W_bar, P, 0, 1, split);
}

// ------------------------------------------------------------------------------------------------
// sampling stage between the data layer and the likelihood (materialised form; the fused form is sample.cu)
extern "C" int gpsa_sample_fwd(long R, int L, const float* kq, const float* eps, float* F, float* var, cudaStream_t st) {
  if (R <= 0 || L <= 0) return GPSA_OK;
  const bool al16 = ((reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(var)) & 15) == 0;
  if (L >= 128 && (L & 3) == 0 && al16 && !g_no_vec4)
    sample_fwd_rows4_kernel<<<(int)(R < 148 * 16 ? R : 148 * 16), 256, 0, st>>>(
        R, L / 4, kq, reinterpret_cast<const float4*>(eps), reinterpret_cast<float4*>(F), reinterpret_cast<float4*>(var));
  else if (L >= 128) sample_fwd_rows_kernel<<<(int)(R < 148 * 16 ? R : 148 * 16), 256, 0, st>>>(R, L, kq, eps, F, var);
  else sample_fwd_kernel<<<grid_for(R * L), 256, 0, st>>>(R * L, L, kq, eps, F, var);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_sample_bwd(long R, int L, const float* F_bar, const float* eps, const float* var, float* q2_bar,
                               float* kq_bar, cudaStream_t st) {
  if (R <= 0 || L <= 0) return GPSA_OK;
  long b = (R + 7) / 8;
  if (b > 148 * 8) b = 148 * 8;
  const bool al16 = ((reinterpret_cast<uintptr_t>(F_bar) | reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(var) |
                      reinterpret_cast<uintptr_t>(q2_bar)) & 15) == 0;
  if (L >= 128 && (L & 3) == 0 && al16 && !g_no_vec4)
    sample_bwd4_kernel<<<(int)b, 256, 0, st>>>(R, L / 4, reinterpret_cast<const float4*>(F_bar), reinterpret_cast<const float4*>(eps),
                                               reinterpret_cast<const float4*>(var), reinterpret_cast<float4*>(q2_bar), kq_bar);
  else
    sample_bwd_kernel<<<(int)b, 256, 0, st>>>(R, L, F_bar, eps, var, q2_bar, kq_bar);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int gpsa_gaussian_ll_fwd(long N, int P, int S, const float* F, const float* Y, const float* log_noise,
                                    double* ll_acc, cudaStream_t st) {
  if (N <= 0 || P <= 0 || S <= 0) return GPSA_OK;
  const long NP = N * P;
  if ((NP & 3) == 0 && ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0 && NP >= 4096 && !g_no_vec4)
    ll_fwd4_kernel<<<grid_for(NP / 4, 256, 148 * 8), 256, 0, st>>>(NP / 4, S, reinterpret_cast<const float4*>(F),
                                                                    reinterpret_cast<const float4*>(Y), log_noise, ll_acc);
  else
    ll_fwd_kernel<<<grid_for(N * P, 256, 148 * 8), 256, 0, st>>>(N, P, S, F, Y, log_noise, ll_acc);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_gaussian_ll_bwd(long N, int P, int S, const float* F, const float* Y, const float* log_noise,
                                    const float* ll_bar, float* F_bar, double* acc_noise, cudaStream_t st) {
  if (N <= 0 || P <= 0 || S <= 0) return GPSA_OK;
  const long NP = N * P;
  if ((NP & 3) == 0 && NP >= 4096 && !g_no_vec4 &&
      ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(F_bar)) & 15) == 0)
    ll_bwd4_kernel<<<grid_for(NP / 4, 256, 148 * 8), 256, 0, st>>>(NP / 4, S, reinterpret_cast<const float4*>(F),
                                                                    reinterpret_cast<const float4*>(Y), log_noise, ll_bar,
                                                                    reinterpret_cast<float4*>(F_bar), acc_noise);
  else
    ll_bwd_kernel<<<grid_for(N * P, 256, 148 * 8), 256, 0, st>>>(N, P, S, F, Y, log_noise, ll_bar, F_bar, acc_noise);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}
