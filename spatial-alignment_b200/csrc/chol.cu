// Batched Cholesky factorisation and triangular inverse of the M x M inducing matrices,
// one CTA per matrix.
//
// Replaces torch.cholesky at reference gpsa/models/vgpsa.py:257 (Omega_G, batch V*D), :320 (K_uu per
// view), :394 (data K_uu), :412 (Omega_F, batch L) and the triangular solves inside
// torch.cholesky_solve (:177) and MultivariateNormal KL (:506-530).
//
// Blocked right-looking factorisation with 32-wide panels: the 32x32 diagonal block is
// factorised by one warp entirely in registers with warp shuffles (lane i owns row i), the panel
// below it is solved one row per thread against the block held in shared memory, and the
// trailing update streams the panel from shared memory.  fp32 for the gene-batched variational
// covariances, fp64 for the handful of ill-conditioned prior matrices K_uu.
#include "common.cuh"
#include "gpsa_b200.h"

namespace {

constexpr int NB = 32;
constexpr int LDP = NB + 1;

template <typename T>
__device__ __forceinline__ T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Factor the (identity-padded) 32x32 block in Dg (ld = LDP) in place; lower triangle out, zeros above.
// Executed by warp 0.  Returns true if a non-positive pivot was met.
template <typename T>
__device__ __forceinline__ bool warp_potrf32(T* Dg, int kn) {
  const int lane = threadIdx.x & 31;
  T r[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) r[j] = (lane < kn && j < kn) ? Dg[lane * LDP + j] : ((lane == j) ? T(1) : T(0));
  bool bad = false;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    T d = shfl(r[j], j);
    if (!(d > T(0))) { bad = true; d = T(1); }
    d = sqrt(d);
    const T inv = T(1) / d;
    if (lane == j) r[j] = d;
    else if (lane > j) r[j] *= inv;
    const T lij = r[j];
#pragma unroll
    for (int c = j + 1; c < NB; ++c) {
      const T lcj = shfl(r[j], c);
      if (lane >= c) r[c] -= lij * lcj;
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) Dg[lane * LDP + j] = (j <= lane) ? r[j] : T(0);
  return bad;
}

template <typename T>
__global__ void __launch_bounds__(256) potrf_kernel(int M, T* A, long stride, T* half_logdet, int* info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Dg = reinterpret_cast<T*>(smem_raw);  // [NB][LDP]
  T* P = Dg + NB * LDP;                    // [M][LDP] panel rows below the diagonal block
  T* Ab = A + (long)blockIdx.x * stride;
  const int tid = threadIdx.x;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int k0 = 0; k0 < M; k0 += NB) {
    const int kn = (M - k0 < NB) ? M - k0 : NB;
    for (int idx = tid; idx < kn * kn; idx += blockDim.x) {
      const int i = idx / kn, j = idx % kn;
      Dg[i * LDP + j] = Ab[(long)(k0 + i) * M + k0 + j];
    }
    __syncthreads();
    if (tid < 32) {
      const bool bad = warp_potrf32<T>(Dg, kn);
      if (bad && tid == 0) s_bad = 1;
    }
    __syncthreads();
    for (int idx = tid; idx < kn * kn; idx += blockDim.x) {
      const int i = idx / kn, j = idx % kn;
      Ab[(long)(k0 + i) * M + k0 + j] = Dg[i * LDP + j];
    }
    const int rem = M - k0 - kn;
    // panel: row i of A21 <- a_i L11^-T
    for (int i = tid; i < rem; i += blockDim.x) {
      T* row = Ab + (long)(k0 + kn + i) * M + k0;
      T x[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] = (c < kn) ? row[c] : T(0);
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (c < kn) {
          T s = x[c];
#pragma unroll
          for (int t = 0; t < c; ++t) s -= x[t] * Dg[c * LDP + t];
          x[c] = s / Dg[c * LDP + c];
        }
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (c < kn) row[c] = x[c];
        P[i * LDP + c] = x[c];
      }
    }
    __syncthreads();
    // trailing update on the lower triangle: A22 -= P P^T, 32x32 tiles, 2x2 per thread
    const int nt = (rem + NB - 1) / NB;
    const int ty = tid >> 4, tx = tid & 15;
    for (int ti = 0; ti < nt; ++ti) {
      for (int tj = 0; tj <= ti; ++tj) {
        const int i0 = ti * NB + ty * 2, j0 = tj * NB + tx * 2;
        T a00 = 0, a01 = 0, a10 = 0, a11 = 0;
        const bool vi0 = i0 < rem, vi1 = i0 + 1 < rem, vj0 = j0 < rem, vj1 = j0 + 1 < rem;
        const T* pi0 = P + (vi0 ? i0 : 0) * LDP;
        const T* pi1 = P + (vi1 ? i0 + 1 : 0) * LDP;
        const T* pj0 = P + (vj0 ? j0 : 0) * LDP;
        const T* pj1 = P + (vj1 ? j0 + 1 : 0) * LDP;
#pragma unroll 8
        for (int c = 0; c < NB; ++c) {
          const T u0 = pi0[c], u1 = pi1[c], w0 = pj0[c], w1 = pj1[c];
          a00 += u0 * w0; a01 += u0 * w1; a10 += u1 * w0; a11 += u1 * w1;
        }
        T* base = Ab + (long)(k0 + kn) * M + (k0 + kn);
        if (vi0 && vj0 && j0 <= i0) base[(long)i0 * M + j0] -= a00;
        if (vi0 && vj1 && j0 + 1 <= i0) base[(long)i0 * M + j0 + 1] -= a01;
        if (vi1 && vj0 && j0 <= i0 + 1) base[(long)(i0 + 1) * M + j0] -= a10;
        if (vi1 && vj1 && j0 + 1 <= i0 + 1) base[(long)(i0 + 1) * M + j0 + 1] -= a11;
      }
    }
    __syncthreads();
  }
  // zero the strict upper triangle (torch.cholesky returns a clean lower factor), sum of log diag
  for (long idx = tid; idx < (long)M * M; idx += blockDim.x) {
    const int i = idx / M, j = idx % M;
    if (j > i) Ab[idx] = T(0);
  }
  T ld = 0;
  for (int i = tid; i < M; i += blockDim.x) ld += log(Ab[(long)i * M + i]);
  __shared__ T red[32];
  ld = block_sum<T>(ld, red);
  if (tid == 0) {
    if (half_logdet) half_logdet[blockIdx.x] = ld;
    if (info) info[blockIdx.x] = s_bad;
  }
}

// X = L^-1 (lower triangular), block row by block row.  X must not alias L.
template <typename T>
__global__ void __launch_bounds__(256) trtri_kernel(int M, const T* L, T* X, long stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Dg = reinterpret_cast<T*>(smem_raw);  // [NB][LDP]  L_II, then X_II
  T* LT = Dg + NB * LDP;                   // [M][NB]    LT[t][i] = L[I0+i][t]
  const T* Lb = L + (long)blockIdx.x * stride;
  T* Xb = X + (long)blockIdx.x * stride;
  const int tid = threadIdx.x;
  for (int I0 = 0; I0 < M; I0 += NB) {
    const int in = (M - I0 < NB) ? M - I0 : NB;
    for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
      const int i = idx / NB, j = idx % NB;
      Dg[i * LDP + j] = (i < in && j < in && j <= i) ? Lb[(long)(I0 + i) * M + I0 + j] : ((i == j) ? T(1) : T(0));
    }
    for (int idx = tid; idx < I0 * NB; idx += blockDim.x) {
      const int t = idx % I0, i = idx / I0;  // consecutive threads walk t (contiguous in L's row)
      LT[t * NB + i] = (i < in) ? Lb[(long)(I0 + i) * M + t] : T(0);
    }
    __syncthreads();
    if (tid < 32) {
      // lane c solves L_II x = e_c by forward substitution
      const int c = tid;
      T x[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        T s = (i == c) ? T(1) : T(0);
#pragma unroll
        for (int t = 0; t < i; ++t) s -= Dg[i * LDP + t] * x[t];
        x[i] = s / Dg[i * LDP + i];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NB; ++i) Dg[i * LDP + c] = x[i];  // column c of X_II (zeros above the diagonal)
    }
    __syncthreads();
    for (int idx = tid; idx < in * in; idx += blockDim.x) {
      const int i = idx / in, j = idx % in;
      Xb[(long)(I0 + i) * M + I0 + j] = Dg[i * LDP + j];
    }
    // off-diagonal part of this block row: X[I, 0:I0] = -X_II (L[I,0:I0] X[0:I0,0:I0])
    for (int c = tid; c < I0; c += blockDim.x) {
      T acc[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) acc[i] = T(0);
      const int t0 = (c / NB) * NB;  // X[t][c] = 0 for t < c; start at the block boundary (warp-uniform)
      for (int t = t0; t < I0; ++t) {
        const T xv = Xb[(long)t * M + c];
        const T* lt = LT + t * NB;
#pragma unroll
        for (int i = 0; i < NB; ++i) acc[i] += lt[i] * xv;
      }
#pragma unroll 4
      for (int i = 0; i < NB; ++i) {
        if (i < in) {
          T s = T(0);
#pragma unroll
          for (int k = 0; k < NB; ++k) s -= Dg[i * LDP + k] * acc[k];
          Xb[(long)(I0 + i) * M + c] = s;
        }
      }
    }
    // strict upper part of this block row is zero
    for (int idx = tid; idx < in * (M - I0); idx += blockDim.x) {
      const int i = idx / (M - I0), j = I0 + idx % (M - I0);
      if (j > I0 + i) Xb[(long)(I0 + i) * M + j] = T(0);
    }
    __syncthreads();
  }
}

template <typename T>
int potrf_launch(int M, int batch, T* A, T* half_logdet, int* info, cudaStream_t st) {
  const size_t smem = (size_t)(NB * LDP + (size_t)M * LDP) * sizeof(T);
  if (smem > 220 * 1024) return GPSA_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(potrf_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return GPSA_ERR_CUDA;
  potrf_kernel<T><<<batch, 256, smem, st>>>(M, A, (long)M * M, half_logdet, info);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

template <typename T>
int trtri_launch(int M, int batch, const T* L, T* X, cudaStream_t st) {
  const size_t smem = (size_t)(NB * LDP + (size_t)M * NB) * sizeof(T);
  if (smem > 220 * 1024) return GPSA_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(trtri_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return GPSA_ERR_CUDA;
  trtri_kernel<T><<<batch, 256, smem, st>>>(M, L, X, (long)M * M);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

}  // namespace

extern "C" int gpsa_potrf_batched_f32(int M, int batch, float* A, float* half_logdet, int* info, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return potrf_launch<float>(M, batch, A, half_logdet, info, st);
}
extern "C" int gpsa_potrf_batched_f64(int M, int batch, double* A, double* half_logdet, int* info, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return potrf_launch<double>(M, batch, A, half_logdet, info, st);
}
extern "C" int gpsa_trtri_batched_f32(int M, int batch, const float* L, float* X, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return trtri_launch<float>(M, batch, L, X, st);
}
extern "C" int gpsa_trtri_batched_f64(int M, int batch, const double* L, double* X, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return trtri_launch<double>(M, batch, L, X, st);
}
