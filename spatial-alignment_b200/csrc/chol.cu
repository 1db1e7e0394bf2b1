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

#include <stdlib.h>

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
    for (int c = 0; c < NB; ++c) {  // static bounds: r[] stays in registers
      if (c > j) {
        const T lcj = shfl(r[j], c);
        if (lane >= c) r[c] -= lij * lcj;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) Dg[lane * LDP + j] = (j <= lane) ? r[j] : T(0);
  return bad;
}

template <typename T, typename TL>
__global__ void __launch_bounds__(256) potrf_kernel(int M, T* A, long stride, TL* half_logdet, int* info) {
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
          for (int t = 0; t < NB; ++t)
            if (t < c) s -= x[t] * Dg[c * LDP + t];
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
  TL ld = 0;
  for (int i = tid; i < M; i += blockDim.x) ld += log((TL)Ab[(long)i * M + i]);
  __shared__ TL red[32];
  ld = block_sum<TL>(ld, red);
  if (tid == 0) {
    // a non-positive pivot poisons the log-determinant: the KL terms, hence the loss, become NaN on the device, so a
    // failed factorisation cannot pass silently even when nobody reads `info` (torch.cholesky raises on the host)
    if (half_logdet) half_logdet[blockIdx.x] = s_bad ? (TL)NAN : ld;
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

// ------------------------------------------------------------------------------------------------
// Packed in-shared-memory variants (one CTA of 512 threads per matrix): the whole lower triangle lives in shared
// memory, row-major packed (element (i,j) at i(i+1)/2 + j), so neither factorisation nor inversion touches global
// memory between the initial load and the final store.  fp64: M <= 208 (174 KB); fp32: M <= 300.
// The panel kernels above round-trip the trailing matrix through L2 once per 32 columns and spend 0.4 ms on a
// single 200 x 200 fp64 matrix; these take 0.156 ms (potrf) / 0.149 ms (trtri), bound by the serial 16-wide
// diagonal blocks and the barriers between panel phases.
// ------------------------------------------------------------------------------------------------
constexpr int PK_THREADS = 512;      // fp64: one CTA per SM (the packed triangle of M = 200 is 161 KB)
constexpr int PK_THREADS_F32 = 256;  // fp32: half the shared memory and half the threads -> two CTAs per SM, so one
                                     // matrix's serial diagonal-block phases overlap the other's trailing updates
__host__ __device__ inline long pk(int i, int j) { return (long)i * (i + 1) / 2 + j; }

constexpr int PB = 16;        // panel width of the packed kernels
constexpr int PLD = PB + 1;

// Factor the identity-padded PB x PB diagonal block held one row per lane (lanes >= kn: identity rows).
// On exit lane i holds row i of the factor in r[0..i].  Returns true on a non-positive pivot.
template <typename T>
__device__ __forceinline__ bool warp_potrf16(T (&r)[PB], int lane) {
  bool bad = false;
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    T d = shfl(r[j], j);
    if (!(d > T(0))) { bad = true; d = T(1); }
    const T inv = rsqrt(d);  // <= 1 ulp; the pivot itself is d * rsqrt(d): no division, no second special function
    d = d * inv;
    if (lane == j) r[j] = d;
    else if (lane > j) r[j] *= inv;
    const T lij = r[j];
#pragma unroll
    for (int c = 0; c < PB; ++c) {  // static bounds: r[] stays in registers
      if (c > j) {
        const T lcj = shfl(r[j], c);
        if (lane >= c) r[c] -= lij * lcj;
      }
    }
  }
  return bad;
}

// Right-looking blocked factorisation on the packed triangle: per 16-column panel, warp 0 factors the diagonal
// block in registers, one thread per row solves the panel against it, and the trailing triangle is updated in
// 8 x 4 register tiles (12 shared-memory loads per 32 FMAs).  The fp64 pipe of one SM (64 FMA/clk) bounds a
// 200 x 200 matrix at ~25 us; the barriers and the serial diagonal blocks add about as much again.
// Ain (lower triangle read) -> Aout (lower factor, zeros above); Aout may be Ain.  TL: type of the log-determinant.
template <typename T, typename TL, int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
potrf_packed_kernel(int M, const T* Ain, T* Aout, long stride, TL* half_logdet, int* info) {
  constexpr int PK_THREADS = THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Lp = reinterpret_cast<T*>(smem_raw);       // packed lower triangle
  T* Dg = Lp + (long)M * (M + 1) / 2;           // [PB][PLD] current diagonal block
  T* dinv = Dg + PB * PLD;                      // [PB]
  const T* Ai = Ain + (long)blockIdx.x * stride;
  T* Ab = Aout + (long)blockIdx.x * stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (long idx = tid; idx < (long)M * M; idx += PK_THREADS) {
    const int i = idx / M, j = idx % M;
    if (j <= i) Lp[pk(i, j)] = Ai[idx];
  }
  __syncthreads();
  for (int k0 = 0; k0 < M; k0 += PB) {
    const int kn = (M - k0 < PB) ? M - k0 : PB;
    const int base = k0 + kn, rem = M - base;
    if (warp == 0) {
      T r[PB];
#pragma unroll
      for (int c = 0; c < PB; ++c)
        r[c] = (lane < kn && c <= lane) ? Lp[pk(k0 + lane, k0 + c)] : ((lane == c) ? T(1) : T(0));
      const bool bad = warp_potrf16<T>(r, lane);
      if (__any_sync(0xffffffffu, bad) && lane == 0) s_bad = 1;
      if (lane < PB) {
#pragma unroll
        for (int c = 0; c < PB; ++c) {
          Dg[lane * PLD + c] = (c <= lane) ? r[c] : T(0);
          if (lane < kn && c <= lane) Lp[pk(k0 + lane, k0 + c)] = r[c];
          if (c == lane) dinv[lane] = T(1) / r[c];
        }
      }
    }
    __syncthreads();
    // panel: row i of A21 <- a_i L11^-T
    for (int row = base + tid; row < M; row += PK_THREADS) {
      T* ar = Lp + pk(row, k0);
      T x[PB];
#pragma unroll
      for (int c = 0; c < PB; ++c) x[c] = (c < kn) ? ar[c] : T(0);
#pragma unroll
      for (int c = 0; c < PB; ++c) {
        T s0 = x[c], s1 = T(0);  // two chains: halves the dependent-FMA latency of the substitution
#pragma unroll
        for (int t = 0; t < PB; ++t) {
          if (t < c) {
            if (t & 1) s1 -= x[t] * Dg[c * PLD + t];
            else s0 -= x[t] * Dg[c * PLD + t];
          }
        }
        x[c] = (s0 + s1) * dinv[c];
      }
#pragma unroll
      for (int c = 0; c < PB; ++c)
        if (c < kn) ar[c] = x[c];
    }
    __syncthreads();
    // trailing update of the lower triangle: A22 -= P P^T in 8 (rows) x 4 (cols) tiles; row tile ti needs the
    // column tiles 0 .. 2 ti + 1, i.e. tiles are enumerated as t = ti^2 + ti + tj
    if (rem > 0) {
      const int nrt = (rem + 7) / 8;
      const int ntiles = nrt * nrt + nrt;
      for (int t = tid; t < ntiles; t += PK_THREADS) {
        int ti = (int)((sqrtf(4.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (ti * ti + ti > t) --ti;
        while ((ti + 1) * (ti + 1) + (ti + 1) <= t) ++ti;
        const int tj = t - ti * ti - ti;
        const int i0 = base + 8 * ti, j0 = base + 4 * tj;
        if (j0 >= M) continue;
        const T* pa[8];
        const T* pb[4];
#pragma unroll
        for (int a = 0; a < 8; ++a) pa[a] = Lp + pk(min(i0 + a, M - 1), k0);
#pragma unroll
        for (int b = 0; b < 4; ++b) pb[b] = Lp + pk(min(j0 + b, M - 1), k0);
        T acc[8][4];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
        for (int c = 0; c < kn; ++c) {
          T va[8], vb[4];
#pragma unroll
          for (int a = 0; a < 8; ++a) va[a] = pa[a][c];
#pragma unroll
          for (int b = 0; b < 4; ++b) vb[b] = pb[b][c];
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += va[a] * vb[b];
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int i = i0 + a;
          if (i >= M) continue;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int j = j0 + b;
            if (j <= i) Lp[pk(i, j)] -= acc[a][b];
          }
        }
      }
    }
    __syncthreads();
  }
  for (long idx = tid; idx < (long)M * M; idx += PK_THREADS) {
    const int i = idx / M, j = idx % M;
    Ab[idx] = (j <= i) ? Lp[pk(i, j)] : T(0);
  }
  TL ld = 0;
  for (int i = tid; i < M; i += PK_THREADS) ld += log((TL)Lp[pk(i, i)]);
  __shared__ TL red[32];
  ld = block_sum<TL>(ld, red);
  if (tid == 0) {
    // a non-positive pivot poisons the log-determinant: the KL terms, hence the loss, become NaN on the device, so a
    // failed factorisation cannot pass silently even when nobody reads `info` (torch.cholesky raises on the host)
    if (half_logdet) half_logdet[blockIdx.x] = s_bad ? (TL)NAN : ld;
    if (info) info[blockIdx.x] = s_bad;
  }
}

// In-place inversion of the packed factor by block rows of 16:
//   X_II = L_II^-1,   X[I, 0:I0] = -X_II (L[I, 0:I0] X[0:I0, 0:I0]).
// The block row of L is staged in a side buffer (it is overwritten by X[I, :]); the product W = L[I,:] X runs in
// 4 x 4 register tiles with the contraction split over 4 adjacent lanes (the early block rows are short and would
// otherwise leave most of the CTA idle), while one warp inverts the diagonal block.
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS) trtri_packed_kernel(int M, const T* L, T* X, long stride) {
  constexpr int PK_THREADS = THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Xp = reinterpret_cast<T*>(smem_raw);   // packed: L on entry, X on exit
  T* stage = Xp + (long)M * (M + 1) / 2;    // [PB][M]  block row of L
  T* Wst = stage + (long)PB * M;            // [PB][M]  W = L[I, 0:I0] X[0:I0, 0:I0]
  T* Dg = Wst + (long)PB * M;               // [PB][PLD] L_II, then X_II
  const T* Lb = L + (long)blockIdx.x * stride;
  T* Xb = X + (long)blockIdx.x * stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (long idx = tid; idx < (long)M * M; idx += PK_THREADS) {
    const int i = idx / M, j = idx % M;
    if (j <= i) Xp[pk(i, j)] = Lb[idx];
  }
  __syncthreads();
  for (int I0 = 0; I0 < M; I0 += PB) {
    const int in = (M - I0 < PB) ? M - I0 : PB;
    for (int idx = tid; idx < PB * I0; idx += PK_THREADS) {
      const int i = idx / I0, t = idx % I0;
      stage[i * M + t] = (i < in) ? Xp[pk(I0 + i, t)] : T(0);
    }
    for (int idx = tid; idx < PB * PB; idx += PK_THREADS) {
      const int i = idx / PB, j = idx % PB;
      Dg[i * PLD + j] = (i < in && j <= i) ? Xp[pk(I0 + i, I0 + j)] : ((i == j) ? T(1) : T(0));
    }
    __syncthreads();
    if (warp == PK_THREADS / 32 - 1) {
      // X_II by forward substitution: lane c solves L_II x = e_c; written back as column c (zeros above the diagonal)
      T xcol[PB];
      const int c = lane & (PB - 1);
#pragma unroll
      for (int i = 0; i < PB; ++i) {
        T sacc = (i == c) ? T(1) : T(0);
#pragma unroll
        for (int t = 0; t < i; ++t) sacc -= Dg[i * PLD + t] * xcol[t];
        xcol[i] = sacc / Dg[i * PLD + i];
      }
      __syncwarp();
      if (lane < PB) {
#pragma unroll
        for (int i = 0; i < PB; ++i) Dg[i * PLD + c] = xcol[i];
      }
    }
    // work item = (row group of 4, column group of 4, contraction split ks of 4); the loop bound is rounded up to a
    // whole warp so that the shuffles below are executed by every lane
    const int ncg = (I0 + 3) / 4;
    const int nwork = 4 * ncg * 4;
    const int nwork32 = (nwork + 31) & ~31;
    for (int item = tid; item < nwork32; item += PK_THREADS) {
      const bool valid = item < nwork;
      const int ks = item & 3;
      const int cg = valid ? (item >> 2) % ncg : 0, rg = valid ? (item >> 2) / ncg : 0;
      const int c0 = cg * 4;
      T acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
      if (valid) {
        for (int t = c0 + ks; t < I0; t += 4) {
          T lv[4], xv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) lv[a] = stage[(rg * 4 + a) * M + t];
          const T* xr = Xp + pk(t, c0);
#pragma unroll
          for (int b = 0; b < 4; ++b) xv[b] = (c0 + b <= t) ? xr[b] : T(0);  // X[t, c] = 0 above the diagonal
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += lv[a] * xv[b];
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          T v = acc[a][b];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          acc[a][b] = v;
        }
      if (valid && ks == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (c0 + b < I0) Wst[(rg * 4 + a) * M + c0 + b] = acc[a][b];
      }
    }
    __syncthreads();
    // X[I, c] = -X_II W[:, c], and the diagonal block
    for (int idx = tid; idx < in * I0; idx += PK_THREADS) {
      const int i = idx / I0, c = idx % I0;
      T sacc = T(0);
#pragma unroll
      for (int k = 0; k < PB; ++k) sacc -= Dg[i * PLD + k] * Wst[k * M + c];
      Xp[pk(I0 + i, c)] = sacc;
    }
    for (int idx = tid; idx < in * in; idx += PK_THREADS) {
      const int i = idx / in, j = idx % in;
      if (j <= i) Xp[pk(I0 + i, I0 + j)] = Dg[i * PLD + j];
    }
    __syncthreads();
  }
  for (long idx = tid; idx < (long)M * M; idx += PK_THREADS) {
    const int i = idx / M, j = idx % M;
    Xb[idx] = (j <= i) ? Xp[pk(i, j)] : T(0);
  }
}

template <typename T>
size_t packed_smem(int M, int extra_rows) {
  return ((size_t)M * (M + 1) / 2 + (size_t)extra_rows * M + PB * PLD + PB) * sizeof(T);
}

template <typename T> struct PkThreads { static constexpr int v = PK_THREADS; };
template <> struct PkThreads<float> { static constexpr int v = PK_THREADS_F32; };

// Ain -> Aout (may alias), log-determinant in TL
template <typename T, typename TL>
int potrf_launch(int M, int batch, const T* Ain, T* Aout, TL* half_logdet, int* info, cudaStream_t st) {
#ifdef GPSA_DEBUG  // experiments only: force the panel kernels
  static const int no_packed = [] { const char* e = getenv("GPSA_NO_PACKED_CHOL"); return e ? atoi(e) : 0; }();
#else
  constexpr int no_packed = 0;
#endif
  constexpr int TH = PkThreads<T>::v;
  const size_t psm = packed_smem<T>(M, 0);
  const size_t cap = sizeof(T) == 4 ? 113 * 1024 : 226 * 1024;  // fp32: two CTAs per SM
  if (!no_packed && psm <= 226 * 1024) {
    static size_t attr[GPSA_MAX_DEVICES] = {};
    const int dev = gpsa_dev();
    if (psm > 48 * 1024 && psm > attr[dev]) {
      if (cudaFuncSetAttribute(potrf_packed_kernel<T, TL, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess)
        return GPSA_ERR_CUDA;
      attr[dev] = 226 * 1024;
    }
    (void)cap;
    potrf_packed_kernel<T, TL, TH><<<batch, TH, psm, st>>>(M, Ain, Aout, (long)M * M, half_logdet, info);
    GPSA_LAUNCH_CHECK();
    return GPSA_OK;
  }
  const size_t smem = (size_t)(NB * LDP + (size_t)M * LDP) * sizeof(T);
  if (smem > 220 * 1024) return GPSA_ERR_UNSUPPORTED;
  if (Ain != Aout &&
      cudaMemcpyAsync(Aout, Ain, sizeof(T) * (size_t)batch * M * M, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return GPSA_ERR_CUDA;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(potrf_kernel<T, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return GPSA_ERR_CUDA;
  potrf_kernel<T, TL><<<batch, 256, smem, st>>>(M, Aout, (long)M * M, half_logdet, info);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

template <typename T>
int trtri_launch(int M, int batch, const T* L, T* X, cudaStream_t st) {
#ifdef GPSA_DEBUG  // experiments only: force the panel kernels
  static const int no_packed = [] { const char* e = getenv("GPSA_NO_PACKED_CHOL"); return e ? atoi(e) : 0; }();
#else
  constexpr int no_packed = 0;
#endif
  constexpr int TH = PkThreads<T>::v;
  const size_t psm = packed_smem<T>(M, 2 * PB);
  if (!no_packed && psm <= 226 * 1024 && L != X) {
    static size_t attr[GPSA_MAX_DEVICES] = {};
    const int dev = gpsa_dev();
    if (psm > 48 * 1024 && psm > attr[dev]) {
      if (cudaFuncSetAttribute(trtri_packed_kernel<T, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess)
        return GPSA_ERR_CUDA;
      attr[dev] = 226 * 1024;
    }
    trtri_packed_kernel<T, TH><<<batch, TH, psm, st>>>(M, L, X, (long)M * M);
    GPSA_LAUNCH_CHECK();
    return GPSA_OK;
  }
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
  return potrf_launch<float, float>(M, batch, A, A, half_logdet, info, st);
}
extern "C" int gpsa_potrf_batched_f64(int M, int batch, double* A, double* half_logdet, int* info, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return potrf_launch<double, double>(M, batch, A, A, half_logdet, info, st);
}
// fp32 factorisation A (lower triangle read) -> L (may alias A) with the log-determinant accumulated in fp64: the form
// the gene-batched variational covariances use (gpsa_omega_prepare with L64 == NULL)
extern "C" int gpsa_potrf_batched_f32_ld64(int M, int batch, const float* A, float* L, double* half_logdet, int* info,
                                           cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return potrf_launch<float, double>(M, batch, A, L, half_logdet, info, st);
}
extern "C" int gpsa_trtri_batched_f32(int M, int batch, const float* L, float* X, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return trtri_launch<float>(M, batch, L, X, st);
}
extern "C" int gpsa_trtri_batched_f64(int M, int batch, const double* L, double* X, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return GPSA_OK;
  return trtri_launch<double>(M, batch, L, X, st);
}
