// Register-tiled SIMT GEMM engine (fp32 / fp64) shared by every non-tensor-core
// contraction of the path: A = K^-1 B, the predictive-mean product, Omega =
// Omega_sqt Omega_sqt^T, the K-bar / Omega-bar backward products, and -- with an
// on-the-fly operand generator -- the implicit-feature quadratic-form GEMMs.
//
//   C tile (BM x BN) per CTA, (BM/TM)*(BN/TN) threads, each a TM x TN micro-tile.
//   Operand tiles are staged k-major in shared memory: As[kk][i], Bs[kk][j], so
//   the inner loop reads TM + TN values with 128-bit loads for TM*TN FMAs.
//   Loaders and the epilogue are functors so that operands can be strided views,
//   generated products, or gathered rows.
#pragma once
#include "common.cuh"

#include <stdlib.h>

template <typename T, int BM_, int BN_, int BK_, int TM_, int TN_>
struct GemmCfg {
  using elem = T;
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static constexpr int PAD = 4;
  static constexpr int LDA = BM + PAD, LDB = BN + PAD;
  static constexpr int SMEM_ELEMS = BK * (LDA + LDB);
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro-tile is built from 4-wide chunks");
};

template <typename T>
__device__ __forceinline__ void ld4(const T* p, T (&v)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ld4<double>(const double* p, double (&v)[4]) {
  double2 a = *reinterpret_cast<const double2*>(p);
  double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// Row r (0..TM) of a thread's micro-tile maps to tile row  (r/4)*(BM/(TM/4)) + ty*4 + r%4.
template <class C>
__device__ __forceinline__ int tile_row(int ty, int r) { return (r >> 2) * (C::BM / (C::TM / 4)) + ty * 4 + (r & 3); }
template <class C>
__device__ __forceinline__ int tile_col(int tx, int c) { return (c >> 2) * (C::BN / (C::TN / 4)) + tx * 4 + (c & 3); }

// acc[r][c] += sum_k A(m0 + row(r), k) * B(k, n0 + col(c)) for k in [k_begin, k_end).
// al.fill(As, m0, k0, k_end) must write As[kk*LDA + i] for kk<BK, i<BM (zero where out of range);
// bl.fill(Bs, n0, k0, k_end) likewise.  All threads of the CTA must call this.
template <class C, class AL, class BL>
__device__ __forceinline__ void gemm_mainloop(const AL& al, const BL& bl, int m0, int n0, long k_begin, long k_end,
                                               typename C::elem* As, typename C::elem* Bs,
                                               typename C::elem (&acc)[C::TM][C::TN]) {
  using T = typename C::elem;
  const int tx = threadIdx.x % C::TX, ty = threadIdx.x / C::TX;
  for (long k0 = k_begin; k0 < k_end; k0 += C::BK) {
    al.fill(As, m0, k0, k_end);
    bl.fill(Bs, n0, k0, k_end);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < C::BK; ++kk) {
      T a[C::TM], b[C::TN];
#pragma unroll
      for (int r = 0; r < C::TM; r += 4) {
        T t[4];
        ld4<T>(As + kk * C::LDA + tile_row<C>(ty, r), t);
        a[r] = t[0]; a[r + 1] = t[1]; a[r + 2] = t[2]; a[r + 3] = t[3];
      }
#pragma unroll
      for (int c = 0; c < C::TN; c += 4) {
        T t[4];
        ld4<T>(Bs + kk * C::LDB + tile_col<C>(tx, c), t);
        b[c] = t[0]; b[c + 1] = t[1]; b[c + 2] = t[2]; b[c + 3] = t[3];
      }
#pragma unroll
      for (int r = 0; r < C::TM; ++r)
#pragma unroll
        for (int c = 0; c < C::TN; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
}

// Same contraction with the NEXT k block's global loads issued (into registers) before the current block's FMAs,
// so that a CTA overlaps its own memory latency even at one CTA per SM.  Needs loaders with load()/store().
template <class C, class AL, class BL>
__device__ __forceinline__ void gemm_mainloop_pf(const AL& al, const BL& bl, int m0, int n0, long k_begin, long k_end,
                                                  typename C::elem* As, typename C::elem* Bs,
                                                  typename C::elem (&acc)[C::TM][C::TN]) {
  using T = typename C::elem;
  const int tx = threadIdx.x % C::TX, ty = threadIdx.x / C::TX;
  T ra[AL::NREG], rb[BL::NREG];
  al.load(ra, m0, k_begin, k_end);
  bl.load(rb, n0, k_begin, k_end);
  for (long k0 = k_begin; k0 < k_end; k0 += C::BK) {
    al.store(As, ra);
    bl.store(Bs, rb);
    __syncthreads();
    if (k0 + C::BK < k_end) {
      al.load(ra, m0, k0 + C::BK, k_end);
      bl.load(rb, n0, k0 + C::BK, k_end);
    }
#pragma unroll
    for (int kk = 0; kk < C::BK; ++kk) {
      T a[C::TM], b[C::TN];
#pragma unroll
      for (int r = 0; r < C::TM; r += 4) {
        T t[4];
        ld4<T>(As + kk * C::LDA + tile_row<C>(ty, r), t);
        a[r] = t[0]; a[r + 1] = t[1]; a[r + 2] = t[2]; a[r + 3] = t[3];
      }
#pragma unroll
      for (int c = 0; c < C::TN; c += 4) {
        T t[4];
        ld4<T>(Bs + kk * C::LDB + tile_col<C>(tx, c), t);
        b[c] = t[0]; b[c + 1] = t[1]; b[c + 2] = t[2]; b[c + 3] = t[3];
      }
#pragma unroll
      for (int r = 0; r < C::TM; ++r)
#pragma unroll
        for (int c = 0; c < C::TN; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
}

// Strided-view loader: element (i, k) at base[i*rs + k*cs], i < rows.  TI may differ from the
// compute type (fp32 data feeding an fp64 accumulation).
template <class C, typename TI, int B /*BM or BN*/, int LD>
struct StridedLoader {
  const TI* base;
  long rs, cs;
  long rows;
  static constexpr int NREG = (B * C::BK + C::NT - 1) / C::NT;
  // load(): this thread's share of the tile into registers; store(): registers -> shared (same index map as fill())
  __device__ __forceinline__ void load(typename C::elem (&reg)[NREG], int r0, long k0, long k_end) const {
    using T = typename C::elem;
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int idx = threadIdx.x + t * C::NT;
      int i, kk;
      if (cs == 1) { kk = idx % C::BK; i = idx / C::BK; } else { i = idx % B; kk = idx / B; }
      const long gi = r0 + i, gk = k0 + kk;
      reg[t] = (idx < B * C::BK && gi < rows && gk < k_end) ? T(base[gi * rs + gk * cs]) : T(0);
    }
  }
  __device__ __forceinline__ void store(typename C::elem* S, const typename C::elem (&reg)[NREG]) const {
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int idx = threadIdx.x + t * C::NT;
      if (idx < B * C::BK) {
        int i, kk;
        if (cs == 1) { kk = idx % C::BK; i = idx / C::BK; } else { i = idx % B; kk = idx / B; }
        S[kk * LD + i] = reg[t];
      }
    }
  }
  __device__ __forceinline__ void fill(typename C::elem* S, int r0, long k0, long k_end) const {
    using T = typename C::elem;
    if (cs == 1) {  // k contiguous in memory: consecutive threads walk k
      for (int idx = threadIdx.x; idx < B * C::BK; idx += C::NT) {
        const int kk = idx % C::BK, i = idx / C::BK;
        const long gi = r0 + i, gk = k0 + kk;
        S[kk * LD + i] = (gi < rows && gk < k_end) ? T(base[gi * rs + gk]) : T(0);
      }
    } else {  // consecutive threads walk the row index
      for (int idx = threadIdx.x; idx < B * C::BK; idx += C::NT) {
        const int i = idx % B, kk = idx / B;
        const long gi = r0 + i, gk = k0 + kk;
        S[kk * LD + i] = (gi < rows && gk < k_end) ? T(base[gi * rs + gk * cs]) : T(0);
      }
    }
  }
};

// C[b] = alpha * A[b] * B[b] + beta * C[b] (+ diag on the diagonal), generic strides.
//   A(i,k) = A[b*sA + i*ars + k*acs],  B(k,j) = B[b*sB + k*brs + j*bcs],  C(i,j) = C[b*sC + i*ldc + j].
//   split_k > 1: gridDim.z = batch*split_k, partial products are atomically added to C
//   (C must be pre-initialised, beta is ignored).
//   lower_only: skip tiles strictly above the diagonal (symmetric/triangular outputs).
//   tri: structure of op(A) used to trim the K range of a tile -- 1: op(A)(i,k) = 0 for k > i (lower triangular),
//        2: op(A)(i,k) = 0 for k < i (upper triangular, e.g. the transpose of a lower factor).
//   alpha_dev: optional device scalar(s) multiplied into alpha (per batch with stride 1, shared with stride 0).
template <class C, typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(C::NT) gemm_strided_kernel(int M, int N, long K, double alpha, const TA* A, long ars,
                                                              long acs, long sA, const TB* B, long brs, long bcs,
                                                              long sB, double beta, TC* Cm, long ldc, long sC,
                                                              int split_k, double diag, int lower_only,
                                                              const float* alpha_dev, int alpha_dev_stride, int tri) {
  using T = typename C::elem;
  __shared__ __align__(16) T smem[C::SMEM_ELEMS];
  T* As = smem;
  T* Bs = smem + C::BK * C::LDA;
  const int b = blockIdx.z / split_k, ks = blockIdx.z % split_k;
  const int m0 = blockIdx.x * C::BM, n0 = blockIdx.y * C::BN;  // rows on x: R-sized row counts exceed the y limit
  if (lower_only && n0 > m0 + C::BM - 1) return;
  const long kchunk = ((K + split_k - 1) / split_k + C::BK - 1) / C::BK * C::BK;
  long k_begin = ks * kchunk;
  long k_end = (k_begin + kchunk < K) ? k_begin + kchunk : K;
  if (tri == 1 && k_end > m0 + C::BM) k_end = m0 + C::BM;
  if (tri == 2 && k_begin < m0) k_begin = (m0 / C::BK) * C::BK;
  StridedLoader<C, TA, C::BM, C::LDA> al{A + b * sA, ars, acs, M};
  StridedLoader<C, TB, C::BN, C::LDB> bl{B + b * sB, bcs, brs, N};  // "row" of the B tile is the column j
  T acc[C::TM][C::TN];
#pragma unroll
  for (int r = 0; r < C::TM; ++r)
#pragma unroll
    for (int c = 0; c < C::TN; ++c) acc[r][c] = T(0);
  if (k_begin < k_end) gemm_mainloop_pf<C>(al, bl, m0, n0, k_begin, k_end, As, Bs, acc);
  const int tx = threadIdx.x % C::TX, ty = threadIdx.x / C::TX;
  TC* Cb = Cm + b * sC;
  if (alpha_dev) alpha *= (double)alpha_dev[(long)b * alpha_dev_stride];  // device-resident scale (e.g. upstream dKL)
#pragma unroll
  for (int r = 0; r < C::TM; ++r) {
    const int i = m0 + tile_row<C>(ty, r);
    if (i >= M) continue;
#pragma unroll
    for (int c = 0; c < C::TN; ++c) {
      const int j = n0 + tile_col<C>(tx, c);
      if (j >= N) continue;
      T v = T(alpha) * acc[r][c];
      if (split_k > 1) {
        if (k_begin < k_end) atomicAdd(&Cb[i * ldc + j], TC(v));
      } else {
        if (beta != 0.0) v += T(beta) * T(Cb[i * ldc + j]);
        if (i == j) v += T(diag);
        Cb[i * ldc + j] = TC(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp64 GEMM on the DMMA path (mma.sync.m8n8k4.f64): same contract as gemm_strided_kernel.
// The fp64 pipe of B200 retires 64 FMA/clk/SM through DFMA and through DMMA alike (measured,
// tools/mma_probe.cu); what DMMA buys is operand reuse in registers: a warp tile of WM x WN 8x8 atoms
// needs WM + WN shared-memory loads per WM*WN atoms (256 FMA each), 5x less LDS traffic than the 8x8
// SIMT micro-tile, which is LDS-bound at a quarter of the pipe.
// ------------------------------------------------------------------------------------------------
template <int BM_, int BN_, int BK_, int WM_, int WN_, int MINB_ = 1>
struct DmmaCfg {
  using elem = double;
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, MINB = MINB_;
  static constexpr int WARPS_M = BM / (8 * WM), WARPS_N = BN / (8 * WN), NT = 32 * WARPS_M * WARPS_N;
  // leading dimensions = 4 (mod 16) doubles: the 16 lanes of a half-warp (4 k x 4 rows) hit 16 distinct 8-byte banks
  static constexpr int pad16(int x) { return x + ((4 - x % 16) + 16) % 16; }
  static constexpr int LDA = pad16(BM), LDB = pad16(BN);
  static constexpr int SMEM_ELEMS = BK * (LDA + LDB);
  static_assert(BM % (8 * WM) == 0 && BN % (8 * WN) == 0 && BK % 4 == 0, "tile / warp-tile mismatch");
};

// loader with the register stage kept in the INPUT type (fp32 operands cost half the registers)
template <class C, typename TI, int B, int LD>
struct DmmaLoader {
  const TI* base;
  long rs, cs;
  long rows;
  static constexpr int NREG = (B * C::BK + C::NT - 1) / C::NT;
  __device__ __forceinline__ void load(TI (&reg)[NREG], int r0, long k0, long k_end) const {
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int idx = threadIdx.x + t * C::NT;
      int i, kk;
      if (cs == 1) { kk = idx % C::BK; i = idx / C::BK; } else { i = idx % B; kk = idx / B; }
      const long gi = r0 + i, gk = k0 + kk;
      reg[t] = (idx < B * C::BK && gi < rows && gk < k_end) ? base[gi * rs + gk * cs] : TI(0);
    }
  }
  __device__ __forceinline__ void store(double* S, const TI (&reg)[NREG]) const {
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int idx = threadIdx.x + t * C::NT;
      if (idx < B * C::BK) {
        int i, kk;
        if (cs == 1) { kk = idx % C::BK; i = idx / C::BK; } else { i = idx % B; kk = idx / B; }
        S[kk * LD + i] = (double)reg[t];
      }
    }
  }
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <class C, typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(C::NT, C::MINB) gemm_dmma_kernel(int M, int N, long K, double alpha, const TA* A, long ars,
                                                             long acs, long sA, const TB* B, long brs, long bcs, long sB,
                                                             double beta, TC* Cm, long ldc, long sC, int split_k,
                                                             double diag, int lower_only, const float* alpha_dev,
                                                             int alpha_dev_stride, int tri) {
  __shared__ __align__(16) double smem[C::SMEM_ELEMS];
  double* As = smem;
  double* Bs = smem + C::BK * C::LDA;
  const int b = blockIdx.z / split_k, ks = blockIdx.z % split_k;
  const int m0 = blockIdx.x * C::BM, n0 = blockIdx.y * C::BN;
  if (lower_only && n0 > m0 + C::BM - 1) return;
  const long kchunk = ((K + split_k - 1) / split_k + C::BK - 1) / C::BK * C::BK;
  long k_begin = ks * kchunk;
  long k_end = (k_begin + kchunk < K) ? k_begin + kchunk : K;
  if (tri == 1 && k_end > m0 + C::BM) k_end = m0 + C::BM;
  if (tri == 2 && k_begin < m0) k_begin = (m0 / C::BK) * C::BK;
  DmmaLoader<C, TA, C::BM, C::LDA> al{A + b * sA, ars, acs, M};
  DmmaLoader<C, TB, C::BN, C::LDB> bl{B + b * sB, bcs, brs, N};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  double acc[C::WM][C::WN][2];
#pragma unroll
  for (int mi = 0; mi < C::WM; ++mi)
#pragma unroll
    for (int ni = 0; ni < C::WN; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  if (k_begin < k_end) {
    TA ra[DmmaLoader<C, TA, C::BM, C::LDA>::NREG];
    TB rb[DmmaLoader<C, TB, C::BN, C::LDB>::NREG];
    al.load(ra, m0, k_begin, k_end);
    bl.load(rb, n0, k_begin, k_end);
    const double* Aw = As + wm * (8 * C::WM) + g + t * C::LDA;
    const double* Bw = Bs + wn * (8 * C::WN) + g + t * C::LDB;
    for (long k0 = k_begin; k0 < k_end; k0 += C::BK) {
      al.store(As, ra);
      bl.store(Bs, rb);
      __syncthreads();
      if (k0 + C::BK < k_end) {
        al.load(ra, m0, k0 + C::BK, k_end);
        bl.load(rb, n0, k0 + C::BK, k_end);
      }
#pragma unroll
      for (int kk = 0; kk < C::BK; kk += 4) {
        double a[C::WM];
#pragma unroll
        for (int mi = 0; mi < C::WM; ++mi) a[mi] = Aw[kk * C::LDA + mi * 8];
        // B fragments in chunks of <= 7 atoms: keeps the live fragment registers low (the 13-warp tile has 104
        // accumulator registers and a 128-register budget)
        constexpr int CH = C::WN > 7 ? (C::WN + 1) / 2 : C::WN;
#pragma unroll
        for (int n0c = 0; n0c < C::WN; n0c += CH) {
          double bf[CH];
#pragma unroll
          for (int ni = 0; ni < CH; ++ni)
            if (n0c + ni < C::WN) bf[ni] = Bw[kk * C::LDB + (n0c + ni) * 8];
#pragma unroll
          for (int ni = 0; ni < CH; ++ni)
            if (n0c + ni < C::WN) {
#pragma unroll
              for (int mi = 0; mi < C::WM; ++mi) dmma884(acc[mi][n0c + ni][0], acc[mi][n0c + ni][1], a[mi], bf[ni]);
            }
        }
      }
      __syncthreads();
    }
  }
  TC* Cb = Cm + b * sC;
  if (alpha_dev) alpha *= (double)alpha_dev[(long)b * alpha_dev_stride];
#pragma unroll
  for (int mi = 0; mi < C::WM; ++mi) {
    const int i = m0 + wm * (8 * C::WM) + mi * 8 + g;
    if (i >= M) continue;
#pragma unroll
    for (int ni = 0; ni < C::WN; ++ni) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = n0 + wn * (8 * C::WN) + ni * 8 + 2 * t + e;
        if (j >= N) continue;
        double v = alpha * acc[mi][ni][e];
        if (split_k > 1) {
          if (k_begin < k_end) atomicAdd(&Cb[i * ldc + j], TC(v));
        } else {
          if (beta != 0.0) v += beta * (double)Cb[i * ldc + j];
          if (i == j) v += diag;
          Cb[i * ldc + j] = TC(v);
        }
      }
    }
  }
}

template <class Cfg, typename TA, typename TB, typename TC>
static void gemm_launch_dmma(cudaStream_t st, int M, int N, long K, double alpha, const TA* A, long ars, long acs, long sA,
                             const TB* B, long brs, long bcs, long sB, double beta, TC* Cm, long ldc, long sC, int batch,
                             int split_k, double diag, int lower_only, const float* alpha_dev, int alpha_dev_stride,
                             int tri) {
  dim3 grid(gpsa_cdiv(M, Cfg::BM), gpsa_cdiv(N, Cfg::BN), batch * split_k);
  gemm_dmma_kernel<Cfg, TA, TB, TC><<<grid, Cfg::NT, 0, st>>>(M, N, K, alpha, A, ars, acs, sA, B, brs, bcs, sB, beta, Cm,
                                                              ldc, sC, split_k, diag, lower_only, alpha_dev,
                                                              alpha_dev_stride, tri);
}

// Host-side launcher.  Picks the tile from the problem shape: fp32 128x128 (8x8 micro-tile) when both extents
// reach 96; fp64 an 8x8-micro-tile configuration (104 or 128 wide, whichever pads M x N less) for the batched
// M x M algebra, else the 64x64 / 4x4 tile.
template <class Cfg, typename TA, typename TB, typename TC>
static void gemm_launch_cfg(cudaStream_t st, int M, int N, long K, double alpha, const TA* A, long ars, long acs, long sA,
                            const TB* B, long brs, long bcs, long sB, double beta, TC* Cm, long ldc, long sC, int batch,
                            int split_k, double diag, int lower_only, const float* alpha_dev, int alpha_dev_stride,
                            int tri) {
  dim3 grid(gpsa_cdiv(M, Cfg::BM), gpsa_cdiv(N, Cfg::BN), batch * split_k);
  gemm_strided_kernel<Cfg, TA, TB, TC><<<grid, Cfg::NT, 0, st>>>(M, N, K, alpha, A, ars, acs, sA, B, brs, bcs, sB, beta,
                                                                 Cm, ldc, sC, split_k, diag, lower_only, alpha_dev,
                                                                 alpha_dev_stride, tri);
}

template <typename T, typename TA, typename TB, typename TC>
static int gemm_strided(cudaStream_t st, int M, int N, long K, double alpha, const TA* A, long ars, long acs, long sA,
                        const TB* B, long brs, long bcs, long sB, double beta, TC* Cm, long ldc, long sC, int batch,
                        int split_k = 1, double diag = 0.0, int lower_only = 0, const float* alpha_dev = nullptr,
                        int alpha_dev_stride = 0, int tri = 0) {
  if (M <= 0 || N <= 0 || batch <= 0) return GPSA_OK;
  if (split_k < 1) split_k = 1;
#define GPSA_GEMM_ARGS st, M, N, K, alpha, A, ars, acs, sA, B, brs, bcs, sB, beta, Cm, ldc, sC, batch, split_k, diag, \
                       lower_only, alpha_dev, alpha_dev_stride, tri
  auto padded = [&](int t) { return (long)gpsa_cdiv(M, t) * t * ((long)gpsa_cdiv(N, t) * t); };
  if (sizeof(T) == 4) {
    // (104 x 104 tiles, which pad M = 200 to 208 instead of 256, were measured on the two gene-batched Omega^-1 Omega_sqt
    // products: 1.33 -> 1.21 ms for Linv Osq but 1.33 -> 1.58 ms for Linv^T Y -- 169-thread CTAs, rows that straddle
    // warps -- so fp32 stays on 128 x 128; profiles/r2_final_launches_c3_summary.txt has the tile that shipped.)
    if (M >= 96 && N >= 96) gemm_launch_cfg<GemmCfg<T, 128, 128, 8, 8, 8>, TA, TB, TC>(GPSA_GEMM_ARGS);
    else gemm_launch_cfg<GemmCfg<T, 64, 64, 16, 4, 4>, TA, TB, TC>(GPSA_GEMM_ARGS);
  } else {
#ifdef GPSA_DEBUG  // experiments only: force a tile family
    static const int force = [] { const char* e = getenv("GPSA_F64_TILE"); return e ? atoi(e) : 0; }();
#else
    constexpr int force = 0;
#endif
    // small problems (the per-view warp-layer products): the 8x8-micro-tile grid would not fill the machine
    const long big_ctas = (long)gpsa_cdiv(M, 104) * gpsa_cdiv(N, 104) * batch * split_k;
    // DMMA path: 104 x 104 CTA tiles (13 warps, 8 x 104 each) when M pads well to 208 (the M = 200 algebra), else
    // 128 x 128 (8 warps, 32 x 64 each, one CTA per SM); split-K is re-derived for the tile so the grid still
    // fills the GPU
    if (force == 0 && M >= 64 && N >= 64) {
      const bool t208 = (long)gpsa_cdiv(M, 208) * 208 * ((long)gpsa_cdiv(N, 104) * 104) <=
                        (long)gpsa_cdiv(M, 128) * 128 * ((long)gpsa_cdiv(N, 128) * 128);
      const long tiles = t208 ? (long)gpsa_cdiv(M, 104) * gpsa_cdiv(N, 104) * batch
                              : (long)gpsa_cdiv(M, 128) * gpsa_cdiv(N, 128) * batch;
      int sk = split_k;
      if (split_k > 1) {
        long want = (2 * 148 + tiles - 1) / tiles;
        const long smax = K / 512 > 0 ? K / 512 : 1;
        if (want > smax) want = smax;
        sk = (int)(want < 1 ? 1 : want);
      }
      if (tiles * sk >= 120) {
        // 104 x 104 tiles, 13 warps of 8 x 104, 72 registers -> two CTAs per SM (measured 1.7x a 208 x 104 tile at one
        // CTA per SM); on the symmetric / triangular M x M algebra they also skip the upper-right quarter of the
        // output (lower_only) or a quarter of the K range (tri)
        if (t208)
          gemm_launch_dmma<DmmaCfg<104, 104, 8, 1, 13, 2>, TA, TB, TC>(st, M, N, K, alpha, A, ars, acs, sA, B, brs, bcs, sB, beta, Cm, ldc, sC, batch, sk, diag, lower_only, alpha_dev, alpha_dev_stride, tri);
        else gemm_launch_dmma<DmmaCfg<128, 128, 16, 4, 4>, TA, TB, TC>(st, M, N, K, alpha, A, ars, acs, sA, B, brs, bcs, sB, beta, Cm, ldc, sC, batch, sk, diag, lower_only, alpha_dev, alpha_dev_stride, tri);
        GPSA_LAUNCH_CHECK();
        return GPSA_OK;
      }
    }
    // single M x M products of the prior / warp-layer chains (16 CTAs of 64 x 64 at M = 200, LDS-bound on 16 SMs):
    // 32 x 32 tiles put the same work on 4x as many SMs, same k order per element
    const long small_ctas = (long)gpsa_cdiv(M, 64) * gpsa_cdiv(N, 64) * batch * split_k;
    if (force == 0 && big_ctas < 2 * 148 && small_ctas <= 74) {
      gemm_launch_cfg<GemmCfg<T, 32, 32, 16, 4, 4>, TA, TB, TC>(GPSA_GEMM_ARGS);
    } else if (force == 64 || (force == 0 && big_ctas < 2 * 148)) {
      gemm_launch_cfg<GemmCfg<T, 64, 64, 16, 4, 4>, TA, TB, TC>(GPSA_GEMM_ARGS);
    } else if (M >= 96 && N >= 96) {
      if (force == 104 || (force == 0 && padded(104) < padded(128))) gemm_launch_cfg<GemmCfg<T, 104, 104, 8, 8, 8>, TA, TB, TC>(GPSA_GEMM_ARGS);
      else gemm_launch_cfg<GemmCfg<T, 128, 128, 8, 8, 8>, TA, TB, TC>(GPSA_GEMM_ARGS);
    } else {
      gemm_launch_cfg<GemmCfg<T, 64, 64, 16, 4, 4>, TA, TB, TC>(GPSA_GEMM_ARGS);
    }
  }
#undef GPSA_GEMM_ARGS
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}
