// The hot contraction of the path: the per-gene marginal-variance quadratic form
//     q2[r,p] = a_r^T Omega_p a_r ,   r = (sample s, spot n),  a_r = column r of A = K_uu^-1 K_uf
// and its two backward products.  The reference computes it by materialising an [S,L,N,M]
// tensor through a broadcast bmm (gpsa/models/vgpsa.py:193-196); here it is one GEMM over
// IMPLICIT features:
//     phi_r[(i,j)] = a_r[i] a_r[j]            (generated on the fly, never stored)
//     W[(i,j), p]  = c_ij Omega_p[i,j]        (packed once per iteration, c = 1 on diagonal blocks, 2 above)
//     q2   = Phi  W            [R,K] x [K,L]      forward
//     H    = Phi^T G           [K,R] x [R,L]      Omega-bar   (G = dLoss/dq2)
//     Psi  = G W^T             [R,L] x [L,K]      contracted with a in the epilogue -> A-bar
// Only the upper block-triangle of 8x8 index blocks is enumerated, so the flop count is the
// symmetric minimum M(M+1) per (r,p) instead of the 2 M^2 of the dense A^T L product.
//
// This file is the fp32 SIMT engine for those three GEMMs (exact fp32 accumulate); the
// tcgen05 engine shares the packing and the feature ordering.
#include "gemm.cuh"
#include "gpsa_b200.h"

namespace {

constexpr int FB = 8;          // index-block edge
constexpr int FBK = FB * FB;   // features per block

__host__ __device__ inline int feat_nb(int M) { return (M + FB - 1) / FB; }
__host__ __device__ inline long feat_nblk(int M) { const long nb = feat_nb(M); return nb * (nb + 1) / 2; }

__device__ __forceinline__ void decode_block(int b, int nb, int& I, int& J) {
  int i = 0, rem = b;
  while (rem >= nb - i) { rem -= nb - i; ++i; }
  I = i; J = i + rem;
}
__device__ __forceinline__ int encode_block(int I, int J, int nb) { return I * nb - I * (I - 1) / 2 + (J - I); }

// W[k,p] from Omega [L,M,M]
__global__ void __launch_bounds__(256) feat_pack_kernel(int M, int L, const float* __restrict__ Omega,
                                                        float* __restrict__ W) {
  const int nb = feat_nb(M);
  const int b = blockIdx.x;
  int I, J;
  decode_block(b, nb, I, J);
  const float c = (I == J) ? 1.f : 2.f;
  for (int p0 = blockIdx.y * 32; p0 < L; p0 += gridDim.y * 32) {
    for (int idx = threadIdx.x; idx < FBK * 32; idx += blockDim.x) {
      const int pl = idx % 32, f = idx / 32;
      const int p = p0 + pl;
      if (p >= L) continue;
      const int i = I * FB + f / FB, j = J * FB + f % FB;
      const float v = (i < M && j < M) ? c * Omega[((long)p * M + i) * M + j] : 0.f;
      W[((long)b * FBK + f) * L + p] = v;
    }
  }
}

// Obar[p,i,j] (symmetric) = H[(i,j),p] + add_scale * Add[i,j]
// One CTA per (index block b = (I,J), 32 genes): the 64 x 32 piece of H is read along the genes (128-byte rows) into
// shared memory and written out as the 8 x 8 block (I,J) of each gene's matrix -- and, off the diagonal, its mirror
// image (J,I) -- in 32-byte row segments.  (A thread per output element reading H[k(i,j), p] directly touched one
// 32-byte sector per value: 0.49 ms at C3 for 0.5 GB of useful traffic.)
__global__ void __launch_bounds__(256) feat_unpack_kernel(int M, int L, const float* __restrict__ H,
                                                          const float* __restrict__ Add, float add_scale,
                                                          const float* __restrict__ add_scale_dev,
                                                          float* __restrict__ Obar) {
  __shared__ float t[FBK][33];
  const int nb = feat_nb(M);
  const int b = blockIdx.x;
  int I, J;
  decode_block(b, nb, I, J);
  if (add_scale_dev) add_scale *= add_scale_dev[0];
  const long MM = (long)M * M;
  for (int p0 = blockIdx.y * 32; p0 < L; p0 += gridDim.y * 32) {
    for (int idx = threadIdx.x; idx < FBK * 32; idx += 256) {
      const int pl = idx & 31, f = idx >> 5;
      t[f][pl] = (p0 + pl < L) ? H[((long)b * FBK + f) * L + p0 + pl] : 0.f;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < FBK * 32; idx += 256) {
      const int f = idx & (FBK - 1), pl = idx >> 6;
      const int p = p0 + pl;
      const int il = f >> 3, jl = f & 7;
      const int i = I * FB + il, j = J * FB + jl;
      if (p < L && i < M && j < M) {
        float v = t[f][pl];
        if (Add) v += add_scale * Add[(long)i * M + j];
        Obar[p * MM + (long)i * M + j] = v;
      }
    }
    if (I != J) {
      for (int idx = threadIdx.x; idx < FBK * 32; idx += 256) {
        const int g = idx & (FBK - 1), pl = idx >> 6;
        const int p = p0 + pl;
        const int jl = g >> 3, il = g & 7;  // il fastest: consecutive threads walk a row of the mirrored block
        const int i = I * FB + il, j = J * FB + jl;
        if (p < L && i < M && j < M) {
          float v = t[il * FB + jl][pl];
          if (Add) v += add_scale * Add[(long)j * M + i];
          Obar[p * MM + (long)j * M + i] = v;
        }
      }
    }
    __syncthreads();
  }
}

using CfgF = GemmCfg<float, 128, 128, 16, 8, 8>;

// A-operand generator for the forward GEMM: rows = r, k = feature.
struct FeatRowLoader {
  const float* A;  // [M,R]
  int M, nb;
  long R;
  mutable int cb = -1, cI = 0, cJ = 0;  // block cursor: k0 advances monotonically, so (I,J) is stepped, not decoded
  __device__ __forceinline__ void fill(float* S, int r0, long k0, long) const {
    // BK = 16 features = two `il` rows of one block
    const int b = (int)(k0 / FBK);
    if (b != cb) {
      if (cb >= 0 && b == cb + 1) {
        if (++cJ == nb) { ++cI; cJ = cI; }
      } else {
        decode_block(b, nb, cI, cJ);
      }
      cb = b;
    }
    const int I = cI, J = cJ;
    const int il0 = (int)(k0 % FBK) / FB;
    for (int idx = threadIdx.x; idx < CfgF::BM * CfgF::BK; idx += CfgF::NT) {
      const int i = idx % CfgF::BM, kk = idx / CfgF::BM;
      const long r = r0 + i;
      const int mi = I * FB + il0 + kk / FB, mj = J * FB + kk % FB;
      float v = 0.f;
      if (r < R && mi < M && mj < M) v = A[(long)mi * R + r] * A[(long)mj * R + r];
      S[kk * CfgF::LDA + i] = v;
    }
  }
};

__global__ void __launch_bounds__(CfgF::NT) feat_fwd_kernel(int M, long R, int L, const float* __restrict__ A,
                                                            const float* __restrict__ W, float* __restrict__ q2) {
  __shared__ __align__(16) float smem[CfgF::SMEM_ELEMS];
  float* As = smem;
  float* Bs = smem + CfgF::BK * CfgF::LDA;
  const int m0 = blockIdx.x * CfgF::BM, n0 = blockIdx.y * CfgF::BN;
  const long K = feat_nblk(M) * FBK;
  FeatRowLoader al;
  al.A = A; al.M = M; al.nb = feat_nb(M); al.R = R;
  StridedLoader<CfgF, float, CfgF::BN, CfgF::LDB> bl{W, 1, (long)L, (long)L};
  float acc[CfgF::TM][CfgF::TN];
#pragma unroll
  for (int r = 0; r < CfgF::TM; ++r)
#pragma unroll
    for (int c = 0; c < CfgF::TN; ++c) acc[r][c] = 0.f;
  gemm_mainloop<CfgF>(al, bl, m0, n0, 0, K, As, Bs, acc);
  const int tx = threadIdx.x % CfgF::TX, ty = threadIdx.x / CfgF::TX;
#pragma unroll
  for (int r = 0; r < CfgF::TM; ++r) {
    const long i = m0 + tile_row<CfgF>(ty, r);
    if (i >= R) continue;
#pragma unroll
    for (int c = 0; c < CfgF::TN; ++c) {
      const int j = n0 + tile_col<CfgF>(tx, c);
      if (j < L) q2[i * L + j] = acc[r][c];
    }
  }
}

// A-operand generator for H = Phi^T G: rows = feature, k = r.
struct FeatColLoader {
  const float* A;
  int M, nb;
  long R;
  long nfeat;
  int bI[CfgF::BM / FBK], bJ[CfgF::BM / FBK];  // the tile's feature blocks, decoded once per CTA
  __device__ __forceinline__ void fill(float* S, int f0, long k0, long k_end) const {
    for (int idx = threadIdx.x; idx < CfgF::BM * CfgF::BK; idx += CfgF::NT) {
      const int kk = idx % CfgF::BK, i = idx / CfgF::BK;  // consecutive threads walk r (contiguous in A)
      const long f = f0 + i, r = k0 + kk;
      float v = 0.f;
      if (f < nfeat && r < k_end) {
        const int I = bI[i / FBK], J = bJ[i / FBK];
        const int mi = I * FB + (int)(f % FBK) / FB, mj = J * FB + (int)(f % FB);
        if (mi < M && mj < M) v = A[(long)mi * R + r] * A[(long)mj * R + r];
      }
      S[kk * CfgF::LDA + i] = v;
    }
  }
};

__global__ void __launch_bounds__(CfgF::NT) feat_bwd_omega_kernel(int M, long R, int L, const float* __restrict__ A,
                                                                  const float* __restrict__ G, float* H, int split_k) {
  __shared__ __align__(16) float smem[CfgF::SMEM_ELEMS];
  float* As = smem;
  float* Bs = smem + CfgF::BK * CfgF::LDA;
  const int m0 = blockIdx.x * CfgF::BM, n0 = blockIdx.y * CfgF::BN;
  const long nfeat = feat_nblk(M) * FBK;
  const long kchunk = ((R + split_k - 1) / split_k + CfgF::BK - 1) / CfgF::BK * CfgF::BK;
  const long k_begin = (long)blockIdx.z * kchunk;
  const long k_end = (k_begin + kchunk < R) ? k_begin + kchunk : R;
  if (k_begin >= k_end) return;
  FeatColLoader al{A, M, feat_nb(M), R, nfeat, {0, 0}, {0, 0}};
  static_assert(CfgF::BM / FBK == 2, "two feature blocks per tile");
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int b = m0 / FBK + q;
    if (b < feat_nblk(M)) decode_block(b, al.nb, al.bI[q], al.bJ[q]);
  }
  StridedLoader<CfgF, float, CfgF::BN, CfgF::LDB> bl{G, 1, (long)L, (long)L};
  float acc[CfgF::TM][CfgF::TN];
#pragma unroll
  for (int r = 0; r < CfgF::TM; ++r)
#pragma unroll
    for (int c = 0; c < CfgF::TN; ++c) acc[r][c] = 0.f;
  gemm_mainloop<CfgF>(al, bl, m0, n0, k_begin, k_end, As, Bs, acc);
  const int tx = threadIdx.x % CfgF::TX, ty = threadIdx.x / CfgF::TX;
#pragma unroll
  for (int r = 0; r < CfgF::TM; ++r) {
    const long f = m0 + tile_row<CfgF>(ty, r);
    if (f >= nfeat) continue;
#pragma unroll
    for (int c = 0; c < CfgF::TN; ++c) {
      const int j = n0 + tile_col<CfgF>(tx, c);
      if (j >= L) continue;
      if (split_k > 1) atomicAdd(&H[f * L + j], acc[r][c]);
      else H[f * L + j] = acc[r][c];
    }
  }
}

// Psi = G W^T with the epilogue contraction  Abar[:, r] += 2 * (sum_p G[r,p] Omega_p) a_r
__global__ void __launch_bounds__(CfgF::NT) feat_bwd_alpha_kernel(int M, long R, int L, const float* __restrict__ A,
                                                                  const float* __restrict__ G,
                                                                  const float* __restrict__ W, float* Abar) {
  __shared__ __align__(16) float smem[CfgF::SMEM_ELEMS];
  float* As = smem;
  float* Bs = smem + CfgF::BK * CfgF::LDA;
  const int m0 = blockIdx.x * CfgF::BM, n0 = blockIdx.y * CfgF::BN;
  const long nfeat = feat_nblk(M) * FBK;
  const int nb = feat_nb(M);
  StridedLoader<CfgF, float, CfgF::BM, CfgF::LDA> al{G, (long)L, 1, R};
  StridedLoader<CfgF, float, CfgF::BN, CfgF::LDB> bl{W, (long)L, 1, nfeat};
  float acc[CfgF::TM][CfgF::TN];
#pragma unroll
  for (int r = 0; r < CfgF::TM; ++r)
#pragma unroll
    for (int c = 0; c < CfgF::TN; ++c) acc[r][c] = 0.f;
  gemm_mainloop<CfgF>(al, bl, m0, n0, 0, L, As, Bs, acc);
  const int tx = threadIdx.x % CfgF::TX, ty = threadIdx.x / CfgF::TX;
  // a thread's 8 columns are two runs of 4 consecutive features: same block, same il, jl..jl+3
#pragma unroll
  for (int cc = 0; cc < CfgF::TN; cc += 4) {
    const long f = n0 + tile_col<CfgF>(tx, cc);
    if (f >= nfeat) continue;
    int I, J;
    decode_block((int)(f / FBK), nb, I, J);
    const int mi = I * FB + (int)(f % FBK) / FB;
    const int mj0 = J * FB + (int)(f % FB);
    if (mi >= M) continue;
    const bool diag = (I == J);
#pragma unroll
    for (int r = 0; r < CfgF::TM; ++r) {
      const long row = m0 + tile_row<CfgF>(ty, r);
      if (row >= R) continue;
      const float ai = A[(long)mi * R + row];
      float si = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int mj = mj0 + q;
        if (mj >= M) continue;
        const float psi = acc[r][cc + q];
        si = fmaf(psi, A[(long)mj * R + row], si);
        if (!diag) atomicAdd(&Abar[(long)mj * R + row], psi * ai);
      }
      atomicAdd(&Abar[(long)mi * R + row], diag ? 2.f * si : si);
    }
  }
}

}  // namespace

extern "C" long gpsa_feat_count(int M) { return feat_nblk(M) * FBK; }

extern "C" int gpsa_feat_pack(int M, int L, const float* Omega, float* W, cudaStream_t st) {
  if (M <= 0 || L <= 0) return GPSA_OK;
  dim3 grid((unsigned)feat_nblk(M), (unsigned)((L + 31) / 32 < 64 ? (L + 31) / 32 : 64));
  feat_pack_kernel<<<grid, 256, 0, st>>>(M, L, Omega, W);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_feat_unpack(int M, int L, const float* H, const float* Add, float add_scale,
                                const float* add_scale_dev, float* Obar, cudaStream_t st) {
  if (M <= 0 || L <= 0) return GPSA_OK;
  dim3 grid((unsigned)feat_nblk(M), (unsigned)((L + 31) / 32 < 64 ? (L + 31) / 32 : 64));
  feat_unpack_kernel<<<grid, 256, 0, st>>>(M, L, H, Add, add_scale, add_scale_dev, Obar);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_quadform_fwd_f32(int M, long R, int L, const float* A, const float* W, float* q2,
                                     cudaStream_t st) {
  if (M <= 0 || R <= 0 || L <= 0) return GPSA_OK;
  dim3 grid(gpsa_cdiv(R, CfgF::BM), gpsa_cdiv(L, CfgF::BN));
  feat_fwd_kernel<<<grid, CfgF::NT, 0, st>>>(M, R, L, A, W, q2);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_quadform_bwd_omega_f32(int M, long R, int L, const float* A, const float* G, float* H,
                                           cudaStream_t st) {
  if (M <= 0 || R <= 0 || L <= 0) return GPSA_OK;
  const long nfeat = feat_nblk(M) * FBK;
  const int tiles = gpsa_cdiv(nfeat, CfgF::BM) * gpsa_cdiv(L, CfgF::BN);
  int split = 1;
  if (tiles < 148 * 2) {
    split = (148 * 2 + tiles - 1) / tiles;
    const long max_split = (R + 255) / 256;
    if (split > max_split) split = (int)max_split;
    if (split < 1) split = 1;
  }
  if (split > 1) {
    if (cudaMemsetAsync(H, 0, sizeof(float) * nfeat * L, st) != cudaSuccess) return GPSA_ERR_CUDA;
  }
  dim3 grid(gpsa_cdiv(nfeat, CfgF::BM), gpsa_cdiv(L, CfgF::BN), split);
  feat_bwd_omega_kernel<<<grid, CfgF::NT, 0, st>>>(M, R, L, A, G, H, split);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_quadform_bwd_alpha_f32(int M, long R, int L, const float* A, const float* G, const float* W,
                                           float* Abar, cudaStream_t st) {
  if (M <= 0 || R <= 0 || L <= 0) return GPSA_OK;
  const long nfeat = feat_nblk(M) * FBK;
  dim3 grid(gpsa_cdiv(R, CfgF::BM), gpsa_cdiv(nfeat, CfgF::BN));
  feat_bwd_alpha_kernel<<<grid, CfgF::NT, 0, st>>>(M, R, L, A, G, W, Abar);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}
