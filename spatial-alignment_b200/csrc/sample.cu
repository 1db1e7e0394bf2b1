// Reparameterised sampling of the data GP fused with the Gaussian log-likelihood, and the counter-based noise
// behind it.
//
//   F[r,p] = mean[r,p] + sqrt(var[r,p]) eps[r,p],  var = kq[r] + q2[r,p] + 2e-5     (reference gpsa/models/vgpsa.py:197-204,
//   LL     = sum_{r,p} log N(Y[n,p]; F[r,p], sigma) / S                              :423-426, :532-538)
//
// The reference materialises eps, F and the log-prob chain as [S,N,L] tensors (1 GB each at C3, 128 GB at C5).  Here
// one kernel reads the predictive mean and the quadratic form once, draws eps from Philox4x32-10 keyed by
// (seed, sample, spot, GLOBAL gene) -- so the draw does not depend on how genes or samples are sharded over ranks --
// accumulates LL and d LL / d log_noise, and overwrites the two input buffers IN PLACE with the only things the
// backward pass needs:
//     U [r,p] = d(-LL)/dF   = -(Y - F) / (sigma^2 S)            (into the mean buffer)
//     Gu[r,p] = d(-LL)/dvar = U eps / (2 sqrt(var))             (into the q2 buffer)
// plus kqb[r] = sum_p Gu[r,p].  F, eps and var are never stored.  The backward multiplies by the upstream gradient
// only if it is not exactly 1 (gpsa_scale_if_not_one: for loss.backward() it is, and the kernel exits at once).
// Explicit noise (`eps` != NULL, parity tests) takes the same path with the draw replaced by a load.
#include "common.cuh"
#include "gpsa_b200.h"

#include <math.h>

namespace {

// ---- Philox4x32-10 (Salmon et al., SC'11): counter (4 x 32 bit), key (2 x 32 bit) -> 4 x 32 random bits ---------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t h0 = __umulhi(M0, c.x), l0 = M0 * c.x;
    const uint32_t h1 = __umulhi(M1, c.z), l1 = M1 * c.z;
    c = make_uint4(h1 ^ c.y ^ k.x, l1, h0 ^ c.w ^ k.y, l0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// two uniforms in (0, 1] -> two standard normals (Box-Muller)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  const float v = (float)b * 2.3283064365386963e-10f;           // [0, 1)
  const float r = sqrtf(-2.0f * __logf(u));
  float s, c;
  __sincosf(6.283185307179586f * v, &s, &c);
  z0 = r * c;
  z1 = r * s;
}
// the four normals of (sample s, spot n, global gene quad q): genes 4q .. 4q+3
__device__ __forceinline__ float4 normal_quad(uint2 key, uint32_t s, uint32_t n, uint32_t q) {
  const uint4 x = philox4x32_10(make_uint4(n, s, q, 0x47505341u /* "GPSA" */), key);
  float4 z;
  box_muller(x.x, x.y, z.x, z.y);
  box_muller(x.z, x.w, z.z, z.w);
  return z;
}
__device__ __forceinline__ float pick(const float4& z, int i) { return i == 0 ? z.x : i == 1 ? z.y : i == 2 ? z.z : z.w; }

struct NoiseSrc {
  const float* eps;        // explicit noise [R, L] or NULL
  const long long* key;    // device: one 64-bit seed (drawn from torch's generator by the caller)
  int gene_off, samp_off;  // global index of local gene 0 / local sample 0 (sharding)
};

// eps[r, p] for all (r, p): the materialised form of the in-kernel draw (parity tests, the non-fused path)
__global__ void __launch_bounds__(256) philox_fill_kernel(long N, int S, int L, NoiseSrc ns, float* __restrict__ out) {
  const uint2 key = make_uint2((uint32_t)ns.key[0], (uint32_t)((unsigned long long)ns.key[0] >> 32));
  const int q0 = ns.gene_off >> 2, q1 = (ns.gene_off + L + 3) >> 2, nq = q1 - q0;
  const long total = (long)S * N * nq;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / nq;
    const int q = q0 + (int)(i - r * nq);
    const int s = (int)(r / N);
    const long n = r - (long)s * N;
    const float4 z = normal_quad(key, (uint32_t)(s + ns.samp_off), (uint32_t)n, (uint32_t)q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = 4 * q + j - ns.gene_off;
      if (p >= 0 && p < L) out[r * L + p] = pick(z, j);
    }
  }
}

// One warp per (spot n, chunk of 128 global gene quads), all S samples: Y is read once per spot.
// VEC: gene_off % 4 == 0, L % 4 == 0 and 16-byte aligned rows -> 128-bit accesses.
template <bool VEC>
__global__ void __launch_bounds__(256) sample_ll_fused_kernel(long N, int S, int L, int nchunk, NoiseSrc ns,
                                                              const float* __restrict__ kq, const float* __restrict__ Y,
                                                              const float* __restrict__ log_noise, float* __restrict__ Fm,
                                                              float* __restrict__ Vq, float* __restrict__ kqb,
                                                              double* nll_acc, double* noise_acc) {
  constexpr int QPW = 128;  // quads per warp item = 4 per lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint2 key = ns.eps ? make_uint2(0, 0) : make_uint2((uint32_t)ns.key[0], (uint32_t)((unsigned long long)ns.key[0] >> 32));
  const float en = expf(log_noise[0]);
  const float sigma = en + GPSA_OFF;  // vgpsa.py:217, used as the Normal SCALE (:534)
  const float inv = 1.f / sigma;
  const float cu = -inv * inv / (float)S;  // U = cu (Y - F)
  const int q0 = ns.gene_off >> 2;
  const long items = N * nchunk;
  double zz = 0.0;  // sum of ((Y - F) / sigma)^2 over this thread's elements
  for (long it = (long)blockIdx.x * (blockDim.x >> 5) + warp; it < items; it += (long)gridDim.x * (blockDim.x >> 5)) {
    const long n = it / nchunk;
    const int ch = (int)(it - n * nchunk);
    float y[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int q = q0 + ch * QPW + k * 32 + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = 4 * q + j - ns.gene_off;
        y[k][j] = (p >= 0 && p < L) ? Y[n * L + p] : 0.f;
      }
    }
    for (int s = 0; s < S; ++s) {
      const long r = (long)s * N + n;
      const float kqr = kq[r];
      float rowsum = 0.f, part = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = q0 + ch * QPW + k * 32 + lane;
        const int p0 = 4 * q - ns.gene_off;
        if (p0 >= L || p0 + 3 < 0) continue;
        float m[4], v[4], e[4];
        if (VEC) {
          const float4 m4 = *reinterpret_cast<const float4*>(Fm + r * L + p0);
          const float4 v4 = *reinterpret_cast<const float4*>(Vq + r * L + p0);
          m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
          v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = p0 + j >= 0 && p0 + j < L;
            m[j] = ok ? Fm[r * L + p0 + j] : 0.f;
            v[j] = ok ? Vq[r * L + p0 + j] : 1.f;
          }
        }
        if (ns.eps) {
          if (VEC) {
            const float4 e4 = *reinterpret_cast<const float4*>(ns.eps + r * L + p0);
            e[0] = e4.x; e[1] = e4.y; e[2] = e4.z; e[3] = e4.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = (p0 + j >= 0 && p0 + j < L) ? ns.eps[r * L + p0 + j] : 0.f;
          }
        } else {
          const float4 z = normal_quad(key, (uint32_t)(s + ns.samp_off), (uint32_t)n, (uint32_t)q);
          e[0] = z.x; e[1] = z.y; e[2] = z.z; e[3] = z.w;
        }
        float u[4], g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = VEC || (p0 + j >= 0 && p0 + j < L);
          const float var = (kqr + v[j] + GPSA_OFF) + GPSA_OFF;  // jitter twice: vgpsa.py:201,:204
          const float sd = sqrtf(var);
          const float f = fmaf(sd, e[j], m[j]);
          const float d = y[k][j] - f;
          u[j] = cu * d;
          g[j] = 0.5f * u[j] * e[j] / sd;
          if (ok) {
            const float z = d * inv;
            part = fmaf(z, z, part);
            rowsum += g[j];
          }
        }
        if (VEC) {
          *reinterpret_cast<float4*>(Fm + r * L + p0) = make_float4(u[0], u[1], u[2], u[3]);
          *reinterpret_cast<float4*>(Vq + r * L + p0) = make_float4(g[0], g[1], g[2], g[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (p0 + j >= 0 && p0 + j < L) { Fm[r * L + p0 + j] = u[j]; Vq[r * L + p0 + j] = g[j]; }
        }
      }
      rowsum = warp_sum(rowsum);
      if (lane == 0) {
        if (nchunk == 1) kqb[r] = rowsum;
        else atomicAdd(&kqb[r], rowsum);  // kqb zeroed by the launcher
      }
      zz += (double)part;
    }
  }
  __shared__ double red[32];
  zz = block_sum<double>(zz, red);
  if (threadIdx.x == 0) {
    // -LL = sum [ z^2/2 + log sigma + log(2 pi)/2 ] / S ;  d(-LL)/dlog_noise = sum [1/sigma - z^2/sigma] exp(log_noise) / S
    double nll = 0.5 * zz, dn = -zz * (double)inv;
    if (blockIdx.x == 0) {
      const double cnt = (double)S * (double)N * (double)L;
      nll += cnt * (log((double)sigma) + 0.91893853320467274178);
      dn += cnt * (double)inv;
    }
    atomicAdd(nll_acc, nll / (double)S);
    atomicAdd(noise_acc, dn * (double)en / (double)S);
  }
}

// x *= *scale unless *scale == 1 (then every thread leaves after one load)
__global__ void __launch_bounds__(256) scale_if_not_one_kernel(long n4, long n, const float* __restrict__ scale,
                                                               float* __restrict__ x) {
  const float s = scale[0];
  if (s == 1.0f) return;
  const long stride = (long)gridDim.x * blockDim.x;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = x4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    x4[i] = v;
  }
  for (long i = 4 * n4 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

// ---- linear model of coregionalisation fused with the likelihood ------------------------------------------------
// F_obs[r, p] = sum_l F_lat[r, l] W[l, p]  (reference gpsa/models/vgpsa.py:428-432), then the same Gaussian NLL as above.
// The [S,N,P] tensor F_obs is never formed: thread = output gene p (its column of W and of W-bar in registers),
// CTA = 256 genes x a chunk of spots; per row the L latent values are broadcast loads, the contributions to
// F_lat-bar are reduced over the warp's 32 genes and added atomically.  Gradients are produced in the forward pass
// (scaled by the upstream gradient later, gpsa_scale_if_not_one).
template <int LMAX>
__global__ void __launch_bounds__(256) lmc_ll_fused_kernel(long N, int S, int L, int P, long spots_per_cta,
                                                           const float* __restrict__ Fl, const float* __restrict__ W,
                                                           const float* __restrict__ Y, const float* __restrict__ log_noise,
                                                           float* Flbar, float* Wbar, double* nll_acc, double* noise_acc) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < P;
  const float en = expf(log_noise[0]);
  const float sigma = en + GPSA_OFF;
  const float inv = 1.f / sigma;
  const float cu = -inv * inv / (float)S;
  float w[LMAX], wb[LMAX];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    w[l] = (valid && l < L) ? W[(long)l * P + p] : 0.f;
    wb[l] = 0.f;
  }
  const long n0 = (long)blockIdx.y * spots_per_cta, n1 = (n0 + spots_per_cta < N) ? n0 + spots_per_cta : N;
  double zz = 0.0;
  for (long n = n0; n < n1; ++n) {
    const float y = valid ? Y[n * P + p] : 0.f;
    float part = 0.f;
    for (int s = 0; s < S; ++s) {
      const float* fr = Fl + ((long)s * N + n) * L;
      float fl[LMAX], f = 0.f;
#pragma unroll
      for (int l = 0; l < LMAX; ++l) {
        fl[l] = (l < L) ? __ldg(fr + l) : 0.f;
        f = fmaf(fl[l], w[l], f);
      }
      const float d = valid ? y - f : 0.f;
      const float u = cu * d;
      const float z = d * inv;
      part = fmaf(z, z, part);
#pragma unroll
      for (int l = 0; l < LMAX; ++l) {
        if (l < L) {  // warp-uniform
          wb[l] = fmaf(fl[l], u, wb[l]);
          const float t = warp_sum(u * w[l]);
          if (lane == 0) atomicAdd(&Flbar[((long)s * N + n) * L + l], t);
        }
      }
    }
    zz += (double)part;
  }
  if (valid) {
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) atomicAdd(&Wbar[(long)l * P + p], wb[l]);
  }
  __shared__ double red[32];
  zz = block_sum<double>(zz, red);
  if (threadIdx.x == 0) {
    double nll = 0.5 * zz, dn = -zz * (double)inv;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      const double cnt = (double)S * (double)N * (double)P;
      nll += cnt * (log((double)sigma) + 0.91893853320467274178);
      dn += cnt * (double)inv;
    }
    atomicAdd(nll_acc, nll / (double)S);
    atomicAdd(noise_acc, dn * (double)en / (double)S);
  }
}

int grid_cap(long work, int per_block, int cap) {
  const long b = (work + per_block - 1) / per_block;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace

extern "C" int gpsa_philox_normal(long N, int S, int L, const long long* key, int gene_off, int samp_off, float* out,
                                  cudaStream_t st) {
  if (N <= 0 || S <= 0 || L <= 0) return GPSA_OK;
  if (!key || gene_off < 0 || samp_off < 0) return GPSA_ERR_ARG;
  NoiseSrc ns = {nullptr, key, gene_off, samp_off};
  const long quads = (long)S * N * (((gene_off + L + 3) >> 2) - (gene_off >> 2));
  philox_fill_kernel<<<grid_cap(quads, 256, 148 * 16), 256, 0, st>>>(N, S, L, ns, out);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_sample_ll_fused(long N, int S, int L, const float* kq, const float* Y, const float* log_noise,
                                    const float* eps, const long long* key, int gene_off, int samp_off, float* mean_U,
                                    float* q2_Gu, float* kqb, double* nll_acc, double* noise_acc, cudaStream_t st) {
  if (N <= 0 || S <= 0 || L <= 0) return GPSA_OK;
  if ((!eps && !key) || gene_off < 0 || samp_off < 0) return GPSA_ERR_ARG;
  NoiseSrc ns = {eps, key, gene_off, samp_off};
  const int nq = ((gene_off + L + 3) >> 2) - (gene_off >> 2);
  const int nchunk = (nq + 127) / 128;
  if (nchunk > 1 && cudaMemsetAsync(kqb, 0, sizeof(float) * (size_t)S * N, st) != cudaSuccess) return GPSA_ERR_CUDA;
  const long items = N * nchunk;
  const int grid = grid_cap(items, 8, 148 * 8);
  const bool vec = (gene_off & 3) == 0 && (L & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(mean_U) | reinterpret_cast<uintptr_t>(q2_Gu) |
                     reinterpret_cast<uintptr_t>(eps)) & 15) == 0;
  if (vec)
    sample_ll_fused_kernel<true><<<grid, 256, 0, st>>>(N, S, L, nchunk, ns, kq, Y, log_noise, mean_U, q2_Gu, kqb, nll_acc, noise_acc);
  else
    sample_ll_fused_kernel<false><<<grid, 256, 0, st>>>(N, S, L, nchunk, ns, kq, Y, log_noise, mean_U, q2_Gu, kqb, nll_acc, noise_acc);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_scale_if_not_one(long n, const float* scale, float* x, cudaStream_t st) {
  if (n <= 0) return GPSA_OK;
  const bool al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const long n4 = al ? n / 4 : 0;
  scale_if_not_one_kernel<<<grid_cap(n / 4 + 1, 256, 148 * 8), 256, 0, st>>>(n4, n, scale, x);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

extern "C" int gpsa_lmc_max_latent(void) { return 32; }

extern "C" int gpsa_lmc_ll_fused(long N, int S, int L, int P, const float* F_lat, const float* W, const float* Y,
                                 const float* log_noise, float* F_lat_bar, float* W_bar, double* nll_acc,
                                 double* noise_acc, cudaStream_t st) {
  if (N <= 0 || S <= 0 || L <= 0 || P <= 0) return GPSA_OK;
  if (L > 32) return GPSA_ERR_UNSUPPORTED;
  if (cudaMemsetAsync(F_lat_bar, 0, sizeof(float) * (size_t)S * N * L, st) != cudaSuccess ||
      cudaMemsetAsync(W_bar, 0, sizeof(float) * (size_t)L * P, st) != cudaSuccess)
    return GPSA_ERR_CUDA;
  const int nx = gpsa_cdiv(P, 256);
  long ny = (148L * 4 + nx - 1) / nx;
  if (ny > N) ny = N;
  if (ny < 1) ny = 1;
  const long per = (N + ny - 1) / ny;
  ny = (N + per - 1) / per;
  const dim3 grid((unsigned)nx, (unsigned)ny);
  if (L <= 8) lmc_ll_fused_kernel<8><<<grid, 256, 0, st>>>(N, S, L, P, per, F_lat, W, Y, log_noise, F_lat_bar, W_bar, nll_acc, noise_acc);
  else if (L <= 16) lmc_ll_fused_kernel<16><<<grid, 256, 0, st>>>(N, S, L, P, per, F_lat, W, Y, log_noise, F_lat_bar, W_bar, nll_acc, noise_acc);
  else lmc_ll_fused_kernel<32><<<grid, 256, 0, st>>>(N, S, L, P, per, F_lat, W, Y, log_noise, F_lat_bar, W_bar, nll_acc, noise_acc);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}
