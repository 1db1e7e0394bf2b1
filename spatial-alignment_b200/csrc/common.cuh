// Shared helpers for the gpsa_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GPSA_OK 0
#define GPSA_ERR_ARG 1
#define GPSA_ERR_CUDA 2
#define GPSA_ERR_UNSUPPORTED 3

#define GPSA_KIND_RBF 0
#define GPSA_KIND_MATERN12 1
#define GPSA_KIND_MATERN32 2
#define GPSA_KIND_EXTERNAL 3  // K_uu / K_uf evaluated by the caller (user-supplied covariance function)

#define GPSA_OFF 1e-5f  // diagonal_offset, reference gpsa/models/gpsa.py:153

// Every kernel launch of the library is followed by this check; it also counts launches so that
// bench.py can report how many of OUR kernels ran inside a timed region (gpsa_launch_count()).
extern long g_gpsa_launches;
#define GPSA_LAUNCH_CHECK()                       \
  do {                                            \
    ++g_gpsa_launches;                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return GPSA_ERR_CUDA; \
  } while (0)

// Optional in-library timing of the hot kernels with CUDA events on the launching stream
// (gpsa_prof_enable / gpsa_prof_read).  slot: 0 = quadform fwd, 1 = quadform bwd (A-bar), 2 = quadform bwd (Omega-bar)
#define GPSA_PROF_SLOTS 4
void gpsa_prof_begin(int slot, cudaStream_t st);
void gpsa_prof_end(int slot, cudaStream_t st);

static inline int gpsa_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute state is per DEVICE: the launchers that opt into large dynamic shared memory cache what they
// have set in arrays indexed by the current device ordinal (one process may drive several GPUs)
#define GPSA_MAX_DEVICES 64
static inline int gpsa_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < GPSA_MAX_DEVICES) ? d : 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` needs >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : T(0);
  if (w == 0) v = warp_sum(v);
  return v;
}
