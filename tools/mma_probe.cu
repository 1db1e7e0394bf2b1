// Micro-probe (B200): how long does a chain of DEPENDENT tcgen05.mma (same TMEM accumulator) take per
// instruction as a function of N, and does interleaving independent accumulators hide that latency?
// Also: throughput of fp64 DFMA vs mma.sync.m8n8k4.f64 per SM.  Build: tools/build_probe.sh; run on the GPU box.
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"
using namespace tc;

__global__ void __launch_bounds__(128, 1) mma_chain_kernel(int N, int chains, int iters, int cstride, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;              // 128 rows x 64 B (SW64), 8 KB
  uint8_t* sB = smem + 8192;       // 256 rows x 64 B, 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (8192 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint64_t a = make_desc_sw64(smem_u32(sA)), b = make_desc_sw64(smem_u32(sB));
    const uint32_t idesc = make_idesc_bf16(128, N);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 = warm-up
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const int c = i % chains;
          umma_bf16(tmem + (uint32_t)(c * cstride), a + ((i & 1) ? 2 : 0), b + ((i & 1) ? 2 : 0), idesc, i >= chains ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// fp64 throughput probes: 8 independent accumulator chains per thread
__global__ void dfma_kernel(int iters, double* out, long long* cyc) {
  double a[8], x = 1.0000001 + threadIdx.x * 1e-9, y = 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void dmma_kernel(int iters, double* out, long long* cyc) {
  double c[8][2], a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  const int smem = 8192 + 16384 + 1024 + 256;
  cudaFuncSetAttribute(mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  printf("tcgen05.mma M=128 K=16 bf16, dependent chains: cycles per MMA\n");
  printf("%6s %8s %8s %8s %8s\n", "N", "1chain", "2chain", "4chain", "N/2");
  const int Ns[] = {16, 32, 64, 96, 112, 128, 160, 192, 208, 256};
  for (int N : Ns) {
    printf("%6d", N);
    for (int chains : {1, 2, 4}) {
      const int cstride = chains == 4 ? 128 : 256;
      if (chains == 4 && N > 128) { printf(" %8s", "-"); continue; }
      mma_chain_kernel<<<1, 128, smem>>>(N, chains, iters, cstride, d_out);
      long long h = 0;
      cudaError_t e = cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf(" err:%s\n", cudaGetErrorString(e)); return 1; }
      printf(" %8.1f", (double)h / iters);
    }
    printf(" %8.1f\n", N / 2.0);
  }
  // all SMs busy: does the per-MMA time change under chip-wide load (power)?
  mma_chain_kernel<<<148, 128, smem>>>(208, 1, iters * 8, 256, d_out);
  long long h = 0;
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  printf("148 CTAs, N=208, 1 chain: %.1f cycles/MMA\n", (double)h / (iters * 8));
  mma_chain_kernel<<<148, 128, smem>>>(112, 2, iters * 8, 256, d_out);
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  printf("148 CTAs, N=112, 2 chains: %.1f cycles/MMA\n", (double)h / (iters * 8));

  double* d_d;
  cudaMalloc(&d_d, 148 * 8 * 1024 * 8);
  for (int threads : {128, 256, 512, 1024}) {
    const int it = 4096;
    dfma_kernel<<<148, threads>>>(it, d_d, d_out);
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    const double fma_per_clk = (double)threads * 8 * it / h;
    dmma_kernel<<<148, threads>>>(it, d_d, d_out);
    long long h2 = 0;
    cudaMemcpy(&h2, d_out, 8, cudaMemcpyDeviceToHost);
    const double mma_fma_per_clk = (double)(threads / 32) * 8 * it * 256 / h2;
    printf("fp64 per SM, %4d threads: DFMA %.1f fma/clk   DMMA(m8n8k4) %.1f fma/clk\n", threads, fma_per_clk, mma_fma_per_clk);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(e));
  return 0;
}
