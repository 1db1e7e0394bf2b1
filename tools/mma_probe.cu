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

// same chain with the A operand in TENSOR MEMORY (tcgen05.mma TS form): does the ~58-cycle floor of small-N MMAs
// (the 4 KB A fetch from shared memory) go away?
// (umma_bf16_ts now lives in tc_common.cuh)

__global__ void __launch_bounds__(128, 1) mma_chain_ts_kernel(int N, int iters, int mix, long long* out, float* check) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8192;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  // B = identity-like pattern: B[n][k] = (n == k) ? 1 : 0 for the first 16 columns (K-major rows of 64 B, SW64)
  for (int i = threadIdx.x; i < (8192 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (threadIdx.x < 16) {
    const int n = threadIdx.x, k = threadIdx.x;  // element (row n, col k) of B, bf16 1.0 = 0x3F80
    const int chunk = (k >> 3) ^ ((n >> 1) & 3);
    reinterpret_cast<uint16_t*>(sB + n * 64 + chunk * 16)[k & 7] = 0x3F80;
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  // A in TMEM columns 448..455: row r (lane) holds bf16 pairs (k = 2c, 2c+1) = (r + 2c, r + 2c + 1) as small integers
  {
    uint32_t v[8];
    const int r = threadIdx.x;
    for (int c = 0; c < 8; ++c) {
      const __nv_bfloat162 h = __floats2bfloat162_rn((float)((r + 2 * c) & 63), (float)((r + 2 * c + 1) & 63));
      v[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 448;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint64_t a = make_desc_sw64(smem_u32(sA)), b = make_desc_sw64(smem_u32(sB));
    const uint32_t idesc = make_idesc_bf16(128, N);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          if (mix && (i % 3 == 1)) umma_bf16(tmem, a, b, idesc, i > 0 ? 1u : 0u);  // 2 TS : 1 SS like the 3-pass product
          else umma_bf16_ts(tmem, tmem + 448, b, idesc, i > 0 ? 1u : 0u);
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
  tc_fence_after();
  // numerical check of the TS layout: one fresh MMA D = A * B^T with N = 16 -> D[r][n] = A[r][n]
  if (check) {
    if (warp == 0) {
      if (elect_one()) {
        umma_bf16_ts(tmem + 256, tmem + 448, make_desc_sw64(smem_u32(sB)), make_idesc_bf16(128, 16), 0u);
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 256, v);
    tmem_ld_wait();
    for (int n = 0; n < 16; ++n) check[threadIdx.x * 16 + n] = __uint_as_float(v[n]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
// cta_group::2 probe: a CTA pair issues ONE tcgen05.mma of M = 256 (each CTA its own 128 rows of A and of D in its
// own TMEM) with the B operand split by rows between the two CTAs' shared memories (CTA 0: rows [0, N/2), CTA 1:
// rows [N/2, N)).  Checks that understanding numerically and times a dependent chain in the leader.
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma2_probe_kernel(int N, int iters, long long* out, float* check) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8192;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int crank = (int)cluster_ctarank();
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (8192 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  {  // A[grow][k] = (grow + k) & 15, grow = 128 * crank + r; K-major rows of 64 B, 64-byte swizzle
    const int r = threadIdx.x, grow = 128 * crank + r;
    for (int k = 0; k < 16; ++k) {
      const int chunk = (k >> 3) ^ ((r >> 1) & 3);
      reinterpret_cast<__nv_bfloat16*>(sA + r * 64 + chunk * 16)[k & 7] = __float2bfloat16_rn((float)((grow + k) & 15));
    }
    // this CTA's half of B: local row j <-> n = crank * N/2 + j;  B[n][k] = (k == n % 16) ? n + 1 : 0
    for (int j = threadIdx.x; j < N / 2; j += blockDim.x) {
      const int n = crank * (N / 2) + j, k = n & 15;
      const int chunk = (k >> 3) ^ ((j >> 1) & 3);
      reinterpret_cast<__nv_bfloat16*>(sB + j * 64 + chunk * 16)[k & 7] = __float2bfloat16_rn((float)(n + 1));
    }
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (crank == 0) {
      const uint64_t a = make_desc_sw64(smem_u32(sA)), b = make_desc_sw64(smem_u32(sB));
      const uint32_t idesc = make_idesc_bf16(256, N);
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t acc = i > 0 ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
              "l"(a), "l"(b), "r"(idesc), "r"(acc)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
      }
      __syncwarp();
    }
    mbar_wait(bar, 0);  // both CTAs: the multicast commit arrives on each CTA's own barrier
    t1 = clock64();
    if (crank == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (check) {  // D[grow][n] for n < 64: lane = row, 32 columns per load
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      for (int n = 0; n < 32; ++n) check[((long)crank * 128 + threadIdx.x) * 64 + c * 32 + n] = __uint_as_float(v[n]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
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

  {
    cudaFuncSetAttribute(mma_chain_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* d_chk;
    cudaMalloc(&d_chk, 128 * 16 * 4);
    printf("A operand in TMEM (TS form): cycles per MMA\n%6s %8s %8s %8s\n", "N", "TS", "2TS:1SS", "N/2");
    for (int N : Ns) {
      printf("%6d", N);
      for (int mix : {0, 1}) {
        mma_chain_ts_kernel<<<1, 128, smem>>>(N, 510, mix, d_out, nullptr);
        long long hh = 0;
        cudaError_t e = cudaMemcpy(&hh, d_out, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf(" err:%s\n", cudaGetErrorString(e)); return 1; }
        printf(" %8.1f", (double)hh / 510);
      }
      printf(" %8.1f\n", N / 2.0);
    }
    mma_chain_ts_kernel<<<1, 128, smem>>>(16, 3, 0, d_out, d_chk);
    float hc[128 * 16];
    cudaError_t e = cudaMemcpy(hc, d_chk, sizeof(hc), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 16; ++n)
        if (hc[r * 16 + n] != (float)((r + n) & 63)) ++bad;
    printf("TS layout check (lane = row, column c = bf16 pair (2c, 2c+1)): %s, %d mismatches; row 5: %g %g %g %g ... (%s)\n",
           bad ? "MISMATCH" : "ok", bad, hc[80], hc[81], hc[82], hc[83], cudaGetErrorString(e));
  }
  {
    cudaFuncSetAttribute(mma2_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* d_chk2;
    cudaMalloc(&d_chk2, 256 * 64 * 4);
    mma2_probe_kernel<<<2, 128, smem>>>(64, 1, d_out, d_chk2);
    static float hc2[256 * 64];
    cudaError_t e = cudaMemcpy(hc2, d_chk2, sizeof(hc2), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 256; ++r)
      for (int n = 0; n < 64; ++n)
        if (hc2[r * 64 + n] != (float)((n + 1) * ((r + (n & 15)) & 15))) ++bad;
    printf("cta_group::2 check (M = 256 over a CTA pair, B rows [0,N/2) in CTA 0 and [N/2,N) in CTA 1, N = 64): %s, %d mismatches "
           "(%s); D[1][0..3] = %g %g %g %g, D[129][32..35] = %g %g %g %g\n",
           bad ? "MISMATCH" : "ok", bad, cudaGetErrorString(e), hc2[64], hc2[65], hc2[66], hc2[67], hc2[129 * 64 + 32],
           hc2[129 * 64 + 33], hc2[129 * 64 + 34], hc2[129 * 64 + 35]);
    if (e == cudaSuccess) {
      printf("cta_group::2 dependent chain, cycles per MMA (M = 256):");
      for (int N : {32, 64, 128, 208, 256}) {
        mma2_probe_kernel<<<2, 128, smem>>>(N, 512, d_out, nullptr);
        long long hh = 0;
        if (cudaMemcpy(&hh, d_out, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf(" err"); break; }
        printf("  N=%d: %.1f", N, (double)hh / 512);
      }
      printf("\n");
    }
  }
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
