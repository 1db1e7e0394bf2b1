// tcgen05 engine for the hot contraction of the path: the per-gene marginal-variance quadratic form
//     q2[r,p] = a_r^T Omega_p a_r          (reference gpsa/models/vgpsa.py:193-196, the [S,L,N,M] bmm)
// and its two backward products.  All three run on the 5th-generation tensor cores as bf16 x bf16 -> fp32
// MMAs with every fp32 operand split into bf16 (hi, lo) and three passes (hi*hi + lo*hi + hi*lo), which
// reproduces an fp32 product to ~2^-16 while keeping fp32 accumulation in TMEM.
//
//   forward      T_p = A_tile^T L_p  (L_p = chol(Omega_p), the reference's own formulation ||a^T L||^2):
//                the 128-row A tile stays resident in shared memory, the transposed factor of each gene is
//                streamed by TMA in reverse K order so that lower-triangular zeros are never multiplied
//                (tcgen05.mma N shrinks with K), sum of squares in the epilogue straight out of TMEM.
//   A-bar        Psi = G W^T over the packed symmetric features of Omega (feat.cu ordering), contracted with
//                a_r in the epilogue:  Abar[:, r] += 2 (sum_p G[r,p] Omega_p) a_r.
//   Omega-bar    H = Phi^T G with the feature operand Phi[r,(i,j)] = a_r[i] a_r[j] generated on the fly into
//                swizzled shared memory by eight generator warps (never stored) from TMA-staged rows of A,
//                G^T streamed by TMA.
//
// Kernel anatomy (all three): warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
// allocator, warps 4-7 = epilogue (TMEM lane quadrant = warp % 4), [warps 8-15 = operand generators];
// persistent CTAs, one per SM, mbarrier full/empty rings for shared memory and for the two TMEM
// accumulator stages.
#include "common.cuh"
#include "gpsa_b200.h"
#include "tc_common.cuh"

#include <stdlib.h>

using namespace tc;

namespace {

constexpr int FB = 8, FBK = 64;  // feature blocks, identical to feat.cu
__host__ __device__ inline int feat_nb(int M) { return (M + FB - 1) / FB; }
__host__ __device__ inline long feat_nblk(int M) { const long nb = feat_nb(M); return nb * (nb + 1) / 2; }
__device__ __forceinline__ void decode_block(int b, int nb, int& I, int& J) {
  int i = 0, rem = b;
  while (rem >= nb - i) { rem -= nb - i; ++i; }
  I = i; J = i + rem;
}

constexpr int TM = 128;                       // rows of one accumulator = TMEM lanes
constexpr int TN = 256;                       // columns of one accumulator
constexpr int A_TILE_BYTES = TM * ROW_BYTES;  // 16 KB
constexpr int B_TILE_BYTES = TN * ROW_BYTES;  // 32 KB
constexpr int NSTAGE = 2;
constexpr int OPER_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // hi/lo of both operands: 96 KB
constexpr int RAW_BYTES = 8192;  // Omega-bar only: 4 row groups x 2 halves of [8 rows x 32 r] fp32, 128B-swizzled

enum { MODE_TEST = 0, MODE_ALPHA = 1, MODE_OMEGA = 2, MODE_FWD = 3 };
// tcgen05.mma adds into its fp32 TMEM accumulator with TRUNCATION: a chain of n dependent MMAs shrinks the magnitude of
// the sum by ~0.3 ulp per instruction (measured: 24 000 chained MMAs -> -5e-4 relative in the C3 Omega-bar product,
// 25 000 -> -3e-4 in the M = 512 forward; it would be ~2 % over the 1.2 M MMAs of a C5 Omega-bar tile).  A K loop longer
// than KB_CHAIN K blocks (12 MMAs each: <= 1536 instructions, bias <= ~3e-5 of the chain's partial sum) is therefore
// accumulated on TWO LEVELS: the MMAs of one chain run into TMEM columns [0, 256); when a chain completes, the epilogue
// warps add it, in fp32 registers with round-to-nearest, into a second accumulator in TMEM columns [256, 512)
// (tcgen05.ld / add / tcgen05.st), and only the last chain of an item goes out to memory.  The MMA warp waits for the
// chain accumulator to be drained (~2 % of a chain's time).  Items with a single chain keep the two alternating
// accumulator stages, i.e. the epilogue of one item runs under the MMAs of the next.
constexpr int KB_CHAIN = 128;
// Lockstep of the persistent CTAs.  In the forward and Omega-bar products every concurrently running item streams the
// SAME B operand (the packed W panel / the G^T slab of its gene tile) K block by K block, and only the first CTA to
// touch a block should pay for it in HBM.  Nothing keeps 148 free-running CTAs within an L2's worth of each other over
// a long K loop: at the C5 rank shape (R = 6.4 M rows, 100 000 K blocks) the Omega-bar product ran at 224 TF/s, exactly
// the HBM bound of every CTA fetching its own copy (72 KB per 0.9 us K block x 148 CTAs = 11.8 TB/s wanted), against
// 400 TF/s at R = 128 000.  The TMA producers therefore meet at a grid-wide counter every SYNC_EVERY K blocks (all CTAs
// are co-resident: one per SM); a producer that waits longer than ~2 s stops synchronising instead of hanging.
constexpr int SYNC_EVERY = 128;
constexpr int N_GEN_WARPS = 8;
// modes whose A operand is generated on the fly from TMA-staged raw rows of A (never stored)
__host__ __device__ constexpr bool mode_gen(int mode) { return mode == MODE_OMEGA || mode == MODE_FWD; }
__host__ __device__ constexpr int stage_bytes(int mode) { return OPER_BYTES + (mode_gen(mode) ? RAW_BYTES : 0); }
__host__ __device__ constexpr int gemm_smem(int mode) { return NSTAGE * stage_bytes(mode) + 1024 /*alignment*/ + 256 /*barriers*/; }
// MODE_ALPHA runs TWO epilogue warp groups (warps 4-7 and 8-11, each warp on the TMEM lane quadrant warp % 4), one per
// accumulator stage: its K dimension is the gene count, so a rank of an 8-way gene split has 4 K blocks (3 us of MMAs)
// per tile and one group's contraction epilogue (four 8x8 feature blocks: TMEM loads, 64 L2 loads, 64 reductions per
// thread) took longer than that
__host__ __device__ constexpr int epi_groups(int mode) { return mode == MODE_ALPHA ? 2 : 1; }
__host__ __device__ constexpr int gemm_threads(int mode) {
  return mode_gen(mode) ? 256 + 32 * N_GEN_WARPS : 128 + 128 * epi_groups(mode);
}

struct GemmParams {
  int n_mt, n_nt, group_m, n_split, kblocks, kb_per;
  long Mrows, Ncols;  // logical output extent
  float* C;           // TEST: C [Mrows, Ncols];  OMEGA: H [NF, L];  FWD: q2 [R, L]
  long ldc;
  int accumulate;     // OMEGA/TEST/FWD: 1 = reductions into C (split-K), 0 = plain store
  int batch;          // TEST: > 0 -> operands are 3-D maps [batch, rows, K], item = (b, tile); C advances by sC per batch
  long sC;
  int trans_add;      // TEST: 1 -> C[col*ldc + row] += alpha*acc (transposed, read-modify-write), 0 -> C[row*ldc + col]
  float alpha;
  const float* Amat;  // ALPHA/OMEGA/FWD: A [Mind, R]
  long R;
  int Mind, nb, nblk;
  float* Abar;        // ALPHA: [Mind, R], added to
  unsigned int* sync; // FWD/OMEGA: grid-wide arrival counter (zeroed by the launcher), NULL = no lockstep
};

// arrive at the grid-wide checkpoint whose cumulative arrival count is `expect`, wait for the others (bounded)
__device__ __forceinline__ bool grid_checkpoint(unsigned int* counter, unsigned int expect) {
  atomicAdd(counter, 1u);
  const long long t0 = clock64();
  for (;;) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= expect) return true;
    if (clock64() - t0 > 4000000000LL) return false;  // ~2 s: give up on lockstep, never hang
    __nanosleep(200);
  }
}

__device__ __forceinline__ void decode_item(const GemmParams& p, int item, int& mt, int& nt, int& ks) {
  const int per_split = p.n_mt * p.n_nt;
  if (p.batch > 0) item %= per_split * p.n_split;  // batch index = item / (tiles * splits), see item_batch()
  ks = item / per_split;
  int w = item - ks * per_split;
  const int gfull = p.group_m * p.n_nt;
  const int g = w / gfull;
  const int m0 = g * p.group_m;
  const int gm = min(p.group_m, p.n_mt - m0);
  w -= g * gfull;
  nt = w / gm;
  mt = m0 + w % gm;
}

// four consecutive floats, 16-byte aligned: one fire-and-forget L2 reduction (REDG.ADD.F32x4, round to nearest)
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ int item_batch(const GemmParams& p, int item) {
  return p.batch > 0 ? item / (p.n_mt * p.n_nt * p.n_split) : 0;
}

// 32 accumulator columns of this thread's TMEM lane: the chain accumulator, plus the second-level accumulator when
// `prev` (see KB_CHAIN)
__device__ __forceinline__ void ld_acc32(uint32_t t_chain, uint32_t t_acc2, bool prev, uint32_t* v) {
  tmem_ld32(t_chain, v);
  if (prev) {
    uint32_t w[32];
    tmem_ld32(t_acc2, w);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
  } else {
    tmem_ld_wait();
  }
}

// -------------------------------------------------------------------------------------------------
// generic 128 x 256 x K tile GEMM, C = (A_hi + A_lo)(B_hi + B_lo)^T in three bf16 passes
// -------------------------------------------------------------------------------------------------
// In MODE_OMEGA / MODE_FWD tmA_hi is the fp32 map of A [M, R] (box 32 r x 8 rows) the generators read from; tmA_lo is
// unused.
//
// MODE_FWD is the forward quadratic form as the same implicit-feature GEMM the two backward products use:
//     q2[r, p] = sum_{(i,j)} Phi[r,(i,j)] W[(i,j), p],   Phi[r,(i,j)] = a_r[i] a_r[j],  W = c_b Omega_p[i,j]
// (reference gpsa/models/vgpsa.py:193-196; the Cholesky factor of Omega is not needed for it).
// Output tile = 128 rows r x 256 genes, K runs over the nblk 8x8 feature blocks (I, J) in feat.cu order.  Per K block
// the producer stages the two 8-row groups a[I*8.., r0..r0+127] and a[J*8.., ...] (8 KB, fp32) and the 256 x 64 block of
// the packed W^T (hi, lo); the generator warps (thread = row r, two threads per row) multiply, split to bf16 (hi, lo)
// and write the swizzled K-major A operand.  Measured against fp64 the three-pass feature form is at least as accurate
// as the reference's own ||a^T L||^2 form evaluated the same way and runs at the issue rate of the plain
// GEMM core.
template <int MODE>
__global__ void __launch_bounds__(gemm_threads(MODE), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const GemmParams p) {
  constexpr int STAGE_BYTES = stage_bytes(MODE);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* full = bars;            // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;  // [NSTAGE]
  uint64_t* tfull = bars + 2 * NSTAGE;      // [2]
  uint64_t* tempty = bars + 2 * NSTAGE + 2; // [2]
  uint64_t* rawfull = bars + 2 * NSTAGE + 4; // [NSTAGE]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NSTAGE + 4);

  // warp index through a shuffle: provably warp-uniform, so the role branches below are convergent for the compiler
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA_hi);
    if (!mode_gen(MODE)) prefetch_tmap(&tmA_lo);
    prefetch_tmap(&tmB_hi);
    prefetch_tmap(&tmB_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], mode_gen(MODE) ? 1 + N_GEN_WARPS : 1);  // TMA expect_tx arrive (+ generator warps)
      mbar_init(&empty[s], 1);
      mbar_init(&rawfull[s], 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.n_mt * p.n_nt * p.n_split * (p.batch > 0 ? p.batch : 1);
  // two-level accumulation (see KB_CHAIN): every item of this launch has the same chain structure
  const bool two_level = p.kb_per > KB_CHAIN;

  // Producer and MMA warps run converged (all lanes wait on the barriers, loop state is warp-uniform); only the
  // TMA / MMA issue is predicated on elect.sync, which keeps the single issuing thread's instruction stream short.
  if (warp == 0) {
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    bool lockstep = p.sync != nullptr && p.n_split == 1 && n_items > 1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int mt, nt, ks;
      decode_item(p, item, mt, nt, ks);
      const int kb0 = ks * p.kb_per;
      const int kb1 = min(p.kblocks, kb0 + p.kb_per);
      int grow[4] = {0, 0, 0, 0};  // MODE_OMEGA: first A row of the (I, J) index groups of the tile's two feature blocks
      if (MODE == MODE_OMEGA) {
        for (int q = 0; q < 2; ++q) {
          const int b = mt * (TM / FBK) + q;
          int I = p.nb, J = p.nb;  // out of range -> rows >= M -> zero filled
          if (b < p.nblk) decode_block(b, p.nb, I, J);
          grow[2 * q] = I * FB;
          grow[2 * q + 1] = J * FB;
        }
      }
      int fI = 0, fJ = 0;  // MODE_FWD: feature block (I, J) of K block kb (feat.cu order: J runs from I to nb - 1)
      if (MODE == MODE_FWD && kb0 > 0) decode_block(kb0, p.nb, fI, fJ);
      // lockstep checkpoints (n_split == 1: every item has the same K range): wave w of the persistent loop has
      // part = min(grid, items left) participants and ncp checkpoints
      const unsigned int wave = (unsigned int)((item - (int)blockIdx.x) / (int)gridDim.x);
      const unsigned int part = (unsigned int)min((int)gridDim.x, n_items - (int)(wave * gridDim.x));
      const unsigned int ncp = (unsigned int)((kb1 - kb0 + SYNC_EVERY - 1) / SYNC_EVERY);
      for (int kb = kb0; kb < kb1; ++kb) {
        if (mode_gen(MODE) && lockstep && (kb - kb0) % SYNC_EVERY == 0) {
          const unsigned int cp = (unsigned int)((kb - kb0) / SYNC_EVERY);
          unsigned int ok = 1;
          if (elect_one()) ok = grid_checkpoint(p.sync, ncp * wave * gridDim.x + (cp + 1) * part) ? 1u : 0u;
          if (!__all_sync(0xffffffffu, ok != 0)) lockstep = false;
        }
        uint8_t* st = smem + stage * STAGE_BYTES;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], mode_gen(MODE) ? 2 * B_TILE_BYTES : OPER_BYTES);
          if (MODE == MODE_FWD) {
            // raw rows of A for this feature block: group 0 = rows I*8.., group 1 = rows J*8.., four 32-r boxes each
            mbar_arrive_expect_tx(&rawfull[stage], RAW_BYTES);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tma_load_2d(st + OPER_BYTES + c * 1024, &tmA_hi, &rawfull[stage], mt * TM + c * 32, fI * FB);
              tma_load_2d(st + OPER_BYTES + (4 + c) * 1024, &tmA_hi, &rawfull[stage], mt * TM + c * 32, fJ * FB);
            }
          } else if (MODE == MODE_OMEGA) {
            mbar_arrive_expect_tx(&rawfull[stage], RAW_BYTES);
#pragma unroll
            for (int gq = 0; gq < 4; ++gq)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                tma_load_2d(st + OPER_BYTES + (gq * 2 + h) * 1024, &tmA_hi, &rawfull[stage], kb * BK + h * 32, grow[gq]);
          } else if (p.batch > 0) {
            const int bz = item_batch(p, item);
            tma_load_3d(st, &tmA_hi, &full[stage], kb * BK, mt * TM, bz);
            tma_load_3d(st + A_TILE_BYTES, &tmA_lo, &full[stage], kb * BK, mt * TM, bz);
            tma_load_3d(st + 2 * A_TILE_BYTES, &tmB_hi, &full[stage], kb * BK, nt * TN, bz);
            tma_load_3d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &tmB_lo, &full[stage], kb * BK, nt * TN, bz);
          } else {
            tma_load_2d(st, &tmA_hi, &full[stage], kb * BK, mt * TM);
            tma_load_2d(st + A_TILE_BYTES, &tmA_lo, &full[stage], kb * BK, mt * TM);
          }
          if (MODE == MODE_OMEGA) {  // G^T is stored K-block-major (pack_Gt_kernel): the tile is one contiguous piece
            tma_load_3d(st + 2 * A_TILE_BYTES, &tmB_hi, &full[stage], 0, nt * TN, kb);
            tma_load_3d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &tmB_lo, &full[stage], 0, nt * TN, kb);
          } else if (!(MODE == MODE_TEST && p.batch > 0)) {
            tma_load_2d(st + 2 * A_TILE_BYTES, &tmB_hi, &full[stage], kb * BK, nt * TN);
            tma_load_2d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &tmB_lo, &full[stage], kb * BK, nt * TN);
          }
        }
        __syncwarp();
        if (MODE == MODE_FWD && ++fJ == p.nb) { ++fI; fJ = fI; }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_bf16(TM, TN);
    const uint64_t desc0 = make_desc_sw128(smem_u32(smem));
    int stage = 0, acc = 0;
    uint32_t phase = 0;
    uint32_t acc_ph[2] = {0, 0};  // completed uses of each accumulator stage, mod 2
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int mt, nt, ks;
      decode_item(p, item, mt, nt, ks);
      const int kb0 = ks * p.kb_per;
      const int kb1 = min(p.kblocks, kb0 + p.kb_per);
      // chains of at most KB_CHAIN K blocks; one-level items are a single chain on the alternating stage `acc`
      for (int c0 = kb0; c0 < kb1; c0 += (two_level ? KB_CHAIN : p.kb_per)) {
        const int c1 = two_level ? min(kb1, c0 + KB_CHAIN) : kb1;
        mbar_wait(&tempty[acc], acc_ph[acc] ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * TN;
        for (int kb = c0; kb < c1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_hi = desc0 + (uint64_t)((uint32_t)(stage * STAGE_BYTES) >> 4);
            const uint64_t a_lo = a_hi + (A_TILE_BYTES >> 4);
            const uint64_t b_hi = a_hi + (2 * A_TILE_BYTES >> 4);
            const uint64_t b_lo = b_hi + (B_TILE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t off = (k * UMMA_K * 2) >> 4;
              umma_bf16(d, a_hi + off, b_hi + off, idesc, (kb > c0 || k > 0) ? 1u : 0u);
              umma_bf16(d, a_lo + off, b_hi + off, idesc, 1u);
              umma_bf16(d, a_hi + off, b_lo + off, idesc, 1u);
            }
            umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull[acc]);
        __syncwarp();
        acc_ph[acc] ^= 1;
        if (!two_level) acc ^= 1;  // two-level: the chain accumulator is always stage 0, stage 1 is the second level
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * epi_groups(MODE)) {
    // ===== epilogue =====
    // NGRP warp groups (MODE_ALPHA: 2).  One-level items alternate between the two accumulator stages, and group g owns
    // stage g: it drains every NGRP-th item of this CTA, so one group's epilogue has two tile times to finish.  Two-level
    // items (chains through stage 0 into the second-level accumulator) are all drained by group 0.
    constexpr int NGRP = epi_groups(MODE);
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    int acc = (NGRP > 1 && !two_level) ? grp : 0;
    uint32_t acc_ph[2] = {0, 0};
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int local = 0;  // index of the item among this CTA's items
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local) {
      if (NGRP > 1 && (two_level ? grp != 0 : (local % NGRP) != grp)) continue;  // another group's item
      int mt, nt, ks;
      decode_item(p, item, mt, nt, ks);
      const int kb0 = ks * p.kb_per;
      const int kb1 = min(p.kblocks, kb0 + p.kb_per);
      bool prev = false;  // the second-level accumulator holds the sum of this item's earlier chains
      if (two_level) {
        // every chain but the last: second level += chain (fp32 registers, round to nearest), release the chain stage
        for (int c0 = kb0; c0 + KB_CHAIN < kb1; c0 += KB_CHAIN) {
          mbar_wait(&tfull[0], acc_ph[0]);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < TN / 32; ++c) {
            uint32_t v[32];
            ld_acc32(lane_base + c * 32, lane_base + TN + c * 32, prev, v);
            tmem_st32(lane_base + TN + c * 32, v);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[0]);
          acc_ph[0] ^= 1;
          prev = true;
        }
      }
      const long row = (long)mt * TM + q * 32 + lane;
      // MODE_ALPHA: this thread's a_r[i], a_r[j] for the tile's feature blocks (I_u, J_u), fetched while the tile's MMAs
      // still run (L2 round trips under the operand stream are ~1 us: behind the TMEM wait they were the epilogue).
      // Consecutive blocks mostly share I (feat.cu order: J runs from I to nb - 1): a_r[I-rows] is fetched once per run.
      constexpr int NBLK_EPI = MODE == MODE_ALPHA ? TN / FBK : 1;
      float aI[NBLK_EPI][FB], aJ[NBLK_EPI][FB];
      int bI[NBLK_EPI], bJ[NBLK_EPI];
      if (MODE == MODE_ALPHA) {
        const bool valid = row < p.R;
#pragma unroll
        for (int u = 0; u < NBLK_EPI; ++u) {
          const int b = nt * (TN / FBK) + u;
          bI[u] = -1;
          bJ[u] = -1;
          if (b < p.nblk) decode_block(b, p.nb, bI[u], bJ[u]);
          const bool same_I = u > 0 && bI[u] == bI[u > 0 ? u - 1 : 0];  // warp-uniform
#pragma unroll
          for (int t = 0; t < FB; ++t) {
            const int mi = bI[u] * FB + t, mj = bJ[u] * FB + t;
            if (same_I) aI[u][t] = aI[u > 0 ? u - 1 : 0][t];
            else aI[u][t] = (valid && bI[u] >= 0 && mi < p.Mind) ? __ldg(&p.Amat[(long)mi * p.R + row]) : 0.f;
            aJ[u][t] = (valid && bI[u] >= 0 && mj < p.Mind) ? __ldg(&p.Amat[(long)mj * p.R + row]) : 0.f;
          }
        }
      }
      mbar_wait(&tfull[acc], acc_ph[acc]);
      tc_fence_after();
      const uint32_t taddr = lane_base + (uint32_t)acc * TN;  // two-level: acc == 0
      const uint32_t taddr2 = lane_base + TN;
      if (MODE == MODE_TEST || MODE == MODE_OMEGA || MODE == MODE_FWD) {
        float* Cb = p.C + (MODE == MODE_TEST ? (long)item_batch(p, item) * p.sC : 0);
        const float alpha = MODE == MODE_TEST ? p.alpha : 1.f;
        const bool al4 = !p.trans_add && (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(Cb) & 15) == 0);
        const bool vec4 = al4 && !p.accumulate;
        const bool red4 = al4 && p.accumulate;
#pragma unroll 1
        for (int c = 0; c < TN / 32; ++c) {
          const long col0 = (long)nt * TN + c * 32;
          if (col0 >= p.Ncols) break;  // warp-uniform
          uint32_t v[32];
          ld_acc32(taddr + c * 32, taddr2 + c * 32, prev, v);
          if (row < p.Mrows) {
            if (MODE == MODE_TEST && p.trans_add) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const long col = col0 + j;
                // lanes = consecutive rows: coalesced.  A reduction (RED.ADD, fire and forget) instead of load-add-store:
                // every element has exactly one writer, so the result is the same, but 200 dependent global round
                // trips per thread (the compiler cannot hoist the loads over the stores) made this epilogue cost
                // ~95 us per tile
                if (col < p.Ncols) atomicAdd(&Cb[col * p.ldc + row], alpha * __uint_as_float(v[j]));
              }
            } else if (vec4 && col0 + 32 <= p.Ncols) {
              float4* dst = reinterpret_cast<float4*>(Cb + row * p.ldc + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                dst[j] = make_float4(alpha * __uint_as_float(v[4 * j]), alpha * __uint_as_float(v[4 * j + 1]),
                                     alpha * __uint_as_float(v[4 * j + 2]), alpha * __uint_as_float(v[4 * j + 3]));
            } else if (red4 && col0 + 32 <= p.Ncols) {
              float* dst = Cb + row * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                red_add_v4(dst + 4 * j, alpha * __uint_as_float(v[4 * j]), alpha * __uint_as_float(v[4 * j + 1]),
                           alpha * __uint_as_float(v[4 * j + 2]), alpha * __uint_as_float(v[4 * j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const long col = col0 + j;
                if (col < p.Ncols) {
                  float* dst = Cb + row * p.ldc + col;
                  if (p.accumulate) atomicAdd(dst, alpha * __uint_as_float(v[j]));
                  else *dst = alpha * __uint_as_float(v[j]);
                }
              }
            }
          }
        }
      } else {  // MODE_ALPHA: Abar[:, r] += contraction of Psi[r, (i,j)] with a_r
        // Abar[I-rows] of a run of blocks with the same I is summed in registers and reduced to memory once per run
        // (the L2 reduction units, not the MMAs, bounded this product at small gene counts): 8 + 8 reductions per block
        // become 8 per block + 8 per run
        const bool valid = row < p.R;
        float sI[FB];
#pragma unroll
        for (int t = 0; t < FB; ++t) sI[t] = 0.f;
#pragma unroll
        for (int u = 0; u < NBLK_EPI; ++u) {
          const int I = bI[u], J = bJ[u];
          if (I < 0) break;  // warp-uniform: past the last feature block
          float sJ[FB];
#pragma unroll
          for (int t = 0; t < FB; ++t) sJ[t] = 0.f;
          const bool diag = (I == J);
#pragma unroll
          for (int half = 0; half < 2; ++half) {  // 4 feature rows il = 32 accumulator columns at a time
            uint32_t v[32];
            ld_acc32(taddr + u * FBK + half * 32, taddr2 + u * FBK + half * 32, prev, v);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int il = half * 4 + i4;
              float si = 0.f;
#pragma unroll
              for (int jl = 0; jl < FB; ++jl) {
                const float psi = __uint_as_float(v[i4 * FB + jl]);
                si = fmaf(psi, aJ[u][jl], si);
                sJ[jl] = fmaf(psi, aI[u][il], sJ[jl]);
              }
              sI[il] += diag ? 2.f * si : si;
            }
          }
          if (!diag) {
#pragma unroll
            for (int jl = 0; jl < FB; ++jl) {
              const int mj = J * FB + jl;
              if (valid && mj < p.Mind) atomicAdd(&p.Abar[(long)mj * p.R + row], sJ[jl]);
            }
          }
          const bool last_of_run = (u + 1 == NBLK_EPI) || bI[u + 1 < NBLK_EPI ? u + 1 : u] != I;  // warp-uniform
          if (last_of_run) {
#pragma unroll
            for (int il = 0; il < FB; ++il) {
              const int mi = I * FB + il;
              if (valid && mi < p.Mind) atomicAdd(&p.Abar[(long)mi * p.R + row], sI[il]);
              sI[il] = 0.f;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc_ph[acc] ^= 1;
      if (NGRP == 1 && !two_level) acc ^= 1;
    }
  } else if (MODE == MODE_OMEGA && warp >= 8) {
    // ===== feature-operand generators =====
    // A-operand row t = feature (i, j) of the tile, columns = 64 consecutive r:  phi = a_i[r] a_j[r], split to bf16
    // (hi, lo) and written in the swizzled K-major layout.  The raw a rows arrive by TMA (128B-swizzled
    // [8 rows x 32 r] fp32 boxes, so the 8 rows a quarter-warp reads at one column land in 8 different bank groups).
    const int gt = threadIdx.x - 256;
    const int t = gt & (TM - 1);   // feature row of the tile
    const int h = gt >> 7;         // which half of the 64 r of a K block
    const int bq = t / FBK, il = (t % FBK) / FB, jl = t % FB;
    const uint32_t raw_i = OPER_BYTES + ((2 * bq) * 2 + h) * 1024 + il * 128;
    const uint32_t raw_j = OPER_BYTES + ((2 * bq + 1) * 2 + h) * 1024 + jl * 128;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int mt, nt, ks;
      decode_item(p, item, mt, nt, ks);
      const int kb0 = ks * p.kb_per, kb1 = min(p.kblocks, kb0 + p.kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        uint8_t* st = smem + stage * STAGE_BYTES;
        mbar_wait(&rawfull[stage], phase);  // armed only after the MMA released this stage
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // output chunk 4h + c = 8 consecutive r = raw chunks 2c, 2c+1 of this half
          const float4 u0 = *reinterpret_cast<const float4*>(st + raw_i + (((2 * c) ^ il) << 4));
          const float4 u1 = *reinterpret_cast<const float4*>(st + raw_i + (((2 * c + 1) ^ il) << 4));
          const float4 w0 = *reinterpret_cast<const float4*>(st + raw_j + (((2 * c) ^ jl) << 4));
          const float4 w1 = *reinterpret_cast<const float4*>(st + raw_j + (((2 * c + 1) ^ jl) << 4));
          uint4 hi, lo;
          split_pair(u0.x * w0.x, u0.y * w0.y, hi.x, lo.x);
          split_pair(u0.z * w0.z, u0.w * w0.w, hi.y, lo.y);
          split_pair(u1.x * w1.x, u1.y * w1.y, hi.z, lo.z);
          split_pair(u1.z * w1.z, u1.w * w1.w, hi.w, lo.w);
          const uint32_t off = sw128_offset(t, 4 * h + c);
          *reinterpret_cast<uint4*>(st + off) = hi;
          *reinterpret_cast<uint4*>(st + A_TILE_BYTES + off) = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (MODE == MODE_FWD && warp >= 8) {
    // ===== feature-operand generators, forward =====
    // A-operand row t = row r of the tile, its 64 K columns = features (il, jl) of block (I, J): phi = a_I[il][r] a_J[jl][r].
    // Thread (t, h) writes the four 16-byte chunks il = 4h .. 4h+3 (chunk il = the eight jl of one il) of row t.
    // Raw layout: box c = t / 32 of group g at OPER_BYTES + (4g + c) KB, 8 rows (m) x 32 r fp32, 128B-swizzled: lanes of
    // a warp read consecutive r of ONE row m -> 32 distinct banks.
    const int gt = threadIdx.x - 256;
    const int t = gt & (TM - 1);
    const int h = gt >> 7;
    const int rc = t & 31;
    const uint32_t boxI = OPER_BYTES + (uint32_t)(t >> 5) * 1024, boxJ = boxI + 4096;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int mt, nt, ks;
      decode_item(p, item, mt, nt, ks);
      const int kb0 = ks * p.kb_per, kb1 = min(p.kblocks, kb0 + p.kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        uint8_t* st = smem + stage * STAGE_BYTES;
        mbar_wait(&rawfull[stage], phase);  // armed only after the MMA released this stage
        float aJ[FB];
#pragma unroll
        for (int jl = 0; jl < FB; ++jl)
          aJ[jl] = *reinterpret_cast<const float*>(st + boxJ + jl * 128 + ((((rc >> 2) ^ jl) << 4) | ((rc & 3) << 2)));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int il = 4 * h + c;
          const float ai = *reinterpret_cast<const float*>(st + boxI + il * 128 + ((((rc >> 2) ^ il) << 4) | ((rc & 3) << 2)));
          uint4 hi, lo;
          split_pair(ai * aJ[0], ai * aJ[1], hi.x, lo.x);
          split_pair(ai * aJ[2], ai * aJ[3], hi.y, lo.y);
          split_pair(ai * aJ[4], ai * aJ[5], hi.z, lo.z);
          split_pair(ai * aJ[6], ai * aJ[7], hi.w, lo.w);
          const uint32_t off = sw128_offset(t, il);
          *reinterpret_cast<uint4*>(st + off) = hi;
          *reinterpret_cast<uint4*>(st + A_TILE_BYTES + off) = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// operand packing (HBM-bound pre-passes): fp32 -> bf16 (hi, lo), K-major, zero padded
// -------------------------------------------------------------------------------------------------
// Wt[(b, il, jl), p] = c_b Omega[p, i, j]   (c = 1 on diagonal blocks, 2 above; same ordering as feat.cu)
__global__ void __launch_bounds__(256) pack_Wt_kernel(int M, int L, int Lp, const float* __restrict__ Omega,
                                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
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
      __nv_bfloat16 h, l;
      split_one(v, h, l);
      const long o = ((long)b * FBK + f) * Lp + p;
      hi[o] = h;
      lo[o] = l;
    }
  }
}

// Wg[p, (b, il, jl)] = c_b Omega[p, i, j]: the same packed symmetric features, gene-major / feature-contiguous (the
// K-major B operand of the forward feature GEMM).  One CTA per gene; for a fixed block row I the blocks (I, I..nb-1)
// are contiguous in the feature order, one thread per (J, il) writes the eight jl of its chunk as one 16-byte store.
__global__ void __launch_bounds__(256) pack_Wg_kernel(int M, long NF, const float* __restrict__ Omega,
                                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int nb = feat_nb(M);
  const int p = blockIdx.x;
  const float* Om = Omega + (long)p * M * M;
  uint4* oh = reinterpret_cast<uint4*>(hi + (long)p * NF);
  uint4* ol = reinterpret_cast<uint4*>(lo + (long)p * NF);
  long chunk0 = 0;  // 16-byte chunk index of block (I, I)
  for (int I = 0; I < nb; ++I) {
    const int nch = (nb - I) * FB;  // chunks of this block row: (J - I) * 8 + il
    for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
      const int J = I + ch / FB, il = ch % FB;
      const int i = I * FB + il, j0 = J * FB;
      const float c = (I == J) ? 1.f : 2.f;
      float v[FB];
#pragma unroll
      for (int jl = 0; jl < FB; ++jl) v[jl] = (i < M && j0 + jl < M) ? c * Om[(long)i * M + j0 + jl] : 0.f;
      uint4 h, l;
      split_pair(v[0], v[1], h.x, l.x);
      split_pair(v[2], v[3], h.y, l.y);
      split_pair(v[4], v[5], h.z, l.z);
      split_pair(v[6], v[7], h.w, l.w);
      oh[chunk0 + ch] = h;
      ol[chunk0 + ch] = l;
    }
    chunk0 += nch;
  }
}

// row-major split: out[r, p] = G[r, p], pitch Lp
__global__ void pack_G_kernel(long R, int L, int Lp, const float* __restrict__ G, __nv_bfloat16* __restrict__ hi,
                              __nv_bfloat16* __restrict__ lo) {
  const long total = R * L;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / L;
    const int pp = (int)(idx - r * L);
    __nv_bfloat16 h, l;
    split_one(G[idx], h, l);
    hi[r * Lp + pp] = h;
    lo[r * Lp + pp] = l;
  }
}

// transposed split in K-BLOCK-MAJOR order: out[r / 64][p][r % 64] = G[r, p]  (rows r >= R of the last block are zero).
// The B operand tile of the Omega-bar product -- 256 genes x 64 consecutive r -- is then ONE contiguous 32 KB piece of
// memory.  With the plain [L, R] transpose the 256 rows of a tile were R * 2 bytes apart: at the C5 rank shape
// (R = 6.4 M) every 128-byte row of every TMA box sat in its own 2 MB page and the product ran at 224 TF/s against
// 403 TF/s at R = 128 000 (the rate fell monotonically with the row stride: 256 KB / 1.3 MB / 1.6 MB / 12.8 MB ->
// 403 / 347 / 318 / 224 TF/s).
__global__ void __launch_bounds__(256) pack_Gt_kernel(long R, int L, const float* __restrict__ G,
                                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  // one 64 r x 64 gene tile per CTA (block 32 x 8): 128-byte reads along the genes, and every (gene, K block) row of the
  // output -- 64 consecutive r = 128 bytes of bf16 -- written by one warp as bf16 pairs
  __shared__ float t[64][65];
  const long r0 = (long)blockIdx.x * 64;
  const int p0 = blockIdx.y * 64;
  for (int yy = threadIdx.y; yy < 64; yy += 8) {
    const long r = r0 + yy;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pp = p0 + threadIdx.x + 32 * h;
      t[yy][threadIdx.x + 32 * h] = (r < R && pp < L) ? G[r * L + pp] : 0.f;  // zero beyond R: the tail of the last block
    }
  }
  __syncthreads();
  for (int yy = threadIdx.y; yy < 64; yy += 8) {
    const int pp = p0 + yy;
    if (pp < L) {
      uint32_t h, l;
      split_pair(t[2 * threadIdx.x][yy], t[2 * threadIdx.x + 1][yy], h, l);
      const long o = ((long)blockIdx.x * L + pp) * 64 + 2 * threadIdx.x;
      *reinterpret_cast<uint32_t*>(hi + o) = h;
      *reinterpret_cast<uint32_t*>(lo + o) = l;
    }
  }
}

// generic batched packs for gpsa_gemm_tc: out[b, row, k] (pitch Kp, bf16 hi/lo) from a row-major fp32 operand
//   rows_major = 1: in[b*sIn + row*ld + k]          (operand already K-contiguous)
//   rows_major = 0: in[b*sIn + k*ld + row]          (operand stored K x rows: transposed on the fly)
__global__ void pack_rows_kernel(long rows, int K, int Kp, long ld, long sIn, const float* __restrict__ in,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int b = blockIdx.z;
  const float* src = in + (long)b * sIn;
  const long total = rows * K, obase = (long)b * rows * Kp;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / K;
    const int k = (int)(idx - r * K);
    __nv_bfloat16 h, l;
    split_one(src[r * ld + k], h, l);
    hi[obase + r * Kp + k] = h;
    lo[obase + r * Kp + k] = l;
  }
}
// 128-bit variant of pack_rows_kernel / pack_G_kernel (K % 4 == 0, ld % 4 == 0, 16-byte aligned source): one warp per
// row, four elements per lane and access, no per-element 64-bit division
__global__ void __launch_bounds__(256) pack_rows4_kernel(long rows, int K4, int Kp, long ld, long sIn,
                                                         const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo) {
  const int b = blockIdx.z;
  const float* src = in + (long)b * sIn;
  const long obase = (long)b * rows * Kp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  for (long r = (long)blockIdx.x * (blockDim.x >> 5) + warp; r < rows; r += nwarps) {
    const float4* s4 = reinterpret_cast<const float4*>(src + r * ld);
    uint2* h2 = reinterpret_cast<uint2*>(hi + obase + r * Kp);
    uint2* l2 = reinterpret_cast<uint2*>(lo + obase + r * Kp);
    for (int k = lane; k < K4; k += 32) {
      const float4 v = s4[k];
      uint2 h, l;
      split_pair(v.x, v.y, h.x, l.x);
      split_pair(v.z, v.w, h.y, l.y);
      h2[k] = h;
      l2[k] = l;
    }
  }
}
// true if the 128-bit pack applies
static inline bool pack4_ok(const float* src, long ld, long sIn, int K, int Kp) {
  return K >= 128 && (K & 3) == 0 && (Kp & 3) == 0 && (ld & 3) == 0 && (sIn & 3) == 0 &&
         (reinterpret_cast<uintptr_t>(src) & 15) == 0;
}

__global__ void __launch_bounds__(256) pack_trans_kernel(long rows, int K, int Kp, long ld, long sIn,
                                                         const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo) {
  // one 64 row x 64 k tile per CTA (block 32 x 8): 128-byte reads along the rows of the K x rows source, 128-byte writes
  // (bf16 pairs) along k
  __shared__ float t[64][65];
  const int b = blockIdx.z;
  const float* src = in + (long)b * sIn;
  const long obase = (long)b * rows * Kp;
  // flattened (row tile, K tile) index on x: either extent can exceed the 65535 limit of grid.y (R-sized operands)
  const unsigned nkt = (unsigned)((K + 63) / 64);
  const long r0 = (long)(blockIdx.x / nkt) * 64;
  const int k0 = (int)(blockIdx.x % nkt) * 64;
  for (int yy = threadIdx.y; yy < 64; yy += 8) {
    const int k = k0 + yy;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long r = r0 + threadIdx.x + 32 * h;
      t[yy][threadIdx.x + 32 * h] = (k < K && r < rows) ? src[(long)k * ld + r] : 0.f;
    }
  }
  __syncthreads();
  for (int yy = threadIdx.y; yy < 64; yy += 8) {
    const long r = r0 + yy;
    const int k = k0 + 2 * threadIdx.x;
    if (r < rows && k < K) {
      const long o = obase + r * Kp + k;  // even: Kp is a multiple of 8
      if (k + 1 < K) {
        uint32_t h, l;
        split_pair(t[2 * threadIdx.x][yy], t[2 * threadIdx.x + 1][yy], h, l);
        *reinterpret_cast<uint32_t*>(hi + o) = h;
        *reinterpret_cast<uint32_t*>(lo + o) = l;
      } else {
        __nv_bfloat16 h, l;
        split_one(t[2 * threadIdx.x][yy], h, l);
        hi[o] = h;
        lo[o] = l;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// bf16 tensor, innermost dimension contiguous, 128-byte swizzle, out-of-bounds elements read as zero
int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
              CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return GPSA_ERR_CUDA;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], es[3];
  for (int d = 0; d < rank; ++d) { gdim[d] = dims[d]; bx[d] = box[d]; es[d] = 1; }
  for (int d = 0; d + 1 < rank; ++d) gstr[d] = strides_bytes[d];
  const CUresult rc = fn(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? GPSA_OK : GPSA_ERR_CUDA;
}
int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_rows) {
  const uint64_t dims[2] = {inner, outer}, str[1] = {pitch_elems * 2};
  const uint32_t box[2] = {(uint32_t)BK, box_rows};
  return make_tmap(tm, base, 2, dims, str, box);
}

int sm_count() {
  static int n[GPSA_MAX_DEVICES] = {};
  const int dev = gpsa_dev();
  if (n[dev] == 0) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v;
  }
  return n[dev];
}

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
inline long rup(long x, long m) { return (x + m - 1) / m * m; }

struct FeatFwdLayout { long NF; size_t wg, apad, total; };
FeatFwdLayout featfwd_layout(int M, long R, int L) {
  FeatFwdLayout f;
  f.NF = feat_nblk(M) * FBK;
  f.wg = al256((size_t)L * f.NF * 2);
  f.apad = (R % 4 == 0) ? 0 : al256((size_t)M * rup(R, 4) * 4);  // TMA needs a 16-byte row pitch: padded copy of A
  f.total = 2 * f.wg + f.apad + 256;  // + the lockstep counter
  return f;
}
struct AlphaLayout { int Lp; long NF; size_t g, w, total; };
AlphaLayout alpha_layout(int M, long R, int L) {
  AlphaLayout a;
  a.Lp = (int)rup(L, 8);
  a.NF = feat_nblk(M) * FBK;
  a.g = al256((size_t)R * a.Lp * 2);
  a.w = al256((size_t)a.NF * a.Lp * 2);
  a.total = 2 * a.g + 2 * a.w;
  return a;
}
struct OmegaLayout { long Rp; size_t gt, apad, total; };
OmegaLayout omega_layout(int M, long R, int L) {
  OmegaLayout o;
  o.Rp = rup(R, BK);  // whole K blocks (pack_Gt_kernel zero-fills the tail)
  o.gt = al256((size_t)L * o.Rp * 2);
  o.apad = (R % 4 == 0) ? 0 : al256((size_t)M * rup(R, 4) * 4);  // TMA needs a 16-byte row pitch: padded copy of A
  o.total = 2 * o.gt + o.apad + 256;  // + the lockstep counter
  return o;
}

template <int MODE>
int launch_gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                const GemmParams& p, cudaStream_t st) {
  static bool attr_set[GPSA_MAX_DEVICES] = {};
  const int dev = gpsa_dev();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(MODE)) !=
        cudaSuccess)
      return GPSA_ERR_CUDA;
    attr_set[dev] = true;
  }
  const int n_items = p.n_mt * p.n_nt * p.n_split * (p.batch > 0 ? p.batch : 1);
  const int grid = n_items < sm_count() ? n_items : sm_count();
  tc_gemm_kernel<MODE><<<grid, gemm_threads(MODE), gemm_smem(MODE), st>>>(a_hi, a_lo, b_hi, b_lo, p);
  GPSA_LAUNCH_CHECK();
  return GPSA_OK;
}

void set_split(GemmParams& p, int want_split) {
  if (want_split < 1) want_split = 1;
  if (want_split > p.kblocks) want_split = p.kblocks;
  p.kb_per = (p.kblocks + want_split - 1) / want_split;
  p.n_split = (p.kblocks + p.kb_per - 1) / p.kb_per;
}

}  // namespace

// =================================================================================================
// exported entry points
// =================================================================================================
extern "C" int gpsa_tc_supported(int M) { return (M >= 16 && M <= 512) ? 1 : 0; }

extern "C" size_t gpsa_gemm_tc_ws_bytes(long Mr, long Nc, int K, int batch);

// scratch for everything the data layer runs on the tcgen05 engine at this shape: the three quadratic-form
// products and the four plain GEMMs (predictive mean, delta-bar, A-bar += delta Fbar^T, Omega-bar Omega_sqt)
extern "C" size_t gpsa_quadform_tc_ws_bytes(int M, long R, int L) {
  if (M <= 0 || R <= 0 || L <= 0) return 0;
  size_t m = featfwd_layout(M, R, L).total;
  const size_t c[] = {alpha_layout(M, R, L).total, omega_layout(M, R, L).total,
                      gpsa_gemm_tc_ws_bytes(R, L, M, 1),
                      gpsa_gemm_tc_ws_bytes(M, L, (int)(R > 2000000000L ? 2000000000L : R), 1), gpsa_gemm_tc_ws_bytes(R, M, L, 1),
                      gpsa_gemm_tc_ws_bytes(M, M, M, L)};
  for (size_t v : c) m = v > m ? v : m;
  return m + 256;
}

// C [Mr, Nc] = A [Mr, K] B[Nc, K]^T with A, B fp32 row-major: split to bf16 in `ws`, then the tcgen05 GEMM core.
// Exists so that the descriptor / TMA / TMEM plumbing can be unit-tested in isolation.
extern "C" int gpsa_tc_gemm_test(int Mr, int Nc, int K, const float* A, const float* B, float* C, int split, void* ws,
                                 size_t ws_bytes, cudaStream_t st) {
  if (Mr <= 0 || Nc <= 0 || K <= 0) return GPSA_ERR_ARG;
  const int Kp = (int)rup(K, 8);
  const size_t sa = al256((size_t)Mr * Kp * 2), sb = al256((size_t)Nc * Kp * 2);
  if (ws_bytes < 2 * sa + 2 * sb) return GPSA_ERR_ARG;
  uint8_t* w = static_cast<uint8_t*>(ws);
  __nv_bfloat16 *a_hi = (__nv_bfloat16*)w, *a_lo = (__nv_bfloat16*)(w + sa), *b_hi = (__nv_bfloat16*)(w + 2 * sa),
                *b_lo = (__nv_bfloat16*)(w + 2 * sa + sb);
  pack_G_kernel<<<148 * 4, 256, 0, st>>>(Mr, K, Kp, A, a_hi, a_lo);
  GPSA_LAUNCH_CHECK();
  pack_G_kernel<<<148 * 4, 256, 0, st>>>(Nc, K, Kp, B, b_hi, b_lo);
  GPSA_LAUNCH_CHECK();
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if (make_tmap_2d(&ta_hi, a_hi, K, Mr, Kp, TM) || make_tmap_2d(&ta_lo, a_lo, K, Mr, Kp, TM) ||
      make_tmap_2d(&tb_hi, b_hi, K, Nc, Kp, TN) || make_tmap_2d(&tb_lo, b_lo, K, Nc, Kp, TN))
    return GPSA_ERR_CUDA;
  GemmParams p = {};
  p.n_mt = gpsa_cdiv(Mr, TM);
  p.n_nt = gpsa_cdiv(Nc, TN);
  p.group_m = p.n_mt;
  p.kblocks = gpsa_cdiv(K, BK);
  set_split(p, split);
  p.Mrows = Mr; p.Ncols = Nc; p.C = C; p.ldc = Nc; p.alpha = 1.f;
  p.accumulate = p.n_split > 1;
  if (p.accumulate && cudaMemsetAsync(C, 0, sizeof(float) * (size_t)Mr * Nc, st) != cudaSuccess) return GPSA_ERR_CUDA;
  return launch_gemm<MODE_TEST>(ta_hi, ta_lo, tb_hi, tb_lo, p, st);
}

// Generic tcgen05 GEMM on fp32 data: C[b] (op)= alpha * A[b] B[b]^T, 3-pass bf16 split, fp32 accumulate.
//   A: Mr x K, element (i,k) at A[b*sA + i*lda + k] (a_rows_major = 1) or A[b*sA + k*lda + i] (= 0);  B likewise, Nc x K.
//   out_mode 0: C[b*sC + i*ldc + j] = v      (split-K > 1: C is zeroed, then atomically accumulated)
//   out_mode 1: C[j*ldc + i] += v            (transposed read-modify-write; batch must be 1, no split)
extern "C" size_t gpsa_gemm_tc_ws_bytes(long Mr, long Nc, int K, int batch) {
  const size_t Kp = (size_t)rup(K, 8);
  if (batch < 1) batch = 1;
  return 2 * al256((size_t)batch * Mr * Kp * 2) + 2 * al256((size_t)batch * Nc * Kp * 2) + 256;
}

extern "C" int gpsa_gemm_tc(long Mr, long Nc, int K, int batch, const float* A, long lda, long sA, int a_rows_major,
                            const float* B, long ldb, long sB, int b_rows_major, float* C, long ldc, long sC, float alpha,
                            int out_mode, int split, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (Mr <= 0 || Nc <= 0 || K <= 0) return GPSA_OK;
  if (batch < 1) batch = 1;
  if (out_mode == 1 && batch != 1) return GPSA_ERR_ARG;
  if (ws_bytes < gpsa_gemm_tc_ws_bytes(Mr, Nc, K, batch)) return GPSA_ERR_ARG;
  const int Kp = (int)rup(K, 8);
  const size_t sa = al256((size_t)batch * Mr * Kp * 2), sb = al256((size_t)batch * Nc * Kp * 2);
  uint8_t* w = static_cast<uint8_t*>(ws);
  __nv_bfloat16 *a_hi = (__nv_bfloat16*)w, *a_lo = (__nv_bfloat16*)(w + sa), *b_hi = (__nv_bfloat16*)(w + 2 * sa),
                *b_lo = (__nv_bfloat16*)(w + 2 * sa + sb);
  auto pack = [&](long rows, const float* src, long ld, long sIn, int rows_major, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    if (rows_major && pack4_ok(src, ld, sIn, K, Kp)) {
      long blocks = (rows + 7) / 8;
      if (blocks > 148 * 8) blocks = 148 * 8;
      dim3 grid((unsigned)blocks, 1, batch);
      pack_rows4_kernel<<<grid, 256, 0, st>>>(rows, K / 4, Kp, ld, sIn, src, hi, lo);
    } else if (rows_major) {
      long blocks = (rows * K + 255) / 256;
      if (blocks > 148 * 8) blocks = 148 * 8;
      dim3 grid((unsigned)blocks, 1, batch);
      pack_rows_kernel<<<grid, 256, 0, st>>>(rows, K, Kp, ld, sIn, src, hi, lo);
    } else {
      dim3 grid((unsigned)((long)gpsa_cdiv(rows, 64) * gpsa_cdiv(K, 64)), 1, batch), block(32, 8);
      pack_trans_kernel<<<grid, block, 0, st>>>(rows, K, Kp, ld, sIn, src, hi, lo);
    }
  };
  pack(Mr, A, lda, sA, a_rows_major, a_hi, a_lo);
  GPSA_LAUNCH_CHECK();
  pack(Nc, B, ldb, sB, b_rows_major, b_hi, b_lo);
  GPSA_LAUNCH_CHECK();
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  const bool batched = batch > 1;
  if (batched) {
    const uint64_t da[3] = {(uint64_t)K, (uint64_t)Mr, (uint64_t)batch}, sa_[2] = {(uint64_t)Kp * 2, (uint64_t)Mr * Kp * 2};
    const uint64_t db[3] = {(uint64_t)K, (uint64_t)Nc, (uint64_t)batch}, sb_[2] = {(uint64_t)Kp * 2, (uint64_t)Nc * Kp * 2};
    const uint32_t ba[3] = {(uint32_t)BK, (uint32_t)TM, 1}, bb[3] = {(uint32_t)BK, (uint32_t)TN, 1};
    if (make_tmap(&ta_hi, a_hi, 3, da, sa_, ba) || make_tmap(&ta_lo, a_lo, 3, da, sa_, ba) ||
        make_tmap(&tb_hi, b_hi, 3, db, sb_, bb) || make_tmap(&tb_lo, b_lo, 3, db, sb_, bb))
      return GPSA_ERR_CUDA;
  } else {
    if (make_tmap_2d(&ta_hi, a_hi, K, Mr, Kp, TM) || make_tmap_2d(&ta_lo, a_lo, K, Mr, Kp, TM) ||
        make_tmap_2d(&tb_hi, b_hi, K, Nc, Kp, TN) || make_tmap_2d(&tb_lo, b_lo, K, Nc, Kp, TN))
      return GPSA_ERR_CUDA;
  }
  GemmParams p = {};
  p.n_mt = gpsa_cdiv(Mr, TM);
  p.n_nt = gpsa_cdiv(Nc, TN);
  p.group_m = p.n_mt < 32 ? p.n_mt : 32;
  p.kblocks = gpsa_cdiv(K, BK);
  if (split <= 0) {  // auto: split K until the grid covers the machine ~4 times, at least 16 K blocks per split
    const long tiles = (long)p.n_mt * p.n_nt * batch;
    split = 1;
    if (tiles < 4L * sm_count()) split = (int)((4L * sm_count() + tiles - 1) / tiles);
    const int max_split = p.kblocks / 16 > 0 ? p.kblocks / 16 : 1;
    if (split > max_split) split = max_split;
  }
  if (out_mode == 1) split = 1;
  set_split(p, split);
  p.Mrows = Mr; p.Ncols = Nc; p.C = C; p.ldc = ldc; p.sC = sC; p.alpha = alpha;
  p.batch = batched ? batch : 0;
  p.trans_add = out_mode == 1;
  p.accumulate = p.n_split > 1;
  if (p.accumulate) {
    if (ldc != Nc) return GPSA_ERR_ARG;
    if (cudaMemsetAsync(C, 0, sizeof(float) * (size_t)batch * Mr * Nc, st) != cudaSuccess) return GPSA_ERR_CUDA;
  }
  return launch_gemm<MODE_TEST>(ta_hi, ta_lo, tb_hi, tb_lo, p, st);
}

// fp32 TMA map of A [M, R] (box 32 r x 8 rows, 128-byte swizzle) for the generator modes; a 16-byte row pitch is
// required, so an R that is not a multiple of 4 goes through a padded copy in `pad` (apad bytes)
static int make_raw_map(CUtensorMap* tm, int M, long R, const float* A, void* pad, cudaStream_t st) {
  const float* Asrc = A;
  long pitch = R;
  if (R % 4 != 0) {
    pitch = rup(R, 4);
    float* Ap = reinterpret_cast<float*>(pad);
    if (cudaMemcpy2DAsync(Ap, pitch * 4, A, R * 4, R * 4, M, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return GPSA_ERR_CUDA;
    Asrc = Ap;
  }
  const uint64_t dims[2] = {(uint64_t)R, (uint64_t)M}, str[1] = {(uint64_t)pitch * 4};
  const uint32_t box[2] = {32, 8};
  return make_tmap(tm, Asrc, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

// Forward quadratic form in its implicit-feature form (MODE_FWD above): q2 [R, L] = Phi(A) W(Omega).
extern "C" int gpsa_quadform_fwd_feat_tc(int M, long R, int L, const float* A, const float* Omega, float* q2, void* ws,
                                         size_t ws_bytes, cudaStream_t st) {
  if (M <= 0 || R <= 0 || L <= 0) return GPSA_OK;
  if (!gpsa_tc_supported(M)) return GPSA_ERR_UNSUPPORTED;
  const FeatFwdLayout f = featfwd_layout(M, R, L);
  if (ws_bytes < f.total) return GPSA_ERR_ARG;
  uint8_t* w = static_cast<uint8_t*>(ws);
  __nv_bfloat16 *wg_hi = (__nv_bfloat16*)w, *wg_lo = (__nv_bfloat16*)(w + f.wg);
  pack_Wg_kernel<<<L, 256, 0, st>>>(M, f.NF, Omega, wg_hi, wg_lo);
  GPSA_LAUNCH_CHECK();
  CUtensorMap tb_hi, tb_lo, ta_raw;
  if (make_tmap_2d(&tb_hi, wg_hi, f.NF, L, f.NF, TN) || make_tmap_2d(&tb_lo, wg_lo, f.NF, L, f.NF, TN)) return GPSA_ERR_CUDA;
  if (make_raw_map(&ta_raw, M, R, A, w + 2 * f.wg, st)) return GPSA_ERR_CUDA;
  GemmParams p = {};
  p.n_mt = gpsa_cdiv(R, TM);
  p.n_nt = gpsa_cdiv(L, TN);
  p.group_m = p.n_mt;  // all row tiles of one 256-gene panel run together: the packed W panel (NF x 256 x 4 B) is shared through L2
  p.kblocks = (int)feat_nblk(M);
  set_split(p, 1);  // one item walks all feature blocks (in chains of KB_CHAIN, combined in TMEM)
  p.sync = reinterpret_cast<unsigned int*>(w + 2 * f.wg + f.apad);
  if (cudaMemsetAsync(p.sync, 0, sizeof(unsigned int), st) != cudaSuccess) return GPSA_ERR_CUDA;
  p.Mrows = R; p.Ncols = L; p.C = q2; p.ldc = L; p.alpha = 1.f;
  p.Amat = A; p.R = R; p.Mind = M; p.nb = feat_nb(M); p.nblk = (int)feat_nblk(M);
  gpsa_prof_begin(0, st);  // bench.py's roofline times the product kernel itself (operand packs excluded)
  const int rc = launch_gemm<MODE_FWD>(ta_raw, ta_raw, tb_hi, tb_lo, p, st);
  gpsa_prof_end(0, st);
  return rc;
}

extern "C" int gpsa_quadform_bwd_alpha_tc(int M, long R, int L, const float* A, const float* G, const float* Omega,
                                          float* Abar, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (M <= 0 || R <= 0 || L <= 0) return GPSA_OK;
  const AlphaLayout a = alpha_layout(M, R, L);
  if (ws_bytes < a.total) return GPSA_ERR_ARG;
  uint8_t* w = static_cast<uint8_t*>(ws);
  __nv_bfloat16 *g_hi = (__nv_bfloat16*)w, *g_lo = (__nv_bfloat16*)(w + a.g), *w_hi = (__nv_bfloat16*)(w + 2 * a.g),
                *w_lo = (__nv_bfloat16*)(w + 2 * a.g + a.w);
  if (pack4_ok(G, L, 0, L, a.Lp)) {
    long blocks = (R + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_rows4_kernel<<<dim3((unsigned)blocks, 1, 1), 256, 0, st>>>(R, L / 4, a.Lp, L, 0, G, g_hi, g_lo);
  } else {
    pack_G_kernel<<<148 * 8, 256, 0, st>>>(R, L, a.Lp, G, g_hi, g_lo);
  }
  GPSA_LAUNCH_CHECK();
  {
    dim3 grid((unsigned)feat_nblk(M), (unsigned)((L + 31) / 32 < 64 ? (L + 31) / 32 : 64));
    pack_Wt_kernel<<<grid, 256, 0, st>>>(M, L, a.Lp, Omega, w_hi, w_lo);
    GPSA_LAUNCH_CHECK();
  }
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if (make_tmap_2d(&ta_hi, g_hi, L, R, a.Lp, TM) || make_tmap_2d(&ta_lo, g_lo, L, R, a.Lp, TM) ||
      make_tmap_2d(&tb_hi, w_hi, L, a.NF, a.Lp, TN) || make_tmap_2d(&tb_lo, w_lo, L, a.NF, a.Lp, TN))
    return GPSA_ERR_CUDA;
  GemmParams p = {};
  p.n_mt = gpsa_cdiv(R, TM);
  p.n_nt = gpsa_cdiv(a.NF, TN);
  // 32 row tiles of G (hi+lo) stay L2-resident while the feature panels of W^T sweep past.  (The packed W^T is re-read
  // once per group -- 5.2 GB of the 6.8 GB DRAM reads ncu sees at C3 -- but DRAM is not the bound here: groups of 60
  // row tiles halved those reads and made the kernel 5 % SLOWER, 26.4 vs 25.0 ms, profiles/r2_products.txt.)
  p.group_m = 32;
  if (p.group_m > p.n_mt) p.group_m = p.n_mt;
  p.kblocks = gpsa_cdiv(L, BK);
  set_split(p, 1);
  p.Mrows = R; p.Ncols = a.NF;
  p.Amat = A; p.R = R; p.Mind = M; p.nb = feat_nb(M); p.nblk = (int)feat_nblk(M); p.Abar = Abar;
  gpsa_prof_begin(1, st);
  const int rc = launch_gemm<MODE_ALPHA>(ta_hi, ta_lo, tb_hi, tb_lo, p, st);
  gpsa_prof_end(1, st);
  return rc;
}

extern "C" int gpsa_quadform_bwd_omega_tc(int M, long R, int L, const float* A, const float* G, float* H, void* ws,
                                          size_t ws_bytes, cudaStream_t st) {
  if (M <= 0 || R <= 0 || L <= 0) return GPSA_OK;
  const OmegaLayout o = omega_layout(M, R, L);
  if (ws_bytes < o.total) return GPSA_ERR_ARG;
  uint8_t* w = static_cast<uint8_t*>(ws);
  __nv_bfloat16 *gt_hi = (__nv_bfloat16*)w, *gt_lo = (__nv_bfloat16*)(w + o.gt);
  {
    dim3 grid((unsigned)(o.Rp / 64), gpsa_cdiv(L, 64)), block(32, 8);
    pack_Gt_kernel<<<grid, block, 0, st>>>(R, L, G, gt_hi, gt_lo);
    GPSA_LAUNCH_CHECK();
  }
  CUtensorMap tb_hi, tb_lo, ta_raw;
  {
    // [K block][gene][64 r]: box = 64 r x 256 genes x 1 block = one contiguous tile
    const uint64_t dims[3] = {(uint64_t)BK, (uint64_t)L, (uint64_t)(o.Rp / BK)};
    const uint64_t str[2] = {(uint64_t)BK * 2, (uint64_t)L * BK * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)TN, 1};
    if (make_tmap(&tb_hi, gt_hi, 3, dims, str, box) || make_tmap(&tb_lo, gt_lo, 3, dims, str, box)) return GPSA_ERR_CUDA;
  }
  if (make_raw_map(&ta_raw, M, R, A, w + 2 * o.gt, st)) return GPSA_ERR_CUDA;
  const long NF = feat_nblk(M) * FBK;
  GemmParams p = {};
  p.n_mt = gpsa_cdiv(NF, TM);
  p.n_nt = gpsa_cdiv(L, TN);
  p.group_m = p.n_mt;  // all feature tiles of one gene chunk run together: the G^T slab is shared through L2
  p.kblocks = gpsa_cdiv(R, BK);
  // split K (spots) until the grid fills the machine a few times over
  const int tiles = p.n_mt * p.n_nt;
  int split = 1;
  if (tiles < 4 * sm_count()) split = (4 * sm_count() + tiles - 1) / tiles;
  const int max_split = p.kblocks / 16 > 0 ? p.kblocks / 16 : 1;
  if (split > max_split) split = max_split;
  set_split(p, split);
  p.Mrows = NF; p.Ncols = L; p.C = H; p.ldc = L;
  p.accumulate = p.n_split > 1;
  if (p.accumulate && cudaMemsetAsync(H, 0, sizeof(float) * (size_t)NF * L, st) != cudaSuccess) return GPSA_ERR_CUDA;
  p.Amat = A; p.R = R; p.Mind = M; p.nb = feat_nb(M); p.nblk = (int)feat_nblk(M);
  if (p.n_split == 1) {
    p.sync = reinterpret_cast<unsigned int*>(w + 2 * o.gt + o.apad);
    if (cudaMemsetAsync(p.sync, 0, sizeof(unsigned int), st) != cudaSuccess) return GPSA_ERR_CUDA;
  }
  gpsa_prof_begin(2, st);
  const int rc = launch_gemm<MODE_OMEGA>(ta_raw, ta_raw, tb_hi, tb_lo, p, st);
  gpsa_prof_end(2, st);
  return rc;
}
