/* gpsa_b200 -- C ABI of the B200-native GPSA variational-ELBO hot path.
 *
 * The reference (andrewcharlesjones/spatial-alignment, gpsa 0.6) is pure Python/PyTorch and has no
 * FFI of its own; this header is the boundary underneath its Python API.  Every entry point names
 * the reference code it replaces.  All pointers are DEVICE pointers to contiguous row-major fp32
 * unless marked otherwise; every call is asynchronous on `stream`, never synchronises with the
 * host and never allocates.  Return value: 0 = ok, 1 = bad argument, 2 = CUDA launch error,
 * 3 = unsupported size/kind.
 *
 * Shapes use the symbols of SURVEY.md:  M inducing points, D spatial dims (1..3), R = S*N rows
 * (Monte-Carlo sample s, spot n; r = s*N + n), L latent outputs (genes), V views.
 * Kernel kinds: 0 = rbf (gpsa/util/util.py:8-23), 1 = matern12 (gpsa/util/util.py:33-47),
 * 2 = matern32 (gpsa/util/util.py:50-66), 3 = external: K_uu / K_uf are evaluated (and differentiated) by the caller --
 * the path for user-supplied covariance callables (kernel_func_warp / kernel_func_data, gpsa/models/vgpsa.py:25-26);
 * the layer entry points then take the matrices and hand back dLoss/dK_uu, dLoss/dK_uf.
 */
#ifndef GPSA_B200_H
#define GPSA_B200_H

#ifdef __CUDACC__
#include <cuda_runtime.h>
#else
#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;
#endif
#endif

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int gpsa_version(void);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------
 * gpsa_launch_count: kernels this library has launched so far in this process.
 * gpsa_prof_enable(1): record CUDA events on the launching stream around the hot kernels
 *   (slot 0 = quadratic form forward, 1 = its A-bar backward, 2 = its Omega-bar backward).
 * gpsa_prof_read: synchronise, return per slot the number of timed launches and their total ms. */
long gpsa_launch_count(void);
/* testing aid: 1 = use the scalar variants of the sampling / log-likelihood kernels where the 128-bit ones apply */
void gpsa_debug_disable_vec4(int off);
void gpsa_prof_enable(int on);
int gpsa_prof_read(int* counts, double* total_ms);

/* ---- covariance functions -------------------------------------------------------------------
 * K[m,r] = k(x1[m], x2[r]);  x1 [M,D], x2 [R,D], K [M,R].  log_ls / log_var: device scalars.
 * Replaces rbf_kernel / matern12_kernel (gpsa/util/util.py:8-23, :33-47) as called at
 * gpsa/models/vgpsa.py:314-318, :390, :409. */
int gpsa_kernel_matrix_fwd(int kind, int D, int M, long R, const float* x1, const float* x2, const float* log_ls,
                           const float* log_var, float* K, cudaStream_t stream);
/* Gradient of sum(Kbar * K).  acc_x1 [M,D] and acc_hyp [2] = (d/dlog_ls, d/dlog_var) are fp64
 * accumulators that are ADDED to; x2bar [R,D] is overwritten (may be NULL); acc_x2 (fp64 [R,D],
 * may be NULL) is added to instead of writing x2bar -- used for k(Z,Z) where x2 is the parameter
 * itself.  Replaces autograd through the same reference lines. */
int gpsa_kernel_matrix_bwd(int kind, int D, int M, long R, const float* x1, const float* x2, const float* log_ls,
                           const float* log_var, const float* Kbar, double* acc_x1, float* x2bar, double* acc_x2,
                           double* acc_hyp, cudaStream_t stream);

/* ---- batched Cholesky / triangular inverse, one CTA per matrix ------------------------------
 * A [batch,M,M] is factorised in place (lower factor, zeros above, like torch.cholesky);
 * half_logdet[b] = sum_i log L_ii; info[b] = 1 if a non-positive pivot was met.  Either may be NULL.
 * Replaces torch.cholesky at gpsa/models/vgpsa.py:257, :320, :394, :412. */
int gpsa_potrf_batched_f32(int M, int batch, float* A, float* half_logdet, int* info, cudaStream_t stream);
int gpsa_potrf_batched_f64(int M, int batch, double* A, double* half_logdet, int* info, cudaStream_t stream);
/* X = L^-1 (lower).  X must not alias L.  Stands in for the triangular solves of
 * torch.cholesky_solve (gpsa/models/vgpsa.py:177) and of the MultivariateNormal KL (:506-530). */
/* fp32 factorisation out of place (A: lower triangle read; L may alias A) with the log-determinant in fp64. */
int gpsa_potrf_batched_f32_ld64(int M, int batch, const float* A, float* L, double* half_logdet, int* info,
                                cudaStream_t stream);
int gpsa_trtri_batched_f32(int M, int batch, const float* L, float* X, cudaStream_t stream);
int gpsa_trtri_batched_f64(int M, int batch, const double* L, double* X, cudaStream_t stream);

/* ---- plain strided batched GEMM (fp32 SIMT) --------------------------------------------------
 * C[b] = alpha * op(A[b]) op(B[b]) + beta * C[b];  element (i,k) of op(A) at A[b*sA + i*ars + k*acs],
 * (k,j) of op(B) at B[b*sB + k*brs + j*bcs], C row-major with leading dimension ldc. */
int gpsa_gemm_f32(int M, int N, long K, float alpha, const float* A, long ars, long acs, long sA, const float* B,
                  long brs, long bcs, long sB, float beta, float* C, long ldc, long sC, int batch,
                  cudaStream_t stream);

/* ---- prior covariance K_uu: fp64 factorisation of k(Z,Z) + 1e-5 I -----------------------------
 * Outputs the fp32 Cholesky factor Lk [M,M] (the reference's cached Kuu_chol_*), K^-1 in fp32
 * (Kinv, may be NULL) and fp64 (Kinv64) and half_logdet (fp64 scalar, sum log diag).
 * ws64: 2*M*M doubles.  info: 1 int.  Replaces gpsa/models/vgpsa.py:314-321 and :390-394. */
int gpsa_prior_prepare(int kind, int D, int M, const float* Z, const float* log_ls, const float* log_var, float* Lk,
                       float* Kinv, double* Kinv64, double* half_logdet, int* info, double* ws64,
                       cudaStream_t stream);

/* Same for kind 3: Kuu [M,M] fp32 (without the 1e-5 jitter) evaluated by the caller. */
int gpsa_prior_prepare_ext(int M, const float* Kuu, float* Lk, float* Kinv, double* Kinv64, double* half_logdet, int* info,
                           double* ws64, cudaStream_t stream);

/* ---- variational covariances Omega = Omega_sqt Omega_sqt^T + 1e-5 I, batched ------------------
 * Replaces get_Omega_from_Omega_sqt + torch.cholesky (gpsa/models/vgpsa.py:206-210, :255-257, :410-412).
 * Omega, Ltril: [B,M,M]; half_logdet [B]; info [B]. */
int gpsa_omega_prepare(int M, int B, const float* Osq, float* Omega, float* Ltril, double* L64, double* half_logdet,
                       int* info, cudaStream_t stream);
/* L64 == NULL selects the fp32 factorisation (large gene batches): Omega is still accumulated in fp64 and rounded once,
 * the Cholesky runs in fp32 (what the reference does, :410-412), the log-determinants are summed in fp64. */
/* Backward: Osq_bar = 2 (Obar + coef[b] * Omega^-1) Osq.  Obar [B,M,M] symmetric; coef: device [B]
 * (the -1/2 dKL factor of the log-det term, 0 for slices without a KL term; NULL = no such term);
 * L64: the fp64 factor from gpsa_omega_prepare; Linv64, Y64: [B,M,M] fp64 scratch. */
int gpsa_omega_grad(int M, int B, const float* Osq, const double* L64, const float* Obar, const float* coef,
                    double* Linv64, double* Y64, float* Osq_bar, cudaStream_t stream);
/* fp32 form for factors from the fp32 branch of gpsa_omega_prepare (Ltril fp32; Linv32, Y32: [B,M,M] fp32 scratch). */
int gpsa_omega_grad_f32(int M, int B, const float* Osq, const float* Ltril, const float* Obar, const float* coef,
                        float* Linv32, float* Y32, float* Osq_bar, void* tc_ws, size_t tc_ws_bytes, cudaStream_t stream);
/* Same, with the 2 Obar Osq product on the tcgen05 engine; tc_ws: gpsa_gemm_tc_ws_bytes(M, M, M, B) bytes. */
int gpsa_omega_grad_tc(int M, int B, const float* Osq, const double* L64, const float* Obar, const float* coef,
                       double* Linv64, double* Y64, float* Osq_bar, void* tc_ws, size_t tc_ws_bytes, cudaStream_t stream);

/* ---- implicit-feature quadratic form (the hot contraction) ------------------------------------
 * q2[r,p] = a_r^T Omega_p a_r with a_r = A[:,r];  replaces the [S,L,N,M] broadcast bmm at
 * gpsa/models/vgpsa.py:193-196 and its autograd.  W is the packed feature matrix [gpsa_feat_count(M), L]. */
long gpsa_feat_count(int M);
int gpsa_feat_pack(int M, int L, const float* Omega, float* W, cudaStream_t stream);
/* Obar[p] = sym(H[:,p]) + add_scale * (*add_scale_dev) * Add   (Add [M,M], may be NULL) */
int gpsa_feat_unpack(int M, int L, const float* H, const float* Add, float add_scale, const float* add_scale_dev,
                     float* Obar, cudaStream_t stream);
int gpsa_quadform_fwd_f32(int M, long R, int L, const float* A, const float* W, float* q2, cudaStream_t stream);
int gpsa_quadform_bwd_omega_f32(int M, long R, int L, const float* A, const float* G, float* H,
                                cudaStream_t stream);
int gpsa_quadform_bwd_alpha_f32(int M, long R, int L, const float* A, const float* G, const float* W, float* Abar,
                                cudaStream_t stream);

/* ---- tensor-core (tcgen05) engine for the same three products ------------------------------------
 * bf16 hi/lo split operands, three MMA passes, fp32 accumulation in TMEM in chains of bounded length (error ~2^-16 per
 * product).  All three products are implicit-feature GEMMs over the packed symmetric features of Omega: the forward is
 * q2 = Phi(A) W(Omega) with Phi[r,(i,j)] = a_r[i] a_r[j] generated on the fly (gpsa/models/vgpsa.py:193-196; the
 * Cholesky factor is not needed for it).  `ws` is caller-provided device scratch of at least
 * gpsa_quadform_tc_ws_bytes(M, R, L) bytes (bf16 copies of the operands); gpsa_tc_supported(M) says whether the engine
 * covers this M (16 <= M <= 512).  G [R,L] = dLoss/dq2, Abar [M,R] is ADDED to, H as for the fp32 engine. */
int gpsa_tc_supported(int M);
size_t gpsa_quadform_tc_ws_bytes(int M, long R, int L);
int gpsa_quadform_fwd_feat_tc(int M, long R, int L, const float* A, const float* Omega, float* q2, void* ws,
                              size_t ws_bytes, cudaStream_t stream);
int gpsa_quadform_bwd_alpha_tc(int M, long R, int L, const float* A, const float* G, const float* Omega, float* Abar,
                               void* ws, size_t ws_bytes, cudaStream_t stream);
int gpsa_quadform_bwd_omega_tc(int M, long R, int L, const float* A, const float* G, float* H, void* ws,
                               size_t ws_bytes, cudaStream_t stream);
/* Generic fp32-in / fp32-out GEMM on the same engine: C[b] (op)= alpha * A[b] B[b]^T.  A is Mr x K with element (i,k) at
 * A[b*sA + i*lda + k] (a_rows_major = 1) or A[b*sA + k*lda + i] (0); B likewise (Nc x K).  out_mode 0: C[b*sC + i*ldc + j] = v
 * (split > 1 or split <= 0 = auto: C zeroed then atomically accumulated, needs ldc == Nc); out_mode 1: C[j*ldc + i] += v
 * (batch 1).  Replaces the predictive-mean matmul (gpsa/models/vgpsa.py:182-184) and the autograd matmuls of the
 * data layer.  ws: gpsa_gemm_tc_ws_bytes(Mr, Nc, K, batch) bytes. */
size_t gpsa_gemm_tc_ws_bytes(long Mr, long Nc, int K, int batch);
int gpsa_gemm_tc(long Mr, long Nc, int K, int batch, const float* A, long lda, long sA, int a_rows_major, const float* B,
                 long ldb, long sB, int b_rows_major, float* C, long ldc, long sC, float alpha, int out_mode, int split,
                 void* ws, size_t ws_bytes, cudaStream_t stream);
/* Plain C [Mr,Nc] = A [Mr,K] B[Nc,K]^T through the same TMA / tcgen05 / TMEM core (unit-test hook). */
int gpsa_tc_gemm_test(int Mr, int Nc, int K, const float* A, const float* B, float* C, int split, void* ws,
                      size_t ws_bytes, cudaStream_t stream);

/* ---- warp layer, one non-fixed view -------------------------------------------------------------
 * Replaces the body of the view loop, gpsa/models/vgpsa.py:275-351 (K_uu, K_uf, compute_mean_and_var
 * :174-204 in its 2-D branch, reparameterised sampling) and this view's KL terms in loss_fn (:498-516).
 * Reference quirks kept: sample scale is the variance (:334-340), variance uses Omega slice v*D+j
 * (:336-339) while the KL uses slice j*V+v (:508), jitter added twice (:191,:204). */
typedef struct {
  int kind, D, M, V, v, S;
  long n;                 /* spots of this view (all modalities concatenated) */
  const float* Z;         /* Xtilde[v]            [M,D] */
  const float* dlt;       /* delta_G_list[v]      [M,D] */
  const float* log_ls;    /* &warp_kernel_lengthscales[v] */
  const float* log_var;   /* &warp_kernel_variances[v]    */
  const float* Omega_G;   /* [V*D,M,M] from gpsa_omega_prepare */
  const double* hld_Omega; /* [V*D] half log-dets of Omega_G */
  const float* X;         /* [n,D] observed coordinates */
  const float* eps;       /* [S,n,D] standard-normal draws */
  float* Lk;              /* out [M,M]  Kuu_chol_list[v] */
  float* Kinv;            /* out [M,M] fp32 copy of K^-1 (may be NULL) */
  double* Kinv64;         /* out [M,M] fp64 K^-1 (saved) */
  double* hld_K;          /* out fp64 scalar */
  int* info;              /* out 1 int */
  double* A;              /* out fp64 [M,n]  K^-1 K_uf   (saved) */
  double* B;              /* out fp64 [M,n]  K_uf        (saved) */
  double* T;              /* out fp64 [D,M,n] Omega_{vD+j} A (saved) */
  double* Ke;             /* out fp64 [D,M]  K^-1 (Z_j - dlt_j)  (saved) */
  float* var;             /* out [n,D] marginal "variance" (saved) */
  float* Gmean;           /* out [n,D] */
  float* Gs;              /* out: sample s at Gs + s*gs_stride, [n,D] each */
  long gs_stride;
  double* kl_acc;         /* fp64 scalar, ADDED to (may be NULL) */
  double* ws64;           /* 2*M*M doubles */
  const float* Kuu_ext;   /* kind 3 only: k(Z,Z) [M,M] and k(Z,X) [M,n] evaluated by the caller; else NULL */
  const float* Kuf_ext;
} gpsa_warp_fwd_args;
int gpsa_warp_view_fwd(const gpsa_warp_fwd_args* a, cudaStream_t stream);

typedef struct {
  int kind, D, M, V, v, S;
  long n;
  const float *Z, *dlt, *log_ls, *log_var, *Omega_G, *X, *eps;
  const double *Kinv64, *A, *B, *T, *Ke;   /* fp64, saved by the forward */
  const float* Gs_bar;    /* sample s at Gs_bar + s*gs_stride, [n,D]; may be NULL */
  long gs_stride;
  const float* Gm_bar;    /* [n,D], may be NULL */
  const float* kl_bar;    /* device scalar: upstream gradient of the KL term */
  double* acc_Z;          /* fp64 [M,D], ADDED to: gradient of Xtilde[v] */
  double* acc_dlt;        /* fp64 [M,D], ADDED to: gradient of delta_G_list[v] */
  double* acc_hyp;        /* fp64 [2]: (log_ls, log_var) of this view */
  float* Obar_G;          /* [V*D,M,M] ADDED to (slices v*D+j and j*V+v) */
  /* scratch */
  float *mubar, *varbar, *q1bar; /* [n,D], [n,D], [n] */
  double *Abar, *C, *AS;         /* fp64 [M,n], [M,n], [D,M,n] */
  double* ws64;                  /* 3*M*M doubles */
  float* Kuu_bar;                /* kind 3 only: out [M,M] dLoss/dK_uu, out [M,n] dLoss/dK_uf; else NULL */
  float* Kuf_bar;
} gpsa_warp_bwd_args;
int gpsa_warp_view_bwd(const gpsa_warp_bwd_args* a, cudaStream_t stream);

/* ---- data layer, one modality -------------------------------------------------------------------
 * Replaces gpsa/models/vgpsa.py:390-421 (K_uu, K_uf for all S samples, compute_mean_and_var in its 3-D branch
 * :192-204 up to the predictive moments) and the modality's KL term (:520-530).  Outputs the predictive mean,
 * the quadratic form q2[r,p] = a_r^T Omega_p a_r and kq[r] = sigma^2 - a_r^T K a_r; the marginal variance is
 * var = kq + q2 + 2e-5 and the reparameterised sample F = mean + sqrt(var) eps (:423-426) is the SAMPLING STAGE below,
 * either materialised (gpsa_sample_fwd / _bwd) or fused with the likelihood (gpsa_sample_ll_fused). */
typedef struct {
  int kind, D, M, L;
  long R;                 /* S*N rows */
  const float* Gt;        /* Gtilde [M,D] */
  const float *log_ls, *log_var;
  const float* dlt;       /* delta_F [M,L] */
  const float* Omega;     /* [L,M,M] from gpsa_omega_prepare */
  const double* hld_Omega; /* [L] */
  const float* G;         /* G_samples flattened [R,D] */
  float *Lk, *Kinv;       /* out [M,M] each (fp32) */
  double* Kinv64;         /* out [M,M] fp64 (saved) */
  double* hld_K;
  int* info;
  float *A, *B;           /* out [M,R] (saved) */
  float* kq;              /* out [R]  K_ff - a^T K a */
  float* W;               /* out [gpsa_feat_count(M), L] (saved; engine 0 only) */
  double* KD;             /* out fp64 [M,L] K^-1 delta (saved) */
  float* mean;            /* out [R,L] predictive mean */
  float* q2;              /* out [R,L] quadratic form */
  double* kl_acc;         /* fp64 scalar, ADDED to; may be NULL (prediction) */
  double* ws64;           /* 2*M*M doubles */
  int engine;             /* 0 = fp32 SIMT quadratic form (launch-bound toy shapes), 1 = tcgen05 split-bf16 */
  void* tc_ws;            /* engine 1: scratch, gpsa_quadform_tc_ws_bytes(M, R, L) bytes */
  size_t tc_ws_bytes;
  const float* Kuu_ext;   /* kind 3 only: k(Gt,Gt) [M,M] evaluated by the caller, and B is an INPUT holding k(Gt,G) [M,R] */
  int prior_ready;        /* 1: Lk, Kinv, Kinv64, hld_K, info are INPUTS, already filled by gpsa_prior_prepare(_ext) -- the
                             caller factorised K_uu ahead of time (it depends on parameters only), e.g. on another stream */
} gpsa_data_fwd_args;
int gpsa_data_layer_fwd(const gpsa_data_fwd_args* a, cudaStream_t stream);

typedef struct {
  int kind, D, M, L;
  long R;
  const float *Gt, *log_ls, *log_var, *dlt, *Omega, *G;
  const float* Kinv;
  const double* Kinv64;
  const float *A, *B, *W;
  const double* KD;
  const float* mean_bar;  /* [R,L] upstream gradient of the predictive mean */
  const float* q2_bar;    /* [R,L] upstream gradient of the quadratic form (= of the marginal variance) */
  const float* kq_bar;    /* [R]   upstream gradient of kq (= sum_p q2_bar[r,p] when var = kq + q2 + const) */
  const float* kl_bar;    /* device scalar */
  float* G_bar;           /* out [R,D] */
  double* acc_Gt;         /* fp64 [M,D] ADDED to */
  double* acc_hyp;        /* fp64 [2] ADDED to */
  float* dlt_bar;         /* out [M,L] */
  float* Obar;            /* out [L,M,M]  (feed to gpsa_omega_grad) */
  /* scratch */
  float* q1bar;           /* [R] */
  float *Abar, *C;        /* [M,R] each */
  float* H;               /* [gpsa_feat_count(M), L] */
  double* ws64;           /* 3*M*M doubles */
  int engine;
  void* tc_ws;            /* engine 1: scratch, gpsa_quadform_tc_ws_bytes(M, R, L) bytes */
  size_t tc_ws_bytes;
  float* Kuu_bar;         /* kind 3 only: out [M,M] dLoss/dK_uu; dLoss/dK_uf is C [M,R]; G_bar / acc_Gt are not written */
} gpsa_data_bwd_args;
int gpsa_data_layer_bwd(const gpsa_data_bwd_args* a, cudaStream_t stream);

/* ---- sampling stage ---------------------------------------------------------------------------------
 * Materialised form (reference gpsa/models/vgpsa.py:197-204, :423-426): in: F = mean, var = q2;
 * out: var = kq[r] + q2 + 2e-5 (jitter added twice like the reference), F = mean + sqrt(var) eps.  R = S*N rows. */
int gpsa_sample_fwd(long R, int L, const float* kq, const float* eps, float* F, float* var, cudaStream_t stream);
/* q2_bar[r,p] = F_bar eps / (2 sqrt(var)),  kq_bar[r] = sum_p q2_bar[r,p]   (mean_bar = F_bar) */
int gpsa_sample_bwd(long R, int L, const float* F_bar, const float* eps, const float* var, float* q2_bar, float* kq_bar,
                    cudaStream_t stream);
/* Counter-based standard-normal noise: out[(s*N + n)*L + p] = the draw of (sample samp_off + s, spot n, gene gene_off + p)
 * under the 64-bit seed *key (device memory) -- Philox4x32-10 + Box-Muller, identical to what gpsa_sample_ll_fused
 * draws in-kernel, independent of how genes / samples are sharded.  Stands in for torch.randn at vgpsa.py:423. */
int gpsa_philox_normal(long N, int S, int L, const long long* key, int gene_off, int samp_off, float* out,
                       cudaStream_t stream);
/* Fused form: sampling + Gaussian negative log-likelihood + its gradients in one pass (vgpsa.py:423-426, :532-538).
 * in:  mean_U = predictive mean [S*N, L], q2_Gu = quadratic form [S*N, L], kq [S*N], Y [N, L], log_noise (device scalar);
 *      noise: eps [S*N, L] if non-NULL, else drawn in-kernel from *key (see gpsa_philox_normal).
 * out (IN PLACE): mean_U = d(-LL)/dF = -(Y - F)/(sigma^2 S),  q2_Gu = d(-LL)/dvar = U eps / (2 sqrt(var)),
 *      kqb [S*N] = sum_p q2_Gu,  nll_acc += -LL (fp64),  noise_acc += d(-LL)/dlog_noise (fp64).
 * sigma = exp(*log_noise) + 1e-5 is used as the Normal SCALE exactly as the reference does (:217, :534). */
int gpsa_sample_ll_fused(long N, int S, int L, const float* kq, const float* Y, const float* log_noise, const float* eps,
                         const long long* key, int gene_off, int samp_off, float* mean_U, float* q2_Gu, float* kqb,
                         double* nll_acc, double* noise_acc, cudaStream_t stream);
/* x[0..n) *= *scale unless *scale == 1 (device scalar): the upstream gradient of loss.backward() is exactly 1. */
int gpsa_scale_if_not_one(long n, const float* scale, float* x, cudaStream_t stream);

/* ---- Gaussian log-likelihood ----------------------------------------------------------------------
 * ll_acc += sum_{r,p} log N(Y[n,p]; F[r,p], sigma) / S,  sigma = exp(*log_noise) + 1e-5 used as the
 * Normal SCALE exactly as the reference does (gpsa/models/vgpsa.py:217, :532-538); r = s*N + n.
 * F [S*N,P], Y [N,P]. */
int gpsa_gaussian_ll_fwd(long N, int P, int S, const float* F, const float* Y, const float* log_noise,
                         double* ll_acc, cudaStream_t stream);
/* F_bar[r,p] = ll_bar * dLL/dF,  acc_noise += ll_bar * dLL/dlog_noise (fp64).  ll_bar: device scalar. */
int gpsa_gaussian_ll_bwd(long N, int P, int S, const float* F, const float* Y, const float* log_noise,
                         const float* ll_bar, float* F_bar, double* acc_noise, cudaStream_t stream);

/* ---- linear model of coregionalisation (LMC), reference gpsa/models/vgpsa.py:167-172, :428-432 ---------------
 * Materialised: F_obs [R,P] = F_lat [R,L] W [L,P] and its backward (W_bar overwritten). */
int gpsa_lmc_fwd(long R, int L, int P, const float* F_lat, const float* W, float* F_obs, cudaStream_t stream);
int gpsa_lmc_bwd(long R, int L, int P, const float* F_lat, const float* W, const float* F_obs_bar, float* F_lat_bar,
                 float* W_bar, cudaStream_t stream);
/* Fused with the Gaussian negative log-likelihood of Y [N,P] (no [S,N,P] tensor is formed), for L <= gpsa_lmc_max_latent():
 * nll_acc += -LL, noise_acc += d(-LL)/dlog_noise, F_lat_bar [S*N,L] = d(-LL)/dF_lat, W_bar [L,P] = d(-LL)/dW (both overwritten). */
int gpsa_lmc_max_latent(void);
int gpsa_lmc_ll_fused(long N, int S, int L, int P, const float* F_lat, const float* W, const float* Y,
                      const float* log_noise, float* F_lat_bar, float* W_bar, double* nll_acc, double* noise_acc,
                      cudaStream_t stream);

/* ---- optimiser step and initialisation -------------------------------------------
 * Adam exactly as torch.optim.Adam(lr, betas, eps) with weight_decay = 0, amsgrad = False, maximize = False -- the
 * optimiser every reference training loop uses (examples/grid_example.py:59,76) -- as one launch over up to
 * GPSA_ADAM_MAX_TENSORS parameter tensors.  `step` is a DEVICE array of `count` floats holding the number of steps each
 * tensor has taken so far; the call increments them first (so the launch pair can be captured in a CUDA graph).
 * g[k] == NULL skips tensor k (its step count does not advance, as in torch). */
#define GPSA_ADAM_MAX_TENSORS 24
typedef struct {
  int count;
  float* p[GPSA_ADAM_MAX_TENSORS];        /* parameters, updated in place */
  const float* g[GPSA_ADAM_MAX_TENSORS];  /* gradients */
  float* m[GPSA_ADAM_MAX_TENSORS];        /* exp_avg */
  float* v[GPSA_ADAM_MAX_TENSORS];        /* exp_avg_sq */
  long n[GPSA_ADAM_MAX_TENSORS];          /* elements */
  double lr, beta1, beta2, eps;           /* doubles like torch's Python scalars: 1 - beta2 must not be formed in fp32 */
  float* step;
} gpsa_adam_args;
int gpsa_adam_step(const gpsa_adam_args* a, cudaStream_t stream);

/* Lloyd's k-means on spot coordinates X [N,D] (D <= 3): `iters` assign + update rounds starting from `centres` [K,D]
 * (in/out), then one more assignment; assign [N] int, sums: scratch K*(D+1) doubles, inertia: 1 double (sum of squared
 * distances to the final centres).  An empty cluster keeps its centre.  Replaces sklearn.cluster.KMeans in the
 * data_init branch of gpsa/models/vgpsa.py:61-92 for large inputs. */
int gpsa_kmeans_lloyd(long N, int D, int K, const float* X, float* centres, int iters, int* assign, double* sums,
                      double* inertia, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GPSA_B200_H */
