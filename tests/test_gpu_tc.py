"""tcgen05 engine parity (-m gpu): the TMA / tcgen05 / TMEM GEMM core and the three quadratic-form products
built on it, each called through the C ABI and compared with a float64 torch evaluation of the same
expression.  Tolerance: the engine multiplies bf16 (hi, lo) splits in three passes, i.e. each product
carries ~2^-16 relative error instead of fp32's 2^-24; sums of K such terms are checked scale-relative
(max |err| / max |ref|) at 1e-4, the tolerance BASELINE.json's north_star states for the ELBO path."""
import ctypes as C
import os

import pytest
import torch

from golden_io import relerr

pytestmark = pytest.mark.gpu

f32, f64, u8 = torch.float32, torch.float64, torch.uint8
TOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from gpsa import _lib

    assert torch.cuda.is_available()
    return _lib


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ws_for(L, M, R, Lg):
    return L.tc_workspace(M, R, Lg, torch.empty(1, device="cuda"))


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Mr,Nc,K,split", [(128, 256, 64, 1), (128, 256, 256, 1), (300, 500, 200, 1),
                                           (1000, 700, 1000, 3), (77, 19, 33, 1), (513, 1030, 4100, 4)])
def test_gemm_core(L, Mr, Nc, K, split):
    g = torch.Generator().manual_seed(Mr + Nc + K)
    A = torch.randn(Mr, K, generator=g)
    B = torch.randn(Nc, K, generator=g)
    ref = A.double() @ B.double().T
    a, b = A.cuda(), B.cuda()
    c = torch.full((Mr, Nc), float("nan"), device="cuda")
    ws = torch.empty(4 * (Mr + Nc) * (K + 8) + 4096, dtype=u8, device="cuda")
    rc = L.lib().gpsa_tc_gemm_test(Mr, Nc, K, a.data_ptr(), b.data_ptr(), c.data_ptr(), split, ws.data_ptr(), ws.numel(),
                                   stream())
    assert rc == 0
    torch.cuda.synchronize()
    err = relerr(c.cpu(), ref)
    assert err < 2e-5, err


def _problem(M, R, Lg, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, R, generator=g) * 0.3
    Osq = torch.randn(Lg, M, M, generator=g) * 0.1
    G = torch.randn(R, Lg, generator=g)
    return A, Osq, G


@pytest.mark.parametrize("M,R,Lg", [(64, 128, 1), (64, 1000, 5), (200, 700, 37), (256, 2049, 9), (50, 300, 3),
                                    (100, 4096, 300), (512, 700, 5), (320, 300, 3), (37, 130, 2), (200, 1027, 260)])
def test_quadform_tc_fwd_bwd(L, M, R, Lg):
    from gpsa import _ops

    A, Osq, G = _problem(M, R, Lg, M * 7 + R)
    Omega, Ltril, L64, hld, info = _ops.omega_prepare(Osq.cuda())
    Ad = A.double().requires_grad_()
    Omd = Omega.double().cpu().requires_grad_()
    q2r = torch.einsum("mr,pmk,kr->rp", Ad, Omd, Ad)
    (q2r * G.double()).sum().backward()

    lib = L.lib()
    assert lib.gpsa_tc_supported(M) == 1
    a, gg = A.cuda(), G.cuda()
    ws = ws_for(L, M, R, Lg)
    # forward: the implicit-feature GEMM (what the data layer runs)
    q2f = torch.full((R, Lg), float("nan"), device="cuda")
    assert lib.gpsa_quadform_fwd_feat_tc(M, R, Lg, a.data_ptr(), Omega.data_ptr(), q2f.data_ptr(), ws.data_ptr(),
                                         ws.numel(), stream()) == 0
    torch.cuda.synchronize()
    assert relerr(q2f.cpu(), q2r.detach()) < TOL

    nf = L.feat_count(M)
    H = torch.full((nf, Lg), float("nan"), device="cuda")
    Obar = torch.empty(Lg, M, M, device="cuda")
    assert lib.gpsa_quadform_bwd_omega_tc(M, R, Lg, a.data_ptr(), gg.data_ptr(), H.data_ptr(), ws.data_ptr(), ws.numel(),
                                          stream()) == 0
    assert lib.gpsa_feat_unpack(M, Lg, H.data_ptr(), None, 0.0, None, Obar.data_ptr(), stream()) == 0
    torch.cuda.synchronize()
    assert relerr(Obar.cpu(), Omd.grad) < TOL

    Abar = torch.zeros(M, R, device="cuda")
    assert lib.gpsa_quadform_bwd_alpha_tc(M, R, Lg, a.data_ptr(), gg.data_ptr(), Omega.data_ptr(), Abar.data_ptr(),
                                          ws.data_ptr(), ws.numel(), stream()) == 0
    torch.cuda.synchronize()
    assert relerr(Abar.cpu(), Ad.grad) < TOL


def test_tc_unsupported_M(L):
    lib = L.lib()
    assert lib.gpsa_tc_supported(512) == 1 and lib.gpsa_tc_supported(1024) == 0
    x = torch.zeros(16, device="cuda")
    assert lib.gpsa_quadform_fwd_feat_tc(1024, 128, 1, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 16, stream()) == 3


@pytest.mark.parametrize("kind", ["rbf", "matern12"])
def test_data_layer_engines_agree(L, kind):
    """The whole data layer (forward samples, KL, every gradient) with the tcgen05 engine against the fp32 SIMT engine
    on the same inputs."""
    from gpsa import _ops

    g = torch.Generator().manual_seed(11)
    M, D, Lg, S, N = 64, 2, 24, 2, 1500
    Gt = torch.rand(M, D, generator=g) * 10
    G = torch.rand(S, N, D, generator=g) * 10
    Osq = torch.randn(Lg, M, M, generator=g) * 0.1
    dlt = torch.randn(M, Lg, generator=g)
    eps = torch.randn(S, N, Lg, generator=g)
    Fbar = torch.randn(S, N, Lg, generator=g)
    ls, var = torch.tensor([0.3]), torch.tensor([0.1])
    outs = {}
    for engine in (0, 1):
        _ops.ENGINE["value"] = engine
        try:
            leaves = [t.clone().cuda().requires_grad_() for t in (Gt, ls, var, dlt, Osq, G)]
            F, kl, Lk, Ltril, info = _ops.DataLayer.apply({"kind": _ops.KINDS[kind], "with_kl": True}, *leaves[:5],
                                                          leaves[5], eps.cuda())
            ((F * Fbar.cuda()).sum() + kl).backward()
            outs[engine] = [F.detach().cpu(), kl.detach().cpu()] + [t.grad.cpu() for t in leaves]
        finally:
            _ops.ENGINE["value"] = "auto"
    names = ["F", "kl", "Gtilde", "log_ls", "log_var", "delta", "Omega_sqt", "G"]
    for name, x0, x1 in zip(names, outs[0], outs[1]):
        assert relerr(x1, x0) < 2e-4, name


@pytest.mark.parametrize("Mr,Nc,K,batch,arm,brm,out_mode,split", [
    (1000, 300, 200, 1, 0, 0, 0, 1),      # predictive mean: both operands stored K x rows
    (200, 300, 5000, 1, 1, 0, 0, 0),      # delta-bar: auto split-K with atomics
    (1000, 200, 300, 1, 1, 1, 1, 1),      # A-bar += ...: transposed read-modify-write
    (200, 200, 200, 7, 1, 0, 0, 1),       # batched Omega-bar Omega_sqt with alpha = 2
    (50, 50, 50, 3, 1, 0, 0, 1),
])
def test_gemm_tc_generic(L, Mr, Nc, K, batch, arm, brm, out_mode, split):
    g = torch.Generator().manual_seed(Mr + Nc + K + batch)
    A = torch.randn(batch, Mr, K, generator=g)
    B = torch.randn(batch, Nc, K, generator=g)
    alpha = 2.0 if batch > 1 else 1.0
    ref = alpha * A.double() @ B.double().transpose(1, 2)
    a = (A if arm else A.transpose(1, 2)).contiguous().cuda()
    b = (B if brm else B.transpose(1, 2)).contiguous().cuda()
    lda, ldb = (K if arm else Mr), (K if brm else Nc)
    if out_mode == 1:
        C0 = torch.randn(Nc, Mr, generator=g)
        c = C0.clone().cuda()
        ref = C0.double() + ref[0].T
        ldc, sC = Mr, 0
    else:
        c = torch.full((batch, Mr, Nc), float("nan"), device="cuda")
        ldc, sC = Nc, Mr * Nc
    ws = torch.empty(int(L.lib().gpsa_gemm_tc_ws_bytes(Mr, Nc, K, batch)), dtype=u8, device="cuda")
    rc = L.lib().gpsa_gemm_tc(Mr, Nc, K, batch, a.data_ptr(), lda, Mr * K, arm, b.data_ptr(), ldb, Nc * K, brm,
                              c.data_ptr(), ldc, sC, alpha, out_mode, split, ws.data_ptr(), ws.numel(), stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert relerr(c.cpu().reshape(ref.shape), ref) < 2e-5


def test_vectorised_elementwise_kernels_match_scalar(L):
    """The 128-bit sampling / log-likelihood kernels (L % 4 == 0, L >= 128) against their scalar variants: same data
    layer and the same Gaussian log-likelihood, forward and backward, with the vector path switched off and on."""
    from gpsa import _ops

    g = torch.Generator().manual_seed(23)
    M, D, Lg, S, N = 64, 2, 128, 2, 1100
    Gt = torch.rand(M, D, generator=g) * 10
    G = torch.rand(S, N, D, generator=g) * 10
    Osq = torch.randn(Lg, M, M, generator=g) * 0.1
    dlt = torch.randn(M, Lg, generator=g)
    eps = torch.randn(S, N, Lg, generator=g)
    Y = torch.randn(N, Lg, generator=g)
    ls, var, ln = torch.tensor([0.3]), torch.tensor([0.1]), torch.tensor([-0.2])
    outs = {}
    for off in (1, 0):
        L.lib().gpsa_debug_disable_vec4(off)
        try:
            leaves = [t.clone().cuda().requires_grad_() for t in (Gt, ls, var, dlt, Osq, G, ln)]
            F, kl, _, _, _ = _ops.DataLayer.apply({"kind": _ops.KINDS["rbf"], "with_kl": True}, *leaves[:5], leaves[5],
                                                  eps.cuda())
            ll = _ops.GaussianLL.apply(F, Y.cuda(), leaves[6])
            (kl - ll).backward()
            outs[off] = [F.detach().cpu(), ll.detach().cpu()] + [t.grad.cpu() for t in leaves]
        finally:
            L.lib().gpsa_debug_disable_vec4(0)
    for name, x0, x1 in zip(["F", "ll", "Gtilde", "log_ls", "log_var", "delta", "Omega_sqt", "G", "log_noise"], outs[1],
                            outs[0]):
        # the A-bar product accumulates with atomics (run-to-run order) and G-bar amplifies it through K^-1: 1e-4
        assert relerr(x1, x0) < (1e-4 if name in ("G", "Gtilde", "log_ls", "log_var") else 1e-5), name


def test_quadform_full_size_properties(L):
    """BASELINE.json's C3 shape (M = 200, R = S*N = 128 000, L = 2000) -- the shape bench.py times -- for all three
    products, through size-independent properties: (1) a random sample of entries against float64, (2) partition
    invariance -- a sub-block of rows x genes computed on its own equals the same entries of the full result,
    (3) q2 >= 0 up to round-off (it is a squared norm).  ~7 GB of device memory, a few seconds."""
    from gpsa import _ops

    M, R, Lg = 200, 128000, 2000
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, R, device="cuda", generator=g) * 0.3
    Osq = torch.randn(Lg, M, M, device="cuda", generator=g) * 0.1
    Omega, Ltril, L64, hld, info = _ops.omega_prepare(Osq)
    del L64, Osq, Ltril
    lib = L.lib()
    ws = ws_for(L, M, R, Lg)
    rows = torch.randint(0, R, (96,), device="cuda", generator=g)
    genes = torch.randint(0, Lg, (48,), device="cuda", generator=g)
    ref = torch.einsum("mr,pmk,kr->rp", A[:, rows].double(), Omega[genes].double(), A[:, rows].double())
    r0, r1, p0, p1 = 4096, 5120, 100, 164
    As = A[:, r0:r1].contiguous()
    ws2 = ws_for(L, M, r1 - r0, p1 - p0)
    q2 = torch.full((R, Lg), float("nan"), device="cuda")
    q2s = torch.full((r1 - r0, p1 - p0), float("nan"), device="cuda")
    Os = Omega[p0:p1].contiguous()
    assert lib.gpsa_quadform_fwd_feat_tc(M, R, Lg, A.data_ptr(), Omega.data_ptr(), q2.data_ptr(), ws.data_ptr(),
                                         ws.numel(), stream()) == 0
    assert lib.gpsa_quadform_fwd_feat_tc(M, r1 - r0, p1 - p0, As.data_ptr(), Os.data_ptr(), q2s.data_ptr(),
                                         ws2.data_ptr(), ws2.numel(), stream()) == 0
    torch.cuda.synchronize()
    assert bool(torch.isfinite(q2).all())
    assert float(q2.min()) >= -1e-5 * float(q2.max())                            # (3)
    got = q2[rows][:, genes].double()
    assert float((got - ref).abs().max() / ref.abs().max()) < TOL               # (1)
    full = q2[r0:r1, p0:p1]
    assert float((q2s - full).abs().max() / full.abs().max()) < 1e-5            # (2)
    # no systematic shrinkage from the truncating TMEM accumulator (chains of bounded length, see tc_quadform.cu)
    assert abs(float(((got - ref) / ref).mean())) < 2e-5
    del q2, q2s

    # ---- backward products at the same shape: G = dLoss/dq2 [R, L]
    G = torch.randn(R, Lg, device="cuda", generator=g)
    # A-bar[:, r] = 2 sum_p G[r,p] Omega_p a_r, checked on sampled rows (all 2000 genes contribute to each)
    Abar = torch.zeros(M, R, device="cuda")
    assert lib.gpsa_quadform_bwd_alpha_tc(M, R, Lg, A.data_ptr(), G.data_ptr(), Omega.data_ptr(), Abar.data_ptr(),
                                          ws.data_ptr(), ws.numel(), stream()) == 0
    torch.cuda.synchronize()
    rs = rows[:24]
    refA = torch.zeros(M, rs.numel(), dtype=f64, device="cuda")
    for c0 in range(0, Lg, 250):  # chunked over genes: keeps the float64 temporaries small
        Oc = Omega[c0:c0 + 250].double()
        refA += 2.0 * torch.einsum("rp,pmk,kr->mr", G[rs, c0:c0 + 250].double(), Oc, A[:, rs].double())
    gotA = Abar[:, rs].double()
    assert float((gotA - refA).abs().max() / refA.abs().max()) < TOL
    assert bool(torch.isfinite(Abar).all())
    del Abar
    # Omega-bar_p = sum_r G[r,p] a_r a_r^T, checked on sampled genes (all 128 000 rows contribute to each)
    nf = L.feat_count(M)
    H = torch.full((nf, Lg), float("nan"), device="cuda")
    Obar = torch.empty(Lg, M, M, device="cuda")
    assert lib.gpsa_quadform_bwd_omega_tc(M, R, Lg, A.data_ptr(), G.data_ptr(), H.data_ptr(), ws.data_ptr(), ws.numel(),
                                          stream()) == 0
    assert lib.gpsa_feat_unpack(M, Lg, H.data_ptr(), None, 0.0, None, Obar.data_ptr(), stream()) == 0
    torch.cuda.synchronize()
    assert bool(torch.isfinite(Obar).all())
    Ad = A.double()
    for pgene in genes[:6].tolist():
        refO = (Ad * G[:, pgene].double()[None, :]) @ Ad.T
        gotO = Obar[pgene].double()
        assert float((gotO - refO).abs().max() / refO.abs().max()) < TOL, pgene
