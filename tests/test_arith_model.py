"""CPU model of the tcgen05 engine's arithmetic (DESIGN.md 3 "Precision", 4.4): the bf16 hi/lo split and the three
passes hi*hi + lo*hi + hi*lo, emulated with torch on the host and compared with float64.  These tests pin the numbers
DESIGN.md quotes and the decision NOT to move the ill-conditioned Omega^-1 Omega_sqt products onto that engine; the
kernels themselves are held to float64 by the -m gpu tests (tests/test_gpu_tc.py, tests/test_gpu_fullsize.py)."""
import pytest
import torch

from oracle import gpsa_oracle as orc

f32, f64 = torch.float32, torch.float64


def split(x):
    """x = hi + lo (+ residual): both bf16, round to nearest -- csrc/tc_common.cuh split_pair / split_one."""
    hi = x.to(torch.bfloat16).to(f32)
    lo = (x - hi).to(torch.bfloat16).to(f32)
    return hi, lo


def mm3(a, b):
    """a @ b in the engine's three passes; products and sums exact (float64), i.e. the error of the split alone."""
    ah, al = split(a)
    bh, bl = split(b)
    return ah.double() @ bh.double() + al.double() @ bh.double() + ah.double() @ bl.double()


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def test_split_keeps_sixteen_bits():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 16, generator=g) * torch.exp(4 * torch.randn(1 << 16, generator=g))
    hi, lo = split(x)
    res = (x.double() - hi.double() - lo.double()).abs()
    assert bool((res <= 2.0 ** -17 * x.abs().double()).all())          # two 8-bit mantissas
    assert bool((lo.abs().double() <= 2.0 ** -8 * x.abs().double()).all())
    # exactly representable values split exactly
    y = torch.tensor([1.0, -0.5, 3.0, 0.0, 2.0 ** -20])
    yh, yl = split(y)
    assert torch.equal(yh, y) and torch.equal(yl, torch.zeros_like(y))


@pytest.mark.parametrize("M,R,L", [(48, 64, 8), (200, 32, 4)])
def test_three_pass_quadratic_form_is_fp32_class(M, R, L):
    """q2[r,p] = a_r^T Omega_p a_r through the implicit-feature form (Phi = a_i a_j split, W = Omega split): the split
    error stays ~1e-5 of the result's scale at the conditioning of the reference's initialisation."""
    g = torch.Generator().manual_seed(M)
    Z = torch.rand(M, 2, generator=g) * 10
    X = torch.rand(R, 2, generator=g) * 10
    ls, var = torch.tensor([1.0]).log(), torch.tensor([0.0])
    Kuu = orc.kernel_matrix("rbf", Z.double(), Z.double(), ls.double(), var.double()) + 1e-5 * torch.eye(M, dtype=f64)
    Kuf = orc.kernel_matrix("rbf", Z.double(), X.double(), ls.double(), var.double())
    A = torch.linalg.solve(Kuu, Kuf).float()                                   # [M, R]
    Osq = 0.1 * torch.randn(L, M, M, generator=g)
    Om = (Osq.double() @ Osq.double().transpose(1, 2) + 1e-5 * torch.eye(M, dtype=f64)).float()
    truth = torch.einsum("mr,pmn,nr->rp", A.double(), Om.double(), A.double())
    iu = torch.triu_indices(M, M)
    c = torch.where(iu[0] == iu[1], 1.0, 2.0)
    Phi = A[iu[0]] * A[iu[1]]                                                   # [NF, R], fp32 product like the generator warps
    W = (Om[:, iu[0], iu[1]] * c).T.contiguous()                                # [NF, L]
    q3 = mm3(Phi.T.contiguous(), W)
    assert relerr(q3, truth) < 2e-5
    # ... and one bf16 pass alone is nowhere near: the reason for the split
    q1 = Phi.T.to(torch.bfloat16).double() @ W.to(torch.bfloat16).double()
    assert relerr(q1, truth) > 1e-3


def test_omega_inverse_products_do_not_belong_on_the_split_engine():
    """Osq_bar = 2 (Obar + c Omega^-1) Osq at the benchmarked M = 200.  With the fp32 factor inverse L^-1
    (cond ~ 1e3) the two products L^-1 Osq and L^-T Y in fp32 are closer to float64 than the reference's own fp32
    autograd; the same products with the 2^-17 split error, or Omega^-1 = L^-T L^-1 folded into Obar, give that margin
    away.  Geometric means over three draws (single draws scatter with the conditioning of the draw)."""
    M, B = 200, 3
    errs = {"ref": [], "ship": [], "split": [], "fold": []}
    for seed in range(3):
        g = torch.Generator().manual_seed(1000 * seed + M + B)
        Osq = torch.randn(B, M, M, generator=g) * 0.1
        Obar = torch.randn(B, M, M, generator=g) * 1e-3
        Obar = Obar + Obar.transpose(1, 2)
        coef = torch.full((B,), -0.5)

        def ref(dtype):
            od = Osq.to(dtype).requires_grad_()
            Om = orc.omega_from_sqt(od)
            hld = torch.log(torch.diagonal(torch.linalg.cholesky(Om), dim1=1, dim2=2)).sum(1)
            ((Obar.to(dtype) * Om).sum() + (2 * coef.to(dtype) * hld).sum()).backward()
            return Om.detach(), od.grad.detach()

        Om64, g64 = ref(f64)
        _, g32 = ref(f32)
        Lf = torch.linalg.cholesky(Om64.float())        # Omega accumulated in fp64, rounded once, fp32 factorisation
        Linv = torch.linalg.solve_triangular(Lf, torch.eye(M).expand(B, M, M), upper=False)
        cB = coef[:, None, None]
        LinvT = Linv.transpose(1, 2).contiguous()
        shipped = 2 * (Obar @ Osq) + 2 * cB * (LinvT @ (Linv @ Osq))
        on_split = (2 * mm3(Obar, Osq) + 2 * cB * mm3(LinvT, mm3(Linv, Osq).float())).float()
        folded = (2 * mm3(Obar + cB * mm3(LinvT, Linv).float(), Osq)).float()
        for k, v in (("ref", g32), ("ship", shipped), ("split", on_split), ("fold", folded)):
            errs[k].append(relerr(v, g64))
    gm = {k: float(torch.tensor(v).log().mean().exp()) for k, v in errs.items()}
    assert gm["ship"] < gm["ref"]             # the shipped arithmetic beats the reference's fp32 ...
    assert gm["ship"] < 2e-3
    assert gm["split"] > 1.5 * gm["ship"]     # ... the rejected variants lose accuracy,
    assert gm["fold"] > gm["split"]           # folding Omega^-1 the most


def _trunc32(x):
    """float64 -> float32 rounded TOWARD ZERO (the TMEM accumulator's behaviour measured in DESIGN.md 3)."""
    import numpy as np

    f = x.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x)
    f[over] = np.nextafter(f[over], np.float32(0))
    return f


def test_two_level_accumulation_bounds_the_truncation_bias():
    """DESIGN.md 3, 'The TMEM accumulator truncates': one chain of 24 000 dependent truncating adds (an Omega-bar tile at
    C3: 2000 K blocks x 12 MMAs) comes out ~5e-4 low; chains of KB_CHAIN x 12 = 1536 instructions combined in fp32 with
    round-to-nearest (what the epilogue warps do between TMEM columns [0,256) and [256,512)) stay below 6e-5."""
    import numpy as np

    rng = np.random.default_rng(0)
    n, chain, lanes = 24000, 1536, 512
    p = (rng.standard_normal((n, lanes)) * 0.3 + 1.0).astype(np.float32)   # partial products with a common sign
    exact = p.astype(np.float64).sum(0)
    one = np.zeros(lanes, np.float32)
    acc = np.zeros(lanes, np.float32)       # second-level accumulator (round to nearest)
    cur = np.zeros(lanes, np.float32)       # chain accumulator (truncating)
    for k in range(n):
        one = _trunc32(one.astype(np.float64) + p[k])
        cur = _trunc32(cur.astype(np.float64) + p[k])
        if (k + 1) % chain == 0 or k == n - 1:
            acc = (acc + cur).astype(np.float32)
            cur[:] = 0
    bias_one = float(np.mean((one - exact) / exact))
    bias_two = float(np.mean((acc - exact) / exact))
    assert -2e-3 < bias_one < -2e-4
    assert abs(bias_two) < 6e-5 and abs(bias_two) < 0.1 * abs(bias_one)
