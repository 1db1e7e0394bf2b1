"""Sampling stage of the data layer (-m gpu): the counter-based noise, the fused sampling + likelihood kernel against
the materialised kernels it replaces, and the lazy-handle contract of forward() / loss_fn()."""
import ctypes as C

import numpy as np
import pytest
import torch

from golden_io import Golden, relerr

pytestmark = pytest.mark.gpu
f32, f64 = torch.float32, torch.float64


def _key(v=1234567890123):
    return torch.tensor([v], dtype=torch.int64, device="cuda")


def test_philox_normal_statistics_and_keying():
    from gpsa import _ops

    S, N, L = 4, 3000, 64
    e = _ops.philox_normal(_key(), S, N, L)
    assert e.shape == (S, N, L) and bool(torch.isfinite(e).all())
    x = e.double().flatten()
    n = x.numel()
    assert abs(float(x.mean())) < 5 / np.sqrt(n)
    assert abs(float(x.var()) - 1.0) < 5 * np.sqrt(2.0 / n)
    assert abs(float((x ** 3).mean())) < 5 * np.sqrt(15.0 / n)           # skewness
    assert abs(float((x ** 4).mean()) - 3.0) < 5 * np.sqrt(96.0 / n)      # kurtosis
    assert float(x.abs().max()) < 7.0
    # independent across genes / spots / samples: lag correlations vanish
    for a, b in ((e[:, :, :-1], e[:, :, 1:]), (e[:, :-1], e[:, 1:]), (e[:-1], e[1:])):
        c = float((a.double() * b.double()).mean())
        assert abs(c) < 5 / np.sqrt(a.numel()), c
    # same key -> same draw; another key -> another draw
    assert torch.equal(e, _ops.philox_normal(_key(), S, N, L))
    assert not torch.equal(e, _ops.philox_normal(_key(99), S, N, L))


@pytest.mark.parametrize("lo,hi", [(0, 16), (3, 7), (5, 30), (4, 64), (61, 64)])
def test_philox_normal_is_keyed_by_global_gene_and_sample(lo, hi):
    """A rank that owns genes [lo, hi) (any alignment) or samples [s0, S) draws exactly its slice of the global noise."""
    from gpsa import _ops

    S, N, L = 3, 517, 64
    full = _ops.philox_normal(_key(7), S, N, L)
    part = _ops.philox_normal(_key(7), S, N, hi - lo, gene_off=lo)
    assert torch.equal(part, full[:, :, lo:hi].contiguous())
    tail = _ops.philox_normal(_key(7), S - 1, N, L, samp_off=1)
    assert torch.equal(tail, full[1:].contiguous())


def _toy(Lg, S=3, N=700, seed=0):
    g = torch.Generator().manual_seed(seed)
    mean = torch.randn(S, N, Lg, generator=g)
    q2 = torch.rand(S, N, Lg, generator=g) * 0.5
    kq = torch.rand(S, N, generator=g) * 0.3 + 0.01
    Y = torch.randn(N, Lg, generator=g)
    eps = torch.randn(S, N, Lg, generator=g)
    ln = torch.tensor([-0.4])
    return mean, q2, kq, Y, eps, ln


@pytest.mark.parametrize("Lg", [1, 5, 30, 128, 516, 1030])
@pytest.mark.parametrize("noise", ["eps", "philox"])
def test_fused_stage_matches_materialised_kernels_and_fp64(Lg, noise):
    """SampleNLL (one kernel) == SampleF + GaussianLL (the materialised chain) == float64 torch, value and gradients
    w.r.t. mean, q2, kq and log_noise; scalar (L % 4 != 0) and 128-bit paths, one and several gene chunks per row."""
    from gpsa import _ops

    mean, q2, kq, Y, eps, ln = _toy(Lg)
    S, N, _ = mean.shape
    key = _key(42)
    if noise == "philox":
        eps = _ops.philox_normal(key, S, N, Lg).cpu()
    scale = 0.7  # upstream gradient != 1: exercises the in-place rescale of the saved buffers

    def leaves():
        return [t.clone().cuda().requires_grad_() for t in (mean, q2, kq, ln)]

    # fused
    a = leaves()
    # (the stage overwrites its inputs in place; hand it clones that own their storage, like DataLayerPre's outputs)
    m_in, q_in = a[0] * 1.0, a[1] * 1.0
    nll_f = _ops.SampleNLL.apply({"gene_off": 0}, m_in, q_in, a[2], Y.cuda(), a[3],
                                 eps.cuda() if noise == "eps" else None, key if noise == "philox" else None)
    (scale * nll_f).backward()
    # materialised
    b = leaves()
    F = _ops.SampleF.apply(b[0] * 1.0, b[1] * 1.0, b[2], eps.cuda())
    nll_m = -_ops.GaussianLL.apply(F, Y.cuda(), b[3])
    (scale * nll_m).backward()
    # float64
    c = [t.clone().double().requires_grad_() for t in (mean, q2, kq, ln)]
    var = c[2].unsqueeze(-1) + c[1] + 2e-5
    F64 = c[0] + torch.sqrt(var) * eps.double()
    sigma = torch.exp(c[3]) + 1e-5
    nll64 = -(-0.5 * ((Y.double() - F64) / sigma) ** 2 - torch.log(sigma) - 0.5 * np.log(2 * np.pi)).sum() / S
    (scale * nll64).backward()

    assert abs(float(nll_f) - float(nll64)) <= 2e-6 * abs(float(nll64))
    assert abs(float(nll_f) - float(nll_m)) <= 2e-6 * abs(float(nll64))
    for name, x, y, z in zip(["mean", "q2", "kq", "log_noise"], a, b, c):
        assert relerr(x.grad.cpu(), z.grad) < 2e-5, (name, "fused vs f64")
        assert relerr(x.grad.cpu(), y.grad.cpu()) < 2e-5, (name, "fused vs materialised")


def _run(model, data_dict, g, eps=None, seed=None):
    view_idx, Ns, _, _ = model.create_view_idx_dict(data_dict)
    X = {m: data_dict[m]["spatial_coords"] for m in g.mods}
    if seed is not None:
        torch.manual_seed(seed)
    ret = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S, _eps=eps)
    loss = model.loss_fn(data_dict, ret[3])
    model.zero_grad()
    loss.backward()
    return ret, loss, {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("name", ["c1_shipped", "multimodal", "v3_d3_free"])
def test_model_philox_fused_equals_materialised(name):
    """Default noise (in-kernel Philox, keyed from torch's generator): the fused path and the materialised path
    (eps = gpsa_philox_normal of the same key, plain tensors end to end) give the same loss, gradients and samples."""
    from test_gpu_parity import build

    g = Golden(name)
    model, data_dict = build(g)
    assert model.rng_mode == "philox" and model.fused_ll
    ret_f, loss_f, gr_f = _run(model, data_dict, g, seed=77)
    from gpsa.lazy import LazySamples
    assert all(isinstance(ret_f[3][m], LazySamples) for m in g.mods)
    model.fused_ll = False
    ret_m, loss_m, gr_m = _run(model, data_dict, g, seed=77)
    assert all(torch.is_tensor(ret_m[3][m]) for m in g.mods)
    assert abs(float(loss_f) - float(loss_m)) <= 1e-5 * abs(float(loss_m))
    for m in g.mods:
        assert relerr(ret_f[2][m].detach().cpu(), ret_m[2][m].detach().cpu()) < 1e-5   # recovered vs materialised samples
    for n in gr_m:
        assert relerr(gr_f[n].cpu(), gr_m[n].cpu()) < 2e-4, n
    # another seed -> another draw
    model.fused_ll = True
    _, loss_2, _ = _run(model, data_dict, g, seed=78)
    assert float(loss_2) != float(loss_f)


def test_lazy_handle_contract():
    """forward() returns handles that answer shape queries without work, materialise on first real use (then ARE that
    tensor, with autograd behind it), and are recognised by loss_fn only for the model's own latest forward."""
    from gpsa.lazy import LazySamples
    from test_gpu_parity import build

    g = Golden("c2_matern")
    model, data_dict = build(g)
    view_idx, Ns, _, _ = model.create_view_idx_dict(data_dict)
    X = {m: data_dict[m]["spatial_coords"] for m in g.mods}
    m = g.mods[0]
    torch.manual_seed(3)
    out = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S)
    F = out[3][m]
    assert isinstance(F, LazySamples) and not F.is_materialised
    assert tuple(F.shape) == (g.S, int(Ns[m]), g.Y[m].shape[1]) and F.dim() == 3 and F.device.type == "cuda"
    assert not F.is_materialised
    # a user-side read: materialises, and the tensor carries the autograd graph of the whole model
    t = F.mean(0)
    assert F.is_materialised and torch.is_tensor(t) and t.requires_grad
    loss_a = model.loss_fn(data_dict, out[3])           # handle already materialised -> plain likelihood kernel
    loss_a.backward()
    ga = model.delta_F_dict[m].grad.detach().clone()
    # same seed, untouched handle -> fused kernel; same result
    model.zero_grad()
    torch.manual_seed(3)
    out2 = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S)
    loss_b = model.loss_fn(data_dict, out2[3])
    assert not out2[3][m].is_materialised
    loss_b.backward()
    assert abs(float(loss_a) - float(loss_b)) <= 1e-5 * abs(float(loss_b))
    assert relerr(model.delta_F_dict[m].grad.cpu(), ga.cpu()) < 1e-4
    # after the fused loss the handle still yields the samples (recovered, detached)
    Fr = out2[3][m].detach()
    assert relerr(Fr.cpu(), F.detach().cpu()) < 1e-5
    # a stale handle (older forward) is not taken for the fused path: it is materialised like any tensor... which its
    # consumed buffers no longer allow, so the samples recovered above are what loss_fn sees
    out3 = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S)
    loss_c = model.loss_fn(data_dict, out2[3])
    assert abs(float(loss_c) - float(loss_b)) <= 1e-4 * abs(float(loss_b))
    del out3
    # plain tensors work as in the reference
    loss_d = model.loss_fn(data_dict, {m: F.detach()})
    assert abs(float(loss_d) - float(loss_a)) <= 1e-5 * abs(float(loss_a))


def test_fused_backward_twice_raises():
    from test_gpu_parity import build

    g = Golden("c2_matern")
    model, data_dict = build(g)
    view_idx, Ns, _, _ = model.create_view_idx_dict(data_dict)
    X = {m: data_dict[m]["spatial_coords"] for m in g.mods}
    out = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S)
    loss = model.loss_fn(data_dict, out[3])
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()
