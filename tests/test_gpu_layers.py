"""Layer-level parity (-m gpu): the data layer and the warp layer, each called through its autograd
wrapper around the C ABI with fixed inputs and a random upstream gradient, against a float64 torch
evaluation of the reference expressions (oracle helpers).  Isolates each layer's own arithmetic from
perturbations inherited from upstream layers."""
import math

import numpy as np
import pytest
import torch

from golden_io import Golden, relerr
from oracle import gpsa_oracle as orc

pytestmark = pytest.mark.gpu
f64 = torch.float64


def _data_layer_ref(kind, Gt, ls, var, delta, Osq, G, eps):
    M = Gt.shape[0]
    Kuu = orc.kernel_matrix(kind, Gt, Gt, ls, var) + orc.OFF * torch.eye(M, dtype=f64)
    Lk = torch.linalg.cholesky(Kuu)
    Kuf = orc.kernel_matrix(kind, Gt, G, ls, var)
    Lom = torch.linalg.cholesky(orc.omega_from_sqt(Osq))
    kff = torch.ones(G.shape[:2], dtype=f64) * torch.exp(var)
    zN = torch.zeros(G.shape[1], delta.shape[1], dtype=f64)
    zM = torch.zeros(M, delta.shape[1], dtype=f64)
    mu, Sig = orc.mean_and_var(kff, Kuf, Lk, zN, zM, delta, Lom, materialise=False)
    F = mu + torch.sqrt(Sig.transpose(1, 2)) * eps
    kl = orc.kl_mvn_tril(delta.t(), Lom, torch.zeros(M, dtype=f64), Lk).sum()
    return F, kl


@pytest.mark.parametrize("name", ["c1_shipped", "c1_named", "c2_matern", "v3_d3_free"])
def test_data_layer_against_float64(name):
    from gpsa import _ops

    g = Golden(name)
    mod = g.mods[0]
    kind = g.cfg.kernel_data
    p = g.params
    names = ["Gtilde", "data_kernel_lengthscale", "data_kernel_variance", f"delta_F_dict.{mod}",
             f"Omega_sqt_F_dict.{mod}"]
    G_np = g.out("G_samples", mod)
    eps_np = g.eps["F"][mod]
    gen = torch.Generator().manual_seed(1)
    Fbar = torch.randn(eps_np.shape, generator=gen)

    ref_in = [torch.tensor(p[k], dtype=f64).requires_grad_() for k in names]
    Gd = torch.tensor(G_np, dtype=f64).requires_grad_()
    F_r, kl_r = _data_layer_ref(kind, ref_in[0], ref_in[1], ref_in[2], ref_in[3], ref_in[4], Gd,
                                torch.tensor(eps_np, dtype=f64))
    ((F_r * Fbar.double()).sum() + 0.7 * kl_r).backward()

    cu_in = [torch.tensor(p[k]).cuda().requires_grad_() for k in names]
    Gc = torch.tensor(G_np).cuda().requires_grad_()
    F_c, kl_c, Lk, Ltril, info = _ops.DataLayer.apply(
        {"kind": _ops.KINDS[kind], "with_kl": True}, cu_in[0], cu_in[1], cu_in[2], cu_in[3], cu_in[4], Gc,
        torch.tensor(eps_np).cuda())
    ((F_c * Fbar.cuda()).sum() + 0.7 * kl_c).backward()
    assert int(info.abs().sum()) == 0
    rows = [("F", relerr(F_c.detach().cpu(), F_r.detach())), ("KL", abs(float(kl_c) - float(kl_r)) / abs(float(kl_r)))]
    for k, a, b in zip(names + ["G"], cu_in + [Gc], ref_in + [Gd]):
        rows.append((f"grad.{k}", relerr(a.grad.cpu(), b.grad)))
    print("\n".join(f"{k:40s} {e:.2e}" for k, e in rows))
    # fp32 data with fp64 K^-1 algebra: well inside the 1e-4 the north star asks for, except where
    # the default RBF lengthscale makes K_uu ill-conditioned (c1_named: cond ~ 1e7)
    tol = 2e-3 if name == "c1_named" else 1e-4
    for k, e in rows:
        assert e < tol, (k, e)


@pytest.mark.parametrize("name", ["c1_shipped", "c2_matern", "v3_d2_fixedlist", "v3_d3_free", "v2_d1"])
def test_warp_layer_against_float64(name):
    from gpsa import _ops

    g = Golden(name)
    mod = g.mods[0]
    cfg = g.cfg
    V, D, M = cfg.n_views, cfg.n_spatial_dims, cfg.m_X_per_view
    p = g.params
    names = ["Xtilde", "delta_G_list", "Omega_sqt_G_list", "warp_kernel_lengthscales", "warp_kernel_variances"]
    free = [v for v in range(V) if not cfg.is_fixed(v)]
    vi = cfg.view_idx()[mod]
    gen = torch.Generator().manual_seed(2)
    gbar = {v: (torch.randn(len(vi[v]), D, generator=gen), torch.randn(g.S, len(vi[v]), D, generator=gen))
            for v in free}

    # float64 reference of the warp layer alone, with the reference's index quirks
    r = [torch.tensor(p[k], dtype=f64).requires_grad_() for k in names]
    Xt, dG, Osq, ls, var = r
    Lom = torch.linalg.cholesky(orc.omega_from_sqt(Osq))
    loss_r = 0.0
    for v in free:
        Xv = torch.tensor(g.X[mod][vi[v]], dtype=f64)
        Kuu = orc.kernel_matrix(cfg.kernel_warp, Xt[v], Xt[v], ls[v], var[v]) + orc.OFF * torch.eye(M, dtype=f64)
        Lk = torch.linalg.cholesky(Kuu)
        Kuf = orc.kernel_matrix(cfg.kernel_warp, Xt[v], Xv, ls[v], var[v])
        kff = torch.ones(Xv.shape[0], dtype=f64) * torch.exp(var[v])
        mu, Sig = orc.mean_and_var(kff, Kuf, Lk, Xv, Xt, dG, Lom)
        Gm = mu[v]
        Gs = Gm.unsqueeze(0) + Sig[v * D:v * D + D].t().unsqueeze(0) * torch.tensor(g.eps["G"][v], dtype=f64)
        loss_r = loss_r + (Gm * gbar[v][0].double()).sum() + (Gs * gbar[v][1].double()).sum()
        for j in range(D):
            loss_r = loss_r + 1.3 * orc.kl_mvn_tril(dG[v, :, j], Lom[j * V + v], Xt[v, :, j], Lk)
    loss_r.backward()

    c = [torch.tensor(p[k]).cuda().requires_grad_() for k in names]
    mask = torch.zeros(V * D)
    for v in free:
        for j in range(D):
            mask[j * V + v] = -0.5
    meta = {"kind": _ops.KINDS[cfg.kernel_warp], "V": V, "S": g.S, "free": free, "with_kl": True,
            "kl_mask": mask.cuda()}
    flat = []
    for v in free:
        flat += [torch.tensor(g.X[mod][vi[v]]).cuda(), torch.tensor(g.eps["G"][v]).cuda()]
    outs = _ops.WarpLayer.apply(meta, *c, *flat)
    loss_c = 1.3 * outs[0]
    for k, v in enumerate(free):
        loss_c = loss_c + (outs[4 + 2 * k] * gbar[v][0].cuda()).sum() + (outs[5 + 2 * k] * gbar[v][1].cuda()).sum()
    loss_c.backward()
    rows = [("loss", abs(float(loss_c) - float(loss_r)) / abs(float(loss_r)))]
    for k, a, b in zip(names, c, r):
        rows.append((f"grad.{k}", relerr(a.grad.cpu(), b.grad)))
    print("\n".join(f"{k:40s} {e:.2e}" for k, e in rows))
    tol = 2e-3 if name == "c1_shipped" else 1e-4
    for k, e in rows:
        assert e < tol, (k, e)


def test_graphed_iteration_matches_eager():
    """CUDA-graph replays of forward + loss + backward + Adam EQUAL the eager iterations started from the same model,
    optimizer state and generator seed (the graph-safe CUDA generator reads seed and offset at replay time): losses
    to 1e-5, parameter updates to 1e-3 (what run-to-run atomics order leaves after Adam's normalisation)."""
    import copy

    from golden_io import Golden
    from gpsa.graph import GraphedIteration
    from test_gpu_parity import build

    g = Golden("c2_matern")
    model, data_dict = build(g)
    view_idx, Ns, _, _ = model.create_view_idx_dict(data_dict)
    X = {m: data_dict[m]["spatial_coords"] for m in g.mods}
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, capturable=True)
    with pytest.raises(ValueError):
        GraphedIteration(model, data_dict, torch.optim.Adam(model.parameters(), lr=1e-2), S=g.S)
    it = GraphedIteration(model, data_dict, opt, S=g.S, warmup=2)   # 2 warm-up iterations already stepped the model
    ref, _ = build(g)                                                # eager twin: same parameters, same Adam state
    ref.load_state_dict(model.state_dict())
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2, capturable=True)
    opt_ref.load_state_dict(copy.deepcopy(opt.state_dict()))
    p0 = [p.detach().clone() for p in model.parameters()]
    for k in range(3):
        torch.manual_seed(500 + k)
        lg = float(it.step())
        torch.manual_seed(500 + k)
        out = ref.forward(X, view_idx=view_idx, Ns=Ns, S=g.S)
        le = ref.loss_fn(data_dict, out[3])
        opt_ref.zero_grad(set_to_none=True)
        le.backward()
        opt_ref.step()
        assert np.isfinite(lg) and abs(lg - float(le)) <= 1e-5 * abs(float(le)), (k, lg, float(le))
    for a, b, c in zip(model.parameters(), ref.parameters(), p0):
        da, db = (a - c).detach(), (b - c).detach()
        assert torch.isfinite(a).all()
        assert float((da - db).abs().max()) <= 1e-3 * max(float(db.abs().max()), 1e-12)


def test_iteration_survives_poisoned_allocator_blocks():
    """The views of the warp layer run on side streams.  Re-run the same iteration after filling the caching
    allocator's free blocks with NaN: any kernel that reads a buffer before its zero-fill / copy was ordered in
    front of it (a fork placed too early) turns the gradients into NaN or changes them."""
    from test_gpu_parity import build, run

    g = Golden("v3_d3_free")
    model, data_dict = build(g)
    _, loss0 = run(g, model, data_dict)
    ref = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    for rep in range(5):
        poison = [torch.full((1 << 22,), float("nan"), device="cuda") for _ in range(8)]
        small = [torch.full((n,), float("nan"), device="cuda") for n in (64, 1000, 40000, 200000) for _ in range(8)]
        del poison, small
        _, loss = run(g, model, data_dict)
        assert torch.isfinite(loss)
        assert abs(float(loss) - float(loss0)) <= 1e-6 * abs(float(loss0))
        for n, p in model.named_parameters():
            assert torch.isfinite(p.grad).all(), n
            assert relerr(p.grad.cpu(), ref[n].cpu()) < 1e-5, n
