"""-m gpu: the pieces around the ELBO iteration -- LMC loadings (materialised and fused with the likelihood), the
one-launch Adam step against torch.optim.Adam, and the device-side k-means initialisation."""
import numpy as np
import pytest
import torch

from golden_io import relerr

pytestmark = pytest.mark.gpu
f64 = torch.float64


@pytest.mark.parametrize("S,N,L,P", [(3, 400, 5, 37), (2, 1300, 12, 700), (4, 257, 32, 1025), (1, 50, 1, 1)])
def test_lmc_ops_match_fp64(S, N, L, P):
    from gpsa import _ops

    g = torch.Generator().manual_seed(S * 100 + L)
    Fl = torch.randn(S, N, L, generator=g)
    W = torch.randn(L, P, generator=g)
    Y = torch.randn(N, P, generator=g)
    ln = torch.tensor([-0.3])
    scale = 1.7
    # float64 reference
    c = [t.clone().double().requires_grad_() for t in (Fl, W, ln)]
    Fo = c[0] @ c[1]
    sigma = torch.exp(c[2]) + 1e-5
    nll64 = -(-0.5 * ((Y.double() - Fo) / sigma) ** 2 - torch.log(sigma) - 0.5 * np.log(2 * np.pi)).sum() / S
    (scale * nll64).backward()
    # materialised: LMCObserve + GaussianLL
    a = [t.clone().cuda().requires_grad_() for t in (Fl, W, ln)]
    Fo_c = _ops.LMCObserve.apply(a[0], a[1])
    assert relerr(Fo_c.detach().cpu(), Fo.detach()) < 1e-5
    nll_m = -_ops.GaussianLL.apply(Fo_c, Y.cuda(), a[2])
    (scale * nll_m).backward()
    # fused
    assert _ops.lmc_fused_supported(L)
    b = [t.clone().cuda().requires_grad_() for t in (Fl, W, ln)]
    nll_f = _ops.LMCNLL.apply(b[0], b[1], Y.cuda(), b[2])
    (scale * nll_f).backward()
    for nll in (nll_m, nll_f):
        assert abs(float(nll) - float(nll64)) <= 3e-6 * abs(float(nll64))
    for name, x, y, z in zip(["F_lat", "W", "log_noise"], a, b, c):
        assert relerr(x.grad.cpu(), z.grad) < 3e-5, (name, "materialised")
        assert relerr(y.grad.cpu(), z.grad) < 3e-5, (name, "fused")


def test_lmc_model_fused_equals_materialised():
    """Golden LMC case end to end: loss_fn on the lazy handle (fused LMC likelihood) vs plain tensors."""
    from golden_io import Golden
    from test_gpu_parity import build, run

    g = Golden("lmc")
    res = {}
    for fused in (True, False):
        model, data_dict = build(g)
        model.fused_ll = fused
        ret, loss = run(g, model, data_dict)
        res[fused] = (float(loss), {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None},
                      ret[3][g.mods[0]].detach().cpu())
    assert abs(res[True][0] - res[False][0]) <= 1e-5 * abs(res[False][0])
    for n in res[False][1]:
        assert relerr(res[True][1][n], res[False][1][n]) < 1e-4, n
    assert relerr(res[True][2], res[False][2]) < 1e-5


@pytest.mark.parametrize("shapes", [[(7,), (33, 5), (4, 4, 4)], [(100000,), (3,), (257, 129)] + [(5, 5)] * 30])
def test_adam_matches_torch(shapes):
    """gpsa.optim.Adam == torch.optim.Adam step for step (same rule, same state), incl. > GPSA_ADAM_MAX_TENSORS tensors,
    a parameter without gradient and odd sizes / alignments."""
    from gpsa.optim import Adam

    g = torch.Generator().manual_seed(len(shapes))
    p0 = [torch.randn(*s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(t.clone().cuda()) for t in p0]
    pb = [torch.nn.Parameter(t.clone().cuda()) for t in p0]
    oa = Adam(pa, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    ob = torch.optim.Adam(pb, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    for step in range(7):
        for k, (x, y) in enumerate(zip(pa, pb)):
            if k == 1 and step < 2:
                x.grad = y.grad = None        # no gradient for this tensor in the first steps
                continue
            gr = torch.randn(*x.shape, generator=g) * (10.0 ** (k % 3 - 1))
            x.grad, y.grad = gr.cuda(), gr.cuda().clone()
        oa.step()
        ob.step()
    for x, y in zip(pa, pb):
        assert relerr(x.detach().cpu(), y.detach().cpu()) < 2e-6
    sa, sb = oa.state[pa[0]], ob.state[pb[0]]
    assert relerr(sa["exp_avg"].cpu(), sb["exp_avg"].cpu()) < 1e-6
    assert relerr(sa["exp_avg_sq"].cpu(), sb["exp_avg_sq"].cpu()) < 1e-6


def test_adam_in_cuda_graph():
    """The step counter lives on the device: a captured step replays as successive Adam steps."""
    from gpsa.optim import Adam

    p = torch.nn.Parameter(torch.ones(1000, device="cuda"))
    q = torch.nn.Parameter(torch.ones(1000, device="cuda"))
    gr = torch.linspace(-1, 1, 1000, device="cuda")
    p.grad, q.grad = gr.clone(), gr.clone()
    oa, ob = Adam([p], lr=1e-2), torch.optim.Adam([q], lr=1e-2)
    oa.step()
    ob.step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        oa.step()
    ob.step()   # the capture itself does not execute: replays below are steps 2, 3, 4
    for _ in range(3):
        graph.replay()
    ob.step()
    ob.step()
    torch.cuda.synchronize()
    assert relerr(p.detach().cpu(), q.detach().cpu()) < 2e-6


@pytest.mark.parametrize("D,K,N", [(2, 50, 5000), (3, 256, 20000), (1, 7, 300)])
def test_kmeans_gpu_quality(D, K, N):
    """Lloyd's iterations on the GPU reach an inertia within 10 % of sklearn's KMeans (k-means++ start, 1 init) and
    every centre is the mean of the points assigned to it (the fixed-point property of Lloyd's algorithm)."""
    from sklearn.cluster import KMeans

    from gpsa.util.util import kmeans_gpu

    rng = np.random.default_rng(D * 10 + K)
    X = (rng.uniform(0, 10, (N, D)) + 0.3 * rng.standard_normal((N, D))).astype(np.float32)
    C, inertia = kmeans_gpu(torch.from_numpy(X), K, iters=40, seed=1)
    assert C.shape == (K, D) and bool(torch.isfinite(C).all())
    ref = KMeans(n_clusters=K, n_init=1, random_state=0).fit(X)
    assert inertia <= 1.10 * ref.inertia_, (inertia, ref.inertia_)
    Cn = C.cpu().numpy().astype(np.float64)
    d2 = ((X[:, None, :].astype(np.float64) - Cn[None]) ** 2).sum(-1)
    lab = d2.argmin(1)
    assert abs(d2.min(1).sum() - inertia) <= 1e-4 * inertia
    moved = 0.0
    for k in range(K):
        if (lab == k).any():
            moved = max(moved, float(np.abs(X[lab == k].mean(0) - Cn[k]).max()))
    assert moved < 0.15  # converged to (near) a fixed point after 40 rounds


def test_model_uses_gpu_kmeans_for_large_inputs(monkeypatch):
    import gpsa
    from gpsa.models import vgpsa as vg

    monkeypatch.setattr(vg, "KMEANS_HOST_MAX", 500)
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 10, (1200, 2)).astype(np.float32)
    Y = rng.standard_normal((1200, 3)).astype(np.float32)
    dd = {"expression": {"spatial_coords": torch.from_numpy(X), "outputs": torch.from_numpy(Y), "n_samples_list": [600, 600]}}
    np.random.seed(0)
    model = gpsa.VariationalGPSA(dd, m_X_per_view=20, m_G=30, data_init=True, n_latent_gps={"expression": None},
                                 fixed_view_idx=0)
    assert model.Xtilde.shape == (2, 20, 2) and model.Gtilde.shape == (30, 2)
    for v in range(2):  # every inducing location sits inside its view's point cloud
        lo, hi = X[600 * v:600 * (v + 1)].min(0), X[600 * v:600 * (v + 1)].max(0)
        z = model.Xtilde[v].detach().numpy()
        assert (z >= lo - 1e-3).all() and (z <= hi + 1e-3).all()
    assert model.Gtilde.dtype == torch.float32 and model.Xtilde.device.type == "cpu"
