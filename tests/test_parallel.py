"""Gene sharding (gpsa.parallel): host logic under gloo with world_size 2 on CPU, and -- marked gpu -- the sharded
ELBO iteration against the unsharded one on the same parameters and noise (two ranks sharing cuda:0 over gloo;
the data path has no collective, the gradient all-reduce is backend-agnostic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_io import Golden, relerr
from test_host_api import _model_from_golden


def _manager():
    """The result dict's server process must be SPAWNED: a fork of this process inherits its CUDA tensors, and the
    child's garbage collector freeing one of them (the allocator then records events for blocks used on several
    streams) aborts with "CUDA error: initialization error" -- seen intermittently on the GPU box."""
    return mp.get_context("spawn").Manager()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_gene_range_partitions_everything():
    from gpsa.parallel import gene_range

    for n in (1, 5, 30, 2000, 2001, 5000):
        for world in (1, 2, 3, 4, 8):
            spans = [gene_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        gene_range(10, 2, 2)


def test_shard_state_and_data():
    from gpsa.parallel import gene_range, shard_data_dict, shard_state_dict

    g = Golden("c1_shipped")
    model, data_dict = _model_from_golden(g)
    sd = model.state_dict()
    P = data_dict["expression"]["outputs"].shape[1]
    parts = [shard_state_dict(sd, 4, r) for r in range(4)]
    assert torch.equal(torch.cat([p["Omega_sqt_F_dict.expression"] for p in parts], 0), sd["Omega_sqt_F_dict.expression"])
    assert torch.equal(torch.cat([p["delta_F_dict.expression"] for p in parts], 1), sd["delta_F_dict.expression"])
    assert all(torch.equal(p["Xtilde"], sd["Xtilde"]) for p in parts)
    dd = shard_data_dict(data_dict, 4, 3)
    lo, hi = gene_range(P, 4, 3)
    assert torch.equal(dd["expression"]["outputs"], data_dict["expression"]["outputs"][:, lo:hi])
    assert dd["expression"]["spatial_coords"] is data_dict["expression"]["spatial_coords"]
    # LMC: the latent GPs are replicated, the loadings are sharded by output column (needs to be told which modalities)
    g2 = Golden("lmc")
    m2, dd2 = _model_from_golden(g2)
    sd2 = m2.state_dict()
    with pytest.raises(ValueError):
        shard_state_dict(sd2, 2, 0)
    mod = g2.mods[0]
    parts2 = [shard_state_dict(sd2, 2, r, n_latent_gps=g2.n_latent) for r in range(2)]
    assert torch.equal(torch.cat([p[f"W_dict.{mod}"] for p in parts2], 1), sd2[f"W_dict.{mod}"])
    assert all(torch.equal(p[f"Omega_sqt_F_dict.{mod}"], sd2[f"Omega_sqt_F_dict.{mod}"]) for p in parts2)
    assert all(torch.equal(p[f"delta_F_dict.{mod}"], sd2[f"delta_F_dict.{mod}"]) for p in parts2)


def _cpu_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpsa.parallel import SHARED, GeneSharding, shard_data_dict, shard_state_dict

        g = Golden("c1_shipped")
        full, data_dict = _model_from_golden(g)
        local_dd = shard_data_dict(data_dict, world, rank)
        g_local, _ = None, None
        import gpsa

        np.random.seed(0)
        torch.manual_seed(0)
        model = gpsa.VariationalGPSA(local_dd, n_spatial_dims=2, m_X_per_view=g.cfg.m_X_per_view, m_G=g.cfg.m_G,
                                     data_init=True, n_latent_gps=g.n_latent, fixed_view_idx=g.fixed)
        model.load_state_dict(shard_state_dict(full.state_dict(), world, rank))
        sh = GeneSharding(model, world, rank)
        assert model._kl_G_scale == 1.0 / world
        named = dict(model.named_parameters())
        ok = True
        for it in range(2):  # the views must survive a second iteration
            sh.zero_grad()
            # a stand-in loss on CPU (the ELBO needs the GPU): rank-dependent weights on shared and local parameters
            loss = sum((rank + 1.0) * (p ** 2).sum() for n, p in named.items() if n in SHARED)
            loss = loss + 3.0 * (named["delta_F_dict.expression"] ** 2).sum()
            loss.backward()
            total = sh.allreduce(loss)
            wsum = sum(r + 1.0 for r in range(world))
            for n, p in named.items():
                if n in SHARED:
                    ok &= torch.allclose(p.grad, 2 * wsum * p.detach(), rtol=1e-6)
                    ok &= p.grad.data_ptr() >= sh.flat.data_ptr()
            ok &= torch.allclose(named["delta_F_dict.expression"].grad, 6.0 * named["delta_F_dict.expression"].detach())
            losses = [torch.zeros(()) for _ in range(world)]
            dist.all_gather(losses, loss.detach())
            ok &= bool(torch.isclose(total, sum(losses), rtol=1e-6))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shared_gradient_allreduce_gloo_world2():
    world = 2
    mgr = _manager()
    ret = mgr.dict()
    mp.spawn(_cpu_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _cpu_hybrid_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpsa
        from gpsa.parallel import SHARED, HybridSharding, gene_range, shard_data_dict, shard_state_dict

        Ws = 2
        Wg, gi, si = world // Ws, rank // Ws, rank % Ws
        g = Golden("c1_shipped")
        full, data_dict = _model_from_golden(g)
        local_dd = shard_data_dict(data_dict, Wg, gi)
        np.random.seed(rank)   # replicas of a slice start DIFFERENT: the sharder must make them equal
        torch.manual_seed(rank)
        model = gpsa.VariationalGPSA(local_dd, n_spatial_dims=2, m_X_per_view=g.cfg.m_X_per_view, m_G=g.cfg.m_G,
                                     data_init=True, n_latent_gps=g.n_latent, fixed_view_idx=g.fixed)
        if si == 0:
            model.load_state_dict(shard_state_dict(full.state_dict(), Wg, gi))
        sh = HybridSharding(model, world, rank, Ws)
        mod = "expression"
        P = full.state_dict()[f"delta_F_dict.{mod}"].shape[1]
        ok = model._kl_G_scale == 1.0 / world and model._kl_F_scale[mod] == 1.0 / Ws
        ok &= model._sample_shard == (si, Ws) and model._gene_off[mod] == gene_range(P, Wg, gi)[0]
        named = dict(model.named_parameters())
        want = shard_state_dict(full.state_dict(), Wg, gi)
        ok &= torch.equal(named[f"delta_F_dict.{mod}"].detach(), want[f"delta_F_dict.{mod}"])        # slice broadcast
        ok &= torch.equal(named["Xtilde"].detach(), full.state_dict()["Xtilde"])                    # world broadcast
        for it in range(2):
            sh.zero_grad()
            loss = sum((rank + 1.0) * (p ** 2).sum() for n, p in named.items() if n in SHARED)
            loss = loss + (rank + 1.0) * (named[f"delta_F_dict.{mod}"] ** 2).sum()
            loss.backward()
            total = sh.allreduce(loss)
            wsum = sum(r + 1.0 for r in range(world))
            lsum = sum(gi * Ws + k + 1.0 for k in range(Ws))   # the replicas of my slice only
            for n, p in named.items():
                if n in SHARED:
                    ok &= torch.allclose(p.grad, 2 * wsum * p.detach(), rtol=1e-6)
            d = named[f"delta_F_dict.{mod}"]
            ok &= torch.allclose(d.grad, 2 * lsum * d.detach(), rtol=1e-6)
            losses = [torch.zeros(()) for _ in range(world)]
            dist.all_gather(losses, loss.detach())
            ok &= bool(torch.isclose(total, sum(losses), rtol=1e-6))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_hybrid_grid_allreduce_gloo_world4():
    """2 gene slices x 2 sample groups: parameters start equal inside a slice (and shared ones everywhere), gene-local
    gradients sum over the slice's replicas only, shared ones over the world."""
    world = 4
    mgr = _manager()
    ret = mgr.dict()
    mp.spawn(_cpu_hybrid_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}


# --------------------------------------------------------------------------------------------------
def _gpu_worker(rank, world, port, name, out, inject=True, mode="gene"):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpsa
        from gpsa.parallel import (GeneSharding, HybridSharding, SampleSharding, gene_range, shard_data_dict,
                                   shard_state_dict)

        # gene-group coordinates: GeneSharding = (world, rank); hybrid 2 x 2 = (world / 2, rank // 2); samples = (1, 0)
        Ws = 2 if mode == "hybrid" else 1
        gw, gr = (world // Ws, rank // Ws) if mode in ("gene", "hybrid") else (1, 0)

        g = Golden(name)
        full, data_dict = _model_from_golden(g)
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in g.params.items() if k not in g.fixed_params}
        full.load_state_dict(sd, strict=True)
        local_dd = shard_data_dict(data_dict, gw, gr)
        kern = {"rbf": gpsa.rbf_kernel, "matern12": gpsa.matern12_kernel, "matern32": gpsa.matern32_kernel}
        np.random.seed(0)
        torch.manual_seed(0)
        model = gpsa.VariationalGPSA(local_dd, n_spatial_dims=g.cfg.n_spatial_dims, m_X_per_view=g.cfg.m_X_per_view,
                                     m_G=g.cfg.m_G, data_init=True, n_latent_gps=g.n_latent,
                                     kernel_func_warp=kern[g.cfg.kernel_warp], kernel_func_data=kern[g.cfg.kernel_data],
                                     fixed_view_idx=g.fixed)
        model.load_state_dict(shard_state_dict(full.state_dict(), gw, gr, n_latent_gps=g.n_latent))
        model = model.to("cuda")
        local_dd = {m: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for m, d in local_dd.items()}
        sh = (GeneSharding(model, world, rank) if mode == "gene" else
              SampleSharding(model, world, rank) if mode == "sample" else HybridSharding(model, world, rank, Ws))
        view_idx, Ns, _, _ = model.create_view_idx_dict(local_dd)
        P = g.Y[g.mods[0]].shape[1]
        lo, hi = gene_range(P, gw, gr)
        lmc = any(v is not None for v in g.n_latent.values())
        if mode == "sample" or lmc:   # every rank holds every latent output: full noise (the model keeps its samples)
            lo_e, hi_e = 0, None
        else:
            lo_e, hi_e = lo, hi
        eps = {"G": {v: torch.from_numpy(e) for v, e in g.eps["G"].items()},  # identical warp noise on every rank
               "F": {m: torch.from_numpy(e[:, :, lo_e:hi_e].copy()) for m, e in g.eps["F"].items()}, "F_test": {}}
        X = {m: local_dd[m]["spatial_coords"] for m in g.mods}
        if inject:
            ret = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S, _eps=eps)
        else:  # default noise: every rank seeds alike; eps_F comes from the in-kernel generator keyed by GLOBAL gene
            torch.manual_seed(4242)
            ret = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S)
        loss = model.loss_fn(local_dd, ret[3])
        sh.zero_grad()
        loss.backward()
        total = sh.allreduce(loss)
        res = {"loss": float(total), "lo": lo, "hi": hi}
        for n, p in model.named_parameters():
            res[n] = p.grad.detach().cpu().numpy()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_shipped", "c2_matern"])
def test_sharded_iteration_equals_unsharded(name):
    from test_gpu_parity import build, run

    g = Golden(name)
    model, data_dict = build(g)
    _, loss = run(g, model, data_dict)
    ref = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
    world = 2
    mgr = _manager()
    out = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    out = dict(out)
    assert abs(out[0]["loss"] - float(loss)) <= 2e-5 * abs(float(loss))
    assert out[0]["loss"] == out[1]["loss"]
    mod = g.mods[0]
    for n, gref in ref.items():
        if n == f"Omega_sqt_F_dict.{mod}":
            got = np.concatenate([out[r][n] for r in range(world)], 0)
        elif n == f"delta_F_dict.{mod}":
            got = np.concatenate([out[r][n] for r in range(world)], 1)
        else:
            got = out[0][n]
            assert np.array_equal(out[0][n], out[1][n]), n  # all-reduced: bitwise identical on every rank
        assert relerr(got, gref) < 5e-5, n


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_shipped"])
def test_sharded_iteration_equals_unsharded_without_injected_noise(name):
    """World-size-invariant Monte-Carlo draw: with the default in-kernel noise (no injected eps) the gene-sharded
    iteration at world 2 reproduces the unsharded loss and gradients for the same torch seed."""
    from test_gpu_parity import build

    g = Golden(name)
    model, data_dict = build(g)
    view_idx, Ns, _, _ = model.create_view_idx_dict(data_dict)
    X = {m: data_dict[m]["spatial_coords"] for m in g.mods}
    torch.manual_seed(4242)
    ret = model.forward(X, view_idx=view_idx, Ns=Ns, S=g.S)
    loss = model.loss_fn(data_dict, ret[3])
    model.zero_grad()
    loss.backward()
    ref = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
    world = 2
    mgr = _manager()
    out = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), name, out, False), nprocs=world, join=True)
    out = dict(out)
    assert abs(out[0]["loss"] - float(loss)) <= 2e-5 * abs(float(loss))
    mod = g.mods[0]
    for n, gref in ref.items():
        if n == f"Omega_sqt_F_dict.{mod}":
            got = np.concatenate([out[r][n] for r in range(world)], 0)
        elif n == f"delta_F_dict.{mod}":
            got = np.concatenate([out[r][n] for r in range(world)], 1)
        else:
            got = out[0][n]
        assert relerr(got, gref) <= 5e-5, n


def _gather(out, world, n, mod, how):
    if how == "rows":
        return np.concatenate([out[r][n] for r in range(world)], 0)
    if how == "cols":
        return np.concatenate([out[r][n] for r in range(world)], 1)
    assert np.array_equal(out[0][n], out[1][n]), n  # all-reduced: bitwise identical on every rank
    return out[0][n]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_shipped", "v3_d3_free"])
def test_sample_sharded_iteration_equals_unsharded(name):
    """Monte-Carlo-sample sharding (fewer genes than GPUs): each of 2 ranks evaluates its share of the S = 5 samples,
    every gradient is all-reduced; loss and gradients equal the unsharded iteration on the same noise."""
    from test_gpu_parity import build, run

    g = Golden(name)
    model, data_dict = build(g)
    _, loss = run(g, model, data_dict)
    ref = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
    world = 2
    mgr = _manager()
    out = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), name, out, True, "sample"), nprocs=world, join=True)
    out = dict(out)
    assert abs(out[0]["loss"] - float(loss)) <= 2e-5 * abs(float(loss))
    for n, gref in ref.items():
        assert relerr(_gather(out, world, n, None, "same"), gref) < 1e-4, n


@pytest.mark.gpu
def test_lmc_gene_sharded_iteration_equals_unsharded():
    """Gene sharding with LMC loadings: the observed outputs (columns of W and of the data) are split over 2 ranks,
    the latent GPs are replicated and their gradients all-reduced."""
    from test_gpu_parity import build, run

    g = Golden("lmc")
    model, data_dict = build(g)
    _, loss = run(g, model, data_dict)
    ref = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
    world = 2
    mgr = _manager()
    out = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), "lmc", out, True, "gene"), nprocs=world, join=True)
    out = dict(out)
    assert abs(out[0]["loss"] - float(loss)) <= 2e-5 * abs(float(loss))
    mod = g.mods[0]
    for n, gref in ref.items():
        how = "cols" if n == f"W_dict.{mod}" else "same"
        assert relerr(_gather(out, world, n, mod, how), gref) < 1e-4, n


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_shipped", "v3_d3_free"])
def test_hybrid_sharded_iteration_equals_unsharded(name):
    """Genes x samples on a 2 x 2 grid of ranks (4 processes sharing cuda:0 over gloo): gene-local gradients are
    all-reduced inside the slice group, shared ones over the world; loss and gradients equal the unsharded iteration."""
    from test_gpu_parity import build, run

    g = Golden(name)
    model, data_dict = build(g)
    _, loss = run(g, model, data_dict)
    ref = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
    world, Ws = 4, 2
    mgr = _manager()
    out = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), name, out, True, "hybrid"), nprocs=world, join=True)
    out = dict(out)
    assert abs(out[0]["loss"] - float(loss)) <= 2e-5 * abs(float(loss))
    mod = g.mods[0]
    for n, gref in ref.items():
        if n in (f"Omega_sqt_F_dict.{mod}", f"delta_F_dict.{mod}"):
            axis = 0 if n.startswith("Omega") else 1
            for gi in range(world // Ws):   # replicas of a gene slice hold bitwise-identical (all-reduced) gradients
                assert np.array_equal(out[gi * Ws][n], out[gi * Ws + 1][n]), n
            got = np.concatenate([out[gi * Ws][n] for gi in range(world // Ws)], axis)
        else:
            assert all(np.array_equal(out[0][n], out[r][n]) for r in range(1, world)), n
            got = out[0][n]
        assert relerr(got, gref) < 1e-4, n
