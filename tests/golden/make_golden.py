"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (`/root/reference`, gpsa 0.6) in the build container.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference ships no golden vectors of its own, so these
fixtures -- inputs, state_dict, the noise it drew, its four returned dicts, the
cached factors loss_fn reads, the loss and every parameter gradient -- are what
pins `oracle/gpsa_oracle.py` and, through it, the CUDA path.  The reference is a
Python package and does not travel to the GPU box; only the .npz files do.

Noise is fixed without patching the reference: `torch.manual_seed(k)` is called
immediately before `forward`, and the same draws are reproduced afterwards in
the reference's draw order and stored.
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

warnings.filterwarnings("ignore")
import gpsa  # noqa: E402  (the reference)
from gpsa import VariationalGPSA, matern12_kernel, matern32_kernel, rbf_kernel  # noqa: E402

assert gpsa.__file__.startswith(REF), gpsa.__file__
torch.autograd.set_detect_anomaly(False)

from oracle import gpsa_oracle as orc  # noqa: E402

KERN = {"rbf": rbf_kernel, "matern12": matern12_kernel, "matern32": matern32_kernel}


def load_example_h5ad():
    """examples/synthetic_data.h5ad is plain contiguous HDF5; h5py/anndata are not
    installed, so read the three datasets at their raw offsets."""
    buf = open(os.path.join(REF, "examples", "synthetic_data.h5ad"), "rb").read()
    Y = np.frombuffer(buf, "<f4", 6000, 2048).reshape(200, 30).copy()
    batch = np.frombuffer(buf, "<i8", 200, 39488).copy()
    X = np.frombuffer(buf, "<f8", 400, 47680).reshape(200, 2).copy()
    assert set(batch) == {0, 1} and (np.diff(batch) >= 0).all()
    return X.astype(np.float32), Y, batch


def synth(rng, n_list, D, P):
    xs = []
    for v, n in enumerate(n_list):
        x = rng.uniform(0, 10, size=(n, D))
        xs.append(x)
    X = np.concatenate(xs).astype(np.float32)
    W = rng.standard_normal((D, P))
    Y = np.sin(X @ W * 0.5) + 0.1 * rng.standard_normal((X.shape[0], P))
    return X, Y.astype(np.float32)


def run_case(name, data, m_X, m_G, S, kern_warp="rbf", kern_data="rbf", fixed=None, n_latent=None,
             seed=0, fwd_seed=1000, tweak=None, with_gtest=False):
    """data = {mod: (X, Y, n_samples_list)}"""
    if ONLY and name not in ONLY:
        return
    np.random.seed(seed)
    torch.manual_seed(seed)
    mods = list(data.keys())
    data_dict = {
        m: {
            "spatial_coords": torch.from_numpy(X).float().clone(),
            "outputs": torch.from_numpy(Y).float().clone(),
            "n_samples_list": list(nl),
        }
        for m, (X, Y, nl) in data.items()
    }
    n_latent = n_latent or {m: None for m in mods}
    model = VariationalGPSA(
        data_dict,
        n_spatial_dims=data[mods[0]][0].shape[1],
        m_X_per_view=m_X,
        m_G=m_G,
        data_init=True,
        n_latent_gps=n_latent,
        mean_function="identity_fixed",
        kernel_func_warp=KERN[kern_warp],
        kernel_func_data=KERN[kern_data],
        fixed_view_idx=fixed,
    )
    if tweak is not None:
        with torch.no_grad():
            tweak(model)
    view_idx, Ns, Ps, _ = model.create_view_idx_dict(data_dict)
    X_in = {m: data_dict[m]["spatial_coords"] for m in mods}

    D = model.n_spatial_dims
    G_test = None
    if with_gtest:
        g = torch.Generator().manual_seed(7)
        G_test = {m: torch.rand(1, 17, D, generator=g) * 10.0 for m in mods}

    torch.manual_seed(fwd_seed)
    ret = model.forward(X_in, view_idx=view_idx, Ns=Ns, S=S, G_test=G_test)
    G_means, G_samples, F_lat, F_obs = ret[:4]
    loss = model.loss_fn(data_dict, F_obs)
    model.zero_grad()
    loss.backward()

    cfg = orc.Config(
        n_views=int(model.n_views),
        n_spatial_dims=int(D),
        modality_names=mods,
        n_samples_lists={m: [int(x) for x in data[m][2]] for m in mods},
        m_X_per_view=m_X,
        m_G=m_G,
        kernel_warp=kern_warp,
        kernel_data=kern_data,
        fixed_view_idx=fixed,
        n_latent_gps=n_latent,
    )
    Ls = {m: int(model.n_latent_outputs[m]) for m in mods}
    eps = orc.draw_noise(cfg, S, Ls, fwd_seed, G_test=G_test)

    out = {}
    out["meta.kernel_warp"], out["meta.kernel_data"] = kern_warp, kern_data
    out["meta.S"], out["meta.m_X"], out["meta.m_G"] = S, m_X, m_G
    out["meta.fixed"] = np.array([-1] if fixed is None else np.atleast_1d(fixed))
    out["meta.fixed_is_list"] = int(isinstance(fixed, (list, tuple)))
    out["meta.mods"] = np.array(mods)
    out["meta.fwd_seed"] = fwd_seed
    out["meta.torch"] = torch.__version__
    for m in mods:
        out[f"in.X.{m}"], out[f"in.Y.{m}"] = data[m][0], data[m][1]
        out[f"in.n_samples.{m}"] = np.array(data[m][2])
        out[f"meta.n_latent.{m}"] = -1 if n_latent[m] is None else n_latent[m]
        if G_test is not None:
            out[f"in.G_test.{m}"] = G_test[m].numpy()
    sd = model.state_dict()
    named = dict(model.named_parameters())
    for k, v in sd.items():
        out[f"param.{k}"] = v.detach().numpy()
        g = named[k].grad if k in named else None
        out[f"grad.{k}"] = (g if g is not None else torch.zeros_like(v)).detach().numpy()
    # fixed_* overrides are plain tensors outside the state_dict
    for k in ("warp_kernel_variances", "warp_kernel_lengthscales", "data_kernel_lengthscale"):
        if f"param.{k}" not in out:
            out[f"param.{k}"] = getattr(model, k).detach().numpy()
            out[f"grad.{k}"] = np.zeros_like(out[f"param.{k}"])
            out[f"meta.fixed_param.{k}"] = 1
    for v, e in eps["G"].items():
        out[f"eps.G.{v}"] = e.numpy()
    for m in mods:
        out[f"eps.F.{m}"] = eps["F"][m].numpy()
        if G_test is not None:
            out[f"eps.F_test.{m}"] = eps["F_test"][m].numpy()
        out[f"out.G_means.{m}"] = G_means[m].detach().numpy()
        out[f"out.G_samples.{m}"] = G_samples[m].detach().numpy()
        out[f"out.F_latent.{m}"] = F_lat[m].detach().numpy()
        out[f"out.F_observed.{m}"] = F_obs[m].detach().numpy()
        out[f"cache.curr_Omega_tril_F.{m}"] = model.curr_Omega_tril_F[m].detach().numpy()
        if G_test is not None:
            out[f"out.F_latent_test.{m}"] = ret[4][m].detach().numpy()
            out[f"out.F_observed_test.{m}"] = ret[5][m].detach().numpy()
    out["out.loss"] = loss.detach().numpy()
    out["cache.Kuu_chol_list"] = model.Kuu_chol_list.detach().numpy()  # NaN rows for fixed views
    out["cache.curr_Omega_tril_list"] = model.curr_Omega_tril_list.detach().numpy()
    out["cache.Kuu_chol_F"] = model.Kuu_chol_F.detach().numpy()
    out["cache.noise_variance_pos"] = model.noise_variance_pos.detach().numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={float(loss):.6f}  -> {os.path.getsize(path)/1024:.0f} KiB")


ONLY = set(sys.argv[1:])  # optional: regenerate only the named cases


def main():
    X, Y, batch = load_example_h5ad()
    nl = [int((batch == 0).sum()), int((batch == 1).sum())]
    np.savez_compressed(os.path.join(HERE, "synthetic_data.npz"), X=X, Y=Y, batch=batch)

    def short_ls(model):
        model.warp_kernel_lengthscales.fill_(float(np.log(1.5)))
        model.data_kernel_lengthscale.fill_(float(np.log(1.2)))
        model.data_kernel_variance.fill_(0.3)

    # C1 as shipped: examples/grid_example.py:13-55 (P=30 from the file, M=25)
    run_case("c1_shipped", {"expression": (X, Y, nl)}, 25, 25, 5, fixed=0)
    # C1 as named in BASELINE.json (first 5 outputs, M=50) -- default RBF init, ill-conditioned
    run_case("c1_named", {"expression": (X, Y[:, :5].copy(), nl)}, 50, 50, 5, fixed=0)
    # C2: Matern-1/2 warp + data kernels on the same data
    run_case("c2_matern", {"expression": (X, Y[:, :5].copy(), nl)}, 50, 50, 5, "matern12", "matern12", fixed=0)
    # Matern-3/2 warp + data kernels (gpsa/util/util.py:50-66), same data
    run_case("c2_matern32", {"expression": (X, Y[:, :5].copy(), nl)}, 50, 50, 5, "matern32", "matern32", fixed=0)
    # well-conditioned RBF (short lengthscales) on the example data
    run_case("c1_rbf_short", {"expression": (X, Y[:, :5].copy(), nl)}, 25, 25, 5, fixed=0, tweak=short_ls)

    rng = np.random.default_rng(11)
    # V=3, D=2, fixed view given as a list: exposes the v*D+j / j*V+v index split
    Xs, Ys = synth(rng, [40, 33, 37], 2, 4)
    run_case("v3_d2_fixedlist", {"expression": (Xs, Ys, [40, 33, 37])}, 12, 14, 3, "matern12", "rbf",
             fixed=[0], tweak=short_ls)
    # V=3, D=3, no fixed view, Matern
    Xs, Ys = synth(rng, [30, 41, 35], 3, 4)
    run_case("v3_d3_free", {"expression": (Xs, Ys, [30, 41, 35])}, 10, 16, 4, "matern12", "matern12")
    # D=1
    Xs, Ys = synth(rng, [25, 28], 1, 3)
    run_case("v2_d1", {"expression": (Xs, Ys, [25, 28])}, 8, 9, 2, "matern12", "matern12", fixed=0)
    # LMC: 3 latent GPs -> 6 outputs
    Xs, Ys = synth(rng, [36, 30], 2, 6)
    run_case("lmc", {"expression": (Xs, Ys, [36, 30])}, 10, 12, 3, "matern12", "matern12", fixed=0,
             n_latent={"expression": 3})
    # two modalities sharing the warp; per-modality noise index (:534)
    Xa, Ya = synth(rng, [30, 26], 2, 4)
    Xb, Yb = synth(rng, [18, 22], 2, 3)
    run_case("multimodal", {"rna": (Xa, Ya, [30, 26]), "protein": (Xb, Yb, [18, 22])}, 9, 11, 3,
             "matern12", "matern12", fixed=0)
    # G_test prediction branch (vgpsa.py:437-477)
    Xs, Ys = synth(rng, [32, 29], 2, 3)
    run_case("gtest", {"expression": (Xs, Ys, [32, 29])}, 10, 10, 2, "matern12", "matern12", fixed=0,
             with_gtest=True)


if __name__ == "__main__":
    main()
