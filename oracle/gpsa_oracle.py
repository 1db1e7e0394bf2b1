"""CPU oracle for the GPSA variational ELBO hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and there only as the checker or as the
reported CPU baseline -- never as the thing shipped.

It restates, in plain torch-on-CPU with an explicit dtype (float32 to mirror
the reference, float64 as ground truth) and EXPLICIT noise draws, what
`/root/reference/gpsa/models/vgpsa.py:212-540` computes, including the
reference's quirks.  Gradients come from torch
autograd over this restatement.

Parity pin: the reference has no golden vectors or known-answer tests
(`/root/reference/tests/test_import.py:1-2`, `test_test.py:1-2`), so the oracle
is pinned against outputs of the reference itself run in the build container:
`tests/golden/make_golden.py` imports the unmodified reference, records its
inputs / state_dict / noise / outputs / loss / gradients into
`tests/golden/*.npz`, and `tests/test_oracle_golden.py` checks this file against
every one of them.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

OFF = 1e-5  # gpsa/models/gpsa.py:153  (diagonal_offset)

RBF = "rbf"
MATERN12 = "matern12"
MATERN32 = "matern32"


# ---------------------------------------------------------------------------
# covariance functions
# ---------------------------------------------------------------------------
def kernel_matrix(kind: str, x1, x2, log_ls, log_var):
    """k(x1, x2) broadcast over leading batch dims of x2.

    rbf      gpsa/util/util.py:8-23   var * exp(-0.5 * sum(((x-z)/ls)^2))
    matern12 gpsa/util/util.py:33-47  var * exp(-0.5 * sqrt(|x-z|^2 + 1e-10) / ls)
    matern32 gpsa/util/util.py:50-66  var * (1+t) * exp(-t), t = sqrt(3)*d/ls
    Parameters are on the log scale.
    """
    ls = torch.exp(log_ls)
    var = torch.exp(log_var)
    d = x1.unsqueeze(-2) - x2.unsqueeze(-3)
    if kind == RBF:
        return var * torch.exp(-0.5 * torch.sum(torch.square(d / ls), dim=-1))
    r = torch.sqrt(torch.sum(torch.square(d), dim=-1) + 1e-10)
    if kind == MATERN12:
        return var * torch.exp(-0.5 * r / ls)
    if kind == MATERN32:
        t = math.sqrt(3.0) * r / ls
        return var * (1 + t) * torch.exp(-t)
    raise ValueError(kind)


def omega_from_sqt(omega_sqt):
    """gpsa/models/vgpsa.py:206-210."""
    m = omega_sqt.shape[-1]
    eye = torch.eye(m, dtype=omega_sqt.dtype)
    return omega_sqt @ omega_sqt.transpose(-1, -2) + OFF * eye


def kl_mvn_tril(m_q, L_q, m_p, L_p):
    """KL(N(m_q, L_q L_q^T) || N(m_p, L_p L_p^T)), batched over leading dims.

    Restates torch/distributions/kl.py::_kl_multivariatenormal_multivariatenormal
    as called from gpsa/models/vgpsa.py:506-516 and :520-530.
    """
    M = m_q.shape[-1]
    half_logdet = torch.log(torch.diagonal(L_p, dim1=-2, dim2=-1)).sum(-1) - torch.log(
        torch.diagonal(L_q, dim1=-2, dim2=-1)
    ).sum(-1)
    A = torch.linalg.solve_triangular(L_p, L_q, upper=False)  # broadcasts over the batch of q
    tr = torch.sum(torch.square(A), dim=(-2, -1))
    e = (m_p - m_q).unsqueeze(-1)
    b = torch.linalg.solve_triangular(L_p, e, upper=False)
    maha = torch.sum(torch.square(b), dim=(-2, -1))
    return half_logdet + 0.5 * (tr + maha - M)


# ---------------------------------------------------------------------------
# sparse-GP predictive moments
# ---------------------------------------------------------------------------
def mean_and_var(kff_diag, Kuf, Kuu_chol, mu_x, mu_z, delta, Omega_tril, materialise=True):
    """gpsa/models/vgpsa.py:174-204.

    2-D Kuf ([M,N], warp layer): Omega_tril is [B,M,M], result var is [B,N].
    3-D Kuf ([S,M,N], data layer): Omega_tril is [L,M,M], result var is [S,L,N],
    computed -- as the reference does -- through an [S,L,N,M] intermediate when
    `materialise` is True (the CPU-baseline cost model), or by an equivalent
    einsum when False (same value, lets the oracle reach larger shapes).
    The jitter is added twice (:191/:201 and :204).
    """
    alpha = torch.cholesky_solve(Kuf, Kuu_chol)  # :177
    aKa = torch.sum(torch.square(alpha.transpose(-1, -2) @ Kuu_chol), dim=-1)  # :179-180
    mu = mu_x.unsqueeze(0) + alpha.transpose(-1, -2) @ (delta - mu_z)  # :182-184
    if alpha.dim() == 2:
        t = alpha.transpose(-1, -2).unsqueeze(0) @ Omega_tril  # :187-189
        aOa = torch.sum(torch.square(t), dim=-1)
        var = kff_diag - aKa + aOa + OFF  # :191
    else:
        if materialise:
            t = alpha.transpose(-1, -2).unsqueeze(1) @ Omega_tril.unsqueeze(0)  # :193-195
            aOa = torch.sum(torch.square(t), dim=-1)  # :196
        else:
            Om = Omega_tril @ Omega_tril.transpose(-1, -2)
            aOa = torch.einsum("smn,lmk,skn->sln", alpha, Om, alpha)
        var = kff_diag.unsqueeze(1) - aKa.unsqueeze(1) + aOa + OFF  # :197-202
    return mu, var + OFF  # :204


# ---------------------------------------------------------------------------
# model description
# ---------------------------------------------------------------------------
@dataclass
class Config:
    """Static description of one model (everything that is not a tensor)."""

    n_views: int
    n_spatial_dims: int
    modality_names: List[str]
    n_samples_lists: Dict[str, List[int]]
    m_X_per_view: int
    m_G: int
    kernel_warp: str = RBF
    kernel_data: str = RBF
    fixed_view_idx: Optional[object] = None  # int, iterable of int, or None
    n_latent_gps: Dict[str, Optional[int]] = field(default_factory=dict)

    def is_fixed(self, v: int) -> bool:
        f = self.fixed_view_idx  # gpsa/models/vgpsa.py:230-234
        if f is None:
            return False
        if isinstance(f, (list, tuple, set, np.ndarray)):
            return v in f
        return int(f) == v

    def view_idx(self) -> Dict[str, List[np.ndarray]]:
        """gpsa/models/gpsa.py:155-183 (contiguous per-view ranges)."""
        out = {}
        for mod in self.modality_names:
            c = np.concatenate([[0], np.cumsum(self.n_samples_lists[mod])])
            out[mod] = [np.arange(c[i], c[i + 1]) for i in range(self.n_views)]
        return out


def noise_shapes(cfg: Config, S: int, Ls: Dict[str, int]):
    """Shapes of the noise the reference draws, in its draw order:
    per free view S draws of [N_v(all modalities), D] (vgpsa.py:346-348), then per
    modality one [S, N, L] (vgpsa.py:423)."""
    shapes = {"G": {}, "F": {}}
    for v in range(cfg.n_views):
        if cfg.is_fixed(v):
            continue
        n_v = sum(cfg.n_samples_lists[m][v] for m in cfg.modality_names)
        if n_v == 0:
            continue
        shapes["G"][v] = (S, n_v, cfg.n_spatial_dims)
    for mod in cfg.modality_names:
        shapes["F"][mod] = (S, int(np.sum(cfg.n_samples_lists[mod])), Ls[mod])
    return shapes


def draw_noise(cfg: Config, S: int, Ls: Dict[str, int], seed: int, dtype=torch.float32, G_test=None):
    """Draw noise exactly as the reference would after `torch.manual_seed(seed)`
    on CPU: Normal.rsample() == torch.randn(shape) there (verified), draw order as
    in `noise_shapes`; the optional G_test draw follows each modality's F draw
    (vgpsa.py:465)."""
    torch.manual_seed(seed)
    sh = noise_shapes(cfg, S, Ls)
    eps = {"G": {}, "F": {}, "F_test": {}}
    for v, (s, n, d) in sh["G"].items():
        eps["G"][v] = torch.stack([torch.randn(n, d) for _ in range(s)]).to(dtype)
    for mod, shape in sh["F"].items():
        eps["F"][mod] = torch.randn(shape).to(dtype)
        if G_test is not None:
            eps["F_test"][mod] = torch.randn(G_test[mod].shape[0], G_test[mod].shape[1], Ls[mod]).to(dtype)
    return eps


# ---------------------------------------------------------------------------
# forward + loss
# ---------------------------------------------------------------------------
def forward(params, cfg: Config, X_spatial, view_idx, S, eps, G_test=None, materialise=True):
    """gpsa/models/vgpsa.py:212-489 with the noise passed in.

    `params` uses the reference's state_dict keys:
    noise_variance, warp_kernel_variances, warp_kernel_lengthscales,
    data_kernel_lengthscale, data_kernel_variance, Xtilde, Gtilde,
    Omega_sqt_G_list, delta_G_list, Omega_sqt_F_dict.<mod>, delta_F_dict.<mod>
    [, W_dict.<mod>].
    Returns (out, cache): `out` holds the four (six) returned dicts, `cache` the
    attributes loss_fn reads.
    """
    V, D, M = cfg.n_views, cfg.n_spatial_dims, cfg.m_X_per_view
    dt = params["Xtilde"].dtype
    cache = {}
    cache["noise_variance_pos"] = torch.exp(params["noise_variance"]) + OFF  # :217

    # :219-235  identity mean function; x100 for fixed views (never read again)
    mu_z = []
    for v in range(V):
        m = params["Xtilde"][v] @ torch.eye(D, dtype=dt)
        mu_z.append(m * 100.0 if cfg.is_fixed(v) else m)
    mu_z = torch.stack(mu_z)
    cache["mu_z_G"] = mu_z

    L_omega_G = torch.linalg.cholesky(omega_from_sqt(params["Omega_sqt_G_list"]))  # :255-257
    cache["curr_Omega_tril_list"] = L_omega_G
    Kuu_chol_list = [None] * V

    mods = cfg.modality_names
    G_means = {m: [None] * V for m in mods}
    G_samples = {m: [None] * V for m in mods}

    for v in range(V):  # :259
        if cfg.is_fixed(v):  # :262-273
            for m in mods:
                xv = X_spatial[m][view_idx[m][v]]
                G_means[m][v] = xv
                G_samples[m][v] = xv.unsqueeze(0).expand(S, -1, -1)
            continue
        xs = [X_spatial[m][view_idx[m][v]] for m in mods]  # :284-294
        sizes = [x.shape[0] for x in xs]
        Xv = torch.cat(xs, dim=0)
        if Xv.shape[0] == 0:  # :296-297
            for m in mods:
                G_means[m][v] = Xv
                G_samples[m][v] = Xv.unsqueeze(0).expand(S, -1, -1)
            continue
        Z = params["Xtilde"][v]
        ls, var = params["warp_kernel_lengthscales"][v], params["warp_kernel_variances"][v]
        kff = torch.ones(Xv.shape[0], dtype=dt) * torch.exp(var)  # :310-312
        Kuu = kernel_matrix(cfg.kernel_warp, Z, Z, ls, var) + OFF * torch.eye(M, dtype=dt)  # :314-316
        Kuf = kernel_matrix(cfg.kernel_warp, Z, Xv, ls, var)  # :318
        Lk = torch.linalg.cholesky(Kuu)  # :320
        Kuu_chol_list[v] = Lk
        mu, Sig = mean_and_var(kff, Kuf, Lk, Xv, mu_z, params["delta_G_list"], L_omega_G)  # :323-331
        mu_v = mu[v]  # :335
        scale = Sig[v * D : v * D + D].t()  # :336-339   variance used as the scale; index v*D+j
        samp = mu_v.unsqueeze(0) + scale.unsqueeze(0) * eps["G"][v]  # :346-348
        o = 0
        for m, n in zip(mods, sizes):  # :342-351
            G_means[m][v] = mu_v[o : o + n]
            G_samples[m][v] = samp[:, o : o + n]
            o += n
    cache["Kuu_chol_list"] = Kuu_chol_list

    def scatter(parts, idx_list, lead):
        n = sum(len(i) for i in idx_list)
        order = torch.as_tensor(np.concatenate(idx_list), dtype=torch.long)
        cat = torch.cat(parts, dim=lead)
        inv = torch.empty(n, dtype=torch.long)
        inv[order] = torch.arange(n)
        return cat.index_select(lead, inv)

    Gm = {m: scatter(G_means[m], view_idx[m], 0) for m in mods}
    Gs = {m: scatter(G_samples[m], view_idx[m], 1) for m in mods}

    # data layer  :382-435
    Gt = params["Gtilde"]
    lsF, varF = params["data_kernel_lengthscale"], params["data_kernel_variance"]
    KuuF = kernel_matrix(cfg.kernel_data, Gt, Gt, lsF, varF) + OFF * torch.eye(cfg.m_G, dtype=dt)
    LkF = torch.linalg.cholesky(KuuF)  # :394
    cache["Kuu_chol_F"] = LkF
    cache["curr_Omega_tril_F"] = {}
    F_lat, F_obs, F_lat_t, F_obs_t = {}, {}, {}, {}

    def data_predict(G, Lom, delta, e):
        kff = torch.ones(G.shape[:2], dtype=dt) * torch.exp(varF)  # :405-407
        Kuf = kernel_matrix(cfg.kernel_data, Gt, G, lsF, varF)  # :409
        zN = torch.zeros(G.shape[1], delta.shape[1], dtype=dt)
        zM = torch.zeros(cfg.m_G, delta.shape[1], dtype=dt)
        mu, Sig = mean_and_var(kff, Kuf, LkF, zN, zM, delta, Lom, materialise)  # :413-421
        return mu + torch.sqrt(Sig.transpose(1, 2)) * e  # :423-426

    for m in mods:
        Lom = torch.linalg.cholesky(omega_from_sqt(params[f"Omega_sqt_F_dict.{m}"]))  # :410-412
        cache["curr_Omega_tril_F"][m] = Lom
        delta = params[f"delta_F_dict.{m}"]
        W = params.get(f"W_dict.{m}")
        F_lat[m] = data_predict(Gs[m], Lom, delta, eps["F"][m])
        F_obs[m] = F_lat[m] @ W if W is not None else F_lat[m]  # :428-432
        if G_test is not None:  # :437-477
            F_lat_t[m] = data_predict(G_test[m], Lom, delta, eps["F_test"][m])
            F_obs_t[m] = F_lat_t[m] @ W if W is not None else F_lat_t[m]

    out = {"G_means": Gm, "G_samples": Gs, "F_latent": F_lat, "F_observed": F_obs}
    if G_test is not None:
        out["F_latent_test"] = F_lat_t
        out["F_observed_test"] = F_obs_t
    return out, cache


def loss_fn(params, cfg: Config, cache, F_samples, outputs):
    """Negative ELBO, gpsa/models/vgpsa.py:491-540.  `outputs` = {mod: Y [N,P]}."""
    V, D = cfg.n_views, cfg.n_spatial_dims
    kl = 0.0
    for v in range(V):  # :498-516
        if cfg.is_fixed(v) or cache["Kuu_chol_list"][v] is None:
            continue
        for j in range(D):
            kl = kl + kl_mvn_tril(
                params["delta_G_list"][v, :, j],
                cache["curr_Omega_tril_list"][j * V + v],  # index j*V+v  (:508)
                cache["mu_z_G"][v, :, j],
                cache["Kuu_chol_list"][v],
            )
    ll = 0.0
    n_mod = len(cfg.modality_names)
    for mm, m in enumerate(cfg.modality_names):  # :523-538
        delta = params[f"delta_F_dict.{m}"]
        kl = kl + kl_mvn_tril(
            delta.t(), cache["curr_Omega_tril_F"][m], torch.zeros(cfg.m_G, dtype=delta.dtype), cache["Kuu_chol_F"]
        ).sum()
        sigma = cache["noise_variance_pos"][-n_mod + mm]  # :534  used as the Normal *scale*
        F = F_samples[m]
        S = F.shape[0]
        lp = -0.5 * torch.square((outputs[m] - F) / sigma) - torch.log(sigma) - 0.5 * math.log(2 * math.pi)
        ll = ll + lp.sum() / S  # :538
    return -ll + kl  # :540


def elbo_and_grads(params_np, cfg: Config, X_np, Y_np, S, eps, dtype=torch.float64, materialise=True, G_test=None,
                   device="cpu"):
    """One full hot-path iteration (forward + loss + backward) at `dtype`.

    Returns (out, cache, loss, grads) with everything detached.  This is the
    function the parity tests call; at float64 it is the ground truth the
    tolerance rule of SURVEY 7.5 is anchored on.

    `device`: where torch evaluates the restatement.  "cpu" everywhere except the
    benchmark-sized parity tests (tests/test_gpu_fullsize.py), which pass "cuda" so
    that the float64 truth of a 128 000-row case takes seconds instead of minutes;
    it is the same code either way (torch factory calls follow the device context).
    """
    with torch.device(device):
        return _elbo_and_grads(params_np, cfg, X_np, Y_np, S, eps, dtype, materialise, G_test)


def _elbo_and_grads(params_np, cfg: Config, X_np, Y_np, S, eps, dtype, materialise, G_test):
    params = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in params_np.items()}
    X = {m: torch.tensor(np.asarray(v), dtype=dtype) for m, v in X_np.items()}
    Y = {m: torch.tensor(np.asarray(v), dtype=dtype) for m, v in Y_np.items()}
    e = {
        "G": {v: torch.as_tensor(np.asarray(t)).to(dtype) for v, t in eps["G"].items()},
        "F": {m: torch.as_tensor(np.asarray(t)).to(dtype) for m, t in eps["F"].items()},
        "F_test": {m: torch.as_tensor(np.asarray(t)).to(dtype) for m, t in eps.get("F_test", {}).items()},
    }
    gt = None if G_test is None else {m: torch.tensor(np.asarray(v), dtype=dtype) for m, v in G_test.items()}
    out, cache = forward(params, cfg, X, cfg.view_idx(), S, e, G_test=gt, materialise=materialise)
    loss = loss_fn(params, cfg, cache, out["F_observed"], Y)
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}

    def det(x):
        if isinstance(x, dict):
            return {k: det(v) for k, v in x.items()}
        if isinstance(x, list):
            return [det(v) for v in x]
        return x.detach() if torch.is_tensor(x) else x

    return det(out), det(cache), loss.detach(), grads


# ---------------------------------------------------------------------------
# parameter initialisation that mirrors the reference's shapes (not its RNG)
# ---------------------------------------------------------------------------
def init_params(cfg: Config, X_np, Ps: Dict[str, int], seed=0, kmeans=True):
    """Create a parameter dict with the reference's names/shapes/initial scales
    (gpsa/models/gpsa.py:86-124, gpsa/models/vgpsa.py:61-172).  Used by tests and
    the bench to build synthetic models without the reference present; the values
    are NOT meant to reproduce the reference's RNG stream."""
    rng = np.random.default_rng(seed)
    V, D, M, MG = cfg.n_views, cfg.n_spatial_dims, cfg.m_X_per_view, cfg.m_G
    vi = cfg.view_idx()
    p = {}
    p["noise_variance"] = rng.standard_normal(2).astype(np.float32) - 1.0
    p["warp_kernel_variances"] = np.zeros(V, np.float32)
    p["warp_kernel_lengthscales"] = np.full(V, math.log(10.0), np.float32)
    p["data_kernel_lengthscale"] = rng.standard_normal(1).astype(np.float32)
    p["data_kernel_variance"] = rng.standard_normal(1).astype(np.float32)

    def centres(x, k):
        if kmeans:
            from sklearn.cluster import KMeans

            return KMeans(n_clusters=k, n_init=1, random_state=int(rng.integers(1 << 30))).fit(x).cluster_centers_
        return x[rng.choice(x.shape[0], k, replace=x.shape[0] < k)]

    Xt = np.zeros((V, M, D), np.float32)
    for v in range(V):
        xv = np.concatenate([np.asarray(X_np[m])[vi[m][v]] for m in cfg.modality_names])
        Xt[v] = centres(xv, M)
    p["Xtilde"] = Xt
    allx = np.concatenate([np.asarray(X_np[m]) for m in cfg.modality_names])
    p["Gtilde"] = centres(allx, MG).astype(np.float32)
    p["Omega_sqt_G_list"] = (0.1 * rng.standard_normal((V * D, M, M))).astype(np.float32)
    p["delta_G_list"] = Xt.copy()
    for m in cfg.modality_names:
        L = cfg.n_latent_gps.get(m) or Ps[m]
        p[f"Omega_sqt_F_dict.{m}"] = (0.1 * rng.standard_normal((L, MG, MG))).astype(np.float32)
        p[f"delta_F_dict.{m}"] = rng.standard_normal((MG, L)).astype(np.float32)
        if cfg.n_latent_gps.get(m) is not None:
            p[f"W_dict.{m}"] = rng.standard_normal((L, Ps[m])).astype(np.float32)
    return p
