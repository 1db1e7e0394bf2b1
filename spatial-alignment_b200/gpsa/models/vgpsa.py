"""Variational GPSA: the deep-GP ELBO hot path on B200.

Public surface of reference gpsa/models/vgpsa.py (constructor :15-34, forward :212, loss_fn :491)
with the arithmetic moved into the CUDA library behind include/gpsa_b200.h.  The reference's
documented quirks are reproduced on purpose; each is marked below.
"""
import os
from collections.abc import Iterable

import numpy as np
import torch
import torch.nn as nn
from sklearn.cluster import KMeans

from .gpsa import GPSA
from .. import _lib, _ops
from ..lazy import LazySamples
from ..util.util import kmeans_gpu, matern12_kernel, matern32_kernel, rbf_kernel

# data_init=True: inputs with more spots than this are clustered on the GPU (gpsa.util.kmeans_gpu) instead of with the
# host KMeans the reference calls; below it the reference's seeded behaviour is reproduced exactly
KMEANS_HOST_MAX = 20000

_DEBUG_CHECKS = os.environ.get("GPSA_B200_CHECK", "0") == "1"


def _kernel_kind(fn):
    """The model receives kernel CALLABLES (reference :25-26).  The three the reference exports are recognised by
    identity and run fused (distances in registers, analytic gradients).  Any other callable with the reference's
    signature k(x1, x2, lengthscale_unconstrained=..., output_variance_unconstrained=...) takes the documented SLOW
    PATH (returns None here): K_uu and K_uf are evaluated by the callable itself with torch, differentiated by
    torch autograd, and enter the layers as matrices (kernel kind 3 of include/gpsa_b200.h); everything downstream of
    the two matrices -- factorisations, solves, the quadratic forms, sampling, likelihood, KL -- stays on the fused
    path."""
    if fn is rbf_kernel:
        return "rbf"
    if fn is matern12_kernel:
        return "matern12"
    if fn is matern32_kernel:
        return "matern32"
    if not callable(fn):
        raise TypeError(f"kernel function must be callable, got {type(fn).__name__}")
    return None


class VariationalGPSA(GPSA):
    def __init__(
        self,
        data_dict,
        m_X_per_view,
        m_G,
        data_init=True,
        minmax_init=False,
        grid_init=False,
        n_spatial_dims=2,
        n_noise_variance_params=2,
        kernel_func_warp=rbf_kernel,
        kernel_func_data=rbf_kernel,
        n_latent_gps=None,
        mean_function="identity_fixed",
        mean_penalty_param=0.0,
        fixed_warp_kernel_variances=None,
        fixed_warp_kernel_lengthscales=None,
        fixed_data_kernel_lengthscales=None,
        fixed_view_idx=None,
    ):
        # quirk 8: n_spatial_dims / n_noise_variance_params / mean_function / minmax_init are ignored
        # (reference :35-46 hard-codes 2 / 2 / the default identity mean)
        super().__init__(
            data_dict,
            data_init=True,
            n_spatial_dims=2,
            n_noise_variance_params=2,
            kernel_func_warp=kernel_func_warp,
            kernel_func_data=kernel_func_data,
            mean_penalty_param=mean_penalty_param,
            fixed_warp_kernel_variances=fixed_warp_kernel_variances,
            fixed_warp_kernel_lengthscales=fixed_warp_kernel_lengthscales,
            fixed_data_kernel_lengthscales=fixed_data_kernel_lengthscales,
        )
        self._kind_warp = _kernel_kind(kernel_func_warp)
        self._kind_data = _kernel_kind(kernel_func_data)
        self.m_X_per_view = m_X_per_view
        self.m_G = m_G
        self.n_latent_gps = n_latent_gps
        self.n_latent_outputs = {}
        for mod in self.modality_names:
            # quirk 8: n_latent_gps must be a dict; the default None raises TypeError here like :54
            k = self.n_latent_gps[mod]
            self.n_latent_outputs[mod] = k if k is not None else self.Ps[mod]
        self.fixed_view_idx = fixed_view_idx
        V, D = self.n_views, self.n_spatial_dims

        n_points = sum(int(data_dict[mod]["spatial_coords"].shape[0]) for mod in self.modality_names)
        if data_init and n_points > KMEANS_HOST_MAX and torch.cuda.is_available():
            # reference :61-92 with sklearn's host KMeans replaced by Lloyd's iterations on the GPU (gpsa.util.kmeans_gpu):
            # at C4 / C5 sizes the host fit is seconds to minutes and dominates the time to the first iteration.
            # Seeded by numpy's global generator, like the reference's KMeans (random_state=None).
            Xtilde = torch.zeros([V, self.m_X_per_view, D])
            for vv in range(V):
                xs = [data_dict[mod]["spatial_coords"][self.view_idx[mod][vv], :] for mod in self.modality_names]
                curr_X = torch.cat(xs, dim=0)
                Xtilde[vv] = kmeans_gpu(curr_X, self.m_X_per_view, seed=int(np.random.randint(1 << 31)))[0].cpu()
            self.Xtilde = nn.Parameter(Xtilde.clone())
            all_X = torch.cat([data_dict[mod]["spatial_coords"] for mod in self.modality_names])
            self.Gtilde = nn.Parameter(kmeans_gpu(all_X, self.m_G, seed=int(np.random.randint(1 << 31)))[0].cpu())
        elif data_init:  # reference :61-92
            Xtilde = torch.zeros([V, self.m_X_per_view, D])
            for vv in range(V):
                xs = [data_dict[mod]["spatial_coords"][self.view_idx[mod][vv], :] for mod in self.modality_names]
                curr_X = torch.cat(xs, dim=0)
                km = KMeans(n_clusters=self.m_X_per_view)
                km.fit(curr_X.detach().cpu().numpy())
                Xtilde[vv] = torch.tensor(km.cluster_centers_)
            self.Xtilde = nn.Parameter(Xtilde.clone())
            # the reference draws (and discards) a random subset here (:81-85); the draw is kept so that
            # seeded construction consumes the numpy RNG identically -- including its ValueError when
            # m_G exceeds the size of the last view
            np.random.choice(np.arange(curr_X.shape[0]), size=self.m_G, replace=False)
            all_X = torch.cat([data_dict[mod]["spatial_coords"] for mod in self.modality_names])
            km = KMeans(n_clusters=self.m_G)
            km.fit(all_X.detach().cpu().numpy())
            # float32 even when the coordinates came in as float64 (the library is fp32 at this boundary; the
            # reference's Gtilde would follow the KMeans dtype and then fail in its own float32 matmuls)
            self.Gtilde = nn.Parameter(torch.tensor(km.cluster_centers_).float())
        elif grid_init:  # reference :94-121 (2-D only)
            if D == 2:
                coords = data_dict[self.modality_names[0]]["spatial_coords"].cpu().numpy()
                (xlow, ylow), (xhigh, yhigh) = coords.min(0), coords.max(0)
                numticks = np.ceil(np.sqrt(self.m_G)).astype(int)
                self.m_G = numticks**2
                self.m_X_per_view = numticks**2
                X1, X2 = np.meshgrid(np.linspace(xlow, xhigh, num=numticks), np.linspace(ylow, yhigh, num=numticks))
                grid = np.vstack([X1.ravel(), X2.ravel()]).T
                Xt = torch.zeros([V, grid.shape[0], D])
                for vv in range(V):
                    Xt[vv] = torch.tensor(grid)
                self.Xtilde = nn.Parameter(Xt.clone())
                self.Gtilde = nn.Parameter(torch.tensor(grid).float())
        else:  # reference :123-128
            self.Xtilde = nn.Parameter(torch.randn([V, self.m_X_per_view, D]))
            self.Gtilde = nn.Parameter(torch.randn([self.m_G, D]))

        M_X, M_G = int(self.m_X_per_view), int(self.m_G)
        # variational covariance square roots; slot j*V + v (reference :131-143)
        Osq_G = torch.zeros([V * D, M_X, M_X])
        for vv in range(V):
            for jj in range(D):
                Osq_G[jj * V + vv] = 0.1 * torch.randn(size=[M_X, M_X])
        self.Omega_sqt_G_list = nn.Parameter(Osq_G)

        self.Omega_sqt_F_dict = nn.ParameterDict()
        for mod in self.modality_names:  # reference :145-153
            L = self.n_latent_outputs[mod]
            curr = torch.zeros([L, M_G, M_G])
            for jj in range(L):
                curr[jj] = 0.1 * torch.randn(size=[M_G, M_G])
            self.Omega_sqt_F_dict[mod] = nn.Parameter(curr)

        # variational means (reference :156-164)
        self.delta_G_list = nn.Parameter(self.Xtilde.detach().clone())
        self.delta_F_dict = nn.ParameterDict()
        for mod in self.modality_names:
            self.delta_F_dict[mod] = nn.Parameter(torch.randn(size=[M_G, self.n_latent_outputs[mod]]))

        # LMC loadings (reference :167-172)
        self.W_dict = nn.ParameterDict()
        for mod in self.modality_names:
            if self.n_latent_gps[mod] is not None:
                self.W_dict[mod] = nn.Parameter(torch.randn([self.n_latent_gps[mod], self.Ps[mod]]))

        scale = torch.ones(V, 1, 1)
        for vv in range(V):
            if self._is_fixed(vv):
                scale[vv] = 100.0  # reference :235
        self.register_buffer("_mu_z_scale", scale, persistent=False)
        kl_mask = torch.zeros(V * D)
        self.register_buffer("_kl_mask", kl_mask, persistent=False)
        self._idx_cache = {}
        self._kl = None
        # weights of the loss terms under gpsa.parallel (all 1 on a single GPU), chosen so that the local losses of the
        # ranks SUM to the negative ELBO: replicated terms carry 1/world, a rank's share of the samples carries S_loc/S
        self._kl_G_scale = 1.0
        self._kl_F_scale = {mod: 1.0 for mod in self.modality_names}
        self._nll_scale = 1.0
        self._sample_shard = None  # (rank, world): this rank evaluates its contiguous share of the S Monte-Carlo samples
        # --- execution knobs of the B200 path (not part of the reference's constructor) ---
        # rng_mode: where the data layer's noise eps_F [S,N,L] comes from when forward() is not handed explicit noise.
        #   "philox" (default): drawn INSIDE the fused kernels from a counter-based generator keyed by
        #       (seed, sample, spot, global gene); the 64-bit seed is itself drawn from torch's CUDA generator, so
        #       torch.manual_seed governs it.  Nothing [S,N,L]-sized is materialised, and the draw does not depend on how
        #       genes are sharded over ranks.
        #   "torch": torch.randn(S, N, L) in the reference's draw order.
        self.rng_mode = "philox"
        # fused_ll: forward() returns lazy handles for F_samples, and loss_fn() on such a handle runs the fused
        # sampling + likelihood kernel (gpsa/lazy.py).  False: forward() returns plain tensors like the reference.
        self.fused_ll = True
        self._gene_off = {mod: 0 for mod in self.modality_names}  # global index of this rank's first gene (gpsa.parallel)
        self._gen = 0

    # ----------------------------------------------------------------------------------------------
    def _is_fixed(self, vv):
        f = self.fixed_view_idx  # reference :230-234
        if f is None:
            return False
        return (vv in f) if isinstance(f, Iterable) else (f == vv)

    def _device_index(self, idx, device):
        """view_idx entries are numpy index arrays; the reference re-uploads them on every indexing
        op (:268-271,:291,:344,:351).  Here each is uploaded once and cached; contiguous ranges
        (what create_view_idx_dict produces) become slices."""
        key = (id(idx), str(device))
        hit = self._idx_cache.get(key)
        if hit is not None and hit[0] is idx:
            return hit[1]
        arr = np.asarray(idx)
        if arr.size > 0 and arr.ndim == 1 and np.array_equal(arr, np.arange(arr[0], arr[0] + arr.size)):
            val = slice(int(arr[0]), int(arr[0]) + int(arr.size))
        elif arr.size == 0:
            val = slice(0, 0)
        else:
            val = torch.as_tensor(arr, dtype=torch.long, device=device)
        self._idx_cache[key] = (idx, val)
        return val

    def get_Omega_from_Omega_sqt(self, Omega_sqt):
        """Omega_sqt Omega_sqt^T + 1e-5 I (reference :206-210), differentiable in Omega_sqt like the reference's."""
        return _ops.OmegaFromSqt.apply(Omega_sqt)

    def compute_mean_and_var(self, *args, **kwargs):
        """Reference :174-204 is an internal helper of forward(); here its arithmetic lives inside the fused warp /
        data layer kernels (gpsa_warp_view_fwd, gpsa_data_layer_fwd) and is not exposed as a separate torch op."""
        raise NotImplementedError(
            "compute_mean_and_var is fused into gpsa_b200's warp / data layer kernels; call forward() "
            "(see INTEGRATION.md, 'API differences')")

    # ----------------------------------------------------------------------------------------------
    def forward(self, X_spatial, view_idx, Ns, S=1, prediction_mode=False, G_test=None, _eps=None):
        """Same contract as reference gpsa/models/vgpsa.py:212-489.

        Returns (G_means, G_samples, F_latent_samples, F_observed_samples) -- plus the two *_test
        dicts when G_test is given -- each a {modality: Tensor} dict.  `_eps` (tests only) injects
        the noise: {"G": {v: [S,n_v,D]}, "F": {mod: [S,N,L]}, "F_test": {mod: [...]}}; without it the
        noise is drawn with torch in exactly the reference's call order.
        """
        if prediction_mode:
            self.eval()
        dev = self.Xtilde.device
        if dev.type != "cuda":
            raise RuntimeError("VariationalGPSA.forward needs the model on a CUDA device: gpsa_b200 has no CPU path")
        V, D, mods = self.n_views, self.n_spatial_dims, self.modality_names
        S = int(S)
        S_total, s0 = S, 0
        if self._sample_shard is not None:  # gpsa.parallel.SampleSharding: this rank's samples are [s0, s0 + S) of S_total
            from ..parallel import gene_range as _split

            s0, s1 = _split(S_total, self._sample_shard[1], self._sample_shard[0])
            S = s1 - s0
            if S == 0:
                raise ValueError(f"sample sharding needs S >= world size (S={S_total}, world={self._sample_shard[1]})")
            self._nll_scale = S / S_total

        self.noise_variance_pos = torch.exp(self.noise_variance) + self.diagonal_offset  # reference :217
        self.mu_z_G = self.Xtilde * self._mu_z_scale  # identity mean; x100 on fixed views (:219-235)

        # ---- gather each view's coordinates over the modalities (reference :284-294)
        free, X_views, sizes = [], {}, {}
        for vv in range(V):
            if self._is_fixed(vv):
                continue
            xs = [X_spatial[mod][self._device_index(view_idx[mod][vv], dev)] for mod in mods]
            sizes[vv] = [int(x.shape[0]) for x in xs]
            Xv = xs[0] if len(xs) == 1 else torch.cat(xs, dim=0)
            if Xv.shape[0] == 0:  # reference :296-297
                continue
            free.append(vv)
            X_views[vv] = Xv

        # ---- noise, in the reference's draw order (per free view S draws, reference :346-348)
        eps_G = {}
        for vv in free:
            if _eps is not None:
                eps_G[vv] = _eps["G"][vv].to(dev, torch.float32)[s0:s0 + S]
            else:
                # all S_total draws are made (reference order, :346-348) and this rank keeps its samples, so the
                # Monte-Carlo draw does not depend on the world size
                e = torch.empty(S_total, X_views[vv].shape[0], D, device=dev)
                for ss in range(S_total):
                    e[ss].normal_()
                eps_G[vv] = e[s0:s0 + S] if S != S_total else e

        if not self._kl_mask_ready(free):
            self._set_kl_mask(free)
        # Omega_F = Omega_sqt Omega_sqt^T + eps I and its factor do not depend on the warp layer: an autograd node of its
        # own (_ops.OmegaChain) on a side stream, so the gene-batched factorisation runs under the warp layer's
        # latency-bound per-view chains -- and its backward (the inverse and three batched GEMMs) under their backward
        chain_F = {}
        cur = torch.cuda.current_stream(dev)
        om_stream = _ops.omega_stream(dev)
        om_stream.wait_stream(cur)
        with torch.cuda.stream(om_stream):
            for mod in mods:
                cm = {}
                chain_F[mod] = _ops.OmegaChain.apply(cm, self.Omega_sqt_F_dict[mod]) + (cm,)

        # ... and so is the data layer's K_uu = k(Gtilde, Gtilde): a single-matrix fp64 factorisation + inverse, pure
        # latency, that would otherwise sit on the critical path between the two layers
        prior_F, pr_stream = None, None
        if self._kind_data is not None:
            pr_stream = _ops.omega_stream(dev, "prior")
            pr_stream.wait_stream(cur)
            with torch.cuda.stream(pr_stream):
                prior_F = _ops.prior_prepare(_ops.KINDS[self._kind_data], self.Gtilde, self.data_kernel_lengthscale,
                                             self.data_kernel_variance)

        ext_w = self._kind_warp is None
        meta = {"kind": _lib.KIND_EXTERNAL if ext_w else _ops.KINDS[self._kind_warp], "V": V, "S": S, "free": free,
                "with_kl": True, "kl_mask": self._kl_mask}
        flat = []
        for vv in free:
            flat += [X_views[vv], eps_G[vv]]
            if ext_w:  # slow path: the user's covariance callable, evaluated and differentiated by torch (:314-318)
                ls_v, var_v = self.warp_kernel_lengthscales[vv], self.warp_kernel_variances[vv]
                flat += [self.kernel_func_warp(self.Xtilde[vv], self.Xtilde[vv], lengthscale_unconstrained=ls_v,
                                               output_variance_unconstrained=var_v),
                         self.kernel_func_warp(self.Xtilde[vv], X_views[vv], lengthscale_unconstrained=ls_v,
                                               output_variance_unconstrained=var_v)]
        outs = _ops.WarpLayer.apply(
            meta, self.Xtilde, self.delta_G_list, self.Omega_sqt_G_list, self.warp_kernel_lengthscales,
            self.warp_kernel_variances, *flat,
        )
        kl_G, self.Kuu_chol_list, self.curr_Omega_tril_list, info_G = outs[:4]
        per_view = {vv: (outs[4 + 2 * k], outs[5 + 2 * k]) for k, vv in enumerate(free)}

        # ---- assemble G_means / G_samples per modality (reference :262-273, :342-351)
        G_means, G_samples = {}, {}
        for mm, mod in enumerate(mods):
            N = int(Ns[mod])
            idxs = [self._device_index(view_idx[mod][vv], dev) for vv in range(V)]
            parts_m, parts_s = [], []
            for vv in range(V):
                if vv in per_view:
                    o = sum(sizes[vv][:mm])
                    n = sizes[vv][mm]
                    parts_m.append(per_view[vv][0][o:o + n])
                    parts_s.append(per_view[vv][1][:, o:o + n])
                else:  # fixed (or empty) view: observed coordinates pass through for every sample
                    xv = X_spatial[mod][idxs[vv]]
                    parts_m.append(xv)
                    parts_s.append(xv.unsqueeze(0).expand(S, -1, -1))
            ordered = all(isinstance(i, slice) for i in idxs) and all(
                idxs[k].stop == idxs[k + 1].start for k in range(V - 1)) and idxs[0].start == 0
            if ordered and idxs[-1].stop == N:
                G_means[mod] = torch.cat(parts_m, dim=0)
                G_samples[mod] = torch.cat(parts_s, dim=1)
            else:
                gm = torch.full((N, D), float("nan"), device=dev)
                gs = torch.full((S, N, D), float("nan"), device=dev)
                for vv in range(V):
                    gm = gm.index_put((self._as_index(idxs[vv], dev),), parts_m[vv])
                    gs = gs.index_put((slice(None), self._as_index(idxs[vv], dev)), parts_s[vv])
                G_means[mod], G_samples[mod] = gm, gs

        # ---- data layer per modality (reference :390-435)
        self._gen += 1
        self.curr_Omega_tril_F = {}
        self.F_latent_samples, self.F_observed_samples = {}, {}
        if G_test is not None:
            self.F_latent_samples_test, self.F_observed_samples_test = {}, {}
        cur.wait_stream(om_stream)  # join the Omega_F chain; its tensors live in the side stream's pool
        for mod in mods:
            for t in chain_F[mod][:5]:
                if t is not None:
                    t.record_stream(cur)
        if prior_F is not None:
            cur.wait_stream(pr_stream)
            for t in prior_F:
                t.record_stream(cur)
        kl = kl_G if self._kl_G_scale == 1.0 else kl_G * self._kl_G_scale
        infos = [info_G]
        ext_d = self._kind_data is None
        kind_d = _lib.KIND_EXTERNAL if ext_d else _ops.KINDS[self._kind_data]

        def _k_ext(Gs):
            """Slow path of the data layer (:390, :409): (k(Gt,Gt) [M,M], k(Gt,G) as [M, S*N]) from the user's callable."""
            if not ext_d:
                return ()
            ls, var = self.data_kernel_lengthscale, self.data_kernel_variance
            Kuu = self.kernel_func_data(self.Gtilde, self.Gtilde, lengthscale_unconstrained=ls,
                                        output_variance_unconstrained=var)
            Kuf = self.kernel_func_data(self.Gtilde, Gs, lengthscale_unconstrained=ls, output_variance_unconstrained=var)
            return Kuu, Kuf.permute(1, 0, 2).reshape(Kuf.shape[1], -1)   # [S,M,N] -> [M, S*N], r = s*N + n

        for mod in mods:
            L = self.n_latent_outputs[mod]
            N = int(Ns[mod])
            Osq = self.Omega_sqt_F_dict[mod]
            Omega_F, hld_F, Ltril_O, L64_O, info_O, chain_meta = chain_F[mod]
            # noise of the sampling stage (reference :423): explicit, torch's stream, or a key for the in-kernel generator
            eps_F, key = None, None
            if _eps is not None:
                eps_F = _eps["F"][mod].to(dev, torch.float32)[s0:s0 + S]
            elif self.rng_mode == "torch":
                eps_F = torch.randn(S_total, N, L, device=dev)[s0:s0 + S]
            elif self.rng_mode == "philox":
                key = torch.randint(-(1 << 62), 1 << 62, (1,), dtype=torch.int64, device=dev)
            else:
                raise ValueError(f"rng_mode must be 'philox' or 'torch', got {self.rng_mode!r}")
            mean, q2, kq, kl_F, self.Kuu_chol_F, Ltril_F, info_F = _ops.DataLayerPre.apply(
                {"kind": kind_d, "with_kl": True, "omega_chain": (Ltril_O, info_O, chain_meta), "prior": prior_F},
                self.Gtilde, self.data_kernel_lengthscale, self.data_kernel_variance, self.delta_F_dict[mod], Osq,
                G_samples[mod], *(_k_ext(G_samples[mod]) or (None, None)), Omega_F, hld_F,
            )
            kl = kl + (kl_F if self._kl_F_scale[mod] == 1.0 else kl_F * self._kl_F_scale[mod])
            infos.append(info_F)
            self.curr_Omega_tril_F[mod] = Ltril_F
            W = self.W_dict[mod] if self.n_latent_gps[mod] is not None else None
            goff = int(self._gene_off.get(mod, 0))

            def _sample(mean=mean, q2=q2, kq=kq, eps_F=eps_F, key=key, S=S, N=N, L=L, goff=goff, s0=s0):
                e = eps_F if eps_F is not None else _ops.philox_normal(key, S, N, L, gene_off=goff, samp_off=s0)
                return _ops.SampleF.apply(mean, q2, kq, e)

            F_lat = LazySamples((S, N, L), torch.float32, dev, _sample, f"F_latent_samples[{mod!r}]")
            if W is None:
                F_obs = F_lat
                if self.fused_ll:
                    F_lat._fused = {"owner": self, "gen": self._gen, "mean": mean, "q2": q2, "kq": kq, "eps": eps_F,
                                    "key": key, "gene_off": goff, "samp_off": s0, "consumed": False}
            else:  # LMC (:428-432)
                F_obs = LazySamples((S, N, self.Ps[mod]), torch.float32, dev,
                                    lambda F_lat=F_lat, W=W: _ops.LMCObserve.apply(F_lat.materialise(), W),
                                    f"F_observed_samples[{mod!r}]")
                if self.fused_ll and _ops.lmc_fused_supported(L):
                    F_obs._fused = {"owner": self, "gen": self._gen, "lmc": (F_lat, W), "consumed": False}
            if not self.fused_ll:
                F_lat = F_lat.materialise()
                F_obs = F_lat if W is None else F_obs.materialise()
            self.F_latent_samples[mod] = F_lat
            self.F_observed_samples[mod] = F_obs
            if G_test is not None:  # prediction at given aligned coordinates (reference :437-477)
                Gt = G_test[mod].to(dev, torch.float32)
                if _eps is not None:
                    eps_t = _eps["F_test"][mod].to(dev, torch.float32)
                else:
                    eps_t = torch.randn(Gt.shape[0], Gt.shape[1], L, device=dev)
                pre = (Omega_F.detach(), Ltril_O, L64_O if L64_O is not None else Ltril_O, hld_F.detach(), info_O)
                F_t = _ops.DataLayer.apply(
                    {"kind": kind_d, "with_kl": False, "omega": pre},
                    self.Gtilde, self.data_kernel_lengthscale, self.data_kernel_variance, self.delta_F_dict[mod],
                    Osq, Gt, eps_t, *_k_ext(Gt),
                )[0]
                self.F_latent_samples_test[mod] = F_t
                self.F_observed_samples_test[mod] = _ops.LMCObserve.apply(F_t, W) if W is not None else F_t
        self._kl = kl
        self._info = infos
        if _DEBUG_CHECKS:
            self.check_factorisations()
        else:
            self._deferred_factorisation_check(infos)

        if G_test is not None:
            return (G_means, G_samples, self.F_latent_samples, self.F_observed_samples,
                    self.F_latent_samples_test, self.F_observed_samples_test)
        return G_means, G_samples, self.F_latent_samples, self.F_observed_samples

    def _deferred_factorisation_check(self, infos):
        """Sync-free stand-in for the exception torch.cholesky raises in the reference: the `info` flags of this
        forward are copied to pinned host memory asynchronously; a later forward (normally the next one) finds them
        arrived and raises.  Independently of this, a failed factorisation makes the loss NaN on the device
        (csrc/chol.cu), so it can never pass silently."""
        if torch.cuda.is_current_stream_capturing():
            return  # no event queries / host copies while a CUDA graph is being captured
        pend = getattr(self, "_pending_info", None)
        if pend is not None and pend[1].query():
            self._pending_info = None
            if bool(pend[0].any()):
                bad = int(torch.nonzero(pend[0])[0, 0])
                raise RuntimeError(f"VariationalGPSA.forward: a Cholesky factorisation of an earlier forward met a "
                                   f"non-positive pivot (flag {bad}): a K_uu or Omega matrix is not positive-definite")
            pend = None
        if pend is not None:
            return
        n = sum(int(i.numel()) for i in infos)
        host = getattr(self, "_info_host", None)
        if host is None or host.numel() < n:
            host = self._info_host = torch.zeros(max(n, 64), dtype=torch.int32).pin_memory()
        o = 0
        for i in infos:
            host[o:o + i.numel()].copy_(i, non_blocking=True)
            o += int(i.numel())
        ev = torch.cuda.Event()
        ev.record()
        self._pending_info = (host[:n], ev)

    def check_factorisations(self):
        """Raise if any Cholesky of the last forward met a non-positive pivot (torch.cholesky raises
        eagerly in the reference; here the flags stay on the device until asked, so the hot path has
        no host synchronisation).  GPSA_B200_CHECK=1 calls this after every forward."""
        for info in self._info:
            _ops.check_info(info, "VariationalGPSA.forward")

    @staticmethod
    def _as_index(i, dev):
        if isinstance(i, slice):
            return torch.arange(i.start, i.stop, device=dev)
        return i

    def _kl_mask_ready(self, free):
        return getattr(self, "_kl_mask_free", None) == tuple(free)

    def _set_kl_mask(self, free):
        V, D = self.n_views, self.n_spatial_dims
        m = torch.zeros(V * D)
        for vv in free:
            for jj in range(D):
                m[jj * V + vv] = -0.5  # quirk 2: the KL uses slice j*V+v (reference :508)
        self._kl_mask.copy_(m)
        self._kl_mask_free = tuple(free)

    # ----------------------------------------------------------------------------------------------
    def loss_fn(self, data_dict, F_samples):
        """Negative ELBO, reference gpsa/models/vgpsa.py:491-540: -LL/S-averaged + KL_G + KL_F, where the
        KL terms are those of the most recent forward (they depend on the parameters only).

        F_samples[mod] may be the lazy handle forward() returned (fused sampling + likelihood kernel: the samples are
        never written to memory) or any [S,N,P] tensor (plain Gaussian log-likelihood kernel), e.g. a handle that
        was materialised because something else read it."""
        if self._kl is None:
            raise RuntimeError("loss_fn needs a preceding forward (it reads the factors cached there)")
        nll = 0
        for mm, mod in enumerate(self.modality_names):
            Y = data_dict[mod]["outputs"]
            F = F_samples[mod]
            idx = self.noise_variance.shape[0] - self.n_modalities + mm  # quirk 6: index -n_modalities+mm (:534)
            log_noise = self.noise_variance[idx:idx + 1]
            fz = getattr(F, "_fused", None) if isinstance(F, LazySamples) else None
            live = (fz is not None and fz["owner"] is self and fz["gen"] == self._gen and not fz["consumed"]
                    and not F.is_materialised)
            if live and "lmc" in fz:
                # LMC: the (small) latent samples are materialised, the [S,N,P] observed samples are not
                fz["consumed"] = True
                F_lat, W = fz["lmc"]
                nll = nll + _ops.LMCNLL.apply(F_lat.materialise(), W, Y.to(F.device, torch.float32), log_noise)
            elif live:
                fz["consumed"] = True
                Yd = Y.to(F.device, torch.float32)
                S = F.shape[0]
                nll = nll + _ops.SampleNLL.apply({"gene_off": fz["gene_off"], "samp_off": fz["samp_off"]}, fz["mean"],
                                                 fz["q2"], fz["kq"], Yd, log_noise, fz["eps"], fz["key"])
                # the buffer of `mean` now holds U = -(Y - F)/(sigma^2 S): a later read of the handle (plotting, tests)
                # recovers the samples from it, detached: F = Y + sigma^2 S U
                sigma = self.noise_variance_pos[idx].detach()

                def _recover(U=fz["mean"], Yd=Yd, sigma=sigma, S=S):
                    with torch.no_grad():
                        return Yd.unsqueeze(0) + (sigma * sigma * S) * U.detach()

                F._produce = _recover
                fz["mean"] = fz["q2"] = fz["kq"] = fz["eps"] = None
            else:
                Ft = F.materialise() if isinstance(F, LazySamples) else F
                nll = nll - _ops.GaussianLL.apply(Ft, Y.to(Ft.device, torch.float32), log_noise)
        return (nll if self._nll_scale == 1.0 else nll * self._nll_scale) + self._kl
