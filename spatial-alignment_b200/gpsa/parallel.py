"""One-process-per-GPU execution of the ELBO iteration: gene sharding, Monte-Carlo-sample sharding and the two combined
(SURVEY.md 8(e)).

Given the warped coordinates, the data GP is independent across output genes: gene p owns its own
Omega_sqt_F[p], delta_F[:, p], outputs[:, p] and noise draws.  Rank r therefore keeps a contiguous slice of
the genes -- parameters, data, Adam state and the whole tcgen05 quadratic-form work for them -- and the cheap
shared front end (warp GP of every view, K_uu / K_uf / A = K^-1 K_uf of the data GP: < 1 % of the flops) is
replicated with identical noise, so the forward needs no collective at all.  The only exchange is one
all-reduce (NCCL over NVLink/NVSwitch; gloo in the CPU tests) per iteration of the gradients of the shared
parameters, which live in ONE flat buffer that autograd accumulates into directly (the parameters' .grad are
views of it), so nothing is packed or unpacked around the collective.

    model = VariationalGPSA(local_data_dict, ...)            # outputs already sliced to this rank's genes
    sharder = GeneSharding(model, world_size, rank)           # after model.to(device)
    loss = model.loss_fn(local_data_dict, F); sharder.zero_grad(); loss.backward(); sharder.allreduce(); opt.step()

The negative ELBO decomposes as  sum_r [ -LL_r + KL_F_r + KL_G / world ]: every rank evaluates the (replicated)
warp-GP KL and weights it by 1/world, so that the SUM over ranks of the local losses -- and of the local
gradients of the shared parameters -- is exactly the unsharded loss / gradient.
"""
import torch
import torch.distributed as dist

# parameters every rank holds a full replica of (reference state_dict names)
SHARED = ("noise_variance", "warp_kernel_variances", "warp_kernel_lengthscales", "data_kernel_lengthscale",
          "data_kernel_variance", "Xtilde", "Gtilde", "Omega_sqt_G_list", "delta_G_list")


def gene_range(n_genes, world, rank):
    """Contiguous, balanced slice [lo, hi) of the genes owned by `rank` (first n_genes % world ranks get one more)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(int(n_genes), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_data_dict(data_dict, world, rank):
    """This rank's view of a reference-style data_dict: same coordinates, the gene slice of `outputs`."""
    out = {}
    for mod, d in data_dict.items():
        lo, hi = gene_range(d["outputs"].shape[1], world, rank)
        out[mod] = dict(d)
        out[mod]["outputs"] = d["outputs"][:, lo:hi].contiguous()
    return out


def shard_state_dict(state_dict, world, rank, n_latent_gps=None):
    """Slice a full (unsharded) reference/gpsa_b200 state_dict down to this rank's genes.

    Without LMC a gene owns Omega_sqt_F[p] and delta_F[:, p].  With LMC loadings (n_latent_gps[mod] set, pass the dict)
    the latent GPs are replicated and only the observed outputs are sharded: W_dict[mod] [L, P] is sliced by columns."""
    out = {}
    lmc = {m for m, k in (n_latent_gps or {}).items() if k is not None}
    for k, v in state_dict.items():
        mod = k.split(".", 1)[1] if "." in k else None
        if k.startswith("Omega_sqt_F_dict.") and mod not in lmc:
            lo, hi = gene_range(v.shape[0], world, rank)
            out[k] = v[lo:hi].clone()
        elif k.startswith("delta_F_dict.") and mod not in lmc:
            lo, hi = gene_range(v.shape[1], world, rank)
            out[k] = v[:, lo:hi].clone()
        elif k.startswith("W_dict."):
            if mod not in lmc:
                raise ValueError("state_dict holds LMC loadings: pass n_latent_gps so that they can be sharded by output column")
            lo, hi = gene_range(v.shape[1], world, rank)
            out[k] = v[:, lo:hi].clone()
        else:
            out[k] = v.clone()
    return out


class _FlatAllReduce:
    """Flat gradient buffer of the parameters every rank replicates, and the per-iteration all-reduce over it."""

    def _build_flat(self, model, names):
        named = dict(model.named_parameters())
        self.shared = [(n, named[n]) for n in names if n in named and named[n].requires_grad]
        # replicas must START identical (an initialisation with run-to-run round-off, e.g. the GPU k-means, would leave
        # the ranks a few ulps apart for good: identical gradients never pull them together): rank 0's values win
        if self.world > 1 and dist.is_available() and dist.is_initialized():
            with torch.no_grad():
                for _, p in self.shared:
                    dist.broadcast(p.data, src=0, group=self.group)
        total = sum(p.numel() for _, p in self.shared)
        ref = self.shared[0][1]
        # [shared grads ..., local loss]: the trailing slot carries the scalar loss so that logging needs no second collective
        self.flat = torch.zeros(total + 1, dtype=ref.dtype, device=ref.device)
        off = 0
        for _, p in self.shared:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()


class GeneSharding(_FlatAllReduce):
    """Gene (output) sharding: owns the flat gradient buffer of the shared parameters and the per-iteration all-reduce.

    Modalities with LMC loadings (n_latent_gps set) shard the OBSERVED outputs -- W's columns and the data's columns --
    and replicate the latent GPs: their Omega_sqt_F / delta_F join the shared parameters and their KL term carries
    1/world, like the warp layer's."""

    def __init__(self, model, world, rank, group=None):
        self.model, self.world, self.rank, self.group = model, int(world), int(rank), group
        names = list(SHARED)
        # global index of this rank's first gene, per modality: the in-kernel noise of the sampling stage is keyed by
        # the GLOBAL gene index, so the Monte-Carlo draw is the same whatever the world size
        for mod in model.modality_names:
            if model.n_latent_gps[mod] is not None:  # LMC: latent space replicated (same noise on every rank)
                names += [f"Omega_sqt_F_dict.{mod}", f"delta_F_dict.{mod}"]
                model._kl_F_scale[mod] = 1.0 / self.world
                model._gene_off[mod] = 0
                continue
            local = int(model.n_latent_outputs[mod])
            counts = [local]
            if self.world > 1 and dist.is_available() and dist.is_initialized():
                counts = [None] * self.world
                dist.all_gather_object(counts, local, group=group)
            model._gene_off[mod] = int(sum(counts[:self.rank])) if len(counts) > 1 else 0
        self._build_flat(model, names)
        model._kl_G_scale = 1.0 / self.world

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def zero_grad(self):
        """Zero in place (the .grad views must survive), drop the gene-local grads like optimizer.zero_grad()."""
        self.flat.zero_()
        shared_ids = {id(p) for _, p in self.shared}
        for p in self.model.parameters():
            if id(p) not in shared_ids:
                p.grad = None

    def allreduce(self, loss=None):
        """Sum the shared-parameter gradients (and the local losses) over ranks, in place.  Returns the global loss
        as a 0-dim tensor when `loss` is given."""
        for _, p in self.shared:
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.nbytes():
                raise RuntimeError("a shared parameter's .grad no longer aliases the flat buffer "
                                   "(use GeneSharding.zero_grad(), not optimizer.zero_grad(set_to_none=True))")
        if loss is not None:
            self.flat[-1] = loss.detach()
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat[-1].clone() if loss is not None else None


class SampleSharding(_FlatAllReduce):
    """Monte-Carlo-sample sharding, the alternative when there are fewer genes than GPUs: every rank
    holds ALL parameters and data and evaluates its contiguous share of the S samples (forward(..., S=S_total) keeps
    its own S_loc of them; the noise is keyed by the global sample index, so the draw is world-size invariant).
    loss_r = (S_loc / S) (-LL_r) + KL / world sums over ranks to the negative ELBO; ALL gradients are all-reduced
    (that includes Omega_sqt_F: 4 L M^2 bytes, the price of this partition)."""

    def __init__(self, model, world, rank, group=None):
        self.model, self.world, self.rank, self.group = model, int(world), int(rank), group
        self._build_flat(model, [n for n, _ in model.named_parameters()])
        model._sample_shard = (self.rank, self.world)
        model._kl_G_scale = 1.0 / self.world
        for mod in model.modality_names:
            model._kl_F_scale[mod] = 1.0 / self.world


for _cls in (SampleSharding,):
    _cls.nbytes = GeneSharding.nbytes
    _cls.zero_grad = GeneSharding.zero_grad
    _cls.allreduce = GeneSharding.allreduce


class HybridSharding(_FlatAllReduce):
    """Genes x Monte-Carlo samples: world = n_gene_groups * n_sample_groups ranks, rank = g * n_sample_groups + s.

    Gene sharding alone leaves the front end (warp layer, K_uf, A = K^-1 K_uf and their backward) replicated on every
    rank; sample sharding alone leaves the gene-batched Omega_F algebra replicated.  On the grid, rank (g, s) holds
    gene slice g (like GeneSharding over n_gene_groups) and evaluates sample share s of it (like SampleSharding over
    n_sample_groups): the per-sample front end is divided by n_sample_groups, the per-gene algebra by n_gene_groups,
    the products by both.  Two all-reduces per iteration: the gene-local gradients (Omega_sqt_F, delta_F of slice g)
    over the n_sample_groups ranks that share the slice -- neighbouring ranks -- and the shared gradients over all.
    loss_r = (S_loc / S) (-LL_r) + KL_F,g / n_sample_groups + KL_G / world sums over ranks to the negative ELBO."""

    def __init__(self, model, world, rank, n_sample_groups, group=None):
        self.model, self.world, self.rank, self.group = model, int(world), int(rank), group
        Ws = int(n_sample_groups)
        if Ws < 1 or self.world % Ws:
            raise ValueError(f"n_sample_groups={Ws} must divide the world size {self.world}")
        Wg = self.world // Ws
        self.g, self.s, self.Wg, self.Ws = self.rank // Ws, self.rank % Ws, Wg, Ws
        if any(model.n_latent_gps[m] is not None for m in model.modality_names):
            raise NotImplementedError("hybrid sharding with LMC loadings is not supported (use GeneSharding)")
        self.slice_group = None  # the ranks that hold my gene slice; every rank creates every group, in the same order
        if self.world > 1 and dist.is_available() and dist.is_initialized():
            for gi in range(Wg):
                grp = dist.new_group([gi * Ws + si for si in range(Ws)]) if Ws > 1 else None
                if gi == self.g:
                    self.slice_group = grp
        for mod in model.modality_names:
            local = int(model.n_latent_outputs[mod])
            counts = [local] * self.world
            if self.world > 1 and dist.is_available() and dist.is_initialized():
                counts = [None] * self.world
                dist.all_gather_object(counts, local, group=group)
            model._gene_off[mod] = int(sum(counts[gi * Ws] for gi in range(self.g)))
            model._kl_F_scale[mod] = 1.0 / Ws
        model._kl_G_scale = 1.0 / self.world
        model._sample_shard = (self.s, Ws) if Ws > 1 else None
        # shared parameters: one flat buffer over the world; gene-local parameters: a second one over the slice group
        self._build_flat(model, list(SHARED))
        named = dict(model.named_parameters())
        shared_ids = {id(p) for _, p in self.shared}
        self.local = [(n, p) for n, p in named.items() if id(p) not in shared_ids and p.requires_grad]
        self.flat_local = None
        if Ws > 1 and self.local:
            ref = self.local[0][1]
            self.flat_local = torch.zeros(sum(p.numel() for _, p in self.local), dtype=ref.dtype, device=ref.device)
            off = 0
            with torch.no_grad():
                for _, p in self.local:
                    if self.slice_group is not None:  # the replicas of a slice start identical
                        dist.broadcast(p.data, src=self.g * Ws, group=self.slice_group)
                    p.grad = self.flat_local[off:off + p.numel()].view_as(p)
                    off += p.numel()

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def zero_grad(self):
        self.flat.zero_()
        if self.flat_local is not None:
            self.flat_local.zero_()
        else:
            shared_ids = {id(p) for _, p in self.shared}
            for p in self.model.parameters():
                if id(p) not in shared_ids:
                    p.grad = None

    def allreduce(self, loss=None):
        if loss is not None:
            self.flat[-1] = loss.detach()
        if self.world > 1:
            if self.flat_local is not None:
                dist.all_reduce(self.flat_local, op=dist.ReduceOp.SUM, group=self.slice_group)
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat[-1].clone() if loss is not None else None
