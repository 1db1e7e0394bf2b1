"""Kernel functions of the reference's gpsa/util/util.py behind the fused CUDA evaluator, plus the
small host-side helpers that `gpsa/__init__.py` re-exports."""
import numpy as np
import torch

from gpsa import _ops


def _as_param(t, like):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t, dtype=torch.float32)
    return t.to(device=like.device, dtype=torch.float32)


def _diag_kernel(fn_name, x1, x2, lengthscale_unconstrained, output_variance_unconstrained):
    # diag=True is never used by the model (reference gpsa/models/vgpsa.py:306-312 are commented out);
    # it is elementwise over paired points, so plain tensor ops express it.
    ls = torch.exp(lengthscale_unconstrained)
    var = torch.exp(output_variance_unconstrained)
    d = x1 - x2
    if fn_name == "rbf":
        return var * torch.exp(-0.5 * torch.sum(torch.square(d / ls), dim=-1))
    r = torch.sqrt(torch.sum(torch.square(d), dim=-1) + 1e-10)
    if fn_name == "matern12":
        return var * torch.exp(-0.5 * r / ls)
    t = np.sqrt(3.0) * r / ls
    return var * (1 + t) * torch.exp(-t)


def rbf_kernel(x1, x2, lengthscale_unconstrained, output_variance_unconstrained, diag=False):
    """var * exp(-0.5 * sum(((x1 - x2) / ls)^2)) with log-scale parameters
    (reference gpsa/util/util.py:8-23).  x1 [n1,D], x2 [..., n2, D] -> [..., n1, n2]."""
    if diag:
        return _diag_kernel("rbf", x1, x2, lengthscale_unconstrained, output_variance_unconstrained)
    return _ops.kernel_matrix(
        "rbf", x1, x2, _as_param(lengthscale_unconstrained, x1), _as_param(output_variance_unconstrained, x1)
    )


def matern12_kernel(x1, x2, lengthscale_unconstrained, output_variance_unconstrained, diag=False):
    """var * exp(-0.5 * sqrt(|x1 - x2|^2 + 1e-10) / ls)  (reference gpsa/util/util.py:33-47)."""
    if diag:
        return _diag_kernel("matern12", x1, x2, lengthscale_unconstrained, output_variance_unconstrained)
    return _ops.kernel_matrix(
        "matern12", x1, x2, _as_param(lengthscale_unconstrained, x1), _as_param(output_variance_unconstrained, x1)
    )


def matern32_kernel(x1, x2, lengthscale_unconstrained, output_variance_unconstrained, diag=False):
    """var * (1 + t) * exp(-t), t = sqrt(3) * sqrt(|x1 - x2|^2 + 1e-10) / ls  (reference gpsa/util/util.py:50-66)."""
    if diag:
        return _diag_kernel("matern32", x1, x2, lengthscale_unconstrained, output_variance_unconstrained)
    return _ops.kernel_matrix(
        "matern32", x1, x2, _as_param(lengthscale_unconstrained, x1), _as_param(output_variance_unconstrained, x1)
    )


def _kmeanspp_seeds(Xd, K, rng):
    """k-means++ seeding (Arthur & Vassilvitskii 2007) on the device: every next centre is drawn with probability
    proportional to the squared distance to the nearest centre chosen so far.  The uniforms come from `rng` (numpy) and
    are uploaded once, so there is no host synchronisation inside the loop."""
    N = Xd.shape[0]
    u = torch.as_tensor(rng.random(K), dtype=torch.float64, device=Xd.device)
    idx = torch.empty(K, dtype=torch.long, device=Xd.device)
    idx[0] = int(rng.integers(N))
    d2 = torch.sum((Xd - Xd[idx[0]]) ** 2, dim=1).double()
    for k in range(1, K):
        cs = torch.cumsum(d2, 0)
        j = torch.searchsorted(cs, u[k] * cs[-1]).clamp_(max=N - 1)
        idx[k] = j
        d2 = torch.minimum(d2, torch.sum((Xd - Xd[j]) ** 2, dim=1).double())
    return Xd[idx].clone()


def kmeans_gpu(X, n_clusters, iters=30, seed=0, n_init=3):
    """Lloyd's k-means on the GPU (csrc/aux.cu: gpsa_kmeans_lloyd): X [N,D] (D <= 3) -> (centres [K,D] float32 CUDA
    tensor, inertia float).  k-means++ seeding, `n_init` restarts (the lowest inertia wins), numpy Generator(seed).
    Used by VariationalGPSA(data_init=True) for inputs too large for the host KMeans of the reference
    (gpsa/models/vgpsa.py:61-92), where the inducing-point initialisation dominates the time to the first iteration."""
    from gpsa import _lib

    if not torch.cuda.is_available():
        raise _lib.GPSALibraryError("kmeans_gpu needs a CUDA device (no CPU fallback exists)")
    Xd = torch.as_tensor(X, dtype=torch.float32)
    Xd = (Xd if Xd.is_cuda else Xd.cuda()).contiguous()
    N, D = Xd.shape
    K = int(n_clusters)
    if K > N:
        raise ValueError(f"n_clusters={K} exceeds the number of points {N}")
    rng = np.random.default_rng(seed)
    assign = torch.empty(N, dtype=torch.int32, device=Xd.device)
    sums = torch.empty(K * (D + 1), dtype=torch.float64, device=Xd.device)
    best = None
    with torch.cuda.device(Xd.device):
        for _ in range(max(1, int(n_init))):
            centres = _kmeanspp_seeds(Xd, K, rng)
            inertia = torch.zeros(1, dtype=torch.float64, device=Xd.device)
            _lib.ops().kmeans_lloyd(N, D, K, Xd, centres, int(iters), assign, sums, inertia)
            val = float(inertia)
            if best is None or val < best[1]:
                best = (centres, val)
    return best


def rbf_kernel_numpy(x, xp, kernel_params):
    """Host-side helper of the data simulators (reference gpsa/util/util.py:26-30)."""
    output_scale = np.exp(kernel_params[0])
    lengthscales = np.exp(kernel_params[1:])
    diffs = np.expand_dims(x / lengthscales, 1) - np.expand_dims(xp / lengthscales, 0)
    return output_scale * np.exp(-0.5 * np.sum(diffs**2, axis=2))


def polar_warp(X, r, theta):
    """reference gpsa/util/util.py:69-70."""
    return np.array([X[:, 0] + r * np.cos(theta), X[:, 1] + r * np.sin(theta)]).T


def get_st_coordinates(df):
    """Coordinates from an 'AxB' spot index (reference gpsa/util/util.py:73-84)."""
    return np.array([[float(t) for t in spot.split("x")] for spot in df.index])


def compute_distance(X1, X2):
    """reference gpsa/util/util.py:87-88."""
    return np.mean(np.sqrt(np.sum((X1 - X2) ** 2, axis=1)))


class LossNotDecreasingChecker:
    """Windowed convergence test on the loss trace (reference gpsa/util/util.py:257-278)."""

    def __init__(self, max_epochs, atol=1e-2, window_size=10):
        self.max_epochs = max_epochs
        self.atol = atol
        self.window_size = window_size
        self.decrease_in_loss = np.zeros(max_epochs)
        self.average_decrease_in_loss = np.zeros(max_epochs)

    def check_loss(self, iternum, loss_trace):
        if iternum >= 1:
            self.decrease_in_loss[iternum] = loss_trace[iternum - 1] - loss_trace[iternum]
            if iternum >= self.window_size:
                window = self.decrease_in_loss[iternum - self.window_size + 1: iternum]
                self.average_decrease_in_loss[iternum] = np.mean(window)
                return self.average_decrease_in_loss[iternum] < self.atol
        return False
