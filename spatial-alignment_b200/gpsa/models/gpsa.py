"""Base generative model: data validation, per-view index ranges and the kernel / noise / mean
parameters.  Mirrors the public behaviour of reference gpsa/models/gpsa.py:25-183 (same
constructor signature, same parameter names, shapes, initial values and RNG consumption order,
same ValueErrors) so that state_dicts are interchangeable with the reference."""
import numpy as np
import torch
import torch.nn as nn

from ..util.util import rbf_kernel


class GPSA(nn.Module):
    """
    Args:
        data_dict (dict): {"modality": {"spatial_coords": X [N,D], "outputs": Y [N,P],
            "n_samples_list": [N_v, ...]}}
        data_init, n_spatial_dims, n_noise_variance_params, kernel_func_warp, kernel_func_data,
        mean_function, mean_penalty_param, fixed_warp_kernel_variances,
        fixed_warp_kernel_lengthscales, fixed_data_kernel_lengthscales: as in the reference
        (gpsa/models/gpsa.py:10-37).
    """

    def __init__(
        self,
        data_dict,
        data_init=True,
        n_spatial_dims=2,
        n_noise_variance_params=2,
        kernel_func_warp=rbf_kernel,
        kernel_func_data=rbf_kernel,
        mean_function="identity_fixed",
        mean_penalty_param=0.0,
        fixed_warp_kernel_variances=None,
        fixed_warp_kernel_lengthscales=None,
        fixed_data_kernel_lengthscales=None,
    ):
        super().__init__()
        self.modality_names = list(data_dict.keys())
        self.n_modalities = len(self.modality_names)
        self.mean_penalty_param = mean_penalty_param

        # every modality must describe the same views (reference :46-53)
        n_views = {len(data_dict[mod]["n_samples_list"]) for mod in self.modality_names}
        if len(n_views) != 1:
            raise ValueError("Each modality must have the same number of views.")
        self.n_views = int(n_views.pop())

        # the number of spatial dimensions is read off the data, not the argument (reference :56-68)
        dims = {int(data_dict[mod]["spatial_coords"].shape[1]) for mod in self.modality_names}
        if len(dims) != 1:
            raise ValueError("Each modality must have the same number of spatial dimensions.")
        self.n_spatial_dims = int(dims.pop())

        self.view_idx, self.Ns, self.Ps, self.n_total = self.create_view_idx_dict(data_dict)

        # 2 warp-kernel parameters per view + 2 for the data kernel (reference :77-83)
        self.n_kernel_params = 2 * self.n_views + 2
        self.n_noise_variance_params = n_noise_variance_params
        self.kernel_func_warp = kernel_func_warp
        self.kernel_func_data = kernel_func_data

        # -- parameters, created in the reference's order so seeded construction matches (:86-124)
        self.noise_variance = nn.Parameter(torch.randn([self.n_noise_variance_params]) - 1)

        n_warp = self.n_kernel_params // 2 - 1
        if fixed_warp_kernel_variances is None:
            self.warp_kernel_variances = nn.Parameter(torch.zeros(n_warp))
        else:  # a constant, kept out of the state_dict like the reference's plain tensor
            self.register_buffer(
                "warp_kernel_variances", torch.log(torch.tensor(fixed_warp_kernel_variances)).float(), persistent=False
            )
        if fixed_warp_kernel_lengthscales is None:
            self.warp_kernel_lengthscales = nn.Parameter(torch.zeros(n_warp) + np.log(10))
        else:
            self.register_buffer(
                "warp_kernel_lengthscales",
                torch.log(torch.tensor(fixed_warp_kernel_lengthscales)).float(),
                persistent=False,
            )
        if fixed_data_kernel_lengthscales is None:
            self.data_kernel_lengthscale = nn.Parameter(torch.log(torch.exp(torch.randn(1))))
        else:
            self.register_buffer(
                "data_kernel_lengthscale",
                torch.log(torch.tensor(fixed_data_kernel_lengthscales).float()).reshape(-1),
                persistent=False,
            )
        self.data_kernel_variance = nn.Parameter(torch.randn(1))

        D, V = self.n_spatial_dims, self.n_views
        if mean_function == "identity_fixed":
            self.register_buffer("mean_slopes", torch.eye(D).unsqueeze(0).repeat(V, 1, 1), persistent=False)
            self.register_buffer("mean_intercepts", torch.zeros([V, D]), persistent=False)
        elif mean_function == "identity_initialized":
            self.mean_slopes = nn.Parameter(torch.randn([V, D, D]))
            self.mean_intercepts = nn.Parameter(torch.zeros([V, D]))
        else:
            self.mean_slopes = nn.Parameter(torch.eye(D).unsqueeze(0).repeat(V, 1, 1))
            self.mean_intercepts = nn.Parameter(torch.randn([V, D]) * 0.1)

        self.diagonal_offset = 1e-5

    def create_view_idx_dict(self, data_dict):
        """Contiguous index range of every view inside each modality's row order.

        Returns (view_idx {mod: [arange ...]}, Ns {mod: numpy int}, Ps {mod: int}, n_total), exactly
        as reference gpsa/models/gpsa.py:155-183."""
        view_idx, Ns, Ps = {}, {}, {}
        n_total = 0
        for mod in self.modality_names:
            n_samples_list = data_dict[mod]["n_samples_list"]
            Ns[mod] = np.sum(n_samples_list)
            n_total += Ns[mod]
            Ps[mod] = data_dict[mod]["outputs"].shape[1]
            edges = np.insert(np.cumsum(n_samples_list), 0, 0)
            view_idx[mod] = [np.arange(edges[ii], edges[ii + 1]) for ii in range(self.n_views)]
        return view_idx, Ns, Ps, n_total

    def compute_mean_penalty(self):
        eye = torch.eye(self.n_spatial_dims, device=self.mean_slopes.device)
        return self.mean_penalty_param * torch.mean(torch.square(self.mean_slopes - eye.unsqueeze(0)))

    def forward(self, X_spatial):
        raise NotImplementedError

    def loss_fn(self, data_dict, Gs, means_G_list, covs_G_list, means_Y, covs_Y):
        raise NotImplementedError
