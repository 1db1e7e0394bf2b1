"""Load the fixtures written by tests/golden/make_golden.py."""
import glob
import os

import numpy as np

from oracle import gpsa_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ALL_CASES = sorted(
    os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if "synthetic_data" not in p
)
# default RBF init (lengthscale 10 on a [0,10] domain): cond(K_uu) ~ 1e6..1e7, the
# reference's own fp32 results are 1e-3 away from fp64 there
ILL_CONDITIONED = {"c1_shipped", "c1_named"}


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.z = z
        self.mods = [str(m) for m in z["meta.mods"]]
        fixed = z["meta.fixed"]
        if fixed[0] < 0:
            self.fixed = None
        elif int(z["meta.fixed_is_list"]):
            self.fixed = [int(f) for f in fixed]
        else:
            self.fixed = int(fixed[0])
        self.S = int(z["meta.S"])
        self.fwd_seed = int(z["meta.fwd_seed"])
        self.X = {m: z[f"in.X.{m}"] for m in self.mods}
        self.Y = {m: z[f"in.Y.{m}"] for m in self.mods}
        self.n_samples = {m: [int(x) for x in z[f"in.n_samples.{m}"]] for m in self.mods}
        self.n_latent = {m: (None if int(z[f"meta.n_latent.{m}"]) < 0 else int(z[f"meta.n_latent.{m}"])) for m in self.mods}
        self.params = {k[len("param."):]: z[k] for k in z.files if k.startswith("param.")}
        self.grads = {k[len("grad."):]: z[k] for k in z.files if k.startswith("grad.")}
        self.fixed_params = [k[len("meta.fixed_param."):] for k in z.files if k.startswith("meta.fixed_param.")]
        self.eps = {
            "G": {int(k.split(".")[-1]): z[k] for k in z.files if k.startswith("eps.G.")},
            "F": {m: z[f"eps.F.{m}"] for m in self.mods},
            "F_test": {m: z[f"eps.F_test.{m}"] for m in self.mods if f"eps.F_test.{m}" in z.files},
        }
        self.G_test = {m: z[f"in.G_test.{m}"] for m in self.mods if f"in.G_test.{m}" in z.files} or None
        self.cfg = orc.Config(
            n_views=len(self.n_samples[self.mods[0]]),
            n_spatial_dims=int(self.X[self.mods[0]].shape[1]),
            modality_names=self.mods,
            n_samples_lists=self.n_samples,
            m_X_per_view=int(z["meta.m_X"]),
            m_G=int(z["meta.m_G"]),
            kernel_warp=str(z["meta.kernel_warp"]),
            kernel_data=str(z["meta.kernel_data"]),
            fixed_view_idx=self.fixed,
            n_latent_gps=self.n_latent,
        )
        self.loss = float(z["out.loss"])

    def out(self, key, mod=None):
        return self.z[f"out.{key}" + (f".{mod}" if mod else "")]

    def cache(self, key, mod=None):
        return self.z[f"cache.{key}" + (f".{mod}" if mod else "")]


def relerr(a, b):
    """max |a-b| / max|b| (scale-relative, robust to entries near zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = max(np.max(np.abs(b)), 1e-30)
    return float(np.max(np.abs(a - b)) / scale)


def parity_ok(new, ref32, truth64, rtol=1e-4, slack=3.0):
    """The acceptance rule (SURVEY.md 7.5 / BASELINE north_star): `new` matches the
    reference's fp32 result within rtol, or -- where the reference's own fp32
    arithmetic is further than that from the float64 truth (ill-conditioned K_uu,
    cancellation in small gradients) -- `new` is no further from the truth than
    `slack` times the reference is.  Returns (ok, err_vs_ref, err_vs_truth, ref_err)."""
    e_ref = relerr(new, ref32)
    e_tru = relerr(new, truth64)
    r_tru = relerr(ref32, truth64)
    ok = e_ref <= rtol or e_tru <= max(rtol, slack * r_tru)
    return ok, e_ref, e_tru, r_tru
