"""CPU-side checks of the drop-in boundary: public names, constructor parity with the reference's
seeded initialisation (golden state_dicts), error behaviour, and that the C-ABI library loads and
exports every symbol include/gpsa_b200.h declares.  No GPU compute here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from golden_io import ALL_CASES, Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_from_golden(g, **kw):
    import gpsa

    kern = {"rbf": gpsa.rbf_kernel, "matern12": gpsa.matern12_kernel, "matern32": gpsa.matern32_kernel}
    data_dict = {
        m: {
            "spatial_coords": torch.from_numpy(g.X[m]).float().clone(),
            "outputs": torch.from_numpy(g.Y[m]).float().clone(),
            "n_samples_list": list(g.n_samples[m]),
        }
        for m in g.mods
    }
    np.random.seed(0)
    torch.manual_seed(0)
    model = gpsa.VariationalGPSA(
        data_dict,
        n_spatial_dims=g.cfg.n_spatial_dims,
        m_X_per_view=g.cfg.m_X_per_view,
        m_G=g.cfg.m_G,
        data_init=True,
        n_latent_gps=g.n_latent,
        mean_function="identity_fixed",
        kernel_func_warp=kern[g.cfg.kernel_warp],
        kernel_func_data=kern[g.cfg.kernel_data],
        fixed_view_idx=g.fixed,
        **kw,
    )
    return model, data_dict


def test_public_names():
    import gpsa

    for name in ("GPSA", "VariationalGPSA", "rbf_kernel", "matern12_kernel", "matern32_kernel", "polar_warp",
                 "get_st_coordinates", "LossNotDecreasingChecker", "rbf_kernel_numpy", "callback_oned",
                 "callback_twod", "callback_twod_aligned_only", "callback_twod_multimodal"):
        assert hasattr(gpsa, name), name
    import gpsa.models.gpsa, gpsa.models.vgpsa, gpsa.util.util  # noqa: F401,E401
    assert gpsa.__file__.startswith(os.path.join(ROOT, "spatial-alignment_b200"))


@pytest.mark.parametrize("name", ["c1_shipped", "c2_matern", "v3_d3_free", "lmc", "multimodal", "v2_d1"])
def test_seeded_construction_matches_reference_state_dict(name):
    """Same seeds -> same parameters as the reference built (KMeans centres, RNG consumption order,
    slot order of Omega_sqt_G_list), same state_dict keys."""
    g = Golden(name)
    model, _ = _model_from_golden(g)
    sd = model.state_dict()
    ref_keys = {k for k in g.params if k not in g.fixed_params}
    assert set(sd.keys()) == ref_keys
    tweaked = {"warp_kernel_lengthscales", "data_kernel_lengthscale", "data_kernel_variance"}
    for k, v in sd.items():
        if name in ("c1_rbf_short", "v3_d2_fixedlist") and k in tweaked:
            continue
        assert v.shape == g.params[k].shape, k
        np.testing.assert_allclose(v.numpy(), g.params[k], rtol=0, atol=1e-6, err_msg=k)


def test_create_view_idx_dict_and_errors():
    import gpsa

    g = Golden("multimodal")
    model, data_dict = _model_from_golden(g)
    view_idx, Ns, Ps, n_total = model.create_view_idx_dict(data_dict)
    assert n_total == sum(sum(g.n_samples[m]) for m in g.mods)
    for m in g.mods:
        assert [len(v) for v in view_idx[m]] == g.n_samples[m]
        assert Ps[m] == g.Y[m].shape[1]
        assert view_idx[m][1][0] == g.n_samples[m][0]
    bad = {k: dict(v) for k, v in data_dict.items()}
    bad["protein"]["n_samples_list"] = [40]
    with pytest.raises(ValueError):
        gpsa.VariationalGPSA(bad, m_X_per_view=4, m_G=4, n_latent_gps={"rna": None, "protein": None})
    with pytest.raises(TypeError):  # n_latent_gps must be a dict (reference :54)
        gpsa.VariationalGPSA(data_dict, m_X_per_view=4, m_G=4)
    # any callable is accepted (unknown ones take the documented slow path, see tests/test_gpu_callable.py) ...
    m = gpsa.VariationalGPSA(data_dict, m_X_per_view=4, m_G=4, n_latent_gps={"rna": None, "protein": None},
                             kernel_func_warp=lambda *a, **k: None)
    assert m._kind_warp is None and m._kind_data == "rbf"
    with pytest.raises(TypeError):  # ... a non-callable is not
        gpsa.VariationalGPSA(data_dict, m_X_per_view=4, m_G=4, n_latent_gps={"rna": None, "protein": None},
                             kernel_func_warp="rbf")


def test_forward_refuses_cpu():
    g = Golden("v2_d1")
    model, data_dict = _model_from_golden(g)
    view_idx, Ns, _, _ = model.create_view_idx_dict(data_dict)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.forward({m: data_dict[m]["spatial_coords"] for m in g.mods}, view_idx=view_idx, Ns=Ns, S=2)


def test_fixed_hyperparameters_stay_out_of_state_dict():
    g = Golden("v2_d1")
    model, _ = _model_from_golden(g, fixed_warp_kernel_variances=[1.0, 1.0], fixed_warp_kernel_lengthscales=[2.0, 2.0],
                                  fixed_data_kernel_lengthscales=[1.5])
    sd = model.state_dict()
    assert "warp_kernel_variances" not in sd and "warp_kernel_lengthscales" not in sd
    assert "data_kernel_lengthscale" not in sd
    assert torch.allclose(model.warp_kernel_lengthscales, torch.log(torch.tensor([2.0, 2.0])))


def test_library_exports_every_declared_symbol():
    from gpsa import _lib

    header = open(os.path.join(ROOT, "include", "gpsa_b200.h")).read()
    declared = set(re.findall(r"^(?:int|long|void|size_t)\s+(gpsa_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built (run __graft_entry__.build())")
    so = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(so, sym), sym
    assert so.gpsa_version() >= 100
    so.gpsa_feat_count.restype = ctypes.c_long
    assert so.gpsa_feat_count(25) == 10 * 64 and so.gpsa_feat_count(200) == 325 * 64


def test_torch_custom_op_layer_registers_every_launcher():
    """TORCH_LIBRARY(gpsa_b200): every stream-taking entry point of include/gpsa_b200.h is a torch custom op with a
    schema (pointers -> Tensor? / Tensor(a!)?, sizes -> int), and CPU tensors are refused with a RuntimeError."""
    import re

    from gpsa import _lib

    ops = _lib.ops()
    header = open(os.path.join(ROOT, "include", "gpsa_b200.h")).read()
    launchers = re.findall(r"\bint\s+gpsa_(\w+)\s*\(([^;]*?)cudaStream_t stream\)\s*;", header, flags=re.S)
    assert len(launchers) >= 30
    for name, _ in launchers:
        assert hasattr(ops, name), f"torch.ops.gpsa_b200.{name} is not registered"
        schema = str(getattr(ops, name).default._schema)
        assert schema.startswith(f"gpsa_b200::{name}(") and schema.endswith("-> ()")
    s = str(ops.sample_fwd.default._schema)
    assert "Tensor(m4!)? a4" in s and "int a0" in s          # outputs are declared mutable
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.sample_fwd(4, 2, torch.zeros(4), torch.zeros(4, 2), torch.zeros(4, 2), torch.zeros(4, 2))
