"""Model-level parity of the tcgen05 engine at the shapes bench.py times (-m gpu).

The golden cases of tests/test_gpu_parity.py are toy-sized (R = S*N <= 1000) and therefore run the fp32 SIMT engine.
These cases have the benchmark's spot counts, inducing-point counts, sample counts and kernel template instances --
C3 (M = 200, R = 128 000), C4-shaped (D = 3, M = 256) and C5-shaped (M = 512) -- with a reduced number of genes, so
that the oracle (oracle/gpsa_oracle.py, float64 = truth, float32 = stand-in for the reference's own arithmetic, to
which it is pinned by tests/test_oracle_golden.py) can evaluate the identical model on identical noise in seconds
(on the GPU: torch float64, same code).  Gene count does not change the per-gene arithmetic: every gene is an
independent column of the same three products.

Acceptance: loss rtol 1e-4 against float64; every returned sample tensor and every parameter gradient
through golden_io.parity_ok with slack = 2 (within rtol 1e-4 of the fp32 reference arithmetic, or no further from the
float64 truth than twice what the fp32 reference arithmetic is).
"""
import os
import sys

import numpy as np
import pytest
import torch

from golden_io import parity_ok
from oracle import gpsa_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

# name: bench.py-style config (V, Nv, D, P, M, S, kernel)
CASES = {
    "c3": dict(V=4, Nv=4000, D=2, P=32, M=200, S=8, kernel="rbf", desc="C3 spots/M/S, 32 genes"),
    "c4": dict(V=8, Nv=2500, D=3, P=16, M=256, S=8, kernel="rbf", desc="C4-shaped: 3-D, M = 256"),
    "c5": dict(V=8, Nv=2000, D=2, P=16, M=512, S=8, kernel="rbf", desc="C5-shaped: M = 512"),
}


def _setup(name, engine):
    import bench
    from gpsa import _ops

    cfg = CASES[name]
    model, data_dict, X, Y, nl = bench.build_model(cfg, seed=3, kmeans=False)
    S, N, P, M, V, D = cfg["S"], cfg["V"] * cfg["Nv"], cfg["P"], cfg["M"], cfg["V"], cfg["D"]
    _ops.ENGINE["value"] = engine
    assert _ops.pick_engine(M, S * N, P) == engine
    ocfg = orc.Config(n_views=V, n_spatial_dims=D, modality_names=["expression"], n_samples_lists={"expression": nl},
                      m_X_per_view=M, m_G=M, kernel_warp=cfg["kernel"], kernel_data=cfg["kernel"], fixed_view_idx=0,
                      n_latent_gps={"expression": None})
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    eps = orc.draw_noise(ocfg, S, {"expression": P}, seed=11)
    return cfg, model, data_dict, X, Y, ocfg, params, eps


@pytest.mark.parametrize("name,engine", [("c3", 1), ("c4", 1), ("c5", 1)])
def test_benchmark_shape_matches_oracle(name, engine):
    from gpsa import _ops

    try:
        cfg, model, data_dict, X, Y, ocfg, params, eps = _setup(name, engine)
        S = cfg["S"]
        data_dev = {"expression": {"spatial_coords": data_dict["expression"]["spatial_coords"].cuda(),
                                   "outputs": data_dict["expression"]["outputs"].cuda(),
                                   "n_samples_list": data_dict["expression"]["n_samples_list"]}}
        view_idx, Ns, _, _ = model.create_view_idx_dict(data_dev)
        ret = model.forward({"expression": data_dev["expression"]["spatial_coords"]}, view_idx=view_idx, Ns=Ns, S=S,
                            _eps={"G": eps["G"], "F": eps["F"], "F_test": {}})
        loss = model.loss_fn(data_dev, ret[3])
        model.zero_grad()
        loss.backward()
        model.check_factorisations()
        torch.cuda.synchronize()
    finally:
        _ops.ENGINE["value"] = "auto"

    Xd, Yd = {"expression": X}, {"expression": Y}
    res = {}
    for dt in (torch.float64, torch.float32):
        o, c, l, g = orc.elbo_and_grads(params, ocfg, Xd, Yd, S, eps, dtype=dt, materialise=False, device="cuda")
        res[dt] = ({k: v["expression"].cpu().numpy() for k, v in o.items()}, float(l), {k: v.cpu().numpy() for k, v in g.items()})
        del o, c, g
        torch.cuda.empty_cache()
    (o64, l64, g64), (o32, l32, g32) = res[torch.float64], res[torch.float32]

    report, good = [], True

    def chk(label, new, ref32, truth):
        nonlocal good
        ok, e_ref, e_tru, r_tru = parity_ok(new.detach().cpu().numpy(), ref32, truth, rtol=1e-4, slack=2.0)
        report.append(f"{label:34s} vs_f32 {e_ref:.1e} vs_f64 {e_tru:.1e} (f32_vs_f64 {r_tru:.1e}) {'ok' if ok else 'FAIL'}")
        good &= ok

    chk("G_means", ret[0]["expression"], o32["G_means"], o64["G_means"])
    chk("G_samples", ret[1]["expression"], o32["G_samples"], o64["G_samples"])
    chk("F_latent", ret[2]["expression"], o32["F_latent"], o64["F_latent"])
    chk("F_observed", ret[3]["expression"], o32["F_observed"], o64["F_observed"])
    rel_loss = abs(float(loss) - l64) / abs(l64)
    report.append(f"{'loss':34s} ours {float(loss):.8e} f64 {l64:.8e} f32 {l32:.8e} rel {rel_loss:.1e}")
    good &= rel_loss <= 1e-4
    named = dict(model.named_parameters())
    for k in sorted(g64):
        if k not in named:
            continue
        got = named[k].grad if named[k].grad is not None else torch.zeros_like(named[k])
        chk(f"grad.{k}", got, g32[k], g64[k])
    print(f"\n[{name} engine {engine}: {cfg['desc']}]\n" + "\n".join(report))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"parity_fullsize_{name}_e{engine}.txt"), "w") as fh:
            fh.write("\n".join(report) + "\n")
    assert good, "\n" + "\n".join(r for r in report if r.endswith("FAIL") or r.startswith("loss"))
