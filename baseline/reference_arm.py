#!/usr/bin/env python
"""Times the UNMODIFIED reference (gpsa 0.6 installed into baseline/_ref with
`pip install --no-index --no-deps --target baseline/_ref /root/reference`) on this box's host cores:
VariationalGPSA.forward + loss_fn + backward + Adam.step through the reference's own public API, on the
synthetic data bench.py builds.  Runs in its own process (CUDA hidden, so the reference's module-level
`device` resolves to "cpu") because the reference package and ours share the name `gpsa`.

The reference materialises an [S, P, N, M] tensor (gpsa/models/vgpsa.py:193-196); at the full gene count of
C3-C5 that is 205 GB - 65 TB, so the step is timed at a few small gene counts P and bench.py extrapolates
linearly in P (stated in the JSON).  Prints one JSON object."""
import argparse
import json
import os
import sys
import time

if "--device=cuda" not in sys.argv:  # default: the CPU arm bench.py reports; CUDA hidden before torch is imported
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--genes", default="")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--anomaly", type=int, default=0)
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"],
                    help="cuda (pass as --device=cuda): the reference's own eager PyTorch path on this box's GPU "
                         "(C1/C2 only fit), the 'same box, library kernels' bar of SURVEY.md 8(d)")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "gpsa")):
        print(json.dumps({"unavailable": "baseline/_ref/gpsa missing (reference not installed)"}))
        return
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)
    import numpy as np
    import torch

    import gpsa  # the reference
    from bench import CONFIGS, make_data

    assert os.path.realpath(gpsa.__file__).startswith(os.path.realpath(REF)), gpsa.__file__
    torch.autograd.set_detect_anomaly(bool(args.anomaly))  # the reference switches it ON at import (vgpsa.py:9)
    torch.set_num_threads(os.cpu_count())
    cfg = CONFIGS[args.config]
    genes = [int(g) for g in args.genes.split(",") if g] or [cfg["P"]]
    kern = gpsa.rbf_kernel if cfg["kernel"] == "rbf" else gpsa.matern12_kernel
    out = {"cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "anomaly": bool(args.anomaly),
           "device": args.device, "runs": []}
    for P in genes:
        X, Y, nl = make_data(cfg, args.seed, genes=P)
        data_dict = {"expression": {"spatial_coords": torch.from_numpy(X), "outputs": torch.from_numpy(Y),
                                    "n_samples_list": nl}}
        np.random.seed(args.seed)
        torch.manual_seed(args.seed)
        model = gpsa.VariationalGPSA(data_dict, n_spatial_dims=cfg["D"], m_X_per_view=cfg["M"], m_G=cfg["M"],
                                     data_init=True, minmax_init=False, grid_init=False,
                                     n_latent_gps={"expression": None}, mean_function="identity_fixed",
                                     kernel_func_warp=kern, kernel_func_data=kern, fixed_view_idx=0)
        if args.device == "cuda":
            model = model.to("cuda")
            for k in ("spatial_coords", "outputs"):
                data_dict["expression"][k] = data_dict["expression"][k].cuda()
        view_idx, Ns, _, _ = model.create_view_idx_dict(data_dict)
        opt = torch.optim.Adam(model.parameters(), lr=1e-2)
        x = data_dict["expression"]["spatial_coords"]

        def step():
            model.train()
            G_means, G_samples, F_latent, F_samples = model.forward({"expression": x}, view_idx=view_idx, Ns=Ns,
                                                                    S=cfg["S"])
            loss = model.loss_fn(data_dict, F_samples)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.item())

        for _ in range(args.warmup):
            step()
        ts = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            loss = step()  # ends in loss.item(): the device is synchronised
            ts.append(time.perf_counter() - t0)
        out["runs"].append({"genes": P, "s_per_step": float(np.median(ts)), "s_min": float(np.min(ts)), "steps": args.steps,
                            "loss": loss})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
