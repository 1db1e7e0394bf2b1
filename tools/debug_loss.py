#!/usr/bin/env python
"""Prints the loss of the first few training steps of a bench configuration (optionally with a reduced gene count
and a chosen quadratic-form engine) -- a divergence / NaN probe."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spatial-alignment_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from gpsa import _ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--genes", type=int, default=None)
ap.add_argument("--lo", type=int, default=0)
ap.add_argument("--engine", default="auto")
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
cfg = dict(bench.CONFIGS[a.config])
_ops.ENGINE["value"] = a.engine if a.engine == "auto" else int(a.engine)
sl = slice(a.lo, a.lo + a.genes) if a.genes else None
model, data_dict, X, Y, nl = bench.build_model(cfg, 3, sl)
dd = {"expression": {"spatial_coords": data_dict["expression"]["spatial_coords"].cuda(),
                     "outputs": data_dict["expression"]["outputs"].cuda(), "n_samples_list": nl}}
view_idx, Ns, _, _ = model.create_view_idx_dict(dd)
opt = torch.optim.Adam(model.parameters(), lr=1e-2)
for it in range(a.steps):
    torch.manual_seed(1000 + it)
    out = model.forward({"expression": dd["expression"]["spatial_coords"]}, view_idx=view_idx, Ns=Ns, S=cfg["S"])
    loss = model.loss_fn(dd, out[3])
    opt.zero_grad(set_to_none=True)
    loss.backward()
    bad = [n for n, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    fin = {k: bool(torch.isfinite(v).all()) for k, v in (("G", out[1]["expression"]), ("F", out[3]["expression"]))}
    print(f"engine={a.engine} genes={a.genes} lo={a.lo} it={it} loss={float(loss):.6e} finite={fin} bad_grads={bad}", flush=True)
    opt.step()
