#!/usr/bin/env python
"""Headline benchmark: VariationalGPSA fwd + loss + bwd + Adam step (one ELBO iteration) on synthetic data
of the shapes BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

Prints ONE JSON line (rank 0).  metric = spot-samples/s = iters/s x S x N_spots (BASELINE.json);
`value` is device-timed with inputs resident in HBM, `e2e` is the same iteration driven from pinned
HOST buffers (H2D copy of coordinates + outputs and a D2H read of the loss every step).  `roofline`
covers the dominant kernels (the implicit-feature quadratic-form GEMMs) timed live with CUDA events
inside the library; `cpu_baseline` is the UNMODIFIED reference (baseline/_ref) timed on this box's host
cores on a bounded sample.  `--impl reference` prints that CPU arm as its own line.

Without --config the headline is C3 (the configuration the metric is quoted on) and, at N = 1, the line also
carries `other_configs`: C1 and C2 (whole iteration replayed from a CUDA graph, next to the reference's own eager
PyTorch path on the same GPU) and C4 (the largest single-GPU shape), each with ms_per_step, roofline and clocks.
C5 is an 8-GPU job: `bench.py --gpus 8 --config c5` under torchrun.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "spatial-alignment_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

# id: (V, N_v, D, P, M, S, kernel)   SURVEY.md 8 / BASELINE.md 4
CONFIGS = {
    "c1": dict(V=2, Nv=100, D=2, P=30, M=25, S=5, kernel="rbf", desc="grid_example as shipped (synthetic stand-in, P=30, M=25)"),
    "c2": dict(V=2, Nv=100, D=2, P=5, M=50, S=5, kernel="matern12", desc="Matern-1/2 warp+data kernels, 2 views, P=5, M=50"),
    "c3": dict(V=4, Nv=4000, D=2, P=2000, M=200, S=8, kernel="rbf", desc="Visium-shaped synthetic: 4 views x 4k spots x 2k genes"),
    "c4": dict(V=8, Nv=10000, D=3, P=500, M=256, S=8, kernel="rbf", desc="3-D serial sections: 8 views x 10k spots x 500 genes"),
    "c5": dict(V=8, Nv=50000, D=2, P=5000, M=512, S=16, kernel="rbf", desc="large-N scaling: 8 views x 50k spots x 5k genes"),
}
GENE_BLOCK = 256  # granularity of the per-gene noise streams of make_data


# --------------------------------------------------------------------------------------------------
# synthetic data: jittered grid on [0,10]^2, views >= 1 warped by a smooth random
# field, outputs = random-Fourier-feature draws of an RBF GP + noise, z-scored per gene per view.
# Coordinates do not depend on the gene count, and every gene's column depends only on (seed, view, gene):
# a rank that owns genes [lo, hi) generates exactly its slice of the same matrix.
# --------------------------------------------------------------------------------------------------
def make_data(cfg, seed, genes=None, gene_range=None):
    rng = np.random.default_rng(seed)
    V, Nv, D = cfg["V"], cfg["Nv"], cfg["D"]
    P = genes if genes is not None else cfg["P"]
    lo, hi = gene_range if gene_range is not None else (0, P)
    side = int(math.ceil(math.sqrt(Nv)))
    base = np.stack(np.meshgrid(np.linspace(0, 10, side), np.linspace(0, 10, side)), -1).reshape(-1, 2)
    Xs, Ys = [], []
    om_w = rng.standard_normal((16, 2)) / 5.0
    ph_w = rng.uniform(0, 2 * np.pi, 16)
    om_y = rng.standard_normal((64, 2))
    ph_y = rng.uniform(0, 2 * np.pi, 64)
    W = np.random.default_rng([seed, 7]).standard_normal((64, P)).astype(np.float32)[:, lo:hi] * math.sqrt(2.0 / 64)
    for v in range(V):
        keep = rng.permutation(base.shape[0])[:Nv]
        x = base[np.sort(keep)] + rng.uniform(-0.3, 0.3, (Nv, 2)) * (10.0 / side)
        feats = np.cos(x @ om_y.T + ph_y).astype(np.float32)
        y = feats @ W
        for b in range(lo // GENE_BLOCK, (hi + GENE_BLOCK - 1) // GENE_BLOCK):
            g0, g1 = b * GENE_BLOCK, min(P, (b + 1) * GENE_BLOCK)
            nz = np.random.default_rng([seed, 100 + v, b]).standard_normal((Nv, g1 - g0), dtype=np.float32)
            a0, a1 = max(g0, lo), min(g1, hi)
            y[:, a0 - lo:a1 - lo] += 0.03 * nz[:, a0 - g0:a1 - g0]
        y = (y - y.mean(0)) / (y.std(0) + 1e-6)
        if v > 0:
            amp = rng.standard_normal((16, 2)) * 0.3 / 4.0
            x = x + np.cos(x @ om_w.T + ph_w + v) @ amp
        if D == 3:
            x = np.concatenate([x, np.full((Nv, 1), float(v))], 1)
        elif D == 1:
            x = x[:, :1]
        Xs.append(x.astype(np.float32))
        Ys.append(y.astype(np.float32))
    return np.concatenate(Xs), np.concatenate(Ys), [Nv] * V


def flops_iter(cfg, genes=None):
    """F_iter of BASELINE.md 4 (algorithmic, triangular-minimum, no credit for recomputation)."""
    V, Nv, D, M, S = cfg["V"], cfg["Nv"], cfg["D"], cfg["M"], cfg["S"]
    L = genes if genes is not None else cfg["P"]
    N = V * Nv
    Vf = V - 1
    q2 = S * N * L * M * M
    total = 3 * (q2 + 2 * S * N * M * L + 2 * S * N * M * M + Vf * Nv * M * M * (2 + D)) + 4 * (L * M**3 + V * D * M**3)
    return float(total), 3.0 * q2


# --------------------------------------------------------------------------------------------------
# CPU arm
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, seed, budget_s=25.0):
    """Fallback when baseline/_ref is absent: the oracle restatement of the reference (oracle/gpsa_oracle.py, float32,
    materialising the [S,L,N,M] tensor exactly like gpsa/models/vgpsa.py:193-196) on every host thread.
    Returns (spot_samples_per_s, description, measured_s, extrapolated)."""
    import torch

    from oracle import gpsa_oracle as orc

    torch.set_num_threads(os.cpu_count())
    V, Nv, D, M, S = cfg["V"], cfg["Nv"], cfg["D"], cfg["M"], cfg["S"]
    N = V * Nv
    full_bytes = 4.0 * S * cfg["P"] * N * M
    small = full_bytes < 2e9
    gene_counts = [cfg["P"]] if small else [1, 2, 3]
    times = []
    t_start = time.time()
    for Pg in gene_counts:
        X, Y, nl = make_data(cfg, seed, genes=Pg)
        ocfg = orc.Config(n_views=V, n_spatial_dims=D, modality_names=["expression"],
                          n_samples_lists={"expression": nl}, m_X_per_view=M, m_G=M,
                          kernel_warp=cfg["kernel"], kernel_data=cfg["kernel"], fixed_view_idx=0,
                          n_latent_gps={"expression": None})
        params = orc.init_params(ocfg, {"expression": X}, {"expression": Pg}, seed=seed, kmeans=False)
        eps = orc.draw_noise(ocfg, S, {"expression": Pg}, seed)
        reps = []
        for it in range(50):
            t0 = time.perf_counter()
            orc.elbo_and_grads(params, ocfg, {"expression": X}, {"expression": Y}, S, eps, dtype=torch.float32,
                               materialise=True)
            dt = time.perf_counter() - t0
            if it > 0 or not small:
                reps.append(dt)
            if (small and it >= 3 and time.time() - t_start > budget_s) or (not small and it >= 1):
                break
        times.append(float(np.median(reps)))
    if small:
        return S * N / times[0], f"oracle port, full {cfg['desc']}: median of {len(reps)} iterations", times[0], False
    c, t0 = np.polyfit(gene_counts, times, 1)
    c = max(float(c), 1e-6)
    t_full = float(t0) + c * cfg["P"]
    sample = (f"oracle port timed at P={gene_counts} genes ({', '.join(f'{t:.2f}s' for t in times)} per iteration; the full shape "
              f"needs a {full_bytes/1e9:.0f} GB [S,P,N,M] tensor), linear fit extrapolated to P={cfg['P']}: {t_full:.1f} s/iter")
    return S * N / t_full, sample, times[-1], True


def reference_rate(cfg, config_name, seed, steps, warmup):
    """The reference arm: the UNMODIFIED reference (baseline/_ref, installed from /root/reference with pip --target)
    driven through its own public API on this box's host cores by baseline/reference_arm.py, in a subprocess with
    CUDA hidden.  Configurations whose [S,P,N,M] tensor cannot exist (C3: 205 GB) are timed at P = 2, 4, 8 genes and
    a least-squares line t0 + c P is extrapolated to the named gene count; the result says so.
    Returns a dict: rate, kind, sample, cores, steps_run, warmup_run, measured_ms (a directly measured step: the full
    shape, or the largest timed gene count), extrapolated (bool), fit (dict or None)."""
    S, N = cfg["S"], cfg["V"] * cfg["Nv"]
    full_bytes = 4.0 * S * cfg["P"] * N * cfg["M"]
    small = full_bytes < 2e9
    genes = [cfg["P"]] if small else [2, 4, 8]
    script = os.path.join(ROOT, "baseline", "reference_arm.py")
    if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gpsa")):
        steps_run = max(1, min(steps, 50 if small else 3))
        warmup_run = max(1, min(warmup, 10 if small else 1))
        cmd = [sys.executable, script, "--config", config_name, "--genes", ",".join(map(str, genes)), "--steps", str(steps_run),
               "--warmup", str(warmup_run), "--seed", str(seed)]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
            out = json.loads(res.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            out = {"unavailable": f"reference arm failed: {e}"}
        if "runs" in out:
            t = [r["s_per_step"] for r in out["runs"]]
            base = (f"unmodified reference (baseline/_ref) through its own API, median of {steps_run} steps after {warmup_run} "
                    f"warm-up, anomaly detection off, {out['torch_threads']} threads")
            if small:
                return dict(rate=S * N / t[0], kind="reference", sample=f"{base}, full {config_name} shape", cores=out["cores"],
                            steps_run=steps_run, warmup_run=warmup_run, measured_ms=1e3 * t[0], extrapolated=False, fit=None)
            # the fit uses the FASTEST step at each gene count (least disturbed by other host activity; favours the
            # reference); a non-positive slope means the timings were disturbed beyond use -> flagged, two-point slope
            t = [r.get("s_min", r["s_per_step"]) for r in out["runs"]]
            c, t0 = np.polyfit(genes, t, 1)
            reliable = c > 0
            if not reliable:
                c = max((max(t) - min(t)) / (genes[-1] - genes[0]), 1e-6)
                t0 = min(t) - c * genes[0]
            c = float(c)
            resid = [float(ti - (t0 + c * g)) for g, ti in zip(genes, t)]
            t_full = float(t0) + c * cfg["P"]
            base = base.replace("median of", "fastest of") + ("" if reliable else " [FIT UNRELIABLE: timings not monotone in P]")
            sample = (f"{base}; timed at P={genes} genes ({', '.join(f'{ti:.2f} s' for ti in t)} per step): the full shape needs a "
                      f"{full_bytes/1e9:.0f} GB [S,P,N,M] tensor, so value = least-squares fit t0 + c*P (t0={t0:.2f} s, c={c:.4f} s/gene, "
                      f"residuals {', '.join(f'{r:+.3f}' for r in resid)} s) EXTRAPOLATED to P={cfg['P']}: {t_full:.1f} s/step")
            return dict(rate=S * N / t_full, kind="reference", sample=sample, cores=out["cores"], steps_run=steps_run,
                        warmup_run=warmup_run, measured_ms=1e3 * t[-1], extrapolated=True,
                        fit={"genes": genes, "s_per_step": t, "t0_s": float(t0), "s_per_gene": c, "residuals_s": resid,
                             "extrapolated_s_per_step": t_full})
    rate, sample, meas, extra = cpu_reference_rate(cfg, seed)
    return dict(rate=rate, kind="port", sample="baseline/_ref unavailable -> " + sample, cores=os.cpu_count(), steps_run=1,
                warmup_run=1, measured_ms=1e3 * meas, extrapolated=extra, fit=None)


def gpu_eager_reference(config_name, seed):
    """The UNMODIFIED reference run through its own eager PyTorch path on this box's GPU (C1/C2 fit): the 'same box,
    library kernels' bar of SURVEY.md 8(d).  Returns a dict or None."""
    script = os.path.join(ROOT, "baseline", "reference_arm.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gpsa")):
        return None
    try:
        res = subprocess.run([sys.executable, script, "--config", config_name, "--steps", "20", "--warmup", "5", "--seed",
                              str(seed), "--device=cuda"], capture_output=True, text=True, timeout=300)
        out = json.loads(res.stdout.strip().splitlines()[-1])
        r = out["runs"][0]
        return {"ms_per_step": 1e3 * r["s_per_step"], "steps": r["steps"],
                "what": "unmodified reference (baseline/_ref), eager PyTorch on this GPU, wall clock around step() incl. loss.item()"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


# --------------------------------------------------------------------------------------------------
def sample_clocks(stop, out):
    try:
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        idx = os.environ.get("LOCAL_RANK", "0")
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", idx],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    try:
        while not stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            out.append(line.strip())
    finally:
        p.terminate()


def summarise_clocks(lines):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [t.strip() for t in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0]))
            mx.append(float(f[1]))
        except ValueError:
            continue
        for nm, val in zip(names, f[3:7]):
            if val.lower().startswith("active"):
                reasons.add(nm)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_model(cfg, seed, genes_slice=None, device="cuda", kmeans=None, gene_range=None):
    """Construct the public-API model on synthetic data.  Inducing locations are initialised from a random
    subset of the spots instead of KMeans for the large configs (init is outside the timed region); `kmeans`
    overrides that choice.  `gene_range` = (lo, hi): generate only that slice of the outputs (gene sharding)."""
    import torch

    import gpsa

    X, Y, nl = make_data(cfg, seed, gene_range=gene_range)
    if genes_slice is not None:
        Y = np.ascontiguousarray(Y[:, genes_slice])
    kern = gpsa.rbf_kernel if cfg["kernel"] == "rbf" else gpsa.matern12_kernel
    data_dict = {"expression": {"spatial_coords": torch.from_numpy(X), "outputs": torch.from_numpy(Y), "n_samples_list": nl}}
    np.random.seed(seed)
    torch.manual_seed(seed)
    # data_init=True: host KMeans like the reference up to 20 k spots, Lloyd's iterations on the GPU above (gpsa.util.kmeans_gpu)
    use_kmeans = True if kmeans is None else bool(kmeans)
    model = gpsa.VariationalGPSA(data_dict, n_spatial_dims=cfg["D"], m_X_per_view=cfg["M"], m_G=cfg["M"],
                                 data_init=use_kmeans, n_latent_gps={"expression": None},
                                 mean_function="identity_fixed", kernel_func_warp=kern, kernel_func_data=kern,
                                 fixed_view_idx=0)
    if not use_kmeans:
        rng = np.random.default_rng(seed)
        with torch.no_grad():
            for v in range(cfg["V"]):
                sel = rng.choice(cfg["Nv"], cfg["M"], replace=False) + v * cfg["Nv"]
                model.Xtilde[v] = torch.from_numpy(X[sel])
            model.delta_G_list.copy_(model.Xtilde)
            model.Gtilde.copy_(torch.from_numpy(X[rng.choice(X.shape[0], cfg["M"], replace=False)]))
    model = model.to(device)
    return model, data_dict, X, Y, nl


def load_traffic(config_name, kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this round
    (profiles/traffic.json, written by tools/ncu_summary.py --traffic); None when that config was not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(config_name, {}).get(kernel_key)
        return (e["bytes_per_launch"], e["source"]) if e else (None, None)
    except Exception:  # noqa: BLE001
        return None, None


# --------------------------------------------------------------------------------------------------
def run_config(config_name, cfg, args, steps, warmup, use_graph, world, rank, with_e2e=True):
    """Build the model of one configuration, time `steps` iterations after `warmup`, return the measured fields."""
    import torch
    import torch.distributed as dist

    from gpsa import _lib, _ops
    from gpsa.parallel import gene_range

    S, N, P = cfg["S"], cfg["V"] * cfg["Nv"], cfg["P"]
    # what to shard: genes when every rank still fills a good part of a 256-gene MMA tile, else the
    # Monte-Carlo samples (all gradients all-reduced) -- e.g. C4's 500 genes over 8 GPUs
    # ranks form a grid: Wg gene slices x Ws sample groups (rank = g * Ws + s); Ws = 1 is pure gene sharding,
    # Ws = world pure sample sharding
    Ws = sample_groups(args, world, S, P)
    Wg = world // Ws
    by_genes = Ws == 1
    lo, hi = gene_range(P, Wg, rank // Ws)
    model, data_dict, X, Y, nl = build_model(cfg, args.seed, gene_range=(lo, hi) if Wg > 1 else None)
    sharder = None
    if world > 1:
        from gpsa import parallel

        sharder = (parallel.GeneSharding(model, world, rank) if Ws == 1 else
                   parallel.SampleSharding(model, world, rank) if Wg == 1 else
                   parallel.HybridSharding(model, world, rank, Ws))
    data_dev = {"expression": {"spatial_coords": data_dict["expression"]["spatial_coords"].cuda(),
                               "outputs": data_dict["expression"]["outputs"].cuda(), "n_samples_list": nl}}
    view_idx, Ns, _, _ = model.create_view_idx_dict(data_dev)
    use_graph = use_graph and world == 1
    from gpsa.optim import Adam

    opt = Adam(model.parameters(), lr=1e-2)  # torch.optim.Adam's rule as one launch of this library (csrc/aux.cu)
    x_dev, y_dev = data_dev["expression"]["spatial_coords"], data_dev["expression"]["outputs"]
    graphed = None
    if use_graph:
        from gpsa.graph import GraphedIteration

        torch.manual_seed(999)
        graphed = GraphedIteration(model, data_dev, opt, S)

    def step(it):
        if graphed is not None:
            return graphed.step()
        torch.manual_seed(1000 + it)
        _, _, _, F = model.forward({"expression": x_dev}, view_idx=view_idx, Ns=Ns, S=S)
        loss = model.loss_fn(data_dev, F)
        if sharder is None:
            opt.zero_grad(set_to_none=True)
            loss.backward()
        else:
            sharder.zero_grad()
            loss.backward()
            loss = sharder.allreduce(loss)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for it in range(warmup):
        step(it)
    barrier()

    # ---- timed region: K steps, CUDA events, inputs resident
    lib = _lib.lib()
    lib.gpsa_prof_enable(1)
    clock_lines, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clock_lines), daemon=True)
    th.start()
    n0 = lib.gpsa_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for it in range(steps):
        loss = step(warmup + it)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.gpsa_launch_count() - n0
    counts = (C.c_int * 4)()
    tot_ms = (C.c_double * 4)()
    lib.gpsa_prof_read(counts, tot_ms)
    lib.gpsa_prof_enable(0)
    stop.set()
    th.join(timeout=2)
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    ms_step = ms / steps
    out = {"ms_per_step": ms_step, "value": S * N / (ms_step * 1e-3), "iters_per_s": 1e3 / ms_step,
           "clocks": summarise_clocks(clock_lines), "cuda_graph": bool(use_graph),
           # under a CUDA graph the library's launchers ran once, at capture: `launches_per_replay` is that count
           "gpu_launches": int(launches) if not use_graph else int(getattr(graphed, "launches", 0)) * steps}

    # ---- e2e: same iteration from pinned host buffers + D2H read of the loss, every step
    host_loss = float(loss.item())
    if with_e2e:
        x_pin = data_dict["expression"]["spatial_coords"].pin_memory()
        y_pin = data_dict["expression"]["outputs"].pin_memory()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for it in range(steps):
            x_dev.copy_(x_pin, non_blocking=True)
            y_dev.copy_(y_pin, non_blocking=True)
            host_loss = float(step(warmup + steps + it).item())
        t1.record()
        barrier()
        te = torch.tensor([t0.elapsed_time(t1)], device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": S * N / (float(te) / steps * 1e-3), "unit": "spot-samples/s",
                      "h2d_bytes_per_step": int(x_pin.numel() * 4 + y_pin.numel() * 4), "d2h_bytes_per_step": 4}
    out["loss_last"] = host_loss

    # ---- roofline of the dominant kernels
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)" if peaks
                else "fallback (B200_PROFILING.md ~1.4 PF sustained)")
    local_genes = hi - lo
    f_iter, f_q2 = flops_iter(cfg, genes=local_genes)
    f_iter, f_q2 = f_iter / Ws, f_q2 / Ws  # this rank's share of the samples
    q_ms = sum(tot_ms[i] for i in range(3))
    q_launch = sum(counts[i] for i in range(3))
    names = ["fwd", "bwd_alpha", "bwd_omega"]
    engine = _ops.pick_engine(cfg["M"], (S // Ws) * N, local_genes)
    tc = engine in _ops.TC_ENGINES
    kern = {"fwd": "tc_gemm_kernel<3> (implicit-feature forward)", "bwd_alpha": "tc_gemm_kernel<1>",
            "bwd_omega": "tc_gemm_kernel<2>"}
    per = {n: tot_ms[i] / max(counts[i], 1) for i, n in enumerate(names)}
    f_one = f_q2 / 3.0  # algorithmic (symmetric-minimum) flops of ONE of the three products, per launch
    passes = 3 if tc else 1
    products = {n: {"kernel": kern[n] if tc else "feat_*_kernel (fp32 SIMT)", "ms_per_launch": per[n] if per[n] > 0 else None,
                    "algorithmic_tflops": f_one / (per[n] * 1e-3) / 1e12 if per[n] > 0 else None,
                    "issued_tflops": passes * f_one / (per[n] * 1e-3) / 1e12 if per[n] > 0 else None} for n in names}
    dom = max(names, key=lambda n: per[n])
    achieved = products[dom]["algorithmic_tflops"]
    traffic, traffic_src = load_traffic(config_name if (world == 1 and args.genes is None and args.samples is None) else "", dom) if tc else (None, None)
    out["roofline"] = {
        "bound": "tensor", "kernel": f"{products[dom]['kernel']} ({dom}: dominant of the three quadratic-form products)",
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
        "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
        "note": ("achieved counts ALGORITHMIC flops (M(M+1) per (sample, spot, gene)); the tcgen05 engine issues 3 bf16 MMA "
                 "passes per product for fp32-class accuracy, so the tensor pipe runs at `issued_tflops` and the ceiling of "
                 "`frac` is 1/3") if tc else "launch-latency-bound configuration: ~1e8 flop behind ~250 dependent kernels",
        "frac_issued": (passes * achieved / peak_tf) if achieved else None,
        "products": products, "share_of_step": q_ms / ms if ms > 0 else None, "launches": int(q_launch),
        "all_three_algorithmic_tflops": (f_q2 * steps) / (q_ms * 1e-3) / 1e12 if q_ms > 0 else None,
        "engine": ("tcgen05, bf16 hi/lo split x 3 passes, fp32 accumulate in TMEM" if tc else "fp32 SIMT (exact fp32 accumulate)"),
        "whole_step_tflops": f_iter / (ms_step * 1e-3) / 1e12,
        "whole_step_frac": f_iter / (ms_step * 1e-3) / 1e12 / peak_tf,
    }
    out["sharding"] = ("none" if world == 1 else
                       f"genes: {P} outputs split over {world} ranks, shared front end replicated, one NCCL all-reduce of shared-parameter grads"
                       if by_genes else
                       f"samples: the {S} Monte-Carlo samples split over {world} ranks, every rank holds all {P} outputs, one NCCL all-reduce of all grads"
                       if Wg == 1 else
                       f"hybrid: {Wg} gene slices x {Ws} sample groups; gene-local grads all-reduced inside a slice's {Ws} replicas, "
                       f"shared grads over all {world} ranks")
    out["rank_grid"] = {"gene_slices": Wg, "sample_groups": Ws}
    del model, opt, graphed, sharder, data_dev
    torch.cuda.empty_cache()
    return out


def sample_groups(args, world, S, P):
    """How many of the `world` ranks share one gene slice and split the Monte-Carlo samples between them."""
    if world == 1:
        return 1
    if args.sharding == "genes":
        return 1
    if args.sharding == "samples":
        return world
    if args.sample_groups:
        Ws = args.sample_groups
        if world % Ws or S % Ws:
            raise SystemExit(f"--sample-groups {Ws} must divide the world size {world} and S = {S}")
        return Ws
    # auto: genes while every rank still fills a good part of a 256-gene MMA tile, else the samples
    return world if (P // world < 128 and S % world == 0) else 1


def workload_of(name, cfg):
    return (f"{name}: {cfg['desc']}, D={cfg['D']}, M_X=M_G={cfg['M']}, S={cfg['S']}, {cfg['kernel']}, "
            f"fixed_view_idx=0; one step = forward + loss_fn + backward + Adam.step")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="default: c3 as the headline plus other_configs (c1, c2, c4) at one GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip other_configs")
    ap.add_argument("--graph", action="store_true",
                    help="replay the whole iteration from one CUDA graph (gpsa.graph.GraphedIteration); for the "
                         "launch-bound toy configurations c1/c2")
    ap.add_argument("--sample-groups", type=int, default=0,
                    help="with --sharding auto: ranks per gene slice that split the Monte-Carlo samples (hybrid grid)")
    ap.add_argument("--sharding", default="auto", choices=["auto", "genes", "samples"],
                    help="multi-GPU partition: output genes, Monte-Carlo samples, or auto (samples when a rank would own < 128 genes)")
    ap.add_argument("--genes", type=int, default=None,
                    help="profiling aid: run with this many output genes (e.g. P/8 to see one rank of an 8-GPU run); "
                         "the JSON line is then NOT the named configuration and says so")
    ap.add_argument("--samples", type=int, default=None,
                    help="profiling aid like --genes: this many Monte-Carlo samples (S/k = one rank of a k-way sample split)")
    ap.add_argument("--engine", type=int, default=None,
                    help="quadratic-form engine: 0 fp32 SIMT, 1 tcgen05 (default: auto)")
    args = ap.parse_args()
    config_name = args.config or "c3"
    cfg = dict(CONFIGS[config_name])
    if args.genes is not None:
        cfg["P"] = args.genes
        cfg["desc"] += f" [REDUCED to {args.genes} genes: profiling aid, not the named configuration]"
    if args.samples is not None:
        cfg["S"] = args.samples
        cfg["desc"] += f" [REDUCED to S={args.samples}: profiling aid, not the named configuration]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    S, N = cfg["S"], cfg["V"] * cfg["Nv"]
    workload = workload_of(config_name, cfg)

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_rate(cfg, config_name, args.seed, args.steps, args.warmup)
        line = {
            "impl": "reference", "metric": "spot_samples_per_s", "value": r["rate"], "unit": "spot-samples/s",
            "n_gpus": args.gpus, "steps": r["steps_run"], "warmup": r["warmup_run"],
            "requested_steps": args.steps, "requested_warmup": args.warmup,
            # ms_per_step is a MEASURED step (the full shape, or the largest gene count that was timed); `value` of a
            # configuration the reference cannot materialise comes from the fit and is flagged as extrapolated
            "ms_per_step": r["measured_ms"], "value_is_extrapolated": r["extrapolated"],
            "extrapolated_ms_per_step": (1e3 * S * N / r["rate"]) if r["extrapolated"] else None, "fit": r["fit"],
            "iters_per_s": r["rate"] / (S * N), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload},
            "cpu_baseline": {"value": r["rate"], "unit": "spot-samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["rate"], "unit": "spot-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from gpsa import _ops

    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries the ONE JSON line: NCCL's own log (NCCL_DEBUG=INFO/VERSION, whatever the caller set) goes to
        # stderr instead of being silenced, so rank / transport lines stay countable
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.engine is not None:
        _ops.ENGINE["value"] = args.engine

    main_res = run_config(config_name, cfg, args, args.steps, args.warmup, args.graph, world, rank)

    others = None
    if args.config is None and world == 1 and not args.no_others and args.genes is None and args.samples is None:
        others = []
        for name, k, w, graph in (("c1", 50, 10, True), ("c2", 50, 10, True), ("c4", min(args.steps, 5), 3, False)):
            ocfg = dict(CONFIGS[name])
            try:
                r = run_config(name, ocfg, args, k, w, graph, 1, 0, with_e2e=False)
                entry = {"config": {"workload": workload_of(name, ocfg)}, "steps": k, "warmup": w,
                         "ms_per_step": r["ms_per_step"], "value": r["value"], "unit": "spot-samples/s",
                         "iters_per_s": r["iters_per_s"], "cuda_graph": r["cuda_graph"], "roofline": r["roofline"],
                         "clocks": r["clocks"], "loss_last": r["loss_last"]}
                if graph:
                    entry["gpu_eager_baseline"] = gpu_eager_reference(name, args.seed)
            except Exception as e:  # noqa: BLE001
                entry = {"config": {"workload": workload_of(name, ocfg)}, "error": f"{type(e).__name__}: {e}"[:300]}
            others.append(entry)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N = 1 only
        r = reference_rate(cfg, config_name, args.seed, 2, 1)
        cpu = {"value": r["rate"], "unit": "spot-samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "value_is_extrapolated": r["extrapolated"]}
    line = {
        "metric": "spot_samples_per_s", "value": main_res["value"], "unit": "spot-samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "iters_per_s": main_res["iters_per_s"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload,
                   "arithmetic": ("fp32 data; quadratic form = bf16 hi/lo split, 3 tcgen05 passes, fp32 accumulate; prior K_uu / warp-layer MxM algebra fp64; "
                                  "gene-batched Omega_F: fp64-accumulated SYRK, fp32 factorisation, fp64 log-det"),
                   "l2": "working set (Omega_sqt 320 MB, two [S,N,L] buffers of 1 GB each at c3) far exceeds the 126 MB L2",
                   "noise": "eps_F drawn in-kernel (Philox4x32-10 keyed by seed, sample, spot, global gene); F_samples never materialised (fused sampling + likelihood)",
                   "optimizer": "gpsa.optim.Adam(lr=1e-2): torch.optim.Adam's update as one launch of this library", "cuda_graph": main_res["cuda_graph"],
                   "sharding": main_res["sharding"], "rank_grid": main_res.get("rank_grid")},
        "e2e": main_res.get("e2e"), "gpu_launches": main_res["gpu_launches"], "roofline": main_res["roofline"],
        "cpu_baseline": cpu, "clocks": main_res["clocks"], "loss_last": main_res["loss_last"],
    }
    if others is not None:
        line["other_configs"] = others
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
