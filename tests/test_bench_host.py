"""Host-side logic of bench.py (CPU, no GPU): the synthetic-data recipe, the flop model, the rank-grid rule, the clock
parser and the reference arm's JSON contract.  The timed GPU path itself is covered by the -m gpu tests and the bench
lines under profiles/."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_cfg(**kw):
    cfg = dict(V=3, Nv=50, D=2, P=600, M=10, S=2, kernel="rbf", desc="test")
    cfg.update(kw)
    return cfg


def test_make_data_is_deterministic_and_shaped():
    cfg = small_cfg(P=7)
    X1, Y1, nl1 = bench.make_data(cfg, seed=3)
    X2, Y2, nl2 = bench.make_data(cfg, seed=3)
    assert np.array_equal(X1, X2) and np.array_equal(Y1, Y2) and nl1 == nl2 == [50, 50, 50]
    assert X1.shape == (150, 2) and Y1.shape == (150, 7) and X1.dtype == Y1.dtype == np.float32
    X3, Y3, _ = bench.make_data(cfg, seed=4)
    assert not np.array_equal(Y1, Y3)
    # z-scored per gene per view
    for v in range(3):
        blk = Y1[50 * v:50 * (v + 1)]
        assert np.allclose(blk.mean(0), 0, atol=1e-5) and np.allclose(blk.std(0), 1, atol=1e-3)


@pytest.mark.parametrize("D", [1, 2, 3])
def test_make_data_dimensions(D):
    X, Y, _ = bench.make_data(small_cfg(D=D, P=3), seed=1)
    assert X.shape == (150, D)
    if D == 3:  # serial sections: the third coordinate is the view index
        assert np.array_equal(X[:, 2], np.repeat(np.arange(3, dtype=np.float32), 50))


@pytest.mark.parametrize("lo,hi", [(0, 256), (256, 600), (100, 300), (511, 513)])
def test_gene_slice_equals_slice_of_full_matrix(lo, hi):
    """A rank that owns genes [lo, hi) generates its slice of the SAME matrix (coordinates identical, columns equal to
    fp32 round-off of the feature GEMM), also when the slice cuts through the 256-gene noise blocks."""
    cfg = small_cfg()
    Xf, Yf, _ = bench.make_data(cfg, seed=3)
    Xs, Ys, _ = bench.make_data(cfg, seed=3, gene_range=(lo, hi))
    assert np.array_equal(Xf, Xs)
    assert Ys.shape == (150, hi - lo)
    assert np.allclose(Ys, Yf[:, lo:hi], rtol=0, atol=2e-5)


def test_coordinates_do_not_depend_on_gene_count():
    Xa, _, _ = bench.make_data(small_cfg(P=2), seed=3)
    Xb, _, _ = bench.make_data(small_cfg(P=600), seed=3)
    Xc, _, _ = bench.make_data(small_cfg(), seed=3, genes=8)
    assert np.array_equal(Xa, Xb) and np.array_equal(Xa, Xc)


def test_flops_model_matches_baseline_formula():
    """the three quadratic-form products are 3 S N L M^2 and dominate C3-C5."""
    for name in ("c3", "c4", "c5"):
        cfg = bench.CONFIGS[name]
        total, q2 = bench.flops_iter(cfg)
        S, N, L, M = cfg["S"], cfg["V"] * cfg["Nv"], cfg["P"], cfg["M"]
        assert q2 == 3.0 * S * N * L * M * M
        assert 0.95 < q2 / total < 1.0
    c3 = bench.CONFIGS["c3"]
    assert bench.flops_iter(c3)[1] == pytest.approx(3.072e13)
    # gene sharding: the products scale with the local gene count
    assert bench.flops_iter(c3, genes=250)[1] == pytest.approx(3.072e13 / 8)


def test_rank_grid_rule():
    def args(sharding="auto", sample_groups=0):
        return types.SimpleNamespace(sharding=sharding, sample_groups=sample_groups)

    assert bench.sample_groups(args(), 1, 8, 2000) == 1
    # C3 on 8 GPUs: 250 genes per rank still fill a 256-gene MMA tile -> genes
    assert bench.sample_groups(args(), 8, 8, 2000) == 1
    # C4 on 8 GPUs: 62 genes per rank -> the Monte-Carlo samples
    assert bench.sample_groups(args(), 8, 8, 500) == 8
    assert bench.sample_groups(args(), 2, 8, 500) == 1
    # S not divisible by the world: genes even when the slices are thin
    assert bench.sample_groups(args(), 3, 8, 300) == 1
    assert bench.sample_groups(args("genes"), 8, 8, 500) == 1
    assert bench.sample_groups(args("samples"), 4, 8, 2000) == 4
    assert bench.sample_groups(args(sample_groups=2), 8, 8, 2000) == 2
    with pytest.raises(SystemExit):
        bench.sample_groups(args(sample_groups=3), 8, 8, 2000)


def test_clock_summary_parses_nvidia_smi_lines():
    lines = ["1560, 1965, 998.1, Not Active, Not Active, Not Active, Active",
             "1575, 1965, 1001.3, Not Active, Not Active, Not Active, Active",
             "1965, 1965, 310.0, Not Active, Not Active, Not Active, Not Active",
             "garbage", "[N/A], 1965, 1, Not Active, Not Active, Not Active, Not Active"]
    c = bench.summarise_clocks(lines)
    assert c == {"sm_mhz": 1575.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    assert bench.summarise_clocks([]) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    hot = bench.summarise_clocks(["1200, 1965, 900, Active, Active, Not Active, Not Active"])
    assert hot["reasons"] == ["hw_slowdown", "hw_thermal_slowdown"]


def test_workload_names_the_configuration():
    w = bench.workload_of("c3", bench.CONFIGS["c3"])
    assert w.startswith("c3: ") and "M_X=M_G=200" in w and "S=8" in w and "forward + loss_fn + backward + Adam.step" in w


def test_traffic_table_covers_the_headline_products():
    for k in ("fwd", "bwd_alpha", "bwd_omega"):
        b, src = bench.load_traffic("c3", k)
        assert b and b > 1e9 and "ncu" in src
    assert bench.load_traffic("", "fwd") == (None, None)


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gpsa")),
                    reason="baseline/_ref not installed in this checkout")
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on the smallest named configuration: the unmodified reference really runs (CPU), and
    the line carries the keys the driver reads, the steps it actually ran and no extrapolation for a shape that fits."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600,
                         env={**os.environ, "WORLD_SIZE": "1", "RANK": "0"})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spot_samples_per_s" and d["unit"] == "spot-samples/s"
    assert d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["steps"] == 3 and d["warmup"] == 1 and d["requested_steps"] == 3
    assert d["value_is_extrapolated"] is False and d["extrapolated_ms_per_step"] is None and d["fit"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "spot-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    S, N = 5, 200
    assert d["value"] == pytest.approx(S * N / (d["ms_per_step"] * 1e-3), rel=1e-6)
    assert d["config"]["workload"] == bench.workload_of("c1", bench.CONFIGS["c1"])


def test_reference_arm_is_silent_on_other_ranks():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120,
                         env={**os.environ, "WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""
