#!/usr/bin/env python
"""Times the three quadratic-form products (either engine) at a given shape with CUDA events, through the
C ABI.  python tools/bench_quadform.py [--M 200 --R 128000 --L 2000 --engine tc|simt --reps 5]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spatial-alignment_b200"))
import torch  # noqa: E402

from gpsa import _lib, _ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=200)
    ap.add_argument("--R", type=int, default=128000)
    ap.add_argument("--L", type=int, default=2000)
    ap.add_argument("--engine", default="tc")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--which", default="fwd,alpha,omega")
    a = ap.parse_args()
    M, R, L = a.M, a.R, a.L
    lib = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, R, device="cuda", generator=g) * 0.3
    Osq = torch.randn(L, M, M, device="cuda", generator=g) * 0.1
    G = torch.randn(R, L, device="cuda", generator=g)
    Omega, Ltril, L64, hld, info = _ops.omega_prepare(Osq)
    del L64
    q2 = torch.empty(R, L, device="cuda")
    nf = _lib.feat_count(M)
    H = torch.empty(nf, L, device="cuda")
    Abar = torch.zeros(M, R, device="cuda")
    ws = _lib.tc_workspace(M, R, L, A)
    W = torch.empty(nf, L, device="cuda")
    lib.gpsa_feat_pack(M, L, Omega.data_ptr(), W.data_ptr(), st)
    flop = 2.0 * R * L * (M * (M + 1) / 2)  # symmetric-minimum flops of one product

    def run(name):
        if a.engine == "tc":
            if name == "fwd":
                return lib.gpsa_quadform_fwd_feat_tc(M, R, L, A.data_ptr(), Omega.data_ptr(), q2.data_ptr(), ws.data_ptr(), ws.numel(), st)
            if name == "alpha":
                return lib.gpsa_quadform_bwd_alpha_tc(M, R, L, A.data_ptr(), G.data_ptr(), Omega.data_ptr(), Abar.data_ptr(), ws.data_ptr(), ws.numel(), st)
            return lib.gpsa_quadform_bwd_omega_tc(M, R, L, A.data_ptr(), G.data_ptr(), H.data_ptr(), ws.data_ptr(), ws.numel(), st)
        if name == "fwd":
            return lib.gpsa_quadform_fwd_f32(M, R, L, A.data_ptr(), W.data_ptr(), q2.data_ptr(), st)
        if name == "alpha":
            return lib.gpsa_quadform_bwd_alpha_f32(M, R, L, A.data_ptr(), G.data_ptr(), W.data_ptr(), Abar.data_ptr(), st)
        return lib.gpsa_quadform_bwd_omega_f32(M, R, L, A.data_ptr(), G.data_ptr(), H.data_ptr(), st)

    for name in a.which.split(","):
        assert run(name) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            assert run(name) == 0
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"{a.engine:5s} {name:6s} M={M} R={R} L={L}: {ms:8.2f} ms  {flop/ms/1e9:8.1f} TFLOP/s algorithmic (incl. operand packing)", flush=True)


if __name__ == "__main__":
    main()
