#!/usr/bin/env python
"""Times omega_prepare (Omega = Osq Osq^T + eps I, fp64 Cholesky) and omega_grad (2 (Obar + c Omega^-1) Osq) at a
given batch/M -- the batched fp64 M x M algebra of the iteration."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spatial-alignment_b200"))
import torch  # noqa: E402

from gpsa import _ops  # noqa: E402

B, M = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, int(sys.argv[2]) if len(sys.argv) > 2 else 200
g = torch.Generator(device="cuda").manual_seed(0)
Osq = torch.randn(B, M, M, device="cuda", generator=g) * 0.1
Obar = torch.randn(B, M, M, device="cuda", generator=g)
Obar = (Obar + Obar.transpose(1, 2)).contiguous()
coef = torch.full((B,), -0.5, device="cuda")


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


t1, pre = timeit(lambda: _ops.omega_prepare(Osq))
t2, _ = timeit(lambda: _ops.omega_grad(Osq, pre[2], Obar, coef, tc=True))
t3, _ = timeit(lambda: _ops.omega_grad(Osq, pre[2], Obar, None, tc=True))
print(f"B={B} M={M} tile={os.environ.get('GPSA_F64_TILE','auto')}: omega_prepare {t1:.2f} ms, omega_grad(with logdet term) {t2:.2f} ms, omega_grad(tc product only) {t3:.2f} ms", flush=True)
