"""Adam for the ELBO iteration as one kernel launch over all parameter tensors (csrc/aux.cu: gpsa_adam_step).

Same update rule and state as torch.optim.Adam(params, lr, betas, eps) with weight_decay = 0 and amsgrad = False -- what
every training loop of the reference uses (examples/grid_example.py:59) -- so runs are interchangeable step for step.
The step count lives on the device, so `step()` can be captured in a CUDA graph (gpsa.graph.GraphedIteration)
without the `capturable=True` variant's extra kernels.
"""
import torch

from . import _lib


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, capturable=True))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.requires_grad]
            if not params:
                continue
            dev = params[0].device
            if dev.type != "cuda":
                raise _lib.GPSALibraryError("gpsa.optim.Adam is CUDA-only (no CPU fallback exists)")
            steps = group.setdefault("step", {})
            keep = []  # temporaries whose pointers sit in a launch that has not been enqueued yet
            for i in range(0, len(params), _lib.ADAM_MAX_TENSORS):
                chunk = params[i:i + _lib.ADAM_MAX_TENSORS]
                if i not in steps:  # device-side step counts, one per tensor of this launch
                    steps[i] = torch.zeros(_lib.ADAM_MAX_TENSORS, dtype=torch.float32, device=dev)
                grads, ms, vs = [], [], []
                for p in chunk:
                    st = self.state[p]
                    if not st:
                        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    g = p.grad
                    if g is not None and (not g.is_contiguous() or g.dtype != torch.float32):
                        g = g.contiguous().float()
                        keep.append(g)
                    grads.append(g)
                    ms.append(st["exp_avg"])
                    vs.append(st["exp_avg_sq"])
                _lib.ops().adam_step([p.data for p in chunk], grads, ms, vs, float(group["lr"]), float(group["betas"][0]),
                                     float(group["betas"][1]), float(group["eps"]), steps[i])
            del keep
        return loss
