"""CUDA-graph capture of one whole ELBO iteration (forward + loss_fn + backward + optimizer.step).

The toy configurations (examples/grid_example.py: 200 spots, 30 genes, M = 25) are ~1 microsecond of arithmetic behind
~250 dependent kernel launches, i.e. bound by launch latency and host-side Python.  Every op of the path is
asynchronous on the current stream, allocates through torch's caching allocator and draws its noise from the
graph-safe CUDA generator, so the iteration can be captured once and replayed:

    opt = torch.optim.Adam(model.parameters(), lr=1e-2, capturable=True)
    it = GraphedIteration(model, data_dict, opt, S=5)
    for _ in range(n): loss = it.step()            # loss: 0-dim CUDA tensor, valid until the next step()

New data of the same shapes can be fed by copying into the tensors of `data_dict` in place.
"""
import torch


class GraphedIteration:
    def __init__(self, model, data_dict, optimizer, S, warmup=3):
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise ValueError("the optimizer must be created with capturable=True to be captured in a CUDA graph")
        self.model, self.data_dict, self.optimizer, self.S = model, data_dict, optimizer, int(S)
        self.view_idx, self.Ns, _, _ = model.create_view_idx_dict(data_dict)
        self.X = {m: data_dict[m]["spatial_coords"] for m in model.modality_names}
        # warm up off the default stream (allocator pools, cudaFuncSetAttribute, index/mask caches), then capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib

        n0 = _lib.lib().gpsa_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._iteration()
        # kernels of THIS library inside one replay (its launchers ran exactly once, during capture)
        self.launches = int(_lib.lib().gpsa_launch_count() - n0)

    def _iteration(self):
        out = self.model.forward(self.X, view_idx=self.view_idx, Ns=self.Ns, S=self.S)
        loss = self.model.loss_fn(self.data_dict, out[3])
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.optimizer.step()
        return loss

    def step(self):
        self.graph.replay()
        return self.loss
