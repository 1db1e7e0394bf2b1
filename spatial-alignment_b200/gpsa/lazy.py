"""Lazy handle for the data GP's Monte-Carlo samples.

The reference's API passes the samples by value: forward() returns F_samples {modality: [S,N,P] tensor} and the
caller hands that dict to loss_fn(data_dict, F_samples) (reference gpsa/models/vgpsa.py:479-491,
examples/grid_example.py:66-71).  At the benchmark shapes those tensors are 1 GB (C3) to 128 GB (C5), and nothing in a
training loop reads them except loss_fn -- so forward() returns this handle instead, and

  * loss_fn(data_dict, F_samples) recognises the handle of ITS OWN last forward and runs the fused
    sampling + likelihood kernel (no F, eps or variance tensor is ever written; SURVEY.md 7.4 "RNG parity vs memory");
  * any other use -- indexing, .detach().cpu(), torch.mean(F), arithmetic -- materialises the samples as an ordinary
    tensor with the ordinary autograd graph behind it, once, and from then on the handle is that tensor.

The handle is not a torch.Tensor subclass; it forwards attribute access, indexing, arithmetic and torch.* functions
(__torch_function__) to the materialised tensor.  `shape`, `dtype`, `device`, `dim()` and `size()` answer without
materialising.
"""
import torch


class LazySamples:
    def __init__(self, shape, dtype, device, produce, describe=""):
        self._shape = torch.Size(shape)
        self._dtype, self._device = dtype, device
        self._produce = produce      # () -> tensor; called at most once
        self._tensor = None
        self._fused = None           # set by the model: payload of the fused path (buffers, noise, owner, generation)
        self._describe = describe

    # ---- answers that do not need the values -------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    @property
    def is_materialised(self):
        return self._tensor is not None

    def size(self, dim=None):
        return self._shape if dim is None else self._shape[dim]

    def dim(self):
        return len(self._shape)

    def __len__(self):
        return self._shape[0]

    def __repr__(self):
        state = "materialised" if self._tensor is not None else "lazy"
        return f"LazySamples({tuple(self._shape)}, {self._dtype}, {self._device}, {state}{', ' + self._describe if self._describe else ''})"

    # ---- everything else goes through the tensor ------------------------------------------------------
    def materialise(self):
        if self._tensor is None:
            self._tensor = self._produce()
            self._produce = None
        return self._tensor

    def __getattr__(self, name):  # only reached for names not defined above
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialise(), name)

    def __getitem__(self, idx):
        return self.materialise()[idx]

    def __iter__(self):
        return iter(self.materialise())

    def __array__(self, dtype=None):
        a = self.materialise().detach().cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def unwrap(x):
            if isinstance(x, LazySamples):
                return x.materialise()
            if isinstance(x, (list, tuple)):
                return type(x)(unwrap(v) for v in x)
            if isinstance(x, dict):
                return {k: unwrap(v) for k, v in x.items()}
            return x

        return func(*unwrap(args), **unwrap(kwargs or {}))


def _binary(name):
    def op(self, other):
        other = other.materialise() if isinstance(other, LazySamples) else other
        return getattr(self.materialise(), name)(other)
    op.__name__ = name
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__matmul__", "__rmatmul__", "__pow__", "__lt__", "__le__", "__gt__", "__ge__", "__eq__", "__ne__"):
    setattr(LazySamples, _n, _binary(_n))
LazySamples.__neg__ = lambda self: -self.materialise()
LazySamples.__hash__ = object.__hash__
