"""LazySamples (gpsa/lazy.py) on CPU: the handle forward() returns in place of the [S,N,P] sample tensors."""
import numpy as np
import torch

from gpsa.lazy import LazySamples


def _handle(calls):
    base = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4).requires_grad_()

    def produce():
        calls.append(1)
        return base * 2.0

    return LazySamples((2, 3, 4), torch.float32, torch.device("cpu"), produce, "test"), base


def test_metadata_does_not_materialise():
    calls = []
    h, _ = _handle(calls)
    assert tuple(h.shape) == (2, 3, 4) and h.size(1) == 3 and h.dim() == 3 and len(h) == 2
    assert h.dtype == torch.float32 and h.device.type == "cpu" and "lazy" in repr(h)
    assert calls == [] and not h.is_materialised


def test_any_use_materialises_once_with_autograd():
    calls = []
    h, base = _handle(calls)
    a = h.detach().numpy()                 # attribute access
    b = h[1, 2]                            # indexing
    c = torch.mean(h, dim=0)               # torch function
    d = (h + 1.0) * h - 2 / (h + 1)        # arithmetic, both sides
    e = torch.cat([h, h], dim=0)           # inside a list argument
    assert calls == [1] and h.is_materialised
    assert a.shape == (2, 3, 4) and b.shape == (4,) and c.shape == (3, 4) and d.shape == (2, 3, 4) and e.shape[0] == 4
    np.testing.assert_allclose(np.asarray(h), a)
    c.sum().backward()
    assert base.grad is not None and float(base.grad.sum()) == 24.0 * 2.0 / 2.0
    assert "materialised" in repr(h)


def test_private_names_are_not_forwarded():
    h, _ = _handle([])
    try:
        h._nonexistent
    except AttributeError:
        pass
    else:
        raise AssertionError("private attribute lookup must not materialise / forward")
    assert not h.is_materialised
