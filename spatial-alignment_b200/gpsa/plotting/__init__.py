"""Plotting callbacks are visualisation only and outside the hot path.
The names of reference gpsa/plotting/__init__.py:1-6 are kept importable; calling one needs
matplotlib/seaborn and is not provided by this build."""


def _unavailable(name):
    def f(*args, **kwargs):
        raise NotImplementedError(
            f"gpsa.plotting.{name} is a matplotlib visualisation helper of the reference and is out of scope "
            "of the B200 hot-path build; use the reference's gpsa/plotting/callbacks.py for plots."
        )

    f.__name__ = name
    return f


callback_oned = _unavailable("callback_oned")
callback_twod = _unavailable("callback_twod")
callback_twod_aligned_only = _unavailable("callback_twod_aligned_only")
callback_twod_multimodal = _unavailable("callback_twod_multimodal")
