"""Norm selectors with the reference's names (paddlexde/utils/ode_utils.py:4-19).

In the reference these are Python callables applied to tensors; a fused kernel cannot call back
into Python, so here they are *selectors*: the solver recognises the object and runs the matching
device-side reduction.  Passing any other callable as options["norm"] raises UnsupportedFieldError."""


def _linf_norm(tensor):  # utils/ode_utils.py:4-5 (used by the out-of-scope Adams solver only, adams.py:499)
    from .. import _tensor as T
    import numpy as np

    return float(np.abs(T.to_host(tensor)).max()) if not T.is_torch(tensor) else tensor.abs().max()


def _rms_norm(tensor):  # utils/ode_utils.py:8-9
    from .. import _tensor as T

    if T.is_torch(tensor):
        return tensor.abs().pow(2).mean().sqrt()
    import numpy as np

    return float(np.sqrt(np.mean(np.square(np.abs(T.to_host(tensor)), dtype=np.float64))))


def _mixed_norm(tensor_tuple):  # utils/ode_utils.py:16-19
    if len(tensor_tuple) == 0:
        return 0.0
    return max([_rms_norm(t) for t in tensor_tuple])


_linf_norm.xde_norm = "linf"  # a selector the fused controllers do not implement: check_norm refuses it loudly
_rms_norm.xde_norm = "rms"
_mixed_norm.xde_norm = "mixed"
