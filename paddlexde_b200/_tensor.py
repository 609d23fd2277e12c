"""Device-buffer plumbing behind the solver classes: ONE small surface, two providers.

  * torch tensors (the caller lives in PyTorch): torch supplies HBM allocations, the current stream and autograd;
    buffers cross to the C ABI as raw pointers (`data_ptr()`).
  * everything else -- numpy arrays, lists, `DeviceArray` -- goes through `_native` (ctypes + numpy over
    csrc/xde_hostapi.cu): no tensor library at all.  This is the provider when torch is not installed, and it is
    chosen per call by the type of the inputs, so a PyTorch-free process can `import paddlexde_b200` and solve.
Nothing numerical happens here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native
from ._native import DeviceArray

try:  # optional: only as the provider for callers that hand in torch tensors
    import torch
except Exception:  # pragma: no cover - exercised by tests/test_torch_free.py in a subprocess
    torch = None

_NP = {"f32": np.float32, "i32": np.int32, "i64": np.int64}


def is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def device():
    if torch is None or not torch.cuda.is_available():
        raise RuntimeError("paddlexde_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def is_host(x) -> bool:
    if isinstance(x, (np.ndarray, list, tuple, float, int)):
        return True
    if is_torch(x):
        return not x.is_cuda
    return False


def prefer_torch(*xs) -> bool:
    """Provider for this call: torch if any input is a torch tensor."""
    return any(is_torch(x) for x in xs)


def to_dev(x, dtype=None, like=None):
    """Contiguous fp32 device view/copy of x.  Provider: native (DeviceArray) when x or `like` is a DeviceArray or when
    PyTorch is not installed; a torch CUDA tensor otherwise (numpy in a PyTorch process rides on torch's allocator
    and current stream, as in round 1)."""
    native = isinstance(x, DeviceArray) or isinstance(like, DeviceArray) or torch is None
    if not native:
        dt = torch.float32 if dtype is None else dtype
        if is_torch(x):
            t = x
        elif isinstance(x, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(x))
        elif hasattr(x, "__dlpack__"):
            t = torch.from_dlpack(x)
        else:
            t = torch.as_tensor(x)
        if t.dtype != dt:
            t = t.to(dt)
        if not t.is_cuda:
            t = t.to(like.device if is_torch(like) and like.is_cuda else device(), non_blocking=True)
        return t.detach().contiguous()
    if isinstance(x, DeviceArray):
        return x if x.dtype == np.float32 else _native.asarray(x)
    return _native.asarray(to_host(x), np.float32)


def empty(shape, like, kind="f32"):
    if is_torch(like):
        return torch.empty(tuple(shape), device=like.device, dtype={"f32": torch.float32, "i32": torch.int32, "i64": torch.int64}[kind])
    return DeviceArray(tuple(shape), _NP[kind])


def zeros(shape, like, kind="f32"):
    if is_torch(like):
        return torch.zeros(tuple(shape), device=like.device, dtype={"f32": torch.float32, "i32": torch.int32, "i64": torch.int64}[kind])
    return DeviceArray.zeros(tuple(shape), _NP[kind])


def from_host(a: np.ndarray, like):
    """numpy -> device buffer of the provider `like` belongs to."""
    if is_torch(like):
        return torch.from_numpy(np.ascontiguousarray(a)).to(like.device)
    return DeviceArray.from_numpy(a)


def to_host(x) -> np.ndarray:
    if is_torch(x):
        return x.detach().cpu().numpy()
    if isinstance(x, DeviceArray):
        return x.numpy()
    return np.asarray(x)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def stream(like=None) -> C.c_void_p:
    """The stream the call is ordered on: torch's current stream when the buffers are torch tensors (or when no buffer
    is named and torch is there), the default stream for DeviceArrays."""
    if isinstance(like, DeviceArray) or torch is None or not torch.cuda.is_available():
        return C.c_void_p(0)
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def like_input(result, template):
    """Return `result` in the container family of `template` (numpy in -> numpy out)."""
    if isinstance(template, np.ndarray) or isinstance(template, (list, tuple)):
        return to_host(result)
    if is_torch(template) and not template.is_cuda:
        return result.cpu() if is_torch(result) else torch.from_numpy(to_host(result))
    return result
