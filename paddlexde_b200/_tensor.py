"""Device-buffer plumbing: torch is used for HBM allocations, streams and DLPack interchange only."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("paddlexde_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def is_host(x) -> bool:
    if isinstance(x, np.ndarray) or isinstance(x, (list, tuple, float, int)):
        return True
    if isinstance(x, torch.Tensor):
        return not x.is_cuda
    return False


def to_dev(x, dtype=torch.float32) -> torch.Tensor:
    """Contiguous fp32 CUDA tensor view/copy of x (torch / numpy / anything exporting DLPack)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(x)
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.to(device(), non_blocking=True)
    return t.detach().contiguous()


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def like_input(result: torch.Tensor, template):
    """Return `result` in the container family of `template` (numpy in -> numpy out)."""
    if isinstance(template, np.ndarray) or isinstance(template, (list, tuple)):
        return result.cpu().numpy()
    if isinstance(template, torch.Tensor) and not template.is_cuda:
        return result.cpu()
    return result
