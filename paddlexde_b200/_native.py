"""Torch-free device buffers: ctypes + numpy over the host surface of libxde_b200 (csrc/xde_hostapi.cu).

north_star boundary clause: "Python host code hands tensors to CUDA via DLPack through a thin C-ABI layer (ctypes ...),
with no PyTorch".  A `DeviceArray` is a contiguous device buffer from the library's stream-ordered pool:
numpy in (`from_numpy`), numpy out (`numpy()`), and zero-copy export to Paddle / PyTorch / CuPy through
`__dlpack__` (DLPack v0.8 capsule built by `xde_dlpack_wrap`) or `__cuda_array_interface__`.  The solver classes
accept it wherever they accept a tensor; numpy inputs go through it when torch is not installed."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

_DT = {np.dtype(np.float32): (2, 32), np.dtype(np.int32): (0, 32), np.dtype(np.int64): (0, 64), np.dtype(np.float64): (2, 64)}


class _Buffer:
    """Owner of one device allocation (freed with the last view / DLPack consumer that shares it)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p(0)
        check(lib().xde_malloc(C.byref(p), max(int(nbytes), 1), None))
        self.ptr = p.value or 0
        self.owned = True

    def __del__(self):
        try:
            if self.owned and self.ptr:
                lib().xde_free(C.c_void_p(self.ptr), None)
        except Exception:
            pass


class DeviceArray:
    is_cuda = True

    def __init__(self, shape, dtype=np.float32, _buf=None, _ptr=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        if self.dtype not in _DT:
            raise TypeError(f"DeviceArray holds float32 / float64 / int32 / int64, not {self.dtype}")
        n = 1
        for s in self.shape:
            n *= s
        self._n = n
        self._buf = _buf if _buf is not None else _Buffer(n * self.dtype.itemsize)
        self._ptr = self._buf.ptr if _ptr is None else _ptr
        d = C.c_int32(0)
        check(lib().xde_get_device(C.byref(d)))
        self.device = ("native", int(d.value))

    # -- construction / transfer --------------------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, a) -> "DeviceArray":
        a = np.ascontiguousarray(a)
        out = cls(a.shape, a.dtype)
        check(lib().xde_memcpy_async(C.c_void_p(out._ptr), a.ctypes.data_as(C.c_void_p), a.nbytes, 0, None))
        check(lib().xde_stream_synchronize(None))  # the source may be a temporary
        return out

    @classmethod
    def zeros(cls, shape, dtype=np.float32) -> "DeviceArray":
        out = cls(shape, dtype)
        check(lib().xde_memset_async(C.c_void_p(out._ptr), 0, out._n * out.dtype.itemsize, None))
        return out

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        check(lib().xde_memcpy_async(out.ctypes.data_as(C.c_void_p), C.c_void_p(self._ptr), out.nbytes, 1, None))
        check(lib().xde_stream_synchronize(None))
        return out

    # -- the small tensor surface the solver classes use ---------------------------------------------------------
    def data_ptr(self) -> int:
        return self._ptr

    def numel(self) -> int:
        return self._n

    def dim(self) -> int:
        return len(self.shape)

    def reshape(self, *shape) -> "DeviceArray":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            known = 1
            for s in shape:
                if s != -1:
                    known *= s
            shape[shape.index(-1)] = self._n // max(known, 1)
        n = 1
        for s in shape:
            n *= s
        if n != self._n:
            raise ValueError(f"cannot reshape {self.shape} to {tuple(shape)}")
        return DeviceArray(shape, self.dtype, _buf=self._buf, _ptr=self._ptr)

    def __getitem__(self, i):  # leading-axis integer index / slice: still contiguous
        if isinstance(i, int):
            if i < 0:
                i += self.shape[0]
            step = self._n // self.shape[0]
            return DeviceArray(self.shape[1:], self.dtype, _buf=self._buf, _ptr=self._ptr + i * step * self.dtype.itemsize)
        if isinstance(i, slice):
            lo, hi, st = i.indices(self.shape[0])
            if st != 1:
                raise IndexError("DeviceArray slices must be contiguous")
            step = self._n // self.shape[0]
            return DeviceArray((max(hi - lo, 0),) + self.shape[1:], self.dtype, _buf=self._buf,
                               _ptr=self._ptr + lo * step * self.dtype.itemsize)
        raise IndexError("DeviceArray supports leading-axis integer / slice indexing only")

    def contiguous(self):
        return self

    def cpu(self):  # `x.cpu().numpy()` reads a result whatever the provider
        return self

    def detach(self):
        return self

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device=cuda:{self.device[1]})"

    # -- zero-copy interchange -------------------------------------------------------------------------------------
    def __dlpack_device__(self):
        return (2, self.device[1])  # kDLCUDA

    def __dlpack__(self, stream=None):
        """DLPack capsule of this buffer.  The capsule BORROWS the allocation and keeps this object alive through the
        capsule's context (a Python reference held until the consumer's deleter has run)."""
        code, bits = _DT[self.dtype]
        shp = (C.c_int64 * max(len(self.shape), 1))(*self.shape)
        m = lib().xde_dlpack_wrap(C.c_void_p(self._ptr), len(self.shape), shp, code, bits, self.device[1], 0)
        if not m:
            raise RuntimeError("xde_dlpack_wrap failed")
        _keepalive[m] = self  # released by the consumer-side deleter hook below or at interpreter exit
        C.pythonapi.PyCapsule_New.restype = C.py_object
        C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        return C.pythonapi.PyCapsule_New(m, b"dltensor", None)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self._ptr, False), "version": 3, "strides": None}


_keepalive: dict = {}


def release_exports() -> int:
    """Drop the references that keep DLPack-exported DeviceArrays alive (the C deleter of a borrowed capsule cannot call
    back into Python, so an exported array otherwise lives until the process ends).  Call it only when every consumer
    of those capsules is gone.  -> number of references dropped."""
    n = len(_keepalive)
    _keepalive.clear()
    return n


def asarray(x, dtype=np.float32) -> DeviceArray:
    if isinstance(x, DeviceArray):
        if x.dtype != np.dtype(dtype):
            return DeviceArray.from_numpy(x.numpy().astype(dtype))
        return x
    return DeviceArray.from_numpy(np.asarray(x, dtype=dtype))
