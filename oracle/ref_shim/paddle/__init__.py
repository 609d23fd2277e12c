"""A NumPy stand-in for the handful of `paddle` operations the reference's hot-path modules use.

TEST INFRASTRUCTURE ONLY (lives under oracle/): PaddlePaddle cannot be installed in the build container, so the
reference (pure Python on Paddle eager ops) could not be run -- and the oracle's restatement of its control flow could
only be pinned to the closed-form fixtures of the reference's tests.  With this module on `sys.path` the reference's OWN,
UNMODIFIED source files (`paddlexde/solver/base_adaptive_solver*.py`, `solver/adaptive_solver/*.py`,
`solver/base_fixed_solver.py`, `solver/fixed_solver/{euler,midpoint,rk4}.py`, `utils/ode_utils.py`,
`xde/base_{xde,ode}.py`, `interpolation/functional/interp_fn.py`; and, with the additions below,
`functional/odeint_adjoint.py`, `interpolation/interpolate_base.py`, `interpolation/interpolate.py`, `xde/base_dde.py`,
`functional/ddeint.py`) import and run here (`tools/make_reference_golden.py`, `make_reference_adjoint_golden.py`,
`make_reference_dde_golden.py`), which turns "the oracle restates the reference" into "the oracle reproduces what the
reference's code computes", bit for bit, for step sequences, batched (B > 1) behaviour, every tableau, `step_t` /
`jump_t`, the fixed solvers and `step_size` grids.

What is an operation of THIS module and not of the reference: how an eager op rounds.  Paddle's CPU kernels do not
document their summation orders, so every op here follows the repository's arithmetic specification (DESIGN.md
section 2) -- fp32 elementwise ops, `sum(axis)` left to right in fp32, `mean()` accumulated sequentially in fp64 with
`sqrt()` of that mean taken in fp64 and rounded once to fp32, `x ** (1/p)` through the specification's `rootp`,
`x ** 3 = (x * x) * x`, `a @ b` = rounded products summed left to right in fp32 (no fused multiply-add), `sum` over
several axes accumulated sequentially in fp64 and rounded once, `autograd.grad` = the vector-Jacobian product the
caller's field attached to its output (there is no tape).  What the
golden vectors therefore pin is everything ABOVE the op level: the reference's formulas, their order, its control flow.
"""
from __future__ import annotations

import builtins
import ctypes as C
import os

import numpy as np

float32, float64, int32, int64 = np.float32, np.float64, np.int32, np.int64
# "spec" (default): every op rounds as the arithmetic specification says -- the mode all golden vectors are made in.
# "libm": the ops whose rounding Paddle does not document take a DIFFERENT plausible implementation instead (x ** (1/p)
# through float32 pow, mean() / sum() through NumPy's pairwise fp32 reductions): used only by
# tools/reference_rounding_sensitivity.py to measure how far the reference's results move with op-level rounding.
ROUNDING = os.environ.get("XDE_SHIM_ROUNDING", "spec")
bool = np.bool_  # noqa: A001  (paddle.bool)

_ORC = None


def _oracle():
    global _ORC
    if _ORC is None:
        here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        _ORC = C.CDLL(os.path.join(here, "libxde_oracle.so"))
        _ORC.orc_rootpf.restype = C.c_float
        _ORC.orc_rootpf.argtypes = [C.c_float, C.c_int32]
    return _ORC


def _np(x):
    return x.a if isinstance(x, Tensor) else x


class Tensor:
    """An eager tensor: a NumPy array plus the methods the reference calls."""
    __array_priority__ = 1000

    def __init__(self, a, wide=False):
        self.a = np.asarray(a)
        self.stop_gradient = True
        self._wide = wide  # an fp64 mean on its way into sqrt() (the RMS norm of the specification)

    # ---- metadata ----
    @property
    def dtype(self):
        return self.a.dtype.type

    @property
    def shape(self):
        return list(self.a.shape)

    def __len__(self):
        return self.a.shape[0]

    def __iter__(self):
        for i in range(self.a.shape[0]):
            yield Tensor(self.a[i])

    def __bool__(self):
        return builtins_bool(self.a)

    def __float__(self):
        return float(self.a)

    def __repr__(self):
        return f"Tensor({self.a!r})"

    def item(self):
        return self.a.item()

    def tolist(self):
        return self.a.tolist()

    def numpy(self):
        return self.a

    def astype(self, dtype=None):
        return Tensor(self.a.astype(dtype))

    cast = astype

    def reshape(self, shape):
        return Tensor(self.a.reshape(shape))

    def flip(self, axis):
        return Tensor(np.flip(self.a, axis))

    def unsqueeze(self, axis):
        return Tensor(np.expand_dims(self.a, axis))

    def squeeze(self, axis=None):
        return Tensor(np.squeeze(self.a, axis))

    def __matmul__(self, o):
        """[..., n, k] @ [..., k, m] (broadcast over the leading dims): every product rounded to fp32, summed left to
        right over k in fp32 -- no fused multiply-add (the interpolation weights ts @ H @ ps, interpolate_base.py:92)."""
        x, y = self.a, _np(o)
        acc = x[..., :, 0:1] * y[..., 0:1, :]
        for j in range(1, x.shape[-1]):
            acc = acc + x[..., :, j:j + 1] * y[..., j:j + 1, :]
        return Tensor(acc)

    # ---- indexing ----
    def __getitem__(self, idx):
        idx = _np(idx) if not isinstance(idx, tuple) else tuple(_np(i) for i in idx)
        return Tensor(self.a[idx])

    def __setitem__(self, idx, value):
        idx = _np(idx) if not isinstance(idx, tuple) else tuple(_np(i) for i in idx)
        self.a[idx] = _np(value)

    # ---- arithmetic: a Python scalar takes the tensor's dtype (as in Paddle), results keep fp32 ----
    def _other(self, o):
        if isinstance(o, Tensor):
            return o.a
        if isinstance(o, (int, float)) and self.a.dtype.kind == "f":
            return self.a.dtype.type(o)
        return o

    def __add__(self, o): return Tensor(self.a + self._other(o))
    def __radd__(self, o): return Tensor(self._other(o) + self.a)
    def __sub__(self, o): return Tensor(self.a - self._other(o))
    def __rsub__(self, o): return Tensor(self._other(o) - self.a)
    def __mul__(self, o): return Tensor(self.a * self._other(o))
    def __rmul__(self, o): return Tensor(self._other(o) * self.a)
    def __truediv__(self, o): return Tensor(self.a / self._other(o))
    def __rtruediv__(self, o): return Tensor(self._other(o) / self.a)
    def __neg__(self): return Tensor(-self.a)

    def __pow__(self, e):
        """x ** e.  e = 1/p (the step-size controller and select_initial_step): the specification's rootp for finite
        positive x, x itself otherwise (inf ** (1/p) = inf, 0 ** (1/p) = 0, NaN stays NaN)."""
        ev = float(_np(e))
        if ev == 2.0:
            return Tensor(self.a * self.a)
        if ev == 3.0:  # the Hermite / Bezier monomials (interpolation/interpolate.py:186,280): (x * x) * x
            return Tensor((self.a * self.a) * self.a)
        if ROUNDING == "libm" and self.a.dtype == np.float32:
            return Tensor(np.power(self.a, np.float32(ev)))
        p = round(1.0 / ev) if ev != 0 else 0
        if p not in (2, 3, 5, 8) or builtins.abs(ev * p - 1.0) > 1e-6 or self.a.dtype != np.float32:
            raise NotImplementedError(f"paddle shim: x ** {ev}")
        flat = self.a.reshape(-1)
        out = np.array([_oracle().orc_rootpf(float(v), p) if (v > 0 and np.isfinite(v)) else v for v in flat], np.float32)
        return Tensor(out.reshape(self.a.shape))

    # ---- comparisons ----
    def __lt__(self, o): return Tensor(self.a < self._other(o))
    def __le__(self, o): return Tensor(self.a <= self._other(o))
    def __gt__(self, o): return Tensor(self.a > self._other(o))
    def __ge__(self, o): return Tensor(self.a >= self._other(o))
    def __eq__(self, o): return Tensor(self.a == self._other(o))  # noqa: E704
    def __ne__(self, o): return Tensor(self.a != self._other(o))
    def __and__(self, o): return Tensor(self.a & _np(o))
    __hash__ = None

    # ---- methods ----
    def abs(self):
        return Tensor(np.abs(self.a), wide=self._wide)

    def pow(self, e):
        return self.__pow__(e)

    def mean(self):
        """fp64, accumulated sequentially over the flattened tensor (oracle: rms_f64)."""
        if ROUNDING == "libm":
            return Tensor(np.mean(self.a, dtype=self.a.dtype))
        flat = self.a.reshape(-1).astype(np.float64)
        acc = np.float64(0.0)
        for v in flat:
            acc = acc + v
        return Tensor(acc / np.float64(flat.size), wide=True)

    def sqrt(self):
        if self._wide:
            return Tensor(np.float32(np.sqrt(self.a)))  # (float) sqrt(acc / n)
        return Tensor(np.sqrt(self.a))

    def max(self):
        return Tensor(self.a.max())

    def all(self):
        return Tensor(self.a.all())

    def any(self):
        return Tensor(self.a.any())

    def clip(self, lo=None, hi=None):
        return Tensor(np.clip(self.a, _np(lo), _np(hi)))

    def reciprocal(self):
        return Tensor(self.a.dtype.type(1.0) / self.a)

    def detach(self):
        return Tensor(self.a)

    # ---- raw-buffer surface (integration/b200.py binds a C ABI with these) ----
    def contiguous(self):
        return self if self.a.flags["C_CONTIGUOUS"] else Tensor(np.ascontiguousarray(self.a))

    def data_ptr(self):
        assert self.a.flags["C_CONTIGUOUS"]
        return self.a.ctypes.data

    def dot(self, o):
        """1-D dot product: products rounded, summed left to right in fp32 (the oracle's grad_t_span terms)."""
        x, y = self.a.reshape(-1), _np(o).reshape(-1)
        acc = x[0] * y[0]
        for i in range(1, x.size):
            acc = acc + x[i] * y[i]
        return Tensor(acc)

    @property
    def trainable(self):
        return True


builtins_bool = builtins.bool


def zeros_like(x):
    return Tensor(np.zeros_like(_np(x)))


def split(x, sizes, axis=0):
    return [Tensor(p) for p in np.split(_np(x), np.cumsum(sizes)[:-1], axis=axis)]


class _CurrentStream:
    cuda_stream = 0


class device:  # noqa: N801  (paddle.device.cuda.current_stream().cuda_stream)
    class cuda:  # noqa: N801
        @staticmethod
        def current_stream():
            return _CurrentStream()


def ones_like(x):
    return Tensor(np.ones_like(_np(x)))


def get_default_dtype():
    return float32


def cast(x, dtype=None):
    return Tensor(np.asarray(_np(x)).astype(dtype))


def stack(xs, axis=0):
    return Tensor(np.stack([_np(v) for v in xs], axis=axis))


def bucketize(x, sorted_sequence, right=False):
    """#{s < x} (right=False): numpy's searchsorted(side='left')."""
    return Tensor(np.searchsorted(_np(sorted_sequence), _np(x), side="right" if right else "left").astype(np.int64))


def index_select(x, index, axis=0):
    return Tensor(np.take(_np(x), _np(index), axis=axis))


class _SparseCoo:
    def __init__(self, indices, values, shape):
        self._dense = np.zeros(shape, np.float32)
        for r, c, v in zip(indices[0], indices[1], values):
            self._dense[r, c] = np.float32(v)

    def to_dense(self):
        return Tensor(self._dense)


class sparse:  # noqa: N801  (paddle.sparse)
    @staticmethod
    def sparse_coo_tensor(indices, values, shape):
        return _SparseCoo(indices, values, shape)


def assign(x):
    return Tensor(np.array(_np(x)))


class set_grad_enabled:
    def __init__(self, mode):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def to_tensor(x, dtype=None, **_):
    return Tensor(np.array(_np(x), dtype=dtype))


def empty(shape, dtype=float32):
    return Tensor(np.zeros([int(s) for s in shape], dtype=dtype))


zeros = empty


def abs(x):  # noqa: A001
    return Tensor(np.abs(_np(x)))


def sum(x, axis=None):  # noqa: A001
    """Left-to-right fp32 accumulation along `axis` (arithmetic specification: the stage sums)."""
    a = _np(x)
    if ROUNDING == "libm":
        return Tensor(np.sum(a, axis=tuple(axis) if isinstance(axis, (list, tuple)) else axis, dtype=a.dtype))
    if isinstance(axis, (list, tuple)):
        # several axes at once (HistoryIndex.backward, xde/base_dde.py:126): the kept axes first, the reduced ones
        # flattened in row-major order and accumulated sequentially in fp64, one rounding to fp32 (oracle:
        # orc_history_gather_bwd)
        red = [ax % a.ndim for ax in axis]
        keep = [ax for ax in range(a.ndim) if ax not in red]
        flat = np.transpose(a, keep + red).reshape([a.shape[ax] for ax in keep] + [-1]).astype(np.float64)
        acc = np.zeros(flat.shape[:-1], np.float64)
        for j in range(flat.shape[-1]):
            acc = acc + flat[..., j]
        return Tensor(acc.astype(np.float32))
    a = np.moveaxis(a, axis, -1)
    s = a[..., 0].copy()
    for j in range(1, a.shape[-1]):
        s = s + a[..., j]
    return Tensor(s)


def concat(xs, axis=0):
    return Tensor(np.concatenate([_np(v) for v in xs], axis=axis))


def fmax(a, b):
    x, y = np.asarray(_np(a)), np.asarray(_np(b))
    dt = x.dtype if x.dtype.kind == "f" else y.dtype
    return Tensor(np.fmax(x.astype(dt), y.astype(dt)))


def fmin(a, b):
    x, y = np.asarray(_np(a)), np.asarray(_np(b))
    dt = x.dtype if x.dtype.kind == "f" else y.dtype
    return Tensor(np.fmin(x.astype(dt), y.astype(dt)))


def max(a, b=None):  # noqa: A001
    if b is None:
        return Tensor(np.max(_np(a)))
    return Tensor(np.maximum(_np(a), _np(b)))  # the degenerate branch of select_initial_step


def sort(x):
    return Tensor(np.sort(_np(x)))


def isfinite(x):
    return Tensor(np.isfinite(_np(x)))


def numel(x):
    return Tensor(np.int64(_np(x).size))


def equal_all(a, b):
    return Tensor(np.array_equal(_np(a), _np(b)))


def ceil(x):
    return Tensor(np.ceil(_np(x)))


def arange(start, end=None, step=1, dtype=None):
    return Tensor(np.arange(_np(start), _np(end), _np(step), dtype=dtype))


def linspace(a, b, n, dtype=float32):
    return Tensor(np.linspace(a, b, int(n), dtype=dtype))


def promote_types(a, b):
    return np.promote_types(a, b).type


class no_grad:
    def __call__(self, fn):
        return fn

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Ctx:
    def save_for_backward(self, *tensors):
        self._saved = tensors

    def saved_tensor(self):
        return self._saved


class _PyLayer:
    """`apply` runs `forward` and leaves the context on the result (`._ctx`), so that a test can call the layer's own
    `backward(ctx, grad)` -- there is no tape in the stand-in."""

    @classmethod
    def apply(cls, *args, **kwargs):
        ctx = _Ctx()
        out = cls.forward(ctx, *args, **kwargs)
        out._ctx = ctx
        return out


class autograd:
    PyLayer = _PyLayer

    @staticmethod
    def grad(outputs, inputs, grad_outputs=None, allow_unused=False, retain_graph=False):
        """The one use the reference makes of autograd (functional/odeint_adjoint.py:108-114): the vector-Jacobian
        product of the caller's field.  `outputs` must come from a field that attached its VJP (`_vjp`: cotangent ->
        gradients in the order of `inputs`, None where the output does not depend on the input)."""
        vjp = getattr(outputs, "_vjp", None)
        if vjp is None:
            raise NotImplementedError("paddle shim: autograd.grad of a tensor without an attached VJP")
        grads = vjp(grad_outputs)
        assert len(grads) == len(inputs)
        return list(grads)


from . import nn  # noqa: E402,F401  (`import paddle.nn as nn`, `from paddle import nn`)
