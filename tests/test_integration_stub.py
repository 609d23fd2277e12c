"""The reference-side binding of INTEGRATION.md (integration/b200.py: `Dopri5B200`, `OdeintAdjointB200` -- ctypes + Paddle
only) EXECUTED: the reference's own `functional/odeint.py` instantiates the stub through its solver-class protocol on
the NumPy `paddle` stand-in, the C ABI is answered by the CPU oracle on the stub's raw pointers (tests/host_dry_run.py),
and the results are compared with the reference-run vectors: the reference with `solver=Dopri5B200` returns what the
reference with its own `Dopri5` returned.  Checks the documented binding (argument order, struct layouts, protocol), not
the kernels.  Needs /root/reference (skipped on the GPU box)."""
import ast
import ctypes
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZF = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_vectors.npz"), allow_pickle=False)
ZA = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_adjoint_vectors.npz"), allow_pickle=False)


class _Entry:
    """A callable that accepts `.restype` / `.argtypes` like a ctypes function pointer."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, *a):
        return self.fn(*a)


@pytest.fixture(scope="module")
def ref_and_stub(oracle):
    from oracle.ref_shim import loader

    if not loader.available():
        pytest.skip("/root/reference is not on this machine")
    from tests.host_dry_run import FakeLib

    ns = loader.load()
    ns.BaseODE.format = lambda self, sol: sol  # repair R1: `xde.format` does not exist at HEAD (functional/odeint.py:33)
    fake = FakeLib()
    handle = types.SimpleNamespace(**{n: _Entry(getattr(fake, n)) for n in
                                      ("xde_dopri5_mlp_f32", "xde_dopri5_mlp_adjoint_f32", "xde_last_error")})
    real_cdll = ctypes.CDLL
    ctypes.CDLL = lambda *a, **k: handle
    try:
        spec = importlib.util.spec_from_file_location("paddlexde_solver_b200", os.path.join(ROOT, "integration", "b200.py"))
        stub = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(stub)  # `import paddle` inside resolves to the stand-in installed by loader.load()
    finally:
        ctypes.CDLL = real_cdll
    return ns, stub


def make_func(P, z, name, pre):
    """An object shaped like example/ode_demo.py's ODEFunc: `.net = Sequential(Linear, Tanh, Linear)` with Paddle's
    [in, out] weights.  The stub never calls it -- it reads the parameters."""
    lin = lambda w, b: types.SimpleNamespace(weight=P.to_tensor(z[f"{name}/{w}"]), bias=P.to_tensor(z[f"{name}/{b}"]))  # noqa: E731
    Tanh = type("Tanh", (), {})
    layers = [lin("w1", "b1"), Tanh(), lin("w2", "b2")]
    return types.SimpleNamespace(net=types.SimpleNamespace(children=lambda: iter(layers)), pre=pre)


@pytest.mark.parametrize("name,controller", [("cfg2_dopri5_B1", 0), ("cfg2_dopri5_B1", 1), ("cfg1_dopri5_B20", 1),
                                             ("dopri5_options", 1), ("dopri5_min_step", 1)])
def test_reference_odeint_with_the_stub_solver_returns_the_reference_run(ref_and_stub, name, controller):
    ns, stub = ref_and_stub
    P = ns.paddle
    meta = ast.literal_eval(str(ZF[f"{name}/meta"]))
    opts = {k: v for k, v in meta.items() if k not in ("kind", "solver", "pre", "rtol", "atol")}
    func = make_func(P, ZF, name, meta["pre"])
    sol = ns.odeint_mod.odeint(func, P.to_tensor(ZF[f"{name}/y0"]), P.to_tensor(ZF[f"{name}/t"]), stub.Dopri5B200,
                               rtol=meta.get("rtol", 1e-7), atol=meta.get("atol", 1e-9),
                               options={"norm": ns.ode_utils._rms_norm, "controller": controller, **opts})
    assert np.array_equal(sol.a, ZF[f"{name}/sol"])


@pytest.mark.parametrize("name", ["b1_cfg2_seminorm", "b1_cfg2_grad_t", "b1_d4_options", "b1_d8", "b1_d2_reverse_span"])
def test_stub_pylayer_backward_returns_the_reference_run_gradients(ref_and_stub, name):
    """B = 1: the stub's per-trajectory controller (XDE_CTRL_TRAJECTORY, seminorm) is the reference's."""
    ns, stub = ref_and_stub
    P = ns.paddle
    meta = ast.literal_eval(str(ZA[f"{name}/meta"]))
    assert meta["adj_norm"] == "seminorm"
    opts = {k: v for k, v in meta.items() if k not in ("pre", "adj_norm", "rtol", "atol")}
    func = make_func(P, ZA, name, meta["pre"])
    t = P.to_tensor(ZA[f"{name}/t"])
    t.stop_gradient = f"{name}/grad_t" not in ZA.files
    holder = dict(func=func, odeint=ns.odeint_mod.odeint, rtol=meta.get("rtol", 1e-7), atol=meta.get("atol", 1e-9),
                  options=opts)
    params = [func.net.children().__next__().weight]  # (the PyLayer's *params only name the autograd inputs)
    sol = stub.OdeintAdjointB200.apply(holder, P.to_tensor(ZA[f"{name}/y0"]), t, *params)
    assert np.array_equal(sol.a, ZA[f"{name}/sol"])
    res = stub.OdeintAdjointB200.backward(sol._ctx, P.to_tensor(ZA[f"{name}/grad_y"]))
    assert res[0] is None
    if t.stop_gradient:
        assert res[1] is None
    else:
        assert np.array_equal(res[1].a, ZA[f"{name}/grad_t"])
    for got, key in zip(res[2:], ("gw1", "gb1", "gw2", "gb2")):
        assert np.array_equal(got.a, ZA[f"{name}/{key}"]), key


def test_integration_md_quotes_the_stub_verbatim():
    src = open(os.path.join(ROOT, "integration", "b200.py")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = src[src.index("import ctypes as C"):]
    adj = body.index("class OdeintAdjointB200")
    assert body[:adj].rstrip() in doc and body[adj:].rstrip() in doc, "INTEGRATION.md section 2 / 2b != integration/b200.py"
