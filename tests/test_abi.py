"""CPU-side checks of the drop-in boundary: libxde_b200.so loads and exports every symbol that
include/xde_b200.h declares, struct layouts agree with the header, and the host logic of the shim
(option defaulting, argument errors, the no-fallback rule) behaves like the reference's.
No compute call is made here (no GPU in this suite)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xde_b200.h")


@pytest.fixture(scope="module")
def so():
    from paddlexde_b200 import _lib

    if not os.path.exists(_lib.library_path()):
        import __graft_entry__ as g

        g.build()
    return _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xde_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(so):
    names = declared_functions()
    assert len(names) >= 11
    handle = C.CDLL(so.library_path())
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/xde_b200.h but not exported"
    assert sorted(so.exported_symbols()) == names, "ctypes binding table and header disagree"
    assert so.lib().xde_abi_version() == 3


def test_struct_layouts(so):
    # sizes implied by the header (LP64): field 4*4 + 4*8; opts 8*4 + 2*4; stats 3*8 + 2*4; log 2*8 + 2*4
    assert C.sizeof(so.MlpFieldC) == 48
    assert C.sizeof(so.CtrlOptsC) == 40
    assert C.sizeof(so.StatsC) == 32
    assert C.sizeof(so.AttemptLogC) == 24
    o = so.CtrlOptsC()
    so.lib().xde_default_ctrl_opts(C.byref(o))
    # solver/base_adaptive_solver_rk.py:32-49 defaults
    assert (o.rtol, o.atol) == (np.float32(1e-7), np.float32(1e-9))
    assert o.min_step == 0.0 and o.max_step == float("inf") and o.first_step != o.first_step
    assert (o.safety, o.ifactor, o.dfactor) == (np.float32(0.9), 10.0, np.float32(0.2))
    assert o.max_num_steps == 2 ** 31 - 1


def test_null_arguments_are_bad_arg_without_a_device(so):
    lib = so.lib()
    rc = lib.xde_dopri5_mlp_f32(None, None, 1, None, 2, None, 0, None, None, None, None)
    assert rc == so.XDE_E_BAD_ARG
    with pytest.raises(ValueError):
        so.check(rc)
    rc = lib.xde_history_gather_f32(7, None, 1, 2, 1, None, None, 1, None, None, None)
    assert rc == so.XDE_E_BAD_ARG
    assert b"null" in lib.xde_last_error()
    assert lib.xde_dde_fuse_f32(None, 0.1, None, 4, None, None) == so.XDE_E_BAD_ARG


def test_status_words_raise_the_reference_asserts(so):
    so.raise_for_status(0)
    for st, text in ((1, "underflow"), (2, "non-finite"), (3, "max_num_steps"), (5, "interpolation")):
        with pytest.raises(AssertionError, match=text):
            so.raise_for_status(st)


def test_no_fallback_for_python_callables():
    import paddlexde_b200 as px

    with pytest.raises(px.UnsupportedFieldError):
        px.field.as_field(lambda t, y: -y)
    with pytest.raises(px.UnsupportedFieldError):
        px.solver.adaptive_solver.check_norm(lambda x: x.abs().max())
    px.solver.adaptive_solver.check_norm(px.utils._rms_norm)


def test_tspan_validation():
    from paddlexde_b200.solver.adaptive_solver import host_tspan

    assert host_tspan([0.0, 1.0, 2.0]).dtype == np.float32
    assert host_tspan([2.0, 1.0]).tolist() == [2.0, 1.0]
    for bad in ([0.0], [0.0, 0.0, 1.0], [0.0, 2.0, 1.0]):
        with pytest.raises(ValueError):
            host_tspan(bad)


def test_fixed_solver_ctor_contract():
    import paddlexde_b200 as px

    class X:
        kind = "ode"

    with pytest.raises(KeyError):  # base_fixed_solver.py:45-47 requires rtol/atol
        px.RK4(xde=X(), y0=None)
    with pytest.raises(ValueError):  # base_fixed_solver.py:58-60
        px.RK4(xde=X(), y0=None, rtol=1, atol=1, step_size=0.1, grid_constructor=lambda *a: None)
    with pytest.raises(NotImplementedError):
        px.ddeint_adjoint()


def test_sort_tvals_single_point_on_a_decreasing_span():
    """found by tools/fuzz_parity.py: a reversed length-1 NumPy view keeps a negative stride"""
    from paddlexde_b200.solver.adaptive_solver import AdaptiveRKSolver

    import torch

    t = np.array([1.5, 1.0, 0.2], np.float32)
    like = torch.empty(0)  # a host tensor: the sorted values land on its device (no GPU in this suite)
    assert AdaptiveRKSolver._sort_tvals([0.5], t, like).tolist() == [0.5]
    assert AdaptiveRKSolver._sort_tvals([0.5, 0.75, 9.0], t, like).tolist() == [0.75, 0.5]
    assert AdaptiveRKSolver._sort_tvals([0.5, 0.25, -1.0], t[::-1].copy(), like).tolist() == [0.25, 0.5]
    assert AdaptiveRKSolver._sort_tvals([9.0], t, like) is None


def test_default_controller_selection(monkeypatch):
    """The controller granularity when the caller does not choose: per trajectory (north star), announced once;
    PADDLEXDE_B200_CONTROLLER flips it globally without a warning."""
    import warnings

    import paddlexde_b200.solver.adaptive_solver as A

    monkeypatch.delenv("PADDLEXDE_B200_CONTROLLER", raising=False)
    A._warned_default = False
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert A.default_controller(64) == "trajectory"
        assert A.default_controller(64) == "trajectory"
    assert sum(issubclass(r.category, A.ControllerDefaultWarning) for r in rec) == 1
    A._warned_default = False
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert A.default_controller(1) == "trajectory"  # B = 1: both controllers coincide, nothing to announce
    assert not rec
    monkeypatch.setenv("PADDLEXDE_B200_CONTROLLER", "batch")
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert A.default_controller(64) == "batch"
    assert not rec
    monkeypatch.setenv("PADDLEXDE_B200_CONTROLLER", "nonsense")
    with pytest.raises(ValueError):
        A.default_controller(64)


def test_fixed_solver_time_grid_follows_the_reference():
    """_grid_constructor_from_step_size (base_fixed_solver.py:66-89): arange(niters) * step + start with the last point
    replaced by t[-1]; the solver uses its first len(t_span) points; the reference's end-point asserts."""
    import paddlexde_b200 as px

    class X:
        kind = "ode"

    t = np.linspace(0, 1, 5).astype(np.float32)
    s = px.RK4(xde=X(), y0=None, rtol=1, atol=1, step_size=0.1)
    g = s._time_grid(t)
    assert g.dtype == np.float32 and g.size == t.size
    assert np.array_equal(g, (np.arange(5, dtype=np.float32) * np.float32(0.1)).astype(np.float32))
    s = px.RK4(xde=X(), y0=None, rtol=1, atol=1, grid_constructor=lambda y, tt: np.array([0, 0.3, 0.5, 0.6, 0.9, 1.0]))
    assert np.array_equal(s._time_grid(t), np.array([0, 0.3, 0.5, 0.6, 0.9], np.float32))
    with pytest.raises(AssertionError):
        px.RK4(xde=X(), y0=None, rtol=1, atol=1, grid_constructor=lambda y, tt: np.array([0.1, 0.5, 1.0]))._time_grid(t)
    with pytest.raises(ValueError):  # fewer grid points than steps to take
        px.RK4(xde=X(), y0=None, rtol=1, atol=1, step_size=0.5)._time_grid(t)
    with pytest.raises(NotImplementedError):
        px.RK4(xde=X(), y0=None, rtol=1, atol=1, step_size=0.1, interp="cubic")


def test_shard_rows_and_hooks_without_a_process_group():
    from paddlexde_b200 import distributed as pxd

    assert [pxd.shard_rows(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert pxd.init_from_env() == (0, 1, 0)
    g = np.ones(3, np.float32)
    assert pxd.grad_allreduce()(g) is g  # a single process: nothing to sum
